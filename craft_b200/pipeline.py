"""Streaming frame pairs through a CRAFT model (the inference loop of evaluate.py, e.g. :1251-1384 / :120-160,
which does `image.cuda()` -> model(image1, image2) -> `.cpu()` one pair at a time on one stream).

`PairStream` keeps that per-pair contract -- every pair's frames travel host -> device and its full-resolution
flow travels device -> host, every pair is one batch-1 forward -- and overlaps what is independent:

* the copies run on their own CUDA streams (one per direction) with staging slots, so the H2D copy of the next
  pairs and the D2H copy of the previous ones run under the forwards (2.75 MB in and 3.67 MB out per 448x1024
  pair, ~0.15 ms that a single stream serialises with the forward);
* `lanes` > 1 keeps that many PAIRS in flight: lane k = one CUDA stream + the model's lane-k workspace and CUDA
  graph (CRAFT.on_lane).  One pair is a serial chain of ~290 kernels most of which cannot fill the GPU (the
  update-block GEMMs run 114 CTAs on 148 SMs, every kernel has a prologue and an epilogue phase); the kernels
  of a second and third pair run in those holes.  448x1024, 12 iterations, one B200: 242 pairs/s with one lane,
  274 with two, 284 with three (profiles/r02_lanes.txt).  The latency of a single pair grows accordingly; results
  do not depend on the number of lanes.

    stream = PairStream(model, iters=12, lanes=3)
    for flow_up in stream.map(pairs):          # pairs: iterable of (uint8|float host tensors [1,3,H,W]) x 2
        ...                                    # flow_up: pinned host tensor [1,2,H,W], valid until the next item is requested
    flows = stream.run_resident(device_pairs)  # inputs already in HBM, outputs stay there (no copies)
"""
import collections

import torch


class PairStream:
    def __init__(self, model, iters=12, device=None, lanes=1):
        self.model = model
        self.iters = iters
        self.lanes = max(1, int(lanes))
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("craft_b200 runs on a CUDA (sm_100a) device only; there is no CPU path")
        # one stream per direction: on a single copy stream the H2D of pair i+1 would queue behind the D2H of pair i,
        # which waits for forward(i) -- and nothing would overlap
        self.h2d_stream = torch.cuda.Stream(device=self.device)
        self.d2h_stream = torch.cuda.Stream(device=self.device)
        # lane 0 is the caller's current stream when there is only one lane (the round-2 behaviour)
        self.lane_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)] if self.lanes > 1 else [None]
        self._slots = None

    # ---------------------------------------------------------------------------------------------
    def _ensure(self, a):
        shape, dtype = tuple(a.shape), a.dtype
        if self._slots is not None and self._slots[0]["key"] == (shape, dtype):
            return
        B, _, H, W = shape
        self._slots = []
        for _ in range(2 * self.lanes):
            self._slots.append(dict(
                key=(shape, dtype),
                a=torch.empty(shape, dtype=dtype, device=self.device), b=torch.empty(shape, dtype=dtype, device=self.device),
                out=torch.empty((B, 2, H, W), dtype=torch.float32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_out=None, ev_done=torch.cuda.Event(), up=None))

    def _lane_stream(self, k):
        return self.lane_streams[k] if self.lane_streams[k] is not None else torch.cuda.current_stream(self.device)

    def _forward(self, k, a, b):
        with torch.no_grad(), self.model.on_lane(k):
            return self.model(a, b, iters=self.iters, test_mode=1)[1]

    def _submit(self, i, a, b):
        """Enqueue pair i: H2D on the copy stream, forward on its lane's stream, D2H on the copy stream."""
        self._ensure(a)
        s = self._slots[i % (2 * self.lanes)]
        k = i % self.lanes
        ls, hs, ds = self._lane_stream(k), self.h2d_stream, self.d2h_stream
        with torch.cuda.stream(hs):
            if s["ev_out"] is not None:
                hs.wait_event(s["ev_out"])         # the forward that last read this slot (pair i - 2*lanes) is done
            s["a"].copy_(a, non_blocking=True)
            s["b"].copy_(b, non_blocking=True)
            s["ev_in"].record(hs)
        with torch.cuda.stream(ls):
            ls.wait_event(s["ev_in"])
            up = self._forward(k, s["a"], s["b"])
            s["up"] = up                       # keeps the device tensor alive until its copy has run
            if s["ev_out"] is None:
                s["ev_out"] = torch.cuda.Event()
            s["ev_out"].record(ls)
        with torch.cuda.stream(ds):
            ds.wait_event(s["ev_out"])
            s["out"].copy_(up, non_blocking=True)     # (the previous use of this host buffer was handed out 2*lanes pairs ago)
            s["ev_done"].record(ds)
        return s

    def map(self, pairs):
        """Yield the full-resolution flow (pinned host tensor) of every pair, in order, `lanes` pairs behind the
        submissions."""
        with torch.cuda.device(self.device):
            self._fork()
            pending = collections.deque()
            for i, (a, b) in enumerate(pairs):
                pending.append(self._submit(i, a, b))
                if len(pending) > self.lanes:
                    s = pending.popleft()
                    s["ev_done"].synchronize()
                    yield s["out"]
            while pending:
                s = pending.popleft()
                s["ev_done"].synchronize()
                yield s["out"]
            self._join()

    def run_resident(self, pairs, keep=True):
        """Forward every (image1, image2) pair of DEVICE tensors, round-robin over the lanes; returns the list of
        full-resolution flows (device tensors; keep=False: only the last flow of every lane -- throughput runs).
        Work issued on the caller's stream before the call is waited for by every lane, and the caller's stream waits
        for every lane at the end, so CUDA events recorded around the call on the caller's stream bracket all of it."""
        outs = []
        last = [None] * self.lanes
        inflight = collections.deque()
        with torch.cuda.device(self.device):
            self._fork()
            for i, (a, b) in enumerate(pairs):
                k = i % self.lanes
                # the host stays at most two rounds ahead of the device (as map() does through its result hand-off):
                # with every launch of a long run queued at once the measured throughput dipped by 2-4 % now and then
                if len(inflight) >= 2 * self.lanes:
                    inflight.popleft().synchronize()
                with torch.cuda.stream(self._lane_stream(k)):
                    o = self._forward(k, a, b)     # (a dropped result's memory is reused in the order of its own lane's stream)
                    ev = torch.cuda.Event()
                    ev.record()
                inflight.append(ev)
                if keep:
                    outs.append(o)
                else:
                    last[k] = o
            if not keep:
                outs = [o for o in last if o is not None]
            self._join()
            if self.lanes > 1:      # allocated on a lane's stream, consumed on the caller's
                cur = torch.cuda.current_stream(self.device)
                for o in outs:
                    o.record_stream(cur)
        return outs

    def _fork(self):
        if self.lanes > 1:
            cur = torch.cuda.current_stream(self.device)
            for ls in self.lane_streams:
                ls.wait_stream(cur)

    def _join(self):
        if self.lanes > 1:
            cur = torch.cuda.current_stream(self.device)
            for ls in self.lane_streams:
                cur.wait_stream(ls)
