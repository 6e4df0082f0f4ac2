#!/usr/bin/env python
"""bench.py -- image-pairs/sec of CRAFT.forward(test_mode=1) at 448x1024, iters=12 (BASELINE.json
configs[1]) on N B200s, one process per GPU, pairs sharded across ranks (no data-path collective).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md section 8 for how each field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, ITERS = 448, 1024, 12
WORKLOAD = "craft-sintel config (craft+f2full+setrans), %dx%d pair, iters=%d, test_mode=1, batch 1" % (H, W, ITERS)
METRIC = "image-pairs/sec at 448x1024 iters=12"


def _peaks():
    """Roofline denominators: the driver-written MEASURED_PEAKS.json when present (key names matched
    loosely, a malformed file must not break the bench), else the profiling guide's fallback figures."""
    out = dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.isfile(p):
        return out
    try:
        d = json.load(open(p))
        flat = {}

        def walk(prefix, v):
            if isinstance(v, dict):
                for k, x in v.items():
                    walk(prefix + "." + str(k).lower(), x)
            elif isinstance(v, (int, float)):
                flat[prefix] = float(v)
        walk("", d)
        for k, v in flat.items():
            if "hbm" in k or ("copy" in k and "gb" in k):
                out["hbm"] = v
            elif "bf16" in k and "sustain" in k:
                out["tf_sust"] = v
            elif "bf16" in k and ("tflop" in k or "tf" in k):
                out["tf_burst"] = v
        out["src"] = "measured"
    except Exception:
        pass
    return out


def _pairs(n, device=None, uint8=False):
    import torch
    from oracle.ref_loader import synthetic_pair
    out = []
    for i in range(n):
        a, b = synthetic_pair(H, W, seed=1234 + i)
        if uint8:
            a, b = a.to(torch.uint8), b.to(torch.uint8)
        if device is not None:
            a, b = a.to(device), b.to(device)
        out.append((a, b))
    return out


def _state_dict():
    """Trained weights when the untracked local copy travelled with the snapshot, else seeded init."""
    import torch
    from craft_b200.network import CRAFT
    from oracle.ref_loader import craft_args
    torch.manual_seed(1234)
    model = CRAFT(craft_args())
    ck = os.path.join(ROOT, "tests", "golden", "_local", "craft-sintel-model.pth")
    src = "random-init (seed 1234)"
    if os.path.isfile(ck):
        model.load_state_dict(torch.load(ck, map_location="cpu"), strict=True)
        src = "craft-sintel.pth"
    return model, src


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def _cpu_port_rate(model_cpu, budget_s, threads):
    """Oracle port (oracle/cpu_forward.py) timed on the host cores; one step = one 448x1024 pair."""
    import torch
    from oracle import cpu_forward
    torch.set_num_threads(threads)
    sd = {k: v for k, v in model_cpu.state_dict().items()}
    pairs = _pairs(2)
    n, t_total = 0, 0.0
    with torch.no_grad():
        while n < 2 or (t_total < budget_s and n < 4):
            a, b = pairs[n % len(pairs)]
            t0 = time.time()
            cpu_forward.craft_forward(sd, a, b, iters=ITERS)
            t_total += time.time() - t0
            n += 1
            if t_total > budget_s:
                break
    return n / t_total, n, t_total


def run_reference(args):
    """`--impl reference`: the reference's CPU path.  The reference is a pure-Python tree that cannot be
    pip-installed or shipped to the GPU box, so this arm times the oracle port (same torch ops, same
    order) on all host threads.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    model, wsrc = _state_dict()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import cpu_forward
    sd = {k: v for k, v in model.state_dict().items()}
    pairs = _pairs(2)
    steps = max(1, min(args.steps, 6))
    warm = max(0, min(args.warmup, 1))
    with torch.no_grad():
        for i in range(warm):
            cpu_forward.craft_forward(sd, *pairs[i % 2], iters=ITERS)
        t0 = time.time()
        done = 0
        for i in range(steps):
            cpu_forward.craft_forward(sd, *pairs[i % 2], iters=ITERS)
            done += 1
            if time.time() - t0 > 150:
                break
        dt = time.time() - t0
    v = done / dt
    line = dict(impl="reference", metric=METRIC, value=v, unit="pairs/s", n_gpus=args.gpus, steps=done, warmup=warm,
                ms_per_step=1000 * dt / done, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic integer-noise pairs, weights: " + wsrc,
                config=dict(workload=WORKLOAD, device="host CPU"),
                cpu_baseline=dict(value=v, unit="pairs/s", cores=threads, kind="port",
                                  sample="%d full-size pairs (448x1024, iters=12)" % done),
                e2e=dict(value=v, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from craft_b200 import _lib, ops
    from craft_b200.ops import TokenGrid
    from craft_b200.setrans import get_workspace

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs CUDA devices; there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    lib = _lib.load()
    model_cpu, wsrc = _state_dict()
    import copy
    model = copy.deepcopy(model_cpu).to(dev).eval()
    pairs = _pairs(4, dev)

    def step(i):
        a, b = pairs[i % len(pairs)]
        return model(a, b, iters=ITERS, test_mode=1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu:
        with torch.no_grad():
            for i in range(max(args.warmup, 3)):
                step(i)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            step(0)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return
    sampler = ClockSampler(local) if rank == 0 else None
    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        barrier()
        if sampler:
            sampler.start()
        n0 = lib.craft_b200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        launches = lib.craft_b200_launch_count() - n0
        if launches == 0:   # CUDA-graph replay: the library is not called at replay time; count kernels per graph
            launches = args.steps * int(getattr(model, "launches_last_forward", 0))
        ms = e0.elapsed_time(e1)
        # ---- end-to-end: pinned host uint8 frames -> H2D -> forward -> D2H of the full-res flow
        host_pairs = [(a.pin_memory(), b.pin_memory()) for a, b in _pairs(4, None, uint8=True)]
        out_host = torch.empty((1, 2, H, W), dtype=torch.float32).pin_memory()

        def e2e_step(i):
            a, b = host_pairs[i % len(host_pairs)]
            _, up = model(a.to(dev, non_blocking=True).float(), b.to(dev, non_blocking=True).float(),
                          iters=ITERS, test_mode=1)
            out_host.copy_(up, non_blocking=True)

        for i in range(2):
            e2e_step(i)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            e2e_step(i)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- roofline of the dominant kernel: the motion aggregator's flash P.V (attn_pv_kernel<32,128,...>),
    #      12 launches per pair.  Timed alone with CUDA events on the launching stream.
    roof = None
    if rank == 0:
        g = TokenGrid(H // 8, W // 8)
        ws = model._workspaces.get(g, dev, model.materialize_level0)
        ks = ws.pv_split(4)
        O = ws.opart(ks, 4, 128)
        tbl = model.att.vispos_encoder.table()
        def pv():   # exactly the call hotpath.value_aggregate makes 12x per pair
            ops.attn_pv(ws.Qa, ws.Ka, ws.Vt, g, M=4, d=32, F=128, w_pos=1.0, pos_table=tbl, R=7,
                        clip=ws.clip_att, lse2=ws.lse2_att, out=O, ksplit=ks, zero_fill=False)
        for _ in range(3):
            pv()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        r0.record()
        for _ in range(reps):
            pv()
        r1.record()
        torch.cuda.synchronize()
        us = 1000 * r0.elapsed_time(r1) / reps
        U = g.U
        flops = 2.0 * U * U * 128 + 2.0 * 4 * U * U * 128          # QK^T (C=128) + P.V (M=4, F=128), algorithmic
        pk = _peaks()
        ach = flops / (us * 1e-6) / 1e12
        # DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/r01_pv_ncu_full.txt)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_pv_traffic.json")
        if os.path.isfile(tpath):
            t = json.load(open(tpath))
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        # second ceiling of this kernel: one MUFU ex2 per (query, key, mode) at 16 per clock per SM
        mufu_us = 4.0 * U * U / (16.0 * 148 * 1.965e9) * 1e6
        roof = dict(bound="tensor", kernel="attn_pv_kernel<32,128,128> (motion aggregator P.V, x12 per pair)",
                    achieved=ach, peak=pk["tf_burst"], unit="TFLOP/s", frac=ach / pk["tf_burst"], traffic=traffic,
                    us_per_launch=us, flops_per_launch=flops, peak_source=pk["src"] + " bf16 burst",
                    note="exp-bound before tensor-bound: 4*U^2 ex2 at the measured 15.2/clk/SM MUFU rate = %.1f us "
                         "per launch (profiles/r01_mb_exp.txt), i.e. the kernel runs at %.2f of its MUFU ceiling" % (
                             mufu_us * 16.0 / 15.2, mufu_us * 16.0 / 15.2 / us))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, n, tt = _cpu_port_rate(model_cpu, 20.0, threads)
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind="port",
                       sample="%d full-size pairs (448x1024, iters=12), %.1f s of CPU time" % (n, tt))
        total = world * args.steps
        line = dict(metric=METRIC, value=total / (ms * 1e-3), unit="pairs/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="bf16",
                    data="synthetic integer-noise pairs (distinct per step), weights: " + wsrc,
                    config=dict(workload=WORKLOAD, parallelism="pairs sharded over %d ranks, no collective" % world,
                                l2="inputs larger than L2: the per-step working set (68 MB pooled correlation pyramid, 30 MB "
                                   "P.V partial sums, >100 MB of encoder activations) exceeds the 126 MB L2 and every "
                                   "buffer is rewritten each step; 4 distinct input pairs rotate",
                                encoders="fnet/cnet (outside the hot path): cuDNN fp16 convolutions with fp32 accumulation "
                                         "+ craft_b200 norm/relu/residual kernels",
                                launch="whole forward replayed as one CUDA graph; craft_b200 kernels use programmatic "
                                       "dependent launch",
                                dead_work="test_mode=1 returns only the last upsampled flow: the mask head + convex "
                                          "upsampling of iterations 1..11 (discarded by the reference) are elided, "
                                          "outputs bit-identical (tests/test_gpu_e2e.py)"),
                    e2e=dict(value=total / (ms_e2e * 1e-3), unit="pairs/s", h2d_bytes_per_step=2 * 3 * H * W,
                             d2h_bytes_per_step=2 * H * W * 4),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling helper: warm up, then run ONE step inside a cudaProfilerStart/Stop range and exit "
                         "(use with `ncu --profile-from-start off ...`); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
