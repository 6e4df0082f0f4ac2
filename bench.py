#!/usr/bin/env python
"""bench.py -- image-pairs/sec of CRAFT.forward(test_mode=1) on N B200s, one process per GPU, pairs sharded
across ranks (no data-path collective).  Default workload = BASELINE.json configs[1]: craft-sintel.pth,
448x1024, iters=12.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config sintel|kitti|gma]

Prints ONE JSON line (rank 0).  DESIGN.md section 8 says how each field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "image-pairs/sec at 448x1024 iters=12"
CONFIGS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on
    "sintel": dict(H=448, W=1024, iters=12, args={}, weights="sintel", metric=METRIC,
                   workload="craft-sintel config (craft+f2full+setrans), 448x1024 pair, iters=12, test_mode=1, batch 1"),
    # BASELINE.json configs[4]: KITTI shape, 24 iterations (evaluate.py:180)
    "kitti": dict(H=384, W=1248, iters=24, args={}, weights="sintel", metric="image-pairs/sec at 384x1248 iters=24",
                  workload="craft-sintel weights, KITTI shape 384x1248 pair, iters=24, test_mode=1, batch 1"),
    # BASELINE.json configs[2]: f2full + GMA Aggregate (no checkpoint ships for this variant: seeded init, gamma 0.5)
    "gma": dict(H=448, W=1024, iters=12, args=dict(use_setrans=False), weights="seeded", metric=METRIC,
                workload="CRAFT f2full + gma.Attention/Aggregate (use_setrans=False), 448x1024 pair, iters=12, "
                         "test_mode=1, batch 1"),
    # BASELINE.json configs[3]: train_ddp.py FlyingThings-shape (400x720 crop, train-craft-f2full.sh:3), bs 2 per GPU,
    # DDP gradient all-reduce the only collective.  A step = forward + backward + AdamW step on one synthetic batch.
    "train": dict(H=400, W=720, iters=12, args={}, weights="seeded", batch=2,
                  metric="training image-pairs/sec at 400x720 iters=12, bs 2/GPU",
                  workload="train_ddp.py shape: CRAFT(craft+f2full+setrans) 400x720, iters=12, batch 2 per GPU, fwd+bwd+AdamW, "
                           "DDP(find_unused_parameters=True)"),
}


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


def _pairs(cfg, indices=None, device=None, uint8=False):
    """Synthetic pairs: a distinct seed per GLOBAL pair index.
    (`_pairs(n, device)` -- the form the profiling scripts use -- means the first n pairs of the default config.)"""
    import torch
    from craft_b200.testing import synthetic_pair
    if isinstance(cfg, int):
        cfg, indices, device = CONFIGS["sintel"], range(cfg), (indices if indices is not None else device)
    out = []
    for i in indices:
        a, b = synthetic_pair(cfg["H"], cfg["W"], seed=1234 + i)
        if uint8:
            a, b = a.to(torch.uint8), b.to(torch.uint8)
        if device is not None:
            a, b = a.to(device), b.to(device)
        out.append((a, b))
    return out


def _build_model(cfg):
    """craft_b200 CRAFT with the trained weights when the untracked local copy travelled with the snapshot."""
    import torch
    from craft_b200.network import CRAFT
    from craft_b200.testing import craft_args
    torch.manual_seed(1234)
    model = CRAFT(craft_args(**cfg["args"]))
    ck = os.path.join(ROOT, "tests", "golden", "_local", "craft-sintel-model.pth")
    src = "random-init (seed 1234)"
    if cfg["weights"] == "sintel" and os.path.isfile(ck):
        model.load_state_dict(torch.load(ck, map_location="cpu"), strict=True)
        src = "craft-sintel.pth"
    elif cfg["weights"] == "seeded" and hasattr(model.update_block.aggregator, "gamma"):
        with torch.no_grad():
            model.update_block.aggregator.gamma.fill_(0.5)      # gamma initialises to 0: make the aggregation count
        src += ", gamma=0.5"
    return model, src


# names the profiling scripts under profiles/ import
H, W, ITERS = CONFIGS["sintel"]["H"], CONFIGS["sintel"]["W"], CONFIGS["sintel"]["iters"]


def _state_dict():
    return _build_model(CONFIGS["sintel"])


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


# ------------------------------------------------------------------------------------------------
# the reference itself (CPU and, informationally, eager CUDA)
# ------------------------------------------------------------------------------------------------
def _reference_model(cfg, mixed_precision=False):
    """The UNMODIFIED reference (askerlee/craft core/network.py CRAFT) from /root/reference or, on the GPU box, from
    the untracked copy __graft_entry__.build() staged under oracle/_ref/reference.  None when neither exists."""
    try:
        from oracle import ref_loader as RL
        if not RL.reference_available():
            return None
        from craft_b200.testing import craft_args
        args = craft_args(mixed_precision=mixed_precision, **cfg["args"])
        ck = "craft-sintel.pth" if cfg["weights"] == "sintel" else None
        model, _ = RL.build_reference_model(args, checkpoint=ck)
        if ck is None:      # seeded variant: same weights as the craft_b200 model
            ours, _ = _build_model(cfg)
            model.load_state_dict(ours.state_dict(), strict=True)
        return model
    except Exception as e:      # a broken staging must not take the bench down: the port is the fallback
        sys.stderr.write("reference unavailable (%r), using the oracle port\n" % (e,))
        return None


def _cpu_step_fn(cfg):
    """-> (callable(pair) running ONE full-size pair on the host cores, kind)."""
    import torch
    ref = _reference_model(cfg)
    if ref is not None:
        def run(pair):
            with torch.no_grad():
                return ref(pair[0], pair[1], iters=cfg["iters"], test_mode=1)
        return run, "reference"
    from oracle import cpu_forward
    model, _ = _build_model(cfg)
    sd = {k: v for k, v in model.state_dict().items()}
    flags = dict(use_setrans=cfg["args"].get("use_setrans", True))

    def run(pair):
        with torch.no_grad():
            return cpu_forward.craft_forward(sd, pair[0], pair[1], iters=cfg["iters"], **flags)
    return run, "port"


def _cpu_rate(cfg, budget_s, threads, max_pairs=4):
    import torch
    torch.set_num_threads(threads)
    run, kind = _cpu_step_fn(cfg)
    pairs = _pairs(cfg, [0, 1])
    n, t_total = 0, 0.0
    while n < max_pairs:
        t0 = time.time()
        run(pairs[n % 2])
        t_total += time.time() - t0
        n += 1
        if n >= 2 and t_total > budget_s:
            break
    return n / t_total, n, t_total, kind


def run_reference(args, cfg):
    """`--impl reference`: the reference's own CPU implementation of the path (its unmodified Python tree,
    fp32, all host threads) on the same workload; the oracle port is the fallback when the tree is absent.
    Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    run, kind = _cpu_step_fn(cfg)
    pairs = _pairs(cfg, [0, 1])
    steps = max(1, min(args.steps, 6))
    warm = max(0, min(args.warmup, 1))
    for i in range(warm):
        run(pairs[i % 2])
    t0 = time.time()
    done = 0
    for i in range(steps):
        run(pairs[i % 2])
        done += 1
        if time.time() - t0 > 150:
            break
    dt = time.time() - t0
    v = done / dt
    _, wsrc = _build_model(cfg)
    line = dict(impl="reference", metric=cfg["metric"], value=v, unit="pairs/s", n_gpus=args.gpus, steps=done, warmup=warm,
                ms_per_step=1000 * dt / done, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic integer-noise pairs, weights: " + wsrc,
                config=dict(workload=cfg["workload"], device="host CPU"),
                cpu_baseline=dict(value=v, unit="pairs/s", cores=threads, kind=kind,
                                  sample="%d full-size pairs (%dx%d, iters=%d)" % (done, cfg["H"], cfg["W"], cfg["iters"])),
                e2e=dict(value=v, unit="pairs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def _gpu_reference(cfg, dev, pairs):
    """Informational: the unmodified reference in eager CUDA on the same B200, fp32 and
    its own fp16 autocast (evaluate.py:1455-1456) -- the honest comparator for the kernels.  CUDA-event timed."""
    import torch
    out = {}
    for tag, amp in (("fp32", False), ("fp16_autocast", True)):
        try:
            ref = _reference_model(cfg, mixed_precision=amp)
            if ref is None:
                return None
            ref = ref.to(dev).eval()
            with torch.no_grad():
                for i in range(2):
                    ref(*pairs[i % len(pairs)], iters=cfg["iters"], test_mode=1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 5
                e0.record()
                for i in range(n):
                    ref(*pairs[i % len(pairs)], iters=cfg["iters"], test_mode=1)
                e1.record()
                torch.cuda.synchronize()
            out[tag] = dict(value=n / (e0.elapsed_time(e1) * 1e-3), unit="pairs/s", steps=n)
            del ref
            torch.cuda.empty_cache()
        except Exception as e:
            out[tag] = dict(error=repr(e)[:200])
    return out


# ------------------------------------------------------------------------------------------------
# per-kernel roofline table
# ------------------------------------------------------------------------------------------------
def _time_us(fn, reps=20):
    """Average device time of one call of `fn`: `reps` calls captured into ONE CUDA graph and replayed, timed with
    CUDA events on the launching stream (an eager call costs ~20 us of host time in the torch dispatcher + ctypes,
    more than most of these kernels run; the model itself is replayed from a graph as well)."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / reps


def _kernel_table(model, cfg, dev, pk):
    """Each hot kernel of the setrans configuration timed ALONE (CUDA events on the launching stream) with the
    arguments the model uses, against the measured peaks.  ALGORITHMIC work only:
    QK^T = 2 U^2 C, P.V = 2 M U^2 F, conv = 2 U Cin Cout kh kw; recomputation is not credited."""
    import torch
    from craft_b200 import hotpath as hp, ops
    from craft_b200.ops import TokenGrid
    g = TokenGrid(cfg["H"] // 8, cfg["W"] // 8)
    ws = model.workspace_for(8 * g.H, 8 * g.W, dev)
    U = float(g.U)
    ub = model.update_block
    uw = ub.weights(g)
    att_tbl = model.att.vispos_encoder.table()
    f2_tbl = model.f2_trans.vispos_encoder.table()
    ks = ws.pv_split(4)
    agg = ub.aggregator.packed()
    rows = []
    up_out = torch.empty((2, cfg["H"], cfg["W"]), device=dev)

    def add(name, per_pair, fn, flops=None, bytes_=None, note=None):
        us = _time_us(fn)
        r = dict(kernel=name, launches_per_pair=per_pair, us_per_launch=us)
        if flops is not None:
            ach = flops / (us * 1e-6) / 1e12
            r.update(bound="tensor", achieved=ach, peak=pk["tf_burst"], unit="TFLOP/s", frac=ach / pk["tf_burst"],
                     flops_per_launch=flops)
        else:
            ach = bytes_ / (us * 1e-6) / 1e9
            r.update(bound="hbm", achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"], bytes_per_launch=bytes_)
        if note:
            r["note"] = note
        rows.append(r)
        return r

    it = cfg["iters"]
    pv = add("attn_pv_kernel<32,128,128> (motion aggregator softmax.V)", it,
             lambda: ops.attn_pv(ws.Qa, ws.Ka, ws.Vt, g, M=4, d=32, F=128, w_pos=1.0, pos_table=att_tbl, R=7, clip=ws.clip_att,
                                 lse2=ws.lse2_att, out=ws.opart(ks, 4, 128), ksplit=ks, zero_fill=False),
             flops=2 * U * U * 128 + 2 * 4 * U * U * 128)
    add("attn_pv_kernel<64,256,64> (F2 transformer softmax.V)", 1,
        lambda: ops.attn_pv(ws.Q2, ws.K2, ws.Vt, g, M=4, d=64, F=256, w_pos=0.5, pos_table=f2_tbl, R=7, clip=ws.clip_f2,
                            lse2=ws.lse2_f2, out=ws.opart(ks, 4, 256), ksplit=ks, zero_fill=False),
        flops=2 * U * U * 256 + 2 * 4 * U * U * 256)

    def corr():
        ws.stat_sum.zero_()
        ops.corr_build(ws.Qc, ws.Kc, g, M=4, d=64, w_agg=ws.corr_meta["w_agg"], w_pos=0.5, pos_table=f2_tbl, R=7,
                       clip=ws.inf_clip, stat_sum=ws.stat_sum[0], stat_max=ws.stat_max[3:4], levels=ws.levels, ksplit=ws.ks_sc,
                       level0_h16=ws.level0_h16)
    add("scores_kernel<SC_CORR> (4-mode correlation volume + pyramid + LN statistics)", 1, corr, flops=2 * U * U * 256,
        note="includes a 2 us stat_sum reset; level 0 mode: " + ws.level0)
    add("scores_kernel<SC_LSE> d=32 (intra-frame attention statistics)", 1,
        lambda: ops.attn_lse(ws.Qa, ws.Ka, g, M=4, d=32, w_pos=1.0, pos_table=att_tbl, R=7, clip=ws.inf_clip,
                             stat_max=ws.stat_max[3:4], lse_part=ws.lse_part, lse2=ws.lse2_att, ksplit=ws.ks_sc),
        flops=2 * U * U * 128, note="includes lse_merge")
    add("scores_kernel<SC_LSE> d=64 (F2 transformer statistics)", 1,
        lambda: ops.attn_lse(ws.Q2, ws.K2, g, M=4, d=64, w_pos=0.5, pos_table=f2_tbl, R=7, clip=ws.inf_clip,
                             stat_max=ws.stat_max[3:4], lse_part=ws.lse_part, lse2=ws.lse2_f2, ksplit=ws.ks_sc),
        flops=2 * U * U * 256, note="includes lse_merge")
    wzr, bzr, wq, bq, taps = uw.gru[0]
    add("shift_gemm<128,GRU_ZR> (SepConvGRU z,r: 1x5 conv 512->256 + gates)", 2 * it,
        lambda: ops.shift_gemm(ws.X, wzr, M=g.Mp, Npad=256, K=512, BN=128, taps=taps, grid=g, epilogue=ops.EPI_GRU_ZR,
                               bias=bzr, out_b=ws.X, colb=512, aux0=ws.Z, aux1=ws.Hm),
        flops=2 * U * 512 * 256 * 5)
    add("shift_gemm<64,GRU_Q> (SepConvGRU q: 1x5 conv 512->128 + state update)", 2 * it,
        lambda: ops.shift_gemm(ws.X, wq, M=g.Mp, Npad=128, K=512, BN=64, taps=taps, a_koff=128, grid=g, epilogue=ops.EPI_GRU_Q,
                               bias=bq, out_b=ws.X, colb=0, aux0=ws.Z, aux1=ws.Hm),
        flops=2 * U * 512 * 128 * 5)

    def update_gemms():        # every tensor-core GEMM of one refinement iteration, in model order, one stream
        hp.motion_encoder(ws, uw, None)
        BK = ops.pv_block_keys(32, 128)
        ops.shift_gemm(agg["w1"], ws.X, M=512, Npad=ops.blocked_keys(g, BK), K=128, BN=BK, b_koff=256, out_b=ws.Vt, b_block_grid=g)
        hp.sep_conv_gru(ws, uw)
        hp.heads(ws, uw, 0, need_mask=False)
    conv = lambda cin, cout, k: 2 * U * cin * cout * k
    gemm_flops = (conv(324, 256, 1) + conv(256, 192, 9) + conv(128, 64, 9) + conv(256, 126, 9)      # motion encoder (convf1 is FMA)
                  + conv(128, 512, 1)                                                              # first_linear (V)
                  + 6 * conv(512, 128, 5)                                                          # SepConvGRU
                  + conv(128, 256, 9) + conv(256, 2, 9))                                           # flow head
    add("shift-GEMM family, one refinement iteration (11 launches + convf1)", it, update_gemms, flops=gemm_flops,
        note="mask head excluded (runs in the last iteration only)")
    if ws.level0_h16 is not None:
        add("corr_lookup_kernel (all 4 levels; level 0 from the fp16 block-ordered volume)", it,
            lambda: ops.corr_lookup(ws.levels, g, ws.coords1, ws.mean_rstd, out_b=ws.CORR, level0_h16=ws.level0_h16),
            bytes_=100 * 2 * U + 3 * 100 * 4 * U + 324 * U * 2)
    else:
        add("corr_lookup0_kernel (level-0 window recomputed from Q/K rows)", it,
            lambda: ops.corr_lookup0(grid=g, coords=ws.coords1, mean_rstd=ws.mean_rstd, out_b=ws.CORR, **ws.corr_meta),
            bytes_=2 * U * 256 * 2 + 81 * U * 2, note="algorithmic bytes = Q + K rows once + 81 bf16 outputs per query; the "
            "kernel itself moves ~51 KB of key rows per query through L2")
        add("corr_lookup_kernel (pooled levels 1-3)", it,
            lambda: ops.corr_lookup(ws.levels, g, ws.coords1, ws.mean_rstd, out_b=ws.CORR, first_level=1),
            bytes_=3 * 100 * 4 * U + 243 * U * 2)
    add("modes_finalize_kernel<128> (mode soft-pool + skip + LayerNorm)", it,
        lambda: ops.modes_finalize(ws.opart(ks, 4, 128), ks, 4, 128, g, w_score=agg["ws"], b_score=agg["bs"], coeff=agg["coeff"],
                                   x_b=ws.X, colx=256, out_b=ws.X, colb=384, pv_bk=128),
        bytes_=4 * U * 128 * 4 + U * 128 * 2 * 2)
    add("upsample_flow_kernel (convex 8x upsampling)", 1 if True else it,
        lambda: ops.upsample_flow(ws.MASKS[0], ws.flow, g, out=up_out),
        bytes_=576 * U * 4 + 2 * U * 4 + 2 * cfg["H"] * cfg["W"] * 4)
    return pv, rows


def run_ours(args, cfg):
    import copy
    import torch
    import torch.distributed as dist
    from craft_b200 import _lib
    from craft_b200.sharding import pairs_for_rank

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
    _lib.load()
    model_cpu, wsrc = _build_model(cfg)
    model = copy.deepcopy(model_cpu).to(dev).eval()
    H, W, ITERS = cfg["H"], cfg["W"], cfg["iters"]
    # pairs are the units of work: rank r owns global pair indices r, r+N, ... (craft_b200/sharding.py) -- every rank
    # sees different frames
    NPAIR = 4
    mine = pairs_for_rank(NPAIR * world, rank, world)
    pairs = _pairs(cfg, mine, dev)

    def step(i):
        a, b = pairs[i % len(pairs)]
        return model(a, b, iters=ITERS, test_mode=1)

    # --lanes L: L independent pairs in flight per GPU (craft_b200.pipeline.PairStream / CRAFT.on_lane): every pair is
    # still one batch-1 forward of the whole path; the kernels of different pairs overlap on the device
    from craft_b200.pipeline import PairStream
    lanes = max(1, args.lanes)
    ps_val = PairStream(model, iters=ITERS, lanes=lanes)

    def run_steps(n):
        if lanes == 1:
            for i in range(n):
                step(i)
        else:
            # results are dropped as they are produced, like the single-lane loop does: holding K output tensors makes
            # the caching allocator call cudaMalloc (a device-wide sync) inside the timed region once K outgrows the warm-up
            ps_val.run_resident((pairs[i % len(pairs)] for i in range(n)), keep=False)

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
        # single-lane reference point (one pair at a time on one stream): the latency of a pair
        for i in range(max(args.warmup, 1)):
            step(i)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            step(i)
        s1.record()
        barrier()
        ms_single = s0.elapsed_time(s1)
        run_steps(max(args.warmup, 2 * lanes))      # every lane captures its graph and replays it once outside the timed region
        barrier()
        if sampler:
            sampler.start()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(args.steps)
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        if launches == 0:   # CUDA-graph replay: the library is not called at replay time; count kernels per graph
            launches = args.steps * int(getattr(model, "launches_last_forward", 0))
        ms = e0.elapsed_time(e1)
        # ---- end-to-end: pinned host uint8 frames -> H2D -> forward -> D2H of the full-res flow
        host_pairs = [(a.pin_memory(), b.pin_memory()) for a, b in _pairs(cfg, mine, None, uint8=True)]
        out_host = torch.empty((1, 2, H, W), dtype=torch.float32).pin_memory()

        if args.e2e_serial:
            # one pair at a time on one stream, as evaluate.py's loops do
            def e2e_step(i):
                a, b = host_pairs[i % len(host_pairs)]
                _, up = model(a.to(dev, non_blocking=True).float(), b.to(dev, non_blocking=True).float(),
                              iters=ITERS, test_mode=1)
                out_host.copy_(up, non_blocking=True)

            def e2e_run(n):
                for i in range(n):
                    e2e_step(i)
        else:
            # craft_b200.pipeline.PairStream: the same per-pair traffic (every pair's frames H2D, its flow D2H, all
            # inside the timed region) with the copies of neighbouring pairs on a second stream under the forward
            ps = PairStream(model, iters=ITERS, lanes=lanes)

            def e2e_run(n):
                got = 0
                for flow in ps.map(host_pairs[i % len(host_pairs)] for i in range(n)):
                    got += 1
                assert got == n

        e2e_run(max(2, lanes))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_run(args.steps)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, ms_e2e, ms_single], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_single = t.tolist()

    roof, gpu_ref, cpu = None, None, None
    if rank == 0:
        pk = _peaks()
        if cfg["args"].get("use_setrans", True):
            from craft_b200 import ops as _ops
            with torch.no_grad(), _ops.precision(model.act_dtype):
                pv, rows = _kernel_table(model, cfg, dev, pk)
            U = (H // 8) * (W // 8)
            traffic = None
            for name in ("r02_pv_traffic.json", "r01_pv_traffic.json"):
                tpath = os.path.join(ROOT, "profiles", name)
                if os.path.isfile(tpath):
                    tj = json.load(open(tpath))
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    break
            mufu_us = 4.0 * U * U / (15.2 * 148 * 1.965e9) * 1e6
            roof = dict(bound="tensor", kernel=pv["kernel"] + ", x%d per pair" % ITERS, achieved=pv["achieved"], peak=pv["peak"],
                        unit="TFLOP/s", frac=pv["frac"], traffic=traffic, us_per_launch=pv["us_per_launch"],
                        flops_per_launch=pv["flops_per_launch"], peak_source=pk["src"] + " bf16 burst (kernel timed alone)",
                        note="second ceiling of this kernel: one exp per (query, key, mode) at the measured 15.2/clk/SM MUFU "
                             "rate = %.1f us per launch (profiles/r01_mb_exp.txt)" % mufu_us,
                        kernels=rows)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, n, tt, kind = _cpu_rate(cfg, 20.0, threads)
            cpu = dict(value=v, unit="pairs/s", cores=threads, kind=kind,
                       sample="%d full-size pairs (%dx%d, iters=%d), %.1f s of CPU time" % (n, H, W, ITERS, tt))
        if world == 1 and args.gpu_reference:
            gpu_ref = _gpu_reference(cfg, dev, pairs)
        total = world * args.steps
        line = dict(metric=cfg["metric"], value=total / (ms * 1e-3), unit="pairs/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype={"bf16": "bf16", "fp16": "fp16", "fp32-parity": "fp16"}[model.precision],
                    data="synthetic integer-noise pairs (distinct per step and per rank), weights: " + wsrc,
                    config=dict(workload=cfg["workload"],
                                precision="%s tier: tensor-core operands %s, fp32 accumulation/state (CRAFT_B200_PRECISION)"
                                          % (model.precision, str(model.act_dtype).replace("torch.", "")),
                                parallelism="pairs sharded over %d ranks (rank r owns pairs r, r+N, ...), no collective" % world,
                                l2="inputs larger than L2: the per-step working set (68 MB pooled correlation pyramid, 30 MB "
                                   "P.V partial sums, >100 MB of encoder activations) exceeds the 126 MB L2 and every "
                                   "buffer is rewritten each step; 4 distinct input pairs rotate per rank",
                                encoders="fnet/cnet (outside the hot path): cuDNN fp16 convolutions with fp32 accumulation "
                                         "+ craft_b200 norm/relu/residual kernels",
                                launch="whole forward replayed as one CUDA graph; craft_b200 kernels use programmatic "
                                       "dependent launch",
                                lanes=lanes,
                                lanes_note="%d independent pairs in flight per GPU, one CUDA stream + workspace + graph each "
                                           "(craft_b200.pipeline.PairStream); every pair is a batch-1 forward of the whole "
                                           "path; `single_lane` is one pair at a time on one stream" % lanes,
                                single_lane=dict(value=total / (ms_single * 1e-3), unit="pairs/s",
                                                 ms_per_pair=ms_single / args.steps),
                                dead_work="test_mode=1 returns only the last upsampled flow: the mask head + convex "
                                          "upsampling of iterations 1..%d (discarded by the reference) are elided, "
                                          "outputs bit-identical (tests/test_gpu_e2e.py)" % (ITERS - 1)),
                    e2e=dict(value=total / (ms_e2e * 1e-3), unit="pairs/s", h2d_bytes_per_step=2 * 3 * H * W,
                             d2h_bytes_per_step=2 * H * W * 4,
                             mode=("serial: H2D -> CRAFT.forward -> D2H on one stream" if args.e2e_serial else
                                   "craft_b200.pipeline.PairStream(lanes=%d): every pair's uint8 frames H2D and its flow D2H "
                                   "inside the timed region, the copies on their own streams under the forwards" % lanes)),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu)
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_train(args, cfg):
    """BASELINE configs[3]: synthetic DDP training step (train_ddp.py:185-256 wiring) -- forward through
    craft_b200/train_path.py, backward, clip, AdamW.  --dropout-prob: the reference's training default keeps
    token/attention dropout on (0.1 / 0.2), which forces the attention blocks onto the PyTorch restatement in
    forward too; 0 runs their forward on the sm_100a kernels."""
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from craft_b200 import _lib
    from craft_b200.network import CRAFT
    from craft_b200.testing import craft_args, synthetic_pair
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29671")
    if world > 1 or "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    kw = dict(cfg["args"], mixed_precision=True)      # every training script of the reference passes --mixed_precision
    if args.dropout_prob is not None:
        kw["dropout_prob"] = args.dropout_prob
    torch.backends.cudnn.benchmark = True
    scaler = torch.amp.GradScaler("cuda", enabled=True)       # train.py:215,231-238
    torch.manual_seed(1234)
    model = CRAFT(craft_args(**kw)).to(dev)
    model.train()
    model.freeze_bn()
    ddp = DDP(model, device_ids=[local], find_unused_parameters=True)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1.25e-4, weight_decay=1e-5, eps=1e-8)
    B, H, W, ITERS = cfg["batch"], cfg["H"], cfg["W"], cfg["iters"]
    batches = []
    for i in range(2):
        a, b = synthetic_pair(H, W, seed=1234 + 10 * rank + i, B=B)
        batches.append((a.pin_memory(), b.pin_memory()))
    gt = torch.zeros(B, 2, H, W, device=dev)
    gt[:, 0], gt[:, 1] = 3.0, 2.0

    def step(i):
        a, b = batches[i % 2]
        preds = ddp(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True), iters=ITERS, test_mode=0)
        loss = sum(0.8 ** (ITERS - k - 1) * (p - gt).abs().mean() for k, p in enumerate(preds))
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(ddp.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        return loss.detach()

    for i in range(args.warmup):
        step(i)
    dist.barrier()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    if rank == 0:
        pairs = world * B * args.steps
        dp = model.intra_trans_config.attention_probs_dropout_prob
        line = dict(metric=cfg["metric"], value=pairs / (ms * 1e-3), unit="pairs/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="fp16 autocast (--mixed_precision, GradScaler), fp32 master weights",
                    data="synthetic integer-noise batches, random-init weights (seed 1234)",
                    config=dict(workload=cfg["workload"], dropout_prob=dp,
                                attention_forward="sm_100a kernels (dropout off)" if dp == 0 else
                                "PyTorch restatement (dropout %.1f live, reference training default)" % dp,
                                collective="DDP gradient all-reduce (NCCL) only"),
                    e2e=dict(value=pairs / (ms * 1e-3), unit="pairs/s", h2d_bytes_per_step=2 * B * 3 * H * W * 4,
                             d2h_bytes_per_step=0, note="host batches are uploaded inside the timed step"),
                    gpu_launches=int(_lib.launch_count() - n0), final_loss=float(loss))
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="sintel", choices=sorted(CONFIGS),
                    help="sintel = BASELINE configs[1] (default, the metric's configuration); kitti = configs[4] shape; "
                         "gma = configs[2] variant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("CRAFT_B200_LANES", "3")),
                    help="independent pairs in flight per GPU (1 = one pair at a time on one stream)")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e leg: one pair at a time on one stream (no PairStream overlap)")
    ap.add_argument("--dropout-prob", type=float, default=None,
                    help="--config train only: override the transformers' dropout (reference training default 0.1/0.2)")
    ap.add_argument("--gpu-reference", action="store_true",
                    help="also time the unmodified reference in eager CUDA (fp32 and fp16 autocast) on the same GPU "
                         "(informational leg `gpu_reference`, N=1 only)")
    ap.add_argument("--ncu", action="store_true",
                    help="profiling helper: warm up, then run ONE step inside a cudaProfilerStart/Stop range and exit "
                         "(use with `ncu --profile-from-start off ...`); prints no bench line")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.config == "train":
        if args.impl == "reference":
            print(json.dumps(dict(impl="reference", unavailable="the reference's training step needs CUDA + datasets; "
                                  "the CPU arm times inference configs only")))
        else:
            run_train(args, cfg)
    elif args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
