"""Per-kernel device-time breakdown of one CRAFT forward (448x1024, iters=12) with CUDA events around
every C-ABI call (warm caches, launch order preserved).  Complements the ncu launch list: ncu times
are cold-cache/serialised, these are in-situ.  Usage (on the GPU box): python profiles/breakdown.py [tag]"""
import json
import os
import sys
import time
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import H, ITERS, W, _pairs, _state_dict  # noqa: E402
from craft_b200 import _lib  # noqa: E402


def main(tag):
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    model, _ = _state_dict()
    model = model.to(dev).eval()
    model.use_cuda_graph = False     # per-call events need eager launches
    a, b = _pairs(1, dev)[0]
    with torch.no_grad():
        for _ in range(3):
            model(a, b, iters=ITERS, test_mode=1)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            model(a, b, iters=ITERS, test_mode=1)
        torch.cuda.synchronize()
        wall = (time.time() - t0) / 5
        events = []
        orig = _lib.call

        def traced(name, *args):
            key = name
            if name == "craft_shift_gemm":
                g = args[0]._obj
                key = "gemm[M=%d,N=%d,K=%d,T=%d,BN=%d,epi=%d]" % (g.M, g.Npad, g.K, g.T, g.BN, g.epilogue)
            elif name == "craft_attn_pv":
                p = args[0]._obj
                key = "attn_pv[d=%d,F=%d,ks=%d]" % (p.d, p.F, p.ksplit)
            elif name in ("craft_corr_build", "craft_attn_lse"):
                p = args[0]._obj
                key = "%s[C=%d,M=%d,flagged=%d]" % (name, p.C, p.M, 1 if p.run_flag else 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig(name, *args)
            e1.record()
            events.append((key, e0, e1))

        _lib.call = traced
        import craft_b200.ops as ops_mod
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        model(a, b, iters=ITERS, test_mode=1)
        s1.record()
        torch.cuda.synchronize()
        _lib.call = orig
        # encoders alone
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(5):
            model._encoders(a, b)
        x1.record()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for k, e0, e1 in events:
        agg[k][0] += 1
        agg[k][1] += 1000 * e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    lines = ["wall per forward (no tracing): %.2f ms ; traced forward span: %.2f ms ; sum of kernel spans: %.2f ms ; "
             "fnet+cnet: %.2f ms" % (1000 * wall, s0.elapsed_time(s1), tot / 1000, x0.elapsed_time(x1) / 5),
             "%-58s %6s %10s %7s %9s" % ("call", "count", "total_us", "share", "avg_us")]
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-58s %6d %10.1f %6.1f%% %9.2f" % (k, c, us, 100 * us / tot, us / c))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "breakdown_%s.txt" % tag), "w").write(out + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
