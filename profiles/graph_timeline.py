"""Kernel timeline of ONE CUDA-graph replay of the forward (448x1024, 12 iterations) via CUPTI
(torch.profiler): true in-situ durations (warm caches, back-to-back) plus the idle gaps between
kernels.  usage: python profiles/graph_timeline.py [tag]"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import ITERS, _pairs, _state_dict  # noqa: E402


def main(tag):
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    model, _ = _state_dict()
    model = model.to(dev).eval()
    a, b = _pairs(1, dev)[0]
    with torch.no_grad():
        for _ in range(4):
            model(a, b, iters=ITERS, test_mode=1)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model(a, b, iters=ITERS, test_mode=1)
            torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.cuda_time_total > 0]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs))
    if not ks:
        print("no kernel events")
        return
    span = ks[-1][1] - ks[0][0]
    busy = sum(e - s for s, e, _ in ks)
    agg = defaultdict(lambda: [0, 0.0])
    for s, e, n in ks:
        n = n.replace("void ", "").split("(")[0][:78]
        agg[n][0] += 1
        agg[n][1] += e - s
    lines = ["graph replay: %d kernels, span %.2f ms, sum of kernel durations %.2f ms (streams overlap / gaps: %.2f ms)"
             % (len(ks), span / 1e3, busy / 1e3, (span - busy) / 1e3),
             "%-80s %5s %9s %6s %8s" % ("kernel", "count", "total_us", "share", "avg_us")]
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append("%-80s %5d %9.1f %5.1f%% %8.2f" % (n, c, us, 100 * us / busy, us / c))
    out = "\n".join(lines)
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "graph_timeline_%s.txt" % tag), "w").write(out + "\n")
    # the raw timeline (start relative to the first kernel, duration, name), for reading overlap and gaps
    t0 = ks[0][0]
    with open(os.path.join(ROOT, "gpurun_out", "graph_timeline_raw_%s.txt" % tag), "w") as f:
        for s, e, n in ks:
            f.write("%9.1f %8.1f  %s\n" % (s - t0, e - s, n.replace("void ", "").split("(")[0][:90]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
