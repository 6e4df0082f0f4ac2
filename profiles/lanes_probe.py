"""Feasibility probe: throughput of L independent pairs in flight on one GPU (one CUDA stream + one model replica
-- own workspace, own graph -- per lane) against one pair at a time.  usage: python profiles/lanes_probe.py [lanes...]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ITERS, _pairs, _state_dict  # noqa: E402


def main(lane_counts):
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    base, _ = _state_dict()
    pairs = _pairs(4, dev)
    N = 48
    for L in lane_counts:
        models = [copy.deepcopy(base).to(dev).eval() for _ in range(L)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(L)]
        prio = os.environ.get("LANE_PRIO")
        if prio:
            streams = [torch.cuda.Stream(device=dev, priority=-1 if (k % 2 == 0) else 0) for k in range(L)]
        with torch.no_grad():
            for k in range(L):
                with torch.cuda.stream(streams[k]):
                    for i in range(3):
                        models[k](*pairs[i % 4], iters=ITERS, test_mode=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main_s = torch.cuda.current_stream()
            e0.record(main_s)
            for s in streams:
                s.wait_event(e0)
            outs = []
            for i in range(N):
                k = i % L
                with torch.cuda.stream(streams[k]):
                    outs.append(models[k](*pairs[i % 4], iters=ITERS, test_mode=1)[1])
            for s in streams:
                main_s.wait_stream(s)
            e1.record(main_s)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("lanes %d: %d pairs in %.2f ms -> %.1f pairs/s (%.3f ms per pair)" % (L, N, ms, N / ms * 1e3, ms / N), flush=True)
        del models
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main([int(x) for x in sys.argv[1:]] or [1, 2, 3])
