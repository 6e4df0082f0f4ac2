"""Micro-benchmark of the shift-GEMM main loop: time vs pipeline depth / tile width / cluster size,
on the GRU-shaped problem (M=7280 tokens, K=512 x 5 taps).  Informs DESIGN.md section 7."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from craft_b200 import ops
from craft_b200.ops import TokenGrid
g = TokenGrid(56, 128)
dev = "cuda"
X = (torch.randn((g.Mp, 640), device=dev) * 0.5).to(torch.bfloat16)
def time_graph(f, reps=20):
    """Device time per call with the host out of the picture: `reps` launches captured into one CUDA
    graph, replayed 3x, timed with events around a replay (back-to-back kernels, warm L2)."""
    for _ in range(2): f()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        f()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for _ in range(reps): f()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return 1000 * e0.elapsed_time(e1) / reps


def run(N, BN, T, K, cluster, stages, lda640=True, reps=20):
    A = X if lda640 else X[:, :K].contiguous()
    Wt = (torch.randn((T * N, K), device=dev) * 0.05).to(torch.bfloat16)
    out = torch.zeros((g.Mp, N), dtype=torch.bfloat16, device=dev)
    taps = list(range(-(T // 2), T // 2 + 1)) if T > 1 else [0]
    f = lambda: ops.shift_gemm(A, Wt, M=g.Mp, Npad=N, K=K, BN=BN, taps=taps, grid=g, out_b=out, cluster=cluster, stages=stages)
    us = time_graph(f, reps)
    fl = 2.0 * g.Mp * N * K * T
    print("N=%3d BN=%3d T=%d K=%3d cluster=%d stages=%2d lda=%s : %7.2f us  %6.1f TFLOP/s  (%.0f ns per k-iter)" % (
        N, BN, T, K, cluster, stages, "640" if lda640 else "K", us, fl / us / 1e6, 1000 * us / (T * K / 64)), flush=True)
for st in (2, 3, 4, 6):
    run(256, 128, 5, 512, 1, st)
for st in (2, 4, 6):
    run(256, 128, 5, 512, 4, st)
for st in (2, 4, 8):
    run(256, 64, 5, 512, 4, st)
run(256, 256, 5, 512, 1, 4)
run(256, 256, 5, 512, 2, 4)
run(256, 128, 1, 512, 1, 6)
run(256, 128, 1, 512, 4, 6)
run(256, 128, 5, 512, 4, 6, lda640=False)
run(128, 64, 5, 512, 4, 8)
run(128, 128, 5, 512, 4, 6)
run(128, 32, 5, 512, 4, 10)
