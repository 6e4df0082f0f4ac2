"""craft_conv3x3_c64 against cuDNN on the encoders' layer-1 shapes (fp16, channels-last), each timed alone:
20 launches in one CUDA graph, CUDA events.  usage: python profiles/conv64_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from bench import _time_us  # noqa: E402
from craft_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    g = torch.Generator(device=dev).manual_seed(1)
    for N, H, W in ((2, 224, 512), (1, 224, 512), (2, 192, 624)):
        x = torch.randn((N, H, W, 64), device=dev, generator=g).half()
        w = torch.randn((64, 64, 3, 3), device=dev, generator=g) * 0.05
        wcl = w.half().contiguous(memory_format=torch.channels_last)
        xn = x.permute(0, 3, 1, 2)
        t = torch.zeros((N, H + 1, W + 2, 64), dtype=torch.float16, device=dev)
        t[:, :H, :W] = x
        xp = ops.PadAct(t.reshape(-1, 64), N, H, W)
        wp = ops.pack_conv64_weight(w)
        b = torch.zeros(64, device=dev)
        us_cudnn = _time_us(lambda: F.conv2d(xn, wcl, None, 1, 1))
        us_own = _time_us(lambda: ops.conv3x3_c64(xp, wp))
        us_own_stats = _time_us(lambda: ops.conv3x3_c64(xp, wp, stats_eps=1e-5))
        us_own_relu = _time_us(lambda: ops.conv3x3_c64(xp, wp, bias=b, relu=True))
        us_stats = _time_us(lambda: ops.instnorm_stats(x))
        ab = ops.instnorm_stats(x)
        us_aff = _time_us(lambda: ops.nhwc_affine(x, ab, relu_in=True))
        us_affp = _time_us(lambda: ops.nhwc_affine_pad(xp, ab, relu_in=True))
        us_affres = _time_us(lambda: ops.nhwc_affine(x, ab, res=x, relu_in=True, relu_out=True))
        us_affpres = _time_us(lambda: ops.nhwc_affine_pad(xp, ab, res=xp, relu_in=True, relu_out=True))
        fl = 2.0 * N * H * W * 64 * 64 * 9
        print("N=%d %dx%d: cuDNN %.1f us (%.0f TF/s) | own %.1f us (%.0f TF/s), +stats %.1f, +bias/relu %.1f | stats kernels %.1f us | "
              "affine dense %.1f / padded %.1f, with residual %.1f / %.1f us"
              % (N, H, W, us_cudnn, fl / us_cudnn / 1e6, us_own, fl / us_own / 1e6, us_own_stats, us_own_relu, us_stats, us_aff, us_affp,
                 us_affres, us_affpres), flush=True)


if __name__ == "__main__":
    with torch.no_grad():
        main()
