"""Host utilities with the reference's names (core/utils/utils.py).  `bilinear_sampler` exists in
the reference only as the backend of CorrBlock.__call__; here that lookup is a fused kernel
(pointwise.cuh corr_lookup), so only the grid / padding helpers remain."""
import os

import torch
import torch.nn.functional as F


def print0(*args, **kwargs):
    if int(os.environ.get("LOCAL_RANK", 0)) == 0:
        print(*args, **kwargs)


class InputPadder:
    """core/utils/utils.py:14-31 -- replicate-pad to a multiple of `mod`."""

    def __init__(self, dims, mode="sintel", mod=8):
        self.ht, self.wd = dims[-2:]
        pad_ht = (((self.ht // mod) + 1) * mod - self.ht) % mod
        pad_wd = (((self.wd // mod) + 1) * mod - self.wd) % mod
        if mode == "sintel":
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        c = [self._pad[2], ht - self._pad[3], self._pad[0], wd - self._pad[1]]
        return x[..., c[0]:c[1], c[2]:c[3]]


def coords_grid(batch, ht, wd, device=None):
    """core/utils/utils.py:82-85 -- channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].expand(batch, -1, -1, -1)


def upflow8(flow, mode="bilinear"):
    new_size = (8 * flow.shape[2], 8 * flow.shape[3])
    return 8 * F.interpolate(flow, size=new_size, mode=mode, align_corners=True)


def forward_interpolate(flow):
    """core/utils/utils.py:34-62 -- warm start between consecutive frames (evaluate.py:146-147): every
    pixel carries its flow to where it lands, the output at a grid point is the flow of the nearest landed
    point.  flow: [2,h,w] tensor.  The reference does this on the CPU with two scipy griddata calls per
    frame; here it is one exact brute-force nearest-neighbour kernel (csrc/hostio.cuh) and the result stays
    on the flow's device.  A CPU tensor is moved to the current CUDA device and back (no CPU path exists)."""
    from .. import ops
    if flow.dim() != 3 or flow.shape[0] != 2:
        raise ValueError("forward_interpolate expects a [2,h,w] flow, got %s" % (list(flow.shape),))
    src = flow.detach()
    was_cpu = not src.is_cuda
    if was_cpu:
        if not torch.cuda.is_available():
            raise RuntimeError("craft_b200.forward_interpolate needs a CUDA device: there is no CPU path")
        src = src.cuda()
    with torch.cuda.device(src.device):
        src = src.float().contiguous()
        out = torch.zeros_like(src)
        ops.OPS.forward_interpolate(src, src.shape[1], src.shape[2], out)
    return out.cpu() if was_cpu else out
