"""Flow file formats with the reference's function names (core/utils/frame_utils.py): Middlebury
`.flo` (readFlow :12-31, writeFlow :70-99) and KITTI 16-bit png (readFlowKITTI :102-108,
writeFlowKITTI :116-120).  numpy arrays work as in the reference; CUDA tensors ([2,H,W] or [H,W,2])
are re-packed on the device (csrc/hostio.cuh flow_encode) and copied out once, already in file order.
"""
import numpy as np
import torch

TAG_FLOAT = 202021.25          # 'PIEH' read as a little-endian float (frame_utils.py:10,17)
TAG_CHAR = b"PIEH"


def _planes(uv):
    """CUDA tensor [2,H,W] or [H,W,2] -> contiguous f32 planes [2,H,W]."""
    t = uv.detach().float()
    if t.dim() != 3:
        raise ValueError("flow must be [2,H,W] or [H,W,2]")
    if t.shape[0] != 2 and t.shape[2] == 2:
        t = t.permute(2, 0, 1)
    return t.contiguous()


def _encode_on_device(uv, mode):
    from .. import ops
    with torch.cuda.device(uv.device):
        p = _planes(uv)
        H, W = p.shape[1:]
        if mode == 0:
            out = torch.empty((H, W, 2), dtype=torch.float32, device=p.device)
        else:
            out = torch.empty((H, W, 3), dtype=torch.int16, device=p.device)      # uint16 payload
        ops.OPS.flow_encode(p, H, W, mode, out)
        host = out.cpu().numpy()
    return host if mode == 0 else host.view(np.uint16)


def writeFlow(filename, uv, v=None):
    """frame_utils.py:70-99: tag 'PIEH', int32 width, int32 height, then rows of interleaved (u, v) float32."""
    if isinstance(uv, torch.Tensor) and uv.is_cuda and v is None:
        data = _encode_on_device(uv, 0)
    else:
        if isinstance(uv, torch.Tensor):
            uv = uv.detach().cpu().numpy()
        if v is None:
            assert uv.ndim == 3 and uv.shape[2] == 2
            u, v = uv[:, :, 0], uv[:, :, 1]
        else:
            u = uv
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
        assert u.shape == v.shape
        data = np.stack([u, v], axis=-1).astype(np.float32)
    height, width = data.shape[:2]
    with open(filename, "wb") as f:
        f.write(TAG_CHAR)
        np.array(width).astype(np.int32).tofile(f)
        np.array(height).astype(np.int32).tofile(f)
        np.ascontiguousarray(data, dtype=np.float32).tofile(f)


def readFlow(fn):
    """frame_utils.py:12-31 -> [H,W,2] float32 (None on a bad magic number, like the reference)."""
    with open(fn, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size == 0 or magic[0] != np.float32(TAG_FLOAT):
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
    return np.resize(data, (h, w, 2))


def writeFlowKITTI(filename, uv):
    """frame_utils.py:116-120: 16-bit png, (64 u + 2^15, 64 v + 2^15, valid = 1), stored B,G,R by cv2."""
    import cv2
    if isinstance(uv, torch.Tensor) and uv.is_cuda:
        bgr = _encode_on_device(uv, 1)
    else:
        if isinstance(uv, torch.Tensor):
            uv = uv.detach().cpu().numpy()
        q = 64.0 * uv + 2 ** 15
        valid = np.ones([uv.shape[0], uv.shape[1], 1])
        bgr = np.concatenate([q, valid], axis=-1).astype(np.uint16)[..., ::-1]
    cv2.imwrite(filename, np.ascontiguousarray(bgr))


def readFlowKITTI(filename):
    """frame_utils.py:102-108 -> (flow [H,W,2] float32, valid [H,W])."""
    import cv2
    flow = cv2.imread(filename, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
    flow = flow[:, :, ::-1].astype(np.float32)
    flow, valid = flow[:, :, :2], flow[:, :, 2]
    return (flow - 2 ** 15) / 64.0, valid
