"""Host I/O around the hot path: .flo / KITTI-png writers and readers against
the reference's byte layout (core/utils/frame_utils.py) and the warm-start forward_interpolate
(core/utils/utils.py:34-62) against scipy's griddata, which the reference calls."""
import os
import struct

import numpy as np
import pytest
import torch

from craft_b200.utils import frame_utils as FU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flow(h=37, w=53, seed=0, scale=30.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn((h, w, 2), generator=g) * scale).numpy().astype(np.float32)


def test_flo_bytes_and_round_trip(tmp_path):
    uv = _flow()
    p = str(tmp_path / "a.flo")
    FU.writeFlow(p, uv)
    raw = open(p, "rb").read()
    # frame_utils.py:88-98: 'PIEH', width, height, interleaved float32 rows
    assert raw[:4] == b"PIEH" and struct.unpack("<ii", raw[4:12]) == (53, 37)
    assert len(raw) == 12 + 37 * 53 * 2 * 4
    assert np.array_equal(np.frombuffer(raw[12:], np.float32).reshape(37, 53, 2), uv)
    assert np.array_equal(FU.readFlow(p), uv)
    FU.writeFlow(p, uv[..., 0], uv[..., 1])          # the (u, v) calling form
    assert np.array_equal(FU.readFlow(p), uv)
    open(p, "wb").write(b"XXXX" + raw[4:])
    assert FU.readFlow(p) is None


def test_kitti_png_quantisation_and_round_trip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    uv = _flow(scale=80.0)
    p = str(tmp_path / "a.png")
    FU.writeFlowKITTI(p, uv)
    img = cv2.imread(p, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
    assert img.dtype == np.uint16 and img.shape == (37, 53, 3)
    # frame_utils.py:116-120: channels R,G,B = (64u+2^15, 64v+2^15, 1) truncated to uint16
    want = np.concatenate([64.0 * uv + 2 ** 15, np.ones((37, 53, 1))], -1).astype(np.uint16)
    assert np.array_equal(img[..., ::-1], want)
    back, valid = FU.readFlowKITTI(p)
    assert np.all(valid == 1) and np.abs(back - uv).max() <= 1.0 / 64.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/core"), reason="reference tree not mounted")
def test_writers_match_the_reference_files_byte_for_byte(tmp_path):
    import importlib.util
    pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("_ref_frame_utils", "/root/reference/core/utils/frame_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    uv = _flow(seed=3)
    a, b = str(tmp_path / "ours.flo"), str(tmp_path / "ref.flo")
    FU.writeFlow(a, uv); ref.writeFlow(b, uv)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert np.array_equal(ref.readFlow(a), FU.readFlow(b))
    a, b = str(tmp_path / "ours.png"), str(tmp_path / "ref.png")
    FU.writeFlowKITTI(a, uv); ref.writeFlowKITTI(b, uv)
    fa, va = ref.readFlowKITTI(a)
    fb, vb = FU.readFlowKITTI(b)
    assert np.array_equal(fa, fb) and np.array_equal(va, vb)


def test_forward_interpolate_refuses_without_a_device():
    from craft_b200.utils.utils import forward_interpolate
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        forward_interpolate(torch.zeros(2, 8, 8))


@pytest.mark.gpu
def test_device_encoders_match_the_host_writers(tmp_path):
    pytest.importorskip("cv2")
    uv = _flow(h=55, w=128, seed=5, scale=60.0)
    t = torch.from_numpy(uv).cuda()
    for make in (lambda: t, lambda: t.permute(2, 0, 1).contiguous()):
        a, b = str(tmp_path / "dev.flo"), str(tmp_path / "host.flo")
        FU.writeFlow(a, make()); FU.writeFlow(b, uv)
        assert open(a, "rb").read() == open(b, "rb").read()
        a, b = str(tmp_path / "dev.png"), str(tmp_path / "host.png")
        FU.writeFlowKITTI(a, make()); FU.writeFlowKITTI(b, uv)
        fa, _ = FU.readFlowKITTI(a)
        fb, _ = FU.readFlowKITTI(b)
        assert np.array_equal(fa, fb)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,scale", [(16, 24, 1.5), (55, 128, 6.0), (48, 156, 20.0)])
def test_forward_interpolate_matches_scipy_griddata(h, w, scale):
    """Same computation as the reference (scipy griddata 'nearest' over the landed points, utils.py:53-60)."""
    from scipy import interpolate
    from craft_b200.utils.utils import forward_interpolate
    g = torch.Generator().manual_seed(h * w)
    flow = torch.randn((2, h, w), generator=g) * scale
    got = forward_interpolate(flow.cuda()).cpu().numpy()
    assert got.shape == (2, h, w)
    f = flow.numpy()
    dx, dy = f[0], f[1]
    x0, y0 = np.meshgrid(np.arange(w), np.arange(h))
    x1, y1 = (x0 + dx).reshape(-1), (y0 + dy).reshape(-1)
    valid = (x1 > 0) & (x1 < w) & (y1 > 0) & (y1 < h)
    ref = np.stack([interpolate.griddata((x1[valid], y1[valid]), d.reshape(-1)[valid], (x0, y0), method="nearest",
                                         fill_value=0) for d in (dx, dy)], 0)
    same = np.all(got == ref, axis=0)
    # exact nearest neighbour: any disagreement can only be a tie (two landed points at the same distance)
    assert same.mean() >= 0.999, same.mean()
    if not same.all():
        ys, xs = np.nonzero(~same)
        px, py = x1[valid], y1[valid]
        vdx, vdy = dx.reshape(-1)[valid], dy.reshape(-1)[valid]
        for yy, xx in zip(ys, xs):
            d = (px - xx) ** 2 + (py - yy) ** 2
            mine = np.nonzero((vdx == got[0, yy, xx]) & (vdy == got[1, yy, xx]))[0]
            assert mine.size and d[mine].min() <= d.min() * (1 + 1e-5) + 1e-6
    # CPU tensors are accepted like in the reference (moved to the device and back)
    assert np.array_equal(forward_interpolate(flow).numpy(), got)
