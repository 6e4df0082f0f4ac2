"""Kernel-level parity on the GPU: every C-ABI entry point against the oracle restatement
(oracle/restate.py, plain torch fp32) on identical, bf16-representable inputs."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from craft_b200 import ops                      # noqa: E402
from craft_b200.ops import TokenGrid            # noqa: E402
from oracle import restate as R                 # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bf16r(t):
    return t.to(torch.bfloat16).float()


def rows_from_nchw(x, grid, cols=None, col0=0, dtype=torch.bfloat16):
    """[C,H,W] -> padded-flat token rows (host-side reference packer)."""
    Cc = x.shape[0]
    buf = torch.zeros((grid.H, grid.Wp, cols or Cc), dtype=torch.float32, device=x.device)
    buf[:, : grid.W, col0:col0 + Cc] = x.permute(1, 2, 0)
    return buf.reshape(grid.Mp, -1).to(dtype).contiguous()


def nchw_from_rows(buf, grid, Cc, col0=0):
    return buf.float().reshape(grid.H, grid.Wp, -1)[:, : grid.W, col0:col0 + Cc].permute(2, 0, 1).contiguous()


@pytest.mark.parametrize("BN,Npad,K,M", [(128, 128, 64, 128), (128, 256, 256, 300), (64, 192, 128, 1000),
                                         (32, 32, 320, 129), (256, 512, 512, 777)])
def test_gemm_plain(BN, Npad, K, M):
    g = torch.Generator(device=DEV).manual_seed(BN + K)
    A = bf16r(torch.randn((M, K), device=DEV, generator=g))
    B = bf16r(torch.randn((Npad, K), device=DEV, generator=g) * 0.1)
    bias = torch.randn((Npad,), device=DEV, generator=g)
    out_b = torch.zeros((M, Npad), dtype=torch.bfloat16, device=DEV)
    out_f = torch.zeros((M, Npad), dtype=torch.float32, device=DEV)
    ops.shift_gemm(A.to(torch.bfloat16), B.to(torch.bfloat16), M=M, Npad=Npad, K=K, BN=BN, bias=bias, act=1,
                   alpha=0.5, out_b=out_b, out_f=out_f)
    torch.cuda.synchronize()
    ref = torch.relu(0.5 * (A @ B.t()) + bias)
    assert torch.allclose(out_f, ref, atol=2e-3, rtol=1e-3), (out_f - ref).abs().max()
    assert torch.allclose(out_b.float(), ref, atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("kh,kw,Cin,Cout,BN", [(3, 3, 64, 128, 128), (1, 5, 128, 64, 64), (5, 1, 192, 256, 128),
                                               (1, 1, 384, 256, 256), (3, 3, 256, 32, 32), (3, 3, 256, 192, 96)])
@pytest.mark.parametrize("a_share", [0, 1])
def test_gemm_conv(kh, kw, Cin, Cout, BN, a_share):
    grid = TokenGrid(13, 21)
    g = torch.Generator(device=DEV).manual_seed(kh * 10 + kw)
    x = bf16r(torch.randn((Cin, grid.H, grid.W), device=DEV, generator=g))
    w = bf16r(torch.randn((Cout, Cin, kh, kw), device=DEV, generator=g) * 0.05)
    b = torch.randn((Cout,), device=DEV, generator=g)
    X = rows_from_nchw(x, grid)
    Wp = ops.pack_conv_weight(w)
    out = grid.zeros(Cout, dtype=torch.float32)
    ops.shift_gemm(X, Wp, M=grid.Mp, Npad=Cout, K=Wp.shape[1], BN=BN, taps=ops.conv_taps(kh, kw, grid), grid=grid,
                   bias=b, act=0, out_f=out, a_share=a_share)
    torch.cuda.synchronize()
    ref = F.conv2d(x[None], w, b, padding=(kh // 2, kw // 2))[0]
    got = nchw_from_rows(out, grid, Cout)
    assert torch.allclose(got, ref, atol=3e-3, rtol=1e-3), (got - ref).abs().max()
    # halo rows must stay untouched (zero)
    assert out.reshape(grid.H, grid.Wp, -1)[:, grid.W:].abs().max() == 0


def test_gemm_gru_epilogues():
    grid = TokenGrid(10, 18)
    g = torch.Generator(device=DEV).manual_seed(5)
    h = torch.tanh(torch.randn((128, grid.H, grid.W), device=DEV, generator=g))
    x = bf16r(torch.randn((384, grid.H, grid.W), device=DEV, generator=g))
    P = {}
    for n in "zrq":
        P["conv%s1.weight" % n] = bf16r(torch.randn((128, 512, 1, 5), device=DEV, generator=g) * 0.03)
        P["conv%s1.bias" % n] = torch.randn((128,), device=DEV, generator=g) * 0.1
    # X buffer: [h | x | rh]
    X = grid.zeros(640)
    X[:, :512] = rows_from_nchw(torch.cat([h, x], 0), grid)
    Hm = rows_from_nchw(h, grid, dtype=torch.float32)
    Z = grid.zeros(128, dtype=torch.float32)
    Wzr = ops.pack_conv_weight(torch.cat([P["convz1.weight"], P["convr1.weight"]], 0))
    bzr = torch.cat([P["convz1.bias"], P["convr1.bias"]]).contiguous()
    perm = list(range(128, 512)) + list(range(0, 128))       # q conv sees [x | r*h]
    Wq = ops.pack_conv_weight(P["convq1.weight"], cin_perm=perm)
    taps = ops.conv_taps(1, 5, grid)
    hb = bf16r(h)
    ops.shift_gemm(X, Wzr, M=grid.Mp, Npad=256, K=512, BN=128, taps=taps, grid=grid, epilogue=ops.EPI_GRU_ZR,
                   bias=bzr, out_b=X, colb=512, aux0=Z, aux1=Hm)
    ops.shift_gemm(X, Wq, M=grid.Mp, Npad=128, K=512, BN=64, taps=taps, a_koff=128, grid=grid,
                   epilogue=ops.EPI_GRU_Q, bias=P["convq1.bias"].contiguous(), out_b=X, colb=0, aux0=Z, aux1=Hm)
    torch.cuda.synchronize()
    hx = torch.cat([hb, x], 0)[None]
    z = torch.sigmoid(F.conv2d(hx, P["convz1.weight"], P["convz1.bias"], padding=(0, 2)))
    r = torch.sigmoid(F.conv2d(hx, P["convr1.weight"], P["convr1.bias"], padding=(0, 2)))
    rh = bf16r(r * h[None])
    q = torch.tanh(F.conv2d(torch.cat([rh, x[None]], 1), P["convq1.weight"], P["convq1.bias"], padding=(0, 2)))
    ref = ((1 - z) * h[None] + z * q)[0]
    got = nchw_from_rows(Hm, grid, 128)
    assert torch.allclose(got, ref, atol=5e-3), (got - ref).abs().max()
    got_b = nchw_from_rows(X, grid, 128)
    assert torch.allclose(got_b, ref, atol=2e-2)


@pytest.mark.parametrize("Cc,mode", [(256, 1), (128, 2), (128, 3), (256, 0)])
def test_pack_unpack(Cc, mode):
    grid = TokenGrid(11, 37)
    x = torch.randn((Cc, grid.H, grid.W), device=DEV) * 3 + 0.5
    ob = grid.zeros(Cc + 64)
    of = grid.zeros(Cc, dtype=torch.float32)
    ops.pack_tokens(x, grid, mode, out_b=ob, colb=64, out_f=of)
    back = ops.unpack_tokens(of, 0, Cc, grid)
    back_b = ops.unpack_tokens(ob, 64, Cc, grid)
    torch.cuda.synchronize()
    if mode == 1:
        ref = R.encode_tokens(x[None])[0].t().reshape(Cc, grid.H, grid.W)
    elif mode == 2:
        ref = torch.tanh(x)
    elif mode == 3:
        ref = torch.relu(x)
    else:
        ref = x
    assert torch.allclose(back, ref, atol=1e-5, rtol=1e-5), (back - ref).abs().max()
    assert torch.allclose(back_b, ref, atol=3e-2, rtol=1e-2)
    assert of.reshape(grid.H, grid.Wp, -1)[:, grid.W:].abs().max() == 0


def _level0_encode(vol):
    """[Mp, H, W] fp32 -> the 16-bit level-0 format of include/craft_b200.h (craft_scores_args.lvl0_h16): fp16 deltas in
    8x8-key-block order followed by the fp32 means of the 4x8 half blocks."""
    Mp, H, W = vol.shape
    nby, nbx = (H + 7) // 8, (W + 7) // 8
    pad = torch.zeros((Mp, nby * 8, nbx * 8), device=vol.device)
    pad[:, :H, :W] = vol
    blk = pad.reshape(Mp, nby, 2, 4, nbx, 8).permute(0, 1, 4, 2, 3, 5)          # [Mp, by, bx, half, 4, 8]
    base = blk.mean(dim=(4, 5))                                                  # [Mp, by, bx, 2]
    delta = (blk - base[..., None, None]).half()
    return torch.cat([delta.reshape(-1), base.contiguous().reshape(-1).view(torch.float16)]).contiguous()


def _level0_decode(buf, grid):
    """-> [Mp, H, W] fp32 volume held by a lvl0_h16 buffer."""
    Mp, H, W = grid.Mp, grid.H, grid.W
    nby, nbx = (H + 7) // 8, (W + 7) // 8
    n = Mp * nby * nbx * 64
    delta = buf[:n].reshape(Mp, nby, nbx, 2, 4, 8).float()
    base = buf[n:n + Mp * nby * nbx * 4].view(torch.float32).reshape(Mp, nby, nbx, 2)
    v = (delta + base[..., None, None]).permute(0, 1, 3, 4, 2, 5).reshape(Mp, nby * 8, nbx * 8)
    return v[:, :H, :W]


@pytest.mark.parametrize("H,W", [(16, 24), (17, 22)])
def test_corr_lookup(H, W):
    grid = TokenGrid(H, W)
    g = torch.Generator(device=DEV).manual_seed(3)
    vol = torch.randn((1, grid.U, H, W), device=DEV, generator=g) * 2 + 0.7
    mean, rstd = 0.3, 1.7
    normed = (vol - mean) * rstd
    pyr_ref = R.corr_pyramid(normed)
    coords = R.coords_grid(1, H, W, DEV) + torch.randn((1, 2, H, W), device=DEV, generator=g) * 4
    ref = R.corr_lookup(pyr_ref, coords)[0]
    # device layout: query rows are padded-flat
    pyr_raw = R.corr_pyramid(vol)
    levels = []
    for lv in pyr_raw:
        hl, wl = lv.shape[-2:]
        buf = torch.zeros((grid.H, grid.Wp, hl * wl), device=DEV)
        buf[:, :W] = lv.reshape(H, W, hl * wl)
        levels.append(buf.reshape(grid.Mp, hl * wl).contiguous())
    cbuf = rows_from_nchw(coords[0], grid, dtype=torch.float32)
    stats = torch.tensor([mean, rstd], device=DEV)
    out_b = grid.zeros(384)
    out_n = torch.zeros((324, H, W), device=DEV)
    ops.corr_lookup(levels, grid, cbuf, stats, out_b=out_b, out_nchw=out_n)
    torch.cuda.synchronize()
    assert torch.allclose(out_n, ref, atol=2e-4, rtol=1e-4), (out_n - ref).abs().max()
    got_b = nchw_from_rows(out_b, grid, 324)
    assert torch.allclose(got_b, ref, atol=5e-2, rtol=1e-2)
    # the same lookup with level 0 held as fp16 8x8-key blocks (the default model path): bit-identical to the
    # fp32 lookup of the fp16-rounded volume
    l0h = _level0_encode(levels[0].reshape(grid.Mp, H, W))
    out_h = torch.zeros((324, H, W), device=DEV)
    ops.corr_lookup([None] + levels[1:], grid, cbuf, stats, out_nchw=out_h, level0_h16=l0h)
    out_r = torch.zeros((324, H, W), device=DEV)
    ops.corr_lookup([_level0_decode(l0h, grid).reshape(grid.Mp, H * W).contiguous()] + levels[1:], grid, cbuf, stats, out_nchw=out_r)
    torch.cuda.synchronize()
    assert torch.allclose(out_h, out_r, atol=1e-5, rtol=1e-5)      # same values, the subtraction of the mean re-associated
    assert torch.allclose(out_h, ref, atol=3e-3, rtol=1e-3), (out_h - ref).abs().max()


def test_upsample_and_small_kernels():
    grid = TokenGrid(9, 14)
    g = torch.Generator(device=DEV).manual_seed(4)
    flow = torch.randn((1, 2, grid.H, grid.W), device=DEV, generator=g) * 3
    mask = torch.randn((1, 576, grid.H, grid.W), device=DEV, generator=g)
    ref = R.upsample_flow(flow, mask)[0]
    fb = rows_from_nchw(flow[0], grid, dtype=torch.float32)
    mb = rows_from_nchw(mask[0], grid, dtype=torch.float32)
    out = ops.upsample_flow(mb, fb, grid)
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4), (out - ref).abs().max()
    # convf1
    w = torch.randn((128, 2, 7, 7), device=DEV, generator=g) * 0.1
    b = torch.randn((128,), device=DEV, generator=g)
    wt = w.permute(1, 2, 3, 0).reshape(98, 128).contiguous()
    ob = grid.zeros(128)
    ops.convf1(fb, wt, b, grid, ob)
    torch.cuda.synchronize()
    ref2 = torch.relu(F.conv2d(flow, w, b, padding=3))[0]
    got2 = nchw_from_rows(ob, grid, 128)
    assert torch.allclose(got2, ref2, atol=3e-2, rtol=1e-2), (got2 - ref2).abs().max()
    # coords init / update
    c1 = grid.zeros(2, dtype=torch.float32)
    fl = grid.zeros(2, dtype=torch.float32)
    finit = torch.randn((2, grid.H, grid.W), device=DEV, generator=g)
    ops.init_coords(c1, finit, grid)
    delta = grid.zeros(32, dtype=torch.float32)
    delta[:, :2] = 0.25
    ops.flow_update(c1, fl, delta, grid)
    torch.cuda.synchronize()
    got = nchw_from_rows(fl, grid, 2)
    assert torch.allclose(got, finit + 0.25, atol=1e-5)


def _rand_qk(grid, Cc, g, scale=1.0):
    q = bf16r(torch.randn((Cc, grid.H, grid.W), device=DEV, generator=g) * scale)
    k = bf16r(torch.randn((Cc, grid.H, grid.W), device=DEV, generator=g) * scale)
    return q, k


def _scores_ref(q, k, M, table, w_pos, clipv):
    Cc, H, W = q.shape
    d = Cc // M
    qq = q.reshape(M, d, H * W).transpose(1, 2)
    kk = k.reshape(M, d, H * W).transpose(1, 2)
    s = qq @ kk.transpose(1, 2) / math.sqrt(d)
    gmax = s.max()
    s = s.clamp(-clipv, clipv)
    if table is not None:
        s = s + w_pos * R.sliding_pos_bias(table, H, W)[None]
    return s, gmax


@pytest.mark.parametrize("H,W,M,d,clipv,w_agg", [(16, 24, 4, 64, float("inf"), 0.13), (13, 22, 4, 64, 2.0, 0.13),
                                                 (8, 16, 1, 256, float("inf"), 0.13), (16, 16, 2, 64, float("inf"), 0.13),
                                                 # a NEGATIVE soft-aggregation weight (feat2score.weight is signed): the
                                                 # specialised epilogue then takes the mode MINIMUM as its softmax pivot
                                                 (16, 24, 4, 64, float("inf"), -0.21), (13, 22, 2, 64, 2.0, -0.21)])
def test_corr_build(H, W, M, d, clipv, w_agg):
    grid = TokenGrid(H, W)
    g = torch.Generator(device=DEV).manual_seed(11)
    q, k = _rand_qk(grid, M * d, g, 0.6)
    table = torch.randn((15, 15), device=DEV, generator=g) if M > 1 else None
    w_pos = 0.5
    s, gmax = _scores_ref(q, k, M, table, w_pos, clipv)
    if M > 1:
        # soft aggregation is shift-invariant to the shared bias, so this matches the reference order
        p = torch.softmax(s * w_agg, dim=0)
        raw = (s * p).sum(0)
    else:
        raw = s[0]
    vol = raw.reshape(1, grid.U, H, W)
    pyr = R.corr_pyramid(vol)
    Q, K = rows_from_nchw(q, grid), rows_from_nchw(k, grid)
    shapes = grid.level_shapes()
    levels = [torch.full((grid.Mp, h * w), float("nan"), device=DEV) for (h, w) in shapes]
    stat_sum = torch.zeros(2, dtype=torch.float64, device=DEV)
    stat_max = torch.full((1,), -float("inf"), device=DEV)
    clip = torch.tensor([clipv], device=DEV)
    nby, nbx = (H + 7) // 8, (W + 7) // 8
    l0h = torch.full((grid.Mp * nby * nbx * 68,), float("nan"), dtype=torch.float16, device=DEV)
    ops.corr_build(Q, K, grid, M=M, d=d, w_agg=w_agg, w_pos=w_pos, pos_table=table, R=7, clip=clip,
                   stat_sum=stat_sum, stat_max=stat_max, levels=levels, level0_h16=l0h)
    mr = torch.zeros(2, device=DEV)
    ops.corr_stats_finalize(stat_sum, grid.U * grid.U, mr)
    torch.cuda.synchronize()
    assert abs(stat_max.item() - gmax.item()) < 1e-3
    n = grid.U * grid.U
    assert abs(stat_sum[0].item() / n - raw.mean().item()) < 1e-4
    assert abs(mr[0].item() - raw.mean().item()) < 1e-4
    rstd_ref = 1.0 / math.sqrt(raw.var(unbiased=False).item() + 1e-12)
    assert abs(mr[1].item() - rstd_ref) < 1e-3 * rstd_ref
    for l, ((h, w), lv) in enumerate(zip(shapes, levels)):
        got = lv.reshape(grid.H, grid.Wp, h * w)[:, :W].reshape(grid.U, h, w)
        ref = pyr[l].reshape(grid.U, h, w)
        assert torch.allclose(got, ref, atol=2e-3, rtol=1e-3), (l, (got - ref).abs().max())
    # level 0 again, as the 16-bit block-ordered copy the default lookup reads (fp16 deltas against fp32 half-block
    # means): within fp16 rounding of the LOCAL variation of the fp32 level 0
    got0 = _level0_decode(l0h, grid).reshape(grid.H, grid.Wp, H, W)[:, :W].reshape(grid.U, H, W)
    f32_0 = levels[0].reshape(grid.H, grid.Wp, H, W)[:, :W].reshape(grid.U, H, W)
    assert torch.allclose(got0, f32_0, atol=3e-3 * float(f32_0.std()), rtol=0), (got0 - f32_0).abs().max()


@pytest.mark.parametrize("H,W,M,d,F_", [(16, 24, 4, 32, 128), (13, 22, 4, 64, 256), (12, 20, 1, 128, 128)])
def test_attn_lse_pv_finalize(H, W, M, d, F_):
    grid = TokenGrid(H, W)
    g = torch.Generator(device=DEV).manual_seed(21)
    q, k = _rand_qk(grid, M * d, g, 0.7)
    table = torch.randn((15, 15), device=DEV, generator=g) if M > 1 else None
    w_pos = 1.0
    clipv = float("inf")
    s, gmax = _scores_ref(q, k, M, table, w_pos, clipv)
    P = torch.softmax(s, dim=-1)                                    # [M,U,U]
    v = bf16r(torch.randn((M, grid.U, F_), device=DEV, generator=g))
    O_ref = P @ v                                                   # [M,U,F]
    Q, K = rows_from_nchw(q, grid), rows_from_nchw(k, grid)
    # V^T in the kernel's key-block order: column = block*BK + (y%8)*BW + x%BW
    BK = ops.pv_block_keys(d, F_)
    BW = BK // 8
    nby, nbx = (H + 7) // 8, (W + BW - 1) // BW
    ldv = nby * nbx * BK
    vt_pad = torch.zeros((M * F_, nby * 8, nbx * BW), device=DEV)
    vt_pad[:, :H, :W] = v.permute(0, 2, 1).reshape(M * F_, H, W)
    Vt = vt_pad.reshape(M * F_, nby, 8, nbx, BW).permute(0, 1, 3, 2, 4).reshape(M * F_, ldv).to(torch.bfloat16).contiguous()
    ks = ops.scores_auto_ksplit(grid)
    lse_part = torch.zeros((ks, M, grid.Mp, 2), device=DEV)
    lse2 = torch.zeros((M, grid.Mp), device=DEV)
    stat_max = torch.full((1,), -float("inf"), device=DEV)
    clip = torch.tensor([clipv], device=DEV)
    ops.attn_lse(Q, K, grid, M=M, d=d, w_pos=w_pos, pos_table=table, R=7, clip=clip, stat_max=stat_max,
                 lse_part=lse_part, lse2=lse2, ksplit=ks)
    torch.cuda.synchronize()
    lse_ref = torch.logsumexp(s, dim=-1) * math.log2(math.e)        # [M,U]
    got = lse2.reshape(M, grid.H, grid.Wp)[:, :, :W].reshape(M, grid.U)
    assert torch.allclose(got, lse_ref, atol=2e-3, rtol=1e-4), (got - lse_ref).abs().max()
    assert abs(stat_max.item() - gmax.item()) < 1e-3
    kp = ops.pv_auto_ksplit(grid, M)      # partial-sum slots the persistent schedule needs
    out = torch.full((kp, M, grid.Mp, F_), float("nan"), device=DEV)
    ops.attn_pv(Q, K, Vt, grid, M=M, d=d, F=F_, w_pos=w_pos, pos_table=table, R=7, clip=clip, lse2=lse2,
                out=out, ksplit=kp)
    torch.cuda.synchronize()
    # partial sums are stored as [slot][M][F/8][Mp][8]
    O = out.reshape(kp, M, F_ // 8, grid.Mp, 8).permute(0, 1, 3, 2, 4).reshape(kp, M, grid.Mp, F_).sum(0)
    O = O.reshape(M, grid.H, grid.Wp, F_)[:, :, :W].reshape(M, grid.U, F_)
    assert torch.allclose(O, O_ref, atol=2e-2, rtol=2e-2), (O - O_ref).abs().max()
    # finalize (setrans flavour) against the restatement fed with the kernel's own O
    if M > 1:
        w_sc = torch.randn((1, F_), device=DEV, generator=g) * 0.1
        b_sc = torch.randn((1,), device=DEV, generator=g)
        coeff = torch.tensor([0.8], device=DEV)
        x = bf16r(torch.randn((grid.U, F_), device=DEV, generator=g))
        xb = torch.zeros((grid.H, grid.Wp, F_), device=DEV)
        xb[:, :W] = x.reshape(H, W, F_)
        xb = xb.reshape(grid.Mp, F_).to(torch.bfloat16)
        yf = grid.zeros(F_, dtype=torch.float32)
        ops.modes_finalize(out, kp, M, F_, grid, w_score=w_sc, b_score=b_sc, coeff=coeff, x_b=xb, out_f=yf)
        torch.cuda.synchronize()
        agg = R.soft_aggregate_feat(O[None], w_sc, b_sc)[0]
        ref = F.layer_norm(0.8 * x + agg, (F_,), eps=1e-12)
        gy = yf.reshape(grid.H, grid.Wp, F_)[:, :W].reshape(grid.U, F_)
        assert torch.allclose(gy, ref, atol=1e-3, rtol=1e-3), (gy - ref).abs().max()
        # production pairing: unused slots stay garbage (NaN here) and the finalize kernel skips them
        out2 = torch.full((kp, M, grid.Mp, F_), float("nan"), device=DEV)
        ops.attn_pv(Q, K, Vt, grid, M=M, d=d, F=F_, w_pos=w_pos, pos_table=table, R=7, clip=clip, lse2=lse2,
                    out=out2, ksplit=kp, zero_fill=False)
        yf2 = grid.zeros(F_, dtype=torch.float32)
        ops.modes_finalize(out2, kp, M, F_, grid, w_score=w_sc, b_score=b_sc, coeff=coeff, x_b=xb, out_f=yf2, pv_bk=BK)
        torch.cuda.synchronize()
        assert torch.equal(yf2, yf)
        # parameters that are 4-byte-aligned views into a flat buffer (nn.DataParallel replicas: broadcast_coalesced)
        flat = torch.zeros(1 + F_ + 1 + 1, device=DEV)
        flat[1:1 + F_] = w_sc[0]
        flat[1 + F_:2 + F_] = b_sc.reshape(-1)
        flat[2 + F_:] = coeff.reshape(-1)
        yf3 = grid.zeros(F_, dtype=torch.float32)
        ops.modes_finalize(out2, kp, M, F_, grid, w_score=flat[1:1 + F_].view(1, F_), b_score=flat[1 + F_:2 + F_],
                           coeff=flat[2 + F_:], x_b=xb, out_f=yf3, pv_bk=BK)
        torch.cuda.synchronize()
        assert torch.equal(yf3, yf)


@pytest.mark.parametrize("Cc", [64, 96, 128, 256])
def test_encoder_norm_kernels(Cc):
    g = torch.Generator(device=DEV).manual_seed(Cc)
    x = (torch.randn((2, Cc, 37, 50), device=DEV, generator=g) * 2 + 0.3).contiguous(memory_format=torch.channels_last)
    r = torch.randn((2, Cc, 37, 50), device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    ab = ops.instnorm_stats(x.permute(0, 2, 3, 1))
    out = torch.empty_like(x)
    ops.nhwc_affine(x.permute(0, 2, 3, 1), ab, r.permute(0, 2, 3, 1), None, True, True, out=out.permute(0, 2, 3, 1))
    torch.cuda.synchronize()
    ref = torch.relu(r + torch.relu(F.instance_norm(x, eps=1e-5)))
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), (out - ref).abs().max()
    # shared (batch-norm style) scale/shift on both branches
    sab = torch.randn((1, Cc, 2), device=DEV, generator=g)
    out2 = torch.empty_like(x)
    ops.nhwc_affine(x.permute(0, 2, 3, 1), sab, r.permute(0, 2, 3, 1), sab, False, False, out=out2.permute(0, 2, 3, 1))
    torch.cuda.synchronize()
    a, b = sab[0, :, 0].view(1, Cc, 1, 1), sab[0, :, 1].view(1, Cc, 1, 1)
    assert torch.allclose(out2, (a * x + b) + (a * r + b), atol=1e-5, rtol=1e-5)
    # statistics are reduced in a fixed order: bit-identical run to run
    assert torch.equal(ab, ops.instnorm_stats(x.permute(0, 2, 3, 1)))
    # fp16 activations (fp32 statistics / arithmetic); Cc must be a multiple of 8
    xh, rh = x.half(), r.half()
    abh = ops.instnorm_stats(xh.permute(0, 2, 3, 1))
    outh = torch.empty_like(xh)
    ops.nhwc_affine(xh.permute(0, 2, 3, 1), abh, rh.permute(0, 2, 3, 1), None, True, True, out=outh.permute(0, 2, 3, 1))
    torch.cuda.synchronize()
    refh = torch.relu(rh.float() + torch.relu(F.instance_norm(xh.float(), eps=1e-5)))
    assert outh.dtype == torch.float16 and torch.allclose(outh.float(), refh, atol=4e-3, rtol=2e-3), (outh.float() - refh).abs().max()


@pytest.mark.parametrize("N,Cc,H,W", [(2, 64, 37, 50), (2, 96, 28, 64), (1, 128, 56, 128), (3, 64, 5, 7), (2, 256, 17, 9),
                                      (2, 64, 224, 512), (1, 64, 400, 640)])
def test_instnorm_one_launch(N, Cc, H, W):
    """craft_nhwc_instnorm_apply (cooperative: slab in shared memory across a grid barrier) against
    F.instance_norm (core/extractor.py:55-64 with norm_fn='instance') and against the three-launch kernels.
    (2,64,224,512) is the 448x1024 layer-1 tensor (198 KB slab per CTA); (1,64,400,640) overflows shared memory."""
    g = torch.Generator(device=DEV).manual_seed(Cc + H)
    cl = torch.channels_last
    x = (torch.randn((N, Cc, H, W), device=DEV, generator=g) * 2 + 0.3).contiguous(memory_format=cl)
    r = torch.randn((N, Cc, H, W), device=DEV, generator=g).contiguous(memory_format=cl)
    nhwc = lambda t: t.permute(0, 2, 3, 1)
    for dt, atol, rtol in ((torch.float32, 2e-4, 1e-4), (torch.float16, 4e-3, 2e-3)):
        if dt == torch.float32 and Cc > 128:
            continue
        xx, rr = x.to(dt), r.to(dt)
        ref_n = F.instance_norm(xx.float(), eps=1e-5)
        # norm + relu
        out = torch.empty_like(xx)
        _, ab = ops.instnorm_apply(nhwc(xx), relu_in=True, out=nhwc(out), return_ab=True)
        torch.cuda.synchronize()
        assert torch.allclose(out.float(), torch.relu(ref_n), atol=atol, rtol=rtol), (out.float() - torch.relu(ref_n)).abs().max()
        ab3 = ops.instnorm_stats(nhwc(xx))
        assert torch.allclose(ab, ab3, atol=1e-5, rtol=1e-4)
        # norm + relu + residual + relu, and an affine on the residual branch
        out2 = ops.instnorm_apply(nhwc(xx), res=nhwc(rr), relu_in=True, relu_out=True)
        ref2 = torch.relu(rr.float() + torch.relu(ref_n))
        assert torch.allclose(nhwc(ref2), out2.float(), atol=atol, rtol=rtol), (nhwc(ref2) - out2.float()).abs().max()
        sab = torch.randn((1, Cc, 2), device=DEV, generator=g)
        out3 = ops.instnorm_apply(nhwc(xx), res=nhwc(rr), rab=sab, relu_in=False, relu_out=False)
        a, b = sab[0, :, 0].view(1, Cc, 1, 1), sab[0, :, 1].view(1, Cc, 1, 1)
        ref3 = (a * rr.float() + b) + ref_n
        assert torch.allclose(nhwc(ref3), out3.float(), atol=2 * atol, rtol=2 * rtol)
        # fixed reduction order: bit-identical run to run
        out4 = ops.instnorm_apply(nhwc(xx), res=nhwc(rr), relu_in=True, relu_out=True)
        assert torch.equal(out2, out4)


def _to_pad(x_nhwc):
    """dense [N,H,W,C] -> ops.PadAct with zero halo cells / gap rows (test helper; the model uses nhwc_affine_pad)."""
    N, H, W, C_ = x_nhwc.shape
    t = torch.zeros((N, H + 1, W + 2, C_), dtype=x_nhwc.dtype, device=x_nhwc.device)
    t[:, :H, :W] = x_nhwc
    return ops.PadAct(t.reshape(-1, C_), N, H, W)


@pytest.mark.parametrize("N,H,W", [(2, 37, 50), (1, 16, 16), (3, 9, 130), (2, 224, 512), (1, 192, 624)])
def test_conv3x3_c64_persistent(N, H, W):
    """craft_conv3x3_c64 (persistent tcgen05 implicit GEMM on the padded-flat layout; core/extractor.py:24-26 layer1
    convolutions) against F.conv2d on the same fp16 operands; halo cells stay zero; bias + ReLU epilogue; the
    InstanceNorm statistics of its epilogue against the statistics kernels on the dense result."""
    g = torch.Generator(device=DEV).manual_seed(N * 1000 + H)
    x = torch.randn((N, H, W, 64), device=DEV, generator=g).half()
    w = (torch.randn((64, 64, 3, 3), device=DEV, generator=g) * 0.06)
    b = torch.randn((64,), device=DEV, generator=g)
    wp = ops.pack_conv64_weight(w)
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w.half().float(), None, padding=1).permute(0, 2, 3, 1)   # [N,H,W,64] fp32
    xp = _to_pad(x)
    y, ab = ops.conv3x3_c64(xp, wp, stats_eps=1e-5)
    torch.cuda.synchronize()
    got = y.dense().float()
    assert torch.allclose(got, ref, atol=2e-2, rtol=1e-2), (got - ref).abs().max()
    full = y.t.view(N, H + 1, W + 2, 64)
    assert float(full[:, H].abs().max()) == 0.0 and float(full[:, :, W:].abs().max()) == 0.0      # padding stays zero
    ab_ref = ops.instnorm_stats(ref.contiguous())                   # statistics of the fp32 result
    assert torch.allclose(ab, ab_ref, atol=2e-3, rtol=2e-3), (ab - ab_ref).abs().max()
    # bias + ReLU (folded eval BatchNorm), no statistics; bit-reproducible
    y2 = ops.conv3x3_c64(xp, wp, bias=b, relu=True)
    ref2 = torch.relu(ref + b.view(1, 1, 1, 64))
    assert torch.allclose(y2.dense().float(), ref2, atol=2e-2, rtol=1e-2)
    full2 = y2.t.view(N, H + 1, W + 2, 64)
    assert float(full2[:, H].abs().max()) == 0.0 and float(full2[:, :, W:].abs().max()) == 0.0
    y3, ab3 = ops.conv3x3_c64(xp, wp, stats_eps=1e-5)
    assert torch.equal(y3.t, y.t) and torch.equal(ab3, ab)
    # layout-converting affine: dense -> padded -> (residual, relu) -> dense
    sab = torch.randn((N, 64, 2), device=DEV, generator=g)
    a_, b_ = sab[:, :, 0].view(N, 1, 1, 64), sab[:, :, 1].view(N, 1, 1, 64)
    p1 = ops.nhwc_affine_pad(x, sab, relu_in=True, out_pad=True)
    assert torch.allclose(p1.dense().float(), torch.relu(a_ * x.float() + b_), atol=4e-3, rtol=2e-3)
    f1 = p1.t.view(N, H + 1, W + 2, 64)
    assert float(f1[:, H].abs().max()) == 0.0 and float(f1[:, :, W:].abs().max()) == 0.0
    d1 = ops.nhwc_affine_pad(y, ab, res=p1, relu_in=True, relu_out=True, out_pad=False)
    refd = torch.relu(p1.dense().float() + torch.relu(ab[:, :, 0].view(N, 1, 1, 64) * got + ab[:, :, 1].view(N, 1, 1, 64)))
    assert d1.shape == (N, H, W, 64) and torch.allclose(d1.float(), refd, atol=8e-3, rtol=4e-3)


@pytest.mark.parametrize("kind", ["instance", "batch"])
def test_fused_encoder_matches_module_path(kind):
    from craft_b200.extractor import BasicEncoder
    torch.manual_seed(3)
    enc = BasicEncoder(output_dim=256, norm_fn=kind).to(DEV).eval()
    if kind == "batch":      # non-trivial running statistics
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    x = torch.rand((2, 3, 128, 192), device=DEV) * 2 - 1
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        enc.use_fused = False
        ref = enc(x)
        enc.use_fused = True
        got = enc(x)
        enc.fused_half = True
        got_h = enc(x)
        enc.fused_half = False
    torch.cuda.synchronize()
    assert torch.allclose(got, ref, atol=2e-3, rtol=1e-3), (got - ref).abs().max()
    # half-precision activations: features are O(1..20); bound the error relative to their scale
    assert got_h.dtype == torch.float32
    assert (got_h - ref).abs().max() <= 2e-2 * max(1.0, ref.abs().max().item()), (got_h - ref).abs().max()
    assert (got_h - ref).abs().mean() <= 4e-3 * max(1.0, ref.abs().mean().item()), (got_h - ref).abs().mean()


@pytest.mark.parametrize("kind", ["instance", "batch"])
def test_s2d_first_convolution_and_direct_token_packing(kind):
    """The inference fast path of the encoders: raw 0..255 frames -> image_s2d (normalise + 2x2 space-to-depth +
    zero border) -> the 7x7/2 convolution as a 4x4/1 one -> ... -> channels-last features handed to the token
    packer.  Checked against the plain module path on the normalised frames (core/network.py:170-171,
    core/extractor.py:173-196) and the NCHW packer."""
    from craft_b200.extractor import BasicEncoder
    torch.manual_seed(5)
    enc = BasicEncoder(output_dim=256, norm_fn=kind).to(DEV).eval()
    img = torch.randint(0, 256, (2, 3, 96, 160), device=DEV).float()
    xn = 2 * (img / 255.0) - 1.0
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        # the input transform alone, against pixel_unshuffle of the normalised frames
        s = ops.image_s2d(img, dtype=torch.float32)
        assert tuple(s.shape) == (2, 96 // 2 + 3, 160 // 2 + 3, 16)
        inner = s[:, 2:-1, 2:-1, :12].reshape(2, 48, 80, 2, 2, 3)            # [n, Y, X, py, px, c]
        back = inner.permute(0, 5, 1, 3, 2, 4).reshape(2, 3, 96, 160)
        assert torch.allclose(back, xn, atol=2.5e-7, rtol=0)      # (torch's CUDA x/255 multiplies by the rounded reciprocal)
        assert s[:, :2].abs().max() == 0 and s[:, -1].abs().max() == 0 and s[:, :, :2].abs().max() == 0
        assert s[:, :, -1].abs().max() == 0 and s[..., 12:].abs().max() == 0
        enc.use_fused = False
        ref = enc(xn)
        enc.use_fused = True
        for half, atol_max, atol_mean in ((False, 2e-3, 2e-4), (True, 2e-2, 4e-3)):
            enc.fused_half = half
            y = enc.forward_nhwc(s2d=ops.image_s2d(img, dtype=enc.fused_dtype()))       # [N, h, w, 256]
            assert y.is_contiguous() and y.dtype == enc.fused_dtype()
            got = y.permute(0, 3, 1, 2).float()
            scale = max(1.0, ref.abs().max().item())
            assert (got - ref).abs().max() <= atol_max * scale, (half, (got - ref).abs().max())
            assert (got - ref).abs().mean() <= atol_mean * max(1.0, ref.abs().mean().item())
            # channels-last features -> token rows, against the NCHW packer on the same values
            grid = TokenGrid(y.shape[1], y.shape[2])
            for mode, c0, Cc in ((ops.PACK_LN, 0, 256), (ops.PACK_TANH, 0, 128), (ops.PACK_RELU_LN, 128, 128), (ops.PACK_RELU, 128, 128)):
                a_b, a_f = grid.zeros(Cc + 64), grid.zeros(Cc, dtype=torch.float32)
                b_b, b_f = grid.zeros(Cc + 64), grid.zeros(Cc, dtype=torch.float32)
                ops.pack_tokens(ops.NhwcFeat(y[0]).window(c0, Cc), grid, mode, out_b=a_b, colb=64, out_f=a_f)
                ops.pack_tokens(got[0, c0:c0 + Cc].contiguous(), grid, mode, out_b=b_b, colb=64, out_f=b_f)
                torch.cuda.synchronize()
                assert torch.allclose(a_f, b_f, atol=2e-5, rtol=1e-5), (mode, (a_f - b_f).abs().max())
                assert torch.allclose(a_b.float(), b_b.float(), atol=2e-2, rtol=1e-2)
                assert a_f.reshape(grid.H, grid.Wp, -1)[:, grid.W:].abs().max() == 0


@pytest.mark.parametrize("H,W,M,d,clipv", [(16, 24, 4, 64, float("inf")), (17, 22, 4, 64, 2.0), (16, 16, 1, 256, float("inf"))])
@pytest.mark.parametrize("spread", [5.0, 0.4, "mixed", "far"])
def test_corr_lookup0_on_demand(H, W, M, d, clipv, spread):
    """Level-0 lookup recomputed from Q/K rows == lookup on the materialised (oracle) level-0 volume."""
    grid = TokenGrid(H, W)
    g = torch.Generator(device=DEV).manual_seed(31)
    q, k = _rand_qk(grid, M * d, g, 0.6)
    table = torch.randn((15, 15), device=DEV, generator=g) if M > 1 else None
    w_agg, w_pos = 0.13, 0.5
    s, _ = _scores_ref(q, k, M, table, w_pos, clipv)
    raw = (s * torch.softmax(s * w_agg, dim=0)).sum(0) if M > 1 else s[0]
    mean, rstd = raw.mean().item(), 1.0 / math.sqrt(raw.var(unbiased=False).item() + 1e-12)
    vol = ((raw - mean) * rstd).reshape(1, grid.U, H, W)
    # spread 5: windows of a 4x2 query patch scatter (direct-from-L2 path); 0.4: smooth flow, the patch's bounding
    # box is staged in shared memory; "mixed": smooth with a motion boundary; "far": windows partly / wholly off-image
    noise = torch.randn((1, 2, H, W), device=DEV, generator=g)
    if spread == "mixed":
        flow = noise * 0.3 + torch.tensor([2.5, -1.5], device=DEV).view(1, 2, 1, 1)
        flow[:, :, :, W // 2:] += torch.tensor([-9.0, 6.0], device=DEV).view(1, 2, 1, 1)
    elif spread == "far":
        flow = noise * 0.5 + torch.tensor([W - 3.0, -(H - 2.0)], device=DEV).view(1, 2, 1, 1)
        flow[:, :, : H // 2] = noise[:, :, : H // 2] * 0.5 + torch.tensor([-6.0, 3.0], device=DEV).view(1, 2, 1, 1)
    else:
        flow = noise * spread
    coords = R.coords_grid(1, H, W, DEV) + flow
    ref = R.corr_lookup([vol.reshape(grid.U, 1, H, W)], coords)[0]          # level 0 only: [81,H,W]
    Q, K = rows_from_nchw(q, grid), rows_from_nchw(k, grid)
    cbuf = rows_from_nchw(coords[0], grid, dtype=torch.float32)
    out_n = torch.zeros((324, H, W), device=DEV)
    ops.corr_lookup0(Q, K, grid, M=M, d=d, w_agg=w_agg, w_pos=w_pos, pos_table=table, R=7,
                     clip=torch.tensor([clipv], device=DEV), coords=cbuf,
                     mean_rstd=torch.tensor([mean, rstd], device=DEV), out_nchw=out_n)
    torch.cuda.synchronize()
    assert torch.allclose(out_n[:81], ref, atol=3e-3, rtol=1e-3), (out_n[:81] - ref).abs().max()


@pytest.mark.skipif(os.environ.get("CRAFT_B200_TEST_EXPERIMENTAL") != "1",
                    reason="experimental big-box TMA mode of the shift-GEMM: written at the end of round 1 without "
                           "GPU time left to validate it; run with CRAFT_B200_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("mode", ["1", "2"])
def test_gemm_bigbox_mode_in_subprocess(mode):
    """CRAFT_GEMM_BIGBOX is read once per process, so the GEMM tests are re-run in a child process with it
    set; the child also checks that the big-box path was really taken (no silent fallback)."""
    import subprocess
    import sys
    env = dict(os.environ, CRAFT_GEMM_BIGBOX=mode, CRAFT_B200_TEST_EXPERIMENTAL="0")
    code = ("import sys, pytest; rc = pytest.main(['-q', '-x', '-m', 'gpu', '-k', 'test_gemm_plain or test_gemm_conv or "
            "test_gemm_gru_epilogues', %r]); "
            "from craft_b200 import _lib; n = _lib.load().craft_b200_bigbox_gemm_count(); print('bigbox launches', n); "
            "sys.exit(int(rc) or (0 if n > 0 else 3))" % os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
