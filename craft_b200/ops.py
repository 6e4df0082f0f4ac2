"""Host-side operator layer: torch tensors in, C-ABI calls out (craft_b200/_lib.py).

torch is used here only for device memory and the current CUDA stream.  Every function launches
hand-written sm_100a kernels from libcraft_b200.so on `torch.cuda.current_stream()` and raises if
the library or a CUDA device is missing -- there is no fallback.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import DenseAttnArgs, GemmArgs, PvArgs, ScoresArgs

EPI_STORE, EPI_GRU_ZR, EPI_GRU_Q, EPI_MOTION = 0, 1, 2, 3
PACK_COPY, PACK_LN, PACK_TANH, PACK_RELU, PACK_RELU_LN = 0, 1, 2, 3, 4


class TokenGrid:
    """Padded-flat token grid: rows p = y*(W+2)+x, two zero halo cells per grid row."""

    def __init__(self, H, W):
        self.H, self.W = int(H), int(W)
        self.Wp = self.W + 2
        self.Mp = self.H * self.Wp
        self.U = self.H * self.W

    def level_shapes(self, n=4):
        out, h, w = [], self.H, self.W
        for _ in range(n):
            out.append((h, w))
            h, w = h // 2, w // 2
        return out

    def zeros(self, cols, dtype=torch.bfloat16, device="cuda"):
        return torch.zeros((self.Mp, cols), dtype=dtype, device=device)


def _stream():
    # current stream of the CURRENT device: every public entry point (CRAFT.forward, the standalone module
    # forwards) runs under `torch.cuda.device(input.device)`, see on_device() below
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(fn):
    """Decorator for module forward()s: make the device of the first tensor argument current for the call,
    so that streams, workspaces and the library's per-device state all refer to it (a model moved to cuda:1
    must not launch on cuda:0's stream; nn.DataParallel calls replicas from one thread per device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kw):
        dev = None
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.Tensor):
                dev = a.device
                break
            if isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor):
                dev = a[0].device
                break
        if dev is None or dev.type != "cuda":
            raise _lib.CraftB200Error("%s.%s needs CUDA tensors: craft_b200 has no CPU path"
                                      % (type(self).__name__, fn.__name__))
        with torch.cuda.device(dev):
            return fn(self, *args, **kw)
    return wrapped


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda, "craft_b200 ops need CUDA tensors (no CPU path exists)"
    return C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.CraftB200Error("%s: expected a CUDA tensor; craft_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s: expected %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def _ld(t):
    return t.shape[-1] if t is not None else 0


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def pack_tokens(src, grid, mode=PACK_COPY, out_b=None, colb=0, out_f=None, colf=0):
    """src [C,H,W] f32 -> token rows (optionally LayerNorm / tanh / relu)."""
    _chk(src, torch.float32, "src")
    _chk(out_b, torch.bfloat16, "out_b")
    _chk(out_f, torch.float32, "out_f")
    Cc = src.shape[0]
    assert src.shape[1] == grid.H and src.shape[2] == grid.W
    _lib.call("craft_pack_tokens", _ptr(src), Cc, grid.H, grid.W, mode, _ptr(out_b), _ld(out_b), colb,
              _ptr(out_f), _ld(out_f), colf, _stream())


def unpack_tokens(buf, col, Cc, grid, out=None):
    if out is None:
        out = torch.empty((Cc, grid.H, grid.W), dtype=torch.float32, device=buf.device)
    is_b = 1 if buf.dtype == torch.bfloat16 else 0
    if not is_b:
        _chk(buf, torch.float32, "buf")
    _lib.call("craft_unpack_tokens", _ptr(buf), is_b, _ld(buf), col, Cc, grid.H, grid.W, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# shift-GEMM
# ------------------------------------------------------------------------------------------------
def shift_gemm(A, Bw, *, M, Npad, K, BN, taps=(0,), a_koff=0, b_koff=0, grid=None, epilogue=EPI_STORE,
               alpha=1.0, act=0, bias=None, out_b=None, colb=0, out_f=None, colf=0, aux0=None, aux1=None,
               b_block_grid=None, cluster=0, stages=0, a_share=0):
    """b_block_grid: B rows are the tokens of that grid and n-tile j is its j-th 8 x BN/8 spatial block.
    a_share=1 (experimental): load the A rows of a kernel row once for all of its taps."""
    _chk(A, torch.bfloat16, "A")
    _chk(Bw, torch.bfloat16, "B")
    _chk(bias, torch.float32, "bias")
    _chk(out_b, torch.bfloat16, "out_b")
    _chk(out_f, torch.float32, "out_f")
    _chk(aux0, torch.float32, "aux0")
    _chk(aux1, torch.float32, "aux1")
    a = GemmArgs()
    a.A, a.a_rows, a.lda, a.a_koff = A.data_ptr(), A.shape[0], A.shape[1], a_koff
    a.B, a.b_rows, a.ldb_, a.b_koff = Bw.data_ptr(), Bw.shape[0], Bw.shape[1], b_koff
    a.M, a.Npad, a.K, a.T, a.BN = M, Npad, K, len(taps), BN
    a.cluster = cluster
    a.stages = stages
    a.a_share = a_share
    if b_block_grid is not None:
        a.b_blocked, a.b_H, a.b_W = 1, b_block_grid.H, b_block_grid.W
    for i, t in enumerate(taps):
        a.tap_off[i] = int(t)
    a.H, a.W = (grid.H, grid.W) if grid is not None else (0, 0)
    a.epilogue, a.alpha, a.act = epilogue, float(alpha), act
    a.bias = bias.data_ptr() if bias is not None else None
    a.out_bf16 = out_b.data_ptr() if out_b is not None else None
    a.ldo_b, a.colo_b = _ld(out_b), colb
    a.out_f32 = out_f.data_ptr() if out_f is not None else None
    a.ldo_f, a.colo_f = _ld(out_f), colf
    a.aux0 = aux0.data_ptr() if aux0 is not None else None
    a.aux1 = aux1.data_ptr() if aux1 is not None else None
    if bias is not None:
        assert bias.numel() >= Npad
    _lib.call("craft_shift_gemm", C.byref(a), _stream())


def conv_taps(kh, kw, grid):
    """Row offsets of a kh x kw 'same' convolution on the padded-flat grid, (ky,kx) row-major."""
    assert kw // 2 <= 2, "halo is two cells wide"
    return [(ky - kh // 2) * grid.Wp + (kx - kw // 2) for ky in range(kh) for kx in range(kw)]


def pack_conv_weight(w, Npad=None, cin_perm=None, Kpad=None):
    """[Cout,Cin,kh,kw] f32 -> bf16 [kh*kw*Npad, Kpad]: tap-major blocks of [Npad, Cin] (zero padded)."""
    Cout, Cin, kh, kw = w.shape
    Npad = Npad or Cout
    Kpad = Kpad or ((Cin + 63) // 64) * 64
    ww = w.detach().float()
    if cin_perm is not None:
        ww = ww[:, cin_perm]
    out = torch.zeros((kh * kw, Npad, Kpad), dtype=torch.float32, device=w.device)
    out[:, :Cout, :Cin] = ww.permute(2, 3, 0, 1).reshape(kh * kw, Cout, Cin)
    return out.reshape(kh * kw * Npad, Kpad).to(torch.bfloat16).contiguous()


def pack_linear_weight(w, Npad=None, Kpad=None):
    """[N,K] f32 -> bf16 [Npad, Kpad]."""
    N, K = w.shape
    Npad = Npad or N
    Kpad = Kpad or ((K + 63) // 64) * 64
    out = torch.zeros((Npad, Kpad), dtype=torch.float32, device=w.device)
    out[:N, :K] = w.detach().float()
    return out.to(torch.bfloat16).contiguous()


def pad_bias(b, Npad):
    out = torch.zeros((Npad,), dtype=torch.float32, device=b.device)
    out[: b.numel()] = b.detach().float()
    return out


# ------------------------------------------------------------------------------------------------
# scores (corr build / attention LSE) and P.V
# ------------------------------------------------------------------------------------------------
def scores_auto_ksplit(grid):
    return _lib.load().craft_scores_auto_ksplit(grid.H, grid.W)


def pv_auto_ksplit(grid, M):
    return _lib.load().craft_pv_auto_ksplit(grid.H, grid.W, M)


def pv_block_keys(d, F):
    bk = _lib.load().craft_pv_block_keys(d, F)
    if bk <= 0:
        raise _lib.CraftB200Error("attn_pv: unsupported (d=%d, F=%d)" % (d, F))
    return bk


def blocked_keys(grid, BK):
    """Number of key columns of a V^T matrix in 8 x BK/8 block order."""
    bw = BK // 8
    return ((grid.H + 7) // 8) * ((grid.W + bw - 1) // bw) * BK


def _scores_args(Q, K, grid, M, d, w_pos, pos_table, R, clip, run_flag, ksplit):
    _chk(Q, torch.bfloat16, "Q")
    _chk(K, torch.bfloat16, "K")
    _chk(pos_table, torch.float32, "pos_table")
    _chk(clip, torch.float32, "clip")
    assert Q.shape == (grid.Mp, M * d) and K.shape == (grid.Mp, M * d)
    a = ScoresArgs()
    a.Q, a.K = Q.data_ptr(), K.data_ptr()
    a.C, a.M, a.d = M * d, M, d
    a.H, a.W = grid.H, grid.W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = pos_table.data_ptr() if pos_table is not None else None
    a.R = R
    a.clip = clip.data_ptr()
    a.run_flag = run_flag.data_ptr() if run_flag is not None else None
    a.ksplit = ksplit
    return a


def corr_build(Q, K, grid, *, M, d, w_agg, w_pos, pos_table, R, clip, stat_sum, stat_max, levels,
               run_flag=None, ksplit=0):
    """levels: list of 4 f32 tensors [Mp, h_l*w_l] (levels[0] may be None)."""
    a = _scores_args(Q, K, grid, M, d, w_pos, pos_table, R, clip, run_flag, ksplit)
    a.w_agg = float(w_agg)
    _chk(stat_sum, torch.float64, "stat_sum")
    _chk(stat_max, torch.float32, "stat_max")
    a.stat_sum, a.stat_max = stat_sum.data_ptr(), stat_max.data_ptr()
    for l in range(4):
        _chk(levels[l], torch.float32, "level")
        a.lvl[l] = levels[l].data_ptr() if levels[l] is not None else None
    _lib.call("craft_corr_build", C.byref(a), _stream())


def attn_lse(Q, K, grid, *, M, d, w_pos, pos_table, R, clip, stat_max, lse_part, lse2, run_flag=None, ksplit=0,
             mask_radius=-1):
    a = _scores_args(Q, K, grid, M, d, w_pos, pos_table, R, clip, run_flag, ksplit)
    a.mask_radius = int(mask_radius)
    _chk(stat_max, torch.float32, "stat_max")
    _chk(lse_part, torch.float32, "lse_part")
    _chk(lse2, torch.float32, "lse2")
    a.stat_max = stat_max.data_ptr()
    a.lse_part, a.lse2 = lse_part.data_ptr(), lse2.data_ptr()
    _lib.call("craft_attn_lse", C.byref(a), _stream())


def corr_stats_finalize(stat_sum, n, mean_rstd, flag=None):
    """stat_sum [2,2] f64: row 0 = unclamped pass, row 1 = clamped re-pass (used when *flag != 0)."""
    _chk(stat_sum, torch.float64, "stat_sum")
    _lib.call("craft_corr_stats_finalize", _ptr(stat_sum), _ptr(flag), float(n), _ptr(mean_rstd), _stream())


def clip_gate(stat_max, attn_clip, clip, flag, diag=None):
    _chk(diag, torch.float32, "diag")
    _lib.call("craft_clip_gate", _ptr(stat_max), float(attn_clip), _ptr(clip), _ptr(flag), _ptr(diag), _stream())


def attn_pv(Q, K, Vt, grid, *, M, d, F, w_pos, pos_table, R, clip, lse2, out, ksplit, zero_fill=True, mask_radius=-1):
    """`out` holds `ksplit` partial-sum slots [ksplit, M, Mp, F]; zero_fill=False leaves the slots a unit does
    not use untouched (pair it with modes_finalize(pv_bk=...), which knows the schedule)."""
    _chk(Q, torch.bfloat16, "Q")
    _chk(K, torch.bfloat16, "K")
    _chk(Vt, torch.bfloat16, "Vt")
    _chk(out, torch.float32, "out")
    _chk(lse2, torch.float32, "lse2")
    assert Vt.shape[0] >= M * F and out.numel() >= ksplit * M * grid.Mp * F
    a = PvArgs()
    a.Q, a.K, a.Vt, a.ldv = Q.data_ptr(), K.data_ptr(), Vt.data_ptr(), Vt.shape[1]
    a.C, a.M, a.d, a.F = M * d, M, d, F
    a.H, a.W = grid.H, grid.W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = pos_table.data_ptr() if pos_table is not None else None
    a.R = R
    a.clip, a.lse2, a.out = clip.data_ptr(), lse2.data_ptr(), out.data_ptr()
    a.ksplit = ksplit
    a.zero_fill = 1 if zero_fill else 0
    a.mask_radius = int(mask_radius)
    _lib.call("craft_attn_pv", C.byref(a), _stream())


def modes_finalize(O, nsum, M, F, grid, *, w_score, b_score, coeff, gma=0, x_b=None, colx=0, x_f=None, colxf=0,
                   out_b=None, colb=0, out_f=None, colf=0, pv_bk=0):
    _chk(O, torch.float32, "O")
    _lib.call("craft_modes_finalize", _ptr(O), nsum, M, F, _ptr(w_score), _ptr(b_score), _ptr(coeff), gma,
              _ptr(x_b), _ld(x_b), colx, _ptr(x_f), _ld(x_f), colxf, grid.H, grid.W,
              _ptr(out_b), _ld(out_b), colb, _ptr(out_f), _ld(out_f), colf, int(pv_bk), _stream())


def soft_aggregate(x, w, b, basis=None, num_feat=1):
    """LearnedSoftAggregate on a dense f32 tensor with the group axis leading: x [M, n] (num_feat 1) or
    [M, n, F]; returns [n] / [n, F]."""
    _chk(x, torch.float32, "x")
    _chk(basis, torch.float32, "basis")
    _chk(w, torch.float32, "w")
    _chk(b, torch.float32, "b")
    M = x.shape[0]
    if num_feat == 1:
        n, F_ = x[0].numel(), 1
        out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
    else:
        F_ = x.shape[-1]
        n = x[0].numel() // F_
        out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
        assert w.numel() == F_
    _lib.call("craft_soft_aggregate", _ptr(x), _ptr(basis), M, n, F_, _ptr(w), _ptr(b), _ptr(out), _stream())
    return out


def attn_dense(Q, K, grid, *, M, d, w_pos, pos_table, R, clip, lse2=None, mask_radius=-1):
    """Dense [M, U, U] scores (lse2 None) or softmax probabilities over the REAL tokens -- small grids only."""
    _chk(Q, torch.bfloat16, "Q")
    _chk(K, torch.bfloat16, "K")
    _chk(lse2, torch.float32, "lse2")
    if grid.U > 4096:
        raise ValueError("attn_dense materialises [M,U,U] and is meant for small grids (U <= 4096), got U=%d" % grid.U)
    out = torch.empty((M, grid.U, grid.U), dtype=torch.float32, device=Q.device)
    a = DenseAttnArgs()
    a.Q, a.K = Q.data_ptr(), K.data_ptr()
    a.C, a.M, a.d, a.H, a.W = M * d, M, d, grid.H, grid.W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = pos_table.data_ptr() if pos_table is not None else None
    a.R = R
    a.clip = clip.data_ptr()
    a.lse2 = lse2.data_ptr() if lse2 is not None else None
    a.mask_radius = int(mask_radius)
    a.out = out.data_ptr()
    _lib.call("craft_attn_dense", C.byref(a), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# lookup / small kernels
# ------------------------------------------------------------------------------------------------
def corr_lookup(levels, grid, coords, mean_rstd, out_b=None, out_nchw=None, first_level=0):
    arr = (C.c_void_p * 4)(*[(lv.data_ptr() if lv is not None else None) for lv in levels])
    _chk(coords, torch.float32, "coords")
    _chk(out_b, torch.bfloat16, "out_b")
    _chk(out_nchw, torch.float32, "out_nchw")
    _lib.call("craft_corr_lookup", arr, grid.H, grid.W, _ptr(coords), _ptr(mean_rstd), _ptr(out_b), _ld(out_b),
              _ptr(out_nchw), first_level, _stream())


def corr_lookup0(Q, K, grid, *, M, d, w_agg, w_pos, pos_table, R, clip, coords, mean_rstd, out_b=None, out_nchw=None):
    """Level-0 lookup channels computed on demand from projected Q/K rows (no level-0 volume)."""
    _chk(Q, torch.bfloat16, "Q")
    _chk(K, torch.bfloat16, "K")
    _chk(out_b, torch.bfloat16, "out_b")
    _chk(out_nchw, torch.float32, "out_nchw")
    _lib.call("craft_corr_lookup0", _ptr(Q), _ptr(K), M, d, 1.0 / math.sqrt(d), float(w_agg), float(w_pos),
              _ptr(pos_table), R, _ptr(clip), grid.H, grid.W, _ptr(coords), _ptr(mean_rstd), _ptr(out_b), _ld(out_b),
              _ptr(out_nchw), _stream())


def convf1(flow, wt, bias, grid, out_b, colo=0):
    _chk(flow, torch.float32, "flow")
    _chk(wt, torch.float32, "wt")
    _chk(out_b, torch.bfloat16, "out_b")
    _lib.call("craft_convf1", _ptr(flow), _ptr(wt), _ptr(bias), grid.H, grid.W, _ptr(out_b), _ld(out_b), colo, _stream())


def flow_update(coords1, flow, delta, grid):
    _lib.call("craft_flow_update", _ptr(coords1), _ptr(flow), _ptr(delta), _ld(delta), grid.H, grid.W, _stream())


def init_coords(coords1, flow_init, grid):
    _chk(flow_init, torch.float32, "flow_init")
    _lib.call("craft_init_coords", _ptr(coords1), _ptr(flow_init), grid.H, grid.W, _stream())


def upsample_flow(mask, flow, grid, out=None):
    if out is None:
        out = torch.empty((2, 8 * grid.H, 8 * grid.W), dtype=torch.float32, device=flow.device)
    _chk(flow, torch.float32, "flow")
    is_b = 1 if mask.dtype == torch.bfloat16 else 0
    _lib.call("craft_upsample_flow", _ptr(mask), is_b, _ld(mask), _ptr(flow), grid.H, grid.W, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# encoder glue (channels-last f32 or f16 tensors)
# ------------------------------------------------------------------------------------------------
def _act_dtype(t, name):
    if t.dtype not in (torch.float32, torch.float16):
        raise TypeError("%s: expected float32 or float16, got %s" % (name, t.dtype))
    _chk(t, t.dtype, name)
    return 1 if t.dtype == torch.float16 else 0


def instnorm_stats(x_nhwc, eps=1e-5):
    """x: [N,H,W,C] contiguous f32/f16 -> ab [N,C,2] f32 = (rstd, -mean*rstd).  Deterministic (no atomics)."""
    half = _act_dtype(x_nhwc, "x")
    N, H, W, Cc = x_nhwc.shape
    part = torch.empty((1024 * N * Cc * 2,), dtype=torch.float32, device=x_nhwc.device)   # per-block partial sums
    ab = torch.empty((N, Cc, 2), dtype=torch.float32, device=x_nhwc.device)
    _lib.call("craft_nhwc_instnorm_stats", _ptr(x_nhwc), half, N, H * W, Cc, float(eps), _ptr(part), part.numel(),
              _ptr(ab), _stream())
    return ab


def nhwc_affine(v, ab=None, res=None, rab=None, relu_in=False, relu_out=False, out=None):
    """out = relu_out([ra*res+rb] + relu_in(a*v+b)); v/res/out [N,H,W,C] f32 or f16; ab/rab f32 [N or 1, C, 2]."""
    half = _act_dtype(v, "v")
    _chk(res, v.dtype, "res")
    N, H, W, Cc = v.shape
    if out is None:
        out = torch.empty_like(v)
    _chk(out, v.dtype, "out")
    st = lambda t: 0 if (t is None or t.shape[0] == 1) else 2 * Cc
    _lib.call("craft_nhwc_affine", _ptr(v), half, _ptr(ab), st(ab), _ptr(res), _ptr(rab), st(rab), int(relu_in),
              int(relu_out), N, H * W, Cc, _ptr(out), _stream())
    return out
