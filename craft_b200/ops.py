"""Host-side operator layer: convenience wrappers (TokenGrid objects, keyword arguments) over the
registered torch custom ops `torch.ops.craft_b200.*` (craft_b200/torch_ops.py), which bind the C ABI
of libcraft_b200.so (include/craft_b200.h) one to one.

torch is used only for device memory, the dispatcher and the current CUDA stream.  Every function
launches hand-written sm_100a kernels on `torch.cuda.current_stream()` and raises if the library or
a CUDA device is missing -- there is no fallback (the ops have no CPU implementation).
"""
import ctypes as C
import math

import torch

from . import _lib
from .torch_ops import OPS

import contextlib
import contextvars

EPI_STORE, EPI_GRU_ZR, EPI_GRU_Q, EPI_MOTION, EPI_FLOW = 0, 1, 2, 3, 4

# Tensor-core operand / activation dtype of the current call: bfloat16 (default, libcraft_b200.so) or float16
# (libcraft_b200_fp16.so, the fp32-parity tier).  CRAFT.forward sets it from the model's `precision`; every
# buffer, packed weight and workspace created below follows it, and torch_ops routes each launch to the
# matching build of the library by the dtype of its operands.
_ACT_DTYPE = contextvars.ContextVar("craft_b200_act_dtype", default=torch.bfloat16)


def act_dtype():
    return _ACT_DTYPE.get()


@contextlib.contextmanager
def precision(dtype):
    if dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("craft_b200 operand dtype must be torch.bfloat16 or torch.float16")
    tok = _ACT_DTYPE.set(dtype)
    try:
        yield
    finally:
        _ACT_DTYPE.reset(tok)
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

    def zeros(self, cols, dtype=None, device="cuda"):
        return torch.zeros((self.Mp, cols), dtype=dtype or act_dtype(), device=device)


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
def _cuda_only(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CraftB200Error("expected CUDA tensors; craft_b200 has no CPU path")


class NhwcFeat:
    """A channels-last feature map [H, W, ldc] (f16 or f32, contiguous) with a channel window [c0, c0 + C): what the
    fused encoders hand to the token packer, so that the NHWC -> NCHW fp32 -> token round trip
    (core/extractor.py output -> core/setrans.py:791-795 / core/network.py:209-211) never happens."""

    def __init__(self, t, c0=0, C=None):
        assert t.dim() == 3 and t.is_contiguous()
        self.t, self.c0, self.C = t, c0, (t.shape[2] - c0 if C is None else C)

    def window(self, c0, C):
        assert c0 + C <= self.C
        return NhwcFeat(self.t, self.c0 + c0, C)

    def nchw(self):
        """[C, H, W] fp32 (tests / debugging)."""
        return self.t[:, :, self.c0:self.c0 + self.C].permute(2, 0, 1).float().contiguous()


def pack_tokens(src, grid, mode=PACK_COPY, out_b=None, colb=0, out_f=None, colf=0):
    """src [C,H,W] f32 (or an NhwcFeat) -> token rows (optionally LayerNorm / tanh / relu)."""
    if isinstance(src, NhwcFeat):
        return pack_tokens_nhwc(src.t, grid, mode, c0=src.c0, C=src.C, out_b=out_b, colb=colb, out_f=out_f, colf=colf)
    _cuda_only(src, out_b, out_f)
    OPS.pack_tokens(src, grid.H, grid.W, mode, out_b, colb, out_f, colf)


def pack_tokens_nhwc(src, grid, mode=PACK_COPY, c0=0, C=None, out_b=None, colb=0, out_f=None, colf=0):
    """src [H,W,ldc] channels-last f16/f32 -> token rows of channels [c0, c0+C) (LayerNorm / tanh / relu as pack_tokens)."""
    _cuda_only(src, out_b, out_f)
    OPS.pack_tokens_nhwc(src, c0, C or src.shape[2], grid.H, grid.W, mode, out_b, colb, out_f, colf)


def unpack_tokens(buf, col, Cc, grid, out=None):
    if out is None:
        out = torch.empty((Cc, grid.H, grid.W), dtype=torch.float32, device=buf.device)
    _cuda_only(buf, out)
    OPS.unpack_tokens(buf, col, Cc, grid.H, grid.W, out)
    return out


# ------------------------------------------------------------------------------------------------
# shift-GEMM
# ------------------------------------------------------------------------------------------------
def shift_gemm(A, Bw, *, M, Npad, K, BN, taps=(0,), a_koff=0, b_koff=0, grid=None, epilogue=EPI_STORE,
               alpha=1.0, act=0, bias=None, out_b=None, colb=0, out_f=None, colf=0, aux0=None, aux1=None,
               b_block_grid=None, cluster=0, stages=0, a_share=0):
    """b_block_grid: B rows are the tokens of that grid and n-tile j is its j-th 8 x BN/8 spatial block.
    a_share=1 (experimental): load the A rows of a kernel row once for all of its taps."""
    _cuda_only(A, Bw)
    H, W = (grid.H, grid.W) if grid is not None else (0, 0)
    bH, bW = (b_block_grid.H, b_block_grid.W) if b_block_grid is not None else (0, 0)
    OPS.shift_gemm(A, Bw, M, Npad, K, BN, [int(t) for t in taps], a_koff, b_koff, H, W, epilogue, float(alpha), act, bias,
                   out_b, colb, out_f, colf, aux0, aux1, bH, bW, cluster, stages, a_share)


def conv_taps(kh, kw, grid):
    """Row offsets of a kh x kw 'same' convolution on the padded-flat grid, (ky,kx) row-major."""
    assert kw // 2 <= 2, "halo is two cells wide"
    return [(ky - kh // 2) * grid.Wp + (kx - kw // 2) for ky in range(kh) for kx in range(kw)]


def pack_conv_weight(w, Npad=None, cin_perm=None, Kpad=None):
    """[Cout,Cin,kh,kw] f32 -> operand dtype [kh*kw*Npad, Kpad]: tap-major blocks of [Npad, Cin] (zero padded)."""
    Cout, Cin, kh, kw = w.shape
    Npad = Npad or Cout
    Kpad = Kpad or ((Cin + 63) // 64) * 64
    ww = w.detach().float()
    if cin_perm is not None:
        ww = ww[:, cin_perm]
    out = torch.zeros((kh * kw, Npad, Kpad), dtype=torch.float32, device=w.device)
    out[:, :Cout, :Cin] = ww.permute(2, 3, 0, 1).reshape(kh * kw, Cout, Cin)
    return out.reshape(kh * kw * Npad, Kpad).to(act_dtype()).contiguous()


def pack_linear_weight(w, Npad=None, Kpad=None):
    """[N,K] f32 -> operand dtype [Npad, Kpad]."""
    N, K = w.shape
    Npad = Npad or N
    Kpad = Kpad or ((K + 63) // 64) * 64
    out = torch.zeros((Npad, Kpad), dtype=torch.float32, device=w.device)
    out[:N, :K] = w.detach().float()
    return out.to(act_dtype()).contiguous()


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


def corr_build(Q, K, grid, *, M, d, w_agg, w_pos, pos_table, R, clip, stat_sum, stat_max, levels,
               run_flag=None, ksplit=0, level0_h16=None):
    """levels: list of 4 f32 tensors [Mp, h_l*w_l] (levels[0] may be None); level0_h16: optional fp16
    [Mp, nblocks*64] level 0 in 8x8 key-block order (see include/craft_b200.h)."""
    _cuda_only(Q, K)
    OPS.corr_build(Q, K, grid.H, grid.W, M, d, float(w_agg), float(w_pos), pos_table, R, clip, stat_sum, stat_max,
                   levels[0], levels[1], levels[2], levels[3], run_flag, ksplit, level0_h16)


def attn_lse(Q, K, grid, *, M, d, w_pos, pos_table, R, clip, stat_max, lse_part, lse2, run_flag=None, ksplit=0,
             mask_radius=-1):
    _cuda_only(Q, K)
    OPS.attn_lse(Q, K, grid.H, grid.W, M, d, float(w_pos), pos_table, R, clip, stat_max, lse_part, lse2, run_flag, ksplit,
                 int(mask_radius))


def corr_stats_finalize(stat_sum, n, mean_rstd, flag=None):
    """stat_sum [2,2] f64: row 0 = unclamped pass, row 1 = clamped re-pass (used when *flag != 0)."""
    OPS.corr_stats_finalize(stat_sum, flag, float(n), mean_rstd)


def clip_gate(stat_max, attn_clip, clip, flag, diag=None):
    OPS.clip_gate(stat_max, float(attn_clip), clip, flag, diag)


def attn_pv(Q, K, Vt, grid, *, M, d, F, w_pos, pos_table, R, clip, lse2, out, ksplit, zero_fill=True, mask_radius=-1):
    """`out` holds `ksplit` partial-sum slots [ksplit, M, Mp, F]; zero_fill=False leaves the slots a unit does
    not use untouched (pair it with modes_finalize(pv_bk=...), which knows the schedule)."""
    _cuda_only(Q, K, Vt, out)
    OPS.attn_pv(Q, K, Vt, grid.H, grid.W, M, d, F, float(w_pos), pos_table, R, clip, lse2, out, ksplit, bool(zero_fill),
                int(mask_radius))


def modes_finalize(O, nsum, M, F, grid, *, w_score, b_score, coeff, gma=0, x_b=None, colx=0, x_f=None, colxf=0,
                   out_b=None, colb=0, out_f=None, colf=0, pv_bk=0):
    OPS.modes_finalize(O, nsum, M, F, grid.H, grid.W, w_score, b_score, coeff, gma, x_b, colx, x_f, colxf, out_b, colb,
                       out_f, colf, int(pv_bk))


def soft_aggregate(x, w, b, basis=None, num_feat=1):
    """LearnedSoftAggregate on a dense f32 tensor with the group axis leading: x [M, n] (num_feat 1) or
    [M, n, F]; returns [n] / [n, F]."""
    _cuda_only(x)
    M = x.shape[0]
    if num_feat == 1:
        n, F_ = x[0].numel(), 1
    else:
        F_ = x.shape[-1]
        n = x[0].numel() // F_
        assert w.numel() == F_
    out = torch.empty(x.shape[1:], dtype=torch.float32, device=x.device)
    OPS.soft_aggregate(x, basis, M, n, F_, w, b, out)
    return out


def attn_dense(Q, K, grid, *, M, d, w_pos, pos_table, R, clip, lse2=None, mask_radius=-1):
    """Dense [M, U, U] scores (lse2 None) or softmax probabilities over the REAL tokens -- small grids only."""
    _cuda_only(Q, K)
    if grid.U > 4096:
        raise ValueError("attn_dense materialises [M,U,U] and is meant for small grids (U <= 4096), got U=%d" % grid.U)
    out = torch.empty((M, grid.U, grid.U), dtype=torch.float32, device=Q.device)
    OPS.attn_dense(Q, K, grid.H, grid.W, M, d, float(w_pos), pos_table, R, clip, lse2, int(mask_radius), out)
    return out


# ------------------------------------------------------------------------------------------------
# lookup / small kernels
# ------------------------------------------------------------------------------------------------
def corr_lookup(levels, grid, coords, mean_rstd, out_b=None, out_nchw=None, first_level=0, level0_h16=None):
    """level0_h16: level 0 in blocked fp16 (corr_build(level0_h16=...)), read instead of levels[0]."""
    OPS.corr_lookup(levels[0], levels[1], levels[2], levels[3], grid.H, grid.W, coords, mean_rstd, out_b, out_nchw,
                    first_level, level0_h16)


def corr_lookup0(Q, K, grid, *, M, d, w_agg, w_pos, pos_table, R, clip, coords, mean_rstd, out_b=None, out_nchw=None):
    """Level-0 lookup channels computed on demand from projected Q/K rows (no level-0 volume)."""
    OPS.corr_lookup0(Q, K, grid.H, grid.W, M, d, float(w_agg), float(w_pos), pos_table, R, clip, coords, mean_rstd,
                     out_b, out_nchw)


def convf1(flow, wt, bias, grid, out_b, colo=0):
    OPS.convf1(flow, wt, bias, grid.H, grid.W, out_b, colo)


def flow_update(coords1, flow, delta, grid):
    OPS.flow_update(coords1, flow, delta, grid.H, grid.W)


def init_coords(coords1, flow_init, grid):
    OPS.init_coords(coords1, flow_init, grid.H, grid.W)


def upsample_flow(mask, flow, grid, out=None):
    if out is None:
        out = torch.empty((2, 8 * grid.H, 8 * grid.W), dtype=torch.float32, device=flow.device)
    _cuda_only(mask, flow, out)
    OPS.upsample_flow(mask, flow, grid.H, grid.W, out)
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
    _act_dtype(x_nhwc, "x")
    N, H, W, Cc = x_nhwc.shape
    part = torch.empty((1024 * N * Cc * 2,), dtype=torch.float32, device=x_nhwc.device)   # per-block partial sums
    ab = torch.empty((N, Cc, 2), dtype=torch.float32, device=x_nhwc.device)
    OPS.nhwc_instnorm_stats(x_nhwc, N, H * W, Cc, float(eps), part, ab)
    return ab


def instnorm_apply(x, res=None, rab=None, relu_in=False, relu_out=False, eps=1e-5, out=None, return_ab=False):
    """out = relu_out([ra*res+rb | res] + relu_in(instance_norm(x))) in one cooperative launch that reads x once
    (include/craft_b200.h craft_nhwc_instnorm_apply).  x/res/out [N,H,W,C] contiguous f32 or f16; rab f32 [N or 1,C,2]."""
    _act_dtype(x, "x")
    _chk(res, x.dtype, "res")
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    _chk(out, x.dtype, "out")
    # per-call scratch (barrier state in the first 64 floats, zero before use; then per-CTA partial sums): never shared
    # between CUDA graphs / lanes
    part = torch.zeros((64 + 256 * 512,), dtype=torch.float32, device=x.device)
    ab = torch.empty((N, Cc, 2), dtype=torch.float32, device=x.device) if return_ab else None
    st = 0 if (rab is None or rab.shape[0] == 1) else 2 * Cc
    OPS.nhwc_instnorm_apply(x, res, rab, st, bool(relu_in), bool(relu_out), N, H * W, Cc, float(eps), part, ab, out)
    return (out, ab) if return_ab else out


class PadAct:
    """A channels-last activation in the padded-flat layout of craft_conv3x3_c64: t [N*(H+1)*(W+2), C], row
    (n*(H+1)+y)*(W+2)+x, zeros in the cells x >= W and in the row y = H of every image."""

    def __init__(self, t, N, H, W):
        self.t, self.N, self.H, self.W = t, N, H, W

    @staticmethod
    def rows(N, H, W):
        return N * (H + 1) * (W + 2)

    def dense(self):
        """[N,H,W,C] copy (tests)."""
        C_ = self.t.shape[1]
        return self.t.view(self.N, self.H + 1, self.W + 2, C_)[:, :self.H, :self.W].contiguous()


def conv3x3_c64(x, w, bias=None, relu=False, stats_eps=None):
    """3x3 'same' convolution 64 -> 64 on a PadAct (f16); w from pack_conv64_weight.  stats_eps: also return the
    InstanceNorm2d (rstd, -mean*rstd) [N,64,2] of the fp32 result (accumulated in the convolution's epilogue)."""
    assert isinstance(x, PadAct) and x.t.dtype == torch.float16 and x.t.shape[1] == 64
    out = torch.empty_like(x.t)
    part = ab = None
    if stats_eps is not None:
        # per-call scratch (one (sum, sumsq) pair per channel and CTA): a buffer cached per stream would be shared by
        # CUDA graphs captured on recycled stream handles and raced on by lanes replaying them side by side
        part = torch.empty((256 * 128,), dtype=torch.float32, device=x.t.device)
        ab = torch.empty((x.N, 64, 2), dtype=torch.float32, device=x.t.device)
    OPS.conv3x3_c64(x.t, w, bias, bool(relu), x.N, x.H, x.W, out, part, ab, float(stats_eps or 0.0))
    o = PadAct(out, x.N, x.H, x.W)
    return (o, ab) if stats_eps is not None else o


def pack_conv64_weight(w, scale=None):
    """[64,64,3,3] -> f16 [576, 64]: tap-major (ky,kx row-major) blocks of [cout][cin]; scale: per-cout factor (folded norm)."""
    ww = w.detach().float()
    if scale is not None:
        ww = ww * scale.view(-1, 1, 1, 1)
    return ww.permute(2, 3, 0, 1).reshape(9 * 64, 64).to(torch.float16).contiguous()


def nhwc_affine_pad(v, ab=None, res=None, rab=None, relu_in=False, relu_out=False, out_pad=True):
    """nhwc_affine between dense [N,H,W,C] tensors and PadAct operands (any combination).  Returns a PadAct when
    out_pad else a dense [N,H,W,C] tensor."""
    vp, rp = isinstance(v, PadAct), isinstance(res, PadAct)
    src = v.t if vp else v
    _act_dtype(src, "v")
    if vp:
        N, H, W, Cc = v.N, v.H, v.W, v.t.shape[1]
    else:
        N, H, W, Cc = v.shape
    rsrc = (res.t if rp else res) if res is not None else None
    _chk(rsrc, src.dtype, "res")
    if out_pad:
        out = torch.empty((PadAct.rows(N, H, W), Cc), dtype=src.dtype, device=src.device)
    else:
        out = torch.empty((N, H, W, Cc), dtype=src.dtype, device=src.device)
    st = lambda t: 0 if (t is None or t.shape[0] == 1) else 2 * Cc
    OPS.nhwc_affine_pad(src, vp, ab, st(ab), rsrc, rp, rab, st(rab), bool(relu_in), bool(relu_out), N, H, W, Cc, out, bool(out_pad))
    return PadAct(out, N, H, W) if out_pad else out


def image_s2d(img, dtype=torch.float16):
    """[N,3,H,W] f32 frames (0..255) -> normalised, 2x2 space-to-depth, channels-last, zero-bordered input of the
    encoders' first convolution: [N, H/2+3, W/2+3, 16] (include/craft_b200.h craft_image_s2d)."""
    _cuda_only(img)
    N, _, H, W = img.shape
    out = torch.empty((N, H // 2 + 3, W // 2 + 3, 16), dtype=dtype, device=img.device)
    OPS.image_s2d(img.contiguous(), out)
    return out


def nhwc_affine(v, ab=None, res=None, rab=None, relu_in=False, relu_out=False, out=None):
    """out = relu_out([ra*res+rb] + relu_in(a*v+b)); v/res/out [N,H,W,C] f32 or f16; ab/rab f32 [N or 1, C, 2]."""
    _act_dtype(v, "v")
    _chk(res, v.dtype, "res")
    N, H, W, Cc = v.shape
    if out is None:
        out = torch.empty_like(v)
    _chk(out, v.dtype, "out")
    st = lambda t: 0 if (t is None or t.shape[0] == 1) else 2 * Cc
    OPS.nhwc_affine(v, ab, st(ab), res, rab, st(rab), bool(relu_in), bool(relu_out), N, H * W, Cc, out)
    return out
