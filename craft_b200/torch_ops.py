"""torch custom-op registration of the sm_100a kernels: `torch.ops.craft_b200.*`.

Each operator is a flat-signature binding of one C-ABI entry point of libcraft_b200.so
(include/craft_b200.h): tensors, ints and floats in, results written into caller-provided tensors
(declared as mutated in the schema).  The CUDA implementation marshals the arguments and launches
the kernel on the current stream; there is deliberately NO CPU / CompositeImplicit implementation
(a CPU tensor reaches the dispatcher's "no kernel for backend CPU" error).  Every op has a fake
(meta) implementation, so the ops can be traced (`torch.compile`, DDP graph capture); autograd is
attached at module level (craft_b200/train_path.py), not per kernel.

craft_b200/ops.py is the convenience layer on top (TokenGrid objects, keyword arguments); the
nn.Modules call through it, so the whole model runs through these registered operators.
"""
import ctypes as C
import math

import torch
from torch.library import Library

from . import _lib
from ._lib import DenseAttnArgs, GemmArgs, PvArgs, ScoresArgs

NS = "craft_b200"
_def = Library(NS, "DEF")
_cuda = Library(NS, "IMPL", "CUDA")
_meta = Library(NS, "IMPL", "Meta")
SCHEMAS = {}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _dp(t):
    return t.data_ptr() if t is not None else None


def _ld(t):
    return t.shape[-1] if t is not None else 0


def _chk(t, dtype, name):
    if t is None:
        return
    if dtype == ACT:
        if t.dtype not in (torch.bfloat16, torch.float16):
            raise TypeError("%s: expected bfloat16 or float16, got %s" % (name, t.dtype))
    elif t.dtype != dtype:
        raise TypeError("%s: expected %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)


def _op(schema):
    """Register `schema` with a CUDA implementation (the decorated function) and a no-op fake: every
    operator writes into caller-provided tensors and returns nothing, so shape inference is trivial."""
    name = schema.split("(")[0]
    SCHEMAS[name] = schema

    def deco(fn):
        _def.define(schema)
        _cuda.impl(name, fn)
        _meta.impl(name, lambda *a, **k: None)
        return fn
    return deco


bf16, f16, f32, f64, i32 = torch.bfloat16, torch.float16, torch.float32, torch.float64, torch.int32
ACT = "act"        # placeholder dtype for _chk: either 16-bit operand type


def _half(*ts):
    """Which build of the library a call goes to: the fp16 one iff its 16-bit tensors are float16."""
    kinds = {t.dtype for t in ts if t is not None and t.dtype in (bf16, f16)}
    if len(kinds) > 1:
        raise TypeError("mixed bf16 / fp16 operands in one call")
    return f16 in kinds


# ------------------------------------------------------------------------------------------------
@_op("pack_tokens(Tensor src, int H, int W, int mode, Tensor(a!)? out_b, int colb, Tensor(b!)? out_f, int colf) -> ()")
def pack_tokens(src, H, W, mode, out_b, colb, out_f, colf):
    _chk(src, f32, "src"); _chk(out_b, ACT, "out_b"); _chk(out_f, f32, "out_f")
    assert src.shape[1] == H and src.shape[2] == W
    _lib.call("craft_pack_tokens", _ptr(src), src.shape[0], H, W, mode, _ptr(out_b), _ld(out_b), colb,
              _ptr(out_f), _ld(out_f), colf, _stream(), fp16=_half(out_b))


@_op("pack_tokens_nhwc(Tensor src, int c0, int C, int H, int W, int mode, Tensor(a!)? out_b, int colb, Tensor(b!)? out_f, int colf) -> ()")
def pack_tokens_nhwc(src, c0, Cc, H, W, mode, out_b, colb, out_f, colf):
    """src [H, W, ldc] channels-last, f16 or f32, contiguous."""
    if src.dtype not in (f16, f32):
        raise TypeError("src: expected float16 or float32, got %s" % src.dtype)
    _chk(src, src.dtype, "src"); _chk(out_b, ACT, "out_b"); _chk(out_f, f32, "out_f")
    assert src.shape[0] == H and src.shape[1] == W
    _lib.call("craft_pack_tokens_nhwc", _ptr(src), 1 if src.dtype == f16 else 0, src.shape[2], c0, Cc, H, W, mode,
              _ptr(out_b), _ld(out_b), colb, _ptr(out_f), _ld(out_f), colf, _stream(), fp16=_half(out_b))


@_op("unpack_tokens(Tensor buf, int col, int C, int H, int W, Tensor(a!) out) -> ()")
def unpack_tokens(buf, col, Cc, H, W, out):
    is_b = 1 if buf.dtype in (bf16, f16) else 0
    if not is_b:
        _chk(buf, f32, "buf")
    _chk(out, f32, "out")
    _lib.call("craft_unpack_tokens", _ptr(buf), is_b, _ld(buf), col, Cc, H, W, _ptr(out), _stream(), fp16=_half(buf))


@_op("shift_gemm(Tensor A, Tensor B, int M, int Npad, int K, int BN, int[] taps, int a_koff, int b_koff, int H, int W, "
     "int epilogue, float alpha, int act, Tensor? bias, Tensor(a!)? out_b, int colb, Tensor(b!)? out_f, int colf, "
     "Tensor(c!)? aux0, Tensor(d!)? aux1, int b_H, int b_W, int cluster, int stages, int a_share) -> ()")
def shift_gemm(A, Bw, M, Npad, K, BN, taps, a_koff, b_koff, H, W, epilogue, alpha, act, bias, out_b, colb, out_f, colf,
               aux0, aux1, b_H, b_W, cluster, stages, a_share):
    _chk(A, ACT, "A"); _chk(Bw, ACT, "B"); _chk(bias, f32, "bias"); _chk(out_b, ACT, "out_b")
    _chk(out_f, f32, "out_f"); _chk(aux0, f32, "aux0"); _chk(aux1, f32, "aux1")
    a = GemmArgs()
    a.A, a.a_rows, a.lda, a.a_koff = A.data_ptr(), A.shape[0], A.shape[1], a_koff
    a.B, a.b_rows, a.ldb_, a.b_koff = Bw.data_ptr(), Bw.shape[0], Bw.shape[1], b_koff
    a.M, a.Npad, a.K, a.T, a.BN = M, Npad, K, len(taps), BN
    a.cluster, a.stages, a.a_share = cluster, stages, a_share
    if b_H > 0:
        a.b_blocked, a.b_H, a.b_W = 1, b_H, b_W
    for i, t in enumerate(taps):
        a.tap_off[i] = int(t)
    a.H, a.W = H, W
    a.epilogue, a.alpha, a.act = epilogue, float(alpha), act
    a.bias = _dp(bias)
    a.out_bf16, a.ldo_b, a.colo_b = _dp(out_b), _ld(out_b), colb
    a.out_f32, a.ldo_f, a.colo_f = _dp(out_f), _ld(out_f), colf
    a.aux0, a.aux1 = _dp(aux0), _dp(aux1)
    if bias is not None:
        assert bias.numel() >= Npad
    _lib.call("craft_shift_gemm", C.byref(a), _stream(), fp16=_half(A, Bw, out_b))


def _scores_args(Q, K, H, W, M, d, w_pos, pos_table, R, clip, run_flag, ksplit):
    _chk(Q, ACT, "Q"); _chk(K, ACT, "K"); _chk(pos_table, f32, "pos_table"); _chk(clip, f32, "clip")
    Mp = H * (W + 2)
    assert tuple(Q.shape) == (Mp, M * d) and tuple(K.shape) == (Mp, M * d)
    a = ScoresArgs()
    a.Q, a.K = Q.data_ptr(), K.data_ptr()
    a.C, a.M, a.d = M * d, M, d
    a.H, a.W = H, W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = _dp(pos_table)
    a.R = R
    a.clip = clip.data_ptr()
    a.run_flag = _dp(run_flag)
    a.ksplit = ksplit
    return a


@_op("corr_build(Tensor Q, Tensor K, int H, int W, int M, int d, float w_agg, float w_pos, Tensor? pos_table, int R, "
     "Tensor clip, Tensor(a!) stat_sum, Tensor(b!) stat_max, Tensor(c!)? lvl0, Tensor(d!) lvl1, Tensor(e!) lvl2, "
     "Tensor(f!) lvl3, Tensor? run_flag, int ksplit, Tensor(g!)? lvl0_h16) -> ()")
def corr_build(Q, K, H, W, M, d, w_agg, w_pos, pos_table, R, clip, stat_sum, stat_max, lvl0, lvl1, lvl2, lvl3, run_flag,
               ksplit, lvl0_h16):
    a = _scores_args(Q, K, H, W, M, d, w_pos, pos_table, R, clip, run_flag, ksplit)
    a.w_agg = float(w_agg)
    _chk(stat_sum, f64, "stat_sum"); _chk(stat_max, f32, "stat_max")
    a.stat_sum, a.stat_max = stat_sum.data_ptr(), stat_max.data_ptr()
    for l, lv in enumerate((lvl0, lvl1, lvl2, lvl3)):
        _chk(lv, f32, "level")
        a.lvl[l] = _dp(lv)
    if lvl0_h16 is not None:
        _chk(lvl0_h16, f16, "lvl0_h16")
        assert lvl0_h16.numel() >= H * (W + 2) * ((H + 7) // 8) * ((W + 7) // 8) * 68   # fp16 deltas + fp32 half-block means
    a.lvl0_h16 = _dp(lvl0_h16)
    _lib.call("craft_corr_build", C.byref(a), _stream(), fp16=_half(Q, K))


@_op("attn_lse(Tensor Q, Tensor K, int H, int W, int M, int d, float w_pos, Tensor? pos_table, int R, Tensor clip, "
     "Tensor(a!) stat_max, Tensor(b!) lse_part, Tensor(c!) lse2, Tensor? run_flag, int ksplit, int mask_radius) -> ()")
def attn_lse(Q, K, H, W, M, d, w_pos, pos_table, R, clip, stat_max, lse_part, lse2, run_flag, ksplit, mask_radius):
    a = _scores_args(Q, K, H, W, M, d, w_pos, pos_table, R, clip, run_flag, ksplit)
    _chk(stat_max, f32, "stat_max"); _chk(lse_part, f32, "lse_part"); _chk(lse2, f32, "lse2")
    a.stat_max = stat_max.data_ptr()
    a.lse_part, a.lse2 = lse_part.data_ptr(), lse2.data_ptr()
    a.mask_radius = int(mask_radius)
    _lib.call("craft_attn_lse", C.byref(a), _stream(), fp16=_half(Q, K))


@_op("attn_pv(Tensor Q, Tensor K, Tensor Vt, int H, int W, int M, int d, int F, float w_pos, Tensor? pos_table, int R, "
     "Tensor clip, Tensor lse2, Tensor(a!) out, int ksplit, bool zero_fill, int mask_radius) -> ()")
def attn_pv(Q, K, Vt, H, W, M, d, F, w_pos, pos_table, R, clip, lse2, out, ksplit, zero_fill, mask_radius):
    _chk(Q, ACT, "Q"); _chk(K, ACT, "K"); _chk(Vt, ACT, "Vt"); _chk(out, f32, "out"); _chk(lse2, f32, "lse2")
    Mp = H * (W + 2)
    assert Vt.shape[0] >= M * F and out.numel() >= ksplit * M * Mp * F
    a = PvArgs()
    a.Q, a.K, a.Vt, a.ldv = Q.data_ptr(), K.data_ptr(), Vt.data_ptr(), Vt.shape[1]
    a.C, a.M, a.d, a.F = M * d, M, d, F
    a.H, a.W = H, W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = _dp(pos_table)
    a.R = R
    a.clip, a.lse2, a.out = clip.data_ptr(), lse2.data_ptr(), out.data_ptr()
    a.ksplit = ksplit
    a.zero_fill = 1 if zero_fill else 0
    a.mask_radius = int(mask_radius)
    _lib.call("craft_attn_pv", C.byref(a), _stream(), fp16=_half(Q, K, Vt))


@_op("modes_finalize(Tensor O, int nsum, int M, int F, int H, int W, Tensor w_score, Tensor b_score, Tensor coeff, int gma, "
     "Tensor? x_b, int colx, Tensor? x_f, int colxf, Tensor(a!)? out_b, int colb, Tensor(b!)? out_f, int colf, int pv_bk) -> ()")
def modes_finalize(O, nsum, M, F, H, W, w_score, b_score, coeff, gma, x_b, colx, x_f, colxf, out_b, colb, out_f, colf,
                   pv_bk):
    _chk(O, f32, "O")
    _lib.call("craft_modes_finalize", _ptr(O), nsum, M, F, _ptr(w_score), _ptr(b_score), _ptr(coeff), gma,
              _ptr(x_b), _ld(x_b), colx, _ptr(x_f), _ld(x_f), colxf, H, W,
              _ptr(out_b), _ld(out_b), colb, _ptr(out_f), _ld(out_f), colf, int(pv_bk), _stream(), fp16=_half(x_b, out_b))


@_op("corr_stats_finalize(Tensor stat_sum, Tensor? flag, float n, Tensor(a!) mean_rstd) -> ()")
def corr_stats_finalize(stat_sum, flag, n, mean_rstd):
    _chk(stat_sum, f64, "stat_sum")
    _lib.call("craft_corr_stats_finalize", _ptr(stat_sum), _ptr(flag), float(n), _ptr(mean_rstd), _stream())


@_op("clip_gate(Tensor stat_max, float attn_clip, Tensor(a!) clip, Tensor(b!) flag, Tensor(c!)? diag) -> ()")
def clip_gate(stat_max, attn_clip, clip, flag, diag):
    _chk(diag, f32, "diag")
    _lib.call("craft_clip_gate", _ptr(stat_max), float(attn_clip), _ptr(clip), _ptr(flag), _ptr(diag), _stream())


@_op("soft_aggregate(Tensor x, Tensor? basis, int M, int n, int F, Tensor w, Tensor b, Tensor(a!) out) -> ()")
def soft_aggregate(x, basis, M, n, F, w, b, out):
    _chk(x, f32, "x"); _chk(basis, f32, "basis"); _chk(w, f32, "w"); _chk(b, f32, "b"); _chk(out, f32, "out")
    _lib.call("craft_soft_aggregate", _ptr(x), _ptr(basis), M, n, F, _ptr(w), _ptr(b), _ptr(out), _stream())


@_op("attn_dense(Tensor Q, Tensor K, int H, int W, int M, int d, float w_pos, Tensor? pos_table, int R, Tensor clip, "
     "Tensor? lse2, int mask_radius, Tensor(a!) out) -> ()")
def attn_dense(Q, K, H, W, M, d, w_pos, pos_table, R, clip, lse2, mask_radius, out):
    _chk(Q, ACT, "Q"); _chk(K, ACT, "K"); _chk(lse2, f32, "lse2"); _chk(out, f32, "out")
    a = DenseAttnArgs()
    a.Q, a.K = Q.data_ptr(), K.data_ptr()
    a.C, a.M, a.d, a.H, a.W = M * d, M, d, H, W
    a.scale, a.w_pos = 1.0 / math.sqrt(d), float(w_pos)
    a.pos_table = _dp(pos_table)
    a.R = R
    a.clip = clip.data_ptr()
    a.lse2 = _dp(lse2)
    a.mask_radius = int(mask_radius)
    a.out = out.data_ptr()
    _lib.call("craft_attn_dense", C.byref(a), _stream(), fp16=_half(Q, K))


@_op("corr_lookup(Tensor? lvl0, Tensor? lvl1, Tensor? lvl2, Tensor? lvl3, int H, int W, Tensor coords, Tensor mean_rstd, "
     "Tensor(a!)? out_b, Tensor(b!)? out_nchw, int first_level, Tensor? lvl0_h16) -> ()")
def corr_lookup(lvl0, lvl1, lvl2, lvl3, H, W, coords, mean_rstd, out_b, out_nchw, first_level, lvl0_h16):
    arr = (C.c_void_p * 4)(*[_dp(lv) for lv in (lvl0, lvl1, lvl2, lvl3)])
    _chk(coords, f32, "coords"); _chk(out_b, ACT, "out_b"); _chk(out_nchw, f32, "out_nchw")
    _chk(lvl0_h16, f16, "lvl0_h16")
    for lv in (lvl0, lvl1, lvl2, lvl3):
        _chk(lv, f32, "level")
    _lib.call("craft_corr_lookup", arr, _ptr(lvl0_h16), H, W, _ptr(coords), _ptr(mean_rstd), _ptr(out_b), _ld(out_b),
              _ptr(out_nchw), first_level, _stream(), fp16=_half(out_b))


@_op("corr_lookup0(Tensor Q, Tensor K, int H, int W, int M, int d, float w_agg, float w_pos, Tensor? pos_table, int R, "
     "Tensor clip, Tensor coords, Tensor mean_rstd, Tensor(a!)? out_b, Tensor(b!)? out_nchw) -> ()")
def corr_lookup0(Q, K, H, W, M, d, w_agg, w_pos, pos_table, R, clip, coords, mean_rstd, out_b, out_nchw):
    _chk(Q, ACT, "Q"); _chk(K, ACT, "K"); _chk(out_b, ACT, "out_b"); _chk(out_nchw, f32, "out_nchw")
    _lib.call("craft_corr_lookup0", _ptr(Q), _ptr(K), M, d, 1.0 / math.sqrt(d), float(w_agg), float(w_pos),
              _ptr(pos_table), R, _ptr(clip), H, W, _ptr(coords), _ptr(mean_rstd), _ptr(out_b), _ld(out_b),
              _ptr(out_nchw), _stream(), fp16=_half(Q, K, out_b))


@_op("convf1(Tensor flow, Tensor wt, Tensor bias, int H, int W, Tensor(a!) out_b, int colo) -> ()")
def convf1(flow, wt, bias, H, W, out_b, colo):
    _chk(flow, f32, "flow"); _chk(wt, f32, "wt"); _chk(out_b, ACT, "out_b")
    _lib.call("craft_convf1", _ptr(flow), _ptr(wt), _ptr(bias), H, W, _ptr(out_b), _ld(out_b), colo, _stream(), fp16=_half(out_b))


@_op("flow_update(Tensor(a!) coords1, Tensor(b!) flow, Tensor? delta, int H, int W) -> ()")
def flow_update(coords1, flow, delta, H, W):
    _lib.call("craft_flow_update", _ptr(coords1), _ptr(flow), _ptr(delta), _ld(delta), H, W, _stream())


@_op("init_coords(Tensor(a!) coords1, Tensor? flow_init, int H, int W) -> ()")
def init_coords(coords1, flow_init, H, W):
    _chk(flow_init, f32, "flow_init")
    _lib.call("craft_init_coords", _ptr(coords1), _ptr(flow_init), H, W, _stream())


@_op("upsample_flow(Tensor mask, Tensor flow, int H, int W, Tensor(a!) out) -> ()")
def upsample_flow(mask, flow, H, W, out):
    _chk(flow, f32, "flow"); _chk(out, f32, "out")
    is_b = 1 if mask.dtype in (bf16, f16) else 0
    _lib.call("craft_upsample_flow", _ptr(mask), is_b, _ld(mask), _ptr(flow), H, W, _ptr(out), _stream(), fp16=_half(mask))


@_op("nhwc_instnorm_stats(Tensor x, int N, int HW, int C, float eps, Tensor(a!) part, Tensor(b!) ab) -> ()")
def nhwc_instnorm_stats(x, N, HW, Cc, eps, part, ab):
    half = 1 if x.dtype == torch.float16 else 0
    _lib.call("craft_nhwc_instnorm_stats", _ptr(x), half, N, HW, Cc, float(eps), _ptr(part), part.numel(), _ptr(ab),
              _stream())


@_op("image_s2d(Tensor img, Tensor(a!) out) -> ()")
def image_s2d(img, out):
    """img [N,3,H,W] f32 (0..255) -> out [N, H/2+3, W/2+3, 16] f16/f32 (see include/craft_b200.h)."""
    _chk(img, f32, "img")
    if out.dtype not in (f16, f32):
        raise TypeError("out: expected float16 or float32, got %s" % out.dtype)
    _chk(out, out.dtype, "out")
    N, _, H, W = img.shape
    assert img.shape[1] == 3 and tuple(out.shape) == (N, H // 2 + 3, W // 2 + 3, 16)
    _lib.call("craft_image_s2d", _ptr(img), N, H, W, _ptr(out), 1 if out.dtype == f16 else 0, _stream())


@_op("nhwc_affine(Tensor v, Tensor? ab, int ab_stride, Tensor? res, Tensor? rab, int rab_stride, bool relu_in, bool relu_out, "
     "int N, int HW, int C, Tensor(a!) out) -> ()")
def nhwc_affine(v, ab, ab_stride, res, rab, rab_stride, relu_in, relu_out, N, HW, Cc, out):
    half = 1 if v.dtype == torch.float16 else 0
    _lib.call("craft_nhwc_affine", _ptr(v), half, _ptr(ab), ab_stride, _ptr(res), _ptr(rab), rab_stride, int(relu_in),
              int(relu_out), N, HW, Cc, _ptr(out), _stream())


@_op("nhwc_instnorm_apply(Tensor x, Tensor? res, Tensor? rab, int rab_stride, bool relu_in, bool relu_out, int N, int HW, int C, "
     "float eps, Tensor(a!) part, Tensor(b!)? ab_out, Tensor(c!) out) -> ()")
def nhwc_instnorm_apply(x, res, rab, rab_stride, relu_in, relu_out, N, HW, Cc, eps, part, ab_out, out):
    half = 1 if x.dtype == torch.float16 else 0
    _lib.call("craft_nhwc_instnorm_apply", _ptr(x), half, N, HW, Cc, float(eps), _ptr(res), _ptr(rab), rab_stride,
              int(relu_in), int(relu_out), _ptr(part), part.numel(), _ptr(ab_out), _ptr(out), _stream())


@_op("conv3x3_c64(Tensor x, Tensor w, Tensor? bias, bool relu, int N, int H, int W, Tensor(a!) out, Tensor(b!)? part, "
     "Tensor(c!)? ab, float eps) -> ()")
def conv3x3_c64(x, w, bias, relu, N, H, W, out, part, ab, eps):
    """x / out: padded-flat f16 [N*(H+1)*(W+2), 64]; w: f16 [576, 64] tap-major (include/craft_b200.h)."""
    for t, nm in ((x, "x"), (w, "w"), (out, "out")):
        _chk(t, f16, nm)
    _chk(bias, f32, "bias"); _chk(part, f32, "part"); _chk(ab, f32, "ab")
    rows = N * (H + 1) * (W + 2)
    assert tuple(x.shape) == (rows, 64) and tuple(out.shape) == (rows, 64) and tuple(w.shape) == (576, 64)
    _lib.call("craft_conv3x3_c64", _ptr(x), _ptr(w), _ptr(bias), int(relu), N, H, W, _ptr(out), _ptr(part),
              part.numel() if part is not None else 0, _ptr(ab), float(eps), _stream(), fp16=True)


@_op("nhwc_affine_pad(Tensor v, bool v_pad, Tensor? ab, int ab_stride, Tensor? res, bool res_pad, Tensor? rab, int rab_stride, "
     "bool relu_in, bool relu_out, int N, int H, int W, int C, Tensor(a!) out, bool out_pad) -> ()")
def nhwc_affine_pad(v, v_pad, ab, ab_stride, res, res_pad, rab, rab_stride, relu_in, relu_out, N, H, W, Cc, out, out_pad):
    half = 1 if v.dtype == torch.float16 else 0
    _lib.call("craft_nhwc_affine_pad", _ptr(v), half, int(v_pad), _ptr(ab), ab_stride, _ptr(res), int(res_pad), _ptr(rab),
              rab_stride, int(relu_in), int(relu_out), N, H, W, Cc, _ptr(out), int(out_pad), _stream(), fp16=True)


@_op("forward_interpolate(Tensor flow, int H, int W, Tensor(a!) out) -> ()")
def forward_interpolate(flow, H, W, out):
    _chk(flow, f32, "flow"); _chk(out, f32, "out")
    _lib.call("craft_forward_interpolate", _ptr(flow), H, W, _ptr(out), _stream())


@_op("flow_encode(Tensor flow, int H, int W, int mode, Tensor(a!) out) -> ()")
def flow_encode(flow, H, W, mode, out):
    _chk(flow, f32, "flow")
    _lib.call("craft_flow_encode", _ptr(flow), H, W, mode, _ptr(out), _stream())


OPS = torch.ops.craft_b200
