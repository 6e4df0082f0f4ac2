"""ctypes binding of libcraft_b200.so (C ABI declared in include/craft_b200.h).

The library is the product: if it is missing or a call fails this module raises -- there is no
PyTorch/CPU fallback anywhere in craft_b200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcraft_b200.so")              # tensor-core operands / activations in bf16
LIB_PATH_FP16 = os.path.join(_HERE, "libcraft_b200_fp16.so")    # same sources built with fp16 operands
MAX_TAPS = 49
ABI_VERSION = 2


class CraftB200Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_rows", C.c_int), ("lda", C.c_int), ("a_koff", C.c_int),
        ("B", C.c_void_p), ("b_rows", C.c_int), ("ldb_", C.c_int), ("b_koff", C.c_int),
        ("b_blocked", C.c_int), ("b_H", C.c_int), ("b_W", C.c_int),
        ("M", C.c_int), ("Npad", C.c_int), ("K", C.c_int), ("T", C.c_int),
        ("BN", C.c_int), ("cluster", C.c_int), ("stages", C.c_int),
        ("tap_off", C.c_int * MAX_TAPS),
        ("H", C.c_int), ("W", C.c_int),
        ("epilogue", C.c_int),
        ("alpha", C.c_float),
        ("act", C.c_int),
        ("bias", C.c_void_p),
        ("out_bf16", C.c_void_p), ("ldo_b", C.c_int), ("colo_b", C.c_int),
        ("out_f32", C.c_void_p), ("ldo_f", C.c_int), ("colo_f", C.c_int),
        ("aux0", C.c_void_p), ("aux1", C.c_void_p),
        ("a_share", C.c_int),
    ]


class ScoresArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("K", C.c_void_p),
        ("C", C.c_int), ("M", C.c_int), ("d", C.c_int),
        ("H", C.c_int), ("W", C.c_int),
        ("scale", C.c_float), ("w_pos", C.c_float),
        ("pos_table", C.c_void_p), ("R", C.c_int),
        ("clip", C.c_void_p), ("run_flag", C.c_void_p),
        ("ksplit", C.c_int),
        ("w_agg", C.c_float),
        ("stat_sum", C.c_void_p), ("stat_max", C.c_void_p),
        ("lvl", C.c_void_p * 4),
        ("lse_part", C.c_void_p), ("lse2", C.c_void_p),
        ("mask_radius", C.c_int),
        ("lvl0_h16", C.c_void_p),
    ]


class PvArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("K", C.c_void_p), ("Vt", C.c_void_p), ("ldv", C.c_int),
        ("C", C.c_int), ("M", C.c_int), ("d", C.c_int), ("F", C.c_int),
        ("H", C.c_int), ("W", C.c_int),
        ("scale", C.c_float), ("w_pos", C.c_float),
        ("pos_table", C.c_void_p), ("R", C.c_int),
        ("clip", C.c_void_p), ("lse2", C.c_void_p), ("out", C.c_void_p),
        ("ksplit", C.c_int), ("zero_fill", C.c_int), ("mask_radius", C.c_int),
    ]


class DenseAttnArgs(C.Structure):
    _fields_ = [
        ("Q", C.c_void_p), ("K", C.c_void_p),
        ("C", C.c_int), ("M", C.c_int), ("d", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("scale", C.c_float), ("w_pos", C.c_float),
        ("pos_table", C.c_void_p), ("R", C.c_int),
        ("clip", C.c_void_p), ("lse2", C.c_void_p),
        ("mask_radius", C.c_int),
        ("out", C.c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/craft_b200.h one to one
_vp, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double
SIGNATURES = {
    "craft_b200_abi_version": (_i, []),
    "craft_b200_last_error": (C.c_char_p, []),
    "craft_b200_launch_count": (C.c_longlong, []),
    "craft_b200_bigbox_gemm_count": (C.c_longlong, []),
    "craft_b200_device_info": (_i, [C.POINTER(_i)]),
    "craft_pack_tokens": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "craft_pack_tokens_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "craft_unpack_tokens": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "craft_shift_gemm": (_i, [C.POINTER(GemmArgs), _vp]),
    "craft_corr_build": (_i, [C.POINTER(ScoresArgs), _vp]),
    "craft_attn_lse": (_i, [C.POINTER(ScoresArgs), _vp]),
    "craft_scores_auto_ksplit": (_i, [_i, _i]),
    "craft_pv_auto_ksplit": (_i, [_i, _i, _i]),
    "craft_pv_block_keys": (_i, [_i, _i]),
    "craft_corr_stats_finalize": (_i, [_vp, _vp, _d, _vp, _vp]),
    "craft_clip_gate": (_i, [_vp, _f, _vp, _vp, _vp, _vp]),
    "craft_attn_pv": (_i, [C.POINTER(PvArgs), _vp]),
    "craft_modes_finalize": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _i, _i, _i,
                                  _vp, _i, _i, _vp, _i, _i, _i, _vp]),
    "craft_soft_aggregate": (_i, [_vp, _vp, _i, C.c_longlong, _i, _vp, _vp, _vp, _vp]),
    "craft_attn_dense": (_i, [C.POINTER(DenseAttnArgs), _vp]),
    "craft_corr_lookup": (_i, [C.POINTER(_vp), _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "craft_corr_lookup0": (_i, [_vp, _vp, _i, _i, _f, _f, _f, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "craft_convf1": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp]),
    "craft_flow_update": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "craft_init_coords": (_i, [_vp, _vp, _i, _i, _vp]),
    "craft_upsample_flow": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    "craft_forward_interpolate": (_i, [_vp, _i, _i, _vp, _vp]),
    "craft_flow_encode": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "craft_nhwc_instnorm_stats": (_i, [_vp, _i, _i, _i, _i, _f, _vp, C.c_longlong, _vp, _vp]),
    "craft_image_s2d": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "craft_nhwc_instnorm_apply": (_i, [_vp, _i, _i, _i, _i, _f, _vp, _vp, _i, _i, _i, _vp, C.c_longlong, _vp, _vp, _vp]),
    "craft_conv3x3_c64": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, C.c_longlong, _vp, _f, _vp]),
    "craft_nhwc_affine_pad": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "craft_nhwc_affine": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

_libs = {}


def load(fp16=False):
    """Load (once) and type one of the two builds of the shared library.  Raises CraftB200Error if it is
    not built."""
    hit = _libs.get(bool(fp16))
    if hit is not None:
        return hit
    path = LIB_PATH_FP16 if fp16 else LIB_PATH
    if not os.path.isfile(path):
        raise CraftB200Error(
            "%s is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C craft_b200/csrc`. There is no fallback path." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.craft_b200_abi_version() != ABI_VERSION:
        raise CraftB200Error("%s: ABI version mismatch" % os.path.basename(path))
    _libs[bool(fp16)] = lib
    return lib


def launch_count():
    """Kernels launched so far by both builds of the library."""
    return sum(lib.craft_b200_launch_count() for lib in _libs.values())


def check(rc, what="", fp16=False):
    if rc != 0:
        msg = load(fp16).craft_b200_last_error()
        raise CraftB200Error("%s failed: %s" % (what or "craft_b200 call", msg.decode() if msg else rc))


def call(name, *args, fp16=False):
    lib = load(fp16)
    check(getattr(lib, name)(*args), name, fp16)
