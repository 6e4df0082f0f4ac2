"""CPU-side checks of the C-ABI boundary: the library loads and exports exactly what
include/craft_b200.h declares, and the ctypes mirror (craft_b200/_lib.py) agrees with it."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "craft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(craft_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _header_functions()
    assert "craft_shift_gemm" in names and "craft_corr_build" in names and "craft_attn_pv" in names
    assert len(names) >= 19


def test_library_exports_every_declared_symbol():
    from craft_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    for fp16 in (False, True):          # both builds: bf16 operands (default) and fp16 operands (fp32-parity tier)
        lib = _lib.load(fp16)
        for name in _header_functions():
            assert hasattr(lib, name), "%s does not export %s" % ("fp16 build" if fp16 else "bf16 build", name)
        assert lib.craft_b200_abi_version() == _lib.ABI_VERSION


def test_ctypes_mirror_matches_header():
    from craft_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_missing_device_fails_loudly():
    """No silent CPU fallback: a CPU tensor must raise, not compute."""
    import torch
    from craft_b200 import _lib, ops
    grid = ops.TokenGrid(4, 4)
    with pytest.raises((_lib.CraftB200Error, AssertionError)):
        ops.pack_tokens(torch.zeros((128, 4, 4)), grid, out_f=torch.zeros((grid.Mp, 128)))


def _ranges(NT, G):
    return [(NT * c // G, NT * (c + 1) // G) for c in range(G)]


@pytest.mark.parametrize("H,W,M", [(56, 128, 4), (48, 156, 4), (55, 128, 4), (16, 16, 4), (16, 16, 1), (50, 90, 1)])
def test_persistent_schedule_slot_counts(H, W, M):
    """Host-side arithmetic of the persistent-CTA schedules (no kernel launch): the slot counts the
    library reports must equal the largest number of CTA ranges that intersect one unit when the
    (unit, key tile) list is cut into equal contiguous ranges -- the same formula the kernels and
    modes_finalize evaluate on the device (attn_pv.cuh, scores.cuh, pointwise.cuh)."""
    from craft_b200 import _lib
    lib = _lib.load()
    info = (3 * __import__("ctypes").c_int)()
    sms = 148                                  # the library falls back to 148 SMs without a device
    if lib.craft_b200_device_info(info) == 0 and info[0] > 0:
        sms = info[0]
    Mp = H * (W + 2)
    nqt = (Mp + 127) // 128

    def worst(nunits, nkt, G):
        NT = nunits * nkt
        rng = _ranges(NT, G)
        assert rng[0][0] == 0 and rng[-1][1] == NT and all(a < b for a, b in rng)      # cover, non-empty
        out = 0
        for u in range(nunits):
            lo, hi = u * nkt, (u + 1) * nkt
            out = max(out, sum(1 for a, b in rng if a < hi and b > lo))
        return out

    # scores kernels: unit = query tile, 8x8 key blocks
    nkt = ((H + 7) // 8) * ((W + 7) // 8)
    G = max(1, min(sms, nqt * nkt))
    assert lib.craft_scores_auto_ksplit(H, W) == worst(nqt, nkt, G)
    # P.V kernel: unit = (query tile, mode); 8x16 or 8x8 key blocks; grid capped at 3 CTAs per unit
    want = 0
    for bw in (16, 8):
        nkt = ((H + 7) // 8) * ((W + bw - 1) // bw)
        G = max(1, min(sms, nqt * M * nkt, 3 * nqt * M))
        want = max(want, worst(nqt * M, nkt, G))
    got = lib.craft_pv_auto_ksplit(H, W, M)
    assert got == want and got <= 4


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The three argument structs are mirrored by hand in craft_b200/_lib.py: compile the header with gcc
    and compare size and the offset of every field, so a field added on one side only cannot go unnoticed."""
    import ctypes as C
    import shutil
    import subprocess
    from craft_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = [("craft_gemm_args", _lib.GemmArgs), ("craft_scores_args", _lib.ScoresArgs), ("craft_pv_args", _lib.PvArgs),
             ("craft_dense_attn_args", _lib.DenseAttnArgs)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "craft_b200.h"', 'int main(void) {']
    for cname, st in pairs:
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in st._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for cname, st in pairs:
        assert got[(cname, "sizeof")] == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert got[(cname, fname)] == getattr(st, fname).offset, (cname, fname)


def test_hot_kernels_are_tcgen05_and_tma_in_sass():
    """The built library must carry Blackwell tensor-core / TMA / TMEM instructions in its hot kernels
    (SASS mnemonics from the profiling guide: UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM/STTM =
    tcgen05.ld/st, UTCBAR = tcgen05.commit) and be compiled for sm_100a only -- a guard against a silent
    fallback to mma.sync / plain loads creeping in."""
    import shutil
    import subprocess
    from craft_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    archs = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in archs and "sm_90" not in archs and "sm_80" not in archs, archs
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per_kernel, name = {}, None
    for ln in sass.splitlines():
        if "Function :" in ln:
            name = ln.split("Function :")[1].strip()
            per_kernel[name] = set()
        elif name:
            for m in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "HMMA"):
                if re.search(r"\b" + m + r"\b", ln):          # whole mnemonic: HMMA must not match UTCHMMA
                    per_kernel[name].add(m)

    def ops_of(fragment):
        hits = [v for k, v in per_kernel.items() if fragment in k]
        assert hits, "no kernel matching %s in the library" % fragment
        return set.union(*hits)
    for frag in ("shift_gemm_kernel", "scores_kernel", "attn_pv_kernel"):
        ops = ops_of(frag)
        assert {"UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"} <= ops, (frag, ops)
        assert "HMMA" not in ops, frag                       # no legacy mma.sync in the tensor kernels
    assert "STTM" in ops_of("attn_pv_kernel")                # P is written back to TMEM (tcgen05.st)


def test_no_global_load_is_hoisted_above_the_pdl_wait():
    """Every kernel is launched with programmatic dependent launch and may start while its predecessor is
    still running; nothing may touch global memory before griddepcontrol.wait (SASS: ACQBULK).  nvcc hoists
    invariant loads (`const T* __restrict__`, __ldg) above the wait when it can -- that silently broke the
    clamp gate under CUDA-graph replay (round 2) -- so the SASS of the built library is audited."""
    import shutil
    import sys
    from craft_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    from audit_pdl_hoist import audit
    for path in (_lib.LIB_PATH, _lib.LIB_PATH_FP16):
        bad = audit(path)
        assert not bad, (path, bad)
