"""CorrBlock / TransCorrBlock with the reference's signatures (core/corr.py), backed by the fused
correlation-volume kernel (scores.cuh SC_CORR) and the pyramid lookup kernel (pointwise.cuh).

The pyramid comes out of the GEMM epilogue directly (no pooling passes over HBM).  Level 0 is held in one of
three ways (hotpath.level0_mode): fp16 in 8x8-key-block order (default; one lookup kernel
for all four levels), never stored (each lookup recomputes its 10x10 window from the projected query/key rows,
pointwise.cuh corr_lookup0), or the reference's dense fp32 volume (SAVECORR / debugging).
The global layer-norm of the volume (core/corr.py:200-204) is applied inside the lookup as a
deferred affine, so no second pass over the volume exists.
"""
import os

import torch
import torch.nn as nn

from . import hotpath as hp
from . import ops
from .ops import TokenGrid
from .setrans import CrossAttFeatTrans, SETransInputFeatEncoder, get_workspace, _require_inference


class _LookupMixin:
    """CorrBlock.__call__ core/corr.py:47-71."""

    def lookup_rows(self, ws, coords_rows, out_b=None, out_nchw=None, part="all"):
        """part: "all", or -- when level 0 is computed on demand -- "level0" (channels 0..80) / "pooled"
        (channels 81..323): the two halves are independent kernels writing disjoint channels, so the
        caller may run them on different streams."""
        if self.radius != 4 or self.num_levels != 4:
            raise NotImplementedError("craft_b200 lookup kernel is built for radius 4, 4 levels (CRAFT default)")
        first = 0
        if ws.levels[0] is None and ws.level0_h16 is None:      # level 0 on demand: the U x U volume was never stored
            if part in ("all", "level0"):
                ops.corr_lookup0(grid=ws.grid, coords=coords_rows, mean_rstd=ws.mean_rstd, out_b=out_b,
                                 out_nchw=out_nchw, **ws.corr_meta)
            first = 1
            if part == "level0":
                return
        elif part == "level0":
            return                    # stored level 0 (fp16 blocks or fp32): the single pass below covers everything
        ops.corr_lookup(ws.levels, ws.grid, coords_rows, ws.mean_rstd, out_b=out_b, out_nchw=out_nchw,
                        first_level=first, level0_h16=ws.level0_h16 if ws.levels[0] is None else None)

    def __call__(self, coords):
        """coords [B,2,h,w] (x,y) -> [B,324,h,w] fp32 (core/corr.py:47-71)."""
        B, _, h, w = coords.shape
        if B != 1:
            raise NotImplementedError("standalone CorrBlock lookup handles one pair per call (one pyramid is held); "
                                      "CRAFT.forward iterates over the batch itself")
        ws = self._ws
        if ws is None:
            raise RuntimeError("call update() / construct the CorrBlock before looking up")
        g = ws.grid
        with torch.cuda.device(coords.device):
            crow = torch.zeros((g.Mp, 2), dtype=torch.float32, device=coords.device)
            crow.view(g.H, g.Wp, 2)[:, :g.W] = coords[0].float().permute(1, 2, 0)
            out = torch.empty((1, 324, h, w), dtype=torch.float32, device=coords.device)
            self.lookup_rows(ws, crow, out_nchw=out[0])
        return out


class CorrBlock(_LookupMixin):
    """core/corr.py:16-81: plain dot-product volume fmap1^T fmap2 / sqrt(C) (RAFT / GMA baselines)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, do_corr_global_norm=False):
        self.num_levels, self.radius = num_levels, radius
        self.do_corr_global_norm = do_corr_global_norm
        B, Cc, h, w = fmap1.shape
        if B != 1:
            raise NotImplementedError("standalone CorrBlock handles one pair per call")
        if Cc != 256:
            raise NotImplementedError("CorrBlock kernel instantiated for 256-channel fnet features")
        grid = TokenGrid(h, w)
        with torch.cuda.device(fmap1.device):
            ws = self._ws = get_workspace(grid, fmap1.device, "SAVECORR" in os.environ)
            ops.pack_tokens(fmap1[0].float().contiguous(), grid, ops.PACK_COPY, out_b=ws.Qc)
            ops.pack_tokens(fmap2[0].float().contiguous(), grid, ops.PACK_COPY, out_b=ws.Kc)
            hp.build_correlation(ws, ws.Qc, ws.Kc, M=1, d=Cc, w_agg=0.0, table=None, w_pos=0.0,
                                 global_norm=do_corr_global_norm)


class TransCorrBlock(_LookupMixin, nn.Module):
    """core/corr.py:132-207."""

    def __init__(self, config, num_levels=4, radius=4, do_corr_global_norm=False):
        nn.Module.__init__(self)
        self.num_levels, self.radius = num_levels, radius
        self.config = config
        self.setrans = CrossAttFeatTrans(self.config, "Inter-frame correlation block")
        self.vispos_encoder = SETransInputFeatEncoder(self.config)
        self.coords2 = None
        self.do_corr_global_norm = do_corr_global_norm
        self._ws = None

    def build_rows(self, ws, T1, T2):
        """tokens (already LayerNorm'ed, bf16 rows) -> pyramid + statistics in the workspace."""
        st = self.setrans
        st.project(ws, T1, T2, ws.Qc, ws.Kc)
        pk = st.packed()
        hp.build_correlation(ws, ws.Qc, ws.Kc, M=st.num_modes, d=st.attention_mode_dim,
                             w_agg=pk.get("w_agg", 0.0), table=self.vispos_encoder.table(),
                             w_pos=st.pos_code_weight, global_norm=self.do_corr_global_norm,
                             attn_clip=st.attn_clip, diag=st.diag(ws.Qc.device))
        self._ws = ws

    @ops.on_device
    def update(self, fmap1, fmap2, fmap1o, fmap2o, coords1, coords2=None):
        if fmap1o is not None and fmap2o is not None:
            raise NotImplementedError("two-way correlation (--f1 shared/private) is an ablation outside the hot path")
        _require_inference(self.setrans.query.weight)
        B, Cc, h, w = fmap1.shape
        if B != 1:
            raise NotImplementedError("standalone TransCorrBlock.update handles one pair per call")
        grid = TokenGrid(h, w)
        ws = get_workspace(grid, fmap1.device, "SAVECORR" in os.environ)
        ops.pack_tokens(fmap1[0].float().contiguous(), grid, ops.PACK_LN, out_b=ws.T1)
        ops.pack_tokens(fmap2[0].float().contiguous(), grid, ops.PACK_LN, out_b=ws.T2f)
        self.build_rows(ws, ws.T1, ws.T2f)
        if "SAVECORR" in os.environ:
            self._save_volume(ws, os.environ["SAVECORR"])

    def _save_volume(self, ws, path):
        """SAVECORR hook core/corr.py:180-184 -- needs the materialised, normalised level-0 volume."""
        g = ws.grid
        if ws.levels[0] is None:
            raise RuntimeError("SAVECORR needs materialize_level0=True")
        mean, rstd = ws.mean_rstd.tolist()
        vol = ws.levels[0].view(g.H, g.Wp, g.H, g.W)[:, :g.W]
        torch.save(((vol - mean) * rstd).reshape(1, g.H, g.W, g.H, g.W).cpu(), path)
