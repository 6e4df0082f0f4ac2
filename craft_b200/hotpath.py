"""Fused execution of the CRAFT hot path on token-row buffers (one image pair at a time).

This is the host-side schedule: which kernel runs on which buffer in which order.  All arithmetic
happens in libcraft_b200.so (craft_b200/ops.py); torch only owns the device memory.  The nn.Module
classes (network.py, setrans.py, corr.py, gma.py, update.py) hold the parameters under the
reference's names and call into this file, so `CRAFT.forward` and the per-module `forward()`s run
exactly the same kernels.

Buffer layout: every activation is a [Mp, ld] row-major matrix over the
padded-flat token grid; the update block's GRU input lives in ONE 640-column bf16 matrix
    X = [ h (0:128) | inp (128:256) | motion (256:384) | motion_global (384:512) | r*h (512:640) ]
so that torch.cat never happens: each producer kernel writes its column slice in place.
"""
import math
import os

import torch

from . import ops
from .ops import TokenGrid

_INF = float("inf")


class PackedWeights:
    """bf16 / re-laid-out copies of module parameters, rebuilt when a parameter changes."""

    def __init__(self):
        self._store = {}

    def get(self, key, params, build):
        ver = (ops.act_dtype(),) + tuple((p.data_ptr(), p._version, p.device) for p in params)
        # nn.DataParallel replicas share this object (replicate() shallow-copies module __dict__s): one entry per device
        key = (key, str(params[0].device) if params else None)
        hit = self._store.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                hit = (ver, build())
            self._store[key] = hit
        return hit[1]


def level0_mode(x=None):
    """How level 0 of the correlation pyramid is held:
      "h16"      (default) the build kernel stores it in 8x8-key-block order as fp16 deltas against the fp32 mean
                 of each 4x8 half block (110 MB at 448x1024) and ONE lookup kernel serves all four levels;
      "ondemand" never stored: each lookup recomputes its 10x10 window from the projected Q/K rows;
      "f32"      the reference's dense fp32 [U,U] volume (SAVECORR / debugging / standalone scores-only calls).
    True -> "f32"; None / False -> $CRAFT_B200_LEVEL0 or "h16"."""
    if x is True:
        return "f32"
    if x is None or x is False:
        x = os.environ.get("CRAFT_B200_LEVEL0", "h16")
    if x not in ("h16", "ondemand", "f32"):
        raise ValueError("level-0 mode must be h16, ondemand or f32 (got %r)" % (x,))
    return x


class Workspace:
    """All device buffers for one (H, W) token grid on one device.  Allocated once and reused, so the
    addresses baked into TMA descriptors stay valid and the whole schedule can live in a CUDA graph."""

    def __init__(self, grid, device, materialize_level0=None):
        g = self.grid = grid
        self.device = device
        self.level0 = level0_mode(materialize_level0)
        act = self.dtype = ops.act_dtype()
        z = lambda cols, dt=act: torch.zeros((g.Mp, cols), dtype=dt, device=device)
        f32 = torch.float32
        self.ldv = (((g.H + 7) // 8) * 8) * (((g.W + 15) // 16) * 16)     # keys in 8x16 (or 8x8) block order
        # --- feature tokens / projections
        self.T1 = z(256); self.T2 = z(256); self.T2f = z(256)
        self.Qc = z(256); self.Kc = z(256)                     # corr_fn projections
        self.Q2 = z(256); self.K2 = z(256)                     # f2_trans projections
        self.Ta = z(128); self.Qa = z(128); self.Ka = z(128)   # intra-frame attention
        self.Vt = torch.zeros((1024, self.ldv), dtype=act, device=device)
        self.ks_sc = ops.scores_auto_ksplit(g)
        self.lse_part = torch.zeros((self.ks_sc, 4, g.Mp, 2), dtype=f32, device=device)
        self.lse2_f2 = torch.zeros((4, g.Mp), dtype=f32, device=device)
        self.lse2_att = torch.zeros((4, g.Mp), dtype=f32, device=device)
        # P.V partial sums: sized ONCE for the largest consumer this grid can see (a captured CUDA graph bakes
        # the address in, so the buffer must never be reallocated): F2 transformer M=4 x F=256, the motion
        # aggregator M=4 x F=128, GMA M=1 x F=128.
        self.ks_pv = {M: ops.pv_auto_ksplit(g, M) for M in (1, 2, 4)}
        n = max(self.ks_pv[4] * 4 * 256, self.ks_pv[2] * 2 * 256, self.ks_pv[1] * 256) * g.Mp
        self.Opart = torch.zeros((n,), dtype=f32, device=device)
        # --- scalars
        self.clip_corr = torch.full((1,), _INF, dtype=f32, device=device)
        self.clip_f2 = torch.full((1,), _INF, dtype=f32, device=device)
        self.clip_att = torch.full((1,), _INF, dtype=f32, device=device)
        self.flag = torch.zeros((3,), dtype=torch.int32, device=device)
        self.stat_max = torch.zeros((4,), dtype=f32, device=device)
        self.stat_sum = torch.zeros((2, 2), dtype=torch.float64, device=device)
        self.mean_rstd = torch.zeros((2,), dtype=f32, device=device)
        self.identity_stats = torch.tensor([0.0, 1.0], dtype=f32, device=device)
        self.inf_clip = torch.full((1,), _INF, dtype=f32, device=device)
        # --- correlation pyramid (level 0 only when asked for)
        shapes = g.level_shapes()
        self.levels = [None] * 4
        for l, (h, w) in enumerate(shapes):
            if l == 0 and self.level0 != "f32":
                continue
            self.levels[l] = torch.zeros((g.Mp, max(h * w, 1)), dtype=f32, device=device)
        self.level0_h16 = None      # fp16 level 0, allocated by the first build_correlation (standalone attention
                                    # calls on the shared workspace never need its 100 MB)
        # --- update block
        self.X = z(640)
        self.Hm = z(128, f32); self.Z = z(128, f32)
        self.CORR = z(384); self.C1 = z(256); self.CF = z(256); self.F1 = z(128)
        self.HD = z(512)
        self.DELTA = z(32, f32)
        # two mask buffers alternate by iteration so that upsample(it-1) on the side stream can overlap
        # the mask head of iteration it
        self.MASKS = [z(576, f32), z(576, f32)]
        self.MASK = self.MASKS[0]
        self.side = torch.cuda.Stream(device=device)    # second lane for independent branches (captured too)
        self.ev_lookup = torch.cuda.Event()
        self.coords1 = z(2, f32); self.flow = z(2, f32)

    def ensure_level0(self):
        if self.level0 == "h16" and self.level0_h16 is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the fp16 level-0 volume must be allocated before CUDA-graph capture (run one eager forward)")
            g = self.grid
            nblk = ((g.H + 7) // 8) * ((g.W + 7) // 8)
            # [Mp][nblk][64] fp16 deltas, then [Mp][nblk][2] fp32 half-block means (include/craft_b200.h)
            self.level0_h16 = torch.zeros((g.Mp * nblk * 68,), dtype=torch.float16, device=self.device)
        return self.level0_h16

    def opart(self, ks, M, F):
        if ks * M * self.grid.Mp * F > self.Opart.numel():
            raise ValueError("P.V partial-sum buffer is sized for M*F <= 1024 (got ks=%d, M=%d, F=%d)" % (ks, M, F))
        return self.Opart

    def pv_split(self, M):
        return self.ks_pv[M]


# ------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------
def _bn_for(npad):
    for bn in (128, 64, 32):
        if npad % bn == 0:
            return bn
    raise ValueError("N=%d is not a multiple of 32" % npad)


def project(grid, A, Wp, bias, out_b, *, K, a_koff=0, Npad=None):
    """out = A[:, a_koff:a_koff+K] @ Wp^T + bias  (nn.Linear on token rows; core/setrans.py:507-508)."""
    Npad = Npad or Wp.shape[0]
    ops.shift_gemm(A, Wp, M=grid.Mp, Npad=Npad, K=K, BN=_bn_for(Npad), a_koff=a_koff, grid=grid, bias=bias,
                   out_b=out_b)


def attention_stats(ws, Q, K, *, M, d, table, w_pos, clip, lse2, slot, attn_clip=100.0, diag=None, mask_radius=-1):
    """Softmax statistics with the reference's data-dependent clamp (core/setrans.py:520-529):
    pass 1 runs unclamped and records the global max; the gate kernel arms `clip`; pass 2 is a
    device-side no-op unless the gate fired.  No host synchronisation.  `diag`: the owning module's
    device-side {max_attn, clamp_count}."""
    g = ws.grid
    clip.fill_(_INF)
    smax = ws.stat_max[slot:slot + 1]
    smax.fill_(-_INF)
    flag = ws.flag[slot:slot + 1]
    ops.attn_lse(Q, K, g, M=M, d=d, w_pos=w_pos, pos_table=table, R=7, clip=clip, stat_max=smax,
                 lse_part=ws.lse_part, lse2=lse2, ksplit=ws.ks_sc, mask_radius=mask_radius)
    ops.clip_gate(smax, attn_clip, clip, flag, diag)
    ops.attn_lse(Q, K, g, M=M, d=d, w_pos=w_pos, pos_table=table, R=7, clip=clip, stat_max=ws.stat_max[3:4],
                 lse_part=ws.lse_part, lse2=lse2, ksplit=ws.ks_sc, run_flag=flag, mask_radius=mask_radius)


def value_aggregate(ws, Q, K, X, x_koff, W1p, *, M, d, F, table, w_pos, clip, lse2, w_score, b_score, coeff,
                    out_b=None, colb=0, out_f=None, colf=0, gma=0, mask_radius=-1):
    """ExpandedFeatTrans.forward core/setrans.py:364-410 without ever forming P:
    V^T = W1 X^T (tcgen05 GEMM) -> flash P.V per mode -> mode soft-pool + skip + LayerNorm."""
    g = ws.grid
    C_in = W1p.shape[1]
    BK = ops.pv_block_keys(d, F)
    ops.shift_gemm(W1p, X, M=M * F, Npad=ops.blocked_keys(g, BK), K=C_in, BN=BK, b_koff=x_koff, out_b=ws.Vt,
                   b_block_grid=g)
    ks = ws.pv_split(M)
    O = ws.opart(ks, M, F)
    ops.attn_pv(Q, K, ws.Vt, g, M=M, d=d, F=F, w_pos=w_pos, pos_table=table, R=7, clip=clip, lse2=lse2,
                out=O, ksplit=ks, zero_fill=False, mask_radius=mask_radius)
    ops.modes_finalize(O, ks, M, F, g, w_score=w_score, b_score=b_score, coeff=coeff, gma=gma, x_b=X, colx=x_koff,
                       out_b=out_b, colb=colb, out_f=out_f, colf=colf, pv_bk=BK)


def build_correlation(ws, Q, K, *, M, d, w_agg, table, w_pos, global_norm, attn_clip=100.0, diag=None):
    """TransCorrBlock.update (core/corr.py:148-207) / CorrBlock.__init__ (:16-45): pyramid + LN stats."""
    g = ws.grid
    # what the on-demand level-0 lookup needs to recompute volume cells (ops.corr_lookup0)
    ws.corr_meta = dict(Q=Q, K=K, M=M, d=d, w_agg=w_agg, w_pos=w_pos, pos_table=table, R=7, clip=ws.clip_corr)
    ws.stat_sum.zero_()
    clip = ws.clip_corr
    clip.fill_(_INF)
    smax = ws.stat_max[0:1]
    smax.fill_(-_INF)
    flag = ws.flag[0:1]
    kw = dict(M=M, d=d, w_agg=w_agg, w_pos=w_pos, pos_table=table, R=7, clip=clip, levels=ws.levels,
              ksplit=ws.ks_sc, level0_h16=ws.ensure_level0())
    ops.corr_build(Q, K, g, stat_sum=ws.stat_sum[0], stat_max=smax, **kw)
    if M > 1:   # the clamp exists only on the transformer path
        ops.clip_gate(smax, attn_clip, clip, flag, diag)
        ops.corr_build(Q, K, g, stat_sum=ws.stat_sum[1], stat_max=ws.stat_max[3:4], run_flag=flag, **kw)
        gate = flag
    else:
        gate = None
    if global_norm:
        ops.corr_stats_finalize(ws.stat_sum, float(g.U) * float(g.U), ws.mean_rstd, flag=gate)
    else:
        ws.mean_rstd.copy_(ws.identity_stats)


# ------------------------------------------------------------------------------------------------
# update block (one refinement iteration)
# ------------------------------------------------------------------------------------------------
def pack_gru(gru, grid):
    """[(w_zr, b_zr, w_q, b_q, taps)] for the 1x5 and the 5x1 pass of SepConvGRU (core/update.py:49-64)."""
    pc = ops.pack_conv_weight
    perm = list(range(128, 512)) + list(range(0, 128))      # q conv reads X[:, 128:640] = [x | r*h]
    out = []
    for tag in ("1", "2"):
        cz, cr, cq = (getattr(gru, "conv%s%s" % (n, tag)) for n in "zrq")
        wzr = pc(torch.cat([cz.weight, cr.weight], 0))
        bzr = torch.cat([cz.bias, cr.bias]).detach().float().contiguous()
        wq = pc(cq.weight, cin_perm=perm)
        bq = cq.bias.detach().float().contiguous()
        kh, kw = cz.weight.shape[2:]
        out.append((wzr, bzr, wq, bq, ops.conv_taps(kh, kw, grid)))
    return out


class EncoderWeights:
    """Packed parameters of BasicMotionEncoder (core/update.py:67-87)."""

    def __init__(self, enc, grid):
        pc, pb = ops.pack_conv_weight, ops.pad_bias
        self.c1_w = pc(enc.convc1.weight, Kpad=384); self.c1_b = pb(enc.convc1.bias, 256)
        self.c2_w = pc(enc.convc2.weight); self.c2_b = pb(enc.convc2.bias, 192)
        self.f1_w = enc.convf1.weight.detach().float().permute(1, 2, 3, 0).reshape(98, 128).contiguous()
        self.f1_b = enc.convf1.bias.detach().float().contiguous()
        self.f2_w = pc(enc.convf2.weight); self.f2_b = pb(enc.convf2.bias, 64)
        self.cv_w = pc(enc.conv.weight, Npad=128); self.cv_b = pb(enc.conv.bias, 128)
        self.taps3 = ops.conv_taps(3, 3, grid)


class EncoderBuffers:
    """The slice of a Workspace that motion_encoder touches (standalone BasicMotionEncoder.forward)."""

    def __init__(self, grid, device):
        self.grid = grid
        z = lambda cols, dt=ops.act_dtype(): torch.zeros((grid.Mp, cols), dtype=dt, device=device)
        self.X = z(640)
        self.CORR = z(384); self.C1 = z(256); self.CF = z(256); self.F1 = z(128)
        self.flow = z(2, torch.float32)
        self.side = torch.cuda.Stream(device=device)
        self.ev_lookup = torch.cuda.Event()


class UpdateWeights(EncoderWeights):
    """Packed parameters of GMAUpdateBlock (core/update.py:116-162)."""

    def __init__(self, ub, grid):
        super().__init__(ub.encoder, grid)
        dev = ub.encoder.convc1.weight.device
        pc, pb = ops.pack_conv_weight, ops.pad_bias
        self.gru = pack_gru(ub.gru, grid)
        fh, mk = ub.flow_head, ub.mask
        self.hd_w = pc(torch.cat([fh.conv1.weight, mk[0].weight], 0))
        self.hd_b = torch.cat([fh.conv1.bias, mk[0].bias]).detach().float().contiguous()
        self.hdf_w = pc(fh.conv1.weight); self.hdf_b = fh.conv1.bias.detach().float().contiguous()   # flow head alone
        self.fl_w = pc(fh.conv2.weight, Npad=32); self.fl_b = pb(fh.conv2.bias, 32)
        self.mk_w = pc(mk[2].weight); self.mk_b = (0.25 * mk[2].bias.detach().float()).contiguous()
        self.device = dev


def motion_encoder(ws, uw, lookup=None):
    """BasicMotionEncoder.forward core/update.py:79-87 -> X[:, 256:384] (126 features + the flow).

    `lookup(part)` (optional) is the correlation lookup of this iteration (core/corr.py:47-71).  The flow
    branch (convf1 -> convf2) depends only on the current flow and runs on the side stream next to the
    lookup and the correlation branch (convc1 -> convc2).  When level 0 is computed on demand the lookup
    is two kernels: the level-0 part runs on the main stream, the pooled levels on the side stream."""
    g = ws.grid
    sg = ops.shift_gemm
    main, side = torch.cuda.current_stream(), ws.side
    split = lookup is not None and getattr(ws, "level0", "f32") == "ondemand"
    side.wait_stream(main)
    with torch.cuda.stream(side):
        if split:
            lookup("pooled")
            ws.ev_lookup.record(side)
        ops.convf1(ws.flow, uw.f1_w, uw.f1_b, g, ws.F1)
        sg(ws.F1, uw.f2_w, M=g.Mp, Npad=64, K=128, BN=64, taps=uw.taps3, grid=g, bias=uw.f2_b, act=1, out_b=ws.CF,
           colb=192)
    if split:
        lookup("level0")
        main.wait_event(ws.ev_lookup)
    elif lookup is not None:
        lookup("all")
    sg(ws.CORR, uw.c1_w, M=g.Mp, Npad=256, K=384, BN=128, grid=g, bias=uw.c1_b, act=1, out_b=ws.C1)
    # N = 192 as two 96-wide tiles: 114 CTAs = one wave (three 64-wide tiles would need 171 CTAs = two waves)
    sg(ws.C1, uw.c2_w, M=g.Mp, Npad=192, K=256, BN=96, taps=uw.taps3, grid=g, bias=uw.c2_b, act=1, out_b=ws.CF)
    main.wait_stream(side)
    sg(ws.CF, uw.cv_w, M=g.Mp, Npad=128, K=256, BN=64, taps=uw.taps3, grid=g, bias=uw.cv_b,
       epilogue=ops.EPI_MOTION, out_b=ws.X, colb=256, aux1=ws.flow)


def sep_conv_gru_rows(g, X, Hm, Z, passes):
    """SepConvGRU.forward core/update.py:49-64 on X = [h | x | r*h] (in place; fp32 master state in Hm)."""
    for (wzr, bzr, wq, bq, taps) in passes:
        ops.shift_gemm(X, wzr, M=g.Mp, Npad=256, K=512, BN=128, taps=taps, grid=g, epilogue=ops.EPI_GRU_ZR,
                       bias=bzr, out_b=X, colb=512, aux0=Z, aux1=Hm)
        ops.shift_gemm(X, wq, M=g.Mp, Npad=128, K=512, BN=64, taps=taps, a_koff=128, grid=g,
                       epilogue=ops.EPI_GRU_Q, bias=bq, out_b=X, colb=0, aux0=Z, aux1=Hm)


def sep_conv_gru(ws, uw):
    sep_conv_gru_rows(ws.grid, ws.X, ws.Hm, ws.Z, uw.gru)


def heads(ws, uw, it=0, need_mask=True, update_flow=False):
    """FlowHead core/update.py:15-16 + mask head core/update.py:124-127,161 -> DELTA[:, :2], MASKS[it & 1].
    The two second-layer convolutions are independent: the flow one runs on the side stream.
    need_mask=False computes the flow head only (the caller discards this iteration's mask).
    update_flow=True: the flow head's epilogue also applies coords1 += delta, flow = coords1 - coords0
    (core/network.py:247,236) instead of a separate kernel."""
    g = ws.grid
    sg = ops.shift_gemm
    fl = dict(M=g.Mp, Npad=32, K=256, BN=32, taps=uw.taps3, grid=g, bias=uw.fl_b, out_f=ws.DELTA)
    if update_flow:
        fl.update(epilogue=ops.EPI_FLOW, aux0=ws.coords1, aux1=ws.flow)
    if not need_mask:
        sg(ws.X, uw.hdf_w, M=g.Mp, Npad=256, K=128, BN=128, taps=uw.taps3, grid=g, bias=uw.hdf_b, act=1, out_b=ws.HD)
        sg(ws.HD, uw.fl_w, **fl)
        return
    sg(ws.X, uw.hd_w, M=g.Mp, Npad=512, K=128, BN=128, taps=uw.taps3, grid=g, bias=uw.hd_b, act=1, out_b=ws.HD)
    main, side = torch.cuda.current_stream(), ws.side
    side.wait_stream(main)
    with torch.cuda.stream(side):
        sg(ws.HD, uw.fl_w, **fl)
    sg(ws.HD, uw.mk_w, M=g.Mp, Npad=576, K=256, BN=64, a_koff=256, grid=g, bias=uw.mk_b, alpha=0.25,
       out_f=ws.MASKS[it & 1])
    main.wait_stream(side)
