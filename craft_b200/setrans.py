"""Squeeze-expansion transformer modules with the reference's class names, constructor arguments,
parameter names and forward() signatures (core/setrans.py), executed by the sm_100a kernels.

What differs from the reference, by design:
  * attention matrices [B,M,U,U] and dense positional-bias tensors [1,1,U,U] are never built.
    Where the reference hands such a tensor from one module to the next, these modules hand over an
    opaque handle (`PosBiasHandle`, `AttentionHandle`) that the consuming module understands.
  * only the configuration space the CRAFT drivers can reach is implemented: pos_code_type 'bias',
    has_FFN False, pool_modes_feat 'softmax', no multi-head ablation.  Anything else raises.
  * the fused kernels are forward-only: dropout / drop-path are identity (eval semantics); calling
    them with grad enabled on parameters that require grad raises (training: CRAFT.forward switches
    to craft_b200/train_path.py).
"""
import collections
import copy
import math

import torch
import torch.nn as nn
from torch.nn import Parameter

from . import hotpath as hp
from . import ops
from .ops import TokenGrid


def print0(*a, **k):
    import os
    if int(os.environ.get("LOCAL_RANK", 0)) == 0 and os.environ.get("CRAFT_B200_VERBOSE"):
        print(*a, **k)


def gen_all_indices(shape, device):
    """core/setrans.py:32-39 -- [*shape, len(shape)] integer coordinates."""
    grids = torch.meshgrid(*[torch.arange(s, device=device) for s in shape], indexing="ij")
    return torch.stack(grids, dim=len(shape))


class SETransConfig(object):
    """Attribute-compatible with core/setrans.py:71-157."""

    def __init__(self):
        self.feat_dim = -1
        self.in_feat_dim = -1
        self.pos_dim = 2
        self.pos_code_weight = 1
        self.num_modes = 4
        self.tie_qk_scheme = "shared"
        self.trans_output_type = "private"
        self.act_fun = torch.nn.functional.gelu
        self.attn_clip = 100
        self.attn_diag_cycles = 1000
        self.base_initializer_range = 0.02
        self.qk_have_bias = False
        self.v_has_bias = False
        self.query_idbias_scale = 10
        self.feattrans_lin1_idbias_scale = 10
        self.pool_modes_feat = "softmax"
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.2
        self.drop_path_prob = 0
        self.pos_code_type = "bias"
        self.ablate_multihead = False
        self.out_attn_probs_only = False
        self.out_attn_scores_only = False
        self.attn_mask_radius = -1

    def try_assign(self, args, *keys):
        src = args if isinstance(args, dict) else args.__dict__
        hit = False
        for k in keys:
            if k in src:
                self.__dict__[k] = src[k]
                hit = True
        return hit

    def update_config(self, args):
        self.try_assign(args, "use_pretrained", "apply_attn_stage", "num_modes", "trans_output_type",
                        "base_initializer_range", "pos_code_type", "ablate_multihead", "attn_clip",
                        "attn_diag_cycles", "tie_qk_scheme", "feattrans_lin1_idbias_scale", "qk_have_bias",
                        "v_has_bias", "out_attn_probs_only", "out_attn_scores_only", "in_feat_dim",
                        "pos_bias_radius")
        if self.try_assign(args, "out_feat_dim"):
            self.feat_dim = self.out_feat_dim
        else:
            self.feat_dim = self.in_feat_dim
        src = args if isinstance(args, dict) else args.__dict__
        if "dropout_prob" in src and src["dropout_prob"] >= 0:
            self.hidden_dropout_prob = src["dropout_prob"]
            self.attention_probs_dropout_prob = src["dropout_prob"]


def _require_inference(*params):
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        raise RuntimeError("craft_b200 kernels are forward-only in this round: wrap the call in torch.no_grad() "
                           "(training/backward is listed as out of scope in DESIGN.md)")


# ------------------------------------------------------------------------------------------------
# opaque handles
# ------------------------------------------------------------------------------------------------
class PosBiasHandle:
    """Stands for the dense [1,1,U,U] tensor of SlidingPosBiases2D.forward (core/setrans.py:690-708):
    bias[(y1,x1),(y2,x2)] = table[y2-y1+R, x2-x1+R] inside the window.  Kernels evaluate it on the fly."""

    def __init__(self, table, shape):
        self.table = table
        self.shape = tuple(shape)


class AttentionHandle:
    """Stands for the [B,M,U,U] attention-probability tensor: projected Q/K rows plus the softmax
    log-sum-exp; the aggregator recomputes P tile by tile (attn_pv.cuh)."""

    def __init__(self, grid, Q, K, lse2, clip, table, w_pos, M, d, mask_radius=-1):
        self.grid, self.Q, self.K, self.lse2, self.clip = grid, Q, K, lse2, clip
        self.table, self.w_pos, self.M, self.d = table, w_pos, M, d
        self.mask_radius = mask_radius
        self.shape = (1, M, grid.U, grid.U)

    def dense(self):
        """Materialise the [1,M,U,U] probabilities (small grids only; debugging and tests)."""
        with torch.cuda.device(self.Q.device):
            return ops.attn_dense(self.Q, self.K, self.grid, M=self.M, d=self.d, w_pos=self.w_pos, pos_table=self.table,
                                  R=7, clip=self.clip, lse2=self.lse2, mask_radius=self.mask_radius)[None]


# ------------------------------------------------------------------------------------------------
class LearnedSoftAggregate(nn.Module):
    """core/setrans.py:279-300.  Inside CRAFT.forward the arithmetic is fused into kernel epilogues
    (scores.cuh for num_feat == 1, modes_finalize for num_feat == F); forward() is the standalone form."""

    def __init__(self, num_feat, group_dim, keepdim=False):
        super().__init__()
        self.group_dim, self.num_feat, self.keepdim = group_dim, num_feat, keepdim
        self.feat2score = nn.Linear(num_feat, 1)

    @ops.on_device
    def forward(self, x, score_basis=None):
        """x [..., M at group_dim, ...(, F)] -> softmax-over-modes weighted sum (core/setrans.py:289-300)."""
        _require_inference(self.feat2score.weight)
        gd = self.group_dim % x.dim()
        if self.num_feat > 1 and gd == x.dim() - 1:
            raise ValueError("group_dim cannot be the feature axis")
        xm = x.float().movedim(gd, 0).contiguous()
        bm = score_basis.float().movedim(gd, 0).contiguous() if score_basis is not None else None
        if xm.shape[0] > 8:
            raise NotImplementedError("craft_b200 LearnedSoftAggregate: at most 8 modes")
        w = self.feat2score.weight.detach().float().reshape(-1).contiguous()
        b = self.feat2score.bias.detach().float().contiguous()
        out = ops.soft_aggregate(xm, w, b, basis=bm, num_feat=self.num_feat)
        return out.unsqueeze(gd) if self.keepdim else out


class SlidingPosBiases2D(nn.Module):
    """core/setrans.py:644-708 -- holds the learnable (2R+1)^2 table.  The reference's four index
    buffers are non-persistent, so the state dict is identical without them."""

    def __init__(self, pos_dim=2, pos_bias_radius=7, max_pos_size=(200, 200)):
        super().__init__()
        assert pos_dim == 2
        self.pos_dim, self.R = pos_dim, pos_bias_radius
        self.biases = Parameter(torch.zeros([2 * pos_bias_radius + 1] * pos_dim))

    def forward(self, feat_shape, device=None):
        return PosBiasHandle(self.biases, feat_shape[-2:])


class SETransInputFeatEncoder(nn.Module):
    """core/setrans.py:710-800 for pos_code_type == 'bias'."""

    def __init__(self, config):
        super().__init__()
        if config.pos_code_type != "bias":
            raise NotImplementedError("craft_b200 implements pos_code_type='bias' only (what CRAFT configures)")
        self.feat_dim = config.in_feat_dim
        self.pos_code_type = config.pos_code_type
        self.pos_code_weight = 0
        self.pos_coder = SlidingPosBiases2D(config.pos_dim, config.pos_bias_radius)

    def table(self):
        return self.pos_coder.biases.detach().float().contiguous()

    @ops.on_device
    def forward(self, vis_feat, voxels_pos=None, return_pos_biases=True):
        """[B,C,h,w] -> [B,h*w,C] LayerNorm'ed tokens (+ PosBiasHandle)."""
        _require_inference()
        B, Cc, h, w = vis_feat.shape
        grid = TokenGrid(h, w)
        out = torch.empty((B, h * w, Cc), dtype=torch.float32, device=vis_feat.device)
        buf = torch.zeros((grid.Mp, Cc), dtype=torch.float32, device=vis_feat.device)
        for b in range(B):
            ops.pack_tokens(vis_feat[b].float().contiguous(), grid, ops.PACK_LN, out_f=buf)
            out[b] = buf.view(h, grid.Wp, Cc)[:, :w].reshape(h * w, Cc)
        if return_pos_biases:
            return out, self.pos_coder(vis_feat.shape)
        return out


class ExpandedFeatTrans(nn.Module):
    """core/setrans.py:304-410 (has_FFN False path)."""

    def __init__(self, config, name):
        super().__init__()
        self.config, self.name = config, name
        self.in_feat_dim, self.feat_dim, self.num_modes = config.in_feat_dim, config.feat_dim, config.num_modes
        self.feat_dim_allmode = self.feat_dim * self.num_modes
        self.first_linear = nn.Linear(self.in_feat_dim, self.feat_dim_allmode, bias=config.v_has_bias)
        self.base_initializer_range = config.base_initializer_range
        self.has_FFN = getattr(config, "has_FFN", True)
        self.has_input_skip = getattr(config, "has_input_skip", False)
        if self.has_FFN or config.v_has_bias or config.pool_modes_feat != "softmax" or not self.has_input_skip \
                or config.drop_path_prob > 0:
            raise NotImplementedError("craft_b200: ExpandedFeatTrans supports the CRAFT configuration only "
                                      "(no FFN, no V bias, softmax mode pooling, input skip)")
        self.pool_modes_feat = config.pool_modes_feat
        self.feat_softaggr = LearnedSoftAggregate(self.feat_dim, group_dim=1, keepdim=False)
        self.input_skip_coeff = Parameter(torch.ones(1))
        self._packed = hp.PackedWeights()

    def add_identity_bias(self):
        if self.config.feattrans_lin1_idbias_scale > 0:
            F_ = self.feat_dim
            eye = torch.eye(F_) * self.base_initializer_range * self.config.feattrans_lin1_idbias_scale
            w = self.first_linear.weight.data
            w[:F_, :F_] = w[:F_, :F_] * 0.5 + eye

    def packed(self):
        def build():
            return dict(w1=ops.pack_linear_weight(self.first_linear.weight),
                        ws=self.feat_softaggr.feat2score.weight.detach().float().reshape(-1).contiguous(),
                        bs=self.feat_softaggr.feat2score.bias.detach().float().contiguous(),
                        coeff=self.input_skip_coeff.detach().float().contiguous())
        return self._packed.get("w", [self.first_linear.weight, self.feat_softaggr.feat2score.weight,
                                      self.feat_softaggr.feat2score.bias, self.input_skip_coeff], build)

    def run(self, ws, att, X, x_koff, out_b=None, colb=0, out_f=None, colf=0):
        """Fused path on token rows: X[:, x_koff:x_koff+C] is `input_feat`."""
        pk = self.packed()
        hp.value_aggregate(ws, att.Q, att.K, X, x_koff, pk["w1"], M=att.M, d=att.d, F=self.feat_dim,
                           table=att.table, w_pos=att.w_pos, clip=att.clip, lse2=att.lse2,
                           w_score=pk["ws"], b_score=pk["bs"], coeff=pk["coeff"],
                           out_b=out_b, colb=colb, out_f=out_f, colf=colf, mask_radius=att.mask_radius)

    @ops.on_device
    def forward(self, input_feat, attention_probs):
        """input_feat [B,U,C]; attention_probs: AttentionHandle from CrossAttFeatTrans/SelfAttVisPosTrans."""
        if not isinstance(attention_probs, (AttentionHandle, list)):
            raise TypeError("craft_b200.ExpandedFeatTrans takes the AttentionHandle produced by the attention "
                            "module, not a dense [B,M,U,U] tensor (which this implementation never builds)")
        _require_inference(self.first_linear.weight)
        handles = attention_probs if isinstance(attention_probs, list) else [attention_probs]
        B, U, Cc = input_feat.shape
        grid = handles[0].grid
        ws = get_workspace(grid, input_feat.device)
        out = torch.empty((B, U, self.feat_dim), dtype=torch.float32, device=input_feat.device)
        xb = torch.zeros((grid.Mp, Cc), dtype=ops.act_dtype(), device=input_feat.device)
        yf = torch.zeros((grid.Mp, self.feat_dim), dtype=torch.float32, device=input_feat.device)
        for b in range(B):
            xb.view(grid.H, grid.Wp, Cc)[:, :grid.W] = input_feat[b].reshape(grid.H, grid.W, Cc).to(xb.dtype)
            self.run(ws, handles[b], xb, 0, out_f=yf)
            out[b] = yf.view(grid.H, grid.Wp, -1)[:, :grid.W].reshape(U, -1)
        return out


class CrossAttFeatTrans(nn.Module):
    """core/setrans.py:412-566."""

    def __init__(self, config, name):
        super().__init__()
        self.config, self.name = config, name
        self.num_modes = config.num_modes
        self.in_feat_dim, self.feat_dim = config.in_feat_dim, config.feat_dim
        self.attention_mode_dim = self.in_feat_dim // self.num_modes
        self.att_size_allmode = self.num_modes * self.attention_mode_dim
        self.query = nn.Linear(self.in_feat_dim, self.att_size_allmode, bias=config.qk_have_bias)
        self.key = nn.Linear(self.in_feat_dim, self.att_size_allmode, bias=config.qk_have_bias)
        self.base_initializer_range = config.base_initializer_range
        self.out_attn_scores_only = config.out_attn_scores_only
        self.out_attn_probs_only = config.out_attn_probs_only
        if config.ablate_multihead:
            raise NotImplementedError("craft_b200: ablate_multihead is an ablation outside the hot path")
        if self.out_attn_scores_only or self.out_attn_probs_only:
            self.out_trans = None
            if self.num_modes > 1:
                self.attn_softaggr = LearnedSoftAggregate(1, group_dim=1, keepdim=True)
        else:
            self.out_trans = ExpandedFeatTrans(config, name + "-out_trans")
        self.tie_qk_scheme = config.tie_qk_scheme
        self.pos_code_weight = config.pos_code_weight if config.pos_code_type == "bias" else 1
        self.attn_clip = config.attn_clip
        self.attn_diag_cycles = getattr(config, "attn_diag_cycles", 1000)
        self.call_count = 0
        self._diag = {}          # device -> f32[2] {max_attn, clamp_count}, updated by the gate kernel
        self._diag_host = [0.0, 0]
        self._init_weights()
        self._packed = hp.PackedWeights()

    # The reference reads the global score maximum back with two .item() calls per forward
    # (core/setrans.py:520-529) to keep these counters; here they are accumulated on the device by the
    # clamp-gate kernel and only read when somebody asks.
    def diag(self, device):
        key = str(device)
        if key not in self._diag:
            self._diag[key] = torch.zeros(2, dtype=torch.float32, device=device)
        return self._diag[key]

    def _diag_read(self):
        mx, cnt = self._diag_host
        for t in self._diag.values():
            v = t.tolist()
            mx, cnt = max(mx, v[0]), cnt + int(v[1])
        return mx, cnt

    @property
    def max_attn(self):
        return self._diag_read()[0]

    @max_attn.setter
    def max_attn(self, v):          # the reference resets both counters to 0 every attn_diag_cycles calls
        for t in self._diag.values():
            t[0] = 0.0
        self._diag_host[0] = float(v)

    @property
    def clamp_count(self):
        return self._diag_read()[1]

    @clamp_count.setter
    def clamp_count(self, v):
        for t in self._diag.values():
            t[1] = 0.0
        self._diag_host[1] = int(v)

    def _init_weights(self):
        std = self.base_initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=std)
                if m.bias is not None:
                    m.bias.data.zero_()
        self.tie_qk()
        d = self.attention_mode_dim
        eye = (torch.eye(d) * std * self.config.query_idbias_scale).repeat(1, self.in_feat_dim // d)
        self.key.weight.data[:d] = self.key.weight.data[:d] * 0.5 + eye
        if self.out_trans is not None:
            self.out_trans.add_identity_bias()

    def tie_qk(self, tie_qk_scheme=None):
        if tie_qk_scheme is not None:
            self.tie_qk_scheme = tie_qk_scheme
        if self.tie_qk_scheme == "shared":
            self.key.weight = self.query.weight
            if self.key.bias is not None:
                self.key.bias = self.query.bias
        elif self.tie_qk_scheme == "loose":
            self.key.weight.data.copy_(self.query.weight)
            if self.key.bias is not None:
                self.key.bias.data.copy_(self.query.bias)

    def packed(self):
        def build():
            d = dict(wq=ops.pack_linear_weight(self.query.weight), wk=ops.pack_linear_weight(self.key.weight),
                     bq=None, bk=None)
            if self.query.bias is not None:
                d["bq"] = self.query.bias.detach().float().contiguous()
                d["bk"] = self.key.bias.detach().float().contiguous()
            if self.out_attn_scores_only and self.num_modes > 1:
                d["w_agg"] = float(self.attn_softaggr.feat2score.weight.detach().float().item())
            return d
        ps = [self.query.weight, self.key.weight]
        if self.query.bias is not None:
            ps += [self.query.bias, self.key.bias]
        if self.out_attn_scores_only and self.num_modes > 1:
            ps.append(self.attn_softaggr.feat2score.weight)
        return self._packed.get("qk", ps, build)

    def project(self, ws, Tq, Tk, Q, K):
        pk = self.packed()
        g = ws.grid
        hp.project(g, Tq, pk["wq"], pk["bq"], Q, K=self.in_feat_dim)
        hp.project(g, Tk, pk["wk"], pk["bk"], K, K=self.in_feat_dim)
        self.call_count += 1

    @ops.on_device
    def forward(self, query_feat, key_feat=None, pos_biases=None, attention_mask=None):
        """core/setrans.py:501-566 on token tensors [B,U,C] (already encoded by SETransInputFeatEncoder).

        `pos_biases` is the PosBiasHandle the encoder returned (it also carries the (h, w) of the token
        grid, which a [B,U,C] tensor alone does not); `attention_mask` is None or an int radius (the
        --f2radius Chebyshev mask of SelfAttVisPosTrans, core/setrans.py:580-584) -- dense mask tensors do
        not exist here.  Returns, per output mode:
          out_attn_scores_only -> dense [B,1,U,U] mode-aggregated scores (the level-0 volume before the
                                  global LayerNorm; fine for the grids a standalone call is used on),
          out_attn_probs_only  -> AttentionHandle (list for B > 1); `.dense()` materialises small ones,
          otherwise            -> [B,U,F] transformed features (out_trans)."""
        _require_inference(self.query.weight)
        if not isinstance(pos_biases, PosBiasHandle):
            raise TypeError("craft_b200.CrossAttFeatTrans.forward needs the PosBiasHandle returned by "
                            "SETransInputFeatEncoder.forward (it carries the token grid's (h, w))")
        if attention_mask is not None and not isinstance(attention_mask, int):
            raise TypeError("attention_mask must be None or the integer mask radius (dense masks are never built)")
        h, w = pos_biases.shape
        B, U, Cc = query_feat.shape
        if U != h * w or Cc != self.in_feat_dim:
            raise ValueError("query_feat %s does not match grid %dx%d / in_feat_dim %d" % (list(query_feat.shape), h, w, self.in_feat_dim))
        grid = TokenGrid(h, w)
        dev = query_feat.device
        ws = get_workspace(grid, dev, materialize_level0=self.out_attn_scores_only)
        table = pos_biases.table.detach().float().contiguous()
        mr = attention_mask if attention_mask is not None else -1
        M, d = self.num_modes, self.attention_mode_dim
        pk = self.packed()
        outs = []
        for b in range(B):
            Tq = tokens_to_rows(query_feat[b], grid)
            Tk = Tq if key_feat is None else tokens_to_rows(key_feat[b], grid)
            Q, K = grid.zeros(Cc, device=dev), grid.zeros(Cc, device=dev)
            self.project(ws, Tq, Tk, Q, K)
            if self.out_attn_scores_only:
                if mr > 0:
                    raise NotImplementedError("scores-only attention with a radius mask is not a CRAFT configuration")
                hp.build_correlation(ws, Q, K, M=M, d=d, w_agg=pk.get("w_agg", 0.0), table=table,
                                     w_pos=self.pos_code_weight, global_norm=False, attn_clip=self.attn_clip,
                                     diag=self.diag(dev))
                vol = ws.levels[0].view(grid.H, grid.Wp, U)[:, :grid.W].reshape(1, 1, U, U)
                outs.append(vol.clone())
                continue
            lse2 = torch.zeros((M, grid.Mp), dtype=torch.float32, device=dev)
            clip = torch.full((1,), float("inf"), dtype=torch.float32, device=dev)
            hp.attention_stats(ws, Q, K, M=M, d=d, table=table, w_pos=self.pos_code_weight, clip=clip, lse2=lse2,
                               slot=1, attn_clip=self.attn_clip, diag=self.diag(dev), mask_radius=mr)
            att = AttentionHandle(grid, Q, K, lse2, clip, table, self.pos_code_weight, M, d, mask_radius=mr)
            if self.out_attn_probs_only:
                outs.append(att)
                continue
            yf = torch.zeros((grid.Mp, self.feat_dim), dtype=torch.float32, device=dev)
            self.out_trans.run(ws, att, Tk, 0, out_f=yf)
            outs.append(rows_to_tokens(yf, grid)[None])
        if self.out_attn_probs_only:
            return outs[0] if B == 1 else outs
        return torch.cat(outs, 0)


def tokens_to_rows(tok, grid, dtype=None):
    """[U, C] token tensor -> padded-flat rows [Mp, C] (halo cells zero) in the operand dtype."""
    dtype = dtype or ops.act_dtype()
    Cc = tok.shape[-1]
    rows = torch.zeros((grid.H, grid.Wp, Cc), dtype=dtype, device=tok.device)
    rows[:, :grid.W] = tok.reshape(grid.H, grid.W, Cc).to(dtype)
    return rows.reshape(grid.Mp, Cc)


def rows_to_tokens(rows, grid):
    """padded-flat rows [Mp, C] -> [U, C] fp32."""
    Cc = rows.shape[-1]
    return rows.float().view(grid.H, grid.Wp, Cc)[:, :grid.W].reshape(grid.U, Cc)


class SelfAttVisPosTrans(nn.Module):
    """core/setrans.py:568-619."""

    def __init__(self, config, name):
        super().__init__()
        self.config = copy.copy(config)
        self.name = name
        self.out_attn_only = config.out_attn_scores_only or config.out_attn_probs_only
        self.attn_mask_radius = config.attn_mask_radius
        if self.attn_mask_radius > 0 and (config.in_feat_dim != 256 or config.num_modes != 4 or self.out_attn_only):
            raise NotImplementedError("craft_b200: the --f2radius mask is built for the F2 transformer (256 channels, "
                                      "4 modes, feature output), the only place the drivers can enable it")
        self.setrans = CrossAttFeatTrans(self.config, name)
        self.vispos_encoder = SETransInputFeatEncoder(self.config)

    def attend(self, ws, feat_chw, T, Q, K, lse2, clip, slot, pack_mode=ops.PACK_LN):
        """tokens -> projections -> softmax statistics; returns the AttentionHandle."""
        g = ws.grid
        st = self.setrans
        ops.pack_tokens(feat_chw, g, pack_mode, out_b=T)
        st.project(ws, T, T, Q, K)
        table = self.vispos_encoder.table()
        mr = self.attn_mask_radius if self.attn_mask_radius > 0 else -1
        hp.attention_stats(ws, Q, K, M=st.num_modes, d=st.attention_mode_dim, table=table,
                           w_pos=st.pos_code_weight, clip=clip, lse2=lse2, slot=slot, attn_clip=st.attn_clip,
                           diag=st.diag(Q.device), mask_radius=mr)
        return AttentionHandle(g, Q, K, lse2, clip, table, st.pos_code_weight, st.num_modes, st.attention_mode_dim,
                               mask_radius=mr)

    @ops.on_device
    def forward(self, x):
        _require_inference(self.setrans.query.weight)
        B, Cc, h, w = x.shape
        grid = TokenGrid(h, w)
        ws = get_workspace(grid, x.device)
        xf = x.float().contiguous()
        dev = x.device
        if self.out_attn_only:
            # a standalone call hands out handles that OWN their Q/K/lse buffers (the shared workspace is
            # overwritten by the next call); CRAFT.forward uses attend() on workspace buffers instead
            outs = []
            for b in range(B):
                Q, K, T = (grid.zeros(Cc, device=dev) for _ in range(3))
                lse2 = torch.zeros((self.setrans.num_modes, grid.Mp), dtype=torch.float32, device=dev)
                clip = torch.full((1,), float("inf"), dtype=torch.float32, device=dev)
                outs.append(self.attend(ws, xf[b], T, Q, K, lse2, clip, slot=2))
            return outs[0] if B == 1 else outs
        out = torch.empty_like(xf)
        yf = torch.zeros((grid.Mp, Cc), dtype=torch.float32, device=dev)
        for b in range(B):
            att = self.attend(ws, xf[b], ws.T2, ws.Q2, ws.K2, ws.lse2_f2, ws.clip_f2, slot=1)
            self.setrans.out_trans.run(ws, att, ws.T2, 0, out_f=yf)
            ops.unpack_tokens(yf, 0, Cc, grid, out=out[b])
        return out


class WorkspaceCache:
    """LRU of hotpath.Workspace objects keyed by (device, H, W, level-0 mode).  Each CRAFT instance
    owns one (so that two models never share buffers a captured CUDA graph has baked in); standalone
    module calls share the module-level one.  Bounded: KITTI-style variable input sizes would otherwise
    accumulate ~150 MB per shape."""

    def __init__(self, capacity=6):
        self.capacity = capacity
        self._lru = collections.OrderedDict()

    def get(self, grid, device, materialize_level0=False, slot=0):
        """slot: independent buffer sets for the same grid (training keeps one correlation pyramid per batch element)."""
        materialize_level0 = hp.level0_mode(materialize_level0)
        key = (str(device), grid.H, grid.W, materialize_level0, ops.act_dtype(), slot)
        ws = self._lru.get(key)
        if ws is None:
            with torch.cuda.device(device):
                ws = hp.Workspace(grid, device, materialize_level0)
            self._lru[key] = ws
            while len(self._lru) > self.capacity:
                self._lru.popitem(last=False)      # a graph that still needs the buffers holds its own reference
        else:
            self._lru.move_to_end(key)
        return ws

    def clear(self):
        self._lru.clear()


_SHARED_WORKSPACES = WorkspaceCache()


def get_workspace(grid, device, materialize_level0=None, cache=None):
    return (cache or _SHARED_WORKSPACES).get(grid, device, materialize_level0)
