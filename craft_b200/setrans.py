"""Squeeze-expansion transformer modules with the reference's class names, constructor arguments,
parameter names and forward() signatures (core/setrans.py), executed by the sm_100a kernels.

What differs from the reference, by design:
  * attention matrices [B,M,U,U] and dense positional-bias tensors [1,1,U,U] are never built.
    Where the reference hands such a tensor from one module to the next, these modules hand over an
    opaque handle (`PosBiasHandle`, `AttentionHandle`) that the consuming module understands.
  * only the configuration space the CRAFT drivers can reach is implemented: pos_code_type 'bias',
    has_FFN False, pool_modes_feat 'softmax', no multi-head ablation.  Anything else raises.
  * inference only in this round: dropout / drop-path are identity (eval semantics) and the ops
    carry no autograd; calling them with grad enabled on parameters that require grad raises.
"""
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

    def __init__(self, grid, Q, K, lse2, clip, table, w_pos, M, d):
        self.grid, self.Q, self.K, self.lse2, self.clip = grid, Q, K, lse2, clip
        self.table, self.w_pos, self.M, self.d = table, w_pos, M, d
        self.shape = (len(Q), M, grid.U, grid.U)


# ------------------------------------------------------------------------------------------------
class LearnedSoftAggregate(nn.Module):
    """core/setrans.py:279-300 -- parameters only; the arithmetic is fused into kernel epilogues."""

    def __init__(self, num_feat, group_dim, keepdim=False):
        super().__init__()
        self.group_dim, self.num_feat, self.keepdim = group_dim, num_feat, keepdim
        self.feat2score = nn.Linear(num_feat, 1)


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
                           out_b=out_b, colb=colb, out_f=out_f, colf=colf)

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
        xb = torch.zeros((grid.Mp, Cc), dtype=torch.bfloat16, device=input_feat.device)
        yf = torch.zeros((grid.Mp, self.feat_dim), dtype=torch.float32, device=input_feat.device)
        for b in range(B):
            xb.view(grid.H, grid.Wp, Cc)[:, :grid.W] = input_feat[b].reshape(grid.H, grid.W, Cc).to(torch.bfloat16)
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
        self.max_attn, self.clamp_count, self.call_count = 0, 0, 0
        self._init_weights()
        self._packed = hp.PackedWeights()

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

    def forward(self, query_feat, key_feat=None, pos_biases=None, attention_mask=None):
        raise NotImplementedError(
            "craft_b200.CrossAttFeatTrans is driven through its owners (TransCorrBlock.update, "
            "SelfAttVisPosTrans.forward): a standalone call would have to return a [B,M,U,U] tensor, which "
            "this implementation never materialises")


class SelfAttVisPosTrans(nn.Module):
    """core/setrans.py:568-619."""

    def __init__(self, config, name):
        super().__init__()
        self.config = copy.copy(config)
        self.name = name
        self.out_attn_only = config.out_attn_scores_only or config.out_attn_probs_only
        self.attn_mask_radius = config.attn_mask_radius
        if self.attn_mask_radius > 0:
            raise NotImplementedError("craft_b200: --f2radius masking (default off) is not implemented")
        self.setrans = CrossAttFeatTrans(self.config, name)
        self.vispos_encoder = SETransInputFeatEncoder(self.config)

    def attend(self, ws, feat_chw, T, Q, K, lse2, clip, slot, pack_mode=ops.PACK_LN):
        """tokens -> projections -> softmax statistics; returns the AttentionHandle."""
        g = ws.grid
        st = self.setrans
        ops.pack_tokens(feat_chw, g, pack_mode, out_b=T)
        st.project(ws, T, T, Q, K)
        table = self.vispos_encoder.table()
        hp.attention_stats(ws, Q, K, M=st.num_modes, d=st.attention_mode_dim, table=table,
                           w_pos=st.pos_code_weight, clip=clip, lse2=lse2, slot=slot, attn_clip=st.attn_clip)
        return AttentionHandle(g, Q, K, lse2, clip, table, st.pos_code_weight, st.num_modes, st.attention_mode_dim)

    def forward(self, x):
        _require_inference(self.setrans.query.weight)
        B, Cc, h, w = x.shape
        grid = TokenGrid(h, w)
        ws = get_workspace(grid, x.device)
        xf = x.float().contiguous()
        if self.out_attn_only:
            if B != 1:
                # each handle owns its Q/K/lse buffers
                outs = []
                for b in range(B):
                    Q, K, T = (torch.zeros((grid.Mp, Cc), dtype=torch.bfloat16, device=x.device) for _ in range(3))
                    lse2 = torch.zeros((self.setrans.num_modes, grid.Mp), dtype=torch.float32, device=x.device)
                    clip = torch.full((1,), float("inf"), dtype=torch.float32, device=x.device)
                    outs.append(self.attend(ws, xf[b], T, Q, K, lse2, clip, slot=2))
                return outs
            return self.attend(ws, xf[0], ws.Ta, ws.Qa, ws.Ka, ws.lse2_att, ws.clip_att, slot=2)
        out = torch.empty_like(xf)
        yf = torch.zeros((grid.Mp, Cc), dtype=torch.float32, device=x.device)
        for b in range(B):
            att = self.attend(ws, xf[b], ws.T2, ws.Q2, ws.K2, ws.lse2_f2, ws.clip_f2, slot=1)
            self.setrans.out_trans.run(ws, att, ws.T2, 0, out_f=yf)
            ops.unpack_tokens(yf, 0, Cc, grid, out=out[b])
        return out


_WORKSPACES = {}


def get_workspace(grid, device, materialize_level0=None):
    import os
    if materialize_level0 is None:
        materialize_level0 = False
    key = (str(device), grid.H, grid.W, bool(materialize_level0))
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = hp.Workspace(grid, device, materialize_level0)
        _WORKSPACES[key] = ws
    return ws
