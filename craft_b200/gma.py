"""GMA single-head attention and aggregation with the reference's names (core/gma.py), for the
non-setrans CRAFT variant (`use_setrans=False`, BASELINE config 3).  Same kernels as setrans:
M = 1 mode of d = 128, no positional bias, no clamp; Aggregate's epilogue is fmap + gamma * (P V).

RelPosEmb (`--position_only`, `--position_and_content`; default off, train.py:368-371) is an
ablation outside the hot path and raises.
"""
import torch
import torch.nn as nn

from . import hotpath as hp
from . import ops
from .ops import TokenGrid
from .setrans import AttentionHandle, get_workspace, _require_inference


class RelPosEmb(nn.Module):
    """core/gma.py:6-50 -- parameters only (so GMA-variant checkpoints load key for key); its forward
    belongs to the --position_only / --position_and_content ablations, which are out of scope."""

    def __init__(self, max_pos_size, dim_head):
        super().__init__()
        self.rel_height = nn.Embedding(2 * max_pos_size - 1, dim_head)
        self.rel_width = nn.Embedding(2 * max_pos_size - 1, dim_head)
        deltas = torch.arange(max_pos_size).view(1, -1) - torch.arange(max_pos_size).view(-1, 1)
        self.register_buffer("rel_ind", deltas + max_pos_size - 1)


class Attention(nn.Module):
    """core/gma.py:53-102."""

    def __init__(self, *, args, dim, max_pos_size=100, heads=4, dim_head=128):
        super().__init__()
        self.args, self.heads = args, heads
        self.scale = dim_head ** -0.5
        if heads != 1 or dim_head != 128 or dim != 128:
            raise NotImplementedError("craft_b200 gma.Attention: CRAFT uses heads=1, dim=dim_head=128")
        if getattr(args, "position_only", False) or getattr(args, "position_and_content", False):
            raise NotImplementedError("RelPosEmb ablations are outside the hot path")
        self.to_qk = nn.Conv2d(dim, heads * dim_head * 2, 1, bias=False)
        self.pos_emb = RelPosEmb(max_pos_size, dim_head)
        self.pos_embed_weight = 1.0
        self._packed = hp.PackedWeights()

    def packed(self):
        def build():
            w = self.to_qk.weight.detach().float().reshape(256, 128)
            return dict(wq=ops.pack_linear_weight(w[:128]), wk=ops.pack_linear_weight(w[128:]))
        return self._packed.get("qk", [self.to_qk.weight], build)

    def attend(self, ws, T, Q, K, lse2, clip):
        """T: bf16 token rows of fmap (no LayerNorm in GMA). softmax(q*scale . k) == scores scaled by 1/sqrt(128)."""
        g = ws.grid
        pk = self.packed()
        hp.project(g, T, pk["wq"], None, Q, K=128)
        hp.project(g, T, pk["wk"], None, K, K=128)
        clip.fill_(float("inf"))
        smax = ws.stat_max[2:3]
        smax.fill_(-float("inf"))
        ops.attn_lse(Q, K, g, M=1, d=128, w_pos=0.0, pos_table=None, R=7, clip=clip, stat_max=smax,
                     lse_part=ws.lse_part, lse2=lse2, ksplit=ws.ks_sc)
        return AttentionHandle(g, Q, K, lse2, clip, None, 0.0, 1, 128)

    @ops.on_device
    def forward(self, fmap):
        """[B,128,h,w] -> AttentionHandle (list for B > 1) standing for softmax(q.k) (core/gma.py:74-102).
        Handles of a standalone call own their buffers."""
        _require_inference(self.to_qk.weight)
        B, Cc, h, w = fmap.shape
        grid = TokenGrid(h, w)
        dev = fmap.device
        ws = get_workspace(grid, dev)
        outs = []
        for b in range(B):
            T, Q, K = (grid.zeros(Cc, device=dev) for _ in range(3))
            lse2 = torch.zeros((1, grid.Mp), dtype=torch.float32, device=dev)
            clip = torch.full((1,), float("inf"), dtype=torch.float32, device=dev)
            ops.pack_tokens(fmap[b].float().contiguous(), grid, ops.PACK_COPY, out_b=T)
            outs.append(self.attend(ws, T, Q, K, lse2, clip))
        return outs[0] if B == 1 else outs


class Aggregate(nn.Module):
    """core/gma.py:105-142."""

    def __init__(self, args, dim, heads=4, dim_head=128):
        super().__init__()
        self.args, self.heads = args, heads
        self.scale = dim_head ** -0.5
        inner = heads * dim_head
        if dim != inner or heads != 1:
            raise NotImplementedError("craft_b200 gma.Aggregate: heads=1, dim == inner_dim (project is None)")
        self.to_v = nn.Conv2d(dim, inner, 1, bias=False)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.project = None
        self._packed = hp.PackedWeights()

    def packed(self):
        def build():
            return dict(wv=ops.pack_linear_weight(self.to_v.weight.detach().float().reshape(128, 128)),
                        gamma=self.gamma.detach().float().contiguous())
        return self._packed.get("v", [self.to_v.weight, self.gamma], build)

    def run(self, ws, att, X, x_koff, out_b=None, colb=0, out_f=None, colf=0):
        pk = self.packed()
        hp.value_aggregate(ws, att.Q, att.K, X, x_koff, pk["wv"], M=1, d=128, F=128, table=None, w_pos=0.0,
                           clip=att.clip, lse2=att.lse2, w_score=pk["gamma"], b_score=pk["gamma"],
                           coeff=pk["gamma"], gma=1, out_b=out_b, colb=colb, out_f=out_f, colf=colf)

    @ops.on_device
    def forward(self, attn, fmap):
        """(AttentionHandle(s), fmap [B,128,h,w]) -> fmap + gamma * (attn @ to_v(fmap)) (core/gma.py:128-142)."""
        atts = attn if isinstance(attn, (list, tuple)) else [attn]
        if not all(isinstance(a, AttentionHandle) for a in atts):
            raise TypeError("craft_b200.gma.Aggregate takes the AttentionHandle returned by gma.Attention")
        _require_inference(self.to_v.weight)
        B, Cc, h, w = fmap.shape
        if len(atts) != B:
            raise ValueError("need one attention handle per batch element")
        grid = atts[0].grid
        dev = fmap.device
        ws = get_workspace(grid, dev)
        xb = grid.zeros(Cc, device=dev)
        yf = grid.zeros(Cc, dtype=torch.float32, device=dev)
        out = torch.empty((B, Cc, h, w), dtype=torch.float32, device=dev)
        for b in range(B):
            ops.pack_tokens(fmap[b].float().contiguous(), grid, ops.PACK_COPY, out_b=xb)
            self.run(ws, atts[b], xb, 0, out_f=yf)
            ops.unpack_tokens(yf, 0, Cc, grid, out=out[b])
        return out
