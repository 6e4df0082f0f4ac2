"""Update block with the reference's class and parameter names (core/update.py): motion encoder,
SepConvGRU, flow head, mask head, aggregator dispatch -- all executed as tcgen05 shift-GEMMs on the
shared X buffer (hotpath.py).  Standalone forward()s convert NCHW tensors at the boundary; inside
CRAFT.forward nothing is converted between iterations.
"""
import torch
import torch.nn as nn

from . import hotpath as hp
from . import ops
from .gma import Aggregate
from .ops import TokenGrid
from .setrans import ExpandedFeatTrans, get_workspace, _require_inference


class FlowHead(nn.Module):
    """core/update.py:8-16 (parameters; computed together with the mask head in hotpath.heads)."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)


class SepConvGRU(nn.Module):
    """core/update.py:37-64."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128):
        super().__init__()
        if hidden_dim != 128 or input_dim != 384:
            raise NotImplementedError("craft_b200 SepConvGRU: hidden 128, input 384 (GMAUpdateBlock)")
        c = hidden_dim + input_dim
        self.convz1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convr1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convq1 = nn.Conv2d(c, hidden_dim, (1, 5), padding=(0, 2))
        self.convz2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convr2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))
        self.convq2 = nn.Conv2d(c, hidden_dim, (5, 1), padding=(2, 0))


class BasicMotionEncoder(nn.Module):
    """core/update.py:67-87."""

    def __init__(self, args):
        super().__init__()
        cor_planes = args.corr_levels * args.corr_multiplier * (2 * args.corr_radius + 1) ** 2
        if cor_planes != 324:
            raise NotImplementedError("craft_b200 motion encoder: 4 levels x 81 taps (radius 4, one-way correlation)")
        self.convc1 = nn.Conv2d(cor_planes, 256, 1, padding=0)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)


class GMAUpdateBlock(nn.Module):
    """core/update.py:116-162."""

    def __init__(self, args, hidden_dim=128):
        super().__init__()
        self.args = args
        self.encoder = BasicMotionEncoder(args)
        self.gru = SepConvGRU(hidden_dim=hidden_dim, input_dim=128 + hidden_dim + hidden_dim)
        self.flow_head = FlowHead(hidden_dim, hidden_dim=256)
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, 64 * 9, 1, padding=0))
        self.use_setrans = args.use_setrans
        if self.use_setrans:
            self.intra_trans_config = args.intra_trans_config
            self.aggregator = ExpandedFeatTrans(self.intra_trans_config, "Motion Aggregator")
        else:
            self.aggregator = Aggregate(args=self.args, dim=128, dim_head=128, heads=self.args.num_heads)
        self._packed = hp.PackedWeights()

    def weights(self, grid):
        ps = [p for n, p in self.named_parameters() if not n.startswith("aggregator.")]
        return self._packed.get(("uw", grid.H, grid.W), ps, lambda: hp.UpdateWeights(self, grid))

    def step(self, ws, attention, it=0, need_mask=True, lookup=None):
        """One refinement iteration on the workspace: CORR (looked-up correlation) and flow are in
        place (or produced by `lookup(part)`, see hotpath.motion_encoder); writes the new hidden state
        (X[:, :128], Hm), DELTA and (when need_mask) MASK."""
        uw = self.weights(ws.grid)
        hp.motion_encoder(ws, uw, lookup)
        self.aggregator.run(ws, attention, ws.X, 256, out_b=ws.X, colb=384)
        hp.sep_conv_gru(ws, uw)
        hp.heads(ws, uw, it, need_mask)

    def forward(self, net, inp, corr, flow, attention):
        """(net, inp, corr, flow, attention) -> (net, mask, delta_flow); NCHW fp32 at the boundary."""
        _require_inference(self.encoder.convc1.weight)
        B, _, h, w = net.shape
        if B != 1:
            raise NotImplementedError("standalone GMAUpdateBlock.forward handles one pair per call")
        grid = TokenGrid(h, w)
        ws = get_workspace(grid, net.device)
        g = grid
        ops.pack_tokens(net[0].float().contiguous(), g, ops.PACK_COPY, out_b=ws.X, colb=0, out_f=ws.Hm)
        ops.pack_tokens(inp[0].float().contiguous(), g, ops.PACK_COPY, out_b=ws.X, colb=128)
        ws.CORR.view(g.H, g.Wp, 384)[:, :g.W, :324] = corr[0].permute(1, 2, 0).to(torch.bfloat16)
        ws.flow.view(g.H, g.Wp, 2)[:, :g.W] = flow[0].permute(1, 2, 0)
        att = attention[0] if isinstance(attention, list) else attention
        self.step(ws, att)
        net_o = ops.unpack_tokens(ws.Hm, 0, 128, g)[None]
        mask_o = ops.unpack_tokens(ws.MASK, 0, 576, g)[None]
        delta_o = ops.unpack_tokens(ws.DELTA, 0, 2, g)[None]
        return net_o, mask_o, delta_o
