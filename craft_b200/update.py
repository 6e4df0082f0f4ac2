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


def _rows_in(x_chw, grid, cols=None, col0=0):
    """[C,h,w] fp32 -> operand-dtype padded-flat rows [Mp, cols] with the channels at col0 (standalone boundary only)."""
    Cc = x_chw.shape[0]
    cols = cols or ((Cc + 63) // 64) * 64
    buf = torch.zeros((grid.H, grid.Wp, cols), dtype=ops.act_dtype(), device=x_chw.device)
    buf[:, :grid.W, col0:col0 + Cc] = x_chw.permute(1, 2, 0).to(buf.dtype)
    return buf.view(grid.Mp, cols)


class FlowHead(nn.Module):
    """core/update.py:8-16.  Inside GMAUpdateBlock the first convolution is computed together with the
    mask head's (hotpath.heads); forward() is the standalone form on the same shift-GEMM kernel."""

    def __init__(self, input_dim=128, hidden_dim=256):
        super().__init__()
        if input_dim != 128 or hidden_dim != 256:
            raise NotImplementedError("craft_b200 FlowHead: 128 -> 256 -> 2 (GMAUpdateBlock)")
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, 2, 3, padding=1)
        self._packed = hp.PackedWeights()

    @ops.on_device
    def forward(self, x):
        """[B,128,h,w] -> [B,2,h,w]: conv2(relu(conv1(x))) (core/update.py:15-16)."""
        _require_inference(self.conv1.weight)
        B, _, h, w = x.shape
        g = TokenGrid(h, w)
        pk = self._packed.get("w", list(self.parameters()), lambda: dict(
            w1=ops.pack_conv_weight(self.conv1.weight), b1=self.conv1.bias.detach().float().contiguous(),
            w2=ops.pack_conv_weight(self.conv2.weight, Npad=32), b2=ops.pad_bias(self.conv2.bias, 32)))
        taps = ops.conv_taps(3, 3, g)
        hid = g.zeros(256, device=x.device)
        d = g.zeros(32, dtype=torch.float32, device=x.device)
        out = torch.empty((B, 2, h, w), dtype=torch.float32, device=x.device)
        for b in range(B):
            X = _rows_in(x[b].float(), g)
            ops.shift_gemm(X, pk["w1"], M=g.Mp, Npad=256, K=128, BN=128, taps=taps, grid=g, bias=pk["b1"], act=1, out_b=hid)
            ops.shift_gemm(hid, pk["w2"], M=g.Mp, Npad=32, K=256, BN=32, taps=taps, grid=g, bias=pk["b2"], out_f=d)
            ops.unpack_tokens(d, 0, 2, g, out=out[b])
        return out


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
        self._packed = hp.PackedWeights()

    def packed_passes(self, grid):
        """[(w_zr, b_zr, w_q, b_q, taps)] for the horizontal and the vertical pass (shared with hotpath.UpdateWeights)."""
        return self._packed.get(("gru", grid.H, grid.W), list(self.parameters()), lambda: hp.pack_gru(self, grid))

    @ops.on_device
    def forward(self, h, x):
        """(h [B,128,h,w], x [B,384,h,w]) -> new h (core/update.py:49-64), fp32 state, bf16 conv operands."""
        _require_inference(self.convz1.weight)
        B, _, hh, ww = h.shape
        g = TokenGrid(hh, ww)
        dev = h.device
        X = g.zeros(640, device=dev)
        Hm, Z = g.zeros(128, dtype=torch.float32, device=dev), g.zeros(128, dtype=torch.float32, device=dev)
        out = torch.empty((B, 128, hh, ww), dtype=torch.float32, device=dev)
        passes = self.packed_passes(g)
        for b in range(B):
            ops.pack_tokens(h[b].float().contiguous(), g, ops.PACK_COPY, out_b=X, colb=0, out_f=Hm)
            X[:, 128:512] = _rows_in(x[b].float(), g)
            hp.sep_conv_gru_rows(g, X, Hm, Z, passes)
            ops.unpack_tokens(Hm, 0, 128, g, out=out[b])
        return out


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
        self._packed = hp.PackedWeights()

    @ops.on_device
    def forward(self, flow, corr):
        """(flow [B,2,h,w], corr [B,324,h,w]) -> [B,128,h,w] = cat(conv(...), flow) (core/update.py:79-87)."""
        _require_inference(self.convc1.weight)
        B, _, h, w = flow.shape
        g = TokenGrid(h, w)
        dev = flow.device
        ew = self._packed.get(("enc", h, w), list(self.parameters()), lambda: hp.EncoderWeights(self, g))
        bufs = hp.EncoderBuffers(g, dev)
        out = torch.empty((B, 128, h, w), dtype=torch.float32, device=dev)
        for b in range(B):
            bufs.CORR[:, :] = _rows_in(corr[b].float(), g, cols=384)
            bufs.flow.view(g.H, g.Wp, 2)[:, :g.W] = flow[b].float().permute(1, 2, 0)
            hp.motion_encoder(bufs, ew)
            ops.unpack_tokens(bufs.X, 256, 128, g, out=out[b])
        return out


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

    def step(self, ws, attention, it=0, need_mask=True, lookup=None, update_flow=False):
        """One refinement iteration on the workspace: CORR (looked-up correlation) and flow are in
        place (or produced by `lookup(part)`, see hotpath.motion_encoder); writes the new hidden state
        (X[:, :128], Hm), DELTA and (when need_mask) MASK."""
        uw = self.weights(ws.grid)
        hp.motion_encoder(ws, uw, lookup)
        self.aggregator.run(ws, attention, ws.X, 256, out_b=ws.X, colb=384)
        hp.sep_conv_gru(ws, uw)
        hp.heads(ws, uw, it, need_mask, update_flow)

    @ops.on_device
    def forward(self, net, inp, corr, flow, attention):
        """(net, inp, corr, flow, attention) -> (net, mask, delta_flow); NCHW fp32 at the boundary
        (core/update.py:137-162).  `attention` is the AttentionHandle (list for B > 1) of self.att."""
        _require_inference(self.encoder.convc1.weight)
        B, _, h, w = net.shape
        grid = TokenGrid(h, w)
        ws = get_workspace(grid, net.device)
        g = grid
        atts = attention if isinstance(attention, (list, tuple)) else [attention] * B
        if len(atts) != B:
            raise ValueError("need one attention handle per batch element")
        net_o = torch.empty((B, 128, h, w), dtype=torch.float32, device=net.device)
        mask_o = torch.empty((B, 576, h, w), dtype=torch.float32, device=net.device)
        delta_o = torch.empty((B, 2, h, w), dtype=torch.float32, device=net.device)
        for b in range(B):
            ops.pack_tokens(net[b].float().contiguous(), g, ops.PACK_COPY, out_b=ws.X, colb=0, out_f=ws.Hm)
            ops.pack_tokens(inp[b].float().contiguous(), g, ops.PACK_COPY, out_b=ws.X, colb=128)
            ws.CORR.view(g.H, g.Wp, 384)[:, :g.W, :324] = corr[b].permute(1, 2, 0).to(ws.CORR.dtype)
            ws.flow.view(g.H, g.Wp, 2)[:, :g.W] = flow[b].float().permute(1, 2, 0)
            self.step(ws, atts[b])
            ops.unpack_tokens(ws.Hm, 0, 128, g, out=net_o[b])
            ops.unpack_tokens(ws.MASK, 0, 576, g, out=mask_o[b])
            ops.unpack_tokens(ws.DELTA, 0, 2, g, out=delta_o[b])
        return net_o, mask_o, delta_o
