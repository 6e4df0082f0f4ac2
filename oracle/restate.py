"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (craft_b200/).

Functional fp32 restatement (plain torch ops, device-agnostic) of the reference algorithm on the
CRAFT hot path.  Each function cites the reference lines it follows.  It is pinned against the
*executed* reference by tests/test_oracle_golden.py: live against /root/reference where it is
mounted, and against the frozen outputs in tests/golden/ everywhere.  The reference has no
tests, golden vectors or fixtures of its own, so the executed reference
is the only pin there is.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# positional biases -- SlidingPosBiases2D.forward core/setrans.py:690-708
# --------------------------------------------------------------------------------------------
def sliding_pos_bias(table: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """bias[(y1,x1),(y2,x2)] = table[y2-y1+R, x2-x1+R] inside the (2R+1)^2 window, else 0 -> [h*w, h*w]."""
    R = (table.shape[0] - 1) // 2
    ys = torch.arange(h, device=table.device)
    xs = torch.arange(w, device=table.device)
    dy = ys[None, :] - ys[:, None]          # [h1, h2]
    dx = xs[None, :] - xs[:, None]          # [w1, w2]
    oky = dy.abs() <= R
    okx = dx.abs() <= R
    iy = (dy + R).clamp(0, 2 * R)
    ix = (dx + R).clamp(0, 2 * R)
    b = table[iy[:, None, :, None], ix[None, :, None, :]]       # [h1, w1, h2, w2]
    b = b * (oky[:, None, :, None] & okx[None, :, None, :])
    return b.reshape(h * w, h * w)


# --------------------------------------------------------------------------------------------
# token encoder -- SETransInputFeatEncoder.forward core/setrans.py:763-800 (pos_code_type='bias')
# --------------------------------------------------------------------------------------------
def encode_tokens(feat: torch.Tensor) -> torch.Tensor:
    """[B,C,h,w] -> [B,h*w,C] with LayerNorm over C (no affine, eps 1e-12); eval mode (no dropout)."""
    B, C, h, w = feat.shape
    tok = feat.reshape(B, C, h * w).transpose(1, 2)
    return F.layer_norm(tok, (C,), eps=1e-12)


# --------------------------------------------------------------------------------------------
# multi-mode scores -- CrossAttFeatTrans.forward core/setrans.py:501-542
# --------------------------------------------------------------------------------------------
def mode_scores(xq, xk, wq, bq, wk, bk, M, pos_bias=None, pos_weight=1.0, attn_clip=100.0, mask=None):
    """xq [B,U1,C], xk [B,U2,C] -> scores [B,M,U1,U2] (after the data-dependent clamp, bias, mask)."""
    B, U1, C = xq.shape
    d = C // M
    q = F.linear(xq, wq, bq).reshape(B, U1, M, d).permute(0, 2, 1, 3)
    k = F.linear(xk, wk, bk).reshape(B, xk.shape[1], M, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    gmax = float(s.max())
    if gmax > attn_clip:                      # core/setrans.py:527-529
        s = s.clamp(-attn_clip, attn_clip)
    if pos_bias is not None:
        s = s + pos_weight * pos_bias         # core/setrans.py:538-540
    if mask is not None:
        s = s + mask
    return s, gmax


# LearnedSoftAggregate core/setrans.py:289-300
def soft_aggregate_scalar(s, w, b):
    """num_feat = 1: p = softmax_m(w*s + b); out = sum_m p*s.  s [B,M,U1,U2] -> [B,1,U1,U2]."""
    p = torch.softmax(s * w + b, dim=1)
    return (s * p).sum(dim=1, keepdim=True)


def soft_aggregate_feat(x, w, b):
    """num_feat = F: p = softmax_m(<w, x_m> + b); out = sum_m p_m x_m.  x [B,M,U,F] -> [B,U,F]."""
    sc = F.linear(x, w, b)                    # [B,M,U,1]
    p = torch.softmax(sc, dim=1)
    return (x * p).sum(dim=1)


# --------------------------------------------------------------------------------------------
# correlation volume -- TransCorrBlock.update/.corr core/corr.py:148-207, CorrBlock core/corr.py:16-45,73-81
# --------------------------------------------------------------------------------------------
def trans_corr_volume(fmap1, fmap2, wqk, bqk, w_agg, b_agg, table, M=4, pos_weight=0.5, attn_clip=100.0):
    """-> (volume [B,U,h,w] after the global layer-norm, raw volume, gmax)."""
    B, C, h, w = fmap1.shape
    t1, t2 = encode_tokens(fmap1), encode_tokens(fmap2)
    bias = sliding_pos_bias(table, h, w)[None, None]
    s, gmax = mode_scores(t1, t2, wqk, bqk, wqk, bqk, M, bias, pos_weight, attn_clip)
    raw = soft_aggregate_scalar(s, w_agg, b_agg) if M > 1 else s   # [B,1,U,U]
    flat = raw.reshape(B, 1, -1)
    normed = F.layer_norm(flat, (flat.shape[2],), eps=1e-12)       # core/corr.py:200-204
    return normed.reshape(B, h * w, h, w), raw.reshape(B, h * w, h, w), gmax


def plain_corr_volume(fmap1, fmap2):
    """CorrBlock.corr core/corr.py:73-81: fmap1^T fmap2 / sqrt(C) -> [B,U,h,w]."""
    B, C, h, w = fmap1.shape
    a = fmap1.reshape(B, C, h * w)
    b = fmap2.reshape(B, C, h * w)
    return (torch.matmul(a.transpose(1, 2), b) / math.sqrt(C)).reshape(B, h * w, h, w)


def corr_pyramid(volume, num_levels=4):
    """core/corr.py:42-45 / 186-189: [B,U,h,w] -> list of [B*U,1,h_l,w_l] (floor-mode avg_pool2d)."""
    B, U, h, w = volume.shape
    lv = volume.reshape(B * U, 1, h, w)
    out = [lv]
    for _ in range(num_levels - 1):
        lv = F.avg_pool2d(lv, 2, stride=2)
        out.append(lv)
    return out


def corr_lookup(pyramid, coords, radius=4):
    """CorrBlock.__call__ core/corr.py:47-71 with bilinear_sampler core/utils/utils.py:65-79.
    coords [B,2,h,w] (x,y) -> [B, L*(2r+1)^2, h, w].  Window axis 0 offsets x, axis 1 offsets y."""
    B, _, h, w = coords.shape
    r = radius
    n = 2 * r + 1
    c = coords.permute(0, 2, 3, 1).reshape(B * h * w, 1, 1, 2)
    off = torch.arange(-r, r + 1, device=coords.device, dtype=coords.dtype)
    dxx = off.view(n, 1).expand(n, n)          # added to x: varies along window axis 0
    dyy = off.view(1, n).expand(n, n)          # added to y: varies along window axis 1
    outs = []
    for l, vol in enumerate(pyramid):
        H, W = vol.shape[-2:]
        x = c[..., 0] / 2 ** l + dxx
        y = c[..., 1] / 2 ** l + dyy
        gx = 2 * x / (W - 1) - 1
        gy = 2 * y / (H - 1) - 1
        grid = torch.stack([gx, gy], dim=-1)
        s = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        outs.append(s.reshape(B, h, w, n * n))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# --------------------------------------------------------------------------------------------
# attention + value aggregation -- CrossAttFeatTrans / ExpandedFeatTrans core/setrans.py:364-410,553-566
# --------------------------------------------------------------------------------------------
def radius_mask(h, w, radius, device="cpu"):
    """SelfAttVisPosTrans.forward core/setrans.py:579-584 (--f2radius): -1e9 where the Chebyshev distance between
    query and key exceeds `radius` -> [1,1,U,U]."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    c = torch.stack([ys, xs], -1).reshape(-1, 2)
    far = (c[None] - c[:, None]).abs().max(dim=2)[0] > radius
    return (far.float() * -1e9)[None, None]


def self_attention_probs(feat, wq, wk, M, table, pos_weight, attn_clip=100.0, mask_radius=-1):
    """SelfAttVisPosTrans.forward core/setrans.py:578-600 up to the softmax (:553) -> ([B,M,U,U], tokens, gmax)."""
    B, C, h, w = feat.shape
    tok = encode_tokens(feat)
    bias = sliding_pos_bias(table, h, w)[None, None]
    mask = radius_mask(h, w, mask_radius, feat.device) if mask_radius > 0 else None
    s, gmax = mode_scores(tok, tok, wq, None, wk, None, M, bias, pos_weight, attn_clip, mask)
    return torch.softmax(s, dim=-1), tok, gmax


def expanded_feat_trans(x, probs, w_first, w_score, b_score, skip_coeff, M):
    """ExpandedFeatTrans.forward core/setrans.py:364-410 (has_FFN False, has_input_skip True).
    x [B,U,C]; probs [B,M,U,U]; w_first [M*F, C] -> [B,U,F]."""
    B, U, C = x.shape
    Fd = w_first.shape[0] // M
    v = F.linear(x, w_first).transpose(1, 2).reshape(B, M, Fd, U).transpose(2, 3)   # [B,M,U,F]
    o = torch.matmul(probs, v)                                                      # [B,M,U,F]
    agg = soft_aggregate_feat(o, w_score, b_score)
    y = skip_coeff * x + agg
    return F.layer_norm(y, (Fd,), eps=1e-12)


def gma_attention(fmap, w_qk, heads=1):
    """gma.Attention.forward core/gma.py:74-102 (content-only branch) -> [B,heads,U,U]."""
    B, C, h, w = fmap.shape
    qk = F.conv2d(fmap, w_qk)
    q, k = qk.chunk(2, dim=1)
    dh = q.shape[1] // heads
    q = q.reshape(B, heads, dh, h * w).transpose(2, 3) * dh ** -0.5
    k = k.reshape(B, heads, dh, h * w).transpose(2, 3)
    return torch.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)


def gma_aggregate(attn, fmap, w_v, gamma, heads=1):
    """gma.Aggregate.forward core/gma.py:128-142 (project is None when dim == inner_dim)."""
    B, C, h, w = fmap.shape
    v = F.conv2d(fmap, w_v)
    dh = v.shape[1] // heads
    v = v.reshape(B, heads, dh, h * w).transpose(2, 3)
    o = torch.matmul(attn, v).transpose(2, 3).reshape(B, heads * dh, h, w)
    return fmap + gamma * o


# --------------------------------------------------------------------------------------------
# update block -- core/update.py
# --------------------------------------------------------------------------------------------
def motion_encoder(flow, corr, P):
    """BasicMotionEncoder.forward core/update.py:79-87.  P: dict of conv weights/biases."""
    cor = F.relu(F.conv2d(corr, P["convc1.weight"], P["convc1.bias"]))
    cor = F.relu(F.conv2d(cor, P["convc2.weight"], P["convc2.bias"], padding=1))
    flo = F.relu(F.conv2d(flow, P["convf1.weight"], P["convf1.bias"], padding=3))
    flo = F.relu(F.conv2d(flo, P["convf2.weight"], P["convf2.bias"], padding=1))
    out = F.relu(F.conv2d(torch.cat([cor, flo], 1), P["conv.weight"], P["conv.bias"], padding=1))
    return torch.cat([out, flow], 1)


def sep_conv_gru(h, x, P):
    """SepConvGRU.forward core/update.py:49-64."""
    for tag, pad in (("1", (0, 2)), ("2", (2, 0))):
        hx = torch.cat([h, x], 1)
        z = torch.sigmoid(F.conv2d(hx, P["convz" + tag + ".weight"], P["convz" + tag + ".bias"], padding=pad))
        r = torch.sigmoid(F.conv2d(hx, P["convr" + tag + ".weight"], P["convr" + tag + ".bias"], padding=pad))
        q = torch.tanh(F.conv2d(torch.cat([r * h, x], 1), P["convq" + tag + ".weight"],
                                P["convq" + tag + ".bias"], padding=pad))
        h = (1 - z) * h + z * q
    return h


def flow_and_mask_heads(net, P):
    """FlowHead.forward core/update.py:15-16 and the mask head core/update.py:124-127,161."""
    d = F.conv2d(F.relu(F.conv2d(net, P["flow_head.conv1.weight"], P["flow_head.conv1.bias"], padding=1)),
                 P["flow_head.conv2.weight"], P["flow_head.conv2.bias"], padding=1)
    m = F.conv2d(F.relu(F.conv2d(net, P["mask.0.weight"], P["mask.0.bias"], padding=1)),
                 P["mask.2.weight"], P["mask.2.bias"])
    return d, 0.25 * m


def upsample_flow(flow, mask):
    """CRAFT.upsample_flow core/network.py:151-162: convex combination of the 3x3 coarse neighbours."""
    N, _, H, W = flow.shape
    m = torch.softmax(mask.reshape(N, 1, 9, 8, 8, H, W), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).reshape(N, 2, 9, 1, 1, H, W)
    up = (m * nb).sum(dim=2)                       # [N,2,8,8,H,W]
    return up.permute(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W)


def coords_grid(B, h, w, device="cpu"):
    """core/utils/utils.py:82-85: channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    return torch.stack([xs, ys], 0).float()[None].expand(B, -1, -1, -1)
