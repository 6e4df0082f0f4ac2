"""Training path of CRAFT.forward (grad mode): what `train.py` / `train_ddp.py` call.

The fused kernels of craft_b200 are forward-only.  Under grad mode the model therefore runs

  * the three attention blocks -- F2 transformer, correlation volume + lookups, intra-frame attention +
    motion aggregator (together 70 % of the hot-path FLOPs) -- as `torch.autograd.Function`s whose FORWARD is
    the sm_100a kernel path (the standalone module forwards) and whose BACKWARD recomputes the block with the
    differentiable PyTorch restatement below on the saved inputs (hard part 5: "backward
    recomputes with reference PyTorch ops first; native backward later");
  * everything whose backward cuDNN / ATen already provide -- encoders, the update block's convolutions,
    convex upsampling -- as ordinary autograd-native PyTorch on the modules' own parameters.

Consumers that reuse one producer many times (12 lookups into one correlation volume, 12 aggregations with
one attention matrix) are wired through a zero-size "token" tensor: each consumer's backward only RECORDS its
incoming gradient, and the producer's backward -- which autograd runs after all of them -- recomputes the
dense [M,U,U] tensor ONCE and back-propagates all recorded gradients through it.

Dropout (attention probabilities p=0.2, tokens p=0.1 in the reference's training configuration,
core/setrans.py:110-111,557,795) has no kernel form: when it is active the attention blocks run the PyTorch
restatement in forward as well, so the reference's training semantics are kept exactly.  `args.dropout_prob = 0`
(or eval mode under grad, e.g. fine-tuning with frozen regularisation) enables the kernel forward.

Everything here is CUDA-only like the rest of the package: CRAFT.forward refuses CPU tensors before it gets
here.  The only collective in multi-GPU training stays DDP's gradient all-reduce (train_ddp.py:198-200).
"""
import math

import torch
import torch.nn.functional as F

from . import ops
from .ops import TokenGrid


# ------------------------------------------------------------------------------------------------
# differentiable PyTorch restatement of the attention blocks (used for backward recomputation, and for
# forward when dropout is active).  Reference lines as in craft_b200/setrans.py / corr.py.
# ------------------------------------------------------------------------------------------------
def _tokens_ln(feat):
    """SETransInputFeatEncoder.forward core/setrans.py:791-794: [B,C,h,w] -> LayerNorm'ed [B,U,C]."""
    B, C, h, w = feat.shape
    return F.layer_norm(feat.reshape(B, C, h * w).transpose(1, 2), (C,), eps=1e-12)


def _pos_bias(table, h, w):
    """SlidingPosBiases2D.forward core/setrans.py:690-708 -> dense [U,U], differentiable w.r.t. the table.
    Written as two one-hot matmuls, bias[(y1,x1),(y2,x2)] = sum_ab [y2-y1+R == a] T[a,b] [x2-x1+R == b]: the
    backward is two small GEMMs.  (A gather `table[iy, ix]` has a scatter-add backward of U^2 = 20 M atomics
    onto 225 addresses -- measured: it dominated the training step.)"""
    R = (table.shape[0] - 1) // 2
    n = 2 * R + 1

    def onehot(m):
        i = torch.arange(m, device=table.device)
        d = i[None, :] - i[:, None] + R                       # [m1, m2]
        return (d[..., None] == torch.arange(n, device=table.device)).to(table.dtype).reshape(m * m, n)
    ay, ax = onehot(h), onehot(w)                              # [h*h, n], [w*w, n]
    b = (ay @ table @ ax.t()).reshape(h, h, w, w)              # (y1, y2, x1, x2)
    return b.permute(0, 2, 1, 3).reshape(h * w, h * w)


def _radius_mask(h, w, radius, device):
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    c = torch.stack([ys, xs], -1).reshape(-1, 2)
    return ((c[None] - c[:, None]).abs().max(dim=2)[0] > radius).float() * -1e9


def _scores(xq, xk, st, bias, mask=None):
    """CrossAttFeatTrans.forward core/setrans.py:507-542 -> [B,M,U,U]."""
    B, U, C = xq.shape
    M = st.num_modes
    d = C // M
    q = st.query(xq).reshape(B, U, M, d).permute(0, 2, 1, 3)
    k = st.key(xk).reshape(B, xk.shape[1], M, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    if s.detach().max().item() > st.attn_clip:                        # core/setrans.py:527-529
        s = s.clamp(-st.attn_clip, st.attn_clip)
    s = s + st.pos_code_weight * bias
    if mask is not None:
        s = s + mask
    return s


def _soft_aggregate(agg, x, dim=1, keepdim=False):
    """LearnedSoftAggregate.forward core/setrans.py:289-300."""
    sc = agg.feat2score(x.unsqueeze(-1)).squeeze(-1) if agg.num_feat == 1 else agg.feat2score(x)
    return (x * sc.softmax(dim=dim)).sum(dim=dim, keepdim=keepdim)


def _expanded_feat_trans(ot, x, probs):
    """ExpandedFeatTrans.forward core/setrans.py:364-410 (no FFN, input skip, softmax mode pooling)."""
    B, U, C = x.shape
    M, Fd = ot.num_modes, ot.feat_dim
    v = ot.first_linear(x).transpose(1, 2).reshape(B, M, Fd, U).transpose(2, 3)
    o = torch.matmul(probs, v)
    y = ot.input_skip_coeff * x + _soft_aggregate(ot.feat_softaggr, o)
    return F.layer_norm(y, (Fd,), eps=1e-12)


def _self_att_probs(sa, feat, training):
    """SelfAttVisPosTrans.forward core/setrans.py:578-600 up to the attention probabilities (:553-557)."""
    B, C, h, w = feat.shape
    tok = _tokens_ln(feat)
    p_tok = sa.config.hidden_dropout_prob
    tok = F.dropout(tok, p_tok, training)
    bias = _pos_bias(sa.vispos_encoder.pos_coder.biases, h, w)
    mask = _radius_mask(h, w, sa.attn_mask_radius, feat.device) if sa.attn_mask_radius > 0 else None
    probs = _scores(tok, tok, sa.setrans, bias, mask).softmax(dim=-1)
    return F.dropout(probs, sa.config.attention_probs_dropout_prob, training), tok


def _f2_trans(sa, feat, training):
    probs, tok = _self_att_probs(sa, feat, training)
    y = _expanded_feat_trans(sa.setrans.out_trans, tok, probs)
    return y.permute(0, 2, 1).reshape(feat.shape)


def _corr_volume(cf, fmap1, fmap2, training):
    """TransCorrBlock.update/.corr core/corr.py:148-207 -> normalised volume [B,U,h,w]."""
    B, C, h, w = fmap1.shape
    p = cf.config.hidden_dropout_prob
    t1, t2 = F.dropout(_tokens_ln(fmap1), p, training), F.dropout(_tokens_ln(fmap2), p, training)
    bias = _pos_bias(cf.vispos_encoder.pos_coder.biases, h, w)
    s = _scores(t1, t2, cf.setrans, bias)
    raw = _soft_aggregate(cf.setrans.attn_softaggr, s, keepdim=True) if cf.setrans.num_modes > 1 else s
    if cf.do_corr_global_norm:
        # F.layer_norm over ONE row of U^2 = 20 M elements (core/corr.py:200-204) runs in a single thread block
        # (34 ms forward + 44 ms backward at 400x720); the same arithmetic as parallel reductions:
        flat = raw.float().reshape(B, -1)
        var, mean = torch.var_mean(flat, dim=1, unbiased=False, keepdim=True)
        raw = (flat - mean) * torch.rsqrt(var + 1e-12)
    return raw.reshape(B, h * w, h, w)


def _plain_volume(fmap1, fmap2):
    B, C, h, w = fmap1.shape
    a, b = fmap1.reshape(B, C, h * w), fmap2.reshape(B, C, h * w)
    return (torch.matmul(a.transpose(1, 2), b) / math.sqrt(C)).reshape(B, h * w, h, w)


def _pyramid(vol, levels=4):
    B, U, h, w = vol.shape
    lv = vol.reshape(B * U, 1, h, w)
    out = [lv]
    for _ in range(levels - 1):
        lv = F.avg_pool2d(lv, 2, stride=2)
        out.append(lv)
    return out


def _lookup(pyr, coords, r=4):
    """CorrBlock.__call__ core/corr.py:47-71 + bilinear_sampler core/utils/utils.py:65-79."""
    B, _, h, w = coords.shape
    n = 2 * r + 1
    c = coords.permute(0, 2, 3, 1).reshape(B * h * w, 1, 1, 2)
    off = torch.arange(-r, r + 1, device=coords.device, dtype=coords.dtype)
    dxx, dyy = off.view(n, 1).expand(n, n), off.view(1, n).expand(n, n)
    outs = []
    for l, vol in enumerate(pyr):
        H, W = vol.shape[-2:]
        gx = 2 * (c[..., 0] / 2 ** l + dxx) / (W - 1) - 1
        gy = 2 * (c[..., 1] / 2 ** l + dyy) / (H - 1) - 1
        s = F.grid_sample(vol, torch.stack([gx, gy], dim=-1), mode="bilinear", padding_mode="zeros", align_corners=True)
        outs.append(s.reshape(B, h, w, n * n))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def _gma_attention(att, fmap):
    B, C, h, w = fmap.shape
    q, k = att.to_qk(fmap).chunk(2, dim=1)
    q = q.reshape(B, 1, -1, h * w).transpose(2, 3) * att.scale
    k = k.reshape(B, 1, -1, h * w).transpose(2, 3)
    return torch.matmul(q, k.transpose(-1, -2)).softmax(dim=-1)


def _gma_aggregate(ag, attn, fmap):
    B, C, h, w = fmap.shape
    v = ag.to_v(fmap).reshape(B, 1, -1, h * w).transpose(2, 3)
    o = torch.matmul(attn, v).transpose(2, 3).reshape(B, -1, h, w)
    return fmap + ag.gamma * o


def upsample_flow(flow, mask):
    """CRAFT.upsample_flow core/network.py:151-162 (autograd-native)."""
    N, _, H, W = flow.shape
    m = torch.softmax(mask.view(N, 1, 9, 8, 8, H, W), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(N, 2, 9, 1, 1, H, W)
    return torch.sum(m * nb, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(N, 2, 8 * H, 8 * W)


# ------------------------------------------------------------------------------------------------
# kernel-forward / recompute-backward functions
# ------------------------------------------------------------------------------------------------
def _params(mod):
    return [p for p in mod.parameters() if p.requires_grad]


def _unique(ts):
    seen, out = set(), []
    for t in ts:
        if id(t) not in seen:
            seen.add(id(t))
            out.append(t)
    return out


class _KernelFwd(torch.autograd.Function):
    """y = kernel_fn(x) in forward (no graph); backward = autograd through torch_fn(x) recomputed with grad.
    Extra tensor arguments after x are the parameters torch_fn touches (so that they receive gradients)."""

    @staticmethod
    def forward(ctx, kernel_fn, torch_fn, x, *params):
        ctx.torch_fn, ctx.params = torch_fn, params
        ctx.save_for_backward(x)
        with torch.no_grad():
            return kernel_fn(x.detach())

    @staticmethod
    def backward(ctx, gy):
        (x,), params = ctx.saved_tensors, ctx.params
        with torch.enable_grad():
            xr = x.detach().requires_grad_(ctx.needs_input_grad[2])
            y = ctx.torch_fn(xr)
            wanted = ([xr] if ctx.needs_input_grad[2] else []) + [p for p, need in zip(params, ctx.needs_input_grad[3:]) if need]
            grads = torch.autograd.grad(y, wanted, gy, allow_unused=True)
        grads = list(grads)
        gx = grads.pop(0) if ctx.needs_input_grad[2] else None
        gp = [grads.pop(0) if need else None for need in ctx.needs_input_grad[3:]]
        return (None, None, gx, *gp)


class _Producer(torch.autograd.Function):
    """Forward: run `kernel_build(*inputs)` (fills kernel-side state), return a 1-element token.  Backward: runs
    after every consumer has recorded (replay_fn, grad_out) in `shared["uses"]`: recompute the dense tensor once
    with `torch_build(*inputs)` and back-propagate sum_k <consumer_k(dense), grad_k>."""

    @staticmethod
    def forward(ctx, shared, kernel_build, torch_build, n_in, *tensors):
        ctx.shared, ctx.torch_build, ctx.n_in = shared, torch_build, n_in
        ctx.params = tensors[n_in:]
        ctx.save_for_backward(*tensors[:n_in])
        with torch.no_grad():
            kernel_build(*[t.detach() for t in tensors[:n_in]])
        return torch.zeros(1, device=tensors[0].device, dtype=torch.float32)

    @staticmethod
    def backward(ctx, _g):
        tensors = tuple(ctx.saved_tensors) + tuple(ctx.params)
        needs = ctx.needs_input_grad[4:]
        with torch.enable_grad():
            ins = [t.detach().requires_grad_(needs[i]) for i, t in enumerate(tensors[:ctx.n_in])]
            dense = ctx.torch_build(*ins)
            total = None
            for replay, g in ctx.shared["uses"]:
                term = (replay(dense) * g).sum()
                total = term if total is None else total + term
            extra = ctx.shared.get("dense_grad")             # gradient w.r.t. the dense tensor itself (aggregator)
            if extra is not None:
                term = (dense * extra).sum()
                total = term if total is None else total + term
            wanted = [t for t, need in zip(list(ins) + list(tensors[ctx.n_in:]), needs) if need]
            grads = list(torch.autograd.grad(total, wanted, allow_unused=True)) if total is not None else [None] * len(wanted)
        ctx.shared["uses"].clear()
        ctx.shared.pop("dense_grad", None)
        ctx.shared.pop("P", None)
        return (None, None, None, None, *[grads.pop(0) if need else None for need in needs])


class _Lookup(torch.autograd.Function):
    """Correlation lookup at detached coordinates: kernel forward, gradient recorded for the producer."""

    @staticmethod
    def forward(ctx, shared, kernel_lookup, token, coords):
        ctx.shared = shared
        ctx.coords = coords.detach()
        with torch.no_grad():
            return kernel_lookup(ctx.coords)

    @staticmethod
    def backward(ctx, g):
        c = ctx.coords
        ctx.shared["uses"].append((lambda vol, c=c: _lookup(_pyramid(vol), c), g))
        return None, None, torch.zeros(1, device=g.device), None


class _Aggregate(torch.autograd.Function):
    """Motion aggregation with a fixed attention matrix: kernel forward (flash P.V, P never formed); backward
    needs P: it is recomputed ONCE per backward pass (cached in `shared`), the gradient w.r.t. the motion
    features and the aggregator's parameters is returned now, the gradient w.r.t. P is accumulated for the
    attention producer."""

    @staticmethod
    def forward(ctx, shared, kernel_fn, torch_fn, dense_fn, token, x, *params):
        ctx.shared, ctx.torch_fn, ctx.dense_fn, ctx.params = shared, torch_fn, dense_fn, params
        ctx.save_for_backward(x)
        with torch.no_grad():
            return kernel_fn(x.detach())

    @staticmethod
    def backward(ctx, gy):
        (x,), params = ctx.saved_tensors, ctx.params
        sh = ctx.shared
        if "P" not in sh:
            with torch.no_grad():
                sh["P"] = ctx.dense_fn()                       # [B,M,U,U] probabilities, recomputed once
        with torch.enable_grad():
            P = sh["P"].detach().requires_grad_(True)
            xr = x.detach().requires_grad_(ctx.needs_input_grad[5])
            y = ctx.torch_fn(xr, P)
            wanted = [P] + ([xr] if ctx.needs_input_grad[5] else []) + [p for p, need in zip(params, ctx.needs_input_grad[6:]) if need]
            grads = list(torch.autograd.grad(y, wanted, gy, allow_unused=True))
        gP = grads.pop(0)
        sh["dense_grad"] = gP if sh.get("dense_grad") is None else sh["dense_grad"] + gP
        gx = grads.pop(0) if ctx.needs_input_grad[5] else None
        gp = [grads.pop(0) if need else None for need in ctx.needs_input_grad[6:]]
        return (None, None, None, None, torch.zeros(1, device=gy.device), gx, *gp)


def _kernel_corr_build(cf, ws, fmap1, fmap2):
    """TransCorrBlock.update on an explicit workspace (one per batch element)."""
    with torch.cuda.device(fmap1.device):
        ops.pack_tokens(fmap1[0].float().contiguous(), ws.grid, ops.PACK_LN, out_b=ws.T1)
        ops.pack_tokens(fmap2[0].float().contiguous(), ws.grid, ops.PACK_LN, out_b=ws.T2f)
        cf.build_rows(ws, ws.T1, ws.T2f)


def _kernel_corr_lookup(cf, ws, coords):
    g = ws.grid
    with torch.cuda.device(coords.device):
        crow = torch.zeros((g.Mp, 2), dtype=torch.float32, device=coords.device)
        crow.view(g.H, g.Wp, 2)[:, :g.W] = coords[0].float().permute(1, 2, 0)
        out = torch.empty((1, 324, g.H, g.W), dtype=torch.float32, device=coords.device)
        cf.lookup_rows(ws, crow, out_nchw=out[0])
    return out


# ------------------------------------------------------------------------------------------------
# CRAFT.forward under grad
# ------------------------------------------------------------------------------------------------
def _dropout_active(model, cfg):
    """True -> the block runs the PyTorch restatement in forward too: dropout is live (no kernel form), or the
    model was told not to use the kernels under grad (`model.train_kernels = False`, a debugging switch)."""
    if not getattr(model, "train_kernels", True):
        return True
    return model.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0)


def forward_train(model, image1, image2, iters=12, flow_init=None, test_mode=0):
    """core/network.py:164-267 with autograd.  Returns what the reference returns for `test_mode`."""
    a = model.args
    amp = bool(getattr(a, "mixed_precision", False))
    training = model.training
    image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
    image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
    with torch.autocast("cuda", enabled=amp):
        fmap1, fmap2 = model.fnet([image1, image2])
    fmap1, fmap2 = fmap1.float(), fmap2.float()
    B, _, h, w = fmap1.shape

    # ---- F2 transformer ------------------------------------------------------------------------
    if a.f2trans != "none":
        f2 = model.f2_trans
        if _dropout_active(model, f2.config):
            with torch.autocast("cuda", enabled=amp):          # core/network.py:185-187
                fmap2 = _f2_trans(f2, fmap2, training).float()
        else:
            fmap2 = _KernelFwd.apply(lambda x: f2(x), lambda x: _f2_trans(f2, x, False), fmap2, *_params(f2))

    # ---- correlation volume (one kernel-side pyramid per batch element) ------------------------------
    if a.craft:
        cf = model.corr_fn
        if _dropout_active(model, cf.config):
            with torch.autocast("cuda", enabled=amp):          # core/network.py:227-228
                vol = _corr_volume(cf, fmap1, fmap2, training)
            pyr = _pyramid(vol.float())
            lookup = lambda c: _lookup(pyr, c)
        else:
            grid = TokenGrid(h, w)
            per = []
            for b in range(B):
                ws = model._workspaces.get(grid, fmap1.device, False, slot=b)
                shared = {"uses": []}
                token = _Producer.apply(shared, lambda f1, f2_, ws=ws: _kernel_corr_build(cf, ws, f1, f2_),
                                        lambda f1, f2_: _corr_volume(cf, f1, f2_, False), 2,
                                        fmap1[b:b + 1], fmap2[b:b + 1], *_params(cf))
                per.append((shared, ws, token))
            lookup = lambda c: torch.cat([_Lookup.apply(sh, lambda cc, ws=ws: _kernel_corr_lookup(cf, ws, cc), tok, c[b:b + 1])
                                          for b, (sh, ws, tok) in enumerate(per)], dim=0)
    else:
        pyr = _pyramid(_plain_volume(fmap1, fmap2))
        lookup = lambda c: _lookup(pyr, c)

    # ---- context, intra-frame attention -----------------------------------------------------------
    with torch.autocast("cuda", enabled=amp):
        cnet = model.cnet(image1)
    net, inp = torch.split(cnet.float(), [model.hidden_dim, model.context_dim], dim=1)
    net, inp = torch.tanh(net), torch.relu(inp)
    ub = model.update_block
    ag = ub.aggregator
    if a.use_setrans:
        att = model.att
        if _dropout_active(model, att.config):
            with torch.autocast("cuda", enabled=amp):          # core/network.py:213-214
                probs, _ = _self_att_probs(att, inp, training)
            aggregate = lambda m3: _expanded_feat_trans(ag, m3, probs)
        else:
            per_att = []
            for b in range(B):
                holder, shared = {}, {"uses": []}
                xb = inp[b:b + 1]
                token = _Producer.apply(shared, lambda x, holder=holder: holder.__setitem__("h", att(x)),
                                        lambda x: _self_att_probs(att, x, False)[0], 1, xb, *_params(att))
                dense = (lambda holder=holder, xb=xb: holder["h"].dense() if holder["h"].grid.U <= 4096
                         else _self_att_probs(att, xb.detach(), False)[0])
                per_att.append((shared, holder, token, dense))
            aggregate = lambda m3: torch.cat([
                _Aggregate.apply(sh, lambda x, ho=ho: ag(x, ho["h"]), lambda x, P: _expanded_feat_trans(ag, x, P), de, tok,
                                 m3[b:b + 1], *_params(ag))
                for b, (sh, ho, tok, de) in enumerate(per_att)], dim=0)
    else:
        attn = _gma_attention(model.att, inp)

    # ---- refinement loop -------------------------------------------------------------------------------
    from .utils.utils import coords_grid
    coords0 = coords_grid(B, h, w, device=image1.device)
    coords1 = coords0.clone()
    if flow_init is not None:
        coords1 = coords1 + flow_init
    flow_predictions = []
    enc, gru = ub.encoder, ub.gru
    for _ in range(iters):
        coords1 = coords1.detach()
        corr = lookup(coords1)
        flow = coords1 - coords0
        with torch.autocast("cuda", enabled=amp):
            cor = F.relu(enc.convc2(F.relu(enc.convc1(corr))))
            flo = F.relu(enc.convf2(F.relu(enc.convf1(flow))))
            motion = torch.cat([F.relu(enc.conv(torch.cat([cor, flo], dim=1))), flow], dim=1)
            if a.use_setrans:
                m3 = motion.float().reshape(B, 128, h * w).permute(0, 2, 1).contiguous()
                glob = aggregate(m3).reshape(B, h, w, 128).permute(0, 3, 1, 2)
            else:
                glob = _gma_aggregate(ag, attn, motion)
            x = torch.cat([inp, motion, glob], dim=1)
            for tag in ("1", "2"):
                hx = torch.cat([net, x], dim=1)
                z = torch.sigmoid(getattr(gru, "convz" + tag)(hx))
                r = torch.sigmoid(getattr(gru, "convr" + tag)(hx))
                q = torch.tanh(getattr(gru, "convq" + tag)(torch.cat([r * net, x], dim=1)))
                net = (1 - z) * net + z * q
            delta = ub.flow_head.conv2(F.relu(ub.flow_head.conv1(net)))
            up_mask = 0.25 * ub.mask(net)
        coords1 = coords1 + delta.float()
        flow_predictions.append(upsample_flow(coords1 - coords0, up_mask.float()))
    if test_mode == 1:
        return coords1 - coords0, flow_predictions[-1]
    if test_mode == 2:
        return coords1 - coords0, flow_predictions
    return flow_predictions
