"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (craft_b200/).

Full CRAFT forward as a functional fp32 program over a state dict, built from oracle/restate.py
(hot path) plus a functional restatement of the BasicEncoder (core/extractor.py:124-196).  It
follows CRAFT.forward core/network.py:164-267 for the three model variants the drivers can build
(setrans, gma aggregator, plain CorrBlock).  Used as
  * the parity checker of __graft_entry__.smoke() and tests/, and
  * the CPU baseline / `--impl reference` arm of bench.py (kind "port": the reference is a Python
    tree that cannot travel to the GPU box, this port can).
Pinned against the executed reference by tests/test_oracle_golden.py.
"""
import torch
import torch.nn.functional as F

from . import restate as R


def _norm(x, sd, prefix, kind):
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


def _res_block(x, sd, p, kind, stride):
    y = F.relu(_norm(F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=stride, padding=1), sd, p + ".norm1", kind))
    y = F.relu(_norm(F.conv2d(y, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1), sd, p + ".norm2", kind))
    if stride != 1:
        x = _norm(F.conv2d(x, sd[p + ".downsample.0.weight"], sd[p + ".downsample.0.bias"], stride=stride), sd, p + ".norm3", kind)
    return F.relu(x + y)


def basic_encoder(x, sd, p, kind):
    """core/extractor.py:173-196 (eval mode)."""
    x = F.relu(_norm(F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], stride=2, padding=3), sd, p + ".norm1", kind))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = _res_block(x, sd, "%s.layer%d.0" % (p, li), kind, stride)
        x = _res_block(x, sd, "%s.layer%d.1" % (p, li), kind, 1)
    return F.conv2d(x, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"])


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def craft_forward(sd, image1, image2, iters=12, flow_init=None, craft=True, use_setrans=True, f2trans=True,
                  M=4, w_inter=0.5, w_f2=0.5, w_intra=1.0, return_all=False, attn_clip=100.0, f2_mask_radius=-1,
                  M_inter=None, M_intra=None, diag=None):
    """-> (flow_lo [B,2,h,w], flow_up [B,2,H,W]) like test_mode=1 (or the list of all flow_up).
    M_inter / M_intra: --inter_num_modes / --intra_num_modes when they differ from M (the F2 transformer keeps M);
    diag (optional dict) receives the global score maxima the reference keeps as max_attn (core/setrans.py:520-529)."""
    M_inter = M_inter or M
    M_intra = M_intra or M
    diag = diag if diag is not None else {}
    image1 = 2 * (image1 / 255.0) - 1.0
    image2 = 2 * (image2 / 255.0) - 1.0
    B = image1.shape[0]
    fm = basic_encoder(torch.cat([image1, image2], 0), sd, "fnet", "instance")
    fmap1, fmap2 = fm[:B], fm[B:]
    if f2trans:      # SelfAttVisPosTrans "F2 transformer" core/network.py:185-187, setrans.py:578-619
        probs, tok, diag["f2_trans"] = R.self_attention_probs(
            fmap2, sd["f2_trans.setrans.query.weight"], sd["f2_trans.setrans.key.weight"], M,
            sd["f2_trans.vispos_encoder.pos_coder.biases"], w_f2, attn_clip, f2_mask_radius)
        y = R.expanded_feat_trans(tok, probs, sd["f2_trans.setrans.out_trans.first_linear.weight"],
                                  sd["f2_trans.setrans.out_trans.feat_softaggr.feat2score.weight"],
                                  sd["f2_trans.setrans.out_trans.feat_softaggr.feat2score.bias"],
                                  sd["f2_trans.setrans.out_trans.input_skip_coeff"], M)
        fmap2 = y.permute(0, 2, 1).reshape(fmap2.shape)
        del probs
    if craft:        # TransCorrBlock.update core/corr.py:148-207
        vol, _, diag["corr_fn"] = R.trans_corr_volume(
            fmap1, fmap2, sd["corr_fn.setrans.query.weight"], sd["corr_fn.setrans.query.bias"],
            sd["corr_fn.setrans.attn_softaggr.feat2score.weight"].reshape(()),
            sd["corr_fn.setrans.attn_softaggr.feat2score.bias"].reshape(()),
            sd["corr_fn.vispos_encoder.pos_coder.biases"], M_inter, w_inter, attn_clip)
    else:            # CorrBlock core/corr.py:16-45
        vol = R.plain_corr_volume(fmap1, fmap2)
    pyramid = R.corr_pyramid(vol)
    del vol
    cnet = basic_encoder(image1, sd, "cnet", "batch")
    net, inp = torch.tanh(cnet[:, :128]), torch.relu(cnet[:, 128:])
    if use_setrans:  # intra-frame attention core/network.py:214
        attn, _, diag["att"] = R.self_attention_probs(inp, sd["att.setrans.query.weight"], sd["att.setrans.key.weight"],
                                                      M_intra, sd["att.vispos_encoder.pos_coder.biases"], w_intra, attn_clip)
    else:
        attn = R.gma_attention(inp, sd["att.to_qk.weight"])
    _, _, h, w = net.shape
    coords0 = R.coords_grid(B, h, w, image1.device)
    coords1 = coords0.clone()
    if flow_init is not None:
        coords1 = coords1 + flow_init
    P = _sub(sd, "update_block.")
    enc, gru = _sub(P, "encoder."), _sub(P, "gru.")
    ups = []
    for _ in range(iters):
        corr = R.corr_lookup(pyramid, coords1)
        flow = coords1 - coords0
        motion = R.motion_encoder(flow, corr, enc)
        if use_setrans:   # GMAUpdateBlock.forward core/update.py:141-148
            m3 = motion.reshape(B, 128, h * w).permute(0, 2, 1)
            glob = R.expanded_feat_trans(m3, attn, P["aggregator.first_linear.weight"],
                                         P["aggregator.feat_softaggr.feat2score.weight"],
                                         P["aggregator.feat_softaggr.feat2score.bias"],
                                         P["aggregator.input_skip_coeff"], M_intra)
            glob = glob.reshape(B, h, w, 128).permute(0, 3, 1, 2)
        else:
            glob = R.gma_aggregate(attn, motion, P["aggregator.to_v.weight"], P["aggregator.gamma"])
        net = R.sep_conv_gru(net, torch.cat([inp, motion, glob], 1), gru)
        delta, mask = R.flow_and_mask_heads(net, P)
        coords1 = coords1 + delta
        ups.append(R.upsample_flow(coords1 - coords0, mask))
    if return_all:
        return coords1 - coords0, ups
    return coords1 - coords0, ups[-1]
