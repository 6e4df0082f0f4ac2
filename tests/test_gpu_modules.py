"""Every reference-facing nn.Module of craft_b200 called STANDALONE through its own forward() /
update() / __call__ / upsample_flow, against the oracle restatement (oracle/restate.py, pinned to the
executed reference by tests/test_oracle_golden.py) on the same seeded weights and inputs.

These are the signatures north_star asks to keep (SURVEY.md section 8b "inner seams"); inside
CRAFT.forward the same kernels run fused on workspace buffers (tests/test_gpu_e2e.py covers that)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from craft_b200.network import CRAFT                       # noqa: E402
from craft_b200.setrans import AttentionHandle, PosBiasHandle   # noqa: E402
from oracle import restate as R                            # noqa: E402
from oracle.ref_loader import craft_args                   # noqa: E402

DEV = "cuda"
H8, W8 = 16, 24            # token grid of the standalone tests (a 128 x 192 image)


def bf16r(t):
    return t.to(torch.bfloat16).float()


def _model(**kw):
    torch.manual_seed(1234)
    m = CRAFT(craft_args(**kw))
    # spread the (zero-initialised) positional-bias tables and the GMA gamma so that every term is visible
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("pos_coder.biases"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            if name.endswith("aggregator.gamma"):
                p.fill_(0.5)
    return m.to(DEV).eval()


def _close(got, ref, rel, name=""):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rel * max(scale, 1e-3), "%s: max abs err %.4g vs scale %.4g" % (name, err, scale)


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(shape, device=DEV, generator=g) * scale


# ------------------------------------------------------------------------------------------------
def test_learned_soft_aggregate_forward():
    """core/setrans.py:289-300, both flavours, batch of 2."""
    from craft_b200.setrans import LearnedSoftAggregate
    torch.manual_seed(3)
    a1 = LearnedSoftAggregate(1, group_dim=1, keepdim=True).to(DEV)
    a1.feat2score.weight.data.fill_(-0.7)
    a1.feat2score.bias.data.fill_(0.3)
    x = _rand((2, 4, 50, 60), 1, 2.0)
    with torch.no_grad():
        got = a1(x)
    ref = R.soft_aggregate_scalar(x, a1.feat2score.weight.reshape(()), a1.feat2score.bias.reshape(()))
    assert got.shape == ref.shape == (2, 1, 50, 60)
    _close(got, ref, 1e-5, "scalar")
    aF = LearnedSoftAggregate(128, group_dim=1, keepdim=False).to(DEV)
    y = _rand((2, 4, 70, 128), 2)
    with torch.no_grad():
        got = aF(y)
    ref = R.soft_aggregate_feat(y, aF.feat2score.weight, aF.feat2score.bias)
    assert got.shape == ref.shape == (2, 70, 128)
    _close(got, ref, 1e-5, "feat")


def test_input_feat_encoder_forward():
    """SETransInputFeatEncoder.forward core/setrans.py:763-800 (+ the PosBiasHandle it returns)."""
    m = _model()
    x = _rand((2, 256, H8, W8), 5, 3.0) + 0.5
    with torch.no_grad():
        tok, pb = m.corr_fn.vispos_encoder(x, None, return_pos_biases=True)
    assert isinstance(pb, PosBiasHandle) and pb.shape == (H8, W8)
    _close(tok, R.encode_tokens(x), 1e-5, "tokens")


@pytest.mark.parametrize("which", ["corr_fn", "f2_trans", "att"])
def test_cross_att_feat_trans_forward(which):
    """CrossAttFeatTrans.forward core/setrans.py:501-566 in its three output modes (scores-only /
    features / probabilities-only), called on encoder tokens exactly as its owners do."""
    m = _model()
    owner = getattr(m, which)
    st, enc = owner.setrans, owner.vispos_encoder
    Cc = st.in_feat_dim
    x = _rand((1, Cc, H8, W8), 7)
    table = enc.pos_coder.biases.detach()
    bias = R.sliding_pos_bias(table, H8, W8)[None, None]
    with torch.no_grad():
        tok, pb = enc(x, None, return_pos_biases=True)
        tok2 = R.encode_tokens(_rand((1, Cc, H8, W8), 8)) if which == "corr_fn" else None
        out = st(tok, tok2, pos_biases=pb)
        wq, wk = st.query.weight, st.key.weight
        bq = st.query.bias
        s, gmax = R.mode_scores(bf16r(tok), bf16r(tok2 if tok2 is not None else tok), bf16r(wq), bq, bf16r(wk),
                                st.key.bias, st.num_modes, bias, st.pos_code_weight)
        if which == "corr_fn":
            ref = R.soft_aggregate_scalar(s, st.attn_softaggr.feat2score.weight.reshape(()),
                                          st.attn_softaggr.feat2score.bias.reshape(()))
            assert out.shape == ref.shape == (1, 1, H8 * W8, H8 * W8)
            _close(out, ref, 2e-2, "aggregated scores")
        elif which == "att":
            assert isinstance(out, AttentionHandle)
            probs = out.dense()
            ref = torch.softmax(s, dim=-1)
            assert probs.shape == ref.shape
            _close(probs, ref, 3e-2, "probabilities")
            assert (probs.sum(-1) - 1).abs().max() < 2e-3
        else:
            ot = st.out_trans
            ref = R.expanded_feat_trans(bf16r(tok), torch.softmax(s, dim=-1), bf16r(ot.first_linear.weight),
                                        ot.feat_softaggr.feat2score.weight, ot.feat_softaggr.feat2score.bias,
                                        ot.input_skip_coeff, st.num_modes)
            assert out.shape == ref.shape == (1, H8 * W8, 256)
            _close(out, ref, 3e-2, "features")
    assert abs(st.max_attn - gmax) <= 2e-2 * max(1.0, gmax) and st.clamp_count == 0


@pytest.mark.parametrize("radius", [-1, 3])
def test_self_att_vis_pos_trans_forward(radius):
    """SelfAttVisPosTrans.forward core/setrans.py:578-619: feature output (F2 transformer, with and without
    the --f2radius mask) and probabilities-only output (intra-frame attention), batch of 2."""
    m = _model(f2_attn_mask_radius=radius)
    x = _rand((2, 256, H8, W8), 11)
    f2 = m.f2_trans
    ot = f2.setrans.out_trans
    with torch.no_grad():
        got = f2(x)
        probs, tok, _ = R.self_attention_probs(x, bf16r(f2.setrans.query.weight), bf16r(f2.setrans.key.weight), 4,
                                               f2.vispos_encoder.pos_coder.biases, 0.5, mask_radius=radius)
        y = R.expanded_feat_trans(bf16r(tok), probs, bf16r(ot.first_linear.weight), ot.feat_softaggr.feat2score.weight,
                                  ot.feat_softaggr.feat2score.bias, ot.input_skip_coeff, 4)
    ref = y.permute(0, 2, 1).reshape(x.shape)
    assert got.shape == x.shape
    _close(got, ref, 3e-2, "f2_trans features")
    if radius > 0:
        return
    inp = _rand((2, 128, H8, W8), 12).relu()
    with torch.no_grad():
        handles = m.att(inp)
        ref, _, _ = R.self_attention_probs(inp, bf16r(m.att.setrans.query.weight), bf16r(m.att.setrans.key.weight), 4,
                                           m.att.vispos_encoder.pos_coder.biases, 1.0)
    assert isinstance(handles, list) and len(handles) == 2
    for b in range(2):
        _close(handles[b].dense()[0], ref[b], 3e-2, "att probabilities %d" % b)


def test_expanded_feat_trans_forward():
    """ExpandedFeatTrans.forward core/setrans.py:364-410 as the motion aggregator: (tokens, attention) -> tokens."""
    m = _model()
    inp = _rand((1, 128, H8, W8), 13).relu()
    motion = _rand((1, H8 * W8, 128), 14)
    ag = m.update_block.aggregator
    with torch.no_grad():
        att = m.att(inp)
        got = ag(motion, att)
        probs, _, _ = R.self_attention_probs(inp, bf16r(m.att.setrans.query.weight), bf16r(m.att.setrans.key.weight), 4,
                                             m.att.vispos_encoder.pos_coder.biases, 1.0)
        ref = R.expanded_feat_trans(bf16r(motion), probs, bf16r(ag.first_linear.weight), ag.feat_softaggr.feat2score.weight,
                                    ag.feat_softaggr.feat2score.bias, ag.input_skip_coeff, 4)
    _close(got, ref, 3e-2, "aggregator")
    with pytest.raises(TypeError):
        ag(motion, torch.zeros(1, 4, 8, 8, device=DEV))        # dense probabilities are not an input format here


def test_gma_attention_and_aggregate_forward():
    """gma.Attention.forward core/gma.py:74-102 and gma.Aggregate.forward core/gma.py:128-142 with gamma != 0."""
    m = _model(use_setrans=False)
    fmap = _rand((2, 128, H8, W8), 15)
    motion = _rand((2, 128, H8, W8), 16)
    ag = m.update_block.aggregator
    assert ag.gamma.item() == 0.5
    with torch.no_grad():
        att = m.att(fmap)
        got = ag(att, motion)
        attn = R.gma_attention(bf16r(fmap), bf16r(m.att.to_qk.weight))
        ref = R.gma_aggregate(attn, bf16r(motion), bf16r(ag.to_v.weight), ag.gamma)
        # the aggregated part alone must be resolved, not only fmap + small correction
        part_ref = ref - bf16r(motion)
    _close(att[0].dense()[0], attn[0], 3e-2, "gma attention")
    _close(got - bf16r(motion), part_ref, 4e-2, "gamma * attn @ v")
    _close(got, ref, 2e-2, "aggregate")


def _update_params(m):
    return {k: v for k, v in m.update_block.state_dict().items()}


def test_motion_encoder_forward():
    """BasicMotionEncoder.forward core/update.py:79-87, batch of 2."""
    m = _model()
    enc = m.update_block.encoder
    flow = _rand((2, 2, H8, W8), 17, 2.0)
    corr = _rand((2, 324, H8, W8), 18)
    P = {k: (bf16r(v) if k.endswith("weight") and "convf1" not in k else v) for k, v in enc.state_dict().items()}
    with torch.no_grad():
        got = enc(flow, corr)
        ref = R.motion_encoder(flow, bf16r(corr), P)
    assert got.shape == ref.shape == (2, 128, H8, W8)
    _close(got[:, :126], ref[:, :126], 3e-2, "motion features")
    _close(got[:, 126:], bf16r(flow), 1e-6, "flow channels")      # X is a bf16 buffer


def test_sep_conv_gru_forward():
    """SepConvGRU.forward core/update.py:49-64 (both passes), batch of 2."""
    m = _model()
    gru = m.update_block.gru
    h = torch.tanh(_rand((2, 128, H8, W8), 19))
    x = _rand((2, 384, H8, W8), 20)
    P = {k: (bf16r(v) if k.endswith("weight") else v) for k, v in gru.state_dict().items()}
    with torch.no_grad():
        got = gru(h, x)
        ref = R.sep_conv_gru(h, bf16r(x), P)
    assert got.shape == ref.shape
    _close(got, ref, 2e-2, "hidden state")


def test_flow_head_forward():
    """FlowHead.forward core/update.py:15-16."""
    m = _model()
    fh = m.update_block.flow_head
    x = _rand((2, 128, H8, W8), 21)
    with torch.no_grad():
        got = fh(x)
        hid = F.relu(F.conv2d(bf16r(x), bf16r(fh.conv1.weight), fh.conv1.bias, padding=1))
        ref = F.conv2d(bf16r(hid), bf16r(fh.conv2.weight), fh.conv2.bias, padding=1)
    assert got.shape == (2, 2, H8, W8)
    _close(got, ref, 1e-2, "delta flow")


@pytest.mark.parametrize("use_setrans", [True, False])
def test_gma_update_block_forward(use_setrans):
    """GMAUpdateBlock.forward core/update.py:137-162: (net, inp, corr, flow, attention) -> (net, mask, delta)."""
    m = _model(use_setrans=use_setrans)
    ub = m.update_block
    net = torch.tanh(_rand((1, 128, H8, W8), 22))
    inp = _rand((1, 128, H8, W8), 23).relu()
    corr = _rand((1, 324, H8, W8), 24)
    flow = _rand((1, 2, H8, W8), 25, 2.0)
    sd = {k: v for k, v in ub.state_dict().items()}
    sub = lambda pre: {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    with torch.no_grad():
        att = m.att(inp)
        net_o, mask_o, delta_o = ub(net, inp, corr, flow, att)
        motion = R.motion_encoder(flow, corr, sub("encoder."))
        if use_setrans:
            probs, _, _ = R.self_attention_probs(inp, m.att.setrans.query.weight, m.att.setrans.key.weight, 4,
                                                 m.att.vispos_encoder.pos_coder.biases, 1.0)
            glob = R.expanded_feat_trans(motion.reshape(1, 128, -1).permute(0, 2, 1), probs, sd["aggregator.first_linear.weight"],
                                         sd["aggregator.feat_softaggr.feat2score.weight"],
                                         sd["aggregator.feat_softaggr.feat2score.bias"], sd["aggregator.input_skip_coeff"], 4)
            glob = glob.reshape(1, H8, W8, 128).permute(0, 3, 1, 2)
        else:
            glob = R.gma_aggregate(R.gma_attention(inp, m.att.to_qk.weight), motion, sd["aggregator.to_v.weight"],
                                   sd["aggregator.gamma"])
        net_r = R.sep_conv_gru(net, torch.cat([inp, motion, glob], 1), sub("gru."))
        delta_r, mask_r = R.flow_and_mask_heads(net_r, sd)
    _close(net_o, net_r, 3e-2, "net")
    _close(delta_o, delta_r, 5e-2, "delta")
    _close(mask_o, mask_r, 5e-2, "mask")


def test_trans_corr_block_update_and_call():
    """TransCorrBlock.update core/corr.py:148-207 then __call__ core/corr.py:47-71 at shifted coordinates."""
    m = _model()
    cf = m.corr_fn
    f1 = _rand((1, 256, H8, W8), 26, 2.0)
    f2 = _rand((1, 256, H8, W8), 27, 2.0)
    coords = R.coords_grid(1, H8, W8, DEV) + _rand((1, 2, H8, W8), 28, 3.0)
    st = cf.setrans
    with torch.no_grad():
        cf.update(f1, f2, None, None, coords)
        got = cf(coords)
        vol, _, _ = R.trans_corr_volume(f1, f2, st.query.weight, st.query.bias, st.attn_softaggr.feat2score.weight.reshape(()),
                                        st.attn_softaggr.feat2score.bias.reshape(()), cf.vispos_encoder.pos_coder.biases, 4, 0.5)
        ref = R.corr_lookup(R.corr_pyramid(vol), coords)
    assert got.shape == ref.shape == (1, 324, H8, W8)
    assert (got - ref).abs().mean().item() <= 2e-2 and (got - ref).abs().max().item() <= 0.25


def test_plain_corr_block_and_call():
    """CorrBlock(fmap1, fmap2) core/corr.py:16-45 + __call__ :47-71 (RAFT / GMA baselines; BASELINE configs[0])."""
    from craft_b200.corr import CorrBlock
    f1 = bf16r(_rand((1, 256, H8, W8), 29))
    f2 = bf16r(_rand((1, 256, H8, W8), 30))
    coords = R.coords_grid(1, H8, W8, DEV) + _rand((1, 2, H8, W8), 31, 2.0)
    with torch.no_grad():
        cb = CorrBlock(f1, f2, num_levels=4, radius=4)
        got = cb(coords)
        ref = R.corr_lookup(R.corr_pyramid(R.plain_corr_volume(f1, f2)), coords)
    _close(got, ref, 1e-2, "plain lookup")     # bf16 only on the output-free path: fp32 accumulate of exact products


def test_upsample_flow_method():
    """CRAFT.upsample_flow core/network.py:151-162 on NCHW tensors, batch of 2."""
    m = _model()
    flow = _rand((2, 2, H8, W8), 32, 3.0)
    mask = _rand((2, 576, H8, W8), 33)
    with torch.no_grad():
        got = m.upsample_flow(flow, mask)
    _close(got, R.upsample_flow(flow, mask), 1e-5, "convex upsampling")


def test_standalone_handles_own_their_buffers():
    """A handle returned by a standalone attention call must survive later calls on the same grid
    (the shared workspace is overwritten)."""
    m = _model()
    a = _rand((1, 128, H8, W8), 34).relu()
    b = _rand((1, 128, H8, W8), 35).relu()
    with torch.no_grad():
        ha = m.att(a)
        pa = ha.dense().clone()
        m.att(b)
        m.f2_trans(_rand((1, 256, H8, W8), 36))
        assert torch.equal(ha.dense(), pa)


def test_modules_follow_their_input_device_when_another_is_current():
    """Per-device library state + device guards: run a model living on the LAST visible GPU while cuda:0 is
    current (a single-GPU box degenerates to the plain case)."""
    n = torch.cuda.device_count()
    dev = torch.device("cuda", n - 1)
    torch.manual_seed(1234)
    m = CRAFT(craft_args()).to(dev).eval()
    from oracle.ref_loader import synthetic_pair
    i1, i2 = synthetic_pair(128, 128)
    with torch.no_grad(), torch.cuda.device(0):
        lo, up = m(i1.to(dev), i2.to(dev), iters=2, test_mode=1)
        torch.cuda.synchronize(dev)
        m0 = CRAFT(craft_args())
        m0.load_state_dict(m.state_dict())
        lo0, up0 = m0.to("cuda:0").eval()(i1.to("cuda:0"), i2.to("cuda:0"), iters=2, test_mode=1)
    assert up.device == dev and torch.isfinite(up).all()
    assert (up.cpu() - up0.cpu()).abs().max().item() <= 1e-3


def test_data_parallel_on_two_gpus_matches_single_device():
    """train.py:183 / evaluate.py:1534 wrap the model in nn.DataParallel: with a batch of two on a 2-GPU box every
    replica (one thread per device, shared module attributes) must reproduce the single-device result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from oracle.ref_loader import smooth_pair, synthetic_pair
    torch.manual_seed(1234)
    m = CRAFT(craft_args()).cuda().eval()
    a1, a2 = synthetic_pair(128, 128)
    b1, b2 = smooth_pair(128, 128)
    i1, i2 = torch.cat([a1, b1]).cuda(), torch.cat([a2, b2]).cuda()
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    with torch.no_grad():
        for _ in range(2):                       # second call: packed-weight caches of both devices are warm
            lo, up = dp(i1, i2, iters=3, test_mode=1)
        lo_a, up_a = m(i1[:1], i2[:1], iters=3, test_mode=1)
        lo_b, up_b = m(i1[1:], i2[1:], iters=3, test_mode=1)
    assert up.shape == (2, 2, 128, 128) and up.device.index == 0
    assert (up[0] - up_a[0]).abs().max().item() <= 2e-3 and (up[1] - up_b[0]).abs().max().item() <= 2e-3
