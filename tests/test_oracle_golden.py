"""CPU tests: the oracle (oracle/restate.py + oracle/cpu_forward.py) reproduces the executed
reference.  Two pins:
  * frozen golden outputs (tests/golden/*.pt) -- runs everywhere;
  * the live reference, when /root/reference is mounted (build container only)."""
import os

import pytest
import torch

from oracle import cpu_forward, restate as R
from oracle.ref_loader import craft_args, reference_available, synthetic_pair

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _seeded_sd(kw, tweak=None):
    from craft_b200.network import CRAFT
    torch.manual_seed(1234)
    sd = {k: v.clone() for k, v in CRAFT(craft_args(**kw)).state_dict().items()}
    for k, v in (tweak or {}).items():
        sd[k].fill_(v)
    return sd


@pytest.mark.parametrize("name,flags", [
    ("seeded_setrans_128", dict(craft=True, use_setrans=True, f2trans=True)),
    ("seeded_gma_128", dict(craft=True, use_setrans=False, f2trans=True)),
    ("seeded_plain_128", dict(craft=False, use_setrans=False, f2trans=False)),
    # branches the default configuration never takes: the data-dependent clamp (all three attentions / two of
    # three), the --f2radius key mask, two modes with a negative soft-aggregation weight
    ("seeded_clip02_128", dict(attn_clip=0.2)),
    ("seeded_clip03_128", dict(attn_clip=0.3)),
    ("seeded_f2radius_128", dict(f2_mask_radius=5)),
    ("seeded_modes2_128", dict(M_inter=2, M_intra=2)),
])
def test_oracle_forward_matches_golden(name, flags):
    rec = torch.load(os.path.join(GOLD, name + ".pt"), map_location="cpu")
    sd = _seeded_sd(rec["args"], rec.get("tweak"))
    i1, i2 = synthetic_pair(rec["H"], rec["W"])
    diag = {}
    with torch.no_grad():
        lo, up = cpu_forward.craft_forward(sd, i1, i2, iters=rec["iters"], diag=diag, **flags)
    assert torch.allclose(lo[0], rec["flow_lo"], atol=2e-4), (lo[0] - rec["flow_lo"]).abs().max()
    assert torch.allclose(up[0], rec["flow_up"], atol=1e-3), (up[0] - rec["flow_up"]).abs().max()
    # the reference's own diagnostics: global score maximum and whether the clamp branch was taken
    for mod, d in rec.get("diag", {}).items():
        assert abs(diag[mod] - d["max_attn"]) < 1e-4, (mod, diag[mod], d)
        clip = flags.get("attn_clip", 100.0)
        assert (diag[mod] > clip) == (d["clamp_count"] == 1), (mod, d)


def test_oracle_seams_match_golden():
    rec = torch.load(os.path.join(GOLD, "seeded_setrans_128.pt"), map_location="cpu")
    sd = _seeded_sd(rec["args"])
    fn = rec["fnet_out"]
    with torch.no_grad():
        probs, tok, _ = R.self_attention_probs(fn[1:2], sd["f2_trans.setrans.query.weight"],
                                               sd["f2_trans.setrans.key.weight"], 4,
                                               sd["f2_trans.vispos_encoder.pos_coder.biases"], 0.5)
        y = R.expanded_feat_trans(tok, probs, sd["f2_trans.setrans.out_trans.first_linear.weight"],
                                  sd["f2_trans.setrans.out_trans.feat_softaggr.feat2score.weight"],
                                  sd["f2_trans.setrans.out_trans.feat_softaggr.feat2score.bias"],
                                  sd["f2_trans.setrans.out_trans.input_skip_coeff"], 4)
        f2 = y.permute(0, 2, 1).reshape(fn[1:2].shape)
        assert torch.allclose(f2, rec["f2_out"], atol=1e-4)
        vol, _, _ = R.trans_corr_volume(fn[0:1], f2, sd["corr_fn.setrans.query.weight"], sd["corr_fn.setrans.query.bias"],
                                        sd["corr_fn.setrans.attn_softaggr.feat2score.weight"].reshape(()),
                                        sd["corr_fn.setrans.attn_softaggr.feat2score.bias"].reshape(()),
                                        sd["corr_fn.vispos_encoder.pos_coder.biases"], 4, 0.5)
        corr = R.corr_lookup(R.corr_pyramid(vol), R.coords_grid(1, 16, 16))
        assert torch.allclose(corr[0], rec["ub_it0.corr"], atol=2e-4), (corr[0] - rec["ub_it0.corr"]).abs().max()
        up = R.upsample_flow(rec["ub_it0.delta"][None], rec["ub_it0.mask"][None])
        assert torch.allclose(up[0], rec["flow_up_first"], atol=1e-4)


def test_pos_bias_restatement_is_the_sliding_window():
    t = torch.arange(225.).reshape(15, 15)
    b = R.sliding_pos_bias(t, 9, 11).reshape(9, 11, 9, 11)
    assert b[4, 5, 4, 5] == t[7, 7] and b[0, 0, 7, 7] == t[14, 14] and b[0, 0, 8, 0] == 0 and b[8, 10, 1, 3] == t[0, 0]


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference_checkpoint():
    from oracle.ref_loader import build_reference_model
    model, _ = build_reference_model()
    sd = {k: v for k, v in model.state_dict().items()}
    i1, i2 = synthetic_pair(128, 160, seed=5)   # smallest legal grid: a 1-wide pyramid level gives NaN in the reference
    with torch.no_grad():
        lo_r, up_r = model(i1, i2, iters=3, test_mode=1)
        lo, up = cpu_forward.craft_forward(sd, i1, i2, iters=3)
    assert torch.allclose(up, up_r, atol=1e-3), (up - up_r).abs().max()
