"""CPU tests of the host-side mirror of the reference interface: constructor / state-dict contract,
config mutation, weight packing, geometry helpers, loud failure without CUDA."""
import argparse
import os

import pytest
import torch

from craft_b200 import ops
from craft_b200.network import CRAFT, RAFTER
from craft_b200.ops import TokenGrid
from oracle.ref_loader import craft_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference checkpoint's hot-path keys, shapes included
EXPECTED = {
    "corr_fn.setrans.query.weight": (256, 256), "corr_fn.setrans.key.bias": (256,),
    "corr_fn.setrans.attn_softaggr.feat2score.weight": (1, 1),
    "corr_fn.vispos_encoder.pos_coder.biases": (15, 15),
    "f2_trans.setrans.out_trans.first_linear.weight": (1024, 256),
    "f2_trans.setrans.out_trans.feat_softaggr.feat2score.weight": (1, 256),
    "f2_trans.setrans.out_trans.input_skip_coeff": (1,),
    "att.setrans.query.weight": (128, 128), "att.setrans.attn_softaggr.feat2score.bias": (1,),
    "update_block.encoder.convc1.weight": (256, 324, 1, 1), "update_block.encoder.conv.weight": (126, 256, 3, 3),
    "update_block.gru.convz1.weight": (128, 512, 1, 5), "update_block.gru.convq2.weight": (128, 512, 5, 1),
    "update_block.flow_head.conv2.weight": (2, 256, 3, 3), "update_block.mask.2.weight": (576, 256, 1, 1),
    "update_block.aggregator.first_linear.weight": (512, 128),
    "update_block.aggregator.feat_softaggr.feat2score.weight": (1, 128),
    "fnet.conv1.weight": (64, 3, 7, 7), "cnet.layer2.0.downsample.1.running_mean": (96,),
}


def test_state_dict_contract():
    args = craft_args()
    m = CRAFT(args)
    sd = m.state_dict()
    assert len(sd) == 202 and sum(v.numel() for v in sd.values()) == 6377020
    for k, shp in EXPECTED.items():
        assert tuple(sd[k].shape) == shp, k
    # tied Q/K of the correlation transformer (tie_qk_scheme 'shared', network.py:53)
    assert m.corr_fn.setrans.key.weight is m.corr_fn.setrans.query.weight
    # ctor mutates args exactly like the reference (network.py:33,57,92,106,127)
    assert args.corr_levels == 4 and args.corr_multiplier == 1
    assert args.inter_trans_config.out_attn_scores_only and args.intra_trans_config.out_attn_probs_only
    assert RAFTER is CRAFT


def test_trained_checkpoint_loads_strict():
    path = os.path.join(ROOT, "tests", "golden", "_local", "craft-sintel-model.pth")
    if not os.path.isfile(path):
        pytest.skip("untracked local copy of craft-sintel.pth not present")
    m = CRAFT(craft_args())
    res = m.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_variants_construct_and_reject_out_of_scope_flags():
    CRAFT(craft_args(use_setrans=False))
    CRAFT(craft_args(craft=False, use_setrans=False, f2trans="none", corr_multiplier=1))
    with pytest.raises(NotImplementedError):
        CRAFT(craft_args(f1trans="shared"))
    m = CRAFT(craft_args(f2_attn_mask_radius=16))          # --f2radius (core/setrans.py:580-584) is supported
    assert m.f2_trans.attn_mask_radius == 16 and m.att.attn_mask_radius == -1
    with pytest.raises(NotImplementedError):
        CRAFT(craft_args(intra_pos_code_type="lsinu"))


def test_forward_refuses_cpu_and_bad_shapes():
    m = CRAFT(craft_args()).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU path"):
            m(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))     # grad mode (training path) is CUDA-only too


def test_geometry_and_weight_packing():
    g = TokenGrid(55, 128)
    assert (g.Wp, g.Mp, g.U) == (130, 55 * 130, 7040)
    assert g.level_shapes() == [(55, 128), (27, 64), (13, 32), (6, 16)]      # floor-mode pooling, corr.py:186-189
    assert ops.conv_taps(1, 5, g) == [-2, -1, 0, 1, 2]
    assert ops.conv_taps(3, 3, g) == [-131, -130, -129, -1, 0, 1, 129, 130, 131]
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_conv_weight(w, Npad=32).float().reshape(9, 32, 64)
    assert torch.equal(p[4, :2, :3], w[:, :, 1, 1]) and p[:, 2:].abs().sum() == 0 and p[:, :, 3:].abs().sum() == 0
    perm = [2, 0, 1]
    p2 = ops.pack_conv_weight(w, Npad=32, cin_perm=perm).float().reshape(9, 32, 64)
    assert torch.equal(p2[0, :2, :3], w[:, perm, 0, 0])
    assert ops.blocked_keys(g, 128) == 7 * 8 * 128 and ops.blocked_keys(g, 64) == 7 * 16 * 64


def test_padded_flat_layout_and_conv64_weight_packing():
    """Host side of craft_conv3x3_c64 (include/craft_b200.h): the padded-flat row index, the tap-major weight layout
    (identical to the update block's pack_conv_weight), the folded-norm scale, and CRAFT.on_lane bookkeeping."""
    N, H, W = 2, 5, 7
    assert ops.PadAct.rows(N, H, W) == N * (H + 1) * (W + 2)
    x = torch.arange(N * H * W * 64, dtype=torch.float32).reshape(N, H, W, 64)
    t = torch.zeros((N, H + 1, W + 2, 64))
    t[:, :H, :W] = x
    pa = ops.PadAct(t.reshape(-1, 64), N, H, W)
    assert torch.equal(pa.dense(), x)
    n, y, xx = 1, 3, 6
    assert torch.equal(pa.t[(n * (H + 1) + y) * (W + 2) + xx], x[n, y, xx])          # row (n*(H+1)+y)*(W+2)+x
    w = torch.randn(64, 64, 3, 3)
    wp = ops.pack_conv64_weight(w)
    assert wp.shape == (576, 64) and wp.dtype == torch.float16
    with ops.precision(torch.float16):
        assert torch.equal(wp, ops.pack_conv_weight(w))                                # same layout as the shift-GEMM weights
    ky, kx, co, ci = 2, 0, 17, 40
    assert wp[(ky * 3 + kx) * 64 + co, ci] == w[co, ci, ky, kx].half()
    a = torch.rand(64) + 0.5
    assert torch.equal(ops.pack_conv64_weight(w, scale=a), ops.pack_conv64_weight(w * a.view(-1, 1, 1, 1)))
    # lanes: a context manager that only selects which workspace / graph a forward uses
    m = CRAFT(craft_args())
    assert m._lane == 0 and m._ws_slot() == 0
    with m.on_lane(2):
        assert m._lane == 2 and m._ws_slot() == ("lane", 2)
        with m.on_lane(0):
            assert m._ws_slot() == 0
        assert m._lane == 2
    assert m._lane == 0
