"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (craft_b200/).

Loads the *unmodified* reference (askerlee/craft) from /root/reference so that
tests/ and oracle/make_golden.py can execute it as the parity oracle.  Only
usable in the build container: /root/reference does not exist on the GPU box,
which is why its outputs are frozen into tests/golden/ by make_golden.py.

Reference entry points exercised:
  core/network.py:26   CRAFT(args)
  core/network.py:164  CRAFT.forward(image1, image2, iters, flow_init, upsample, test_mode)
"""
import argparse
import contextlib
import io
import os
import sys
import warnings

REF_ROOT = os.environ.get("CRAFT_REFERENCE_ROOT", "/root/reference")
REF_CORE = os.path.join(REF_ROOT, "core")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_CORE, "network.py"))


def craft_args(**overrides) -> argparse.Namespace:
    """Namespace matching evaluate.py's flags for the shipped checkpoints
    (SURVEY.md section 8b / 8d config 2)."""
    d = dict(
        craft=True, use_setrans=True, f2trans="full", f1trans="none",
        corr_radius=4, pos_bias_radius=7, mixed_precision=False, num_heads=1,
        position_only=False, position_and_content=False,
        f2_attn_mask_radius=-1, f2_num_modes=4, f2_pos_code_weight=0.5,
        inter_num_modes=4, inter_qk_have_bias=True, inter_pos_code_type="bias",
        inter_pos_code_weight=0.5, intra_num_modes=4, intra_pos_code_type="bias",
        intra_pos_code_weight=1.0, dropout=0.0,
    )
    d.update(overrides)
    return argparse.Namespace(**d)


@contextlib.contextmanager
def _ref_on_path():
    """Temporarily put the reference's core/ first on sys.path and hide any
    same-named modules of ours (network, corr, setrans, ...)."""
    names = ["network", "corr", "setrans", "setrans_ablation", "gma", "update",
             "extractor", "utils", "utils.utils", "raft"]
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    sys.path.insert(0, REF_CORE)
    try:
        yield
    finally:
        sys.path.remove(REF_CORE)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)


def load_reference_modules():
    """Returns a dict of the reference's modules (network, corr, setrans, gma, update, utils)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    with _ref_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import network, corr, setrans, gma, update, extractor  # noqa
        import utils.utils as uu
        return dict(network=network, corr=corr, setrans=setrans, gma=gma,
                    update=update, extractor=extractor, utils=uu)


def build_reference_model(args=None, checkpoint="craft-sintel.pth", seed=1234, quiet=True):
    """CRAFT(args) from the reference with optional checkpoint, eval mode, CPU fp32."""
    import torch
    mods = load_reference_modules()
    args = args or craft_args()
    torch.manual_seed(seed)
    sink = io.StringIO()
    with warnings.catch_warnings(), contextlib.redirect_stdout(sink if quiet else sys.stdout):
        warnings.simplefilter("ignore")
        model = mods["network"].CRAFT(args)
    if checkpoint:
        path = os.path.join(REF_ROOT, "checkpoints", checkpoint)
        ck = torch.load(path, map_location="cpu", weights_only=False)
        sd = ck["model"] if "model" in ck else ck
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
        missing, unexpected = model.load_state_dict(sd, strict=False)
        model._load_report = (list(missing), list(unexpected))
    model.eval()
    return model, mods


def synthetic_pair(H, W, seed=1234, B=1):
    """SURVEY.md section 8d synthetic inputs: integer noise + (2,3) roll => true flow (3,2)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    image1 = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
    image2 = torch.roll(image1, shifts=(2, 3), dims=(2, 3))
    return image1, image2


def smooth_pair(H, W, seed=1234, B=1, blur=5):
    """Box-blurred noise variant -- textured but smooth."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((B, 3, H + 16, W + 16), generator=g)
    k = torch.ones(3, 1, blur, blur) / (blur * blur)
    for _ in range(2):
        x = F.conv2d(F.pad(x, (blur // 2,) * 4, mode="reflect"), k, groups=3)
    x = (x - x.amin()) / (x.amax() - x.amin()) * 255.0
    image1 = x[:, :, 8:8 + H, 8:8 + W].contiguous()
    image2 = x[:, :, 6:6 + H, 5:5 + W].contiguous()   # image2(y,x) = image1(y-2, x-3)
    return image1, image2
