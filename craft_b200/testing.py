"""Synthetic inputs and the argparse.Namespace the reference drivers build -- shared by bench.py, the tests and
the oracle's loaders.  Plain torch, no kernels, no oracle code."""
import argparse


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
