"""Generates tests/golden/*.pt by EXECUTING the unmodified reference (/root/reference) on CPU fp32.
Run in the build container only:   python tests/golden/make_golden.py

Two weight sources:
  * "seeded": craft_b200.network.CRAFT(args) constructed under torch.manual_seed(1234) on CPU; its
    state dict is loaded into the reference model (identical keys), so the GPU box can rebuild the
    very same weights without any file.
  * "sintel": the reference's checkpoints/craft-sintel.pth.  Its 'model' dict is also copied to
    tests/golden/_local/ (git-ignored) so it can travel to the GPU box with gpurun; tests that
    need it skip when it is absent.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import (REF_ROOT, build_reference_model, craft_args, smooth_pair,  # noqa: E402
                               synthetic_pair)

OUT = os.path.join(ROOT, "tests", "golden")
LOCAL = os.path.join(OUT, "_local")


def seeded_state(args_kw, tweak=None):
    """tweak: {state-dict key: scalar} -- parameters overwritten with a constant after the seeded init (e.g. GMA's
    gamma, which initialises to 0 and would make the aggregation invisible)."""
    from craft_b200.network import CRAFT
    torch.manual_seed(1234)
    m = CRAFT(craft_args(**args_kw))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for k, v in (tweak or {}).items():
        sd[k].fill_(v)
    return sd


def capture(model, image1, image2, iters, flow_init=None, seams=False):
    """Runs the reference; returns outputs (+ per-seam tensors from hooks when seams=True)."""
    rec = {}
    hooks = []
    if seams:
        def save(name, idx=None):
            def fn(mod, inp, out):
                if name in rec:
                    return
                rec[name] = (out[idx] if idx is not None else out)
            return fn
        def fnet_hook(m, i, o):
            rec.setdefault("fnet_out", torch.cat(list(o), 0))
        hooks.append(model.fnet.register_forward_hook(fnet_hook))
        hooks.append(model.cnet.register_forward_hook(save("cnet_out")))
        if hasattr(model, "f2_trans"):
            hooks.append(model.f2_trans.register_forward_hook(save("f2_out")))
        ub = model.update_block
        hooks.append(ub.encoder.register_forward_hook(save("motion_it0")))
        hooks.append(ub.aggregator.register_forward_hook(save("aggr_it0")))
        hooks.append(ub.gru.register_forward_hook(save("net_it0")))
        def ub_hook(m, i, o):
            rec.setdefault("ub_it0", dict(corr=i[2].clone(), flow=i[3].clone(), mask=o[1].clone(), delta=o[2].clone()))
        hooks.append(ub.register_forward_hook(ub_hook))
    with torch.no_grad():
        flow_lo, ups = model(image1, image2, iters=iters, flow_init=flow_init, test_mode=2)
    for h in hooks:
        h.remove()
    out = dict(flow_lo=flow_lo[0].clone(), flow_up=ups[-1][0].clone(),
               flow_up_first=ups[0][0].clone(),
               # every iteration's upsampled flow on a stride-8 lattice: the error TRAJECTORY over the refinement loop
               flow_up_iters_s8=torch.stack([u[0][:, ::8, ::8] for u in ups]).clone())
    for k, v in rec.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out["ub_it0." + kk] = vv[0].clone() if vv.dim() == 4 else vv.clone()
        else:
            out[k] = v.detach().clone()
    return out


def main():
    os.makedirs(LOCAL, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    # local, untracked copy of the trained weights + the shipped frame pair for the GPU box
    ck = torch.load(os.path.join(REF_ROOT, "checkpoints", "craft-sintel.pth"), map_location="cpu", weights_only=False)
    sd = ck["model"] if "model" in ck else ck
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    torch.save(sd, os.path.join(LOCAL, "craft-sintel-model.pth"))
    import shutil
    for f in ("frame_0047.png", "frame_0048.png"):
        shutil.copy(os.path.join(REF_ROOT, "imgs", f), os.path.join(LOCAL, f))

    cases = [
        # name, weights, args overrides, H, W, iters, input kind, seams
        ("seeded_setrans_128", "seeded", {}, 128, 128, 4, "noise", True),
        ("seeded_gma_128", "seeded", dict(use_setrans=False), 128, 128, 4, "noise", True),
        ("seeded_gma_448x1024", "seeded", dict(use_setrans=False), 448, 1024, 12, "noise", False),
        # attn_clip below the seeded score maxima (f2 0.51, corr 1.62, att 0.22): the data-dependent clamp of
        # core/setrans.py:527-529 fires in all three attentions
        ("seeded_clip02_128", "seeded", dict(attn_clip=0.2), 128, 128, 4, "noise", True),
        # clamp fires in f2_trans and corr_fn but NOT in the intra-frame attention
        ("seeded_clip03_128", "seeded", dict(attn_clip=0.3), 128, 128, 4, "noise", False),
        ("seeded_f2radius_128", "seeded", dict(f2_attn_mask_radius=5), 128, 128, 4, "noise", True),
        ("seeded_modes2_128", "seeded", dict(inter_num_modes=2, intra_num_modes=2), 128, 128, 4, "noise", True),
        ("seeded_plain_128", "seeded", dict(craft=False, use_setrans=False, f2trans="none", corr_multiplier=1),
         128, 128, 4, "noise", False),
        ("sintel_128", "sintel", {}, 128, 128, 4, "noise", True),
        ("sintel_smooth_256x320", "sintel", {}, 256, 320, 12, "smooth", False),
        ("sintel_flowinit_192x256", "sintel", {}, 192, 256, 6, "smooth_init", False),
        ("sintel_448x1024", "sintel", {}, 448, 1024, 12, "noise", False),
        ("sintel_kitti_384x1248", "sintel", {}, 384, 1248, 24, "smooth", False),
        ("sintel_frames_440x1024", "sintel", {}, 440, 1024, 12, "frames", True),
    ]
    tweaks = {
        "seeded_gma_128": {"update_block.aggregator.gamma": 0.5},
        "seeded_gma_448x1024": {"update_block.aggregator.gamma": 0.5},
        # negative soft-aggregation weight: with M == 2 a -inf sentinel for the unused modes would turn into NaN
        "seeded_modes2_128": {"corr_fn.setrans.attn_softaggr.feat2score.weight": -0.8},
    }
    only = sys.argv[1:]
    for name, wsrc, kw, H, W, iters, kind, seams in cases:
        if only and name not in only:
            continue
        args = craft_args(**kw)
        model, _ = build_reference_model(args, checkpoint=("craft-sintel.pth" if wsrc == "sintel" else None))
        if wsrc == "seeded":
            model.load_state_dict(seeded_state(kw, tweaks.get(name)), strict=True)
        flow_init = None
        if kind == "noise":
            i1, i2 = synthetic_pair(H, W)
        elif kind in ("smooth", "smooth_init"):
            i1, i2 = smooth_pair(H, W)
            if kind == "smooth_init":
                g = torch.Generator().manual_seed(7)
                flow_init = torch.randn((1, 2, H // 8, W // 8), generator=g) * 0.5 + torch.tensor([0.3, 0.2]).view(1, 2, 1, 1)
        else:
            import numpy as np
            from PIL import Image
            from craft_b200.utils.utils import InputPadder
            a = torch.from_numpy(np.array(Image.open(os.path.join(LOCAL, "frame_0047.png")))).permute(2, 0, 1).float()[None]
            b = torch.from_numpy(np.array(Image.open(os.path.join(LOCAL, "frame_0048.png")))).permute(2, 0, 1).float()[None]
            i1, i2 = InputPadder(a.shape).pad(a, b)
        out = capture(model, i1, i2, iters, flow_init, seams)
        big = H * W > 256 * 320
        rec = dict(name=name, weights=wsrc, args=kw, tweak=tweaks.get(name), H=H, W=W, iters=iters, kind=kind,
                   flow_lo=out["flow_lo"], flow_up_mean=out["flow_up"].mean((1, 2)),
                   flow_up_absmax=out["flow_up"].abs().max())
        if big:
            rec["flow_up_s4"] = out["flow_up"][:, ::4, ::4].contiguous()
        else:
            rec["flow_up"] = out["flow_up"]
            rec["flow_up_first"] = out["flow_up_first"]
        if flow_init is not None:
            rec["flow_init"] = flow_init
        if kind == "frames":
            # yardstick for this ill-conditioned real pair (occlusions, 200-px motions): how far the
            # reference's OWN reduced-precision mode (bf16 autocast) lands from its fp32 result.
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                _, up16 = model(i1, i2, iters=iters, test_mode=1)
            e = (out["flow_up"] - up16[0].float()).pow(2).sum(0).sqrt()
            rec["ref_bf16_autocast_epe_mean"] = float(e.mean())
            rec["ref_bf16_autocast_epe_median"] = float(e.median())
        for k, v in out.items():
            if k not in ("flow_lo", "flow_up", "flow_up_first"):
                if big and k != "flow_up_iters_s8":
                    if k in ("fnet_out", "cnet_out"):
                        continue                      # 20 MB at this size; the seam report runs our own encoders
                    v = v[..., ::4, ::4].contiguous() if v.dim() >= 3 and k != "aggr_it0" else v
                    if k == "aggr_it0":               # [1,U,128] tokens -> stride-4 lattice like the NCHW seams
                        v = v.reshape(1, H // 8, W // 8, -1)[:, ::4, ::4].contiguous()
                    k = k + "_s4"
                rec[k] = v
        # the reference's attention diagnostics (core/setrans.py:524-529) after this single forward
        diag = {}
        for mod_name in ("corr_fn", "f2_trans", "att"):
            mod = getattr(model, mod_name, None)
            st = getattr(mod, "setrans", None)
            if st is not None:
                diag[mod_name] = dict(max_attn=float(st.max_attn), clamp_count=int(st.clamp_count))
        rec["diag"] = diag
        torch.save(rec, os.path.join(OUT, name + ".pt"))
        print(name, "flow_up mean", rec["flow_up_mean"].tolist(), "absmax", float(rec["flow_up_absmax"]),
              "keys", len(rec), flush=True)


GRAD_KEYS = ["fnet.conv1.weight", "fnet.layer3.1.conv2.weight", "cnet.conv1.weight",
             "f2_trans.setrans.query.weight", "f2_trans.setrans.key.weight", "f2_trans.setrans.out_trans.first_linear.weight",
             "f2_trans.setrans.out_trans.input_skip_coeff", "f2_trans.vispos_encoder.pos_coder.biases",
             "corr_fn.setrans.query.weight", "corr_fn.setrans.query.bias", "corr_fn.setrans.attn_softaggr.feat2score.weight",
             "corr_fn.vispos_encoder.pos_coder.biases", "att.setrans.query.weight", "att.setrans.key.weight",
             "att.vispos_encoder.pos_coder.biases", "update_block.aggregator.first_linear.weight",
             "update_block.aggregator.feat_softaggr.feat2score.weight", "update_block.aggregator.input_skip_coeff",
             "update_block.encoder.convc1.weight", "update_block.gru.convz1.weight", "update_block.gru.convq2.weight",
             "update_block.flow_head.conv2.weight", "update_block.mask.2.weight"]


def training_loss(flow_preds, flow_gt, gamma=0.8):
    """train.py:44-73 sequence_loss without the validity mask: exponentially weighted L1 over the iterations."""
    n = len(flow_preds)
    return sum(gamma ** (n - i - 1) * (flow_preds[i] - flow_gt).abs().mean() for i in range(n))


def grad_case():
    """Gradients of the EXECUTED reference in training mode (dropout_prob = 0 so that the run is deterministic,
    BatchNorm frozen as train.py does after the chairs stage): the pin of craft_b200/train_path.py."""
    kw = dict(dropout_prob=0.0)
    H = W = 128
    iters = 3
    args = craft_args(**kw)
    model, _ = build_reference_model(args, checkpoint=None)
    sd = seeded_state(kw)
    # seeded positional-bias tables are all-zero and gamma-like scalars sit at their init: spread the tables so
    # that their gradient paths are exercised
    g = torch.Generator().manual_seed(77)
    for k in sd:
        if k.endswith("pos_coder.biases"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.3
    model.load_state_dict(sd, strict=True)
    model.train()
    model.freeze_bn()
    i1, i2 = synthetic_pair(H, W)
    flow_gt = torch.zeros(1, 2, H, W)
    flow_gt[:, 0], flow_gt[:, 1] = 3.0, 2.0
    preds = model(i1, i2, iters=iters, test_mode=0)
    loss = training_loss(preds, flow_gt)
    loss.backward()
    named = dict(model.named_parameters())
    rec = dict(name="seeded_setrans_128_grad", args=kw, H=H, W=W, iters=iters, loss=float(loss),
               pos_seed=77, flow_last=preds[-1][0].detach().clone(),
               grads={k: named[k].grad.detach().clone() for k in GRAD_KEYS},
               grad_norms={k: float(v.grad.norm()) for k, v in named.items() if v.grad is not None})
    torch.save(rec, os.path.join(OUT, "seeded_setrans_128_grad.pt"))
    print("seeded_setrans_128_grad loss", float(loss), "params with grad", len(rec["grad_norms"]))


if __name__ == "__main__":
    if sys.argv[1:] == ["grad"]:
        grad_case()
    else:
        main()
