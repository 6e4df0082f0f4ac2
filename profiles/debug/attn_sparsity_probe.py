"""How sparse is the intra-frame attention at tile granularity?  For the bench input (and the real frame pair):
fraction of (128-query tile, mode, 8x16 key block) tiles whose largest probability is below 2^-25 (the value
below which a float16 P rounds to zero) / 2^-20 / 2^-15."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from bench import CONFIGS, _build_model, _pairs
from oracle import restate as R

dev = torch.device("cuda", 0)
cfg = CONFIGS["sintel"]
model, _ = _build_model(cfg)
model = model.to(dev).eval()


def probe(tag, i1, i2):
    with torch.no_grad():
        f1, f2, cn = model._encoders(i1.float(), i2.float())
        inp = torch.relu(cn[:, 128:]).float()
        B, C, h, w = inp.shape
        probs, _, gmax = R.self_attention_probs(inp, model.att.setrans.query.weight, model.att.setrans.key.weight, 4,
                                                model.att.vispos_encoder.pos_coder.biases, 1.0)
        P = probs[0]                                   # [4,U,U]
        U = h * w
        nq = (U + 127) // 128
        Pk = P.reshape(4, U, h, w)
        hb, wb = (h + 7) // 8, (w + 15) // 16
        pad = torch.zeros((4, U, hb * 8, wb * 16), device=dev)
        pad[:, :, :h, :w] = Pk
        tile_max = pad.reshape(4, U, hb, 8, wb, 16).amax(dim=(3, 5))           # [4,U,hb,wb]
        qpad = torch.zeros((4, nq * 128, hb, wb), device=dev)
        qpad[:, :U] = tile_max
        tmax = qpad.reshape(4, nq, 128, hb, wb).amax(dim=2)                    # [4,nq,hb,wb]
        out = {"gmax": gmax}
        for e in (15, 20, 25):
            out["frac_tiles_below_2^-%d" % e] = (tmax < 2.0 ** -e).float().mean().item()
        out["mass_in_top_1pct_keys"] = P.flatten(0, 1).topk(U // 100, dim=-1).values.sum(-1).mean().item()
        print(tag, out, flush=True)


a, b = _pairs(cfg, [0], dev)[0]
probe("synthetic-noise", a, b)
local = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "_local")
if os.path.isfile(os.path.join(local, "frame_0047.png")):
    import numpy as np
    from PIL import Image
    from craft_b200.utils.utils import InputPadder
    x = torch.from_numpy(np.array(Image.open(os.path.join(local, "frame_0047.png")))).permute(2, 0, 1).float()[None]
    y = torch.from_numpy(np.array(Image.open(os.path.join(local, "frame_0048.png")))).permute(2, 0, 1).float()[None]
    x, y = InputPadder(x.shape).pad(x, y)
    probe("sintel-frames", x.to(dev), y.to(dev))
