"""Runs the reference's UNMODIFIED evaluate.py on top of the drop-in shim directory (craft_b200/dropin), in
a fresh interpreter (tests/test_dropin.py starts it as a subprocess):

    python tests/dropin_driver.py <reference root> imports            # CPU: only resolve the imports
    python tests/dropin_driver.py <reference root> gen_flow <out.npy>  # GPU: evaluate.gen_flow on the shipped pair

The only things supplied from outside are what a user of the reference supplies too: sys.path (the shim
directory ahead of the reference's core/), stubs for the three optional packages this image lacks (imageio,
fvcore, matplotlib -- evaluate.py imports them at module level) and the command-line Namespace."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref_root, mode = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ref_root, "core"))             # what evaluate.py's sys.path.append("core") means
sys.path.insert(0, os.path.join(ROOT, "craft_b200", "dropin"))   # ... shadowed by the shim
def _missing(top):
    if top in sys.modules:
        return isinstance(sys.modules[top], types.ModuleType) and getattr(sys.modules[top], "__spec__", None) is None
    return importlib.util.find_spec(top) is None


for name in ("imageio", "fvcore", "fvcore.nn", "matplotlib", "matplotlib.pyplot"):
    if _missing(name.split(".")[0]):
        sys.modules[name] = types.ModuleType(name)
if "fvcore.nn" in sys.modules and not hasattr(sys.modules["fvcore.nn"], "FlopCountAnalysis"):
    sys.modules["fvcore.nn"].FlopCountAnalysis = object
    sys.modules["fvcore"].nn = sys.modules["fvcore.nn"]
os.chdir(ref_root)
spec = importlib.util.spec_from_file_location("evaluate", os.path.join(ref_root, "evaluate.py"))
evaluate = importlib.util.module_from_spec(spec)
sys.modules["evaluate"] = evaluate
spec.loader.exec_module(evaluate)

import craft_b200.network as ours          # noqa: E402
import network                              # noqa: E402  (the module name the drivers use)
assert evaluate.CRAFT is ours.CRAFT, "from network import CRAFT did not resolve to craft_b200"
assert network.RAFTER is ours.CRAFT                     # evaluate.py:19 installed the alias on OUR module
import corr, extractor, gma, setrans, update            # noqa: E402,E401
assert corr.TransCorrBlock.__module__ == "craft_b200.corr" and update.GMAUpdateBlock.__module__ == "craft_b200.update"
assert setrans.SETransConfig.__module__ == "craft_b200.setrans" and gma.Aggregate.__module__ == "craft_b200.gma"
# names the hot path does not replace come from the reference's own files (raft.py / craft_nogma.py need them)
assert update.BasicUpdateBlock.__module__.startswith("_craft_ref_")
assert evaluate.RAFT.__module__ == "raft" and evaluate.InputPadder.__module__ == "craft_b200.utils.utils"
assert evaluate.frame_utils.writeFlow.__module__ == "craft_b200.utils.frame_utils" and hasattr(evaluate.frame_utils, "read_gen")
print("imports ok")
if mode == "imports":
    sys.exit(0)

import numpy as np          # noqa: E402
import torch                # noqa: E402
import torch.nn as nn       # noqa: E402
from craft_b200.testing import craft_args   # noqa: E402

out_path = sys.argv[3]
local = os.path.join(ROOT, "tests", "golden", "_local")
args = craft_args()
model = nn.DataParallel(evaluate.CRAFT(args), device_ids=[0])                       # evaluate.py:1534
sd = torch.load(os.path.join(local, "craft-sintel-model.pth"), map_location="cpu")
msg = model.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=False)  # evaluate.py:1540-1557 (prefixed keys)
assert not msg.missing_keys and not msg.unexpected_keys, msg
model.cuda()
model.eval()
grabbed = {}
def _grab(mod, inputs, outputs):          # must return None: a hook's return value replaces the module output
    grabbed.setdefault("flow", outputs[1].detach().cpu())


model.module.register_forward_hook(_grab)
outdir = os.path.join(os.path.dirname(out_path), "gen_flow_out")
evaluate.gen_flow(model, "craft", 12, os.path.join(local, "frame_0047.png"), os.path.join(local, "frame_0048.png"),
                  output_path=outdir, test_mode=1)
pngs = [f for f in os.listdir(outdir) if f.endswith(".png")]
assert pngs, "gen_flow wrote no flow image"
np.save(out_path, grabbed["flow"][0].numpy())
print("gen_flow ok", pngs)
