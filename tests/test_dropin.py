"""Drop-in check: the reference's drivers import `from network import CRAFT` after
`sys.path.append('core')`; with craft_b200/dropin ahead on sys.path the UNMODIFIED evaluate.py must import,
build, load the checkpoint and run its single-pair entry point gen_flow (evaluate.py:1251-1384)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "dropin_driver.py")


def _ref_root():
    from oracle import ref_loader as RL
    if not RL.reference_available() or not os.path.isfile(os.path.join(RL.REF_ROOT, "evaluate.py")):
        pytest.skip("reference tree (evaluate.py) neither mounted nor staged under oracle/_ref")
    return RL.REF_ROOT


def _run(*argv):
    env = dict(os.environ, PYTHONPATH="")
    return subprocess.run([sys.executable, DRIVER, *argv], capture_output=True, text=True, env=env, timeout=900)


def test_reference_driver_imports_resolve_to_craft_b200():
    r = _run(_ref_root(), "imports")
    assert r.returncode == 0 and "imports ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_unmodified_evaluate_gen_flow_runs_on_the_kernels(tmp_path):
    import numpy as np
    import torch
    local = os.path.join(ROOT, "tests", "golden", "_local")
    if not os.path.isfile(os.path.join(local, "frame_0047.png")):
        pytest.skip("frame pair / weights not present")
    out = str(tmp_path / "flow.npy")
    r = _run(_ref_root(), "gen_flow", out)
    assert r.returncode == 0 and "gen_flow ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    flow = torch.from_numpy(np.load(out))                       # padded [2,440,1024], as the model returned it
    rec = torch.load(os.path.join(ROOT, "tests", "golden", "sintel_frames_440x1024.pt"), map_location="cpu")
    err = (flow[:, ::4, ::4] - rec["flow_up_s4"]).pow(2).sum(0).sqrt()
    # same bounds as tests/test_gpu_e2e.py for this (ill-conditioned) real pair
    assert err.median().item() <= 1e-2 and err.mean().item() <= rec["ref_bf16_autocast_epe_mean"]
