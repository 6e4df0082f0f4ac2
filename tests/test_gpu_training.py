"""Training path (BASELINE configs[3]): CRAFT.forward under grad against the gradients
of the EXECUTED reference (tests/golden/seeded_setrans_128_grad.pt, made by tests/golden/make_golden.py grad:
training mode, dropout_prob = 0, frozen BatchNorm, sequence loss of train.py:44-73)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

from craft_b200.network import CRAFT                     # noqa: E402
from craft_b200.testing import craft_args, synthetic_pair   # noqa: E402


def _setup(rec, **extra):
    torch.manual_seed(1234)
    m = CRAFT(craft_args(**dict(rec["args"], **extra)))
    g = torch.Generator().manual_seed(rec["pos_seed"])
    sd = m.state_dict()
    for k in sd:
        if k.endswith("pos_coder.biases"):
            sd[k].copy_(torch.randn(sd[k].shape, generator=g) * 0.3)
    m = m.cuda()
    m.train()
    m.freeze_bn()
    return m


def _loss(preds, gamma=0.8):
    gt = torch.zeros_like(preds[0])
    gt[:, 0], gt[:, 1] = 3.0, 2.0
    n = len(preds)
    return sum(gamma ** (n - i - 1) * (preds[i] - gt).abs().mean() for i in range(n))


def _run(m, rec):
    i1, i2 = synthetic_pair(rec["H"], rec["W"])
    preds = m(i1.cuda(), i2.cuda(), iters=rec["iters"], test_mode=0)
    assert isinstance(preds, list) and len(preds) == rec["iters"]
    loss = _loss(preds)
    loss.backward()
    return loss.item(), preds


def test_pytorch_restatement_matches_reference_gradients():
    """train_kernels=False: every block runs the differentiable restatement -> fp32 agreement with the reference."""
    rec = torch.load(os.path.join(GOLD, "seeded_setrans_128_grad.pt"), map_location="cpu")
    m = _setup(rec)
    m.train_kernels = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    loss, preds = _run(m, rec)
    assert abs(loss - rec["loss"]) <= 1e-4 * rec["loss"]
    assert (preds[-1][0].detach().cpu() - rec["flow_last"]).abs().max().item() <= 1e-3
    named = dict(m.named_parameters())
    for k, gref in rec["grads"].items():
        g = named[k].grad.cpu()
        assert (g - gref).norm().item() <= 1e-3 * gref.norm().item() + 1e-9, k
    # exactly the parameters the reference touches receive a gradient (DDP runs with find_unused_parameters=True)
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == set(rec["grad_norms"]), got ^ set(rec["grad_norms"])


def test_kernel_forward_recompute_backward_matches_reference_gradients():
    """Default training path with dropout off: sm_100a kernels in forward (bf16 operands), recompute-in-PyTorch
    backward.  Gradients must point the same way as the reference's and have its size."""
    from craft_b200 import _lib
    rec = torch.load(os.path.join(GOLD, "seeded_setrans_128_grad.pt"), map_location="cpu")
    m = _setup(rec)
    n0 = _lib.launch_count()
    loss, preds = _run(m, rec)
    launches = _lib.launch_count() - n0
    named = dict(m.named_parameters())
    report, bad = {}, []
    for k, gref in rec["grads"].items():
        if named[k].grad is None:
            report[k] = (float("nan"), float("nan"))
            bad.append(k)
            continue
        g = named[k].grad.cpu()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), gref.flatten(), dim=0).item()
        ratio = g.norm().item() / (gref.norm().item() + 1e-20)
        report[k] = (cos, ratio)
        if not (cos >= 0.98 and 0.9 <= ratio <= 1.1):
            bad.append(k)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_grad_report.txt"), "w") as f:
        f.write("loss %.6f (reference %.6f), %d kernel launches\n" % (loss, rec["loss"], launches))
        for k, v in report.items():
            f.write("%-70s cos %.4f  |g|/|g_ref| %.3f\n" % (k, v[0], v[1]))
    assert launches > 30, "the kernel forward did not run"
    assert abs(loss - rec["loss"]) <= 5e-3 * rec["loss"], (loss, rec["loss"])
    assert not bad, {k: report[k] for k in bad}
    got = {k for k, p in named.items() if p.grad is not None}
    assert got == set(rec["grad_norms"]), got ^ set(rec["grad_norms"])


def test_reference_training_configuration_runs_with_dropout_and_batch_of_two():
    """The reference's own training defaults (token dropout 0.1, attention dropout 0.2, core/setrans.py:110-111)
    on a batch of two: dropout is live, so two forwards differ, and every gradient is finite."""
    torch.manual_seed(7)
    m = CRAFT(craft_args()).cuda()
    m.train()
    a1, a2 = synthetic_pair(128, 128, B=2)
    out = []
    for _ in range(2):
        m.zero_grad()
        preds = m(a1.cuda(), a2.cuda(), iters=2, test_mode=0)
        assert preds[0].shape == (2, 2, 128, 128)
        loss = _loss(preds)
        loss.backward()
        out.append(loss.item())
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    assert out[0] != out[1]
    # batch of two through the kernel-forward functions (dropout off)
    m2 = CRAFT(craft_args(dropout_prob=0.0)).cuda()
    m2.train()
    m2.freeze_bn()                     # batch statistics would tie the two samples together (train.py does this too)
    b1, b2 = torch.cat([a1[:1], torch.roll(a1[:1], 5, 3)]), torch.cat([a2[:1], torch.roll(a2[:1], 7, 3)])   # two different pairs
    preds = m2(b1.cuda(), b2.cuda(), iters=2, test_mode=0)
    _loss(preds).backward()
    assert all(torch.isfinite(p.grad).all() for p in m2.parameters() if p.grad is not None)
    # a batch of two equals the two single-pair runs (kernel path, no dropout, one pyramid / handle per sample)
    for b in range(2):
        pb = m2(b1[b:b + 1].cuda(), b2[b:b + 1].cuda(), iters=2, test_mode=0)[-1]
        assert (pb - preds[-1][b:b + 1]).abs().max().item() <= 2e-2, b


def test_ddp_training_step_gradient_allreduce_is_the_only_collective():
    """train_ddp.py:198-200: DistributedDataParallel(find_unused_parameters=True) around CRAFT; one optimiser step
    on every visible GPU (1 on a single-GPU box, 2 under `gpurun --gpus 2`)."""
    n = min(torch.cuda.device_count(), 2)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29653",
                        os.path.join(ROOT, "tests", "ddp_train_worker.py")], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0 and "ddp step ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
