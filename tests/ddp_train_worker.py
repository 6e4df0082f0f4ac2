"""One DDP training step of craft_b200.CRAFT per rank (launched by tests/test_gpu_training.py under torchrun):
the wiring of train_ddp.py:185-256 -- NCCL process group, DDP(find_unused_parameters=True), AdamW, sequence
loss, clip_grad_norm_ -- on a synthetic batch.  Checks that the averaged gradients are identical on all ranks."""
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from craft_b200.network import CRAFT          # noqa: E402
from craft_b200.testing import craft_args, synthetic_pair   # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(1234)
model = CRAFT(craft_args(dropout_prob=0.0)).cuda()
model.train()
ddp = DDP(model, device_ids=[local], find_unused_parameters=True)
opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=1e-5, eps=1e-8)
i1, i2 = synthetic_pair(128, 128, seed=100 + rank, B=2)      # different data per rank
gt = torch.zeros(2, 2, 128, 128, device="cuda")
gt[:, 0], gt[:, 1] = 3.0, 2.0
before = [p.detach().clone() for p in model.parameters()]
preds = ddp(i1.cuda(), i2.cuda(), iters=3, test_mode=0)
loss = sum(0.8 ** (len(preds) - i - 1) * (p - gt).abs().mean() for i, p in enumerate(preds))
loss.backward()
torch.nn.utils.clip_grad_norm_(ddp.parameters(), 1.0)
# after DDP's all-reduce every rank holds the same gradient
flat = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
ref = flat.clone()
dist.broadcast(ref, src=0)
assert torch.equal(flat, ref), "gradients differ across ranks after the all-reduce"
opt.step()
moved = sum(int((a != b.detach()).any()) for a, b in zip(before, model.parameters()))
assert moved > 100 and torch.isfinite(loss)
dist.barrier()
if rank == 0:
    print("ddp step ok: world %d, loss %.4f, %d tensors updated" % (world, loss.item(), moved))
dist.destroy_process_group()
