"""Where does a training step go?  torch.profiler over 2 steps of bench.py's --config train step (1 GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from craft_b200.network import CRAFT
from craft_b200.testing import craft_args, synthetic_pair

dp = float(sys.argv[1]) if len(sys.argv) > 1 else None
kw = dict(mixed_precision=True)
if dp is not None:
    kw["dropout_prob"] = dp
torch.manual_seed(1234)
m = CRAFT(craft_args(**kw)).cuda()
m.train(); m.freeze_bn()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
scaler = torch.amp.GradScaler("cuda")
a, b = synthetic_pair(400, 720, B=2)
a, b = a.cuda(), b.cuda()
gt = torch.zeros(2, 2, 400, 720, device="cuda")


def step():
    preds = m(a, b, iters=12, test_mode=0)
    loss = sum(0.8 ** (12 - k - 1) * (p - gt).abs().mean() for k, p in enumerate(preds))
    opt.zero_grad(set_to_none=True)
    scaler.scale(loss).backward()
    scaler.step(opt); scaler.update()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
