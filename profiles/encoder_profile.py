"""Which kernels make up fnet+cnet (outside the hot path)? torch.profiler on one encoder call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from bench import _pairs, _state_dict
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model, _ = _state_dict()
model = model.to(dev).eval()
a, b = _pairs(1, dev)[0]
with torch.no_grad():
    for _ in range(3):
        model._encoders(a, b)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model._encoders(a, b)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
