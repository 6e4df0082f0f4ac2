import torch
x = torch.randn(1<<20, device='cuda'); y = x*2; torch.cuda.synchronize(); print("ok", y.sum().item())
