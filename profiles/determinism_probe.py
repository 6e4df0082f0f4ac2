import os, sys, torch
sys.path.insert(0, "/root/repo")
from craft_b200.network import CRAFT
from oracle.ref_loader import craft_args, synthetic_pair
torch.manual_seed(1234)
m = CRAFT(craft_args()).cuda().eval()
i1, i2 = (t.cuda() for t in synthetic_pair(128, 128))
with torch.no_grad():
    outs = []
    for k in range(4):
        m.elide_dead_upsample = (k % 2 == 0)
        lo, up = m(i1, i2, iters=4, test_mode=1)
        outs.append((lo.clone(), up.clone()))
    for k in range(1, 4):
        print("run", k, "vs 0: lo maxdiff", (outs[k][0] - outs[0][0]).abs().max().item(), "up maxdiff", (outs[k][1] - outs[0][1]).abs().max().item())
    m.use_cuda_graph = False
    a = m(i1, i2, iters=4, test_mode=1)
    b = m(i1, i2, iters=4, test_mode=1)
    print("eager vs eager", (a[1] - b[1]).abs().max().item(), "eager vs graph", (a[1] - outs[1][1]).abs().max().item())
