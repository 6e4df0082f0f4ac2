import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from craft_b200.network import CRAFT
from craft_b200.ops import TokenGrid
from oracle.ref_loader import craft_args, synthetic_pair

torch.manual_seed(1234)
m = CRAFT(craft_args(attn_clip=0.2)).cuda().eval()
i1, i2 = synthetic_pair(128, 128)
i1, i2 = i1.cuda(), i2.cuda()
g = TokenGrid(16, 16)
for graph in (False, True):
    m.use_cuda_graph = graph
    with torch.no_grad():
        lo, up = m(i1, i2, iters=4, test_mode=1)
    torch.cuda.synchronize()
    ws = m.workspace_for(8 * g.H, 8 * g.W)
    print("graph", graph, "mean flow", up.mean((0, 2, 3)).tolist(), "stat_max", ws.stat_max.tolist(), "flag", ws.flag.tolist(),
          "clips", ws.clip_corr.item(), ws.clip_f2.item(), ws.clip_att.item(),
          "diag", [(n, getattr(m, n).setrans.max_attn, getattr(m, n).setrans.clamp_count) for n in ("corr_fn", "f2_trans", "att")])
