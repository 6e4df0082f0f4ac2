"""Probe of the clamp gate (hotpath.build_correlation / attention_stats) with a low attn_clip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from craft_b200 import hotpath as hp, ops
from craft_b200.ops import TokenGrid
from craft_b200.setrans import get_workspace

dev = torch.device("cuda", 0)
g = TokenGrid(16, 16)
ws = get_workspace(g, dev)
gen = torch.Generator(device=dev).manual_seed(1)
Q = (torch.randn((g.Mp, 256), device=dev, generator=gen) * 0.6).to(torch.bfloat16)
K = (torch.randn((g.Mp, 256), device=dev, generator=gen) * 0.6).to(torch.bfloat16)
table = torch.randn((15, 15), device=dev, generator=gen)
diag = torch.zeros(2, device=dev)
for rep in range(3):
    hp.build_correlation(ws, Q, K, M=4, d=64, w_agg=0.1, table=table, w_pos=0.5, global_norm=True, attn_clip=0.2, diag=diag)
    torch.cuda.synchronize()
    print("corr rep", rep, "stat_max", ws.stat_max.tolist(), "flag", ws.flag.tolist(), "clip", ws.clip_corr.item(), "diag", diag.tolist(),
          "sums", ws.stat_sum.tolist(), "mean_rstd", ws.mean_rstd.tolist())
Qa, Ka = Q[:, :128].contiguous(), K[:, :128].contiguous()
for rep in range(2):
    hp.attention_stats(ws, Qa, Ka, M=4, d=32, table=table, w_pos=1.0, clip=ws.clip_att, lse2=ws.lse2_att, slot=2, attn_clip=0.2, diag=diag)
    torch.cuda.synchronize()
    print("lse rep", rep, "stat_max", ws.stat_max.tolist(), "flag", ws.flag.tolist(), "clip", ws.clip_att.item(), "diag", diag.tolist())
