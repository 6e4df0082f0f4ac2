import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/profiles")
import torch
from craft_b200 import ops
from craft_b200.ops import TokenGrid
import importlib.util
src = open("/root/repo/profiles/gemm_sweep.py").read().split("for st in (2, 3, 4, 6):")[0]
exec(src)
for st in (2, 3):
    run(256, 128, 5, 512, 1, st)
run(256, 128, 5, 512, 4, 3)
run(256, 64, 5, 512, 1, 4)
run(256, 256, 5, 512, 1, 2)
run(256, 128, 1, 512, 1, 3)
run(128, 64, 5, 512, 1, 4)
run(128, 128, 5, 512, 1, 3)
run(576, 64, 1, 256, 1, 4)
run(512, 128, 1, 256, 1, 3)
run(512, 128, 9, 128, 1, 3)
run(192, 64, 9, 256, 1, 4)
