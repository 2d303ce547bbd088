"""A few launches of the bit-packed GoL kernel for ncu (python tools/prof_gol_bits.py [T] [rps] [launches])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesa_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 4
G = 16384
g = torch.Generator(device="cuda").manual_seed(0)
a = (torch.rand((G, G), device="cuda", generator=g) < 0.2).to(torch.uint8)
cur = ops.gol_pack(a)
nxt = torch.empty_like(cur)
for _ in range(launches):
    ops.gol_step_bits(cur, nxt, T, torus=True, rows_per_segment=rps)
    cur, nxt = nxt, cur
torch.cuda.synchronize()
print("alive", int(ops.gol_unpack(cur).sum()))
