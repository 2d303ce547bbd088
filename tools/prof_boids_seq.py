"""Two sequential Boids steps at 10^7 boids -- the target of `ncu -k regex:boids_seq_chain` and of a launch list."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
side = 150.0 * math.sqrt(N / 400.0)
g = torch.Generator(device="cuda").manual_seed(3)
pos = torch.rand((N, 2), device="cuda", generator=g) * side
d = torch.rand((N, 2), device="cuda", generator=g) * 2 - 1
reach = ops.boids_reach(d, 1.0)
ws = ops.BoidsSeqWorkspace(N, (side, side), 15.0, reach, "cuda")
p2, d2 = torch.empty_like(pos), torch.empty_like(d)
for e in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ops.boids_step_sequential(pos, d, (side, side), seed=11, epoch=e, vision=15.0, separation=2.0, reach=reach, ws=ws, pos_out=p2, dir_out=d2)
    torch.cuda.synchronize()
    print("step", e, (time.perf_counter() - t0) * 1e3, "ms", flush=True)
    pos, p2 = p2, pos
    d, d2 = d2, d
