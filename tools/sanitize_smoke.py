"""Small invocation of every kernel family, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mesa_b200 import ops

rng = np.random.default_rng(0)
dev = "cuda"
# GoL: vector + scalar kernels, torus and clamped
for shape in [(80, 1040), (37, 23)]:
    a = torch.from_numpy((rng.random(shape) < 0.3).astype(np.uint8)).to(dev)
    for torus in (True, False):
        ops.gol_step(a, torus=torus)
# Schelling happiness r=1 / r=2 / scalar, relocation with argwhere fallback
for shape, r, h in [((72, 528), 1, 0.4), ((72, 528), 2, 0.55), ((37, 23), 3, 0.3)]:
    u = rng.random(shape)
    t = torch.from_numpy(np.where(u < 0.2, -1, np.where(u < 0.6, 0, 1)).astype(np.int8)).to(dev)
    ops.schelling_happy(t, h, r)
grid = rng.integers(0, 2, size=(24, 32)).astype(np.int8)
grid.reshape(-1)[rng.choice(24 * 32, 20, replace=False)] = -1
t = torch.from_numpy(grid).to(dev)
n = int((t >= 0).sum())
aid = torch.zeros(t.numel(), dtype=torch.int32, device=dev)
aid[(t >= 0).reshape(-1)] = torch.arange(n, dtype=torch.int32, device=dev)
happy, cnt = ops.schelling_happy(t, 0.9, 1)
ws = ops.RelocationWorkspace(t, n)
cell = torch.zeros(n, dtype=torch.int32, device=dev)
for e in range(3):
    ops.schelling_relocate(t, happy, aid.reshape(24, 32), ws, 5, e, agent_cell=cell)
    ops.schelling_happy(t, 0.9, 1, out=happy, count=cnt)
# Boltzmann
c = torch.from_numpy(rng.integers(0, 30 * 20, 3000).astype(np.int32)).to(dev)
w = torch.ones(3000, dtype=torch.int32, device=dev)
bws = ops.BoltzmannWorkspace(c, 30, 20)
for e in range(3):
    ops.boltzmann_step(c, w, bws, 3, e)
# continuous space: tiled + generic Boids, radius query
for n, side in [(3000, 300.0), (200, 20.0)]:
    pos = torch.from_numpy((rng.random((n, 2)) * side).astype(np.float32)).to(dev)
    d = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).to(dev)
    ops.boids_step(pos, d, (side, side), vision=10.0, separation=2.0)
    ops.radius_query(pos, pos[:50].contiguous(), (side, side), 10.0, True)
# layers + sort
x = torch.from_numpy(rng.random((33, 77)).astype(np.float32)).to(dev)
ops.layer_diffuse(x, 0.3, True); ops.layer_reduce(x, "sum"); ops.layer_binary(x, x, "min"); ops.layer_affine_clamp(x, 2.0, 1.0, 0.0, 2.0)
ops.neighborhood_sum((x > 0.5).to(torch.uint8), 2, False, False, True)
ops.sort_pairs(torch.from_numpy(rng.integers(0, 2**31 - 1, 10000).astype(np.int32)).to(dev), torch.arange(10000, dtype=torch.int32, device=dev))
torch.cuda.synchronize()
print("sanitize smoke done")
