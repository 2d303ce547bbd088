"""A few Wolf-Sheep steps at the bench size (10^6 sheep, 2 x 10^5 wolves on 2048^2) -- the target of an ncu launch list."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200.examples.wolf_sheep import WolfSheep, WolfSheepScenario
m = WolfSheep(WolfSheepScenario(width=2048, height=2048, initial_sheep=1_000_000, initial_wolves=200_000, rng=3))
m.step(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    m.step()
torch.cuda.synchronize()
print("ms per step", (time.perf_counter() - t0) / 4 * 1e3, "passes", m.last_sheep_passes, m.last_wolf_passes)
