"""Per-step device time of the single-launch Schelling step on 100 x 100 (BASELINE config 1), with movers / passes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200.examples import Schelling, SchellingScenario

m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
m.step(); torch.cuda.synchronize()
m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(101)]
rows = []
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(100):
    ev[s].record()
    m.step()
    torch.cuda.synchronize()
    rows.append((m.last_movers, m.last_iterations))
ev[100].record(); torch.cuda.synchronize()
m2 = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
e = [torch.cuda.Event(enable_timing=True) for _ in range(101)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(100):
    e[s].record(); m2.step()
e[100].record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
per = [e[s].elapsed_time(e[s + 1]) * 1e3 for s in range(100)]
print(json.dumps({"issue_us_per_step": t_issue / 100 * 1e6, "wall_us_per_step": t_all / 100 * 1e6,
                  "device_us_first10": [round(x, 1) for x in per[:10]], "device_us_last10": [round(x, 1) for x in per[-10:]],
                  "device_us_mean": sum(per) / 100, "movers_passes_first10": rows[:10], "movers_passes_last5": rows[-5:]}))
