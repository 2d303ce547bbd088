"""Latency of the small-state models: Schelling 100x100 (fused single launch vs multi-launch) and Boltzmann 10^6."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200.examples import BoltzmannScenario, BoltzmannWealth, Schelling, SchellingScenario


def timed(fn, steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e6


out = {}
for fused in (True, False):
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False, fused=fused)
    for _ in range(5):
        m.step()
    out[f"schelling_c1_us_fused={fused}"] = round(timed(m.step, 95), 2)
    out[f"happy_after_100_fused={fused}"] = m.happy
    if fused:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mm = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
        for _ in range(5):
            mm.step()
        torch.cuda.synchronize(); e0.record()
        for _ in range(95):
            mm.step()
        e1.record(); torch.cuda.synchronize()
        out["schelling_c1_device_us"] = round(e0.elapsed_time(e1) * 1e3 / 95, 2)
        out["passes_last"] = m.last_iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m2 = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
        m2.step(); torch.cuda.synchronize()
        e0.record(); m2.step(); e1.record(); torch.cuda.synchronize()
        out["schelling_c1_second_step_kernel_us"] = round(e0.elapsed_time(e1) * 1e3, 2)
m = BoltzmannWealth(BoltzmannScenario(n=1_000_000, width=4096, height=4096, rng=1))
m.run_for(3)
out["boltzmann_1e6_us"] = round(timed(m.step, 50), 2)
out["boltzmann_rounds"] = m.last_iterations
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    m.step()
e1.record(); torch.cuda.synchronize()
out["boltzmann_1e6_device_us"] = round(e0.elapsed_time(e1) * 1e3 / 20, 2)
print(json.dumps(out), flush=True)
