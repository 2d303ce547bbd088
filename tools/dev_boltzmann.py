"""Developer micro-benchmark: Boltzmann wealth at BASELINE.json config 3 (10^6 agents on 4096^2)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
g = torch.Generator(device="cuda").manual_seed(0)
cell = torch.randint(0, side * side, (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.int32)
wealth = torch.ones(n, dtype=torch.int32, device="cuda")
ws = ops.BoltzmannWorkspace(cell, side, side)
for e in range(3):
    ops.boltzmann_step(cell, wealth, ws, 1, e)
torch.cuda.synchronize(); t0 = time.perf_counter()
iters = 0
for e in range(3, 3 + steps):
    ops.boltzmann_step(cell, wealth, ws, 1, e)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
ws.check(); iters = int(ws.iterations.item()) * steps
print(json.dumps({"n": n, "side": side, "ms_per_step": dt / steps * 1e3, "agent_updates_per_s": n * steps / dt,
                  "mean_iters": iters / steps, "max_wealth": int(wealth.max()), "sum": int(wealth.sum())}))
