"""Developer micro-benchmark: GoL kernel at full size, CUDA-event timed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
g = torch.Generator(device="cuda").manual_seed(0)
a = (torch.rand((n, n), device="cuda", generator=g) < 0.2).to(torch.uint8)
b = torch.empty_like(a)
for _ in range(10):
    ops.gol_step(a, out=b); a, b = b, a
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ops.gol_step(a, out=b); a, b = b, a
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"n": n, "ms_per_step": ms, "cells_per_s": n * n / ms * 1e3, "GBps": 2 * n * n / ms / 1e6,
                  "alive": int(a.sum())}))
