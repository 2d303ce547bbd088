"""Developer timing of the Schelling happiness kernel on 16384^2 (CUDA events, 50 launches): radius 1 (h = 0.4) and
radius 2 (h = 1.0).  MB_HAPPY_BITS=0 selects the byte-packed kernel, MB_HAPPY_BITS_ROWS=n the rows per warp of the
bit-plane kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mesa_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.rand((n, n), device="cuda", generator=g)
t = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
del u
happy = torch.empty((n, n), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
out = {"bits": os.environ.get("MB_HAPPY_BITS", "1"), "rows": os.environ.get("MB_HAPPY_BITS_ROWS", "")}
for radius, h in [(1, 0.4), (2, 1.0), (2, 0.55)]:
    tbl = ops.happy_threshold_table(h, radius)
    for _ in range(5):
        ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    out[f"r{radius}_h{h}_us"] = round(us, 1)
    out[f"r{radius}_h{h}_count"] = int(cnt.item())
print(json.dumps(out))
