"""One launch of the Schelling happiness kernel for radius 1 and 2 on 16384^2 (after warm launches) -- ncu target."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from mesa_b200 import ops  # noqa: E402

n = 16384
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.rand((n, n), device="cuda", generator=g)
t = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
del u
happy = torch.empty((n, n), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for radius, h in [(1, 0.4), (2, 1.0)]:
    tbl = ops.happy_threshold_table(h, radius)
    for _ in range(3):
        ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
    torch.cuda.synchronize()
print("ok")
