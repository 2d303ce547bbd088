"""Developer micro-benchmark: Schelling happiness kernel + relocation at scale."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
radius = int(sys.argv[2]) if len(sys.argv) > 2 else 1
h = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.rand((n, n), device="cuda", generator=g)
t = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
del u
occ = (t >= 0).reshape(-1)
nag = int(occ.sum())
aid = torch.zeros(n * n, dtype=torch.int32, device="cuda")
aid[occ] = torch.arange(nag, dtype=torch.int32, device="cuda")
aid = aid.reshape(n, n)
happy = torch.empty((n, n), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
tbl = ops.happy_threshold_table(h, radius)
for _ in range(3):
    ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(json.dumps({"kernel": "happy", "n": n, "radius": radius, "ms": ms, "GBps": 2 * n * n / ms / 1e6, "happy": int(cnt.item()), "agents": nag}))
ws = ops.RelocationWorkspace(t, nag)
for s in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    movers, iters = ops.schelling_relocate(t, happy, aid, ws, 1, s)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ops.schelling_happy(t, h, radius, False, out=happy, count=cnt, table=tbl)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"step": s, "movers": movers, "iters": iters, "relocate_ms": (t1 - t0) * 1e3, "happy_ms": (t2 - t1) * 1e3, "happy": int(cnt.item())}))
