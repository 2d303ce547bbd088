"""Timing of the Schelling relocation at scale on one GPU: first step from a random state + the following ones.
    python tools/dev_reloc.py [rows] [cols] [steps]          (MB_RELOC_DEBUG=1, MB_RELOC_BATCH=1|2|4, MB_RELOC_GENERIC=1)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
C = int(sys.argv[2]) if len(sys.argv) > 2 else R
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
u = torch.rand((R, C), device=dev, generator=g)
t8 = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
del u
occ = (t8 >= 0).reshape(-1)
n = int(occ.sum().item())
aid = torch.where(occ, torch.cumsum(occ.to(torch.int32), 0, dtype=torch.int32) - 1, 0).reshape(R, C).contiguous()
del occ
h, c = ops.schelling_happy(t8, 0.4, 1)
ws = ops.RelocationWorkspace(t8, n)
per = []
for epoch in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mv, its = ops.schelling_relocate(t8, h, aid, ws, 7, epoch)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ops.schelling_happy(t8, 0.4, 1, out=h, count=c)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    per.append({"movers": mv, "iterations": its, "relocate_ms": round((t1 - t0) * 1e3, 3), "assign_ms": round((t2 - t1) * 1e3, 3)})
digest = int((t8.to(torch.int64) * torch.arange(t8.numel(), device=dev).reshape(R, C) % 1000003).sum().item())
print(json.dumps({"grid": [R, C], "agents": n, "batch": os.environ.get("MB_RELOC_BATCH", "4"), "generic": os.environ.get("MB_RELOC_GENERIC"),
                  "digest": digest, "happy": int(c.item()), "per_step": per}), flush=True)
