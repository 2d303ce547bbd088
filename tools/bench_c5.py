"""BASELINE.json config 5: Schelling on a row-sharded grid, 8192 x 65536 cells per GPU (65536^2 on 8 GPUs).

    torchrun --nproc-per-node N tools/bench_c5.py [rows_per_gpu cols steps]
Weak scaling: every rank owns rows_per_gpu x cols cells.  Reports per-step times (device-side work incl. the
host protocol, max over ranks) and agent-updates/s = agents / step time."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mesa_b200.sharded import ShardedSchelling

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if not dist.is_initialized():
    if "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29599", rank=0, world_size=1, device_id=dev)
rows_per = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
h, radius = 0.4, 1
g = torch.Generator(device=dev).manual_seed(100 + rank)
u = torch.rand((rows_per, cols), device=dev, generator=g)
v = torch.rand((rows_per, cols), device=dev, generator=g)
types = torch.where(u < 0.8, (v < 0.5).to(torch.int8), torch.tensor(-1, dtype=torch.int8, device=dev))
del u, v
t0 = time.perf_counter()
m = ShardedSchelling(rows_per * world, cols, types, h, radius, seed=7)
del types
torch.cuda.synchronize(); dist.barrier()
init_s = time.perf_counter() - t0
rows = []
for s in range(steps):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    m.move_unhappy()
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    happy = m.assign_state()
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    rows.append({"step": s, "movers": m.last_movers, "passes": m.last_passes, "relocate_ms": round((t1 - t0) * 1e3, 3),
                 "assign_ms": round((t2 - t1) * 1e3, 3), "happy": happy})
    if rank == 0:
        print(json.dumps(rows[-1]), flush=True)
        if getattr(m, "profile", None):
            print(json.dumps({"phases_ms": {k: [v[0], round(v[1], 2)] for k, v in m.profile.items()},
                              "events": m.last_events}), flush=True)
            m.profile.clear()
if rank == 0:
    total = sum(r["relocate_ms"] + r["assign_ms"] for r in rows) / 1e3
    print(json.dumps({"workload": f"schelling {rows_per * world}x{cols} d=0.8 h=0.4 r=1, row-sharded x{world}",
                      "agents": m.n_agents, "steps": steps, "init_s": round(init_s, 2), "sec": round(total, 4),
                      "agent_updates_per_s": m.n_agents * steps / total,
                      "mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
dist.destroy_process_group()
