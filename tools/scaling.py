"""Scaling study, one world size per invocation (torchrun or plain python for N=1).  Writes gpurun_out/scaling_N<world>.json.

  GoL      weak: 16384^2 per GPU (fused peer kernel for N>1)           -> bench.py prints that line itself
  C5       weak: Schelling 8192 x 65536 per GPU, d=0.8 h=0.4 r=1, 5 steps (65536^2 at 8 GPUs)
  Boids    strong: 10^7 boids total, vision 15, reference 'large' density, 20 steps
"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mesa_b200.sharded import ShardedBoids, ShardedSchelling

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if "RANK" in os.environ:
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29598", rank=0, world_size=1, device_id=dev)
out = {"world": world}


def sync():
    torch.cuda.synchronize(); dist.barrier()


# ---- C5 weak
rows_per, cols, steps = 8192, 65536, 5
g = torch.Generator(device=dev).manual_seed(100 + rank)
u = torch.rand((rows_per, cols), device=dev, generator=g)
v = torch.rand((rows_per, cols), device=dev, generator=g)
types = torch.where(u < 0.8, (v < 0.5).to(torch.int8), torch.tensor(-1, dtype=torch.int8, device=dev))
del u, v
m = ShardedSchelling(rows_per * world, cols, types, 0.4, 1, seed=7)
del types
sync()
per_step = []
for s in range(steps):
    sync(); t0 = time.perf_counter()
    m.move_unhappy()
    sync(); t1 = time.perf_counter()
    m.assign_state()
    sync(); t2 = time.perf_counter()
    per_step.append({"movers": m.last_movers, "passes": m.last_passes, "relocate_ms": (t1 - t0) * 1e3, "assign_ms": (t2 - t1) * 1e3})
total = sum(r["relocate_ms"] + r["assign_ms"] for r in per_step) / 1e3
out["schelling_c5_weak"] = {"grid": [rows_per * world, cols], "agents": m.n_agents, "steps": steps, "sec": total,
                            "agent_updates_per_s": m.n_agents * steps / total, "per_step": per_step,
                            "assign_state_ms": min(r["assign_ms"] for r in per_step)}
del m
torch.cuda.empty_cache()

# ---- Boids strong
N = 10_000_000
side = 150.0 * math.sqrt(N / 400.0)
n_local = N // world
pos = torch.rand((n_local, 2), device=dev, generator=g)
pos[:, 0] *= side
pos[:, 1] = (pos[:, 1] + rank) * (side / world)
dirs = torch.rand((n_local, 2), device=dev, generator=g) * 2 - 1
b = ShardedBoids(side, side, pos, dirs, torch.arange(n_local, device=dev) + rank * n_local, vision=15.0, separation=2.0)
for _ in range(3):
    b.step()
sync(); t0 = time.perf_counter()
for _ in range(20):
    b.step()
sync(); dt = time.perf_counter() - t0
out["boids_1e7_strong"] = {"ms_per_step": dt / 20 * 1e3, "agent_updates_per_s": N * 20 / dt}
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/scaling_N{world}.json", "w"), indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "per_step"}) for k, v in out.items()}))
dist.destroy_process_group()
