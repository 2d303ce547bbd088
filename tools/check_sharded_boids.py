"""Multi-GPU check (torchrun): domain-decomposed Boids == the synchronous oracle step, any GPU count; + timing."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import oracle
from mesa_b200.sharded import ShardedBoids

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def build(n, side, seed, vision):
    rng = np.random.default_rng(seed)
    pos = (rng.random((n, 2)) * side).astype(np.float32)
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    owner = ShardedBoids.owner_of(torch.from_numpy(pos[:, 1]), side, world).numpy()
    mine = owner == rank
    ids = np.arange(n)
    return pos, dirs, ShardedBoids(side, side, pos[mine], dirs[mine], ids[mine], vision=vision, separation=2.0)


ok_all = True
for n, side, vision, steps in [(4000, 400.0, 10.0, 4), (20000, 600.0, 15.0, 3)]:
    pos, dirs, m = build(n, side, 5, vision)
    ok = True
    for s in range(steps):
        m.step()
        ids, gp, gd = m.gather()
        if rank == 0:
            wp, wd, _ = oracle.boids_step(pos, dirs, (side, side), vision=vision, separation=2.0)
            gp, gd = gp.cpu().numpy(), gd.cpu().numpy()
            ok &= np.array_equal(ids.cpu().numpy(), np.arange(n))
            ok &= float(np.max(np.abs(gp - wp)) / side) <= 1e-5 and float(np.max(np.abs(gd - wd))) <= 1e-5
            pos, dirs = gp.copy(), gd.copy()      # chain from the device state (fp32) like the single-GPU tests
    if rank == 0:
        print(json.dumps({"case": f"boids n={n} side={side}", "world": world, "ok": bool(ok)}), flush=True)
        ok_all &= ok

# timing: BASELINE config 4 density, N total boids
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
side = 150.0 * math.sqrt(N / 400.0)
g = torch.Generator(device="cuda").manual_seed(rank)
n_local = N // world
pos = torch.rand((n_local, 2), device="cuda", generator=g)
pos[:, 0] *= side
pos[:, 1] = (pos[:, 1] + rank) * (side / world)
dirs = torch.rand((n_local, 2), device="cuda", generator=g) * 2 - 1
m = ShardedBoids(side, side, pos, dirs, torch.arange(n_local, device="cuda") + rank * n_local, vision=15.0, separation=2.0)
for _ in range(3):
    m.step()
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
steps = 20
for _ in range(steps):
    m.step()
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"workload": f"boids N={N} side={side:.0f} vision=15 x{world} GPUs", "ms_per_step": dt / steps * 1e3,
                      "agent_updates_per_s": N * steps / dt}), flush=True)
    print("ALL OK" if ok_all else "FAILED", flush=True)
dist.destroy_process_group()
