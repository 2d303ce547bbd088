"""Multi-GPU parity check (run under torchrun): row-sharded Schelling == sequential oracle, any GPU count.

    gpurun --gpus 2 -- 'python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_schelling.py'
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import oracle
from mesa_b200.sharded import ShardedSchelling
from mesa_b200.sharding import block_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def make_grid(rows, cols, density, seed):
    rng = np.random.default_rng(seed)
    u = rng.random((rows, cols))
    v = rng.random((rows, cols))
    return np.where(u < density, (v < 0.5).astype(np.int8), np.int8(-1)).astype(np.int8)


def run_case(name, rows, cols, density, h, radius, steps, seed, empties=None):
    grid = make_grid(rows, cols, density, seed)
    if empties is not None:  # nearly full grid: provokes the argwhere fallback
        grid = np.random.default_rng(seed).integers(0, 2, size=(rows, cols)).astype(np.int8)
        idx = np.random.default_rng(seed + 1).choice(rows * cols, size=empties, replace=False)
        grid.reshape(-1)[idx] = -1
    lo, hi = block_bounds(rows, world, rank)
    m = ShardedSchelling(rows, cols, grid[lo:hi], h, radius, seed=seed)
    ok, fallbacks = True, 0
    st = oracle.SchellingState.from_type_grid(grid) if rank == 0 else None
    if rank == 0:
        ok &= st.assign_state(h, radius) == m.happy_count
    t0 = time.perf_counter()
    for s in range(steps):
        happy = m.step()
        types = m.gather_types().cpu().numpy()
        ids = m.gather_layer(m._id_buf).cpu().numpy()
        hl = m.gather_layer(m._happy_buf).cpu().numpy()
        if rank == 0:
            movers, nfb = st.move_unhappy(seed, s)
            fallbacks += nfb
            want_happy = st.assign_state(h, radius)
            ok &= want_happy == happy and movers == m.last_movers
            ok &= np.array_equal(types, st.type_grid())
            occ = st.type_grid() >= 0
            ok &= np.array_equal(ids.view(np.uint32)[occ].astype(np.int64), st.id_grid()[occ])
            ok &= np.array_equal(hl, st.happy_grid())
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"case": name, "world": world, "ok": bool(ok), "agents": st.n, "steps": steps,
                          "oracle_fallbacks": fallbacks, "last_passes": m.last_passes, "sec": round(dt, 3)}), flush=True)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.broadcast(flag, 0)
    return bool(flag.item())


results = [
    run_case("64x48 r1 h0.4", 64, 48, 0.8, 0.4, 1, 12, 3),
    run_case("33x16 r2 h0.6 uneven", 33, 16, 0.85, 0.6, 2, 8, 4),
    run_case("24x32 dense fallback", 24, 32, 0.0, 0.9, 1, 5, 5, empties=20),
    run_case("512x512 r1", 512, 512, 0.8, 0.4, 1, 4, 6),
]
if rank == 0:
    print("ALL OK" if all(results) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if all(results) else 1)
