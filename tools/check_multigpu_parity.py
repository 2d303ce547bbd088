"""Multi-GPU parity under torchrun: bench.py's parity_checks (sharded GoL / Schelling / Boids, any GPU count) plus the
barrier-free ghost refresh against the barrier version.  Used by tests/test_multigpu_gpu.py.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_multigpu_parity.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = bench.parity_checks(dev, world, rank)

# the flag hand-shake refresh == the two-barrier refresh == single GPU, over several exchanges, clamped and torus,
# with a CUDA graph of the launch sequence replayed twice
from mesa_b200 import ops
from mesa_b200.sharding import PeerBitLife

ok = True
for torus, T, ghost, gens in [(True, 8, 8, 40), (False, 4, 16, 50), (True, 8, 64, 20)]:
    rows, cols = 512, 2048
    g = torch.Generator(device="cpu").manual_seed(77 + T)
    full = (torch.rand((rows * world, cols), generator=g) < 0.3).to(torch.uint8).to(dev)
    res = []
    for transport in ("peer", "barrier"):
        life = PeerBitLife(rows, cols, torus, dev, generations_per_launch=T, ghost=ghost, transport=transport, overlap=(transport == "peer"))
        life.load(full[rank * rows:(rank + 1) * rows])
        life.run(gens)
        life.check()
        res.append(life.store())
    # graph of an even launch plan, replayed twice
    life = PeerBitLife(rows, cols, torus, dev, generations_per_launch=T, ghost=ghost)
    life.load(full[rank * rows:(rank + 1) * rows])
    life.run(gens, even_launches=True)          # loads the kernels
    life.load(full[rank * rows:(rank + 1) * rows])
    torch.cuda.synchronize(); dist.barrier()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        life.run(gens, even_launches=True)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph.replay(); graph.replay()
    torch.cuda.synchronize()
    life.check()
    twice = life.store()
    a, b = full, torch.empty_like(full)
    for k in range(2 * gens):
        ops.gol_step(a, out=b, torus=torus)
        a, b = b, a
        if k + 1 == gens:
            want1 = a[rank * rows:(rank + 1) * rows].clone()
    want2 = a[rank * rows:(rank + 1) * rows]
    case_ok = torch.equal(res[0], want1) and torch.equal(res[1], want1) and torch.equal(twice, want2)
    t = torch.tensor([int(case_ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = ok and bool(t.item())
    if rank == 0:
        print(json.dumps({"case": f"ghost refresh torus={torus} T={T} ghost={ghost} gens={gens}", "world": world, "ok": bool(t.item())}), flush=True)
out["ghost_refresh_ok"] = ok
if rank == 0:
    print(json.dumps(out), flush=True)
    print("ALL OK" if (out["parity_ok"] and ok) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if (out["parity_ok"] and ok) else 1)
