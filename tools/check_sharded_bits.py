"""Multi-GPU check (torchrun): row-sharded bit-packed Game of Life == the single-GPU result, any GPU count; + timing.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_bits.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mesa_b200 import ops
from mesa_b200.sharding import PeerBitLife

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

ok_all = True
for rows, cols, torus, T, gens, ghost in [(96, 256, True, 4, 30, 8), (64, 1024, False, 3, 17, 6), (2048, 4096, True, 8, 100, 64)]:
    g = torch.Generator(device="cpu").manual_seed(rows + cols)
    full = (torch.rand((rows * world, cols), generator=g) < 0.3).to(torch.uint8)   # same on every rank
    life = PeerBitLife(rows, cols, torus, dev, generations_per_launch=T, ghost=ghost)
    life.load(full[rank * rows:(rank + 1) * rows].to(dev))
    life.run(gens)
    mine = life.store()
    # reference: the whole stacked grid on this GPU with the uint8 kernel
    a = full.to(dev)
    b = torch.empty_like(a)
    for _ in range(gens):
        ops.gol_step(a, out=b, torus=torus)
        a, b = b, a
    ok = torch.equal(mine, a[rank * rows:(rank + 1) * rows])
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_all &= bool(flag.item())
    if rank == 0:
        print(json.dumps({"case": f"bits {rows}x{cols} per rank torus={torus} T={T} ghost={ghost} gens={gens}", "world": world, "ok": bool(flag.item())}), flush=True)

# timing at the benchmark size: 16384^2 per GPU, weak scaling
G = 16384
gen = torch.Generator(device=dev).manual_seed(rank)
blk = (torch.rand((G, G), device=dev, generator=gen) < 0.2).to(torch.uint8)
GHOSTS = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [8, 64, 256]
for T, ghost in [(8, gh) for gh in GHOSTS]:
    life = PeerBitLife(G, G, True, dev, generations_per_launch=T, ghost=ghost)
    life.load(blk)
    life.run(4 * T)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    life.run(1024)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        us = float(t.item()) * 1e3 / 1024
        print(json.dumps({"T": T, "ghost": ghost, "exchanges": life.exchanges, "world": world, "us_per_generation": us, "updates_per_s": G * G * world / us * 1e6}), flush=True)
    del life
if rank == 0:
    print("ALL OK" if ok_all else "MISMATCH", flush=True)
dist.destroy_process_group()
