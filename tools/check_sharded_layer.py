"""Multi-GPU check (torchrun): row-sharded property layers == the single-GPU result, any GPU count; + timing.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded_layer.py
ShardedLayer (halo rows by NCCL send/recv) for neighbourhood sums / diffusion / aggregates, and the same diffusion
kernel with its halo rows read in place from the neighbours' HBM (PeerHaloLayer, symmetric memory)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mesa_b200 import ops
from mesa_b200.sharding import PeerHaloLayer, ShardedLayer, block_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def agree(ok: bool) -> bool:
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


ok_all = True
for rows, cols, torus in [(101, 96, True), (101, 96, False), (515, 1028, True)]:
    g = torch.Generator(device="cpu").manual_seed(rows + cols)
    counts = (torch.rand((rows, cols), generator=g) * 256).to(torch.uint8).to(dev)        # same on every rank
    heat = torch.rand((rows, cols), generator=g).to(dev)
    lo, hi = block_bounds(rows, world, rank)
    layer = ShardedLayer(counts[lo:hi].clone(), torus=torus)
    ok = True
    for radius, moore, inc in [(1, True, False), (2, False, True), (3, True, True)]:
        ok &= torch.equal(layer.neighborhood_sum(radius, moore, inc), ops.neighborhood_sum(counts, radius, moore, inc, torus)[lo:hi])
    ok &= layer.reduce("sum") == int(counts.sum().item()) and layer.reduce("max") == int(counts.max().item())
    field = ShardedLayer(heat[lo:hi].clone(), torus=torus)
    want = heat
    for _ in range(4):
        field.diffuse(0.35)
        want = ops.layer_diffuse(want, 0.35, torus)
    ok &= torch.equal(field.block, want[lo:hi])
    ok = agree(ok)
    ok_all &= ok
    if rank == 0:
        print(json.dumps({"case": f"ShardedLayer {rows}x{cols} torus={torus}", "world": world, "ok": ok}), flush=True)

# peer-memory halos: equal blocks (symmetric allocation)
rows, cols = 64, 512
for torus in (True, False):
    g = torch.Generator(device="cpu").manual_seed(7)
    heat = torch.rand((rows * world, cols), generator=g).to(dev)
    peer = PeerHaloLayer(rows, cols, torch.float32, 1, torus, dev)
    peer.state.copy_(heat[rank * rows:(rank + 1) * rows])
    peer.hdl.barrier()
    want = heat
    for _ in range(4):
        peer.step(lambda src, dst, top, bot: ops.layer_diffuse(src, 0.35, torus, out=dst, halo_top=top, halo_bot=bot))
        want = ops.layer_diffuse(want, 0.35, torus)
    ok = agree(torch.equal(peer.state, want[rank * rows:(rank + 1) * rows]))
    ok_all &= ok
    if rank == 0:
        print(json.dumps({"case": f"PeerHaloLayer diffusion {rows}x{cols} per rank torus={torus}", "world": world, "ok": ok}), flush=True)

# timing: 16384^2 float32 per GPU, weak scaling
G = 16384
gen = torch.Generator(device=dev).manual_seed(rank)
blk = torch.rand((G, G), device=dev, generator=gen)


def timed(fn, n):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


field = ShardedLayer(blk.clone(), torus=True)
ms = timed(lambda: field.diffuse(0.3), 20)
if rank == 0:
    print(json.dumps({"timing": "diffusion f32 16384^2 per GPU, NCCL halo exchange", "world": world, "ms_per_step": ms,
                      "cell_updates_per_s": G * G * world / ms * 1e3}), flush=True)
del field
peer = PeerHaloLayer(G, G, torch.float32, 1, True, dev)
peer.state.copy_(blk)
peer.hdl.barrier()
ms = timed(lambda: peer.step(lambda src, dst, top, bot: ops.layer_diffuse(src, 0.3, True, out=dst, halo_top=top, halo_bot=bot)), 20)
if rank == 0:
    print(json.dumps({"timing": "diffusion f32 16384^2 per GPU, peer-memory halos + device barrier", "world": world,
                      "ms_per_step": ms, "cell_updates_per_s": G * G * world / ms * 1e3}), flush=True)
    print("ALL OK" if ok_all else "MISMATCH", flush=True)
dist.destroy_process_group()
