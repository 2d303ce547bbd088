"""Dev: where does a ShardedBoids step spend its time? (torchrun)"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from mesa_b200 import ops
from mesa_b200.sharded import ShardedBoids

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
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
from torch.profiler import profile, ProfilerActivity
torch.cuda.synchronize(); dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        m.step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
dist.destroy_process_group()
