"""Developer micro-benchmark: Boids step at BASELINE.json config 4 scale (N agents, reference 'large' density)."""
import sys, os, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mesa_b200 import ops

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
side = 150.0 * math.sqrt(n / 400.0)
vision = 15.0
g = torch.Generator(device="cuda").manual_seed(0)
pos = torch.rand((n, 2), device="cuda", generator=g) * side
dirs = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
ws = ops.CellListWorkspace(n, (side, side), vision, "cuda")
po, do = torch.empty_like(pos), torch.empty_like(dirs)
cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.boids_step(pos, dirs, (side, side), vision=vision, separation=2.0, ws=ws, pos_out=po, dir_out=do, neighbor_count=cnt)
    pos, po = po, pos; dirs, do = do, dirs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ops.boids_step(pos, dirs, (side, side), vision=vision, separation=2.0, ws=ws, pos_out=po, dir_out=do)
    pos, po = po, pos; dirs, do = do, dirs
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"n": n, "side": side, "ms_per_step": ms, "agent_updates_per_s": n / ms * 1e3,
                  "mean_neighbors": float(cnt.float().mean()), "ws_MB": ws.nbytes / 1e6}))

# the same with the storage order re-sorted spatially every `every` steps (what BoidFlockers(activation="synchronous") does);
# every variant starts from the SAME initial flock (the step gets slower as the boids cluster, so simulation time matters)
g = torch.Generator(device="cuda").manual_seed(0)
pos0 = torch.rand((n, 2), device="cuda", generator=g) * side
dirs0 = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
for every in (0, 8, 4, 16, 1):
    pos, dirs = pos0.clone(), dirs0.clone()
    po, do = torch.empty_like(pos), torch.empty_like(dirs)
    calls = 0
    for timed in (False, True):        # 3 untimed steps, then `steps` timed ones
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps if timed else 3):
            if every and calls % every == 0:
                pos, dirs, _, _ = ops.boids_resort(pos, dirs, (side, side), vision=vision, ws=ws)
                po, do = torch.empty_like(pos), torch.empty_like(dirs)
            calls += 1
            ops.boids_step(pos, dirs, (side, side), vision=vision, separation=2.0, ws=ws, pos_out=po, dir_out=do)
            pos, po = po, pos; dirs, do = do, dirs
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"n": n, "resort_every": every, "ms_per_step": ms, "agent_updates_per_s": n / ms * 1e3}), flush=True)

# sequential activation (the reference's shuffle_do semantics): passes and time per step
if len(sys.argv) > 3 and sys.argv[3] == "seq":
    reach = ops.boids_reach(dirs, 1.0)
    sws = ops.BoidsSeqWorkspace(n, (side, side), vision, reach, "cuda")
    import time
    for epoch in range(2):
        ops.boids_step_sequential(pos, dirs, (side, side), seed=7, epoch=epoch, vision=vision, separation=2.0, reach=reach,
                                  ws=sws, pos_out=po, dir_out=do)
        pos, po = po, pos; dirs, do = do, dirs
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    passes = []
    for epoch in range(2, 2 + max(3, steps // 4)):
        _, _, p = ops.boids_step_sequential(pos, dirs, (side, side), seed=7, epoch=epoch, vision=vision, separation=2.0,
                                            reach=reach, ws=sws, pos_out=po, dir_out=do)
        passes.append(p)
        pos, po = po, pos; dirs, do = do, dirs
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / len(passes) * 1e3
    print(json.dumps({"mode": "sequential", "n": n, "ms_per_step": ms, "agent_updates_per_s": n / ms * 1e3,
                      "passes": passes, "reach": reach, "ws_MB": sws.nbytes / 1e6}))
