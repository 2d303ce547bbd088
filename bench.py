#!/usr/bin/env python
"""bench.py -- headline benchmark of the mesa_b200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1], Conway's Game of Life on a 16384 x 16384
uint8 torus layer; one "step" = one synchronous generation = 16384^2 agent-updates per GPU.
The layer is resident on the device in the stepping format of the backend: bit-packed (1 bit per
cell, csrc/gol_bits.cu), MB_GOL_T generations fused per launch (default 8); MB_GOL_PATH=u8 times
the uint8-layer kernel instead (csrc/gol.cu, the HBM-bound stencil; also reported under
`workloads`).  With N GPUs the grid is row-sharded weak-scaling style: N blocks of 16384 x 16384
stacked into a [16384*N, 16384] torus, one block per rank, deep ghost zones refreshed from the
neighbours' HBM over NVLink once per 64 generations (mesa_b200/sharding.py, PeerBitLife).

Prints ONE JSON line (rank 0): `value` = agent-updates/s over all ranks with inputs resident in
HBM (CUDA events, max over ranks); `e2e` = the same metric through the host-buffer API (pinned
host -> device copy of the layer, one generation, device -> host copy, every step);
`roofline` = algorithmic bytes (2 B per cell-update) / measured kernel time against
MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle port on a bounded sample (rank 0, N=1).

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the reference
itself is pure Python and is not on the GPU box) with all host threads on the same metric.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = 16384
ALG_BYTES_PER_UPDATE = 2  # uint8 read + uint8 write per cell-update
L2_NOTE = "ping-pong layers 2 x 268 MB per GPU > 126 MB L2 (inputs larger than L2, no flush needed)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.002):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }

            def run():
                while not self._stop.is_set():
                    try:
                        self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, nm in names.items():
                            if mask & bit:
                                self.reasons.add(nm)
                    except Exception:
                        pass
                    self._stop.wait(self.period)

            self._thr = threading.Thread(target=run, daemon=True)
            self._thr.start()
        except Exception:
            self._thr = None
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def _physical_gpu_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# --------------------------------------------------------------------------------------------
# CPU side (oracle port): cpu_baseline leg and --impl reference
# --------------------------------------------------------------------------------------------
def cpu_gol_rate(rows: int, cols: int, steps: int, warmup: int, threads: int):
    import numpy as np

    import oracle

    oracle.set_threads(threads)
    rng = np.random.Generator(np.random.Philox(0))
    a = (rng.random((rows, cols)) < 0.2).astype(np.uint8)
    for _ in range(warmup):
        a = oracle.gol_step(a, True)
    t0 = time.perf_counter()
    for _ in range(steps):
        a = oracle.gol_step(a, True)
    dt = time.perf_counter() - t0
    return rows * cols * steps / dt, dt / steps


def run_reference(args):
    """CPU arm: the oracle port of the reference's GoL step, all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # calibrate on a small slab, then size the per-step sample so K+W steps fit in ~100 s
    rate, _ = cpu_gol_rate(256, GRID, 1, 1, threads)
    budget = 100.0 / max(1, args.steps + args.warmup)
    rows = int(max(64, min(GRID, rate * budget / GRID)))
    rows = max(64, rows // 64 * 64)
    value, sec_per_step = cpu_gol_rate(rows, GRID, args.steps, args.warmup, threads)
    sample = f"gol torus slab {rows}x{GRID} uint8 per step (of the {GRID}x{GRID} layer), C oracle port, {threads} threads"
    line = {
        "impl": "reference", "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "gol_16384x16384_u8_torus", "grid": [GRID, GRID], "torus": True, "sample_rows": rows},
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------
def _free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bind_to_gpu_numa_node(physical_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (so that pinned host buffers are allocated
    there: with 8 ranks on a two-socket host every H2D / D2H copy otherwise crosses the socket interconnect).
    Returns (node, previous affinity) or (None, None)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(physical_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if not cpus:
            return None, None
        os.sched_setaffinity(0, cpus)
        return node, prev
    except Exception:
        return None, None


def _digest_u64(values, first_index: int):
    """Order-independent 64-bit digest of an integer tensor block whose first element has global index `first_index`:
    sum_i (v_i + 2) * (2 * mix(i) + 1)  mod 2^64  (sums of blocks add up, so it is the same for every sharding)."""
    import torch

    v = values.reshape(-1).to(torch.int64) + 2
    idx = torch.arange(first_index, first_index + v.numel(), dtype=torch.int64, device=v.device)
    w = (idx * -7046029254386353131 + 2135587861) | 1      # 0x9E3779B97F4A7C15 as int64; wraps mod 2^64
    return (v * w).sum()


def parity_checks(dev, world, rank):
    """Multi-GPU parity, OUTSIDE every timed region (results must be identical for every GPU count):
    (a) the row-sharded bit-packed Game of Life against a single-GPU recompute of the whole grid with the uint8 kernel;
    (b) row-sharded Schelling 1024^2, 5 steps, against the committed digests of the sequential CPU oracle
        (tests/golden/bench_parity.json; the same digests for 1, 2, 4 and 8 GPUs);
    (c) domain-decomposed Boids, 10^5 boids, 5 steps, against the single-GPU step (state within 1e-5).
    Returns flat scalars."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from mesa_b200 import ops
    from mesa_b200.sharded import ShardedBoids, ShardedSchelling
    from mesa_b200.sharding import PeerBitLife, block_bounds

    out = {}

    def all_ok(flag: bool) -> bool:
        t = torch.tensor([int(bool(flag))], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # (a) Game of Life: 1024 rows per rank, 100 generations = the initial ghost refresh + one more mid-run
    rows, cols, gens = 1024, 1024, 100
    g = torch.Generator(device="cpu").manual_seed(2024)
    full = (torch.rand((rows * world, cols), generator=g) < 0.3).to(torch.uint8).to(dev)
    if world > 1:
        life = PeerBitLife(rows, cols, True, dev, generations_per_launch=8, ghost=64)
    else:
        life = ops.BitLife(rows, cols, dev, torus=True, generations_per_launch=8)
    life.load(full[rank * rows:(rank + 1) * rows])
    life.run(gens)
    mine = life.store()
    a, b = full, torch.empty_like(full)
    for _ in range(gens):
        ops.gol_step(a, out=b, torus=True)
        a, b = b, a
    if world > 1:
        life.check()
        out["parity_gol_exchanges"] = life.exchanges
    out["parity_gol_ok"] = all_ok(torch.equal(mine, a[rank * rows:(rank + 1) * rows]))
    del life, full, a, b

    # (b) Schelling 1024^2 d=0.8 h=0.4 r=1, 5 steps
    with open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")) as f:
        gold = json.load(f)["schelling_1024"]
    R = C = gold["rows"]
    rng = np.random.Generator(np.random.Philox(gold["seed"]))
    u, v = rng.random((R, C)), rng.random((R, C))
    grid = np.where(u < 0.8, (v < 0.5).astype(np.int8), np.int8(-1)).astype(np.int8)
    lo, hi = block_bounds(R, world, rank)
    m = ShardedSchelling(R, C, grid[lo:hi], 0.4, 1, seed=gold["seed"], device=dev)
    ok = m.happy_count == gold["happy0"]
    digests = []
    for s in range(len(gold["steps"])):
        happy = m.step()
        occ = (m.type >= 0)
        d = torch.stack([_digest_u64(m.type, lo * C), _digest_u64(torch.where(occ, m.agent_id.to(torch.int64) & 0xFFFFFFFF, -1), lo * C)])
        dist.all_reduce(d)
        got = {"movers": m.last_movers, "happy": happy, "types": int(d[0].item()) & (2**64 - 1), "ids": int(d[1].item()) & (2**64 - 1)}
        digests.append(got)
        ok = ok and got == gold["steps"][s]
    out["parity_schelling_ok"] = all_ok(ok)
    out["parity_schelling_digest"] = f"{digests[-1]['types']:016x}"
    del m

    # (c) Boids 10^5, reference "large" density, vision 15, 5 steps
    n = 100_000
    side = 150.0 * math.sqrt(n / 400.0)
    rng = np.random.Generator(np.random.Philox(5))
    pos = (rng.random((n, 2)) * side).astype(np.float32)
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    owner = ShardedBoids.owner_of(torch.from_numpy(pos[:, 1]), side, world).numpy()
    sel = owner == rank
    sb = ShardedBoids(side, side, pos[sel], dirs[sel], np.arange(n)[sel], vision=15.0, separation=2.0, device=dev)
    p1, d1 = torch.from_numpy(pos).to(dev), torch.from_numpy(dirs).to(dev)
    for _ in range(5):
        sb.step()
        p1, d1 = ops.boids_step(p1, d1, (side, side), vision=15.0, separation=2.0)
    ids, gp, gd = sb.gather()
    err_p = float((gp - p1).abs().max().item()) / side
    err_d = float((gd - d1).abs().max().item())
    out["parity_boids_ok"] = all_ok(bool(torch.equal(ids, torch.arange(n, device=dev))) and err_p <= 1e-5 and err_d <= 1e-5)
    out["parity_boids_max_err"] = max(err_p, err_d)
    del sb
    out["parity_ok"] = bool(out["parity_gol_ok"] and out["parity_schelling_ok"] and out["parity_boids_ok"])
    return out


def sharded_workloads(dev, world, rank):
    """The other two workloads BASELINE.json's metric names, at this GPU count (flat scalars):
    Schelling config 5 weak-scaled (8192 x 65536 per GPU = 65536^2 at 8 GPUs, d=0.8 h=0.4 r=1, 5 steps from a random
    state) and Boids config 4 strong-scaled (10^7 boids in total, vision 15, synchronous activation)."""
    import torch
    import torch.distributed as dist

    from mesa_b200.sharded import ShardedBoids, ShardedSchelling

    out = {}

    def sync():
        torch.cuda.synchronize()
        dist.barrier()

    rows_per, cols, steps = 8192, 65536, 5
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    u = torch.rand((rows_per, cols), device=dev, generator=g)
    types = torch.where(u < 0.4, 0, torch.where(u < 0.8, 1, -1)).to(torch.int8)
    del u
    m = ShardedSchelling(rows_per * world, cols, types, 0.4, 1, seed=7, device=dev)
    del types
    reloc, assign, movers, passes = [], [], [], []
    for _ in range(steps):
        sync()
        t0 = time.perf_counter()
        m.move_unhappy()
        sync()
        t1 = time.perf_counter()
        m.assign_state()
        sync()
        t2 = time.perf_counter()
        reloc.append((t1 - t0) * 1e3)
        assign.append((t2 - t1) * 1e3)
        movers.append(m.last_movers)
        passes.append(m.last_passes)
    total = (sum(reloc) + sum(assign)) / 1e3
    out.update({"schelling_c5_grid_rows": rows_per * world, "schelling_c5_agents": m.n_agents,
                "schelling_c5_updates_per_s": m.n_agents * steps / total, "schelling_c5_steps": steps,
                "schelling_c5_sec": total, "schelling_relocate_ms_first": reloc[0], "schelling_relocate_ms_last": reloc[-1],
                "schelling_assign_ms": min(assign), "schelling_c5_movers_first": movers[0],
                "schelling_c5_passes_first": passes[0]})
    del m
    torch.cuda.empty_cache()

    N = 10_000_000
    side = 150.0 * math.sqrt(N / 400.0)
    n_local = N // world
    pos = torch.rand((n_local, 2), device=dev, generator=g)
    pos[:, 0] *= side
    pos[:, 1] = (pos[:, 1] + rank) * (side / world)
    dirs = torch.rand((n_local, 2), device=dev, generator=g) * 2 - 1
    b = ShardedBoids(side, side, pos, dirs, torch.arange(n_local, device=dev) + rank * n_local, vision=15.0,
                     separation=2.0, device=dev)
    for _ in range(3):
        b.step()
    sync()
    t0 = time.perf_counter()
    for _ in range(20):
        b.step()
    sync()
    dt = time.perf_counter() - t0
    out.update({"boids_1e7_ms_per_step": dt / 20 * 1e3, "boids_1e7_updates_per_s": N * 20 / dt})
    del b
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mesa_b200 import ops
    from mesa_b200.host_api import GolHostStepper, ShardedGolHostStepper
    from mesa_b200.sharding import FusedPeerGol, HaloExchanger, PeerBitLife, PeerHaloLayer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mesa_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]   # the image's default prints a banner on stdout; rank 0 prints ONE JSON line
    have_group = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        # a 1-rank group: the sharded drivers (parity checks, Schelling / Boids workloads) need symmetric memory
        try:
            dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                                    device_id=dev)
        except Exception:
            have_group = False
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # synthetic initial state: alive with p=0.2 (conways_game_of_life/model.py:9-15), Philox-seeded
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    a = (torch.rand((GRID, GRID), device=dev, generator=g) < 0.2).to(torch.uint8)
    b = torch.empty_like(a)
    path = os.environ.get("MB_GOL_PATH", "bits")
    gens_per_launch = int(os.environ.get("MB_GOL_T", "8")) if path == "bits" else 1
    halo_mode = os.environ.get("MB_HALO", "fused") if (world > 1 and path != "bits") else "none"
    hx = HaloExchanger(GRID, torch.uint8, 1, True, dev) if halo_mode not in ("none", "fused", "peer") else None
    peer = fused = life = None
    if path == "bits":
        # bit-packed planes, gens_per_launch generations per launch; N > 1: deep ghost zones (64 rows of each ring
        # neighbour, copied out of its HBM over NVLink once per 64 generations by one kernel with a flag hand-shake)
        if world > 1:
            life = PeerBitLife(GRID, GRID, True, dev, generations_per_launch=gens_per_launch,
                               ghost=int(os.environ.get("MB_GOL_GHOST", "0")) or None,
                               transport=os.environ.get("MB_GHOST_TRANSPORT", "peer"))
        else:
            life = ops.BitLife(GRID, GRID, dev, torus=True, generations_per_launch=gens_per_launch)
        life.load(a)
    elif halo_mode == "fused":
        fused = FusedPeerGol(GRID, GRID, True, dev)
        fused.load(a)
    elif halo_mode == "peer":
        peer = PeerHaloLayer(GRID, GRID, torch.uint8, 1, True, dev)
        peer.state.copy_(a)

    def gol_rows(src, dst, top, bot):
        ops.gol_step(src, out=dst, torus=False, col_torus=True,
                     halo_top=None if top is None else top[-1], halo_bot=None if bot is None else bot[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def advance():
        nonlocal a, b
        if fused is not None:
            fused.step()
        elif peer is not None:
            peer.step(gol_rows)
        elif hx is not None:
            top, bot = hx.exchange(a)
            ops.gol_step(a, out=b, torus=True, halo_top=top[0], halo_bot=bot[0])
            a, b = b, a
        else:
            ops.gol_step(a, out=b, torus=True)
            a, b = b, a

    def kernel_count():
        if life is None:
            return 0
        n = life.launches
        if world > 1:
            n = life.kernel_launches + life.layer.exchanges + life.layer.waits
        return n

    warm = max(args.warmup, 3)
    use_graph = life is not None and os.environ.get("MB_BENCH_GRAPH", "1") != "0"
    graph = None
    if life is not None:
        life.run(warm)
        # the launch sequence of the timed region, once, untimed: loads every kernel variant it uses (CUDA loads
        # kernels lazily) -- the T-generation stencil, the shorter tail variants and, N > 1, the ghost exchange
        if world > 1:
            life.layer.invalidate()
        life.run(args.steps, even_launches=use_graph)
        barrier()
        if use_graph:
            # the K generations are a static launch sequence (ping-pong planes, T generations per launch; N > 1: ONE ghost
            # refresh with its device-side hand-shake first): captured once, replayed once untimed (uploads the graph),
            # then replayed inside the timed region.  The same at every GPU count.
            if world > 1:
                life.layer.invalidate()           # the timed window starts with stale ghost rows: it contains the exchange
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            k0 = kernel_count()
            x0 = life.exchanges if world > 1 else 0
            with torch.cuda.graph(graph, stream=side):
                life.run(args.steps, even_launches=True)          # captured, not executed
            torch.cuda.current_stream(dev).wait_stream(side)
            launches = kernel_count() - k0
            exchanges = (life.exchanges - x0) if world > 1 else 0
            graph.replay()
            barrier()
    else:
        for _ in range(warm):
            advance()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(_physical_gpu_index(local)) as clocks:
        if world > 1 and life is not None and getattr(life.layer, "hdl", None) is not None:
            # the host barrier above leaves the ranks tens of microseconds apart, which a 0.2 ms window that opens with an
            # exchange would book as waiting time of the early ranks: line the GPUs up with a device-side barrier
            # (symmetric memory, one small kernel per rank) BEFORE the first event is recorded
            life.layer.hdl.barrier()
        if graph is not None:
            e0.record()
            graph.replay()
            e1.record()
        elif life is not None:
            if world > 1:
                life.layer.invalidate()
            k0, x0 = kernel_count(), (life.exchanges if world > 1 else 0)
            e0.record()
            life.run(args.steps)
            e1.record()
            launches = kernel_count() - k0
            exchanges = (life.exchanges - x0) if world > 1 else 0
        else:
            e0.record()
            for _ in range(args.steps):
                advance()
            e1.record()
            launches, exchanges = args.steps, (args.steps if world > 1 else 0)
        barrier()
    if fused is not None:
        fused.check()
    if life is not None and world > 1:
        life.check()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    updates = GRID * GRID * args.steps * world
    value = updates / (ms * 1e-3)
    ms_per_step = ms / args.steps
    stencil_launches = len([t for t in ops.gol_launch_plan(args.steps, gens_per_launch, even=use_graph) if t]) if life is not None else args.steps

    # post-timing verification of the replayed / stepped state: one more generation on the bit planes must equal one
    # uint8-kernel generation of the unpacked state (N = 1), and the e2e leg below checks the host path against it
    # ---- e2e: host buffers in, host buffers out, every step (public API: mesa_b200.host_api; H2D / stencil / D2H
    # pipelined over row chunks so both PCIe directions stay busy).  N > 1 is the REAL sharded generation: every rank
    # steps its own block of the [16384 N, 16384] torus from / to its own pinned host buffers, the halo rows come
    # from the ring neighbours' edge rows (uploaded first into symmetric memory, read in place over NVLink).
    e2e_steps = max(3, min(args.steps, 20))
    node, prev_aff = _bind_to_gpu_numa_node(_physical_gpu_index(local))
    h_in = torch.empty((GRID, GRID), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((GRID, GRID), dtype=torch.uint8).pin_memory()
    cur_state = life.store() if life is not None else fused.state if fused is not None else peer.state if peer is not None else a
    h_in.copy_(cur_state.cpu())
    h_out.zero_()
    if prev_aff is not None:
        os.sched_setaffinity(0, prev_aff)
    if world > 1:
        stepper = ShardedGolHostStepper(GRID, GRID, dev, chunks=16, torus=True)
    else:
        stepper = GolHostStepper(GRID, GRID, dev, chunks=16)

    def e2e_step():
        stepper.step(h_in, h_out)
        torch.cuda.current_stream().synchronize()   # the result is in host memory when the step returns

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = GRID * GRID * e2e_steps * world / e2e_dt
    # the host path's generation == the resident path's generation of the same state (every rank, bit for bit)
    e2e_ok = None
    if life is not None:
        life.run(1)
        same = torch.tensor([int(torch.equal(h_out.to(dev), life.store()))], device=dev)
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        e2e_ok = bool(same.item())
    del h_in, h_out, stepper, cur_state

    extras, notes = {}, []
    if not args.no_others and have_group:
        for fn in (parity_checks, sharded_workloads):
            try:
                extras.update(fn(dev, world, rank))
            except Exception as exc:  # reported, never fatal for the headline line
                notes.append(f"{fn.__name__}: {exc!r}"[:300])
                if fn is parity_checks:
                    extras["parity_ok"] = False
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass

    if rank == 0:
        peak, peak_src = _peaks()
        details = []
        if world == 1 and not args.no_others:
            try:
                details = other_workloads(dev) + layer_workloads(dev) + agent_workloads(dev)
            except Exception as exc:
                notes.append(f"workloads: {exc!r}"[:300])
        flat = _flat_workload_keys(details)
        flat.update(extras)
        if life is not None:
            kernel = f"gol_bits_kernel<{gens_per_launch}>"
            par = (f"row-sharded x{world}, bit-packed planes, {life.ghost} ghost rows of each ring neighbour copied out of its "
                   f"HBM over NVLink once per {life.ghost} generations by one kernel with a flag hand-shake in peer memory "
                   "(csrc/ghost.cu); no barrier, no collective") if world > 1 else "single GPU"
            # physical roofline of the bit-packed kernel: its state is 1 bit per cell, so one launch must read and write
            # one 32 MiB plane whatever the number of fused generations; the contract's 2 B per cell-update figure
            # (the uint8 layer) is reported beside it as `contract_gbs`
            us_launch = ms * 1e3 / stencil_launches
            alg_bytes = 2 * GRID * GRID / 8
            achieved = alg_bytes / (us_launch * 1e-6) / 1e9
            sm_mhz = clocks.summary()["sm_mhz"] or 1965
            alu_peak = 148 * 4 * 16 * sm_mhz * 1e6                      # thread-instructions / s on the ALU pipe
            alu_achieved = 12.0 * (GRID * GRID // 32) / (ms_per_step * 1e-3)   # 12 LOP3 / SHF per word-generation
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": _ncu_traffic("gol_bits_kernel"), "peak_source": peak_src, "kernel": kernel,
                    "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": us_launch,
                    "binding_pipe": "alu (integer LOP3/SHF issue), not HBM: the planes are L2-resident",
                    "alu_pipe_frac": alu_achieved / alu_peak,
                    "contract_gbs": ALG_BYTES_PER_UPDATE * GRID * GRID / (ms_per_step * 1e-3) / 1e9,
                    "note": "frac = bit-plane bytes one launch must move / launch time / HBM peak (physical, small by design: "
                            "8 generations are fused per launch and the kernel is bound by the integer pipe, alu_pipe_frac); "
                            "the HBM-bound grid stencils are gol_u8_hbm_frac / happy_r1_hbm_frac / happy_r2_hbm_frac"}
        else:
            kernel = "gol_step_u8_vec"
            par = (f"row-sharded x{world}, ring halos " + ("read in place from peer HBM over NVLink, hand-shake fused into the stencil kernel (1 launch / generation, no collective)" if fused is not None else "read in place from peer HBM over NVLink (symmetric memory) + device barrier" if peer is not None else "exchanged with NCCL send/recv")) if world > 1 else "single GPU"
            alg_bytes = ALG_BYTES_PER_UPDATE * GRID * GRID
            achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": _ncu_traffic(kernel), "peak_source": peak_src, "kernel": kernel,
                    "algorithmic_bytes_per_launch": alg_bytes, "note": "uint8 layer, one generation per launch"}
        for k in ("gol_u8_hbm_frac", "happy_r1_hbm_frac", "happy_r2_hbm_frac"):
            if k in flat:
                roof[k] = flat[k]
        line = {
            "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "gol_16384x16384_u8_torus", "grid_rows_per_gpu": GRID, "grid_cols": GRID,
                       "global_rows": GRID * world, "torus": True, "device_format": "bits" if life is not None else "u8",
                       "generations_per_launch": gens_per_launch, "parallelism": par,
                       "launch": "one CUDA graph replay of the K-generation launch sequence (same at every GPU count)" if graph is not None else "stream launches",
                       "exchanges": exchanges,
                       "l2": L2_NOTE if life is None else "bit planes (2 x 32 MiB) are L2-resident by design; no flush: the resident working set IS the workload (1000 consecutive generations of one layer)"},
            "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": GRID * GRID * world,
                    "d2h_bytes_per_step": GRID * GRID * world, "steps": e2e_steps, "ok": e2e_ok, "numa_node": node,
                    "api": ("mesa_b200.host_api.ShardedGolHostStepper.step (every rank: pinned host block in/out, halo rows from the "
                            "ring neighbours over NVLink, 16 row chunks, 3 streams)") if world > 1 else
                           "mesa_b200.host_api.GolHostStepper.step (pinned host in/out, 16 row chunks, 3 streams)"},
            "gpu_launches": launches,
            "exchanges": exchanges,
            "clocks": clocks.summary(),
            "roofline": roof,
        }
        line.update(flat)
        line["config"].update({k: v for k, v in flat.items() if k.startswith("parity")})
        line["roofline"].update({k: v for k, v in flat.items() if not k.startswith("parity") and k not in roof})
        if notes:
            line["notes"] = notes
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, _ = cpu_gol_rate(256, GRID, 1, 1, threads)
            rows = max(64, min(GRID, int(rate * 5.0 / GRID)) // 64 * 64)  # ~5 s per step, 3 steps
            v, _ = cpu_gol_rate(rows, GRID, 3, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "agent-updates/s", "cores": threads, "kind": "port",
                                    "sample": f"gol torus slab {rows}x{GRID} uint8, 3 steps, C oracle port"}
        if details:
            # the per-workload detail goes to a side file and to stderr; stdout carries ONE compact JSON line
            try:
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                with open(os.path.join(ROOT, "gpurun_out", f"bench_workloads_n{world}.json"), "w") as f:
                    json.dump(details, f, indent=1)
            except Exception:
                pass
            print(json.dumps({"workloads": details}), file=sys.stderr, flush=True)
        print(json.dumps(line), flush=True)
    if dist.is_initialized():
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()


def _flat_workload_keys(details):
    """Flat scalars of the single-GPU workload list that the driver's record keeps (it drops nested lists)."""
    flat = {}
    names = {"gol_u8_layer": "gol_u8", "schelling_assign_state_16384x16384_r1": "happy_r1",
             "schelling_assign_state_16384x16384_r2_h1": "happy_r2"}
    for w in details:
        name = w.get("workload", "")
        for key, short in names.items():
            if name.startswith(key) and "hbm_frac_of_measured_peak" in w:
                flat[f"{short}_hbm_frac"] = w["hbm_frac_of_measured_peak"]
                flat[f"{short}_us"] = w["ms_per_step"] * 1e3
        if name.startswith("schelling_100x100") and "with_datacollector" not in name:
            flat["schelling_c1_us_per_step"] = w["ms_per_step"] * 1e3
            flat["schelling_c1_updates_per_s"] = w["agent_updates_per_s"]
            if "ms_per_step_call_by_call" in w:
                flat["schelling_c1_us_per_step_call_by_call"] = w["ms_per_step_call_by_call"] * 1e3
        if name.startswith("schelling_8192"):
            flat["schelling_8192_first_step_ms"] = w["first_step_ms"]
        if name.startswith("boltzmann_1e6"):
            flat["boltzmann_1e6_ms_per_step"] = w["ms_per_step"]
            flat["boltzmann_1e6_updates_per_s"] = w["agent_updates_per_s"]
        if name.startswith("wolf_sheep_1e6") and "ms_per_step" in w:
            flat["wolf_sheep_ms_per_step"] = w["ms_per_step"]
            flat["wolf_sheep_updates_per_s"] = w["agent_updates_per_s"]
        if name.startswith("sugarscape_g1mt_512") and "ms_per_step" in w:
            flat["sugarscape_512_ms_per_step"] = w["ms_per_step"]
            flat["sugarscape_512_updates_per_s"] = w["agent_updates_per_s"]
        if name == "epstein_1024x1024_vision7" and "ms_per_step" in w:
            flat["epstein_1024_ms_per_step"] = w["ms_per_step"]
            flat["epstein_1024_updates_per_s"] = w["agent_updates_per_s"]
        if name == "epstein_40x40_vision7" and "ms_per_step" in w:
            flat["epstein_40_ms_per_step"] = w["ms_per_step"]
        if name == "boids_1e7_vision15_synchronous":
            flat["boids_1e7_single_gpu_ms_per_step"] = w["ms_per_step"]
        if name == "boids_1e7_vision15_synchronous_random_storage_order":
            flat["boids_1e7_unsorted_ms_per_step"] = w["ms_per_step"]
        if name.startswith("boids_1e7_vision15_sequential"):
            flat["boids_1e7_sequential_ms_per_step"] = w["ms_per_step"]
    return flat


def other_workloads(dev):
    """Short measurements of the other BASELINE.json configs on one GPU (same run, CUDA-event / wall timed).
    The headline `value` stays the GoL line; these are reported for context (agent-updates/s)."""
    import math

    import torch

    from mesa_b200 import ops
    from mesa_b200.examples import BoltzmannScenario, BoltzmannWealth, Schelling, SchellingScenario

    out = []

    def timed(fn, steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps

    # the uint8-layer GoL stencil (HBM-bound: 2 B per cell-update, layers larger than L2)
    peak, _ = _peaks()
    g0 = torch.Generator(device=dev).manual_seed(7)
    ua = (torch.rand((GRID, GRID), device=dev, generator=g0) < 0.2).to(torch.uint8)
    ub = torch.empty_like(ua)
    for _ in range(5):
        ops.gol_step(ua, out=ub, torus=True)
        ua, ub = ub, ua
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        ops.gol_step(ua, out=ub, torus=True)
        ua, ub = ub, ua
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    out.append({"workload": "gol_u8_layer_16384x16384 (gol_step_u8_vec, 1 generation per launch)", "ms_per_step": ms,
                "agent_updates_per_s": GRID * GRID / ms * 1e3, "hbm_gbs": 2 * GRID * GRID / ms / 1e6,
                "hbm_frac_of_measured_peak": 2 * GRID * GRID / ms / 1e6 / peak, "traffic": _ncu_traffic("gol_step_u8_vec")})
    del ua, ub
    # config 0: Schelling 100x100, density 0.8, 100 steps (launch-latency bound: 10 KB of state)
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
    n = len(m.agents)
    t_call = timed(m.step, 100)
    # the reference's own benchmark protocol is model.run_for(steps) (benchmarks/global_benchmark.py:17-84): on the
    # single-launch path the 100 steps are ONE launch
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
    m.run_for(2)
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.run_for(100)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 100
    out.append({"workload": "schelling_100x100_d0.8_h0.4_r1_100steps", "agents": n, "ms_per_step": t * 1e3,
                "agent_updates_per_s": n / t, "ms_per_step_call_by_call": t_call * 1e3, "happy_after_100": int(m.happy),
                "note": "model.run_for(100) = one launch (csrc/schelling_small.cu); call by call = 100 launches of the same "
                        "kernel issued from Python; latency-bound (L2-resident), roofline fraction not meaningful"})
    # the same with the reference's DataCollector inside step() (model.py:92; 4 model reporters + 1 agent column per step)
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=True)
    t = timed(m.step, 100)
    out.append({"workload": "schelling_100x100_d0.8_h0.4_r1_100steps_with_datacollector", "agents": n,
                "ms_per_step": t * 1e3, "agent_updates_per_s": n / t})
    # Schelling 8192^2, d=0.8 h=0.4 r=1: relocation at scale on one GPU -- the first step from a random state
    # (~15 M movers) and the four after it (shuffle_do("step") + do("assign_state") each)
    g = torch.Generator(device=dev).manual_seed(0)
    R8 = 8192
    u = torch.rand((R8, R8), device=dev, generator=g)
    t8 = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
    del u
    occ = (t8 >= 0).reshape(-1)
    n8 = int(occ.sum().item())
    aid = torch.where(occ, torch.cumsum(occ.to(torch.int32), 0, dtype=torch.int32) - 1, 0).reshape(R8, R8).contiguous()
    del occ
    h8, c8 = ops.schelling_happy(t8, 0.4, 1)
    ws8 = ops.RelocationWorkspace(t8, n8)
    # untimed warm-up on a small crowded grid: loads every kernel of the relocation (CUDA loads modules lazily; the event
    # loop and the fallback kernels only run when movers miss all 50 probes) so that the first timed step measures the step
    gw = torch.Generator(device=dev).manual_seed(5)
    uw = torch.rand((256, 256), device=dev, generator=gw)
    tw = torch.where(uw < 0.09, -1, torch.where(uw < 0.55, 0, 1)).to(torch.int8)
    occw = (tw >= 0).reshape(-1)
    aidw = torch.where(occw, torch.cumsum(occw.to(torch.int32), 0, dtype=torch.int32) - 1, 0).reshape(256, 256).contiguous()
    hw, _ = ops.schelling_happy(tw, 0.9, 1)
    ops.schelling_relocate(tw, hw, aidw, ops.RelocationWorkspace(tw, int(occw.sum().item())), 7, 0)
    del uw, tw, occw, aidw, hw
    per = []
    for epoch in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mv, its = ops.schelling_relocate(t8, h8, aid, ws8, 7, epoch)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ops.schelling_happy(t8, 0.4, 1, out=h8, count=c8)
        torch.cuda.synchronize()
        per.append({"movers": mv, "iterations": its, "relocate_ms": (t1 - t0) * 1e3, "assign_ms": (time.perf_counter() - t1) * 1e3})
    tot = sum(p["relocate_ms"] + p["assign_ms"] for p in per) / 1e3
    out.append({"workload": "schelling_8192x8192_d0.8_h0.4_r1_5steps", "agents": n8, "first_step_ms": per[0]["relocate_ms"] + per[0]["assign_ms"],
                "ms_per_step": tot / 5 * 1e3, "agent_updates_per_s": n8 * 5 / tot, "per_step": per})
    del t8, h8, aid, ws8
    torch.cuda.empty_cache()
    # Schelling happiness stencil alone on 16384^2 (the >= 60 % target of the north star)
    u = torch.rand((GRID, GRID), device=dev, generator=g)
    types = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
    del u
    happy = torch.empty((GRID, GRID), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    tbl = ops.happy_threshold_table(0.4, 1)
    for _ in range(3):
        ops.schelling_happy(types, 0.4, 1, False, out=happy, count=cnt, table=tbl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.schelling_happy(types, 0.4, 1, False, out=happy, count=cnt, table=tbl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    peak, _ = _peaks()
    occupied = int((types >= 0).sum().item())
    out.append({"workload": "schelling_assign_state_16384x16384_r1", "ms_per_step": ms, "agents": occupied,
                "agent_updates_per_s": occupied / ms * 1e3, "hbm_gbs": 2 * GRID * GRID / ms / 1e6,
                "hbm_frac_of_measured_peak": 2 * GRID * GRID / ms / 1e6 / peak})
    # the reference's "large" Schelling variant: radius 2, homophily 1 (benchmarks/configurations.py:42-49)
    tbl2 = ops.happy_threshold_table(1.0, 2)
    for _ in range(3):
        ops.schelling_happy(types, 1.0, 2, False, out=happy, count=cnt, table=tbl2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.schelling_happy(types, 1.0, 2, False, out=happy, count=cnt, table=tbl2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    out.append({"workload": "schelling_assign_state_16384x16384_r2_h1", "ms_per_step": ms, "agents": occupied,
                "agent_updates_per_s": occupied / ms * 1e3, "hbm_gbs": 2 * GRID * GRID / ms / 1e6,
                "hbm_frac_of_measured_peak": 2 * GRID * GRID / ms / 1e6 / peak})
    del types, happy
    # Wolf-Sheep predation (typed activation + predation on dynamic populations), 10^6 sheep and
    # 2 x 10^5 wolves on 2048^2 with grass; agent-updates = animals activated
    try:
        from mesa_b200.examples.wolf_sheep import Sheep, Wolf, WolfSheep, WolfSheepScenario

        ws_m = WolfSheep(WolfSheepScenario(width=2048, height=2048, initial_sheep=1_000_000, initial_wolves=200_000, rng=3))
        ws_m.step()
        acted, passes = 0, []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            acted += len(ws_m.agents_by_type[Sheep]) + len(ws_m.agents_by_type[Wolf])
            ws_m.step()
            passes.append((ws_m.last_sheep_passes, ws_m.last_wolf_passes))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append({"workload": "wolf_sheep_1e6_sheep_2e5_wolves_2048x2048_grass", "ms_per_step": dt / 10 * 1e3,
                    "agent_updates_per_s": acted / dt, "sheep": len(ws_m.agents_by_type[Sheep]),
                    "wolves": len(ws_m.agents_by_type[Wolf]), "jacobi_passes_sheep_wolf": passes[-1],
                    "note": "sequential shuffle_do semantics of both phases as Jacobi fixed points (csrc/wolf_sheep.cu); "
                            "host-driven passes + population compaction: latency-bound at this size"})
        del ws_m
    except Exception as exc:  # reported, never fatal
        out.append({"workload": "wolf_sheep", "error": repr(exc)[:200]})
    # Epstein civil violence (radius-7 von Neumann neighbourhoods, citizens + cops in one sequential activation)
    try:
        import numpy as np

        from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence, EpsteinScenario

        for side in (40, 1024):
            rs = np.random.default_rng(1)
            u = rs.random(side * side)
            occupied = np.flatnonzero(u < 0.774)                       # the scenario's densities: 0.7 citizens + 0.074 cops
            kind = (u[occupied] >= 0.7).astype(np.uint8)
            init = {"kind": kind, "cell": occupied, "hardship": np.where(kind == 0, rs.random(len(kind)), 0.0),
                    "risk_aversion": np.where(kind == 0, rs.random(len(kind)), 0.0)}
            em = EpsteinCivilViolence(side, side, EpsteinScenario(rng=1), initial_state=init)
            for _ in range(3):
                em.agents.shuffle_do("step")
            t = timed(lambda: em.agents.shuffle_do("step"), 20)
            out.append({"workload": f"epstein_{side}x{side}_vision7", "ms_per_step": t * 1e3, "agents": em.agents.n,
                        "agent_updates_per_s": em.agents.n / t, "active_quiet_arrested": [em.ACTIVE, em.QUIET, em.ARRESTED],
                        "note": "one launch: rank-ordered dependency graph, one warp per agent (csrc/epstein.cu); bound by the "
                                "depth of the dependency chains (2 x vision reach), not by bandwidth"})
            del em
    except Exception as exc:  # reported, never fatal
        out.append({"workload": "epstein", "error": repr(exc)[:200]})
    # Sugarscape G1MT (regrowth, move / eat / die, trade) on a synthetic 512 x 512 landscape, 40 000 traders
    try:
        import numpy as np

        from mesa_b200.examples.sugarscape_g1mt import SugarScapeScenario, SugarscapeG1mt

        rs = np.random.default_rng(1)
        ns, side = 40_000, 512
        init = {"cell": rs.choice(side * side, size=ns, replace=False), "sugar": rs.integers(25, 51, ns), "spice": rs.integers(25, 51, ns),
                "metabolism_sugar": rs.integers(1, 6, ns), "metabolism_spice": rs.integers(1, 6, ns), "vision": rs.integers(1, 6, ns)}
        sm = SugarscapeG1mt(SugarScapeScenario(enable_trade=True, rng=1), width=side, height=side, initial_state=init)
        for _ in range(2):
            sm.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acted = trades = 0
        for _ in range(10):
            acted += 2 * len(sm.agents)                                # two activations per step (move, trade)
            sm.step()
            trades += len(sm.trade_records()[2])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append({"workload": "sugarscape_g1mt_512x512_40000_traders_with_trade", "ms_per_step": dt / 10 * 1e3,
                    "agent_updates_per_s": acted / dt, "traders_end": len(sm.agents), "trades_per_step": trades / 10,
                    "inexact_welfares": sm.inexact_welfares,
                    "note": "two rank-ordered dependency graphs per step (csrc/sugarscape.cu) + the host side of the reference's "
                            "collector (per-trader partner lists); the starved are compacted away every step"})
        del sm
    except Exception as exc:  # reported, never fatal
        out.append({"workload": "sugarscape", "error": repr(exc)[:200]})
    # config 2: Boltzmann wealth, 10^6 agents on 4096^2
    m = BoltzmannWealth(BoltzmannScenario(n=1_000_000, width=4096, height=4096, rng=1))
    m.run_for(3)
    t = timed(m.step, 30)
    out.append({"workload": "boltzmann_1e6_agents_4096x4096", "ms_per_step": t * 1e3, "agent_updates_per_s": 1e6 / t})
    del m
    # config 3: Boids, 10^7 agents at the reference "large" density (vision 15).  Both variants below start from the SAME
    # initial flock (the step gets slower as the boids cluster, so simulation time matters): 3 untimed + 16 timed steps
    N = 10_000_000
    side = 150.0 * math.sqrt(N / 400.0)
    pos0 = torch.rand((N, 2), device=dev, generator=g) * side
    dirs0 = torch.rand((N, 2), device=dev, generator=g) * 2 - 1
    ws = ops.CellListWorkspace(N, (side, side), 15.0, dev)
    state = [None, None, None, None]

    def boids_run(resort_every):
        state[0], state[1] = pos0.clone(), dirs0.clone()
        state[2], state[3] = torch.empty_like(pos0), torch.empty_like(dirs0)
        calls = [0]

        def one():
            # what BoidFlockers(activation="synchronous") does: every 8th step the storage order is rewritten in
            # cell-list order (ops.boids_resort, inside the timed region), which keeps the gathers / scatters coalesced
            if resort_every and calls[0] % resort_every == 0:
                state[0], state[1], _, _ = ops.boids_resort(state[0], state[1], (side, side), vision=15.0, ws=ws)
                state[2], state[3] = torch.empty_like(state[0]), torch.empty_like(state[1])
            calls[0] += 1
            ops.boids_step(state[0], state[1], (side, side), vision=15.0, separation=2.0, ws=ws, pos_out=state[2], dir_out=state[3])
            state[0], state[2] = state[2], state[0]
            state[1], state[3] = state[3], state[1]

        for _ in range(3):
            one()
        return timed(one, 16)

    t = boids_run(8)
    out.append({"workload": "boids_1e7_vision15_synchronous", "ms_per_step": t * 1e3, "agent_updates_per_s": N / t,
                "note": "storage order re-sorted spatially every 8 steps (2 re-sorts inside the 16 timed steps)"})
    t = boids_run(0)
    out.append({"workload": "boids_1e7_vision15_synchronous_random_storage_order", "ms_per_step": t * 1e3,
                "agent_updates_per_s": N / t, "note": "the round-1 figure: storage order unrelated to space"})
    # the same flock in the reference's SEQUENTIAL shuffle_do order (exact replay mode, rank-ordered chained kernel), from
    # the same point of the simulation as round 1 measured it (12 synchronous steps after the start)
    state[0], state[1] = pos0.clone(), dirs0.clone()
    state[2], state[3] = torch.empty_like(pos0), torch.empty_like(dirs0)
    for _ in range(12):
        ops.boids_step(state[0], state[1], (side, side), vision=15.0, separation=2.0, ws=ws, pos_out=state[2], dir_out=state[3])
        state[0], state[2] = state[2], state[0]
        state[1], state[3] = state[3], state[1]
    del pos0, dirs0
    reach = ops.boids_reach(state[1], 1.0)
    sws = ops.BoidsSeqWorkspace(N, (side, side), 15.0, reach, dev)
    epoch = [0]

    def boids_seq():
        ops.boids_step_sequential(state[0], state[1], (side, side), seed=11, epoch=epoch[0], vision=15.0, separation=2.0,
                                  reach=reach, ws=sws, pos_out=state[2], dir_out=state[3])
        epoch[0] += 1
        state[0], state[2] = state[2], state[0]
        state[1], state[3] = state[3], state[1]

    boids_seq()
    t = timed(boids_seq, 4)
    out.append({"workload": "boids_1e7_vision15_sequential_shuffle_do", "ms_per_step": t * 1e3, "agent_updates_per_s": N / t,
                "note": "identical to the reference's sequential activation in the Philox order (dependency-chained kernel)"})
    return out


def layer_workloads(dev, grid=GRID, reps=20):
    """Property-layer kernels (csrc/layers.cu) on grid x grid layers: CUDA-event time per launch and the
    ALGORITHMIC bytes each launch must move (inputs read once + outputs written once) against the HBM peak."""
    import torch

    from mesa_b200 import ops

    peak, _ = _peaks()
    out = []
    n = grid * grid

    def run(name, fn, alg_bytes):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = alg_bytes / ms / 1e6
        out.append({"workload": f"layer_{name}_{grid}x{grid}", "ms_per_step": ms, "cell_updates_per_s": n / ms * 1e3,
                    "algorithmic_bytes": alg_bytes, "hbm_gbs": gbs, "hbm_frac_of_measured_peak": gbs / peak})

    g = torch.Generator(device=dev).manual_seed(3)
    u8 = (torch.rand((grid, grid), device=dev, generator=g) * 8).to(torch.uint8)
    u8b = torch.empty_like(u8)
    i32 = torch.empty((grid, grid), dtype=torch.int32, device=dev)
    run("fill_u8", lambda: ops.layer_fill(u8b, 3), n)
    run("affine_clamp_u8 (sugar regrowth: min(x + 1, cap))", lambda: ops.layer_affine_clamp(u8, 1.0, 1.0, 0.0, 4.0, out=u8b), 2 * n)
    run("binary_min_u8", lambda: ops.layer_binary(u8, u8b, "min", out=u8b), 3 * n)
    run("reduce_count_nonzero_u8 (number of empty cells)", lambda: ops.layer_reduce(u8, "count_nonzero"), n)
    run("reduce_sum_u8", lambda: ops.layer_reduce(u8, "sum"), n)
    run("reduce_max_u8", lambda: ops.layer_reduce(u8, "max"), n)
    run("neighborhood_sum_u8_moore_r1_torus", lambda: ops.neighborhood_sum(u8, 1, True, False, True, out=i32), 5 * n)
    run("neighborhood_sum_u8_vonneumann_r2_clamped", lambda: ops.neighborhood_sum(u8, 2, False, False, False, out=i32), 5 * n)
    del u8, u8b, i32
    f32 = torch.rand((grid, grid), device=dev, generator=g)
    f32b = torch.empty_like(f32)
    f64 = torch.empty((grid, grid), dtype=torch.float64, device=dev)
    run("fill_f32", lambda: ops.layer_fill(f32b, 0.5), 4 * n)
    run("affine_clamp_f32", lambda: ops.layer_affine_clamp(f32, 0.9, 0.01, 0.0, 1.0, out=f32b), 8 * n)
    run("binary_add_f32", lambda: ops.layer_binary(f32, f32b, "add", out=f32b), 12 * n)
    run("reduce_sum_f32", lambda: ops.layer_reduce(f32, "sum"), 4 * n)
    run("diffuse_f32_torus", lambda: ops.layer_diffuse(f32, 0.3, True, out=f32b), 8 * n)
    run("neighborhood_sum_f32_moore_r1_torus", lambda: ops.neighborhood_sum(f32, 1, True, False, True, out=f64), 12 * n)
    del f32, f32b
    # float64 is the default dtype of Mesa's property layers (grid.py:130-147: dtype=float)
    d64 = torch.rand((grid, grid), device=dev, generator=g, dtype=torch.float64)
    run("reduce_sum_f64", lambda: ops.layer_reduce(d64, "sum"), 8 * n)
    run("diffuse_f64_torus", lambda: ops.layer_diffuse(d64, 0.3, True, out=f64), 16 * n)
    run("neighborhood_sum_f64_moore_r1_torus", lambda: ops.neighborhood_sum(d64, 1, True, False, True, out=f64), 16 * n)
    return out


def agent_workloads(dev, n=10_000_000, reps=20):
    """Dynamic populations (csrc/agents.cu): removing 10 % of 10^7 agents -- the compaction plan (flags + exclusive scan)
    and the stable row scatter of a 16-byte column (position + direction of a boid).  Algorithmic bytes of the scatter:
    16 B read per agent + 16 B written per survivor + 5 B of flag / offset per agent.  A failure here is reported in the
    entry instead of ending the bench line."""
    import torch

    from mesa_b200 import ops

    peak, _ = _peaks()
    try:
        g = torch.Generator(device=dev).manual_seed(1)
        keep = torch.rand(n, device=dev, generator=g) < 0.9
        col = torch.zeros((n, 4), dtype=torch.float32, device=dev)

        def run(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_plan = run(lambda: ops.CompactionPlan(keep))
        plan = ops.CompactionPlan(keep)
        alg = 16 * n + 16 * plan.kept + 5 * n
        ms_rows = run(lambda: plan.apply(col))
        return [{"workload": f"remove_agents_plan_{n} (keep flags + exclusive scan)", "ms_per_step": ms_plan,
                 "agents_per_s": n / ms_plan * 1e3, "note": "three small launches: latency-bound"},
                {"workload": f"remove_agents_compact_rows_f32x4_{n} (90 % kept)", "ms_per_step": ms_rows,
                 "agents_per_s": n / ms_rows * 1e3, "algorithmic_bytes": alg, "hbm_gbs": alg / ms_rows / 1e6,
                 "hbm_frac_of_measured_peak": alg / ms_rows / 1e6 / peak}]
    except Exception as exc:  # pragma: no cover - reported, not fatal
        return [{"workload": f"remove_agents_{n}", "error": repr(exc)}]


def _ncu_traffic(kernel="gol_step_u8_vec"):
    """dram bytes per launch of the GoL kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short Schelling / Boltzmann / Boids measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
