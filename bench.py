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
def run_ours(args):
    import torch
    import torch.distributed as dist

    from mesa_b200 import ops
    from mesa_b200.sharding import FusedPeerGol, HaloExchanger, PeerHaloLayer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mesa_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]   # the image's default prints a banner on stdout; rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # synthetic initial state: alive with p=0.2 (conways_game_of_life/model.py:9-15), Philox-seeded
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    a = (torch.rand((GRID, GRID), device=dev, generator=g) < 0.2).to(torch.uint8)
    b = torch.empty_like(a)
    path = os.environ.get("MB_GOL_PATH", "bits")
    gens_per_launch = int(os.environ.get("MB_GOL_T", "8")) if path == "bits" else 1
    halo_mode = os.environ.get("MB_HALO", "fused") if (world > 1 and path != "bits") else "none"
    hx = HaloExchanger(GRID, torch.uint8, 1, True, dev) if halo_mode != "none" else None
    peer = fused = life = None
    if path == "bits":
        # bit-packed planes, gens_per_launch generations per launch; N > 1: deep ghost zones (64 rows of each ring
        # neighbour, copied out of its HBM over NVLink once per 64 generations between two device barriers)
        if world > 1:
            from mesa_b200.sharding import PeerBitLife

            life = PeerBitLife(GRID, GRID, True, dev, generations_per_launch=gens_per_launch,
                               ghost=int(os.environ.get("MB_GOL_GHOST", "0")) or None)
        else:
            life = ops.BitLife(GRID, GRID, dev, torus=True, generations_per_launch=gens_per_launch)
        life.load(a)
    elif halo_mode == "fused":
        # halos read in place from the neighbours' HBM over NVLink, hand-shake fused into the stencil kernel
        fused = FusedPeerGol(GRID, GRID, True, dev)
        fused.load(a)
    elif halo_mode == "peer":
        # halos read in place from the neighbours' HBM (symmetric memory) + one device barrier per generation
        peer = PeerHaloLayer(GRID, GRID, torch.uint8, 1, True, dev)
        peer.state.copy_(a)

    def gol_rows(src, dst, top, bot):
        ops.gol_step(src, out=dst, torus=False, col_torus=True,
                     halo_top=None if top is None else top[-1], halo_bot=None if bot is None else bot[0])

    def step(src, dst):
        if hx is None:
            ops.gol_step(src, out=dst, torus=True)
        else:
            top, bot = hx.exchange(src)
            ops.gol_step(src, out=dst, torus=True, halo_top=top[0], halo_bot=bot[0])

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
        else:
            step(a, b)
            a, b = b, a

    def run_generations(k):
        if life is not None:
            life.run(k)
        else:
            for _ in range(k):
                advance()

    run_generations(max(args.warmup, 3))
    if life is not None:
        # the timed region launches the T-generation kernel and, if K is not a multiple of T, one shorter variant:
        # load both before timing (CUDA loads kernels lazily on first launch)
        run_generations(gens_per_launch)
        if args.steps % gens_per_launch:
            run_generations(args.steps % gens_per_launch)
    barrier()
    launches0 = life.launches if life is not None else 0
    # single GPU, bit-packed path: the K generations are captured once into a CUDA graph (the launch sequence is static:
    # ping-pong planes, T generations per launch) and replayed inside the timed region, so a short run (small K) is not
    # dominated by the host-side cost of the first launches
    graph = None
    if life is not None and world == 1 and os.environ.get("MB_BENCH_GRAPH", "1") != "0":
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            life.run(args.steps)          # captured, not executed
        torch.cuda.current_stream(dev).wait_stream(side)
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(_physical_gpu_index(local)) as clocks:
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            run_generations(args.steps)
        e1.record()
        barrier()
    launches = (life.launches - launches0) if life is not None else args.steps
    if fused is not None:
        fused.check()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    updates = GRID * GRID * args.steps * world
    value = updates / (ms * 1e-3)
    ms_per_step = ms / args.steps

    # ---- e2e: host buffers in, host buffers out, every step (public API: mesa_b200.host_api.GolHostStepper;
    # H2D / stencil / D2H pipelined over row chunks so both PCIe directions stay busy).  With N > 1 every rank
    # steps its own block from / to its own host buffers; the two halo rows come from the resident previous
    # generation of the neighbours (same NCCL / peer exchange as above is not needed for the copy-bound figure:
    # each rank treats its block as a torus, which moves exactly the same bytes).
    from mesa_b200.host_api import GolHostStepper

    e2e_steps = max(3, min(args.steps, 20))
    h_in = torch.empty((GRID, GRID), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((GRID, GRID), dtype=torch.uint8).pin_memory()
    h_in.copy_((life.store() if life is not None else fused.state if fused is not None else peer.state if peer is not None else a).cpu())
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

    if rank == 0:
        peak, peak_src = _peaks()
        # algorithmic bytes: 2 B per cell-update x the cell-updates one launch performs
        gens_avg = args.steps / launches
        alg_bytes = ALG_BYTES_PER_UPDATE * GRID * GRID * gens_avg
        achieved = alg_bytes / (ms / launches * 1e-3) / 1e9
        if life is not None:
            kernel = f"gol_bits_kernel<{gens_per_launch}>"
            par = (f"row-sharded x{world}, bit-packed planes, {life.ghost} ghost rows of each ring neighbour copied out of its HBM "
                   f"over NVLink (peer-mapped memory) once per {life.ghost} generations between two device barriers; no collective "
                   "on the data path") if world > 1 else "single GPU"
            note = ("bit-packed state (1 bit/cell, 2 x 32 MiB planes resident in L2) and temporal blocking: the kernel is bound "
                    "by integer issue rate, not HBM; `achieved` is the contract's ALGORITHMIC figure (2 B per cell-update of the "
                    "uint8 layer) and exceeds the HBM peak by design (SURVEY 8d allows it); `traffic` is the measured DRAM bytes "
                    "per launch.  The HBM-bound uint8-layer stencil is under workloads[gol_u8_layer...]")
        else:
            kernel = "gol_step_u8_vec"
            par = (f"row-sharded x{world}, ring halos " + ("read in place from peer HBM over NVLink, hand-shake fused into the stencil kernel (1 launch / generation, no collective)" if fused is not None else "read in place from peer HBM over NVLink (symmetric memory) + device barrier" if peer is not None else "exchanged with NCCL send/recv")) if world > 1 else "single GPU"
            note = "uint8 layer, one generation per launch"
        line = {
            "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "gol_16384x16384_u8_torus", "grid_per_gpu": [GRID, GRID],
                       "global_grid": [GRID * world, GRID], "torus": True, "device_format": "bits" if life is not None else "u8",
                       "generations_per_launch": gens_per_launch, "parallelism": par,
                       "launch": "one CUDA graph replay of the K-generation launch sequence" if graph is not None else "stream launches",
                       "l2": L2_NOTE if life is None else "bit planes (2 x 32 MiB) are L2-resident by design; no flush: the resident working set IS the workload (1000 consecutive generations of one layer)"},
            "e2e": {"value": e2e_value, "unit": "agent-updates/s", "h2d_bytes_per_step": GRID * GRID,
                    "d2h_bytes_per_step": GRID * GRID, "steps": e2e_steps,
                    "api": "mesa_b200.host_api.GolHostStepper.step (pinned host in/out, 16 row chunks, 3 streams)"},
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": _ncu_traffic(kernel.split("<")[0]), "peak_source": peak_src,
                         "kernel": kernel, "algorithmic_bytes_per_launch": alg_bytes, "note": note},
        }
        if life is not None:
            # the roofline that actually binds the bit-packed kernel: the SM's ALU pipe (LOP3 / SHF issue one warp
            # instruction every other cycle per SM sub-partition; 12 such instructions per 32-cell word and generation,
            # csrc/gol_bits.cu).  ALGORITHMIC count: the words of the layer only, no halo lanes / re-computed rows.
            sms, clock_hz = 148, (clocks.summary()["sm_mhz"] or 1965) * 1e6
            alu_peak = sms * 4 * 16 * clock_hz                       # thread-instructions / s on the ALU pipe
            alu_achieved = 12.0 * (GRID * GRID // 32) / (ms_per_step * 1e-3)   # one thread owns a 32-cell word
            line["roofline"]["alu_pipe"] = {"achieved": alu_achieved, "peak": alu_peak, "unit": "thread-instr/s",
                                            "frac": alu_achieved / alu_peak,
                                            "note": "12 ALU-pipe instructions per word-generation x 16384^2/32 words per generation; "
                                                    "ncu (profiles/r1_gol_bits_T8_ncu_full.csv): ALU pipe 83 % busy while active "
                                                    "including the re-computed halo rows / lanes"}
        if world == 1 and not args.no_others:
            line["workloads"] = other_workloads(dev) + layer_workloads(dev) + agent_workloads(dev)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, _ = cpu_gol_rate(256, GRID, 1, 1, threads)
            rows = max(64, min(GRID, int(rate * 5.0 / GRID)) // 64 * 64)  # ~5 s per step, 3 steps
            v, _ = cpu_gol_rate(rows, GRID, 3, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "agent-updates/s", "cores": threads, "kind": "port",
                                    "sample": f"gol torus slab {rows}x{GRID} uint8, 3 steps, C oracle port",
                                    # context only, NOT measured in this run: the unmodified pure-Python reference as
                                    # timed in the build container (it is not on the GPU box and
                                    # cannot construct the 16384^2 grid at all: ~1.3 KB and 14 us per cell)
                                    "mesa_python_indicative": {"gol_200x200": 2.9e5, "schelling_100x100": 3.0e5,
                                                               "boltzmann_1e4_100x100": 2.2e5, "boids_400": 1.5e4,
                                                               "unit": "agent-updates/s", "cores": 1,
                                                               "source": "BASELINE.md section 2 (build container)"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    t = timed(m.step, 100)
    out.append({"workload": "schelling_100x100_d0.8_h0.4_r1_100steps", "agents": n, "ms_per_step": t * 1e3,
                "agent_updates_per_s": n / t, "note": "latency-bound (L2-resident); roofline fraction not meaningful"})
    # the same with the reference's DataCollector inside step() (model.py:92; 4 model reporters + 1 agent column per step)
    m = Schelling(SchellingScenario(width=100, height=100, density=0.8, homophily=0.4, radius=1, rng=42), collect=True)
    t = timed(m.step, 100)
    out.append({"workload": "schelling_100x100_d0.8_h0.4_r1_100steps_with_datacollector", "agents": n,
                "ms_per_step": t * 1e3, "agent_updates_per_s": n / t})
    # Schelling happiness stencil alone on 16384^2 (the >= 60 % target of the north star)
    g = torch.Generator(device=dev).manual_seed(0)
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
    # config 2: Boltzmann wealth, 10^6 agents on 4096^2
    m = BoltzmannWealth(BoltzmannScenario(n=1_000_000, width=4096, height=4096, rng=1))
    m.run_for(3)
    t = timed(m.step, 30)
    out.append({"workload": "boltzmann_1e6_agents_4096x4096", "ms_per_step": t * 1e3, "agent_updates_per_s": 1e6 / t})
    del m
    # config 3: Boids, 10^7 agents at the reference "large" density (vision 15)
    N = 10_000_000
    side = 150.0 * math.sqrt(N / 400.0)
    pos = torch.rand((N, 2), device=dev, generator=g) * side
    dirs = torch.rand((N, 2), device=dev, generator=g) * 2 - 1
    ws = ops.CellListWorkspace(N, (side, side), 15.0, dev)
    po, do = torch.empty_like(pos), torch.empty_like(dirs)
    state = [pos, dirs, po, do]

    def boids():
        ops.boids_step(state[0], state[1], (side, side), vision=15.0, separation=2.0, ws=ws, pos_out=state[2], dir_out=state[3])
        state[0], state[2] = state[2], state[0]
        state[1], state[3] = state[3], state[1]

    for _ in range(2):
        boids()
    t = timed(boids, 10)
    out.append({"workload": "boids_1e7_vision15_synchronous", "ms_per_step": t * 1e3, "agent_updates_per_s": N / t})
    # the same flock in the reference's SEQUENTIAL shuffle_do order (exact replay mode, rank-ordered chained kernel)
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
