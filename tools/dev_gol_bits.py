"""Developer sweep of the bit-packed Game of Life kernel (generations per launch x rows per segment) on 16384^2,
checked against the uint8 kernel.  python tools/dev_gol_bits.py [grid]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesa_b200 import ops  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
a = (torch.rand((G, G), device=dev, generator=g) < 0.2).to(torch.uint8)
b = torch.empty_like(a)

# reference: 16 generations with the uint8 kernel
x, y = a.clone(), b
for _ in range(16):
    ops.gol_step(x, out=y, torus=True)
    x, y = y, x
want = x.clone()
del y

bits = ops.gol_pack(a)
results = []
# (the kernel variants of the first sweeps -- prefetch depth, warps per CTA, register caps -- have been removed from the
#  library; `variant` is kept in the records for comparison with profiles/r1_gol_bits_sweep*.jsonl)
SWEEP = [(14, T, rps) for T in (1, 2, 3, 4, 5, 6, 7, 8) for rps in (0, 32, 48, 56, 64, 72, 88, 96, 128)]
if len(sys.argv) > 2 and sys.argv[2] == "fine":
    SWEEP = [(14, T, rps) for T in (5, 6, 7, 8) for rps in (0, 48, 50, 52, 56, 64, 72, 80, 96, 100, 104, 112, 128)]
if len(sys.argv) > 2 and sys.argv[2] == "quick":
    SWEEP = [(14, 4, 0), (14, 8, 0), (14, 8, 64), (14, 8, 72)]
for variant, T, rps in SWEEP:
    if True:
        cur, nxt = bits.clone(), torch.empty_like(bits)
        gens = 0
        while gens < 16:  # correctness: 16 generations in launches of T
            t = min(T, 16 - gens)
            ops.gol_step_bits(cur, nxt, t, torus=True, rows_per_segment=rps)
            cur, nxt = nxt, cur
            gens += t
        ok = bool(torch.equal(ops.gol_unpack(cur), want))
        launches = max(8, 240 // T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.gol_step_bits(cur, nxt, T, torus=True, rows_per_segment=rps)
            cur, nxt = nxt, cur
        e0.record()
        for _ in range(launches):
            ops.gol_step_bits(cur, nxt, T, torus=True, rows_per_segment=rps)
            cur, nxt = nxt, cur
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (launches * T)
        results.append({"variant": variant, "T": T, "rps": rps, "ok": ok, "us_per_generation": us, "updates_per_s": G * G / us * 1e6})
        print(json.dumps(results[-1]), flush=True)
# pack / unpack cost
for name, fn in (("pack", lambda: ops.gol_pack(a, out=bits)), ("unpack", lambda: ops.gol_unpack(bits, out=b))):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({name + "_us": e0.elapsed_time(e1) * 1e3 / 20}), flush=True)
best = min((r for r in results if r["ok"]), key=lambda r: r["us_per_generation"], default=None)
print("BEST", json.dumps(best))
assert all(r["ok"] for r in results), "bit-packed kernel disagrees with the uint8 kernel"
