"""Time the dynamic-population kernels (csrc/agents.cu) on one GPU: the compaction plan (flags -> exclusive scan) and
the row scatter for the column widths of the example models; one JSON object per line.

    python tools/dev_agents.py [n_agents]

Algorithmic bytes: plan = 1 B flag read + 4 B offset written per agent; rows = row_bytes read per agent (a dropped
row still shares its sectors with kept neighbours) + row_bytes written per survivor + 1 B flag + 4 B offset per agent."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from mesa_b200 import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda:0")
    peak, _ = bench._peaks()
    g = torch.Generator(device=dev).manual_seed(1)
    for frac in (0.9, 0.5):
        keep = torch.rand(n, device=dev, generator=g) < frac
        ms = timed(lambda: ops.CompactionPlan(keep))
        print(json.dumps({"workload": f"compact_plan_{n}", "keep_fraction": frac, "ms": ms, "algorithmic_bytes": 5 * n,
                          "hbm_gbs": 5 * n / ms / 1e6, "hbm_frac_of_measured_peak": 5 * n / ms / 1e6 / peak}), flush=True)
        plan = ops.CompactionPlan(keep)
        kept = plan.kept
        for name, dtype, trailing in [("u8", torch.uint8, ()), ("i32", torch.int32, ()), ("f32x2", torch.float32, (2,)),
                                      ("f32x4", torch.float32, (4,)), ("f64x4", torch.float64, (4,))]:
            col = torch.zeros((n, *trailing), dtype=dtype, device=dev)
            rb = col.element_size() * (col[0].numel() if trailing else 1)
            ms = timed(lambda: plan.apply(col))
            alg = rb * n + rb * kept + 5 * n
            print(json.dumps({"workload": f"compact_rows_{name}_{n}", "row_bytes": rb, "keep_fraction": frac, "ms": ms,
                              "agents_per_s": n / ms * 1e3, "algorithmic_bytes": alg, "hbm_gbs": alg / ms / 1e6,
                              "hbm_frac_of_measured_peak": alg / ms / 1e6 / peak}), flush=True)
            del col
