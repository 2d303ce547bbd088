"""Golden digests for bench.py's multi-GPU parity check (tests/golden/bench_parity.json).

Test infrastructure: runs the sequential CPU oracle (oracle/mesa_oracle.c, itself pinned to the live reference by
tests/test_oracle_golden.py and tests/test_live_reference.py) on the exact inputs bench.py's `parity_checks` builds,
and records per-step digests.  The row-sharded device path must reproduce them for EVERY GPU count ("parity
at 1024^2 on 1/2/4/8 GPUs must be identical across GPU counts and equal to the restatement").

    python oracle/make_bench_parity.py          # rewrites tests/golden/bench_parity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def digest_u64(values: np.ndarray) -> int:
    """numpy twin of bench.py::_digest_u64 (whole grid, first index 0)."""
    v = (values.reshape(-1).astype(np.int64) + 2).astype(np.uint64)
    idx = np.arange(v.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        w = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(2135587861)) | np.uint64(1)
        return int((v * w).sum(dtype=np.uint64))


def schelling_case(rows=1024, seed=11, steps=5, density=0.8, homophily=0.4, radius=1):
    rng = np.random.Generator(np.random.Philox(seed))
    u, v = rng.random((rows, rows)), rng.random((rows, rows))
    grid = np.where(u < density, (v < 0.5).astype(np.int8), np.int8(-1)).astype(np.int8)
    st = oracle.SchellingState.from_type_grid(grid)
    out = {"rows": rows, "seed": seed, "happy0": st.assign_state(homophily, radius), "steps": []}
    for s in range(steps):
        movers, nfb = st.move_unhappy(seed, s)
        happy = st.assign_state(homophily, radius)
        out["steps"].append({"movers": movers, "happy": happy, "types": digest_u64(st.type_grid()),
                             "ids": digest_u64(st.id_grid())})
    return out


def main():
    data = {"schelling_1024": schelling_case()}
    path = os.path.join(ROOT, "tests", "golden", "bench_parity.json")
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(path, json.dumps(data)[:400])


if __name__ == "__main__":
    main()
