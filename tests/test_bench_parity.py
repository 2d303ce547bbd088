"""bench.py's multi-GPU parity digests: the torch digest (what the ranks compute) equals the numpy one (what the golden
file was made with), adds up over row blocks, and the committed golden file is what the oracle produces today."""
import importlib.util
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_digest_torch_equals_numpy_and_is_block_additive():
    bench = _load("bench_mod", os.path.join(ROOT, "bench.py"))
    mk = _load("make_bench_parity", os.path.join(ROOT, "oracle", "make_bench_parity.py"))
    rng = np.random.default_rng(0)
    grid = rng.integers(-1, 2**31, size=(37, 53)).astype(np.int64)
    want = mk.digest_u64(grid)
    whole = int(bench._digest_u64(torch.from_numpy(grid), 0).item()) & (2**64 - 1)
    assert whole == want
    parts = 0
    for lo, hi in [(0, 5), (5, 20), (20, 37)]:
        parts += int(bench._digest_u64(torch.from_numpy(grid[lo:hi]), lo * 53).item())
    assert parts & (2**64 - 1) == want


def test_golden_digests_are_the_oracles():
    mk = _load("make_bench_parity", os.path.join(ROOT, "oracle", "make_bench_parity.py"))
    with open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")) as f:
        gold = json.load(f)
    assert mk.schelling_case() == gold["schelling_1024"]


def test_launch_plan():
    from mesa_b200.ops import gol_launch_plan as plan

    assert plan(20, 8) == [8, 8, 4]
    assert plan(20, 8, even=True) == [8, 4, 4, 4]        # the LONGEST launch is split (two launches of 4 beat 2 + 2)
    assert plan(20, 8, 0, 64, even=True) == [0, 8, 4, 4, 4]
    p = plan(1000, 8, 5, 64)
    assert sum(p) == 1000 and p[:3] == [5, 0, 8] and p.count(0) == 16
    for k in (2, 7, 16, 17, 63, 64, 65, 129):
        q = plan(k, 8, 0, 64, even=True)
        assert sum(q) == k and sum(1 for t in q if t) % 2 == 0 and max(q) <= 8
