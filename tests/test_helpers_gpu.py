"""Device helpers that replaced torch library calls on product paths: stable argsort of a column
(radix sort of csrc/sort.cu on order-preserving key bits), ascending indices of the set entries of a mask (stable
compaction of csrc/agents.cu), CSR offsets from counts (mb_offsets_from_counts)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64, torch.float32, torch.float64, torch.uint8, torch.bool])
@pytest.mark.parametrize("n", [0, 1, 33, 5000, 70001])
@pytest.mark.parametrize("descending", [False, True])
def test_argsort_stable_matches_python_sorted(dtype, n, descending):
    from mesa_b200 import ops

    rng = np.random.default_rng(n + 7)
    if dtype in (torch.float32, torch.float64):
        x = rng.integers(-20, 20, size=n).astype(np.float64) / 4.0           # many ties, both signs, +-0
        if n > 4:
            x[:4] = [0.0, -0.0, np.inf, -np.inf]
        t = torch.from_numpy(x).to(dtype)
    elif dtype == torch.bool:
        t = torch.from_numpy(rng.integers(0, 2, size=n).astype(bool))
    elif dtype == torch.uint8:
        t = torch.from_numpy(rng.integers(0, 256, size=n).astype(np.uint8))
    else:
        lim = 2**62 if dtype == torch.int64 else 2**30
        x = rng.integers(-lim, lim, size=n)
        x[: n // 2] = rng.integers(-3, 3, size=n // 2)
        t = torch.from_numpy(x).to(dtype)
    got = ops.argsort_stable(t.cuda(), descending=descending).cpu().tolist()
    vals = t.tolist()
    want = sorted(range(n), key=lambda i: vals[i], reverse=descending)      # Python's sort is stable, also with reverse
    assert got == want


@pytest.mark.parametrize("shape", [(0,), (1,), (37, 23), (4096,), (300, 301)])
def test_nonzero_indices_are_the_row_major_argwhere(shape):
    from mesa_b200 import ops

    rng = np.random.default_rng(3)
    m = rng.random(shape) < 0.3
    got = ops.nonzero_indices(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, np.flatnonzero(m))
    assert ops.nonzero_indices(torch.zeros(shape, dtype=torch.bool, device="cuda")).numel() == 0


@pytest.mark.parametrize("n", [0, 1, 4095, 4096, 4097, 100003])
def test_offsets_from_counts(n):
    from mesa_b200 import ops

    rng = np.random.default_rng(n)
    c = rng.integers(0, 50, size=n).astype(np.int32)
    got = ops.offsets_from_counts(torch.from_numpy(c).cuda()).cpu().numpy()
    assert got.dtype == np.int64 and np.array_equal(got, np.concatenate([[0], np.cumsum(c, dtype=np.int64)]))
