"""GPU parity: property-layer kernels and neighbourhood sums vs numpy / the reference's neighbourhoods."""
import os

import numpy as np
import pytest
import torch

from oracle import layers as olayers

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NP = {torch.uint8: np.uint8, torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64}


def test_neighborhood_sum_matches_reference_neighbourhoods():
    from mesa_b200 import ops

    g = np.load(os.path.join(GOLDEN, "neighborhoods.npz"))
    for k, (moore, torus, rows, cols, radius, inc) in enumerate(g["cases"]):
        layer = torch.from_numpy(g[f"layer_{k}"]).cuda()
        got = ops.neighborhood_sum(layer, int(radius), bool(moore), bool(inc), bool(torus))
        assert np.array_equal(got.cpu().numpy(), g[f"sums_{k}"]), k
        ones = ops.neighborhood_sum(torch.ones_like(layer).to(torch.uint8), int(radius), bool(moore), bool(inc), bool(torus))
        assert np.array_equal(ones.cpu().numpy(), g[f"sizes_{k}"]), k


@pytest.mark.parametrize("shape", [(64, 64), (33, 77), (257, 1030)])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.float32, torch.float64])
def test_neighborhood_sum_large_matches_restatement(shape, dtype):
    from mesa_b200 import ops

    rng = np.random.default_rng(0)
    a = rng.integers(0, 4, size=shape).astype(NP[dtype])
    for radius, moore, torus in [(1, True, True), (2, False, False), (4, True, False), (3, False, True)]:
        got = ops.neighborhood_sum(torch.from_numpy(a).cuda(), radius, moore, False, torus).cpu().numpy()
        assert np.array_equal(got, olayers.neighborhood_sum(a, radius, moore, False, torus))


def test_neighborhood_radius_zero_is_value_error():
    from mesa_b200 import ops

    with pytest.raises(ValueError):
        ops.neighborhood_sum(torch.zeros((4, 4), dtype=torch.uint8, device="cuda"), radius=0)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 15, 16, 1000, 100_003])
def test_elementwise_and_reduce(dtype, n):
    from mesa_b200 import ops

    rng = np.random.default_rng(n)
    a = rng.integers(0, 20, size=n).astype(NP[dtype])
    b = rng.integers(1, 9, size=n).astype(NP[dtype])
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    # sugarscape regrowth: np.minimum(sugar + 1, cap)   (sugarscape_g1mt/model.py:123-124)
    grown = ops.layer_binary(ops.layer_affine_clamp(ta, 1.0, 1.0, out=torch.empty_like(ta)), tb, "min")
    assert np.array_equal(grown.cpu().numpy(), np.minimum(a + 1, b))
    for op, f in [("add", np.add), ("sub", None), ("mul", np.multiply), ("max", np.maximum)]:
        if op == "sub":
            if dtype == torch.uint8:
                continue
            want = a - b
        else:
            want = f(a, b)
        assert np.array_equal(ops.layer_binary(ta, tb, op).cpu().numpy(), want.astype(NP[dtype]))
    assert np.array_equal(ops.layer_affine_clamp(ta, 2.0, -3.0, 0.0, 10.0, out=torch.empty_like(ta)).cpu().numpy(),
                          np.clip(2.0 * a.astype(np.float64) - 3.0, 0.0, 10.0).astype(NP[dtype]))
    assert int(ops.layer_reduce(ta, "count_nonzero").item()) == int(np.count_nonzero(a))
    assert ops.layer_reduce(ta, "sum").item() == a.astype(np.float64).sum()
    assert ops.layer_reduce(ta, "min").item() == a.min()
    assert ops.layer_reduce(ta, "max").item() == a.max()
    filled = ops.layer_fill(torch.empty_like(ta), 7)
    assert np.array_equal(filled.cpu().numpy(), np.full(n, 7, dtype=NP[dtype]))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.float64, 1e-14)])
@pytest.mark.parametrize("torus", [True, False])
def test_diffuse_matches_restatement_and_conserves_mass(dtype, tol, torus):
    from mesa_b200 import ops

    rng = np.random.default_rng(0)
    a = rng.random((37, 53)).astype(NP[dtype])
    got = ops.layer_diffuse(torch.from_numpy(a).cuda(), 0.3, torus).cpu().numpy()
    want = olayers.diffuse(a, 0.3, torus)
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol)   # float tolerance: accumulation in float64, one rounding
    assert abs(float(got.astype(np.float64).sum()) - float(a.astype(np.float64).sum())) < 1e-3


# ---- the column-walker fast paths (radius 1 / 2) and the vectorised elementwise / reduce paths -------------
@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (40, 1), (3, 3), (5, 8), (16, 128), (17, 132), (50, 515), (130, 1028)])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.float32, torch.float64])
def test_neighborhood_sum_walker_all_modes(shape, dtype):
    """Every (radius <= 2, Moore / von Neumann, torus / clamped, centre in / out) combination on ragged shapes:
    uint8 with the full value range (sums need > 8 bits), float64 with random reals -- the kernel adds in the
    oracle's (di, dj) order, so even float64 sums are bit-identical."""
    from mesa_b200 import ops

    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    if dtype == torch.uint8:
        a = rng.integers(0, 256, size=shape).astype(np.uint8)
    elif dtype == torch.int32:
        a = rng.integers(-1000, 1000, size=shape).astype(np.int32)
    else:
        a = rng.standard_normal(shape).astype(NP[dtype])
    t = torch.from_numpy(a).cuda()
    rows, cols = shape
    for radius in (1, 2):
        for moore in (True, False):
            for torus in (False, True):
                if torus and (rows < 2 * radius + 1 or cols < 2 * radius + 1):
                    continue
                for inc in (False, True):
                    got = ops.neighborhood_sum(t, radius, moore, inc, torus).cpu().numpy()
                    want = olayers.neighborhood_sum(a, radius, moore, inc, torus)
                    assert np.array_equal(got, want), (radius, moore, torus, inc)


def test_neighborhood_sum_row_block_with_halos_equals_whole_layer():
    """A row block with explicit halo rows (the row-sharded layout) gives the rows of the whole-layer result."""
    from mesa_b200 import ops

    rng = np.random.default_rng(5)
    a = rng.integers(0, 200, size=(70, 96)).astype(np.uint8)
    t = torch.from_numpy(a).cuda()
    for radius in (1, 2, 3):
        whole = ops.neighborhood_sum(t, radius, True, False, True)
        lo, hi = 20, 53
        top = t[lo - radius:lo].contiguous()
        bot = t[hi:hi + radius].contiguous()
        part = ops.neighborhood_sum(t[lo:hi].contiguous(), radius, True, False, True, halo_top=top, halo_bot=bot)
        assert torch.equal(part, whole[lo:hi])
        out = torch.empty_like(whole)
        assert ops.neighborhood_sum(t, radius, True, False, True, out=out) is out and torch.equal(out, whole)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int32, torch.float32, torch.float64])
@pytest.mark.parametrize("offset", [0, 1, 3])
def test_elementwise_and_reduce_unaligned_and_large(dtype, offset):
    """Views that do not start on a 16-byte boundary, sizes spanning many vectors per thread."""
    from mesa_b200 import ops

    n = 1_300_007
    rng = np.random.default_rng(offset)
    hi = 256 if dtype == torch.uint8 else 1000
    a = rng.integers(0, hi, size=n + offset).astype(NP[dtype])
    b = rng.integers(0, hi, size=n + offset).astype(NP[dtype])
    ta, tb = torch.from_numpy(a).cuda()[offset:], torch.from_numpy(b).cuda()[offset:]
    a, b = a[offset:], b[offset:]
    assert np.array_equal(ops.layer_binary(ta, tb, "max").cpu().numpy(), np.maximum(a, b))
    assert np.array_equal(ops.layer_binary(ta, tb, "min").cpu().numpy(), np.minimum(a, b))
    assert np.array_equal(ops.layer_binary(ta, tb, "add").cpu().numpy(), (a + b).astype(NP[dtype]))
    got = ops.layer_affine_clamp(ta, 0.5, 1.25, 2.0, 200.0, out=torch.empty_like(ta)).cpu().numpy()
    assert np.array_equal(got, np.clip(0.5 * a.astype(np.float64) + 1.25, 2.0, 200.0).astype(NP[dtype]))
    assert int(ops.layer_reduce(ta, "count_nonzero").item()) == int(np.count_nonzero(a))
    assert ops.layer_reduce(ta, "sum").item() == a.astype(np.float64).sum()
    assert ops.layer_reduce(ta, "min").item() == a.min()
    assert ops.layer_reduce(ta, "max").item() == a.max()
    buf = torch.zeros(n + offset + 5, dtype=dtype, device="cuda")
    ops.layer_fill(buf[offset:offset + n], 9)
    want = np.zeros(n + offset + 5, dtype=NP[dtype])
    want[offset:offset + n] = 9
    assert np.array_equal(buf.cpu().numpy(), want)   # nothing written outside the view


def test_reduce_min_max_of_sparse_uint8_layer():
    """min / max must see a single outlier byte anywhere in a 16-byte vector."""
    from mesa_b200 import ops

    for pos in (0, 1, 7, 15, 16, 4095, 65537):
        a = np.full(70_001, 100, dtype=np.uint8)
        a[pos] = 3
        assert ops.layer_reduce(torch.from_numpy(a).cuda(), "min").item() == 3
        a[pos] = 251
        assert ops.layer_reduce(torch.from_numpy(a).cuda(), "max").item() == 251
        assert ops.layer_reduce(torch.from_numpy(a).cuda(), "min").item() == 100


@pytest.mark.parametrize("shape", [(1, 5), (5, 1), (3, 3), (16, 128), (33, 130), (100, 257)])
@pytest.mark.parametrize("torus", [True, False])
def test_diffuse_walker_ragged_shapes(shape, torus):
    from mesa_b200 import ops

    rows, cols = shape
    if torus and (rows < 3 or cols < 3):
        return
    rng = np.random.default_rng(1)
    a = rng.random(shape)
    got = ops.layer_diffuse(torch.from_numpy(a).cuda(), 0.7, torus).cpu().numpy()
    np.testing.assert_allclose(got, olayers.diffuse(a, 0.7, torus), rtol=1e-14, atol=1e-14)



def test_nd_grids_match_reference_neighbourhoods():
    """1-, 3- and 4-dimensional device grids (grid.py:457-462, 488-501): whole-grid neighbourhood sums / sizes and the
    ORDERED neighbour lists of probe cells against the reference's Cell.get_neighborhood."""
    import random

    from mesa_b200.space import DeviceMooreGrid, DeviceVonNeumannGrid

    g = np.load(os.path.join(GOLDEN, "neighborhoods_nd.npz"))
    for k, (moore, torus, radius, inc) in enumerate(g["cases"]):
        dims = tuple(int(d) for d in g[f"dims_{k}"])
        grid = (DeviceMooreGrid if moore else DeviceVonNeumannGrid)(dims, torus=bool(torus), random=random.Random(1))
        grid.add_property_layer("v", g[f"layer_{k}"])
        got = grid.neighborhood_sum("v", int(radius), bool(inc))
        assert tuple(got.shape) == dims and np.array_equal(got.cpu().numpy(), g[f"sums_{k}"]), (k, dims)
        ones = grid.neighborhood_sum(torch.ones(dims, dtype=torch.uint8, device="cuda"), int(radius), bool(inc))
        assert np.array_equal(ones.cpu().numpy(), g[f"sizes_{k}"]), (k, dims)
        probes = [tuple(0 for _ in dims), tuple(d - 1 for d in dims), tuple(d // 2 for d in dims)]
        for q, pc in enumerate(probes):
            want = [tuple(int(x) for x in row) for row in g[f"probe_{k}_{q}"]]
            assert grid.neighborhood_coordinates(pc, int(radius), bool(inc)) == want, (k, dims, pc)
            mask = grid.get_neighborhood_mask(pc, include_center=bool(inc), radius=int(radius))
            assert int(mask.sum().item()) == len(want) and all(bool(mask[c].item()) for c in want[:5])
    # float layers and the cell API on a 3-d grid
    grid = DeviceMooreGrid((4, 5, 6), torus=True, random=random.Random(2))
    a = np.random.default_rng(0).standard_normal((4, 5, 6))
    got = grid.neighborhood_sum(torch.from_numpy(a).cuda(), 1).cpu().numpy()
    assert np.array_equal(got, olayers.neighborhood_sum_nd(a, 1, True, False, True))
    assert len(grid) == 120 and grid[(3, 4, 5)].flat == 119 and grid.find_nearest_cell([4.5, 0.2, 7.9]).coordinate == (0, 0, 1)
    assert len(list(grid.all_cells)) == 120 and len(grid.empties) == 120
    assert grid.select_random_empty_cell().coordinate in {c.coordinate for c in grid.all_cells}
    with pytest.raises(Exception):
        grid[(1, 2)]
    with pytest.raises(ValueError):
        DeviceMooreGrid((4, 2, 6), torus=True, random=random.Random(2)).neighborhood_sum("empty", 1)
