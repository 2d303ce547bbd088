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
