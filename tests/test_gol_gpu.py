"""GPU parity: Game of Life kernel vs the CPU oracle (bit-exact)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _run(state_np, torus, steps=1):
    from mesa_b200 import ops

    a = torch.from_numpy(state_np).cuda()
    b = torch.empty_like(a)
    for _ in range(steps):
        ops.gol_step(a, out=b, torus=torus)
        a, b = b, a
    return a.cpu().numpy()


@pytest.mark.parametrize("shape", [(3, 3), (5, 7), (37, 23), (64, 64), (100, 100), (65, 512), (130, 528),
                                   (256, 1024), (67, 1040), (1, 16), (2, 32), (64, 16)])
@pytest.mark.parametrize("torus", [True, False])
def test_gol_matches_oracle(shape, torus):
    if torus and min(shape) < 3:
        pytest.skip("torus needs dims >= 3")
    rng = np.random.default_rng(hash(shape) % 2**32)
    s = (rng.random(shape) < 0.35).astype(np.uint8)
    want = s
    for _ in range(3):
        want = oracle.gol_step(want, torus)
    got = _run(s, torus, steps=3)
    assert np.array_equal(got, want)


def test_gol_rejects_small_torus():
    from mesa_b200 import ops

    s = torch.zeros((8, 2), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.gol_step(s, torus=True)


def test_gol_empty():
    from mesa_b200 import ops

    s = torch.zeros((0, 16), dtype=torch.uint8, device="cuda")
    assert ops.gol_step(s, torus=False).shape == (0, 16)


@pytest.mark.parametrize("shape,chunks", [((64, 64), 4), ((100, 48), 7), ((37, 23), 3), ((8, 16), 16)])
def test_gol_host_stepper_matches_oracle(shape, chunks):
    from mesa_b200.host_api import GolHostStepper, gol_step_host

    rng = np.random.default_rng(5)
    s = (rng.random(shape) < 0.4).astype(np.uint8)
    h_in = torch.from_numpy(s).pin_memory()
    stepper = GolHostStepper(shape[0], shape[1], chunks=chunks)
    want = s
    for _ in range(3):
        h_out = gol_step_host(h_in, stepper=stepper)
        want = oracle.gol_step(want, True)
        assert np.array_equal(h_out.numpy(), want)
        h_in = h_out.clone().pin_memory()
