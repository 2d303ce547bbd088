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


@pytest.mark.parametrize("shape,chunks", [((64, 64), 4), ((100, 48), 7), ((37, 23), 3), ((8, 16), 16), ((1000, 64), 3),
                                          ((521, 48), 5)])
def test_gol_host_stepper_matches_oracle(shape, chunks):
    from mesa_b200.host_api import GolHostStepper, gol_step_host

    rng = np.random.default_rng(5)
    s = (rng.random(shape) < 0.4).astype(np.uint8)
    h_in = torch.from_numpy(s).pin_memory()
    stepper = GolHostStepper(shape[0], shape[1], chunks=chunks, ramp=shape[0] > 500)
    want = s
    for _ in range(3):
        h_out = gol_step_host(h_in, stepper=stepper)
        want = oracle.gol_step(want, True)
        assert np.array_equal(h_out.numpy(), want)
        h_in = h_out.clone().pin_memory()


# ---- bit-packed layer, several generations per launch (csrc/gol_bits.cu) --------------------------------
def _oracle_gens(s, torus, n):
    for _ in range(n):
        s = oracle.gol_step(s, torus)
    return s


@pytest.mark.parametrize("shape", [(3, 32), (5, 64), (40, 96), (70, 1024), (33, 2048), (8, 992), (131, 4096)])
def test_gol_pack_unpack_roundtrip(shape):
    from mesa_b200 import ops

    rng = np.random.default_rng(3)
    s = (rng.random(shape) < 0.5).astype(np.uint8)
    bits = ops.gol_pack(torch.from_numpy(s).cuda())
    want = np.packbits(s.reshape(shape[0], -1, 32), axis=2, bitorder="little").view(np.uint32).reshape(shape[0], -1)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), want)
    assert np.array_equal(ops.gol_unpack(bits).cpu().numpy(), s)


@pytest.mark.parametrize("shape", [(3, 32), (9, 64), (40, 96), (70, 1024), (33, 2048), (8, 992), (131, 4096)])
@pytest.mark.parametrize("torus", [True, False])
@pytest.mark.parametrize("gens", [1, 2, 3, 4, 5, 6, 7, 8])
def test_gol_bits_matches_oracle(shape, torus, gens):
    from mesa_b200 import ops

    if torus and shape[0] < max(3, gens):
        pytest.skip("torus needs rows >= generations")
    rng = np.random.default_rng(hash((shape, gens)) % 2**32)
    s = (rng.random(shape) < 0.4).astype(np.uint8)
    want = _oracle_gens(s, torus, gens)
    bits = ops.gol_pack(torch.from_numpy(s).cuda())
    for rps in (0, 2, 10):
        out = torch.empty_like(bits)
        ops.gol_step_bits(bits, out, gens, torus=torus, rows_per_segment=rps)
        assert np.array_equal(ops.gol_unpack(out).cpu().numpy(), want), (shape, torus, gens, rps)


@pytest.mark.parametrize("torus", [True, False])
def test_gol_bits_row_blocks_with_halos(torus):
    """Two row blocks stepped separately with each other's edge rows as halos == the whole layer."""
    from mesa_b200 import ops

    rows, cols, gens = 48, 160, 4
    rng = np.random.default_rng(11)
    s = (rng.random((rows, cols)) < 0.4).astype(np.uint8)
    want = _oracle_gens(s, torus, gens)
    bits = ops.gol_pack(torch.from_numpy(s).cuda())
    top, bot = bits[:20].contiguous(), bits[20:].contiguous()
    o1, o2 = torch.empty_like(top), torch.empty_like(bot)
    if torus:
        ops.gol_step_bits(top, o1, gens, torus=False, col_torus=True, halo_top=bot[-gens:], halo_bot=bot[:gens])
        ops.gol_step_bits(bot, o2, gens, torus=False, col_torus=True, halo_top=top[-gens:], halo_bot=top[:gens])
    else:
        ops.gol_step_bits(top, o1, gens, torus=False, halo_bot=bot[:gens])
        ops.gol_step_bits(bot, o2, gens, torus=False, halo_top=top[-gens:])
    got = ops.gol_unpack(torch.cat([o1, o2])).cpu().numpy()
    assert np.array_equal(got, want)


def test_bitlife_run_and_errors():
    from mesa_b200 import ops

    rng = np.random.default_rng(2)
    s = (rng.random((64, 256)) < 0.3).astype(np.uint8)
    life = ops.BitLife(64, 256, "cuda", torus=True, generations_per_launch=4)
    life.load(torch.from_numpy(s).cuda())
    life.run(11)
    assert life.launches == 3
    assert np.array_equal(life.store().cpu().numpy(), _oracle_gens(s, True, 11))
    with pytest.raises(ValueError):
        ops.gol_pack(torch.zeros((4, 40), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        b = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
        ops.gol_step_bits(b, torch.empty_like(b), 9)
    with pytest.raises(ValueError):
        b = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
        ops.gol_step_bits(b, b, 1)


def test_gol_model_packed_equals_unpacked():
    from mesa_b200.examples import ConwaysGameOfLife

    rng = np.random.default_rng(4)
    s0 = rng.random((40, 64)) < 0.3
    a = ConwaysGameOfLife(width=40, height=64, initial_state=s0, packed=True)
    b = ConwaysGameOfLife(width=40, height=64, initial_state=s0, packed=False)
    assert a._bits is not None and b._bits is None
    for _ in range(3):
        a.step()
        b.step()
    a.run_for(9)
    b.run_for(9)
    want = _oracle_gens(s0.astype(np.uint8), True, 12)
    assert a.steps == b.steps == 12 and a.time == 12.0  # every step() call advances time by one unit (model.py:152-154)
    assert np.array_equal(a.grid.state.cpu().numpy(), want)
    assert np.array_equal(b.grid.state.cpu().numpy(), want)
    assert np.array_equal(a.agents.get("state").cpu().numpy(), want.reshape(-1))
    # in-place edits of the layer are picked up by the next step
    a.grid.state[:] = 0
    a.grid.state[10, 10:13] = 1  # blinker
    a.step()
    got = a.grid.state.cpu().numpy()
    assert got.sum() == 3 and got[9:12, 11].all()
