"""The row-sharded Schelling driver (mesa_b200/sharded.py) with a 1-rank NCCL group: exercises the
phase-level C ABI + symmetric memory path on the single-GPU CI box.  Multi-GPU parity (2 GPUs, uneven
blocks, argwhere fallback across ranks) is checked by tools/check_sharded_schelling.py under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def one_rank_group():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    yield
    if created:
        dist.destroy_process_group()


@pytest.mark.parametrize("rows,cols,density,h,radius,steps", [(48, 64, 0.8, 0.4, 1, 8), (30, 32, 0.97, 0.9, 1, 4),
                                                              (40, 48, 0.85, 0.6, 2, 5)])
def test_sharded_driver_world1_matches_oracle(one_rank_group, rows, cols, density, h, radius, steps):
    from mesa_b200.sharded import ShardedSchelling

    rng = np.random.default_rng(rows)
    grid = np.where(rng.random((rows, cols)) < density, (rng.random((rows, cols)) < 0.5).astype(np.int8), np.int8(-1)).astype(np.int8)
    m = ShardedSchelling(rows, cols, grid, h, radius, seed=9)
    st = oracle.SchellingState.from_type_grid(grid)
    assert st.assign_state(h, radius) == m.happy_count
    for s in range(steps):
        happy = m.step()
        movers, _ = st.move_unhappy(9, s)
        assert movers == m.last_movers
        assert st.assign_state(h, radius) == happy
        assert np.array_equal(m.type.cpu().numpy(), st.type_grid())
        assert np.array_equal(m.happy.cpu().numpy(), st.happy_grid())


def test_sharded_boids_world1_matches_oracle(one_rank_group):
    """Single-rank run of the domain-decomposed driver (no neighbours): the slab is the whole torus."""
    from mesa_b200.sharded import ShardedBoids

    rng = np.random.default_rng(3)
    n, side = 1500, 200.0
    pos = (rng.random((n, 2)) * side).astype(np.float32)
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    m = ShardedBoids(side, side, pos, dirs, np.arange(n), vision=10.0, separation=2.0)
    for _ in range(3):
        m.step()
        ids, gp, gd = m.gather()
        wp, wd, _ = oracle.boids_step(pos, dirs, (side, side), vision=10.0, separation=2.0)
        gp, gd = gp.cpu().numpy(), gd.cpu().numpy()
        assert np.array_equal(ids.cpu().numpy(), np.arange(n))
        assert np.max(np.abs(gp - wp)) / side <= 1e-5 and np.max(np.abs(gd - wd)) <= 1e-5
        pos, dirs = gp.copy(), gd.copy()


@pytest.mark.parametrize("torus", [True, False])
@pytest.mark.parametrize("ghost,launches,exchanges", [(64, 3, 1), (8, 3, 2), (4, 3, 3)])
@pytest.mark.parametrize("overlap", [True, False])
def test_peer_bit_life_world1_matches_oracle(one_rank_group, torus, ghost, launches, exchanges, overlap):
    """Row-sharded bit-packed Game of Life with a 1-rank group: the ring neighbour is the rank itself (torus) or
    absent (clamped); deep ghost zones refreshed every `ghost` generations must equal the oracle."""
    from mesa_b200.sharding import PeerBitLife

    rng = np.random.default_rng(8)
    s = (rng.random((72, 192)) < 0.35).astype(np.uint8)
    # overlap=True: the launch after a refresh is split into an interior launch + the exchange and two edge strips on a
    # side stream (PeerBitLife._step_with_refresh)
    life = PeerBitLife(72, 192, torus, "cuda", generations_per_launch=4, ghost=ghost, overlap=overlap)
    life.load(torch.from_numpy(s).cuda())
    life.run(10)
    assert life.kernel_launches == launches + (2 * exchanges if overlap else 0)
    want = s
    for _ in range(10):
        want = oracle.gol_step(want, torus)
    assert (life.launches, life.exchanges) == (launches, exchanges)
    assert np.array_equal(life.store().cpu().numpy(), want)


@pytest.mark.parametrize("torus", [True, False])
def test_sharded_layer_world1_matches_oracle(one_rank_group, torus):
    """Row-sharded property layer with a 1-rank group (the ring neighbour is the rank itself on a torus): halo rows go
    through HaloExchanger into halo_top / halo_bot of the walker / box kernels; results equal the oracle's."""
    from mesa_b200.sharding import ShardedLayer
    from oracle import layers as olayers

    rng = np.random.default_rng(11)
    counts = rng.integers(0, 256, size=(70, 96)).astype(np.uint8)
    layer = ShardedLayer(torch.from_numpy(counts).cuda(), torus=torus)
    for radius, moore, inc in [(1, True, False), (2, False, True), (3, True, False)]:
        got = layer.neighborhood_sum(radius, moore, inc).cpu().numpy()
        assert np.array_equal(got, olayers.neighborhood_sum(counts, radius, moore, inc, torus)), (radius, moore, inc)
    assert layer.reduce("sum") == int(counts.sum()) and layer.reduce("max") == int(counts.max())
    assert layer.reduce("count_nonzero") == int(np.count_nonzero(counts))
    heat = rng.random((70, 96)).astype(np.float32)
    field = ShardedLayer(torch.from_numpy(heat).cuda(), torus=torus)
    want = heat
    for _ in range(3):
        field.diffuse(0.4)
        want = olayers.diffuse(want, 0.4, torus)
    np.testing.assert_allclose(field.block.cpu().numpy(), want, rtol=2e-6, atol=2e-6)
    assert abs(field.reduce("sum") - float(heat.astype(np.float64).sum())) < 1e-2
