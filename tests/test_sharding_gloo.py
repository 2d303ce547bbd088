"""Row-block sharding + halo exchange on CPU with the gloo backend, world_size 2 and 3.

The exchange logic (mesa_b200/sharding.py) is host-side and backend-agnostic; the compute on each
block is done here by the CPU oracle so the test runs without a GPU: a sharded run must equal the
unsharded one, torus and clamped, radius 1 and 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gol_block(block, top, bot):
    """One generation of a row block given its halo rows (None = dead), columns wrap (torus)."""
    rows, cols = block.shape
    ext = np.zeros((rows + 2, cols), dtype=np.uint8)
    ext[1:-1] = block
    if top is not None:
        ext[0] = top
    if bot is not None:
        ext[-1] = bot
    n = np.zeros_like(ext, dtype=np.int32)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di or dj:
                n += np.roll(np.roll(ext, di, 0), dj, 1)
    nxt = ((n == 3) | ((ext == 1) & (n == 2))).astype(np.uint8)
    return nxt[1:-1]


def _worker(rank, world, port, torus, steps, rows, cols, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesa_b200.sharding import HaloExchanger, block_bounds

    rng = np.random.default_rng(0)
    full = (rng.random((rows, cols)) < 0.3).astype(np.uint8)
    lo, hi = block_bounds(rows, world, rank)
    block = torch.from_numpy(full[lo:hi].copy())
    hx = HaloExchanger(cols, torch.uint8, 1, torus, "cpu")
    for _ in range(steps):
        top, bot = hx.exchange(block)
        nxt = _gol_block(block.numpy(), None if top is None else top[0].numpy(), None if bot is None else bot[0].numpy())
        block = torch.from_numpy(nxt)
    np.save(os.path.join(out_dir, f"block_{rank}.npy"), block.numpy())
    # the global happy / alive count is a 1-element all-reduce
    total = torch.tensor([int(block.sum())])
    dist.all_reduce(total)
    np.save(os.path.join(out_dir, f"total_{rank}.npy"), total.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,torus", [(2, True), (2, False), (3, True)])
def test_sharded_gol_equals_unsharded(tmp_path, world, torus):
    import oracle

    rows, cols, steps = 31, 16, 4
    mp.spawn(_worker, args=(world, _free_port(), torus, steps, rows, cols, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    full = (rng.random((rows, cols)) < 0.3).astype(np.uint8)
    want = full
    for _ in range(steps):
        if torus:
            want = oracle.gol_step(want, True)
        else:  # rows clamped, columns wrap: emulate with the block helper on the whole grid
            want = _gol_block(want, None, None)
    got = np.concatenate([np.load(tmp_path / f"block_{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
    assert all(int(np.load(tmp_path / f"total_{r}.npy")[0]) == int(want.sum()) for r in range(world))


def test_block_bounds_cover_all_rows():
    from mesa_b200.sharding import block_bounds

    for rows, world in [(31, 2), (65536, 8), (10, 3), (7, 7)]:
        spans = [block_bounds(rows, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == rows
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _owner_worker(rank, world, port, out_dir):
    """Host-side partition logic of the sharded drivers on a gloo group: every cell / boid has exactly one
    owner, owners agree across ranks, and boids crossing a slab edge are routed to the adjacent rank."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesa_b200.sharded import ShardedBoids
    from mesa_b200.sharding import block_bounds

    rows, cols = 37, 5
    lo, hi = block_bounds(rows, world, rank)
    owned = torch.zeros(rows * cols, dtype=torch.int64)
    owned[lo * cols: hi * cols] = 1
    dist.all_reduce(owned)
    ok = bool((owned == 1).all())           # row blocks tile the grid exactly once
    # global agent ids: exclusive scan of the local counts in rank order (ShardedSchelling.__init__)
    rng = np.random.default_rng(1)
    grid = rng.random((rows, cols)) < 0.7
    counts = torch.zeros(world, dtype=torch.int64)
    counts[rank] = int(grid[lo:hi].sum())
    dist.all_reduce(counts)
    first = int(counts[:rank].sum())
    want_first = int(grid[:lo].sum())
    ok &= first == want_first and int(counts.sum()) == int(grid.sum())
    # slab ownership of boids (ShardedBoids.owner_of) incl. the inclusive top edge
    height = 90.0
    y = torch.tensor([0.0, 29.999, 30.0, 44.9, 45.0, 89.99, 90.0])
    own = ShardedBoids.owner_of(y, height, world)
    slab = height / world
    ok &= bool(((own.double() * slab <= y.double()) | (own == 0)).all()) and int(own.max()) == world - 1
    ok &= bool((own[1:] >= own[:-1]).all())
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_logic_of_sharded_drivers(tmp_path, world):
    mp.spawn(_owner_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(bool(np.load(tmp_path / f"ok_{r}.npy")[0]) for r in range(world))


# ---- deep ghost zones (mesa_b200.sharding.DeepGhostLayer, the scheme PeerBitLife runs on the GPUs) ---------------
def _ghost_worker(rank, world, port, torus, gens, ghost, per_launch, rows, cols, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesa_b200.sharding import DeepGhostLayer

    rng = np.random.default_rng(1)
    full = (rng.random((rows * world, cols)) < 0.35).astype(np.uint8)
    layer = DeepGhostLayer(rows, cols, torch.uint8, ghost, torus, "cpu", transport="p2p")
    layer.state.copy_(torch.from_numpy(full[rank * rows:(rank + 1) * rows]))
    left, launches = gens, 0
    while left > 0:
        if layer.valid == 0:
            layer.refresh()
        t = min(left, per_launch, layer.valid)
        src, dst = layer.planes()
        x = src.numpy().copy()
        for _ in range(t):                      # the extended block is stepped as a clamped grid (dead rows outside)
            x = _gol_block(x, None, None) if torus else _gol_clamped(x)
        dst.copy_(torch.from_numpy(x))
        layer.advance(t)
        left -= t
        launches += 1
    np.save(os.path.join(out_dir, f"ghost_{rank}.npy"), layer.state.numpy())
    np.save(os.path.join(out_dir, f"ghost_stats_{rank}.npy"), np.array([launches, layer.exchanges]))
    dist.destroy_process_group()


def _gol_clamped(ext):
    """One generation of a block whose outside is dead in both directions."""
    p = np.pad(ext.astype(np.int32), 1)
    n = sum(p[1 + di:1 + di + ext.shape[0], 1 + dj:1 + dj + ext.shape[1]] for di in (-1, 0, 1) for dj in (-1, 0, 1)) - ext
    return ((n == 3) | ((ext == 1) & (n == 2))).astype(np.uint8)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("torus", [True, False])
@pytest.mark.parametrize("ghost,per_launch,gens", [(4, 3, 11), (6, 8, 13)])
def test_deep_ghost_zones_equal_unsharded(tmp_path, world, torus, ghost, per_launch, gens):
    """Every rank keeps `ghost` rows of its neighbours and exchanges them once per `ghost` generations: the sharded
    run equals the unsharded one (torus: rows and columns wrap; clamped: dead outside, no ghost rows at the edge)."""
    rows, cols = 12, 20
    mp.spawn(_ghost_worker, args=(world, _free_port(), torus, gens, ghost, per_launch, rows, cols, str(tmp_path)),
             nprocs=world, join=True)
    rng = np.random.default_rng(1)
    want = (rng.random((rows * world, cols)) < 0.35).astype(np.uint8)
    for _ in range(gens):
        if torus:
            n = sum(np.roll(np.roll(want.astype(np.int32), di, 0), dj, 1) for di in (-1, 0, 1) for dj in (-1, 0, 1)) - want
            want = ((n == 3) | ((want == 1) & (n == 2))).astype(np.uint8)
        else:
            want = _gol_clamped(want)
    got = np.concatenate([np.load(tmp_path / f"ghost_{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
    launches, exchanges = np.load(tmp_path / "ghost_stats_0.npy")
    assert exchanges == -(-gens // ghost) and launches >= exchanges


# ---- ShardedLayer (row-sharded property layers: neighbourhood sums, diffusion, aggregates) on gloo -------------------
class _CpuLayerKernels:
    """CPU stand-ins with the signature of mesa_b200.ops (the oracle restatements on a halo-extended block)."""

    @staticmethod
    def _extend(block, radius, torus, top, bot, fill_exists=False):
        x = np.ones(block.shape) if fill_exists else block.numpy().astype(np.float64)
        rows, cols = x.shape

        def halo(h):
            if h is None:
                return np.zeros((radius, cols))
            return np.ones((radius, cols)) if fill_exists else h.numpy().reshape(radius, cols).astype(np.float64)

        ext = np.concatenate([halo(top), x, halo(bot)], axis=0)
        return np.pad(ext, ((0, 0), (radius, radius)), mode="wrap" if torus else "constant")

    @classmethod
    def neighborhood_sum(cls, block, radius, moore, include_center, torus, halo_top=None, halo_bot=None):
        from oracle import layers

        ext = cls._extend(block, radius, torus, halo_top, halo_bot)
        out = layers.neighborhood_sum(ext, radius, moore, include_center, False)[radius:-radius, radius:-radius]
        return torch.from_numpy(out.astype(np.int64 if block.dtype in (torch.uint8, torch.int32) else np.float64))

    @classmethod
    def layer_diffuse(cls, block, rate, torus, out=None, halo_top=None, halo_bot=None):
        from oracle import layers

        acc = layers.neighborhood_sum(cls._extend(block, 1, torus, halo_top, halo_bot), 1, True, False, False)[1:-1, 1:-1]
        k = layers.neighborhood_sum(cls._extend(block, 1, torus, halo_top, halo_bot, True), 1, True, False, False)[1:-1, 1:-1]
        x = block.numpy().astype(np.float64)
        out.copy_(torch.from_numpy((x * (1.0 - rate * k / 8.0) + rate / 8.0 * acc).astype(block.numpy().dtype)))
        return out

    @staticmethod
    def layer_reduce(block, op):
        f = {"sum": np.sum, "min": np.min, "max": np.max, "count_nonzero": np.count_nonzero}[op]
        return torch.tensor([float(f(block.numpy().astype(np.float64)))], dtype=torch.float64)


def _layer_worker(rank, world, port, torus, rows, cols, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mesa_b200.sharding import ShardedLayer, block_bounds

    rng = np.random.default_rng(3)
    counts = rng.integers(0, 9, size=(rows, cols)).astype(np.int32)
    heat = rng.random((rows, cols))
    lo, hi = block_bounds(rows, world, rank)
    layer = ShardedLayer(torch.from_numpy(counts[lo:hi].copy()), torus=torus, kernels=_CpuLayerKernels)
    np.save(os.path.join(out_dir, f"nsum1_{rank}.npy"), layer.neighborhood_sum(1, True, False).numpy())
    np.save(os.path.join(out_dir, f"nsum2_{rank}.npy"), layer.neighborhood_sum(2, False, True).numpy())
    agg = [layer.reduce("sum"), layer.reduce("max"), layer.reduce("min"), layer.reduce("count_nonzero")]
    np.save(os.path.join(out_dir, f"agg_{rank}.npy"), np.array(agg))
    field = ShardedLayer(torch.from_numpy(heat[lo:hi].copy()), torus=torus, kernels=_CpuLayerKernels)
    for _ in range(3):
        field.diffuse(0.4)
    np.save(os.path.join(out_dir, f"heat_{rank}.npy"), field.block.numpy())
    np.save(os.path.join(out_dir, f"mass_{rank}.npy"), np.array([field.reduce("sum")]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,torus", [(2, True), (2, False), (3, True), (3, False)])
def test_sharded_layer_equals_unsharded(tmp_path, world, torus):
    """Neighbourhood sums (r=1 Moore, r=2 von Neumann), three diffusion steps and the aggregates of a row-sharded
    layer equal the unsharded restatements (uneven blocks: 23 rows over 2 / 3 ranks)."""
    from oracle import layers

    rows, cols = 23, 11
    mp.spawn(_layer_worker, args=(world, _free_port(), torus, rows, cols, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    counts = rng.integers(0, 9, size=(rows, cols)).astype(np.int32)
    heat = rng.random((rows, cols))
    cat = lambda name: np.concatenate([np.load(tmp_path / f"{name}_{r}.npy") for r in range(world)])  # noqa: E731
    assert np.array_equal(cat("nsum1"), layers.neighborhood_sum(counts, 1, True, False, torus))
    assert np.array_equal(cat("nsum2"), layers.neighborhood_sum(counts, 2, False, True, torus))
    want = [counts.sum(), counts.max(), counts.min(), np.count_nonzero(counts)]
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"agg_{r}.npy"), np.array(want, dtype=np.float64))
    for _ in range(3):
        heat = layers.diffuse(heat, 0.4, torus)
    np.testing.assert_allclose(cat("heat"), heat, rtol=1e-13, atol=1e-13)
    assert abs(float(np.load(tmp_path / "mass_0.npy")[0]) - heat.sum()) < 1e-9
