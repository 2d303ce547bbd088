"""GPU parity: radix sort, cell-list radius queries and the Boids step vs golden fixtures / oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BOIDS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "boids_*.npz")))
RTOL = 1e-5  # north_star: fp32 Boids position/velocity within 1e-5 relative of the float64 reference


def _kw(g):
    return dict(vision=float(g["vision"]), separation=float(g["separation"]), cohere=float(g["cohere"]),
                separate=float(g["separate"]), match=float(g["match"]), speed=float(g["speed"]), torus=True)


@pytest.mark.parametrize("n,bits", [(1, 32), (1000, 32), (4096, 16), (100_003, 24), (70_000, 64)])
def test_radix_sort_is_stable_and_sorted(n, bits):
    from mesa_b200 import ops

    rng = np.random.default_rng(n)
    if bits == 64:
        keys = rng.integers(0, 2**62, size=n, dtype=np.int64)
        keys[::3] = keys[0]  # many duplicates
    else:
        keys = rng.integers(0, min(2**bits, 2**31 - 1), size=n, dtype=np.int64).astype(np.int32)
        keys[::5] = 7 % (2**bits)
    vals = np.arange(n, dtype=np.int32)
    k, v = ops.sort_pairs(torch.from_numpy(keys).cuda(), torch.from_numpy(vals).cuda(), 0, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k.cpu().numpy(), keys[order])
    assert np.array_equal(v.cpu().numpy(), vals[order])


@pytest.mark.parametrize("name", BOIDS)
def test_boids_step_matches_reference(name):
    from mesa_b200 import ops

    g = np.load(os.path.join(GOLDEN, name))
    size = g["size"]
    for t in range(int(g["steps"])):
        pos = torch.from_numpy(g[f"pos_in_{t}"]).cuda()
        dirs = torch.from_numpy(g[f"dir_in_{t}"]).cuda()
        cnt = torch.zeros(int(g["n"]), dtype=torch.int32, device="cuda")
        po, do = ops.boids_step(pos, dirs, size, neighbor_count=cnt, **_kw(g))
        # neighbour sets are bit-exact (counts here, full CSR below)
        assert np.array_equal(cnt.cpu().numpy(), g[f"nbr_count_{t}"])
        want_p, want_d = g[f"pos_out_{t}"], g[f"dir_out_{t}"]
        # norm-wise relative error: |dp| / size, |dd| / |d| (unit vectors)
        assert np.max(np.abs(po.cpu().numpy().astype(np.float64) - want_p) / size) <= RTOL
        assert np.max(np.abs(do.cpu().numpy().astype(np.float64) - want_d)) <= RTOL
        off, idx, dist = ops.radius_query(pos, pos, size, float(g["vision"]), True,
                                          exclude=torch.arange(int(g["n"]), dtype=torch.int32, device="cuda"))
        assert np.array_equal(np.diff(off.cpu().numpy()), g[f"nbr_count_{t}"])
        assert np.array_equal(idx.cpu().numpy(), g[f"nbr_idx_{t}"])


@pytest.mark.parametrize("n,side,vision,torus", [(2000, 200.0, 10.0, True), (500, 25.0, 10.0, True),
                                                 (3000, 300.0, 7.5, False), (64, 8.0, 5.0, True), (1, 10.0, 3.0, True)])
def test_boids_and_radius_match_oracle(n, side, vision, torus):
    from mesa_b200 import ops

    rng = np.random.default_rng(n)
    size = np.array([side, side * 0.75])
    pos = (rng.random((n, 2)) * size).astype(np.float32)
    pos[: min(n, 4)] = [[0.0, 0.0], [side, 0.0], [0.0, side * 0.75], [side, side * 0.75]][: min(n, 4)]
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    kw = dict(vision=vision, separation=2.0, torus=torus)
    wp, wd, wc = oracle.boids_step(pos, dirs, size, **kw) if torus else (None, None, None)
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    if torus:
        po, do = ops.boids_step(torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda(), size,
                                neighbor_count=cnt, **kw)
        assert np.array_equal(cnt.cpu().numpy(), wc)
        assert np.max(np.abs(po.cpu().numpy() - wp) / size) <= RTOL
        assert np.max(np.abs(do.cpu().numpy() - wd)) <= RTOL
    q = pos[:: max(1, n // 50)]
    off, idx, dist = ops.radius_query(torch.from_numpy(pos).cuda(), torch.from_numpy(q.copy()).cuda(), size, vision, torus)
    off, idx, dist = off.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy()
    for i, p in enumerate(q):
        wi, wdist = oracle.radius_query(pos.astype(np.float64), p.astype(np.float64), size, vision, torus)
        assert np.array_equal(idx[off[i]:off[i + 1]], wi)
        assert np.array_equal(dist[off[i]:off[i + 1]], wdist)  # float64 distances are bit-exact


# ---- sequential activation: the reference's shuffle_do semantics (csrc/continuous.cu, boids_seq_pass) ----------
@pytest.mark.parametrize("name", BOIDS)
def test_boids_sequential_matches_reference_model_step(name):
    """One model.step() of the UNMODIFIED reference under the replayed activation order (golden seq_*_1)."""
    from mesa_b200 import ops

    g = np.load(os.path.join(GOLDEN, name))
    n, seed, size = int(g["n"]), int(g["philox_seed"]), g["size"]
    pos0, dir0 = g["seq_pos_0"], g["seq_dir_0"]
    assert np.array_equal(pos0, pos0.astype(np.float32)) and np.array_equal(dir0, dir0.astype(np.float32))
    po, do, passes = ops.boids_step_sequential(torch.from_numpy(pos0.astype(np.float32)).cuda(),
                                               torch.from_numpy(dir0.astype(np.float32)).cuda(), size, seed=seed,
                                               epoch=0, **_kw(g))
    assert passes >= 1
    assert np.max(np.abs(po.cpu().numpy().astype(np.float64) - g["seq_pos_1"]) / size) <= RTOL
    assert np.max(np.abs(do.cpu().numpy().astype(np.float64) - g["seq_dir_1"])) <= RTOL


@pytest.mark.parametrize("n,side,vision,torus,speed", [(2000, 200.0, 10.0, True, 1.0), (500, 25.0, 10.0, True, 1.0),
                                                       (3000, 300.0, 7.5, False, 1.0), (64, 8.0, 5.0, True, 1.0),
                                                       (1, 10.0, 3.0, True, 1.0), (4000, 260.0, 9.0, True, 2.5)])
def test_boids_sequential_matches_oracle(n, side, vision, torus, speed):
    """Several steps against the oracle's sequential mode (pinned to the reference in tests/test_oracle_golden.py)
    in the same activation order; neighbour counts exact, state within the north-star tolerance."""
    from mesa_b200 import ops

    rng = np.random.default_rng(n + 1)
    size = np.array([side, side * 0.75])
    pos = (rng.random((n, 2)) * size).astype(np.float32)
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    seed = 1234
    kw = dict(vision=vision, separation=2.0, torus=torus, speed=speed)
    if not torus:
        # keep the flock inside the clamped space for the steps tested (the reference raises when a boid leaves)
        pos = (size * 0.25 + rng.random((n, 2)) * size * 0.5).astype(np.float32)
    reach = ops.boids_reach(torch.from_numpy(dirs).cuda(), speed)
    for epoch in range(3):
        order = oracle.activation_order(seed, epoch, n)
        wp, wd, wc = oracle.boids_step(pos, dirs, size, order=order, **kw)
        cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
        po, do, passes = ops.boids_step_sequential(torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda(), size,
                                                   seed=seed, epoch=epoch, reach=reach, neighbor_count=cnt, **kw)
        assert np.array_equal(cnt.cpu().numpy(), wc), (epoch, passes)
        assert np.max(np.abs(po.cpu().numpy() - wp) / size) <= RTOL
        assert np.max(np.abs(do.cpu().numpy() - wd)) <= RTOL
        pos, dirs = po.cpu().numpy(), do.cpu().numpy()   # float32 state at the step boundary on both sides


def test_boids_sequential_differs_from_synchronous_and_rejects_short_reach():
    from mesa_b200 import ops

    rng = np.random.default_rng(0)
    size = (150.0, 150.0)
    pos = torch.from_numpy((rng.random((3000, 2)) * 150).astype(np.float32)).cuda()
    dirs = torch.from_numpy(rng.uniform(-1, 1, (3000, 2)).astype(np.float32)).cuda()
    ps, ds = ops.boids_step(pos, dirs, size, vision=10.0, separation=2.0)
    pq, dq, passes = ops.boids_step_sequential(pos, dirs, size, seed=5, epoch=0, vision=10.0, separation=2.0)
    assert passes >= 1 and float((ps - pq).abs().max()) > 1e-3   # Gauss-Seidel is not Jacobi
    with pytest.raises(ValueError):
        ops.boids_step_sequential(pos, dirs, size, seed=5, epoch=0, vision=10.0, separation=2.0, reach=0.5)


@pytest.mark.parametrize("n", [1, 1000, 300_000, 3_000_000])
def test_activation_order_matches_oracle(n):
    """The device permutation (radix sort on the priority's high word + tie fix-up by agent index) equals the
    oracle's; at n = 3e6 about a thousand pairs of agents share a high word, so the tie path is exercised."""
    from mesa_b200 import ops

    for seed, epoch in ((42, 0), (2**63 + 5, 7)):
        got = ops.activation_order(seed, epoch, n, "cuda").cpu().numpy()
        want = oracle.activation_order(seed, epoch, n)
        assert np.array_equal(got, want)
    if n >= 3_000_000:
        import oracle.philox as ph

        hi = (ph.priority64(42, 0, np.arange(n, dtype=np.uint64)) >> np.uint64(32))
        assert len(np.unique(hi)) < n  # ties exist in this case


def test_boids_sequential_schemes_agree(monkeypatch):
    """The three schedules of the sequential step -- the rank-ordered chained kernel (default), priority rounds in
    cell order (MB_BOIDS_SEQ=rounds) and one round (MB_BOIDS_SEQ=passes) -- are the same function."""
    from mesa_b200 import ops

    rng = np.random.default_rng(12)
    n, size = 20000, (700.0, 500.0)
    pos = torch.from_numpy((rng.random((n, 2)) * np.array(size)).astype(np.float32)).cuda()
    dirs = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).cuda()
    kw = dict(seed=3, epoch=1, vision=12.0, separation=2.0)
    monkeypatch.delenv("MB_BOIDS_SEQ", raising=False)
    results = []
    for env in ({"MB_BOIDS_SEQ": "rounds", "MB_BOIDS_SEQ_ROUNDS": "16"},
                {"MB_BOIDS_SEQ": "rounds", "MB_BOIDS_SEQ_ROUNDS": "4", "MB_BOIDS_SEQ_SUBPASSES": "1"},
                {"MB_BOIDS_SEQ": "passes"}, {}):
        for k in ("MB_BOIDS_SEQ", "MB_BOIDS_SEQ_ROUNDS", "MB_BOIDS_SEQ_SUBPASSES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        results.append(ops.boids_step_sequential(pos, dirs, size, **kw))
    assert results[0][2] >= 32 and results[3][2] == 1
    for p_, d_, _ in results[1:]:
        assert torch.equal(results[0][0], p_) and torch.equal(results[0][1], d_)


def test_cell_list_with_crowded_cells_matches_oracle():
    """Dense clusters: cells with more than 32 agents take the cooperative ordering path of the counting-sort cell
    list (sort_long_cells); neighbour sets, query rows and the Boids step must still equal the oracle."""
    from mesa_b200 import ops

    rng = np.random.default_rng(21)
    size = np.array([400.0, 300.0])
    spread = (rng.random((1500, 2)) * size).astype(np.float32)
    blob1 = (np.array([120.0, 80.0]) + rng.random((700, 2)) * 6.0).astype(np.float32)      # ~700 boids in one or two cells
    blob2 = (np.array([399.0, 299.0]) + rng.random((90, 2)) * 1.0).astype(np.float32)      # a crowded corner cell at the seam
    pos = np.concatenate([spread, blob1, np.minimum(blob2, size.astype(np.float32))]).astype(np.float32)
    perm = rng.permutation(len(pos))                                                       # crowded cells hold scattered ids
    pos = pos[perm]
    n = len(pos)
    dirs = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    kw = dict(vision=10.0, separation=2.0, torus=True)
    wp, wd, wc = oracle.boids_step(pos, dirs, size, **kw)
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    po, do = ops.boids_step(torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda(), size, neighbor_count=cnt, **kw)
    assert int(wc.max()) > 600
    assert np.array_equal(cnt.cpu().numpy(), wc)
    assert np.max(np.abs(po.cpu().numpy() - wp) / size) <= RTOL
    assert np.max(np.abs(do.cpu().numpy() - wd)) <= RTOL
    q = pos[::37]
    off, idx, dist = ops.radius_query(torch.from_numpy(pos).cuda(), torch.from_numpy(q.copy()).cuda(), size, 10.0, True)
    off, idx, dist = off.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy()
    for i, p in enumerate(q):
        wi, wdist = oracle.radius_query(pos.astype(np.float64), p.astype(np.float64), size, 10.0, True)
        assert np.array_equal(idx[off[i]:off[i + 1]], wi)
        assert np.array_equal(dist[off[i]:off[i + 1]], wdist)
    # the sequential step runs on the same cell list
    order = oracle.activation_order(77, 0, n)
    sp, sd, sc = oracle.boids_step(pos, dirs, size, order=order, **kw)
    qo, qd, _ = ops.boids_step_sequential(torch.from_numpy(pos).cuda(), torch.from_numpy(dirs).cuda(), size, seed=77, epoch=0,
                                          neighbor_count=cnt, **kw)
    assert np.array_equal(cnt.cpu().numpy(), sc)
    assert np.max(np.abs(qo.cpu().numpy() - sp) / size) <= RTOL and np.max(np.abs(qd.cpu().numpy() - sd)) <= RTOL


# ---- point queries of the drop-in space vs the UNMODIFIED reference (golden) and the oracle -----------------
def test_continuous_queries_match_reference_golden():
    """calculate_distances, calculate_difference_vector, get_agents_in_radius, get_k_nearest_agents and
    get_nearest_neighbors of DeviceContinuousSpace against ContinuousSpace itself: indices identical,
    float64 distances / vectors bit-identical (positions are float32-representable)."""
    import random

    from mesa_b200.continuous_space import ContinuousSpace

    g = np.load(os.path.join(GOLDEN, "continuous_queries.npz"))
    for c, (n, torus, radius, k, kk) in enumerate(g["cases"]):
        n, torus, k, kk = int(n), bool(torus), int(k), int(kk)
        pos, points, dims, sub = g[f"pos_{c}"], g[f"points_{c}"], g[f"dims_{c}"], g[f"subset_{c}"]
        space = ContinuousSpace(dims, torus=torus, random=random.Random(1), n_agents=n)
        space.add_agents(pos)
        for q, pt in enumerate(points):
            d, ids = space.calculate_distances(pt)
            assert ids == list(range(n)) and np.array_equal(d, g[f"dist_all_{c}"][q])
            assert np.array_equal(space.calculate_difference_vector(pt), g[f"delta_all_{c}"][q])
            d, ids = space.calculate_distances(pt, sub.tolist())
            assert ids == sub.tolist() and np.array_equal(d, g[f"dist_sub_{c}"][q])
            assert np.array_equal(space.calculate_difference_vector(pt, sub.tolist()), g[f"delta_sub_{c}"][q])
            lo, hi = g[f"rad_off_{c}"][q], g[f"rad_off_{c}"][q + 1]
            idx, dist = space.get_agents_in_radius(pt, radius)
            assert idx == g[f"rad_idx_{c}"][lo:hi].tolist() and np.array_equal(dist, g[f"rad_dist_{c}"][lo:hi])
            idx, dist = space.get_k_nearest_agents(pt, k)
            assert idx == g[f"knn_idx_{c}"][q].tolist() and np.array_equal(dist, g[f"knn_dist_{c}"][q])
        # batched: all points in one launch, and every agent's own nearest neighbours
        idx, dist = space.k_nearest_query(points, k)
        assert np.array_equal(idx.cpu().numpy(), g[f"knn_idx_{c}"]) and np.array_equal(dist.cpu().numpy(), g[f"knn_dist_{c}"])
        idx, dist = space.nearest_neighbors(kk)
        m = g[f"nn_idx_{c}"].shape[0]
        assert np.array_equal(idx[:m].cpu().numpy(), g[f"nn_idx_{c}"]) and np.array_equal(dist[:m].cpu().numpy(), g[f"nn_dist_{c}"])


def test_continuous_queries_nd_match_reference_golden():
    """1-, 3- and 4-dimensional spaces (tests/experimental/test_continuous_space.py:38-56 builds a 3-d one): the point
    queries of DeviceContinuousSpace against ContinuousSpace itself, float64 results bit-identical."""
    import random

    from mesa_b200 import ops
    from mesa_b200.continuous_space import ContinuousSpace

    g = np.load(os.path.join(GOLDEN, "continuous_queries_nd.npz"))
    for c, (n, torus, radius, k) in enumerate(g["cases"]):
        n, torus, k = int(n), bool(torus), int(k)
        pos, points, dims, sub = g[f"pos_{c}"], g[f"points_{c}"], g[f"dims_{c}"], g[f"subset_{c}"]
        space = ContinuousSpace(dims, torus=torus, random=random.Random(1), n_agents=4)
        space.add_agents(pos[: n // 2])
        space.add_agents(pos[n // 2:])                                   # growth keeps the earlier rows
        assert space.ndims == dims.shape[0] and tuple(space.agent_positions.shape) == (n, dims.shape[0])
        for q, pt in enumerate(points):
            d, ids = space.calculate_distances(pt)
            assert ids == list(range(n)) and np.array_equal(d, g[f"dist_all_{c}"][q])
            assert np.array_equal(space.calculate_difference_vector(pt), g[f"delta_all_{c}"][q])
            d, ids = space.calculate_distances(pt, sub.tolist())
            assert ids == sub.tolist() and np.array_equal(d, g[f"dist_sub_{c}"][q])
            assert np.array_equal(space.calculate_difference_vector(pt, sub.tolist()), g[f"delta_sub_{c}"][q])
            lo, hi = g[f"rad_off_{c}"][q], g[f"rad_off_{c}"][q + 1]
            idx, dist = space.get_agents_in_radius(pt, radius)
            assert idx == g[f"rad_idx_{c}"][lo:hi].tolist() and np.array_equal(dist, g[f"rad_dist_{c}"][lo:hi])
            idx, dist = space.get_k_nearest_agents(pt, k)
            assert idx == g[f"knn_idx_{c}"][q].tolist() and np.array_equal(dist, g[f"knn_dist_{c}"][q])
        with pytest.raises(NotImplementedError):
            space.neighbors_in_radius(radius)                            # the batched queries run on the 2-d cell list
    # the 3-d space of the reference's own test: fields and bounds
    space = ContinuousSpace([[0, 2], [-2, 0], [-2, 2]], torus=False, random=random.Random(1))
    assert space.ndims == 3 and space.size.tolist() == [2, 2, 4] and space.center.tolist() == [1, -1, 0]
    assert space.in_bounds([1, -1, 0]) and not space.in_bounds([1, -1, 3]) and not space.in_bounds([2.5, -1, 0])
    with pytest.raises(ValueError):
        space.add_agents([[1.0, -1.0, 3.0]])
    # in two dimensions the n-d kernel and the 2-d kernel are the same function
    g2 = np.load(os.path.join(GOLDEN, "continuous_queries.npz"))
    pos2 = torch.from_numpy(g2["pos_0"].astype(np.float32)).cuda()
    size2 = g2["dims_0"][:, 1] - g2["dims_0"][:, 0]
    for pt in g2["points_0"]:
        a = ops.point_distances(pos2, pt, size2, True, distances=True, deltas=True)
        b = ops.point_distances_nd(pos2, pt, size2, True, distances=True, deltas=True)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("n,side,torus", [(20_000, (300.0, 200.0), True), (20_000, (300.0, 200.0), False),
                                          (5_000, (40.0, 2.0), True), (3, (10.0, 10.0), True)])
@pytest.mark.parametrize("k", [1, 4, 16])
def test_knn_query_matches_oracle(n, side, torus, k):
    """Ring search on the cell list vs brute force; clustered agents (empty cells force wide rings), k > n padding."""
    from mesa_b200 import ops

    rng = np.random.default_rng(n + k)
    pos = rng.random((n, 2)) * side
    pos[: n // 2] = pos[: n // 2] * 0.05 + np.asarray(side) * 0.6        # half of the agents in a small cluster
    pos = pos.astype(np.float32)
    queries = (rng.random((300, 2)) * side).astype(np.float32)
    idx, dist = ops.knn_query(torch.from_numpy(pos).cuda(), torch.from_numpy(queries).cuda(), side, k, torus)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for q in range(queries.shape[0]):
        wi, wd = oracle.knn_query(pos.astype(np.float64), queries[q].astype(np.float64), side, k, torus)
        m = len(wi)
        assert np.array_equal(idx[q, :m], wi) and np.array_equal(dist[q, :m], wd), q
        assert np.all(idx[q, m:] == -1) and np.all(np.isinf(dist[q, m:]))


def test_knn_query_ties_rank_by_storage_index():
    """Agents on a lattice: many exactly equal distances; the earlier storage index wins (continuous_space.py:256-257)."""
    from mesa_b200 import ops

    xs, ys = np.meshgrid(np.arange(0.5, 16.0, 1.0), np.arange(0.5, 16.0, 1.0), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.float32)
    pos = pos[np.random.default_rng(0).permutation(len(pos))]
    me = torch.arange(len(pos), dtype=torch.int32, device="cuda")
    for torus in (True, False):
        for k, edge in ((4, None), (8, 0.7), (9, 3.0)):
            idx, dist = ops.knn_query(torch.from_numpy(pos).cuda(), torch.from_numpy(pos).cuda(), (16.0, 16.0), k, torus,
                                      exclude=me, cell_edge=edge)
            idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
            for q in range(0, len(pos), 7):
                wi, wd = oracle.knn_query(pos.astype(np.float64), pos[q].astype(np.float64), (16.0, 16.0), k, torus, exclude=q)
                assert np.array_equal(idx[q], wi) and np.array_equal(dist[q], wd), (torus, k, q)


def test_knn_and_point_queries_on_empty_space():
    import random

    from mesa_b200.continuous_space import ContinuousSpace

    space = ContinuousSpace([[0, 1], [0, 1]], torus=True, random=random.Random(1))
    assert space.get_k_nearest_agents([0.5, 0.5], k=3)[0] == []
    d, ids = space.calculate_distances([0.5, 0.5])
    assert len(d) == 0 and ids == []
    assert space.calculate_difference_vector([0.5, 0.5]).shape == (0, 2)
    space.add_agents([[0.2, 0.2]])
    assert space.get_k_nearest_agents([0.5, 0.5], k=0)[0] == []
    idx, dist = space.k_nearest_query(np.array([[0.9, 0.9]]), 2)
    assert idx.cpu().tolist() == [[0, -1]] and np.isinf(dist[0, 1].item())


def test_resort_keeps_the_flock_and_its_ids():
    """ops.boids_resort rewrites the population in cell-list order: the same boids (ids travel with the rows), and a model that
    re-sorts every few steps follows the same trajectory as one that never does (neighbour sets exact, state within 1e-5: the
    float32 neighbour sums run in another order)."""
    from mesa_b200 import ops
    from mesa_b200.examples import BoidFlockers, BoidsScenario

    rng = np.random.default_rng(4)
    n, side = 20000, 600.0
    pos = torch.from_numpy((rng.random((n, 2)) * side).astype(np.float32)).cuda()
    dirs = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32)).cuda()
    ids = torch.arange(100, 100 + n, dtype=torch.int64, device="cuda")
    p2, d2, i2, perm = ops.boids_resort(pos, dirs, (side, side), vision=15.0, ids=ids)
    assert torch.equal(p2, pos[perm.long()]) and torch.equal(d2, dirs[perm.long()]) and torch.equal(i2, ids[perm.long()])
    assert torch.equal(torch.sort(perm.long()).values, torch.arange(n, device="cuda"))
    ncx = int(side // (15.0 * (1 + 1e-6)))
    cell = (torch.floor(p2[:, 1].double() * ncx / side).long().clamp(0, ncx - 1) * ncx
            + torch.floor(p2[:, 0].double() * ncx / side).long().clamp(0, ncx - 1))
    assert bool((cell[1:] >= cell[:-1]).all())                       # cell-list order
    same_cell = cell[1:] == cell[:-1]
    assert bool((perm[1:][same_cell] > perm[:-1][same_cell]).all())  # previous order inside a cell

    sc = BoidsScenario(population_size=3000, width=300, height=300, vision=10.0, rng=5)
    a = BoidFlockers(sc, activation="synchronous", resort_every=0)
    b = BoidFlockers(sc, activation="synchronous", resort_every=3)
    for _ in range(10):
        a.step()
        b.step()
    order = torch.argsort(b.agents.unique_ids())
    assert torch.equal(b.agents.unique_ids()[order], a.agents.unique_ids())
    pa, pb = a.agents.get("position"), b.agents.get("position")[order]
    da, db = a.agents.get("direction"), b.agents.get("direction")[order]
    assert float((pa - pb).abs().max()) / 300.0 <= 1e-5 and float((da - db).abs().max()) <= 1e-5
    assert torch.equal(a._nbr_count, b._nbr_count[order])
