"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run there).

* Game of Life 16384^2: equivariance under torus translations and transposition, known oscillators /
  still lifes / a glider embedded in an empty 16384^2 torus, idempotence on still lifes.
* Schelling 8192^2 (one GPU's share of config 5): the packed happiness kernel equals a recomputation from
  the independent generic neighbourhood-sum kernel; relocation is a permutation (types conserved, nobody
  lands on an occupied cell, happy agents stay put, agent_id layer <-> agent_cell array stay a bijection).
* Boltzmann 10^6 agents on 4096^2: wealth is conserved and non-negative, every agent moved to a Moore neighbour.
* Boids 10^6: neighbour counts equal those of the independent radius-query kernel; directions are unit vectors
  for boids with neighbours; positions stay in bounds.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gol(a, steps=1):
    from mesa_b200 import ops

    b = torch.empty_like(a)
    for _ in range(steps):
        ops.gol_step(a, out=b, torus=True)
        a, b = b, a
    return a


def test_gol_full_size_equivariance():
    n = 16384
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.rand((n, n), device="cuda", generator=g) < 0.2).to(torch.uint8)
    y = _gol(x.clone(), 2)
    # torus translation by an odd offset that is not a multiple of the 16-cell vector width / 16-row segment
    sh = (1237, -4099)
    assert torch.equal(_gol(torch.roll(x, sh, (0, 1)).contiguous(), 2), torch.roll(y, sh, (0, 1)))
    # transposition and reflection commute with the rule (Moore neighbourhood is symmetric)
    assert torch.equal(_gol(x.t().contiguous(), 2), y.t())
    assert torch.equal(_gol(torch.flip(x, (1,)).contiguous(), 2), torch.flip(y, (1,)))


def test_gol_full_size_known_patterns():
    n = 16384
    x = torch.zeros((n, n), dtype=torch.uint8, device="cuda")
    # block (still life) across the torus seam, blinker (period 2) across a strip edge, glider (period 4, moves +1,+1)
    for i, j in [(n - 1, n - 1), (n - 1, 0), (0, n - 1), (0, 0)]:
        x[i, j] = 1
    x[5000, 511:514] = 1
    glider = [(100, 101), (101, 102), (102, 100), (102, 101), (102, 102)]
    for i, j in glider:
        x[8000 + i, 9000 + j] = 1
    y1 = _gol(x.clone(), 1)
    assert int(y1[n - 1, n - 1]) == 1 and int(y1[0, 0]) == 1 and int(y1[n - 1, 0]) == 1       # block survives
    assert y1[4999:5002, 512].tolist() == [1, 1, 1] and int(y1[5000, 511]) == 0                 # blinker flipped
    y4 = _gol(x.clone(), 4)
    want = x.clone()
    for i, j in glider:
        want[8000 + i, 9000 + j] = 0
    for i, j in glider:
        want[8000 + i + 1, 9000 + j + 1] = 1
    assert torch.equal(y4, want)                                                                # glider moved by (1, 1)
    assert int(y4.sum()) == int(x.sum())


def test_schelling_full_size_happiness_matches_independent_kernel_and_relocation_is_a_permutation():
    from mesa_b200 import ops

    n, h, radius = 8192, 0.4, 1
    g = torch.Generator(device="cuda").manual_seed(2)
    u = torch.rand((n, n), device="cuda", generator=g)
    types = torch.where(u < 0.2, -1, torch.where(u < 0.6, 0, 1)).to(torch.int8)
    del u
    happy, count = ops.schelling_happy(types, h, radius)

    def recompute(types):
        occ = (types >= 0).to(torch.uint8)
        one = (types == 1).to(torch.uint8)
        nocc = ops.neighborhood_sum(occ, radius, True, False, False)
        none = ops.neighborhood_sum(one, radius, True, False, False)
        similar = torch.where(types == 1, none, nocc - none)
        tbl = torch.from_numpy(ops.happy_threshold_table(h, radius).astype(np.int32)).cuda()
        return ((similar >= tbl[nocc.long()]) & (types >= 0)).to(torch.uint8)

    assert torch.equal(happy, recompute(types))
    assert int(count.item()) == int(happy.sum().item())

    occupied = types.reshape(-1) >= 0
    nag = int(occupied.sum())
    aid = torch.zeros(n * n, dtype=torch.int32, device="cuda")
    aid[occupied] = torch.arange(nag, dtype=torch.int32, device="cuda")
    aid = aid.reshape(n, n)
    cell = torch.nonzero(occupied).reshape(-1).to(torch.int32)
    ws = ops.RelocationWorkspace(types, nag)
    before_t, before_h = types.clone(), happy.clone()
    movers, passes = ops.schelling_relocate(types, happy, aid, ws, 11, 0, agent_cell=cell)
    assert movers == int(((before_t >= 0) & (before_h == 0)).sum())
    for t in (0, 1):
        assert int((types == t).sum()) == int((before_t == t).sum())                      # types conserved
    stayed = (before_t >= 0) & (before_h == 1)
    assert torch.equal(types[stayed], before_t[stayed])                                    # happy agents stay put
    newly = (types >= 0) & ~stayed
    assert bool(((before_t[newly] < 0) | (before_h[newly] == 0)).all())                    # arrivals only on empty / vacated cells
    flat_t, flat_id = types.reshape(-1), aid.reshape(-1)
    assert bool((flat_id[cell.long()] == torch.arange(nag, dtype=torch.int32, device="cuda")).all())   # bijection
    assert bool((flat_t[cell.long()] >= 0).all()) and int((flat_t >= 0).sum()) == nag
    # determinism: the same step from the same state gives the same result
    types2, happy2, aid2 = before_t.clone(), before_h.clone(), torch.zeros_like(aid)
    aid2.reshape(-1)[occupied] = torch.arange(nag, dtype=torch.int32, device="cuda")
    ws2 = ops.RelocationWorkspace(types2, nag)
    ops.schelling_relocate(types2, happy2, aid2, ws2, 11, 0)
    assert torch.equal(types2, types) and torch.equal(aid2[types2 >= 0], aid[types >= 0])
    assert torch.equal(ops.schelling_happy(types, h, radius)[0], recompute(types))


def test_boltzmann_full_size_conservation():
    from mesa_b200 import ops

    n, side = 1_000_000, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    cell = torch.randint(0, side * side, (n,), device="cuda", generator=g).to(torch.int32)
    wealth = torch.ones(n, dtype=torch.int32, device="cuda")
    ws = ops.BoltzmannWorkspace(cell, side, side)
    for e in range(5):
        prev = cell.clone()
        ops.boltzmann_step(cell, wealth, ws, 4, e)
        assert int(wealth.sum()) == n and int(wealth.min()) >= 0                          # wealth conserved, never negative
        di = (cell // side - prev // side).abs()
        dj = (cell % side - prev % side).abs()
        assert int(torch.maximum(di, dj).max()) == 1 and int(torch.maximum(di, dj).min()) == 1   # exactly one Moore step
        assert int(cell.min()) >= 0 and int(cell.max()) < side * side
    assert int(wealth.max()) > 1


def test_boids_full_size_consistency():
    from mesa_b200 import ops

    n, side, vision = 1_000_000, 7500.0, 15.0
    g = torch.Generator(device="cuda").manual_seed(4)
    pos = torch.rand((n, 2), device="cuda", generator=g) * side
    dirs = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    po, do = ops.boids_step(pos, dirs, (side, side), vision=vision, separation=2.0, neighbor_count=cnt)
    off, idx, dist = ops.radius_query(pos, pos, (side, side), vision, True,
                                      exclude=torch.arange(n, dtype=torch.int32, device="cuda"))
    assert torch.equal(torch.diff(off).to(torch.int32), cnt)                               # tiled kernel == query kernel
    assert float(dist.max()) <= vision
    has = cnt > 0
    norm = torch.linalg.norm(do[has].double(), dim=1)
    assert float((norm - 1).abs().max()) < 1e-6                                            # unit directions
    assert torch.equal(do[~has], dirs[~has])                                               # lonely boids keep their heading
    assert float(po.min()) >= 0.0 and float(po.max()) <= side
    step = po.double() - pos.double()
    step = torch.where(step > side / 2, step - side, torch.where(step < -side / 2, step + side, step))
    assert float((torch.linalg.norm(step, dim=1) - torch.where(has, 1.0, torch.linalg.norm(dirs.double(), dim=1))).abs().max()) < 2e-3


def test_gol_full_size_bit_packed_1000_generations_equal_uint8_kernel():
    """BASELINE config 2 at full size and length: 1000 generations of 16384^2 on the bit-packed planes (8 per launch,
    csrc/gol_bits.cu) equal 1000 launches of the uint8-layer kernel (csrc/gol.cu) bit for bit; the two kernels share
    nothing but the rule, and the uint8 kernel is pinned to the oracle / the live reference at small sizes."""
    from mesa_b200 import ops

    n = 16384
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (torch.rand((n, n), device="cuda", generator=g) < 0.2).to(torch.uint8)
    life = ops.BitLife(n, n, "cuda", torus=True)
    life.load(x)
    y = x.clone()
    checked = 0
    for chunk in (1, 7, 64, 928):   # 1000 generations in total, compared at four points
        life.run(chunk)
        y = _gol(y, chunk)
        checked += chunk
        assert torch.equal(life.store(), y), f"bit-packed and uint8 kernels diverge by generation {checked}"
    assert checked == 1000 and int(y.sum()) > 0


def test_gol_full_size_bit_packed_patterns_and_translation():
    """The same pattern / equivariance properties on the bit-packed path (glider across word, strip and torus seams)."""
    from mesa_b200 import ops

    n = 16384
    x = torch.zeros((n, n), dtype=torch.uint8, device="cuda")
    glider = [(0, 1), (1, 2), (2, 0), (2, 1), (2, 2)]
    for oi, oj in [(n - 2, n - 2), (5000, 958), (63, 30), (8191, 16383 - 1)]:   # torus corner, strip edge (30 words), word edge
        for i, j in glider:
            x[(oi + i) % n, (oj + j) % n] = 1
    life = ops.BitLife(n, n, "cuda", torus=True, generations_per_launch=8)
    life.load(x)
    life.run(40)                                   # a glider moves (+1, +1) every 4 generations
    assert torch.equal(life.store(), torch.roll(x, (10, 10), (0, 1)))
    g = torch.Generator(device="cuda").manual_seed(3)
    r = (torch.rand((n, n), device="cuda", generator=g) < 0.3).to(torch.uint8)
    a = ops.BitLife(n, n, "cuda", torus=True); a.load(r); a.run(12)
    b = ops.BitLife(n, n, "cuda", torus=True); b.load(torch.roll(r, (777, -3333), (0, 1)).contiguous()); b.run(12)
    assert torch.equal(b.store(), torch.roll(a.store(), (777, -3333), (0, 1)))


# ---- property layers at 16384^2 (the oracle restatements take minutes there) ------------------------------------
def test_layer_kernels_full_size_properties():
    """Neighbourhood sums: closed-form sizes of a clamped grid, linearity, equivariance under torus translations and
    transposition, the walker kernel against the independent generic (radius 3) kernel through the identity
    box(r=1) of box(r=1) = weighted r=2 box; reductions: a checksum of row-block checksums; elementwise: involution."""
    from mesa_b200 import ops

    n = 16384
    g = torch.Generator(device="cuda").manual_seed(5)
    a = (torch.rand((n, n), device="cuda", generator=g) * 256).to(torch.uint8)
    # sizes of the radius-1 / radius-2 neighbourhoods of a clamped grid (cell.py:225-276): (rows seen) * (cols seen) - 1
    ones = torch.ones((n, n), dtype=torch.uint8, device="cuda")
    idx = torch.arange(n, device="cuda")
    for r in (1, 2):
        seen = torch.minimum(idx, torch.full_like(idx, r)) + 1 + torch.minimum(n - 1 - idx, torch.full_like(idx, r))
        want = (seen[:, None] * seen[None, :] - 1).to(torch.int32)
        assert torch.equal(ops.neighborhood_sum(ones, r, True, False, False), want)
    del ones, want
    s1 = ops.neighborhood_sum(a, 1, True, False, True)
    # translation / transposition equivariance on the torus (offsets that are no multiple of a vector or a segment)
    sh = (1237, -4099)
    assert torch.equal(ops.neighborhood_sum(torch.roll(a, sh, (0, 1)).contiguous(), 1, True, False, True), torch.roll(s1, sh, (0, 1)))
    assert torch.equal(ops.neighborhood_sum(a.t().contiguous(), 1, True, False, True), s1.t())
    # linearity over int32 layers: sum(a) + sum(b) = sum(a + b)
    b = (torch.rand((n, n), device="cuda", generator=g) * 1000).to(torch.int32)
    ai = a.to(torch.int32)
    lhs = ops.neighborhood_sum(ops.layer_binary(ai, b, "add"), 2, False, True, True)
    rhs = ops.layer_binary(ops.neighborhood_sum(ai, 2, False, True, True), ops.neighborhood_sum(b, 2, False, True, True), "add")
    assert torch.equal(lhs, rhs)
    del b, lhs, rhs
    # checksum of checksums: the whole-layer sum equals the sum over 64 row blocks; count_nonzero likewise
    total = int(ops.layer_reduce(s1, "sum").item())
    assert total == 8 * int(ops.layer_reduce(a, "sum").item())        # every cell is counted by its 8 neighbours
    parts = sum(int(ops.layer_reduce(s1[k * 256:(k + 1) * 256], "sum").item()) for k in range(64))
    assert parts == total
    assert int(ops.layer_reduce(a, "count_nonzero").item()) == int((a != 0).sum().item())
    assert int(ops.layer_reduce(a, "max").item()) == 255 and int(ops.layer_reduce(a, "min").item()) == 0
    # elementwise: x -> 255 - x twice is the identity (uint8 table path), min / max split a + b
    inv = ops.layer_affine_clamp(a, -1.0, 255.0, out=torch.empty_like(a))
    assert torch.equal(ops.layer_affine_clamp(inv, -1.0, 255.0, out=torch.empty_like(a)), a)
    lo, hi = ops.layer_binary(a, inv, "min"), ops.layer_binary(a, inv, "max")
    assert torch.equal(ops.layer_binary(lo, hi, "add"), torch.full_like(a, 255))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_diffusion_full_size_properties(dtype):
    """Diffusion on a 16384^2 torus: a constant layer is a fixed point, mass is conserved, the step commutes with
    translations, and rate 0 is the identity; on a clamped grid mass is conserved as well (edge cells keep the
    shares that have nowhere to go)."""
    from mesa_b200 import ops

    n = 16384
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand((n, n), device="cuda", generator=g, dtype=dtype)
    tol = 1e-6 if dtype == torch.float32 else 1e-13
    const = torch.full((n, n), 0.75, dtype=dtype, device="cuda")
    assert torch.equal(ops.layer_diffuse(const, 0.5, True), const)
    del const
    assert torch.equal(ops.layer_diffuse(x, 0.0, True), x)
    y = ops.layer_diffuse(x, 0.4, True)
    mass = float(ops.layer_reduce(x, "sum").item())
    assert abs(float(ops.layer_reduce(y, "sum").item()) - mass) <= tol * mass
    sh = (-517, 2049)
    assert torch.equal(ops.layer_diffuse(torch.roll(x, sh, (0, 1)).contiguous(), 0.4, True), torch.roll(y, sh, (0, 1)))
    assert float(y.min().item()) >= 0.0 and float(y.max().item()) <= 1.0      # a convex combination of the 9 cells
    z = ops.layer_diffuse(x, 0.4, False)
    assert abs(float(ops.layer_reduce(z, "sum").item()) - mass) <= tol * mass
