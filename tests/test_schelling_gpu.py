"""GPU parity: Schelling happiness + relocation vs the golden fixtures (live reference) and the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SCHELLING = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "schelling_*.npz")))


class DeviceSchelling:
    """Minimal driver over the ops (the drop-in classes are tested in test_dropin_gpu.py)."""

    def __init__(self, st: oracle.SchellingState, homophily, radius, torus=False):
        from mesa_b200 import ops

        self.ops = ops
        self.rows, self.cols = st.rows, st.cols
        self.h, self.r, self.torus = homophily, radius, torus
        self.type = torch.from_numpy(st.type_grid()).cuda()
        ids = st.id_grid()
        ids[ids < 0] = 0
        self.agent_id = torch.from_numpy(ids.astype(np.uint32).view(np.int32)).cuda().view(torch.uint32)
        self.agent_cell = torch.from_numpy(st.cell.astype(np.uint32).view(np.int32)).cuda().view(torch.uint32)
        self.happy = torch.zeros((self.rows, self.cols), dtype=torch.uint8, device="cuda")
        self.count = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.ws = ops.RelocationWorkspace(self.type, st.n)

    def assign(self):
        self.ops.schelling_happy(self.type, self.h, self.r, self.torus, out=self.happy, count=self.count)
        return int(self.count.item())

    def move(self, seed, epoch):
        return self.ops.schelling_relocate(self.type, self.happy, self.agent_id, self.ws, seed, epoch,
                                           agent_cell=self.agent_cell)

    def cells(self):
        return self.agent_cell.cpu().view(torch.int32).numpy().view(np.uint32).astype(np.int64)

    def happy_by_agent(self):
        return self.happy.reshape(-1).cpu().numpy()[self.cells()]


@pytest.mark.parametrize("name", SCHELLING)
def test_schelling_matches_reference_trajectory(name):
    g = np.load(os.path.join(GOLDEN, name))
    rows, cols = int(g["width"]), int(g["height"])
    h, r, seed = float(g["homophily"]), int(g["radius"]), int(g["philox_seed"])
    st = oracle.SchellingState(rows, cols, g["cell0"], g["atype"])
    dev = DeviceSchelling(st, h, r)
    assert dev.assign() == int(g["happy_count0"])
    assert np.array_equal(dev.happy_by_agent(), g["happy0"])
    full = set(g["full_steps"].tolist())
    for s in range(1, len(g["happy_count"]) + 1):
        movers, iters = dev.move(seed, s - 1)
        assert movers == int(g["movers"][s - 1]), s
        happy = dev.assign()
        assert happy == int(g["happy_count"][s - 1]), s
        if s in full:
            assert np.array_equal(dev.cells(), g[f"cell_{s}"]), s
            assert np.array_equal(dev.happy_by_agent(), g[f"happy_{s}"]), s
            # the layers stay consistent with the per-agent view
            tg = dev.type.cpu().numpy().reshape(-1)
            assert np.array_equal(tg[dev.cells()], g["atype"])
            assert (tg >= 0).sum() == len(g["atype"])


@pytest.mark.parametrize("shape,radius,h,torus", [
    ((64, 64), 1, 0.4, False), ((100, 100), 2, 1.0, False), ((48, 528), 2, 0.55, False),
    ((130, 1040), 1, 0.375, True), ((33, 512), 2, 0.5, True), ((37, 23), 1, 0.4, False),
    ((37, 23), 3, 0.3, True), ((16, 16), 1, 0.0, False), ((16, 32), 1, 1.5, False),
    # widths that are multiples of 32 take the bit-plane kernel (csrc/schelling_bits.cu): partial strips, several strips,
    # fewer rows than a row group, single rows, rules with and without a small integer form
    ((130, 1056), 1, 0.375, True), ((67, 2080), 2, 0.55, False), ((5, 32), 2, 0.4, True), ((1, 64), 1, 0.4, False),
    ((70, 1024), 2, 1.0, True), ((200, 96), 1, 0.7, False), ((3, 1024), 1, 1.0, True), ((64, 64), 2, 0.0, False),
    ((41, 160), 1, 0.123, False), ((41, 160), 2, 0.9, True), ((29, 32), 1, 0.5, True), ((29, 64), 2, 0.3333333333333333, False),
])
def test_happy_layer_matches_oracle(shape, radius, h, torus):
    rng = np.random.default_rng(1)
    u = rng.random(shape)
    grid = np.where(u < 0.25, -1, np.where(u < 0.6, 0, 1)).astype(np.int8)
    st = oracle.SchellingState.from_type_grid(grid)
    want_count = st.assign_state(h, radius, torus)
    from mesa_b200 import ops

    happy, count = ops.schelling_happy(torch.from_numpy(grid).cuda(), h, radius, torus)
    assert int(count.item()) == want_count
    assert np.array_equal(happy.cpu().numpy(), st.happy_grid())


@pytest.mark.parametrize("shape,radius,h", [((96, 128), 1, 0.4), ((96, 128), 2, 0.55), ((50, 1056), 2, 1.0), ((50, 48), 1, 0.4)])
def test_happy_row_blocks_with_halo_rows_equal_the_whole_grid(shape, radius, h):
    """The row-sharded form: every block gets the `radius` rows above / below it as halo rows (None at the outer edges);
    the blocks' happy layers and counts add up to the whole grid's."""
    from mesa_b200 import ops

    rng = np.random.default_rng(5)
    u = rng.random(shape)
    grid = torch.from_numpy(np.where(u < 0.25, -1, np.where(u < 0.6, 0, 1)).astype(np.int8)).cuda()
    whole, count = ops.schelling_happy(grid, h, radius, False)
    cuts = [0, 7, 7 + radius, 40, shape[0]]
    total = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        top = grid[lo - radius:lo].contiguous() if lo >= radius else None
        bot = grid[hi:hi + radius].contiguous() if hi + radius <= shape[0] else None
        part, c = ops.schelling_happy(grid[lo:hi].contiguous(), h, radius, False, halo_top=top, halo_bot=bot)
        assert torch.equal(part, whole[lo:hi]), (lo, hi)
        total += int(c.item())
    assert total == int(count.item())


def test_relocate_dense_grid_uses_argwhere_fallback():
    # 97% full: most movers miss all 50 probes (grid.py:308-316 fallback)
    rng = np.random.default_rng(3)
    rows, cols = 24, 32
    grid = rng.integers(0, 2, size=(rows, cols)).astype(np.int8)
    empty = rng.choice(rows * cols, size=20, replace=False)
    grid.reshape(-1)[empty] = -1
    st = oracle.SchellingState.from_type_grid(grid)
    dev = DeviceSchelling(st, 0.9, 1)
    st.assign_state(0.9, 1)
    dev.assign()
    total_fb = 0
    for epoch in range(4):
        movers, nfb = st.move_unhappy(5, epoch)
        total_fb += nfb
        got_movers, _ = dev.move(5, epoch)
        assert got_movers == movers
        assert np.array_equal(dev.cells(), st.cell), epoch
        assert st.assign_state(0.9, 1) == dev.assign()
    assert total_fb > 0


@pytest.mark.parametrize("shape,empty_frac,h", [((256, 256), 0.09, 0.9), ((192, 320), 0.12, 1.0)])
def test_relocate_hundreds_of_fallbacks_go_through_the_batched_event_loop(shape, empty_frac, h):
    """A crowded grid where a few hundred movers per step miss all 50 probes (between one batch and the 2048-entry limit of
    the event loop): the time-ordered event loop takes them in batches of up to 32 (csrc/relocate.cu), and every cell must
    still equal the sequential oracle's."""
    rng = np.random.default_rng(11)
    rows, cols = shape
    grid = rng.integers(0, 2, size=shape).astype(np.int8)
    grid.reshape(-1)[rng.choice(rows * cols, size=int(rows * cols * empty_frac), replace=False)] = -1
    st = oracle.SchellingState.from_type_grid(grid)
    dev = DeviceSchelling(st, h, 1)
    st.assign_state(h, 1)
    dev.assign()
    seen = []
    for epoch in range(3):
        movers, nfb = st.move_unhappy(21, epoch)
        seen.append(nfb)
        got_movers, _ = dev.move(21, epoch)
        assert got_movers == movers
        assert np.array_equal(dev.cells(), st.cell), epoch
        assert st.assign_state(h, 1) == dev.assign()
    assert max(seen) > 64 and max(seen) <= 2048, seen


def test_relocate_full_grid_raises_value_error():
    grid = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=np.int8)
    st = oracle.SchellingState.from_type_grid(grid)
    dev = DeviceSchelling(st, 1.0, 1)
    dev.assign()
    with pytest.raises(ValueError):
        dev.move(1, 0)


# ---- the whole step of a small grid in ONE launch (csrc/schelling_small.cu) ------------------------------------------
class DeviceSchellingSmall(DeviceSchelling):
    """The same state stepped by mb_schelling_step_small: relocation + happiness fused, nothing read back per step."""

    def __init__(self, st, homophily, radius, torus=False):
        super().__init__(st, homophily, radius, torus)
        self.small = self.ops.SchellingSmallStep(self.type, st.n, homophily, radius, torus)
        self.empty = torch.ones((self.rows, self.cols), dtype=torch.bool, device="cuda")

    def assign(self):
        self.small.step(self.type, self.happy, self.agent_id.view(torch.int32), 0, 0, agent_cell=self.agent_cell.view(torch.int32),
                        empty_layer=self.empty, relocate=False)
        return self.small.check()[0]

    def step(self, seed, epoch):
        self.small.step(self.type, self.happy, self.agent_id.view(torch.int32), seed, epoch,
                        agent_cell=self.agent_cell.view(torch.int32), empty_layer=self.empty, relocate=True)
        return self.small.check()      # (happy, movers, passes)


@pytest.mark.parametrize("name", SCHELLING)
def test_small_step_matches_reference_trajectory(name):
    g = np.load(os.path.join(GOLDEN, name))
    rows, cols = int(g["width"]), int(g["height"])
    from mesa_b200 import ops

    if not ops.SchellingSmallStep.supports(rows * cols):
        pytest.skip("grid larger than the single-launch path")
    h, r, seed = float(g["homophily"]), int(g["radius"]), int(g["philox_seed"])
    st = oracle.SchellingState(rows, cols, g["cell0"], g["atype"])
    dev = DeviceSchellingSmall(st, h, r)
    assert dev.assign() == int(g["happy_count0"])
    assert np.array_equal(dev.happy_by_agent(), g["happy0"])
    full = set(g["full_steps"].tolist())
    for s in range(1, len(g["happy_count"]) + 1):
        happy, movers, _ = dev.step(seed, s - 1)
        assert movers == int(g["movers"][s - 1]), s
        assert happy == int(g["happy_count"][s - 1]), s
        if s in full:
            assert np.array_equal(dev.cells(), g[f"cell_{s}"]), s
            assert np.array_equal(dev.happy_by_agent(), g[f"happy_{s}"]), s
            tg = dev.type.cpu().numpy().reshape(-1)
            assert np.array_equal(tg[dev.cells()], g["atype"])
            assert np.array_equal(dev.empty.cpu().numpy().reshape(-1), tg < 0)


@pytest.mark.parametrize("shape,density,h,radius,torus,steps", [
    ((24, 32), None, 0.9, 1, False, 5),       # 97 % full: the argwhere fallback dominates
    ((100, 128), 0.8, 0.4, 1, False, 6),      # the largest grid the path takes
    ((37, 23), 0.7, 0.3, 3, True, 6), ((16, 16), 0.5, 0.0, 1, False, 2), ((9, 11), 0.9, 1.5, 2, False, 4),
    ((1, 40), 0.6, 0.5, 1, False, 4), ((50, 3), 0.8, 0.6, 1, True, 4),
])
def test_small_step_matches_oracle(shape, density, h, radius, torus, steps):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    rows, cols = shape
    if density is None:
        grid = rng.integers(0, 2, size=shape).astype(np.int8)
        grid.reshape(-1)[rng.choice(rows * cols, size=20, replace=False)] = -1
    else:
        u = rng.random(shape)
        grid = np.where(u < density, (rng.random(shape) < 0.5).astype(np.int8), np.int8(-1)).astype(np.int8)
    st = oracle.SchellingState.from_type_grid(grid)
    dev = DeviceSchellingSmall(st, h, radius, torus)
    assert st.assign_state(h, radius, torus) == dev.assign()
    fallbacks = 0
    for epoch in range(steps):
        movers, nfb = st.move_unhappy(21, epoch)
        fallbacks += nfb
        want = st.assign_state(h, radius, torus)
        happy, got_movers, _ = dev.step(21, epoch)
        assert (happy, got_movers) == (want, movers), epoch
        assert np.array_equal(dev.cells(), st.cell), epoch
        assert np.array_equal(dev.type.cpu().numpy(), st.type_grid()) and np.array_equal(dev.happy.cpu().numpy(), st.happy_grid())
    if density is None:
        assert fallbacks > 0


def test_small_step_full_grid_raises_value_error():
    grid = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=np.int8)
    st = oracle.SchellingState.from_type_grid(grid)
    dev = DeviceSchellingSmall(st, 1.0, 1)
    dev.assign()
    with pytest.raises(ValueError):
        dev.step(1, 0)


def test_example_model_fused_equals_unfused_and_flushes_deferred_steps():
    """Schelling(fused=True) (one launch per step, lazily read) == Schelling(fused=False) (multi-launch path); a
    shuffle_do("step") observed before its do("assign_state") is flushed."""
    from mesa_b200.examples import Schelling, SchellingScenario

    sc = SchellingScenario(width=50, height=40, density=0.85, homophily=0.5, radius=1, rng=9)
    a, b = Schelling(sc, collect=False, fused=True), Schelling(sc, collect=False, fused=False)
    assert a._small is not None and b._small is None and a.happy == b.happy
    for _ in range(6):
        a.step()
        b.step()
        assert a.happy == b.happy and a.last_movers == b.last_movers and a.running == b.running
        assert torch.equal(a.grid.type, b.grid.type) and torch.equal(a.agents.get("cell"), b.agents.get("cell"))
        assert torch.equal(a.grid.empty, b.grid.empty)
    a.agents.shuffle_do("step")
    b.agents.shuffle_do("step")
    assert torch.equal(a.grid.type, b.grid.type)          # the deferred step is flushed by the layer access
    a.agents.do("assign_state")
    b.agents.do("assign_state")
    assert a.happy == b.happy
    a.run_for(5)
    b.run_for(5)
    assert a.happy == b.happy and torch.equal(a.grid.type, b.grid.type)


def test_relocation_with_the_confirming_iteration_gives_the_same_trajectories():
    """MB_RELOC_CONFIRM=1 re-enables the confirming FULL iteration after the event loop (csrc/relocate.cu: skipped by default
    when the bucket is settled by construction); the golden trajectories must come out the same either way.  The switch is
    read once per process, hence the subprocess."""
    import subprocess
    import sys

    env = dict(os.environ, MB_RELOC_CONFIRM="1", MB_RELOC_DEBUG="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "matches_reference_trajectory or dense_grid or sharded"], env=env, capture_output=True, text=True,
                       timeout=900, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "the confirming full iteration" not in r.stderr          # it never finds work (debug message of relocate.cu)


def test_relocation_in_rank_ordered_buckets_gives_the_same_trajectories():
    """MB_RELOC_BUCKETS=4 forces the path that huge grids take (> 16 M movers per GPU): the movers are processed in four
    rank-ordered buckets through per-bucket index lists (mb_reloc_bucket_lists / mb_reloc_pass_list).  Same goldens, same
    oracle comparisons, single-GPU and row-sharded (1-rank group) drivers."""
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MB_RELOC_BUCKETS="4")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), os.path.join(here, "test_sharded_gpu.py"), "-q", "-x",
                        "-m", "gpu", "-k", "matches_reference_trajectory or dense_grid or sharded_driver"], env=env,
                       capture_output=True, text=True, timeout=900, cwd=os.path.dirname(here))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_run_for_in_one_launch_follows_the_reference_trajectory():
    """model.run_for(k) on the single-launch path = k steps inside ONE launch (nsteps = k): BASELINE config 1 (100 x 100, the
    live reference's golden trajectory) in chunks of 1, 9, 40 and 50 steps, cells and happiness at the golden checkpoints."""
    from mesa_b200.examples import Schelling, SchellingScenario

    g = np.load(os.path.join(GOLDEN, "schelling_100x100_c1.npz"))
    W, H = int(g["width"]), int(g["height"])
    types = np.full(W * H, -1, dtype=np.int8)
    types[g["cell0"]] = g["atype"]
    m = Schelling(SchellingScenario(width=W, height=H, density=0.8, homophily=float(g["homophily"]), radius=int(g["radius"]), rng=1),
                  collect=False, initial_types=types.reshape(W, H))
    m.philox_seed = int(g["philox_seed"])
    assert m._small is not None and m.happy == int(g["happy_count0"])
    done = 0
    for chunk in (1, 9, 40, 50):
        m.run_for(chunk)
        done += chunk
        assert m.time == done and m.steps == done and m._epoch == done
        assert m.happy == int(g["happy_count"][done - 1]), done
        assert m.last_movers == int(g["movers"][done - 1]), done
        if done in set(g["full_steps"].tolist()):
            assert np.array_equal(m.agents.get("cell").cpu().numpy().astype(np.int64), g[f"cell_{done}"]), done
    assert m.running == (m.happy < len(m.agents))


def test_settled_steps_are_skipped_only_while_the_layers_are_untouched():
    """A step without movers returns at once when the happy layer is known to be current; writing into a layer from outside
    (any access through grid.property_layers / grid.<layer>) makes the next step evaluate again."""
    from mesa_b200.examples import Schelling, SchellingScenario

    sc = SchellingScenario(width=30, height=30, density=0.6, homophily=0.3, radius=1, rng=2)
    a, b = Schelling(sc, collect=False, fused=True), Schelling(sc, collect=False, fused=False)
    for _ in range(60):
        a.step()
        b.step()
        if a.last_movers == 0:
            break
    assert a.last_movers == 0 and b.last_movers == 0 and a.happy == b.happy == len(a.agents)
    a.run_for(5)                                     # five settled steps in one launch: nothing changes
    b.run_for(5)
    assert a.happy == b.happy and torch.equal(a.grid.type, b.grid.type)
    for m in (a, b):                                 # an outside write: flip the type of every agent in the first rows
        t = m.grid.type
        t[:6] = torch.where(t[:6] >= 0, 1 - t[:6], t[:6])
    for _ in range(3):
        a.step()
        b.step()
        assert a.happy == b.happy and a.last_movers == b.last_movers
        assert torch.equal(a.grid.type, b.grid.type) and torch.equal(a.grid.happy, b.grid.happy)
    assert a.time == b.time
