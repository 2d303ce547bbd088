"""The drop-in surface (DeviceGrid / DeviceAgentSet / DeviceContinuousSpace / example models) behaves like
the reference's: these read like tests/discrete_space/test_discrete_space.py, tests/test_agentset.py and
tests/experimental/test_continuous_space.py of the reference, with the same known answers."""
import random

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


# --------------------------------------------------------------------------- grids
def test_orthogonal_grid_neumann_known_neighbours():
    from mesa_b200.space import OrthogonalVonNeumannGrid

    grid = OrthogonalVonNeumannGrid((10, 10), torus=False, random=random.Random(42))
    assert len(grid) == 100
    assert set(grid.neighborhood_coordinates((0, 0))) == {(0, 1), (1, 0)}          # test_discrete_space.py:36-60
    assert set(grid.neighborhood_coordinates((9, 9))) == {(9, 8), (8, 9)}
    assert set(grid.neighborhood_coordinates((5, 5))) == {(4, 5), (5, 4), (5, 6), (6, 5)}
    torus = OrthogonalVonNeumannGrid((10, 10), torus=True, random=random.Random(42))
    assert set(torus.neighborhood_coordinates((0, 0))) == {(0, 1), (1, 0), (0, 9), (9, 0)}
    assert set(torus.neighborhood_coordinates((9, 9))) == {(9, 8), (8, 9), (9, 0), (0, 9)}


def test_orthogonal_grid_moore_known_neighbours():
    from mesa_b200.space import OrthogonalMooreGrid

    grid = OrthogonalMooreGrid((10, 10), torus=False, random=random.Random(42))
    assert grid.neighborhood_coordinates((0, 0)) == [(0, 1), (1, 0), (1, 1)]        # connection order
    assert len(grid.neighborhood_coordinates((5, 5))) == 8
    torus = OrthogonalMooreGrid((10, 10), torus=True, random=random.Random(42))
    assert set(torus.neighborhood_coordinates((0, 0))) == {(9, 9), (9, 0), (9, 1), (0, 9), (0, 1), (1, 9), (1, 0), (1, 1)}


def test_neighbourhood_sizes_by_radius():
    # test_discrete_space.py:380-401
    from mesa_b200.space import OrthogonalMooreGrid, OrthogonalVonNeumannGrid

    vn = OrthogonalVonNeumannGrid((10, 10), torus=False, random=random.Random(1))
    moore = OrthogonalMooreGrid((10, 10), torus=False, random=random.Random(1))
    for radius, a, b in [(1, 2, 3), (2, 5, 8), (3, 9, 15)]:
        assert len(vn[(0, 0)].get_neighborhood(radius=radius)) == a
        assert len(moore[(0, 0)].get_neighborhood(radius=radius)) == b
    with pytest.raises(ValueError):
        vn[(0, 0)].get_neighborhood(radius=0)
    # whole-grid sums agree with the per-cell sets
    ones = torch.ones((10, 10), dtype=torch.uint8, device="cuda")
    assert int(vn.neighborhood_sum(ones, radius=2)[0, 0]) == 5
    assert int(moore.neighborhood_sum(ones, radius=3)[0, 0]) == 15
    assert int(moore.neighborhood_sum(ones, radius=2)[5, 5]) == 24
    assert int(vn.neighborhood_sum(ones, radius=2)[5, 5]) == 12


def test_grid_parameter_validation_and_missing_cells():
    from mesa_b200.space import CellMissingException, OrthogonalMooreGrid

    with pytest.raises(ValueError):
        OrthogonalMooreGrid((10, -1), random=random.Random(1))
    with pytest.raises(ValueError):
        OrthogonalMooreGrid((10.0, 10), random=random.Random(1))
    with pytest.raises(TypeError):
        OrthogonalMooreGrid((10, 10), torus="yes", random=random.Random(1))
    with pytest.raises(TypeError):
        OrthogonalMooreGrid((10, 10), capacity="1", random=random.Random(1))
    grid = OrthogonalMooreGrid((4, 5), random=random.Random(1))
    assert grid.width == 4 and grid.height == 5
    with pytest.raises(CellMissingException):
        grid[(4, 0)]
    assert grid.find_nearest_cell((1.7, 2.2)).coordinate == (1, 2)
    with pytest.raises(ValueError):
        grid.find_nearest_cell((9.0, 0.0))
    assert OrthogonalMooreGrid((4, 5), torus=True, random=random.Random(1)).find_nearest_cell((5.5, -0.5)).coordinate == (1, 4)


def test_property_layers():
    # test_discrete_space.py:871-997
    from mesa_b200.space import OrthogonalMooreGrid

    grid = OrthogonalMooreGrid((10, 12), torus=False, random=random.Random(42))
    elevation = grid.create_property_layer("elevation", default_value=0.0)
    assert elevation.shape == (10, 12) and "elevation" in grid.property_layers and "empty" in grid.property_layers
    assert grid.elevation is elevation
    assert grid[(2, 3)].elevation == 0.0
    grid[(2, 3)].elevation = 5.5
    assert float(elevation[2, 3]) == 5.5
    with pytest.raises(ValueError, match="already exists"):
        grid.create_property_layer("elevation")
    with pytest.raises(ValueError, match="conflicts"):
        grid.create_property_layer("width")
    with pytest.raises(ValueError, match="does not match"):
        grid.add_property_layer("bad", np.zeros((3, 3)))
    grid.add_property_layer("water", np.full((10, 12), 2, dtype=np.int32))
    grid.water[:] = torch.clamp(grid.water * 3 - 1, max=4)         # in-place numpy-style expression
    assert int(grid.water.sum()) == 4 * 120
    grid.remove_property_layer("water")
    assert "water" not in grid.property_layers
    with pytest.raises(KeyError):
        grid.remove_property_layer("water")
    ro = grid.create_property_layer("ro", 1.0, read_only=True)
    assert grid[(0, 0)].ro == 1.0 and ro is grid.ro
    with pytest.raises(AttributeError):
        grid[(0, 0)].ro = 3.0
    mask = grid.get_neighborhood_mask((0, 0), include_center=True, radius=1)
    assert int(mask.sum()) == 4 and bool(mask[1, 1]) and not bool(mask[2, 2])


def test_select_random_empty_cell_fallback_and_full_grid():
    # test_discrete_space.py:1164-1210
    from mesa_b200.space import OrthogonalMooreGrid

    grid = OrthogonalMooreGrid((10, 10), random=random.Random(42))
    grid.empty[:] = False
    with pytest.raises(ValueError, match="completely full"):
        grid.select_random_empty_cell()
    grid.empty[7, 3] = True
    assert grid.select_random_empty_cell().coordinate == (7, 3)   # 50 probes almost surely miss -> argwhere fallback
    assert [c.coordinate for c in grid.empties] == [(7, 3)]


# --------------------------------------------------------------------------- continuous space
def test_continuous_space_known_answers():
    from mesa_b200.continuous_space import ContinuousSpace

    dims = np.asarray([[0, 1], [-1, 0]])
    space = ContinuousSpace(dims, torus=False, random=random.Random(1))
    assert space.ndims == 2 and np.all(space.size == [1, 1]) and np.all(space.center == [0.5, -0.5])
    assert space.in_bounds([0.5, -0.5]) and not space.in_bounds([-0.5, -0.5]) and not space.in_bounds([0.5, 0.5])
    torus = ContinuousSpace(dims, torus=True, random=random.Random(1))
    for p in ([-0.5, 0.5], [0.5, -0.5], [1.5, -0.5], [0.5, -1.5]):
        assert np.all(torus.torus_correct(p) == [0.5, -0.5])
    with pytest.raises(ValueError):
        space.add_agents([[1.1, 1.1]])

    space = ContinuousSpace([[0, 1], [0, 1]], torus=False, random=random.Random(1))
    space.add_agents([[0.1, 0.1]])
    d, ids = space.calculate_distances([0.1, 0.9])
    assert np.allclose(d, [0.8]) and ids == [0]
    torus = ContinuousSpace([[0, 1], [0, 1]], torus=True, random=random.Random(1))
    torus.add_agents([[0.1, 0.1]])
    assert np.allclose(torus.calculate_distances([0.1, 0.9])[0], [0.2])
    assert np.allclose(torus.calculate_distances([0.9, 0.9])[0], [0.2 * 2**0.5])
    assert np.allclose(torus.calculate_difference_vector([0.9, 0.9]), [[0.2, 0.2]], atol=1e-6)

    space = ContinuousSpace([[0, 1], [0, 1]], torus=False, random=random.Random(1))
    idx = space.add_agents([[0.1, 0.1], [0.1, 0.9], [0.9, 0.1], [0.9, 0.9], [0.5, 0.5]])
    assert idx.tolist() == [0, 1, 2, 3, 4]
    agents, dist = space.get_agents_in_radius([0.1, 0.1], radius=0.1)
    assert agents == [0] and np.allclose(dist, [0])
    agents, dist = space.get_agents_in_radius([0.5, 0.1], radius=0.4)     # inclusive boundary
    assert agents == [0, 2, 4] and np.allclose(dist, [0.4, 0.4, 0.4])
    assert len(space.get_agents_in_radius([0.5, 0.5], radius=1)[0]) == 5
    near, nd = space.get_k_nearest_agents([0.45, 0.45], k=2)
    assert near[0] == 4 and len(near) == 2
    with pytest.warns(UserWarning):
        assert len(space.get_k_nearest_agents([0.5, 0.5], k=9)[0]) == 5
    # growth keeps earlier rows (continuous_space.py:121-129), swap-with-last removal (:141-167)
    small = ContinuousSpace([[0, 1], [0, 1]], torus=True, random=random.Random(1), n_agents=2)
    small.add_agents(np.full((5, 2), 0.25))
    small.add_agents([[0.75, 0.75]])
    assert small.agent_positions.shape == (6, 2) and float(small.agent_positions[5, 0]) == 0.75
    assert small.remove_agent(0) == 5 and float(small.agent_positions[0, 0]) == 0.75 and small.agent_positions.shape[0] == 5
    off, nbr, _ = small.neighbors_in_radius(0.1)
    assert np.diff(off.cpu().numpy()).tolist() == [0, 3, 3, 3, 3]


# --------------------------------------------------------------------------- agent sets
def test_agentset_contracts():
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet, batch_operator
    from mesa_b200.model import DeviceModel

    class Walker(DeviceAgent):
        pass

    calls = []

    @batch_operator(Walker, "step")
    def _step(agents, epoch=None, by=1, **_):
        calls.append(epoch)
        agents.columns["x"] += by

    model = DeviceModel(rng=3)
    cols = {"x": torch.arange(10, device="cuda", dtype=torch.int32)}
    agents = DeviceAgentSet(model, Walker, 10, cols)
    assert len(agents) == 10 and agents[0].unique_id == 1 and agents[-1].x == 9
    assert agents.do("step") is agents and agents.shuffle_do("step", by=2) is agents     # test_agentset.py:776-840
    assert calls == [None, 0] and agents.get("x").tolist() == list(range(3, 13))
    agents.shuffle_do("step")
    assert calls[-1] == 1                                             # every shuffle is a fresh epoch
    with pytest.raises(NotImplementedError):
        agents.do("fly")                                              # no per-agent CPU fallback
    with pytest.raises(NotImplementedError):
        agents.do(lambda a: None)
    assert agents.agg("x", "max") == 13 and agents.agg("x", torch.sum).item() == sum(range(4, 14))
    sel = agents.select(lambda c: c["x"] % 2 == 0)
    assert len(sel) == 5 and all(a.x % 2 == 0 for a in sel) and sel[0] in agents
    agents.set("x", 7)
    assert agents.get("x").tolist() == [7] * 10
    shuffled = agents.shuffle()
    assert sorted(a.unique_id for a in shuffled) == list(range(1, 11))
    from mesa_b200 import rng

    want = oracle.activation_order(model.philox_seed, model._epoch - 1, 10)
    assert [a.unique_id - 1 for a in shuffled] == want.tolist()       # host priorities == device/oracle protocol
    assert [a.unique_id for a in agents.sort("x", ascending=True)] == list(range(1, 11))


# --------------------------------------------------------------------------- example models end to end
def test_example_game_of_life_matches_oracle():
    from mesa_b200.examples import ConwaysGameOfLife

    m = ConwaysGameOfLife(width=96, height=128, rng=1)
    s = m.grid.state.cpu().numpy().copy()
    assert 0.1 < s.mean() < 0.3 and len(m.agents) == 96 * 128
    m.run_for(5)
    for _ in range(5):
        s = oracle.gol_step(s, True)
    assert np.array_equal(m.grid.state.cpu().numpy(), s) and m.steps == 5
    assert m.agents[5].state == s.reshape(-1)[5] and m.agents[5].is_alive == bool(s.reshape(-1)[5])


def test_example_schelling_matches_oracle():
    from mesa_b200.examples import Schelling, SchellingScenario

    m = Schelling(SchellingScenario(width=60, height=40, density=0.8, homophily=0.5, radius=1, rng=11))
    st = oracle.SchellingState.from_type_grid(m.grid.type.cpu().numpy())
    assert st.assign_state(0.5, 1) == m.happy and len(m.agents) == st.n
    for step in range(8):
        m.step()
        st.move_unhappy(m.philox_seed, step)
        assert st.assign_state(0.5, 1) == m.happy
        assert np.array_equal(m.agents.get("cell").cpu().numpy(), st.cell)
        assert np.array_equal(m.agents.get("happy").cpu().numpy(), st.happy)
        assert np.array_equal(m.grid.empty.cpu().numpy(), st.type_grid() < 0)
    a = m.agents[3]
    assert a.cell.coordinate == (st.cell[3] // 40, st.cell[3] % 40) and a.type == st.atype[3]


def test_example_boltzmann_matches_oracle():
    from mesa_b200.examples import BoltzmannScenario, BoltzmannWealth

    m = BoltzmannWealth(BoltzmannScenario(n=3000, width=30, height=20, rng=5))
    st = oracle.BoltzmannState(30, 20, m.agents.get("cell").cpu().numpy())
    for step in range(6):
        m.step()
        st.step(m.philox_seed, step)
    assert np.array_equal(m.agents.get("cell").cpu().numpy(), st.cell)
    assert np.array_equal(m.agents.get("wealth").cpu().numpy(), st.wealth)
    x = np.sort(st.wealth)
    n = st.n
    gini = 1 + 1 / n - 2 * sum(xi * (n - i) for i, xi in enumerate(x)) / (n * x.sum())
    assert abs(m.gini - gini) < 1e-12


def test_example_boids_matches_oracle():
    from mesa_b200.examples import BoidFlockers, BoidsScenario

    sc = BoidsScenario(population_size=500, width=120, height=90, vision=10.0, rng=2)
    m = BoidFlockers(sc, activation="synchronous")
    pos = m.agents.get("position").cpu().numpy().copy()
    dirs = m.agents.get("direction").cpu().numpy().copy()
    for _ in range(3):
        m.step()
        wp, wd, _ = oracle.boids_step(pos, dirs, m.space.size, vision=10.0, separation=2.0)
        # the synchronous model keeps its storage spatially sorted (resort_every): rows are found through the unique ids
        by_id = np.argsort(m.agents.unique_ids().cpu().numpy())
        got_p, got_d = m.agents.get("position").cpu().numpy()[by_id], m.agents.get("direction").cpu().numpy()[by_id]
        assert np.max(np.abs(got_p - wp) / m.space.size) <= 1e-5 and np.max(np.abs(got_d - wd)) <= 1e-5
        pos, dirs = got_p.copy(), got_d.copy()
    assert m.resort_every == 8 and not np.array_equal(by_id, np.arange(500))       # the first step re-sorted the rows
    assert m.agent_angles.shape == (500,) and -np.pi <= m.average_heading <= np.pi


def test_example_boids_sequential_activation_matches_oracle():
    """BoidFlockers(activation="sequential") == the oracle's sequential shuffle_do in the model's Philox order."""
    from mesa_b200.examples import BoidFlockers, BoidsScenario

    sc = BoidsScenario(population_size=400, width=100, height=80, vision=10.0, rng=3)
    m = BoidFlockers(sc, activation="sequential")
    pos = m.agents.get("position").cpu().numpy().copy()
    dirs = m.agents.get("direction").cpu().numpy().copy()
    for epoch in range(3):
        m.step()
        order = oracle.activation_order(m.philox_seed, epoch, 400)
        wp, wd, _ = oracle.boids_step(pos, dirs, m.space.size, vision=10.0, separation=2.0, order=order)
        got_p, got_d = m.agents.get("position").cpu().numpy(), m.agents.get("direction").cpu().numpy()
        assert np.max(np.abs(got_p - wp) / m.space.size) <= 1e-5 and np.max(np.abs(got_d - wd)) <= 1e-5
        assert m.last_passes >= 1
        pos, dirs = got_p.copy(), got_d.copy()


def test_capacity_api_matches_reference_draw_protocol():
    """cells_with_capacity / select_random_cell_with_capacity / Cell.is_full (grid.py:318-378, cell.py:178-182) on a
    grid whose cells hold up to 2 agents: the same cells, the same consumption of the stdlib random stream
    (<= 50 `choice(cells)` probes, then one `choice(available)`), IndexError when every cell is full."""
    from mesa_b200.space import DeviceMooreGrid

    grid = DeviceMooreGrid((6, 7), torus=False, capacity=2, random=random.Random(5))
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 3, size=(6, 7)).astype(np.int32)
    grid.add_property_layer("agent_count", counts)
    room = counts < 2
    got = sorted(c.coordinate for c in grid.cells_with_capacity)
    assert got == sorted(map(tuple, np.argwhere(room).tolist()))
    assert grid[(0, 0)].is_full == bool(counts[0, 0] >= 2)
    replay = random.Random(5)
    for _ in range(20):
        want = None
        for _ in range(50):
            k = replay._randbelow(42)
            if room.reshape(-1)[k]:
                want = (k // 7, k % 7)
                break
        assert want is not None and grid.select_random_cell_with_capacity().coordinate == want
    # nearly full: only (5, 6) has room -> the 50 probes usually miss and the fallback draws among 1 cell
    counts[:] = 2
    counts[5, 6] = 1
    grid.property_layers["agent_count"].copy_(torch.from_numpy(counts))
    for _ in range(5):
        assert grid.select_random_cell_with_capacity().coordinate == (5, 6)
    grid.property_layers["agent_count"].fill_(2)
    assert list(grid.cells_with_capacity) == []
    with pytest.raises(IndexError):
        grid.select_random_cell_with_capacity()
    # capacity None: nothing is ever full, one draw per call; capacity 1 without a count layer: 1 - empty
    free = DeviceMooreGrid((3, 3), random=random.Random(1))
    assert not free[(1, 1)].is_full and len(list(free.cells_with_capacity)) == 9
    assert free.select_random_cell_with_capacity().coordinate == divmod(random.Random(1)._randbelow(9), 3)
    single = DeviceMooreGrid((2, 2), capacity=1, random=random.Random(1))
    single.empty[0, 1] = False
    assert single[(0, 1)].is_full and sorted(c.coordinate for c in single.cells_with_capacity) == [(0, 0), (1, 0), (1, 1)]


def test_agents_by_type_typed_activation():
    """model.agents_by_type / agent_types (model.py:221-247): typed activation like wolf_sheep/model.py:140-141."""
    from mesa_b200.examples import Schelling, SchellingScenario
    from mesa_b200.examples.schelling import SchellingAgent

    m = Schelling(SchellingScenario(width=20, height=20, density=0.8, homophily=0.4, radius=1, rng=3), collect=False)
    assert m.agent_types == [SchellingAgent] and m.agents_by_type[SchellingAgent] is m.agents
    before = m.happy
    assert m.agents_by_type[SchellingAgent].do("assign_state") is m.agents and m.happy == before
    m.agents_by_type[SchellingAgent].shuffle_do("step")
    m.agents.do("assign_state")
    assert m.happy != before or m.last_movers == 0


def test_cell_agents_through_the_model_hook():
    """cell.py:192-195 `Cell.agents`, cell_collection.py:97-136 `CellCollection.agents / select_random_agent`: which agent
    sits in a cell is model state on the device, served through `grid.cell_agents`."""
    from mesa_b200.examples import ConwaysGameOfLife, Schelling, SchellingScenario

    m = Schelling(SchellingScenario(width=12, height=9, density=0.7, homophily=0.4, radius=1, rng=3))
    types = m.grid.type.cpu().numpy()
    cells = m.agents.get("cell").cpu().numpy()
    for a in list(m.agents)[:20]:
        assert [b.unique_id for b in a.cell.agents] == [a.unique_id]     # capacity 1: the cell's agent is the agent
    empty = [c for c in m.grid.all_cells if types[c.coordinate] < 0]
    assert empty and empty[0].agents == []
    nb = m.grid[(5, 4)].get_neighborhood()
    want = [int(np.nonzero(cells == c.flat)[0][0]) + 1 for c in nb if types[c.coordinate] >= 0]
    assert [a.unique_id for a in nb.agents] == want
    if want:
        assert nb.select_random_agent().unique_id in want
    g = ConwaysGameOfLife(width=8, height=32, rng=1)
    assert [a.unique_id for a in g.grid[(2, 3)].agents] == [2 * 32 + 3 + 1]
    assert len(g.grid[(0, 0)].get_neighborhood().agents) == 8


def test_host_mirror_of_a_running_model():
    """SURVEY 8(f) next-3 on the device: the host snapshot of a stepped Schelling model holds the layers as numpy arrays and
    one agent view per agent whose cell / position agree with the device columns; `grid.agents` lists them cell by cell."""
    import numpy as np

    from mesa_b200.examples import Schelling, SchellingScenario

    m = Schelling(SchellingScenario(width=20, height=15, density=0.7, rng=3))
    for _ in range(3):
        m.step()
    space = m.grid.host_mirror()
    cells = m.agents.get("cell").cpu().numpy().astype(np.int64)
    types = m.agents.get("type").cpu().numpy()
    assert len(space.agents) == len(m.agents) == len(cells)
    for a, c, t in zip(space.agents, cells, types):
        assert a.cell.coordinate == tuple(int(x) for x in np.unravel_index(c, (20, 15))) and a.type == t
        assert np.array_equal(a.cell.position, np.asarray(a.cell.coordinate, dtype=float))
    assert np.array_equal(space.property_layers["type"], m.grid.type.cpu().numpy())
    assert space.property_layers["empty"].dtype == bool
    assert np.array_equal(space.property_layers["empty"], m.grid.type.cpu().numpy() < 0)
    flat = [a.cell.coordinate[0] * 15 + a.cell.coordinate[1] for a in m.grid.agents]
    assert flat == sorted(flat) and len(flat) == len(cells)
