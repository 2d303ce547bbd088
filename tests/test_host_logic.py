"""Host-side logic of the drop-in classes that needs no GPU: neighbour ordering of the n-dimensional grids against the
reference's Cell.get_neighborhood (golden probes), chunking of the host stepper, cell-edge heuristic of the k-nearest
search, DeviceModel bookkeeping (clock, epochs, agents_by_type)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _bare_grid(cls, dims, torus):
    """A device grid without its device layers: only the fields neighborhood_coordinates reads."""
    g = cls.__new__(cls)
    g.__dict__.update(dimensions=tuple(dims), torus=bool(torus), _ndims=len(dims))
    if len(dims) != 2:
        g.__dict__["_offsets"] = cls._offsets_nd(len(dims))
    return g


def test_nd_neighbour_order_matches_reference_probes():
    from mesa_b200.space import DeviceMooreGrid, DeviceVonNeumannGrid

    g = np.load(os.path.join(GOLDEN, "neighborhoods_nd.npz"))
    for k, (moore, torus, radius, inc) in enumerate(g["cases"]):
        dims = tuple(int(d) for d in g[f"dims_{k}"])
        grid = _bare_grid(DeviceMooreGrid if moore else DeviceVonNeumannGrid, dims, torus)
        probes = [tuple(0 for _ in dims), tuple(d - 1 for d in dims), tuple(d // 2 for d in dims)]
        for q, pc in enumerate(probes):
            want = [tuple(int(x) for x in row) for row in g[f"probe_{k}_{q}"]]
            assert grid.neighborhood_coordinates(pc, int(radius), bool(inc)) == want, (k, dims, pc)


def test_2d_neighbour_order_and_torus_dedup():
    """connection order at a clamped corner, wrapped duplicates collapse, a dimension of 1 makes the
    cell its own neighbour on a torus."""
    from mesa_b200.space import DeviceMooreGrid, DeviceVonNeumannGrid

    assert _bare_grid(DeviceMooreGrid, (5, 5), False).neighborhood_coordinates((0, 0)) == [(0, 1), (1, 0), (1, 1)]
    assert _bare_grid(DeviceVonNeumannGrid, (5, 5), False).neighborhood_coordinates((2, 2)) == [(1, 2), (2, 1), (2, 3), (3, 2)]
    assert len(_bare_grid(DeviceMooreGrid, (2, 5), True).neighborhood_coordinates((0, 0))) == 5
    assert len(_bare_grid(DeviceMooreGrid, (2, 2), True).neighborhood_coordinates((0, 0))) == 3
    assert (0, 1) in _bare_grid(DeviceMooreGrid, (1, 4), True).neighborhood_coordinates((0, 1))
    assert len(_bare_grid(DeviceMooreGrid, (7, 7), True).neighborhood_coordinates((3, 3), radius=5)) == 48
    assert len(_bare_grid(DeviceVonNeumannGrid, (9, 9), False).neighborhood_coordinates((4, 4), radius=2)) == 12


def test_host_stepper_chunk_bounds_tile_the_rows():
    from mesa_b200.host_api import GolHostStepper

    for rows in (8, 37, 64, 100, 1000, 4096, 16384):
        for chunks in (1, 3, 7, 16, 64):
            for ramp in (False, True):
                b = GolHostStepper.chunk_bounds(rows, chunks, ramp)
                assert b[0][0] == 0 and b[-1][1] == rows
                assert all(hi > lo for lo, hi in b) and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    ramped = GolHostStepper.chunk_bounds(16384, 16, True)
    assert len(ramped) == 22 and ramped[0] == (0, 128) and ramped[-1] == (16384 - 128, 16384)


def test_knn_cell_edge_is_bounded_and_scales():
    from mesa_b200 import ops

    e1 = ops.knn_cell_edge(10_000, (100.0, 100.0), 4)
    assert 0 < e1 <= 100.0 and abs(e1 - (100.0 * 100.0 * 5 / (9 * 10_000)) ** 0.5) < 1e-12
    assert ops.knn_cell_edge(0, (10.0, 5.0), 1) <= 10.0            # empty space: one cell at most
    assert ops.knn_cell_edge(10**9, (1.0, 1.0), 1) >= 1e-6          # never degenerate
    assert ops.knn_cell_edge(100, (10.0, 10.0), 16) > ops.knn_cell_edge(100, (10.0, 10.0), 1)


def test_device_model_clock_epochs_and_agents_by_type():
    from mesa_b200.model import DeviceModel

    class Dummy:
        pass

    class Set:
        agent_type = Dummy

    class M(DeviceModel):
        def __init__(self):
            super().__init__(rng=7)
            self.calls = 0
            self.agents = Set()

        def step(self):
            self.calls += 1

    m = M()
    m.step()
    m.run_for(3)
    assert (m.steps, m.time, m.calls) == (4, 4.0, 4)                 # model.py:152-154: every step advances time by 1
    assert [m._next_epoch() for _ in range(3)] == [0, 1, 2]          # one epoch per shuffle
    assert m.agent_types == [Dummy] and m.agents_by_type[Dummy] is m.agents
    m2 = M()
    assert m2.philox_seed == m.philox_seed and m2.random.random() == M().random.random()   # seeded identically


def _host_sort_pairs(keys, vals, begin_bit=0, end_bit=None):
    """Test double for ops.sort_pairs (the stable radix sort of csrc/sort.cu, keys compared as UNSIGNED integers)."""
    import torch

    wide = keys.dtype == torch.int64
    order = torch.argsort(keys ^ (-2**63 if wide else -2**31), stable=True)
    return keys[order], vals[order]


class _HostPlan:
    """Test double for ops.CompactionPlan (the device compaction of csrc/agents.cu): lets the bookkeeping of
    DeviceAgentSet (unique ids, order, generations, hooks) run on CPU tensors.  Not a product path."""

    def __init__(self, keep):
        self.keep = keep.to(bool)
        self.n = keep.numel()
        self.kept = int(self.keep.sum())

    def apply(self, col):
        return col[self.keep].clone()


def test_dynamic_population_bookkeeping_matches_the_reference_script(monkeypatch):
    """tests/golden/dynamic_agents.npz: Agent.create_agents / Agent.remove on the live Model (agent.py:73-160)."""
    from _dynamic_script import replay
    from mesa_b200 import ops

    monkeypatch.setattr(ops, "CompactionPlan", _HostPlan)
    for by_mask in (True, False):
        model, agents = replay("cpu", by_mask)
    # numpy masks and numpy index arrays mean what the torch ones mean
    n = len(agents)
    ids = agents.unique_ids().tolist()
    mask = np.zeros(n, dtype=bool)
    mask[[0, n - 1]] = True
    agents.remove_agents(mask)
    assert agents.unique_ids().tolist() == ids[1:-1]
    agents.remove_agents(np.array([1, 2]))
    assert agents.unique_ids().tolist() == [ids[1]] + ids[4:-1]


def test_population_growth_needs_no_kernel_and_removal_has_no_cpu_fallback(monkeypatch):
    import pytest
    import torch

    from mesa_b200 import ops
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class Wolf(DeviceAgent):
        pass

    model = DeviceModel(rng=1)
    agents = DeviceAgentSet(model, Wolf, 3, {"energy": torch.tensor([4, 5, 6]), "pos": torch.zeros((3, 2))})
    model.agents = agents
    seen = []
    agents.resize_hooks.append(lambda s, plan: seen.append((len(s), plan)))
    cubs = agents.add_agents(2, energy=7, pos=torch.tensor([1.0, 2.0]))       # one shared value per column
    assert len(agents) == 5 and cubs.unique_ids().tolist() == [4, 5] and seen == [(5, None)]
    assert agents.get("energy").tolist() == [4, 5, 6, 7, 7] and agents.get("pos")[3:].tolist() == [[1.0, 2.0]] * 2
    assert agents.add_agents(1).get("energy").tolist() == [0]                   # unnamed columns start at zero
    with pytest.raises(AttributeError):
        agents.add_agents(1, speed=3)
    with pytest.raises(ValueError):
        agents.add_agents(-1)
    assert len(agents) == 6                                                     # failed calls change nothing
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        agents.remove(agents[0])                                                # the compaction is a device kernel
    assert len(agents) == 6 and agents.unique_ids().tolist() == [1, 2, 3, 4, 5, 6]
    fixed = DeviceAgentSet(model, Wolf, 3, {"energy": torch.tensor([4, 5, 6])}, resizable=False)
    with pytest.raises(NotImplementedError):
        fixed.add_agents(1)
    monkeypatch.setattr(ops, "sort_pairs", _host_sort_pairs)                    # shuffle() orders the view with the device sort
    with pytest.raises(NotImplementedError):
        agents.shuffle().remove_agents([0])                                     # views do not resize the population


def test_groupby_bookkeeping_on_the_host(monkeypatch):
    """Group order = first appearance, set order inside a group (agentset.py:305-313); the stable device sort is replaced
    by a host double here, the device test is tests/test_agents_gpu.py::test_groupby_matches_the_reference_grouping."""
    import torch

    from mesa_b200 import ops
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    monkeypatch.setattr(ops, "sort_pairs", _host_sort_pairs)
    monkeypatch.setattr(ops, "CompactionPlan", _HostPlan)

    class Citizen(DeviceAgent):
        pass

    model = DeviceModel(rng=1)
    state = torch.tensor([2, 0, 2, 1, 0, 2, 1], dtype=torch.int32)
    agents = DeviceAgentSet(model, Citizen, 7, {"state": state, "x": torch.arange(7.0)})
    g = agents.groupby("state")
    assert list(g.groups) == [2, 0, 1] and g.count() == {2: 3, 0: 2, 1: 2}
    assert g.get_group(2).unique_ids().tolist() == [1, 3, 6] and g.get_group(1).get("x").tolist() == [3.0, 6.0]
    assert g.agg("x", torch.sum) == {2: torch.tensor(7.0), 0: torch.tensor(5.0), 1: torch.tensor(9.0)}
    shuffled = agents.sort("x", ascending=False)                     # a reordered view groups in ITS order
    g2 = shuffled.groupby("state", result_type="list")
    assert list(g2.groups) == [1, 2, 0] and [a.unique_id for a in g2.get_group(2)] == [6, 3, 1]


def test_cell_collection_matches_the_reference(monkeypatch):
    """tests/golden/cell_collection.npz: CellCollection of a neighbourhood on the live grids -- ordered cells, the draws
    of select_random_cell under random.Random(123) (cell_collection.py:101-119), select(filter / at_most) (:138-175)."""
    import random

    import pytest

    from mesa_b200.space import CellCollectionProxy, DeviceMooreGrid, DeviceVonNeumannGrid

    g = np.load(os.path.join(GOLDEN, "cell_collection.npz"))
    grids = {}
    for k, (moore, torus, radius, pi, pj) in enumerate(g["cases"]):
        key = (int(moore), int(torus), int(radius))
        if key not in grids:                       # the reference drew all probes of a grid from ONE generator
            grid = _bare_grid(DeviceMooreGrid if moore else DeviceVonNeumannGrid, (7, 5), torus)
            grid.__dict__["random"] = random.Random(123)
            grids[key] = grid
        nb = grids[key][(int(pi), int(pj))].get_neighborhood(radius=int(radius))
        assert isinstance(nb, CellCollectionProxy) and [c.coordinate for c in nb.cells] == [tuple(r) for r in g[f"cells_{k}"]]
        assert [nb.select_random_cell().coordinate for _ in range(6)] == [tuple(r) for r in g[f"draws_{k}"]]
        even = nb.select(lambda c: (c.coordinate[0] + c.coordinate[1]) % 2 == 0)
        assert [c.coordinate for c in even] == [tuple(r) for r in g[f"even_{k}"]]
        assert [c.coordinate for c in nb.select(at_most=3)] == [tuple(r) for r in g[f"first3_{k}"]]
        assert [c.coordinate for c in nb.select(at_most=0.5)] == [tuple(r) for r in g[f"half_{k}"]]
        assert nb.select() is nb and nb == nb.cells                       # it is the list of its cells
    empty = CellCollectionProxy([], random.Random(1))
    assert empty.select_random_cell(default=None) is None and empty.agents == []
    with pytest.raises(LookupError):
        empty.select_random_cell()
    with pytest.raises(LookupError):
        empty.select_random_agent()
    with pytest.raises(NotImplementedError):
        nb.agents                                                        # no model registered grid.cell_agents


def test_model_clock_run_for_run_until_like_the_reference():
    """Probed on the live Model (mesa/model.py:156-190, 395-422): steps fall due at the integer times; run_for(2.5)
    runs 2 steps and leaves time 2.5; run_until in the past warns and does nothing; step() itself adds one unit."""
    import pytest

    from mesa_b200.model import DeviceModel

    class M(DeviceModel):
        def __init__(self):
            super().__init__(rng=1)
            self.n = 0
            self.seen = []

        def step(self):
            self.n += 1
            self.seen.append(self.time)

    m = M()
    m.run_for(2.5)
    assert (m.n, m.time) == (2, 2.5)
    m.run_for(0.5)
    assert (m.n, m.time) == (3, 3.0)
    m.run_until(5)
    assert (m.n, m.time) == (5, 5.0)
    with pytest.warns(RuntimeWarning):
        m.run_until(3)
    assert (m.n, m.time) == (5, 5.0)
    m.step()
    assert (m.n, m.time) == (6, 6.0)
    m.run_until(7.2)
    assert (m.n, m.time) == (7, 7.2)
    m.run_for(1)
    assert (m.n, m.time) == (8, 8.2) and m.seen == [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0] and m.steps == 8


def test_select_at_most_semantics_of_the_reference():
    """agentset.py:95-117 (tests/test_agentset.py::test_agentset select cases): an int caps the count, a float <= 1.0
    is a fraction of the ORIGINAL set rounded down -- also after filtering -- and 1.0 keeps everything."""
    import torch

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class A(DeviceAgent):
        pass

    agents = DeviceAgentSet(DeviceModel(rng=1), A, 10, {"x": torch.arange(10)})
    assert len(agents.select(at_most=3)) == 3 and len(agents.select(at_most=0.25)) == 2
    assert len(agents.select(at_most=1.0)) == 10 and len(agents.select(at_most=1)) == 1
    even = lambda c: c["x"] % 2 == 0                                               # noqa: E731
    assert agents.select(even, at_most=0.3).get("x").tolist() == [0, 2, 4]          # 30 % of 10, not of the 5 matches
    assert agents.select(even, at_most=0.9).get("x").tolist() == [0, 2, 4, 6, 8]
    assert agents.select(even, at_most=2).get("x").tolist() == [0, 2]
    assert len(agents.select(agent_type=DeviceModel)) == 0 and len(agents.select(agent_type=DeviceAgent)) == 10


def test_set_creates_missing_columns_like_setattr():
    """agentset.py:215-227: `set` does setattr on every agent, so a new name becomes a new attribute (column)."""
    import torch

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class A(DeviceAgent):
        pass

    agents = DeviceAgentSet(DeviceModel(rng=1), A, 4, {"x": torch.arange(4)})
    assert agents.set("x", 7) is agents and agents.get("x").tolist() == [7] * 4
    agents.set("mood", 2.5)
    assert agents.get("mood").tolist() == [2.5] * 4 and agents[1].mood == 2.5
    agents.set("velocity", [1.0, -1.0])                                   # one shared vector value
    assert agents.get("velocity").tolist() == [[1.0, -1.0]] * 4
    agents.select(lambda c: c["x"] == 7, at_most=2).set("flag", True)     # through a view: the others stay zero
    assert agents.get("flag").tolist() == [True, True, False, False]
    assert agents.add_agents(1, x=3).get("mood").tolist() == [0.0] and len(agents.get("velocity")) == 5


def test_get_orientation_and_missing_handling(host_kernels):
    """agentset.py:171-213 (tests/test_agentset.py::test_agentset_get_attribute): several names give one row per agent."""
    import pytest
    import torch

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class A(DeviceAgent):
        pass

    agents = DeviceAgentSet(DeviceModel(rng=1), A, 3, {"a": torch.tensor([1, 2, 3]), "b": torch.tensor([0.5, 1.5, 2.5]),
                                                       "pos": torch.zeros((3, 2))})
    assert agents.get("a").tolist() == [1, 2, 3]
    assert agents.get(["a", "b"]).tolist() == [[1.0, 0.5], [2.0, 1.5], [3.0, 2.5]]
    assert agents.get(["a", "nope"], handle_missing="default", default_value=-1).tolist() == [[1, -1], [2, -1], [3, -1]]
    with pytest.raises(AttributeError):
        agents.get("nope")
    with pytest.raises(AttributeError):
        agents.get(["a", "nope"])
    with pytest.raises(ValueError):
        agents.get("a", handle_missing="ignore")
    with pytest.raises(ValueError):
        agents.get(["a", "pos"])
    assert agents.sort("b", ascending=False).get(["a"]).tolist() == [[3], [2], [1]]


def test_agg_maps_well_known_functions_to_device_reductions(monkeypatch):
    """tests/test_agentset.py::test_agentset_agg of the reference passes min / max / sum / np.mean / lists of them; those
    must not iterate a device tensor: they dispatch to the named device reduction (host double here)."""
    import torch

    from mesa_b200 import ops
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    calls = []

    def host_reduce(layer, op="sum"):
        calls.append(op)
        fn = {"sum": torch.sum, "min": torch.min, "max": torch.max, "count_nonzero": torch.count_nonzero}[op]
        return fn(layer).reshape(1)

    monkeypatch.setattr(ops, "layer_reduce", host_reduce)

    class A(DeviceAgent):
        pass

    agents = DeviceAgentSet(DeviceModel(rng=1), A, 10, {"energy": torch.arange(1, 11, dtype=torch.int32),
                                                        "wealth": 10 * torch.arange(1, 11, dtype=torch.int32)})
    assert agents.agg("energy", min) == 1 and agents.agg("energy", max) == 10 and agents.agg("energy", sum) == 55
    assert agents.agg("wealth", np.mean) == 55.0 and agents.agg("energy", [min, max]) == [1, 10]
    assert agents.agg("energy", (min, max, sum)) == [1, 10, 55] and agents.agg("energy", len) == 10
    assert calls == ["min", "max", "sum", "sum", "min", "max", "min", "max", "sum"]

    def custom(values):                                               # any other callable sees the column
        return float(values.sum()) / len(values)

    assert agents.agg("energy", custom) == 5.5 and agents.agg("wealth", [min, custom]) == [10, 55.0]
    assert agents.select(lambda c: c["energy"] > 100).agg("energy", sum) == 0


class _PickledAgent:
    pass


def test_agent_proxies_and_views_pickle():
    """agentset.py:617-632 / test_agentset.py::test_agentset_serialization: agents and sets survive pickle."""
    import pickle

    import torch

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    global _PickledAgent

    class _PickledAgent(DeviceAgent):  # noqa: F811 - module-level name so that pickle can find the class
        pass

    _PickledAgent.__qualname__ = "_PickledAgent"
    agents = DeviceAgentSet(DeviceModel(rng=1), _PickledAgent, 5, {"x": torch.arange(5) * 2})
    a, view = agents[3], agents.select(lambda c: c["x"] > 2)
    b, view2, agents2 = pickle.loads(pickle.dumps((a, view, agents)))
    assert b.unique_id == 4 and b.x == 6 and b in agents2 and view2.columns is agents2.columns
    assert [p.unique_id for p in view2] == [3, 4, 5] and len(agents2) == 5


def test_host_mirror_serves_the_visualization_consumers(host_kernels):
    """what visualization/backends/matplotlib_backend.py:104-163, 340-356 and space_renderer.py:339 read
    from a space -- `for agent in space.agents`, `agent.cell.position` (abstract_renderer.py:38-41), `space.property_layers`
    as numpy arrays (`layer.astype(float)` for bool layers, `(space.width, space.height) == data.shape`, np.min / np.max) --
    from a host snapshot of a device grid (CPU tensors here)."""
    import torch

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel
    from mesa_b200.space import DeviceMooreGrid, HostSpaceMirror

    class Ant(DeviceAgent):
        pass

    model = DeviceModel(rng=3)
    grid = DeviceMooreGrid((4, 6), torus=False, random=model.random, device="cpu")
    grid.create_property_layer("sugar", default_value=1.5)
    grid.sugar[2, 3] = 9.0
    cells = torch.tensor([23, 0, 7, 7], dtype=torch.int32)                  # flat ids: (3, 5), (0, 0), (1, 1), (1, 1)
    ants = DeviceAgentSet(model, Ant, 4, {"cell": cells, "load": torch.tensor([0.5, 1.5, 2.5, 3.5])})
    grid.attach_population(ants)
    grid.empty.view(-1)[cells.long()] = False

    space = grid.host_mirror()
    assert isinstance(space, HostSpaceMirror) and (space.width, space.height) == (4, 6)
    # property layers: numpy, the renderer's own expressions work on them
    assert set(space.property_layers) == {"empty", "sugar"}
    for name, layer in space.property_layers.items():
        data = layer.astype(float) if layer.dtype == bool else layer
        assert isinstance(data, np.ndarray) and (space.width, space.height) == data.shape
    assert np.max(space.property_layers["sugar"]) == 9.0 and np.min(space.property_layers["sugar"]) == 1.5
    assert space.property_layers["empty"].dtype == bool and int((~space.property_layers["empty"]).sum()) == 3
    # agents: portrayal-style access
    seen = [(a.unique_id, a.cell.coordinate, tuple(a.cell.position), a.load) for a in space.agents]
    assert seen == [(1, (3, 5), (3.0, 5.0), 0.5), (2, (0, 0), (0.0, 0.0), 1.5), (3, (1, 1), (1.0, 1.0), 2.5),
                    (4, (1, 1), (1.0, 1.0), 3.5)]
    assert all(isinstance(a.cell.position, np.ndarray) and a.agent_type is Ant for a in space.agents)
    # grid.agents: cell by cell (row-major), registration order inside a cell -- discrete_space.py:84-87
    assert [a.unique_id for a in grid.agents] == [2, 3, 4, 1]
    # the snapshot is detached from the device state
    grid.sugar[0, 0] = -1.0
    assert space.property_layers["sugar"][0, 0] == 1.5
    # restricted snapshots (one layer, one attribute)
    small = grid.host_mirror(layers=["sugar"], attributes=[])
    assert list(small.property_layers) == ["sugar"] and not hasattr(small.agents[0], "load")
