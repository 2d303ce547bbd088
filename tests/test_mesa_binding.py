"""The binding executed against a LIVE mesa (build container only; skipped where /root/reference is absent):

1. a `mesa.Model` subclass whose `model.agents` IS a DeviceAgentSet (mesa/model.py:143-148, 269-274), stepped by mesa's own
   event loop (`run_for`), with mesa's own DataCollector reading it;
2. the reference's own hot-path known-answer tests (tests/discrete_space/test_discrete_space.py, the functions SURVEY.md
   8c names) run UNCHANGED against the drop-in grids: the test module is imported from the reference tree and its
   `OrthogonalMooreGrid / OrthogonalVonNeumannGrid` names are pointed at the device classes.

No GPU here: tensors live on the CPU and the three device calls these paths make are replaced by CPU stand-ins
(`layer_fill`, `layer_reduce`, `CompactionPlan`); nothing else of the product is stubbed."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mesa")), reason="needs the reference tree")


class _HostPlan:
    """Test double of ops.CompactionPlan (the stable compaction of csrc/agents.cu) on CPU tensors."""

    def __init__(self, keep):
        self.keep = keep.to(bool)
        self.n, self.kept = keep.numel(), int(self.keep.sum())

    def apply(self, col):
        return col[self.keep].clone()


@pytest.fixture()
def live_mesa(monkeypatch):
    if REF not in sys.path:
        monkeypatch.syspath_prepend(REF)
    import mesa  # noqa: F401

    import mesa_b200.mesa_binding  # noqa: F401  (registers the virtual subclasses, mesa's exception types)
    from mesa_b200 import ops

    monkeypatch.setattr(ops, "layer_fill", lambda layer, value: layer.fill_(value))
    monkeypatch.setattr(ops, "layer_reduce", lambda t, op: {"sum": t.sum, "min": t.min, "max": t.max,
                                                            "count_nonzero": t.count_nonzero}[op]())
    monkeypatch.setattr(ops, "CompactionPlan", _HostPlan)
    yield


def _host_order(seed, epoch, n):
    """Agents in ascending Philox priority64 (csrc/philox.cuh) -- the host view of mesa_b200/rng.py."""
    from mesa_b200 import rng

    pri = rng.priority64(seed, epoch, torch.arange(n)) ^ (-2**63)     # unsigned order on the int64 bit pattern
    return torch.argsort(pri, stable=True)


def test_mesa_model_steps_through_a_device_agentset(live_mesa):
    import mesa
    from mesa.agentset import AbstractAgentSet
    from mesa.discrete_space import DiscreteSpace

    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet, batch_operator
    from mesa_b200.mesa_binding import DeviceBackedModel
    from mesa_b200.space import DeviceMooreGrid

    class Walker(DeviceAgent):
        __slots__ = ()

    # CPU stand-ins of two batch kernels (torch on CPU tensors): "step" records the activation rank of every agent of
    # this shuffle_do (the Philox priority order a kernel would use), "earn" adds to a column and honours views
    @batch_operator(Walker, "step")
    def _step(agents, *, epoch=None):
        order = _host_order(agents.model.philox_seed, epoch, agents.n)
        rank = torch.empty(agents.n, dtype=torch.int64)
        rank[order] = torch.arange(agents.n)
        agents.columns["rank"].copy_(rank)
        agents.columns["steps"] += 1

    @batch_operator(Walker, "earn", views=True)
    def _earn(agents, amount=1, *, epoch=None):
        col = agents.columns["wealth"]
        if agents.index is None:
            col += amount
        else:
            col[agents.index] += amount

    class WalkerModel(DeviceBackedModel):
        def __init__(self, n=50, rng=None):
            super().__init__(rng=rng)
            self.grid = DeviceMooreGrid((8, 8), torus=True, random=self.random, device="cpu")
            cols = {"wealth": torch.ones(n, dtype=torch.int64), "rank": torch.zeros(n, dtype=torch.int64),
                    "steps": torch.zeros(n, dtype=torch.int64)}
            self.adopt(DeviceAgentSet(self, Walker, n, cols, self.random))
            self.datacollector = mesa.DataCollector(model_reporters={"total": lambda m: m.agents.agg("wealth", sum)})

        def step(self):
            self.agents.shuffle_do("step")      # model.py / examples: the reference's own idiom, unchanged
            self.agents.do("earn", 2)
            self.datacollector.collect(self)

    m = WalkerModel(rng=7)
    assert isinstance(m, mesa.Model) and isinstance(m.agents, AbstractAgentSet) and isinstance(m.agents, DeviceAgentSet)
    assert isinstance(m.grid, DiscreteSpace) and isinstance(m.grid.host_mirror(), DiscreteSpace)
    assert m.agents_by_type[Walker] is m.agents and m.agent_types == [Walker] and len(m.agents) == 50
    m.run_for(3)                                  # mesa's event loop calls our step three times
    assert m.time == 3.0 and m._epoch == 3
    assert m.agents.get("steps").tolist() == [3] * 50 and m.agents.get("wealth").tolist() == [7] * 50
    # the recorded ranks are a permutation: the order of the LAST activation (epoch 2), replayable on the host
    assert torch.equal(torch.argsort(m.agents.get("rank")), _host_order(m.philox_seed, 2, 50))
    import oracle

    assert np.array_equal(_host_order(m.philox_seed, 2, 50).numpy(), oracle.activation_order(m.philox_seed, 2, 50))
    assert m.datacollector.get_model_vars_dataframe()["total"].tolist() == [150, 250, 350]
    # do() on a select() view reaches the operator with the view's rows; an operator without views=True refuses
    rich = m.agents.select(lambda cols: torch.arange(50) < 10)   # column predicate: the first ten registered agents
    assert len(rich) == 10
    rich.do("earn", 5)
    assert m.agents.get("wealth").tolist() == [12] * 10 + [7] * 40
    with pytest.raises(NotImplementedError):
        rich.shuffle_do("step")
    with pytest.raises(NotImplementedError):
        m.agents.do("no_such_method")
    with pytest.raises(AttributeError):           # model.py:214-219: model.agents cannot be assigned
        m.agents = None
    m.remove_all_agents()
    assert len(m.agents) == 0


def _reference_tests(monkeypatch, relpath: str, **names):
    path = os.path.join(REF, relpath)
    spec = importlib.util.spec_from_file_location("ref_" + os.path.basename(relpath)[:-3], path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for k, v in names.items():
        monkeypatch.setattr(mod, k, v)
    return mod


GRID_KATS = [
    "test_orthogonal_grid_neumann", "test_orthogonal_grid_neumann_3d", "test_orthogonal_grid_moore",
    "test_orthogonal_grid_moore_3d", "test_orthogonal_grid_moore_4d", "test_orthogonal_grid_moore_1d",
    "test_cell_neighborhood", "test_cell_missing_exception", "test_grid_validate_parameters",
    "test_property_layer_integration", "test_multiple_property_layers", "test_get_neighborhood_mask",
    "test_large_radius_neighborhood",
]


@pytest.mark.parametrize("name", GRID_KATS)
def test_reference_grid_kats_pass_on_the_device_grids(live_mesa, monkeypatch, name):
    from mesa_b200.space import DeviceMooreGrid, DeviceVonNeumannGrid

    class CpuMoore(DeviceMooreGrid):
        def __init__(self, dimensions, torus=False, capacity=None, random=None, **kw):
            super().__init__(dimensions, torus=torus, capacity=capacity, random=random, device="cpu")

    class CpuVonNeumann(DeviceVonNeumannGrid):
        def __init__(self, dimensions, torus=False, capacity=None, random=None, **kw):
            super().__init__(dimensions, torus=torus, capacity=capacity, random=random, device="cpu")

    mod = _reference_tests(monkeypatch, "tests/discrete_space/test_discrete_space.py",
                           OrthogonalMooreGrid=CpuMoore, OrthogonalVonNeumannGrid=CpuVonNeumann)
    getattr(mod, name)()
