"""Dynamic populations on the device: the stable-compaction kernels (csrc/agents.cu) against numpy, and
DeviceAgentSet.add_agents / remove_agents / remove / discard against the live reference's Agent.create_agents /
Agent.remove (tests/golden/dynamic_agents.npz; reference tests: tests/test_agentset.py::test_agentset remove / discard,
tests/test_agent.py::test_agent_removal)."""
import numpy as np
import pytest
import torch

from _dynamic_script import replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 255, 4096, 4097, 100_003])
def test_compaction_plan_is_the_exclusive_scan_of_the_keep_flags(n):
    from mesa_b200 import ops

    rng = np.random.default_rng(n)
    for density in (0.0, 0.3, 1.0):
        keep = rng.random(n) < density
        plan = ops.CompactionPlan(torch.from_numpy(keep).cuda())
        want = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
        assert np.array_equal(plan.offsets.cpu().numpy().astype(np.int64), want)
        assert plan.kept == int(keep.sum())


@pytest.mark.parametrize("dtype,trailing", [(torch.uint8, ()), (torch.uint8, (3,)), (torch.int16, ()), (torch.int16, (3,)),
                                            (torch.int32, ()), (torch.float32, (3,)), (torch.int64, ()), (torch.float32, (2,)),
                                            (torch.float64, (2,)), (torch.float32, (6,)), (torch.float64, (3, 2))])
def test_compact_rows_matches_boolean_indexing(dtype, trailing):
    from mesa_b200 import ops

    rng = np.random.default_rng(7)
    for n in (1, 33, 1000, 65_537):
        keep = rng.random(n) < 0.6
        plan = ops.CompactionPlan(torch.from_numpy(keep).cuda())
        raw = rng.integers(0, 120, size=(n, *trailing))
        col = torch.from_numpy(raw).to(dtype).cuda()
        got = plan.apply(col)
        assert got.dtype == dtype and tuple(got.shape) == (int(keep.sum()), *trailing)
        assert torch.equal(got.cpu(), col.cpu()[torch.from_numpy(keep)])
        # a column that starts at an odd byte offset takes the narrower copy unit
        if dtype == torch.uint8 and trailing == ():
            shifted = torch.empty(n + 1, dtype=dtype, device="cuda")[1:]
            shifted.copy_(col)
            assert torch.equal(plan.apply(shifted).cpu(), got.cpu())


def test_compaction_keep_all_keep_none_and_errors():
    from mesa_b200 import ops

    col = torch.arange(1000, dtype=torch.int32, device="cuda")
    assert torch.equal(ops.CompactionPlan(torch.ones(1000, dtype=torch.bool, device="cuda")).apply(col), col)
    none = ops.CompactionPlan(torch.zeros(1000, dtype=torch.bool, device="cuda"))
    assert none.kept == 0 and none.apply(col).shape == (0,)
    with pytest.raises(ValueError):
        none.apply(col[:10])
    with pytest.raises(RuntimeError):
        ops.CompactionPlan(torch.ones(4, dtype=torch.bool))             # CPU tensor: there is no CPU fallback
    with pytest.raises(TypeError):
        ops.CompactionPlan(torch.ones(4, dtype=torch.int32, device="cuda"))


def test_full_size_compaction_properties():
    """10^7 agents x (int32, float32[2], uint8): order preserved (ids stay ascending), every survivor carried its own
    row (checksum of id-derived payloads), and compacting twice == compacting by the combined mask."""
    from mesa_b200 import ops

    n = 10_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    ids = torch.arange(n, dtype=torch.int32, device="cuda")
    pos = torch.stack([ids.to(torch.float32) * 0.5, ids.to(torch.float32) * 0.25 + 1], dim=1).contiguous()
    tag = (ids % 251).to(torch.uint8)
    keep1 = torch.rand(n, device="cuda", generator=g) < 0.7
    p1 = ops.CompactionPlan(keep1)
    ids1, pos1, tag1 = p1.apply(ids), p1.apply(pos), p1.apply(tag)
    k1 = int(keep1.sum().item())
    assert p1.kept == k1 and ids1.shape[0] == k1
    assert bool((ids1[1:] > ids1[:-1]).all().item())
    assert int(ids1.to(torch.int64).sum().item()) == int(ids.to(torch.int64)[keep1].sum().item())
    assert torch.equal(pos1[:, 0], ids1.to(torch.float32) * 0.5) and torch.equal(pos1[:, 1], ids1.to(torch.float32) * 0.25 + 1)
    assert torch.equal(tag1, (ids1 % 251).to(torch.uint8))
    keep2 = torch.rand(k1, device="cuda", generator=g) < 0.5
    ids2 = ops.CompactionPlan(keep2).apply(ids1)
    both = keep1.clone()
    both[keep1] = keep2
    assert torch.equal(ids2, ops.CompactionPlan(both).apply(ids))


@pytest.mark.parametrize("by_mask", [True, False])
def test_dynamic_population_matches_the_reference_script(by_mask):
    model, agents = replay("cuda", by_mask)
    assert model.agents_by_type[agents.agent_type] is agents


def test_remove_discard_proxies_and_views():
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class Sheep(DeviceAgent):
        pass

    model = DeviceModel(rng=2)
    agents = DeviceAgentSet(model, Sheep, 6, {"energy": torch.tensor([5, 0, 3, 0, 9, 1], device="cuda", dtype=torch.int32)})
    model.agents = agents
    a2, a4 = agents[2], agents[4]
    view = agents.select(lambda c: c["energy"] > 0)
    agents.remove(agents[1])                                           # test_agentset.py: remove -> not in set
    assert len(agents) == 5 and agents.unique_ids().tolist() == [1, 3, 4, 5, 6]
    with pytest.raises(ReferenceError):
        _ = a2.energy                                                  # rows were renumbered
    assert a4 not in agents
    with pytest.raises(KeyError):
        agents.remove(a4)                                              # like `del dict[key]` (agentset.py:709-711)
    agents.discard(a4)                                                 # agentset.py:704-707: silent
    with pytest.raises(ReferenceError):
        len(list(view))
    starving = agents.select(lambda c: c["energy"] == 0)               # WolfSheep-style death: all at once
    agents.remove_agents(starving)
    assert agents.unique_ids().tolist() == [1, 3, 5, 6] and agents.get("energy").tolist() == [5, 3, 9, 1]
    lambs = agents.add_agents(2, energy=torch.tensor([4, 4]))          # reproduction: new ids continue the counter
    assert lambs.unique_ids().tolist() == [7, 8] and [a.unique_id for a in agents] == [1, 3, 5, 6, 7, 8]
    agents[0].remove()                                                 # agent.py:73-108 through the proxy
    assert agents.unique_ids().tolist() == [3, 5, 6, 7, 8]
    agents.add(agents[0])                                              # adding a member is a no-op
    with pytest.raises(TypeError):
        agents.add(a2)
    with pytest.raises(IndexError):
        agents.remove_agents([17])
    model.remove_all_agents()                                          # model.py:307-320
    assert len(agents) == 0 and agents.unique_ids().tolist() == []
    assert agents.add_agents(1).unique_ids().tolist() == [9]


def test_fixed_populations_refuse_resizing():
    from mesa_b200.examples import BoltzmannScenario, BoltzmannWealth

    m = BoltzmannWealth(BoltzmannScenario(n=50, width=10, height=10, rng=1))
    with pytest.raises(NotImplementedError):
        m.agents.remove(m.agents[0])
    with pytest.raises(NotImplementedError):
        m.agents.add_agents(3)


def test_datacollector_reports_surviving_unique_ids():
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.datacollection import DeviceDataCollector
    from mesa_b200.model import DeviceModel

    class Sheep(DeviceAgent):
        pass

    model = DeviceModel(rng=2)
    model.agents = DeviceAgentSet(model, Sheep, 4, {"energy": torch.tensor([1, 2, 3, 4], device="cuda", dtype=torch.int32)})
    dc = DeviceDataCollector(agent_reporters={"energy": "energy"})
    dc.collect(model)
    model.agents.remove_agents([1])
    model.time = 1.0
    dc.collect(model)
    df = dc.get_agent_vars_dataframe().reset_index()
    assert df[df.Step == 1.0].AgentID.tolist() == [1, 3, 4] and df[df.Step == 1.0].energy.tolist() == [1, 3, 4]


def _reference_groups(keys):
    """agentset.py:305-313: dict of lists in first-appearance order."""
    groups = {}
    for i, k in enumerate(keys):
        groups.setdefault(k, []).append(i)
    return groups


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64, torch.uint8, torch.bool, torch.float32, torch.float64])
def test_groupby_matches_the_reference_grouping(dtype):
    """agentset.py:282-320 + GroupBy (:837-970); epstein_civil_violence/model.py:122 uses groupby("state").count()."""
    from mesa_b200.agentset import DeviceAgent, DeviceAgentSet
    from mesa_b200.model import DeviceModel

    class Citizen(DeviceAgent):
        pass

    rng = np.random.default_rng(3)
    n = 5000
    raw = rng.integers(0, 2 if dtype == torch.bool else 7, size=n)
    if dtype in (torch.float32, torch.float64):
        host = (raw * 0.5 - 1.0)
        host[host == 0.0] = np.where(rng.random((host == 0.0).sum()) < 0.5, 0.0, -0.0)     # both zeros: one key
        state = torch.from_numpy(host).to(dtype)
    else:
        state = torch.from_numpy(raw).to(dtype)
    wealth = torch.from_numpy(rng.integers(0, 100, size=n)).to(torch.int32)
    model = DeviceModel(rng=1)
    agents = DeviceAgentSet(model, Citizen, n, {"state": state.cuda(), "wealth": wealth.cuda()})
    want = _reference_groups(state.tolist())
    got = agents.groupby("state")
    assert list(got.groups) == list(want) and got.count() == {k: len(v) for k, v in want.items()}
    assert len(got) == len(want) and repr(got).startswith("DeviceGroupBy(")
    for k, members in got:
        assert members.unique_ids().tolist() == [i + 1 for i in want[k]]                     # set order inside a group
    sums = got.agg("wealth", "sum")
    assert sums == {k: int(wealth[v].sum()) for k, v in want.items()}
    assert got.map("__len__") == got.count() and got.do(lambda g: None) is got
    first = next(iter(want))
    assert got.get_group(first) is got.groups[first] and got.get_group("missing", default=None) is None
    with pytest.raises(KeyError):
        got.get_group("missing")
    # a callable key over the columns, on a selected view, as lists
    rich = agents.select(lambda c: c["wealth"] >= 50)
    by_parity = rich.groupby(lambda c: c["wealth"] % 2, result_type="list")
    rows = [i for i in range(n) if wealth[i] >= 50]
    want2 = _reference_groups([int(wealth[i]) % 2 for i in rows])
    assert list(by_parity.groups) == list(want2)
    for k, members in by_parity:
        assert [a.unique_id - 1 for a in members] == [rows[j] for j in want2[k]]
    with pytest.raises(AttributeError):
        agents.groupby("mood")
    with pytest.raises(ValueError):
        agents.groupby("state", result_type="tuple")
    assert len(agents.select(lambda c: c["wealth"] > 1000).groupby("state")) == 0
