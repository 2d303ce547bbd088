"""DeviceDataCollector (mesa_b200/datacollection.py): the reference DataCollector's API over device columns.

CPU part: reporter forms, tables, data-frame layout, errors (torch CPU tensors stand in for the device columns --
the collector only copies columns to the host).  GPU part: the device Schelling model's collector reproduces the
UNMODIFIED reference model's own DataCollector output (golden fixture from oracle/make_golden.py)."""
import os
import random

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class _Agent:
    pass


class _FakeModel:
    def __init__(self, n):
        from mesa_b200.agentset import DeviceAgentSet

        self.time = 0.0
        self.random = random.Random(0)
        self.wealth = torch.arange(n, dtype=torch.int32)
        self.agents = DeviceAgentSet(self, _Agent, n, {"wealth": self.wealth, "pos": torch.zeros((n, 2))}, self.random)
        self.total = 0

    def gini(self):
        return float(self.wealth.float().mean())


def test_collector_reporter_forms_and_frames():
    from mesa_b200.datacollection import DeviceDataCollector

    m = _FakeModel(5)
    dc = DeviceDataCollector(
        model_reporters={"total": "total", "n": lambda mm: len(mm.agents), "gini": m.gini,
                         "scaled": [lambda a, b: a * b, [3, 4]], "sum": lambda mm: mm.wealth.sum()},
        agent_reporters={"wealth": "wealth", "double": lambda agents: agents.get("wealth") * 2,
                         "plus": [lambda agents, k: agents.get("wealth") + k, [10]], "pos": "pos"},
        tables={"Lifespan": ["unique_id", "age"]},
    )
    for step in range(3):
        m.time, m.total = float(step), step * 7
        m.wealth += 1
        dc.collect(m)
    mv = dc.get_model_vars_dataframe()
    assert list(mv.columns) == ["total", "n", "gini", "scaled", "sum"] and len(mv) == 3
    assert list(mv["total"]) == [0, 7, 14] and list(mv["n"]) == [5, 5, 5] and list(mv["scaled"]) == [12, 12, 12]
    assert list(mv["sum"]) == [15, 20, 25] and all(isinstance(v, int) for v in mv["sum"])
    av = dc.get_agent_vars_dataframe()
    assert list(av.index.names) == ["Step", "AgentID"] and list(av.columns) == ["wealth", "double", "plus", "pos"]
    assert len(av) == 15 and av.loc[(2.0, 5), "wealth"] == 7 and av.loc[(0.0, 1), "double"] == 2 and av.loc[(1.0, 3), "plus"] == 14
    assert np.array_equal(av.loc[(0.0, 1), "pos"], [0.0, 0.0])
    # recorded values are copies: later changes of the columns do not alter earlier records
    m.wealth += 100
    assert dc.get_agent_vars_dataframe().loc[(0.0, 1), "wealth"] == 1
    dc.add_table_row("Lifespan", {"unique_id": 3, "age": 9})
    dc.add_table_row("Lifespan", {"unique_id": 4}, ignore_missing=True)
    assert dc.tables["Lifespan"] == {"unique_id": [3, 4], "age": [9, None]}
    assert list(dc.get_table_dataframe("Lifespan")["unique_id"]) == [3, 4]


def test_collector_errors_match_reference():
    from mesa_b200.datacollection import DeviceDataCollector, TableMissingException

    with pytest.raises(ValueError):
        DeviceDataCollector(model_reporters={"bad": [lambda: 1]})                      # datacollection.py:143-155
    with pytest.raises(ValueError):
        DeviceDataCollector(model_reporters={"bad": [lambda a: a, 5]})
    with pytest.raises(TypeError):
        DeviceDataCollector(agent_reporters={"bad": 3})
    dc = DeviceDataCollector(agent_reporters={"w": "missing_column"}, tables={"T": ["a"]})
    with pytest.raises(AttributeError):
        dc.collect(_FakeModel(3))
    with pytest.raises(ValueError):
        DeviceDataCollector(agent_reporters={"w": lambda agents: torch.zeros(2)}).collect(_FakeModel(3))
    with pytest.raises(TableMissingException):
        dc.add_table_row("nope", {})
    with pytest.raises(ValueError):
        dc.add_table_row("T", {})
    with pytest.warns(UserWarning):
        assert DeviceDataCollector().get_model_vars_dataframe().empty
    with pytest.warns(UserWarning):
        assert DeviceDataCollector().get_agent_vars_dataframe().empty
    typed = DeviceDataCollector(agenttype_reporters={_Agent: {"w": "wealth"}})
    m = _FakeModel(4)
    typed.collect(m)
    assert list(typed.get_agenttype_vars_dataframe(_Agent)["w"]) == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        DeviceDataCollector(agenttype_reporters={int: {"w": "wealth"}}).collect(m)


@pytest.mark.gpu
def test_schelling_collector_reproduces_reference_datacollector():
    """model.py:53-70 + :85, :92: same reporters, same collection points, same frames as the unmodified reference run
    with the replayed activation order."""
    from mesa_b200.examples import Schelling, SchellingScenario

    g = np.load(os.path.join(GOLDEN, "datacollector_schelling_20x20.npz"))
    sc = SchellingScenario(width=int(g["width"]), height=int(g["height"]), density=float(g["density"]),
                           homophily=float(g["homophily"]), radius=int(g["radius"]), rng=int(g["philox_seed"]))
    m = Schelling(sc, initial_types=g["types0"])
    for _ in range(int(g["steps"])):
        m.step()
    mv = m.datacollector.get_model_vars_dataframe()
    assert list(mv.columns) == [str(c) for c in g["model_columns"]]
    for c in mv.columns:
        assert np.array_equal(mv[c].to_numpy(dtype=np.float64), g["model_" + c]), c
    av = m.datacollector.get_agent_vars_dataframe().reset_index()
    assert np.array_equal(av["Step"].to_numpy(dtype=np.float64), g["agent_step"])
    assert np.array_equal(av["AgentID"].to_numpy(dtype=np.int64), g["agent_id"])
    assert np.array_equal(av["agent_type"].to_numpy(dtype=np.int64), g["agent_type"])
    assert m.time == float(g["steps"]) and m.steps == int(g["steps"])


def test_ragged_column_is_lazy_and_materialises_to_lists():
    """datacollection.RaggedColumn: a per-agent column of lists kept as CSR until the DataFrame is built."""
    from mesa_b200.datacollection import DeviceDataCollector, RaggedColumn

    col = RaggedColumn([0, 2, 2, 5], [7, 8, 1, 2, 3])
    assert len(col) == 3 and col.shape == (3,) and [list(x) for x in col.materialize()] == [[7, 8], [], [1, 2, 3]]
    assert col.copy().offsets is not col.offsets
    with pytest.raises(ValueError):
        RaggedColumn([0, 2], [1])
    assert DeviceDataCollector._concat([col]) is col                       # one population: stays lazy
    both = DeviceDataCollector._concat([col, np.array([[9], [9]], dtype=object).reshape(2, 1)[:, 0]])
    assert len(both) == 5 and list(both[0]) == [7, 8]
