"""Checkpoint / resume (the reference pickles models: grid.py:401-431, agentset.py:617-632): a device model
survives a pickle round trip -- layers, agent columns, workspaces, RNG state, activation epoch -- and the restored copy
continues on exactly the same trajectory as the original."""
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roundtrip(m):
    return pickle.loads(pickle.dumps(m))


def test_pickle_game_of_life_packed_and_unpacked():
    from mesa_b200.examples import ConwaysGameOfLife

    for packed in (True, False):
        m = ConwaysGameOfLife(width=48, height=64, rng=3, packed=packed)
        m.run_for(5)
        c = _roundtrip(m)
        assert c.steps == m.steps == 5 and torch.equal(c.grid.state, m.grid.state)
        m.run_for(7)
        c.run_for(7)
        assert torch.equal(c.grid.state, m.grid.state)


def test_pickle_schelling_boltzmann_boids_continue_identically():
    from mesa_b200.examples import (BoidFlockers, BoidsScenario, BoltzmannScenario, BoltzmannWealth, Schelling,
                                    SchellingScenario)

    s = Schelling(SchellingScenario(width=40, height=30, density=0.8, homophily=0.5, radius=1, rng=5), collect=False)
    b = BoltzmannWealth(BoltzmannScenario(n=2000, width=30, height=20, rng=5))
    f = BoidFlockers(BoidsScenario(population_size=600, width=120, height=90, vision=10.0, rng=2), activation="synchronous")
    q = BoidFlockers(BoidsScenario(population_size=300, width=100, height=80, vision=10.0, rng=4), activation="sequential")
    for m in (s, b, f, q):
        for _ in range(3):
            m.step()
    copies = [_roundtrip(m) for m in (s, b, f, q)]
    for m, c in zip((s, b, f, q), copies):
        for _ in range(4):
            m.step()
            c.step()
        assert c.time == m.time and c._epoch == m._epoch
    assert torch.equal(copies[0].grid.type, s.grid.type) and copies[0].happy == s.happy
    assert torch.equal(copies[0].agents.get("cell"), s.agents.get("cell"))
    assert torch.equal(copies[1].agents.get("wealth"), b.agents.get("wealth")) and torch.equal(copies[1].agents.get("cell"), b.agents.get("cell"))
    assert torch.equal(copies[2].agents.get("position"), f.agents.get("position"))
    assert torch.equal(copies[3].agents.get("direction"), q.agents.get("direction"))


def test_pickle_epstein_and_sugarscape_continue_identically():
    """The dependency-graph models (one population with views; a dynamic population with trade records) survive the round
    trip and continue on the same trajectory."""
    from mesa_b200.examples.epstein_civil_violence import Citizen, EpsteinCivilViolence, EpsteinScenario
    from mesa_b200.examples.sugarscape_g1mt import SugarScapeScenario, SugarscapeG1mt
    from mesa_b200.examples.wolf_sheep import Sheep, WolfSheep, WolfSheepScenario

    e = EpsteinCivilViolence(30, 24, EpsteinScenario(legitimacy=0.6, max_jail_term=5, cop_vision=4, citizen_vision=4, rng=3))
    s = SugarscapeG1mt(SugarScapeScenario(initial_population=300, enable_trade=True, rng=4), width=40, height=36)
    w = WolfSheep(WolfSheepScenario(width=24, height=20, initial_sheep=150, initial_wolves=30, rng=6))
    for m in (e, s, w):
        for _ in range(3):
            m.step()
    ce, cs, cw = _roundtrip(e), _roundtrip(s), _roundtrip(w)
    for _ in range(4):
        for m in (e, s, w, ce, cs, cw):
            m.step()
    assert torch.equal(cw.agents_by_type[Sheep].unique_ids(), w.agents_by_type[Sheep].unique_ids())
    assert torch.equal(cw.agents_by_type[Sheep].get("energy"), w.agents_by_type[Sheep].get("energy"))
    assert torch.equal(cw.grid.grass_grown, w.grid.grass_grown)
    assert ce.time == e.time and ce._epoch == e._epoch and (ce.ACTIVE, ce.QUIET, ce.ARRESTED) == (e.ACTIVE, e.QUIET, e.ARRESTED)
    for col in ("cell", "state", "jail_sentence"):
        assert torch.equal(ce.agents.get(col), e.agents.get(col)), col
    assert len(ce.agents_by_type[Citizen]) == len(e.agents_by_type[Citizen])
    assert torch.equal(cs.agents.unique_ids(), s.agents.unique_ids())
    for col in ("cell", "sugar", "spice"):
        assert torch.equal(cs.agents.get(col), s.agents.get(col)), col
    assert torch.equal(cs.grid.sugar, s.grid.sugar) and np.array_equal(cs.trade_records()[2], s.trade_records()[2])


def test_dependency_graph_models_without_agents():
    """Empty populations step without a launch failure (every citizen density 0; every trader starved)."""
    from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence, EpsteinScenario
    from mesa_b200.examples.sugarscape_g1mt import SugarScapeScenario, SugarscapeG1mt

    e = EpsteinCivilViolence(20, 20, EpsteinScenario(citizen_density=0.0, cop_density=0.0, rng=1))
    e.step()
    assert len(e.agents) == 0 and (e.ACTIVE, e.QUIET, e.ARRESTED) == (0, 0, 0)
    init = {"cell": np.arange(5), "sugar": np.ones(5), "spice": np.ones(5), "metabolism_sugar": np.full(5, 5),
            "metabolism_spice": np.full(5, 5), "vision": np.ones(5, dtype=np.int64)}
    s = SugarscapeG1mt(SugarScapeScenario(enable_trade=True, rng=1), sugar_map=np.zeros((12, 12)), initial_state=init)
    s.step()
    assert len(s.agents) == 0                 # nothing to eat: everybody starves in the first step
    s.step()
    assert len(s.agents) == 0 and len(s.trade_records()[2]) == 0
