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
