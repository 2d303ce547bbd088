"""mesa_b200.reference_init: the same seed gives the same streams and the same INITIAL model state as the reference
(scenario.py:156-158, 214-221; schelling/model.py:71-81; conways_game_of_life/model.py:19-27;
boltzmann_wealth_model/model.py:70-74; boid_flockers/model.py:79-82).  Pinned by the initial states stored in the golden
fixtures that oracle/make_golden.py produced by constructing the unmodified reference models from these seeds."""
import os
import random

import numpy as np
import pytest

from mesa_b200 import reference_init as ri

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_mt_uniforms_is_random_random_and_advances_the_generator():
    a, b = random.Random(123), random.Random(123)
    for n in (1, 7, 624, 1000):
        want = [a.random() for _ in range(n)]
        assert ri.mt_uniforms(b, n).tolist() == want
        assert a.getstate() == b.getstate()
    peek = ri.mt_uniforms(b, 5, advance=False)
    assert a.getstate() == b.getstate() and peek.tolist() == [a.random() for _ in range(5)]


def test_seed_streams_match_scenario_seeding():
    gen, stdlib = ri.seed_streams(42)
    ss = np.random.SeedSequence(42)
    assert stdlib == int(ss.generate_state(1, dtype=np.uint32)[0])
    assert gen.random() == np.random.default_rng(42).random()
    g = np.random.default_rng(7)
    assert ri.seed_streams(g)[0] is g                                     # a Generator is used as it is


@pytest.mark.parametrize("name,seed", [("gol_37x23.npz", 42), ("gol_64x48.npz", 7)])
def test_game_of_life_initial_state_is_the_reference(name, seed):
    g = np.load(os.path.join(GOLDEN, name))
    s0 = g["states"][0]
    rnd = random.Random(ri.seed_streams(seed)[1])
    assert np.array_equal(ri.gol_initial_state(rnd, s0.shape[0], s0.shape[1], 0.2), s0.astype(bool))


@pytest.mark.parametrize("name,seed,density", [("schelling_20x20_r1.npz", 42, 0.8), ("schelling_16x48_r2.npz", 5, 0.9),
                                               ("schelling_100x100_c1.npz", 42, 0.8)])
def test_schelling_initial_grid_is_the_reference(name, seed, density):
    g = np.load(os.path.join(GOLDEN, name))
    w, h = int(g["width"]), int(g["height"])
    want = np.full(w * h, -1, dtype=np.int8)
    want[g["cell0"]] = g["atype"]
    rnd = random.Random(ri.seed_streams(seed)[1])
    check = random.Random(ri.seed_streams(seed)[1])
    got = ri.schelling_initial_types(rnd, w, h, density, 0.5)
    assert np.array_equal(got.reshape(-1), want)
    # the generator stands exactly where the reference's stands after its per-cell loop
    for _ in range(w * h):
        if check.random() < density:
            check.random()
    assert rnd.getstate() == check.getstate()


@pytest.mark.parametrize("name,seed", [("boltzmann_100_10x10.npz", 42), ("boltzmann_2000_20x30.npz", 7),
                                       ("boltzmann_10000_100x100.npz", 1)])
def test_boltzmann_initial_cells_are_the_reference(name, seed):
    g = np.load(os.path.join(GOLDEN, name))
    n, ncells = int(g["n"]), int(g["width"]) * int(g["height"])
    rnd = random.Random(ri.seed_streams(seed)[1])
    assert np.array_equal(ri.boltzmann_initial_cells(rnd, n, ncells), g["cell_0"].astype(np.int64))


def test_example_models_start_from_the_reference_initial_state(monkeypatch):
    """Model level, on CPU tensors with the device work stubbed out (the constructors only): `Model(rng=seed)` holds the
    reference's initial state for that seed."""
    import torch

    from mesa_b200 import ops
    from mesa_b200.examples import (BoltzmannScenario, BoltzmannWealth, ConwaysGameOfLife, Schelling, SchellingScenario)

    class _Stub:
        def __init__(self, *a, **k):
            pass

    monkeypatch.setattr(ops, "RelocationWorkspace", _Stub)
    monkeypatch.setattr(ops, "BoltzmannWorkspace", _Stub)
    monkeypatch.setattr(ops, "layer_fill", lambda layer, value: layer.fill_(value))
    monkeypatch.setattr(ops, "schelling_happy", lambda *a, **k: (None, torch.zeros(1, dtype=torch.int64)))

    g = np.load(os.path.join(GOLDEN, "gol_37x23.npz"))
    m = ConwaysGameOfLife(width=37, height=23, rng=42, device="cpu", packed=False)
    assert np.array_equal(m.grid.state.numpy(), g["states"][0])

    g = np.load(os.path.join(GOLDEN, "schelling_20x20_r1.npz"))
    want = np.full(400, -1, dtype=np.int8)
    want[g["cell0"]] = g["atype"]
    s = Schelling(SchellingScenario(width=20, height=20, density=0.8, homophily=0.4, radius=1, rng=42), device="cpu",
                  collect=False, fused=False)
    assert np.array_equal(s.grid.type.numpy().reshape(-1), want)
    assert np.array_equal(s.agents.get("cell").numpy(), g["cell0"]) and np.array_equal(s.agents.get("type").numpy(), g["atype"])

    g = np.load(os.path.join(GOLDEN, "boltzmann_2000_20x30.npz"))
    b = BoltzmannWealth(BoltzmannScenario(n=2000, width=20, height=30, rng=7), device="cpu")
    assert np.array_equal(b.agents.get("cell").numpy(), g["cell_0"]) and b.agents.get("wealth").tolist() == [1] * 2000


@pytest.mark.parametrize("density", [0.0, 0.05, 0.5, 0.95, 1.0])
@pytest.mark.parametrize("minority", [0.0, 0.3, 1.0])
def test_schelling_vectorised_draws_equal_the_per_cell_loop(density, minority):
    """The running-maximum formulation against the reference's literal loop (schelling/model.py:71-81), including the
    extremes where every cell takes one draw (density 0) or two (density 1: all 2 * ncells uniforms are consumed)."""
    for seed in (0, 1, 2, 12345):
        for (w, h) in [(1, 1), (3, 7), (16, 16), (25, 9)]:
            a, b = random.Random(seed), random.Random(seed)
            want = np.full(w * h, -1, dtype=np.int8)
            for c in range(w * h):
                if a.random() < density:
                    want[c] = 1 if a.random() < minority else 0
            got = ri.schelling_initial_types(b, w, h, density, minority)
            assert got.shape == (w, h) and got.dtype == np.int8 and np.array_equal(got.reshape(-1), want)
            assert a.getstate() == b.getstate()


@pytest.mark.parametrize("name", ["epstein_40x40.npz", "epstein_30x22_lively.npz", "epstein_17x19_moore.npz", "epstein_16x16_still.npz"])
def test_epstein_initial_agents_are_the_reference(name):
    """epstein_civil_violence/model.py:84-98 + agents.py:55-56 from the fixture's seed: kinds, cells, hardship, risk aversion."""
    g = np.load(os.path.join(GOLDEN, name))
    st = ri.epstein_initial_state(random.Random(ri.seed_streams(int(g["mt_seed"]))[1]), int(g["width"]), int(g["height"]),
                                  float(g["citizen_density"]), float(g["cop_density"]))
    assert np.array_equal(st["kind"], g["kind"].astype(np.uint8)) and np.array_equal(st["cell"], g["cell_0"])
    assert np.array_equal(st["hardship"], g["hardship"]) and np.array_equal(st["risk_aversion"], g["risk_aversion"])


@pytest.mark.parametrize("name", ["sugarscape_200.npz", "sugarscape_600_crowded.npz", "sugarscape_80_farsighted.npz",
                                  "sugarscape_200_trade.npz", "sugarscape_500_trade_crowded.npz"])
def test_sugarscape_initial_traders_are_the_reference(name):
    """sugarscape_g1mt/model.py:80-116 from the fixture's seed (stdlib stream for the cells, numpy stream for the rest)."""
    g = np.load(os.path.join(GOLDEN, name))
    n, emin, emax, mmin, mmax, vmin, vmax = (int(v) for v in g["scenario"])
    gen, stdlib = ri.seed_streams(int(g["mt_seed"]))
    st = ri.sugarscape_initial_state(random.Random(stdlib), gen, int(g["width"]) * int(g["height"]), n, emin, emax, mmin, mmax, vmin, vmax)
    assert np.array_equal(st["cell"], g["cell_0"])
    assert np.array_equal(st["sugar"], g["sugar_0"]) and np.array_equal(st["spice"], g["spice_0"])
    assert np.array_equal(st["metabolism_sugar"], g["metabolism_sugar"]) and np.array_equal(st["metabolism_spice"], g["metabolism_spice"])
    assert np.array_equal(st["vision"], g["vision"])
