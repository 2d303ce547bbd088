"""Epstein's civil-violence model on the device (mesa_b200/examples/epstein_civil_violence.py, csrc/epstein.cu) against the
unmodified reference model run under the replayed Philox order / draws (tests/golden/epstein_*.npz, oracle/make_golden.py):
after every step every agent's cell, state, jail sentence and float64 arrest probability, bit for bit; at sizes the
reference cannot reach in seconds, against the sequential restatement (oracle/epstein.py, itself pinned by the goldens)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = ["epstein_40x40.npz", "epstein_30x22_lively.npz", "epstein_17x19_moore.npz", "epstein_16x16_still.npz"]


def _scenario(g, **kw):
    from mesa_b200.examples.epstein_civil_violence import EpsteinScenario

    return EpsteinScenario(citizen_density=float(g["citizen_density"]), cop_density=float(g["cop_density"]),
                           citizen_vision=int(g["vision"]), cop_vision=int(g["vision"]), legitimacy=float(g["legitimacy"]),
                           max_jail_term=int(g["max_jail_term"]), active_threshold=float(g["active_threshold"]),
                           arrest_prob_constant=float(g["arrest_prob_constant"]), movement=bool(g["movement"]),
                           grid_type="Moore" if bool(g["moore"]) else "Von Neumann", **kw)


def _check_step(m, g, s, name):
    a = m.agents
    cit = g["kind"] == 0
    assert np.array_equal(a.get("cell").cpu().numpy().astype(np.int64), g[f"cell_{s}"]), (name, s, "cell")
    assert np.array_equal(a.get("state").cpu().numpy()[cit], g[f"state_{s}"][cit].astype(np.uint8)), (name, s, "state")
    assert np.array_equal(a.get("jail_sentence").cpu().numpy().astype(np.int64), g[f"jail_{s}"]), (name, s, "jail")
    ap, want = a.get("arrest_probability").cpu().numpy(), g[f"ap_{s}"]
    assert np.array_equal(np.isnan(ap), np.isnan(want)), (name, s, "arrest probability set")
    assert np.array_equal(ap[~np.isnan(want)], want[~np.isnan(want)]), (name, s, "arrest probability")   # float64, bit for bit
    assert (m.ACTIVE, m.QUIET, m.ARRESTED) == tuple(int(v) for v in g["counts"][s]), (name, s, "counts")


@pytest.mark.parametrize("name", FIXTURES)
def test_epstein_matches_the_reference_trajectory(name):
    from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence

    g = np.load(os.path.join(GOLDEN, name))
    init = {"kind": g["kind"], "cell": g["cell_0"], "hardship": g["hardship"], "risk_aversion": g["risk_aversion"]}
    m = EpsteinCivilViolence(int(g["width"]), int(g["height"]), _scenario(g, rng=1), initial_state=init)
    m.philox_seed = int(g["philox_seed"])
    for s in range(1, int(g["steps"]) + 1):
        m.step()
        _check_step(m, g, s, name)
    counts = m.datacollector.get_model_vars_dataframe().to_numpy()
    assert np.array_equal(counts, g["counts"])              # active / quiet / arrested as the reference's DataCollector saw them
    # the occupancy words agree with the agent table, `grid.empty` follows them
    occ = m._occ.cpu().numpy()
    cell = m.agents.get("cell").cpu().numpy()
    assert np.array_equal(occ[cell] >> 2, np.arange(len(cell))) and (occ >= 0).sum() == len(cell)
    assert np.array_equal(m.grid.empty.cpu().numpy().reshape(-1), occ < 0)


@pytest.mark.parametrize("name", FIXTURES)
def test_same_seed_gives_the_reference_initial_state(name):
    """model.py:84-98 replayed call for call on the model's own generator."""
    from mesa_b200.examples.epstein_civil_violence import Citizen, Cop, EpsteinCivilViolence

    g = np.load(os.path.join(GOLDEN, name))
    m = EpsteinCivilViolence(int(g["width"]), int(g["height"]), _scenario(g, rng=int(g["mt_seed"])))
    a = m.agents
    assert np.array_equal(a.get("kind").cpu().numpy(), g["kind"].astype(np.uint8))
    assert np.array_equal(a.get("cell").cpu().numpy().astype(np.int64), g["cell_0"])
    assert np.array_equal(a.get("hardship").cpu().numpy(), g["hardship"])
    assert np.array_equal(a.get("risk_aversion").cpu().numpy(), g["risk_aversion"])
    assert len(m.agents_by_type[Citizen]) == int((g["kind"] == 0).sum()) and len(m.agents_by_type[Cop]) == int((g["kind"] == 1).sum())
    assert (m.ACTIVE, m.QUIET, m.ARRESTED) == tuple(int(v) for v in g["counts"][0])


@pytest.mark.parametrize("moore,vision,side", [(False, 7, 160), (True, 3, 128), (False, 2, 200)])
def test_epstein_matches_the_sequential_restatement_on_larger_grids(moore, vision, side):
    """Tens of thousands of agents, lively parameters: device == oracle/epstein.py after every step."""
    from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence, EpsteinScenario
    from oracle.epstein import EpsteinOracle

    rs = np.random.default_rng(5 + vision)
    ncells = side * (side + 3)
    u = rs.random(ncells)
    occupied = np.flatnonzero(u < 0.78)
    kind = (u[occupied] >= 0.72).astype(np.uint8)
    hardship = np.where(kind == 0, rs.random(len(kind)), 0.0)
    risk = np.where(kind == 0, rs.random(len(kind)), 0.0)
    sc = EpsteinScenario(citizen_vision=vision, cop_vision=vision, legitimacy=0.55, max_jail_term=5,
                         grid_type="Moore" if moore else "Von Neumann", rng=9)
    m = EpsteinCivilViolence(side, side + 3, sc, initial_state={"kind": kind, "cell": occupied, "hardship": hardship, "risk_aversion": risk})
    o = EpsteinOracle(side, side + 3, kind, occupied, hardship, risk, moore=moore, vision=vision, legitimacy=0.55,
                      max_jail_term=5, philox_seed=m.philox_seed)
    for s in range(3):
        m.step()
        o.step()
        a = m.agents
        assert np.array_equal(a.get("cell").cpu().numpy().astype(np.int64), o.cell), s
        cit = kind == 0
        assert np.array_equal(a.get("state").cpu().numpy()[cit], o.state[cit].astype(np.uint8)), s
        assert np.array_equal(a.get("jail_sentence").cpu().numpy().astype(np.int64), o.jail), s
        assert np.array_equal(a.get("arrest_probability").cpu().numpy()[cit], o.arrest_probability[cit]), s
        assert (m.ACTIVE, m.QUIET, m.ARRESTED) == o.counts(), s
    assert m.ACTIVE > 0 and m.ARRESTED > 0


def test_registration_order_activation():
    """activation_order "Sequential" = agents.do("step"): registration order; draws keyed on the step's own epoch."""
    from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence, EpsteinScenario
    from oracle.epstein import EpsteinOracle
    from oracle import philox

    g = np.load(os.path.join(GOLDEN, "epstein_30x22_lively.npz"))
    init = {"kind": g["kind"], "cell": g["cell_0"], "hardship": g["hardship"], "risk_aversion": g["risk_aversion"]}
    m = EpsteinCivilViolence(int(g["width"]), int(g["height"]), _scenario(g, rng=4, activation_order="Sequential"), initial_state=init)
    o = EpsteinOracle(int(g["width"]), int(g["height"]), g["kind"], g["cell_0"], g["hardship"], g["risk_aversion"],
                      vision=int(g["vision"]), legitimacy=float(g["legitimacy"]), max_jail_term=int(g["max_jail_term"]),
                      philox_seed=m.philox_seed)
    real = philox.permutation
    philox.permutation = lambda seed, epoch, n: np.arange(n)
    try:
        for s in range(6):
            m.step()
            o.step()
            assert np.array_equal(m.agents.get("cell").cpu().numpy().astype(np.int64), o.cell), s
            assert np.array_equal(m.agents.get("jail_sentence").cpu().numpy().astype(np.int64), o.jail), s
    finally:
        philox.permutation = real
