"""Wolf-Sheep on the device (mesa_b200/examples/wolf_sheep.py, csrc/wolf_sheep.cu) against the unmodified reference model
run under the replayed Philox order / draws (tests/golden/wolf_sheep_*.npz, oracle/make_golden.py): after every step the same
sheep and wolves (unique ids in registration order) in the same cells with bit-identical float64 energies, the same grass."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FIXTURES = ["wolf_sheep_20x20.npz", "wolf_sheep_30x30.npz", "wolf_sheep_12x15_crowded.npz", "wolf_sheep_9x7_nograss.npz"]


def _scenario(g, **kw):
    from mesa_b200.examples.wolf_sheep import WolfSheepScenario

    return WolfSheepScenario(width=int(g["width"]), height=int(g["height"]), initial_sheep=len(g["sheep_id_0"]),
                             initial_wolves=len(g["wolf_id_0"]), sheep_reproduce=float(g["sheep_reproduce"]),
                             wolf_reproduce=float(g["wolf_reproduce"]), wolf_gain_from_food=float(g["wolf_gain"]),
                             grass=bool(g["grass"]), grass_regrowth_time=int(g["regrowth"]),
                             sheep_gain_from_food=float(g["sheep_gain"]), **kw)


def _table(model, klass):
    a = model.agents_by_type[klass]
    return a.unique_ids().cpu().numpy(), a.get("cell").cpu().numpy().astype(np.int64), a.get("energy").cpu().numpy()


@pytest.mark.parametrize("name", FIXTURES)
def test_wolf_sheep_matches_the_reference_trajectory(name):
    from mesa_b200.examples.wolf_sheep import Sheep, Wolf, WolfSheep

    g = np.load(os.path.join(GOLDEN, name))
    init = {"sheep_id": g["sheep_id_0"], "sheep_cell": g["sheep_cell_0"], "sheep_energy": g["sheep_energy_0"],
            "wolf_id": g["wolf_id_0"], "wolf_cell": g["wolf_cell_0"], "wolf_energy": g["wolf_energy_0"],
            "grown": g["grown_0"], "regrow_at": g["regrow_at_0"], "next_id": int(g["next_id_0"])}
    m = WolfSheep(_scenario(g, rng=1), initial_state=init)
    m.philox_seed = int(g["philox_seed"])
    for s in range(1, int(g["steps"]) + 1):
        m.step()
        for kind, klass in (("sheep", Sheep), ("wolf", Wolf)):
            ids, cells, energy = _table(m, klass)
            assert np.array_equal(ids, g[f"{kind}_id_{s}"]), (name, s, kind)
            assert np.array_equal(cells, g[f"{kind}_cell_{s}"]), (name, s, kind)
            assert np.array_equal(energy, g[f"{kind}_energy_{s}"]), (name, s, kind)     # float64, bit for bit
        if bool(g["grass"]):
            assert np.array_equal(m.grid.grass_grown.cpu().numpy(), g[f"grown_{s}"]), (name, s)
    counts = m.datacollector.get_model_vars_dataframe().to_numpy()
    assert np.array_equal(counts, g["counts"])          # Wolves, Sheep, Grass as the reference's DataCollector saw them
    assert m.time == int(g["steps"]) and m.last_sheep_passes >= 1 and m.last_wolf_passes >= 1


@pytest.mark.parametrize("name", FIXTURES)
def test_same_seed_gives_the_reference_initial_state(name):
    """model.py:95-131 replayed call for call on the model's own generators."""
    from mesa_b200.examples.wolf_sheep import Sheep, Wolf, WolfSheep

    g = np.load(os.path.join(GOLDEN, name))
    m = WolfSheep(_scenario(g, rng=int(g["mt_seed"])))
    for kind, klass in (("sheep", Sheep), ("wolf", Wolf)):
        ids, cells, energy = _table(m, klass)
        assert np.array_equal(ids, g[f"{kind}_id_0"]) and np.array_equal(cells, g[f"{kind}_cell_0"])
        assert np.array_equal(energy, g[f"{kind}_energy_0"])
    if bool(g["grass"]):
        assert np.array_equal(m.grid.grass_grown.cpu().numpy(), g["grown_0"])
        assert np.array_equal(m.grid.grass_regrow_at.cpu().numpy(), g["regrow_at_0"])
    assert m.agent_id_counter == int(g["next_id_0"])


def test_large_population_conserves_and_is_reproducible():
    """10^5 sheep / 2 x 10^4 wolves on 512 x 512: two runs from the same seed agree exactly; ids stay unique; every sheep
    that disappears was eaten or starved (the population bookkeeping adds up)."""
    from mesa_b200.examples.wolf_sheep import Sheep, Wolf, WolfSheep, WolfSheepScenario

    def run():
        m = WolfSheep(WolfSheepScenario(width=512, height=512, initial_sheep=100_000, initial_wolves=20_000, rng=11))
        for _ in range(5):
            m.step()
        return m

    a, b = run(), run()
    for klass in (Sheep, Wolf):
        ia, ca, ea = _table(a, klass)
        ib, cb, eb = _table(b, klass)
        assert np.array_equal(ia, ib) and np.array_equal(ca, cb) and np.array_equal(ea, eb)
        assert len(np.unique(ia)) == len(ia) and (ea >= 0).all()
    assert torch.equal(a.grid.grass_grown, b.grid.grass_grown)
