"""Generate tests/golden/*.npz from the LIVE reference -- run in the build container only.

    python -m oracle.make_golden

Imports the unmodified reference from /root/reference, runs the four example models (the
Schelling / Boltzmann activations replayed with the keyed Philox draws, oracle/replay.py),
and stores small input/output vectors.  The fixtures pin the CPU oracle
(tests/test_oracle_golden.py) and the CUDA path (tests/test_*_gpu.py); they travel to the GPU
box, the reference does not.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle.replay import ReplayRandom, activation_context, import_reference, patched_model_random  # noqa: E402


def digest(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


# ------------------------------------------------------------------------------ Game of Life
def gol_fixture(width, height, seed, steps, name):
    from mesa.examples.basic.conways_game_of_life.model import ConwaysGameOfLife

    m = ConwaysGameOfLife(width=width, height=height, rng=seed)

    def state():
        s = np.zeros((width, height), np.uint8)
        for a in m.agents:
            s[a.cell.coordinate] = a.state
        return s

    states = [state()]
    for _ in range(steps):
        m.step()
        states.append(state())
    np.savez_compressed(os.path.join(GOLDEN, name), states=np.stack(states), torus=np.array(True))
    print(name, np.stack(states).shape)


# ------------------------------------------------------------------------------ Schelling
def schelling_fixture(name, width, height, density, homophily, radius, steps, mt_seed, philox_seed,
                      full_steps):
    from mesa.examples.basic.schelling.agents import SchellingAgent
    from mesa.examples.basic.schelling.model import Schelling, SchellingScenario

    with patched_model_random(philox_seed):
        m = Schelling(SchellingScenario(width=width, height=height, density=density, minority_pc=0.5,
                                        homophily=homophily, radius=radius, rng=mt_seed))
    rng = m.random
    assert isinstance(rng, ReplayRandom) and m.grid.random is rng
    agents = sorted(m.agents, key=lambda a: a.unique_id)
    assert [a.unique_id for a in agents] == list(range(1, len(agents) + 1))

    def cells():
        return np.array([a.cell.coordinate[0] * height + a.cell.coordinate[1] for a in agents], dtype=np.uint32)

    def happys():
        return np.array([a.happy for a in agents], dtype=np.uint8)

    out = {
        "width": width, "height": height, "homophily": homophily, "radius": radius,
        "philox_seed": np.uint64(philox_seed),
        "atype": np.array([a.type for a in agents], dtype=np.int8),
        "cell0": cells(), "happy0": happys(), "happy_count0": m.happy,
    }
    happy_counts, movers, cell_digest, happy_digest = [], [], [], []
    rng.replay = True
    with activation_context(SchellingAgent, "step"):
        for s in range(1, steps + 1):
            before = cells()
            m.step()
            after = cells()
            happy_counts.append(m.happy)
            movers.append(int((before != after).sum()))
            cell_digest.append(digest(after))
            happy_digest.append(digest(happys()))
            if s in full_steps:
                out[f"cell_{s}"] = after
                out[f"happy_{s}"] = happys()
            assert np.array_equal(m.grid.empty.reshape(-1), np.bincount(after, minlength=width * height) == 0)
    rng.replay = False
    out.update(happy_count=np.array(happy_counts), movers=np.array(movers),
               cell_digest=np.array(cell_digest, dtype=np.uint64),
               happy_digest=np.array(happy_digest, dtype=np.uint64), full_steps=np.array(sorted(full_steps)))
    np.savez_compressed(os.path.join(GOLDEN, name), **out)
    print(name, "agents", len(agents), "happy", happy_counts[:5], "...", happy_counts[-1], "movers", movers[:5])


def main():
    import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    gol_fixture(37, 23, 42, 6, "gol_37x23.npz")
    gol_fixture(64, 48, 7, 20, "gol_64x48.npz")
    schelling_fixture("schelling_20x20_r1.npz", 20, 20, 0.8, 0.4, 1, 30, 42, 2024, full_steps=range(1, 31))
    schelling_fixture("schelling_16x48_r2.npz", 16, 48, 0.9, 0.6, 2, 12, 5, 99, full_steps=range(1, 13))
    # BASELINE.json configs[0]: 100x100, density 0.8, 100 steps (h=0.4, r=1)
    schelling_fixture("schelling_100x100_c1.npz", 100, 100, 0.8, 0.4, 1, 100, 42, 7, full_steps=[1, 2, 10, 50, 100])
    # the reference's own "large" benchmark variant (benchmarks/configurations.py:42-49)
    schelling_fixture("schelling_100x100_large.npz", 100, 100, 0.8, 1, 2, 10, 42, 11, full_steps=[1, 5, 10])


if __name__ == "__main__":
    main()
