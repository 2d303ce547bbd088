"""Developer micro-benchmark: Epstein civil violence (csrc/epstein.cu) on a side x side torus, synthetic initial state."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mesa_b200.examples.epstein_civil_violence import EpsteinCivilViolence, EpsteinScenario


def make(side, vision=7, moore=False, legitimacy=0.8, max_jail_term=1000, seed=1):
    rs = np.random.default_rng(seed)
    u = rs.random(side * side)
    occupied = np.flatnonzero(u < 0.774)
    kind = (u[occupied] >= 0.7).astype(np.uint8)
    hardship = np.where(kind == 0, rs.random(len(kind)), 0.0)
    risk = np.where(kind == 0, rs.random(len(kind)), 0.0)
    sc = EpsteinScenario(citizen_vision=vision, cop_vision=vision, legitimacy=legitimacy, max_jail_term=max_jail_term,
                         grid_type="Moore" if moore else "Von Neumann", rng=seed)
    return EpsteinCivilViolence(side, side, sc, initial_state={"kind": kind, "cell": occupied, "hardship": hardship, "risk_aversion": risk})


def bench(side, steps=10, **kw):
    m = make(side, **kw)
    for _ in range(3):
        m.agents.shuffle_do("step")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        m.agents.shuffle_do("step")          # the activation alone (no collector)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"side": side, "agents": m.agents.n, "ms_per_step": dt * 1e3, "agent_updates_per_s": m.agents.n / dt,
            "counts": [m.ACTIVE, m.QUIET, m.ARRESTED], **{k: v for k, v in kw.items()}}


if __name__ == "__main__":
    sides = [int(s) for s in sys.argv[1:]] or [40, 256, 1024, 2048]
    for side in sides:
        print(json.dumps(bench(side)), flush=True)
    print(json.dumps(bench(1024, legitimacy=0.6, max_jail_term=30)), flush=True)
    print(json.dumps(bench(1024, vision=3, moore=True, legitimacy=0.6, max_jail_term=30)), flush=True)
