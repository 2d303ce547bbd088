"""Developer micro-benchmark: Sugarscape G1MT (csrc/sugarscape.cu) on synthetic landscapes, movement and trade."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mesa_b200.examples.sugarscape_g1mt import SugarScapeScenario, SugarscapeG1mt


def run(side, n, trade, steps=10):
    rs = np.random.default_rng(1)
    cells = rs.choice(side * side, size=n, replace=False)
    init = {"cell": cells, "sugar": rs.integers(25, 51, n), "spice": rs.integers(25, 51, n), "metabolism_sugar": rs.integers(1, 6, n),
            "metabolism_spice": rs.integers(1, 6, n), "vision": rs.integers(1, 6, n)}
    m = SugarscapeG1mt(SugarScapeScenario(enable_trade=trade, rng=1), width=side, height=side, initial_state=init)
    for _ in range(2):
        m.step()
    torch.cuda.synchronize(); t0 = time.perf_counter(); acted = 0; trades = 0
    for _ in range(steps):
        acted += len(m.agents) * (2 if trade else 1)
        m.step()
        trades += len(m.trade_records()[2])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return {"side": side, "traders_start": n, "traders_end": len(m.agents), "trade": trade, "ms_per_step": dt / steps * 1e3,
            "agent_updates_per_s": acted / dt, "trades_per_step": trades / steps, "inexact_welfares": m.inexact_welfares}


if __name__ == "__main__":
    for side, n in ((50, 200), (512, 40000), (2048, 600000)):
        for trade in (False, True):
            print(json.dumps(run(side, n, trade)), flush=True)
