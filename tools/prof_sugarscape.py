import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mesa_b200.examples.sugarscape_g1mt import SugarScapeScenario, SugarscapeG1mt
side, n = 512, 40000
rs = np.random.default_rng(1)
cells = rs.choice(side * side, size=n, replace=False)
init = {"cell": cells, "sugar": rs.integers(25, 51, n), "spice": rs.integers(25, 51, n), "metabolism_sugar": rs.integers(1, 6, n),
        "metabolism_spice": rs.integers(1, 6, n), "vision": rs.integers(1, 6, n)}
m = SugarscapeG1mt(SugarScapeScenario(enable_trade=True, rng=1), width=side, height=side, initial_state=init)
for _ in range(4):
    m.step()
torch.cuda.synchronize()
