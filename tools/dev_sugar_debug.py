import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mesa_b200.examples.sugarscape_g1mt import SugarscapeG1mt, SugarScapeScenario
g = np.load("tests/golden/sugarscape_200_trade.npz")
n, emin, emax, mmin, mmax, vmin, vmax = (int(v) for v in g["scenario"])
sc = SugarScapeScenario(initial_population=n, endowment_min=emin, endowment_max=emax, metabolism_min=mmin, metabolism_max=mmax, vision_min=vmin, vision_max=vmax, enable_trade=True, rng=1)
init = {"unique_id": g["id_0"], "cell": g["cell_0"], "sugar": g["sugar_0"], "spice": g["spice_0"], "metabolism_sugar": g["metabolism_sugar"], "metabolism_spice": g["metabolism_spice"], "vision": g["vision"]}
m = SugarscapeG1mt(sc, sugar_map=g["sugar_max"], initial_state=init)
m.philox_seed = int(g["philox_seed"])
m.step()
t, p, pr = m.trade_records()
want = g["trades_1"]
print("device trades", len(t), "reference", len(want))
k = 0
while k < min(len(t), len(want)) and t[k] == want[k, 0] and p[k] == want[k, 1] and pr[k] == want[k, 2]:
    k += 1
print("first difference at record", k)
for j in range(max(0, k - 3), min(k + 6, max(len(t), len(want)))):
    d = (int(t[j]), int(p[j]), float(pr[j])) if j < len(t) else None
    w = tuple(want[j]) if j < len(want) else None
    print(j, "device", d, "reference", w)
ids = m.agents.unique_ids().cpu().numpy()
sug = m.agents.get("sugar").cpu().numpy(); spi = m.agents.get("spice").cpu().numpy()
bad = np.flatnonzero((sug != g["sugar_1"]) | (spi != g["spice_1"]))
print("rows with wrong amounts", [(int(ids[b]), float(sug[b]), float(g["sugar_1"][b]), float(spi[b]), float(g["spice_1"][b])) for b in bad])
