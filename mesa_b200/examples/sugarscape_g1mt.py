"""Sugarscape G1MT on the device (mirror of mesa/examples/advanced/sugarscape_g1mt/model.py:30-134, agents.py:20-301): sugar
and spice regrow, every trader moves to the best empty cell it can see, eats, starves when it runs out of either, and then
trades sugar for spice with the traders it can see until nobody gains.

The reference model's own step runs unchanged --

    self.grid.sugar[:] = np.minimum(self.grid.sugar + 1, self.sugar_distribution)      # and spice
    self.agents.shuffle_do("step")
    if self.enable_trade: self.agents.shuffle_do("trade_with_neighbors")

-- on one DYNAMIC device population (the starved are compacted away after the first activation, `unique_id`s persist).  The
sequential semantics of both activations are computed exactly, each by one launch of a rank-ordered dependency graph
(csrc/sugarscape.cu); the same seed gives the reference's initial traders, and with the replayed Philox order / draws the
reference's trajectory (tests/golden/sugarscape_*.npz: ids, cells, sugar and spice of every trader, both layers, every trade
with its price, bit for bit; the DataCollector's trade volume and geometric-mean price).

The landscape is the caller's (`sugar_map`: a [width, height] array or a whitespace-separated text file like the reference's
`sugar-map.txt`); without one a synthetic two-hill landscape with levels 0..4 is used.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _native as N
from ..agentset import DeviceAgent, DeviceAgentSet, batch_operator
from ..datacollection import DeviceDataCollector
from ..model import DeviceModel
from ..space import DeviceVonNeumannGrid
from ._scenario import Scenario


class SugarScapeScenario(Scenario):
    """model.py:30-40."""

    initial_population: int = 200
    endowment_min: int = 25
    endowment_max: int = 50
    metabolism_min: int = 1
    metabolism_max: int = 5
    vision_min: int = 1
    vision_max: int = 5
    enable_trade: bool = True


class Trader(DeviceAgent):
    """agents.py:20-45."""


def von_neumann_bfs_offsets(radius: int) -> np.ndarray:
    """(dx, dy) of Cell._neighborhood(radius) on a von Neumann grid in the reference's order (cell.py:240-276; connection
    order grid.py:478-482), centre excluded.  On a CLAMPED von Neumann grid a cell's neighbourhood is this list with the
    absent cells dropped (checked against the live grid), and the list of radius r is a prefix of the list of radius r + 1."""
    conn = [(-1, 0), (0, -1), (0, 1), (1, 0)]
    visited = {(0, 0)}
    layer, found = [], []
    for c in conn:
        visited.add(c)
        layer.append(c)
        found.append(c)
    for _ in range(radius - 1):
        nxt = []
        for c in layer:
            for d in conn:
                nb = (c[0] + d[0], c[1] + d[1])
                if nb not in visited:
                    visited.add(nb)
                    nxt.append(nb)
                    found.append(nb)
        layer = nxt
    return np.asarray(found, dtype=np.int16).reshape(-1, 2)


def welfare_power_table(max_metabolism: int, max_amount: int) -> np.ndarray:
    """table[ms - 1][mp - 1][x] = x ** (ms / (ms + mp)) for x = 0 .. max_amount - 1, evaluated by THIS host's libm through
    `math.pow` -- the function the reference's `sugar ** (self.metabolism_sugar / m_total)` ends in (agents.py:57-70).  The
    kernels look the welfare up here, so the strict welfare comparisons of the trade criteria (agents.py:138-141) see the
    reference's bits even on exact mathematical ties."""
    import math

    table = np.empty((max_metabolism, max_metabolism, max_amount), dtype=np.float64)
    for ms in range(1, max_metabolism + 1):
        for mp in range(1, max_metabolism + 1):
            e = ms / (ms + mp)
            table[ms - 1, mp - 1] = [math.pow(x, e) for x in range(max_amount)]
    return table


def synthetic_sugar_map(width: int, height: int) -> np.ndarray:
    """Two sugar hills with levels 0..4 (the shape of the classic landscape; NOT the reference's data file)."""
    x, y = np.meshgrid(np.arange(width), np.arange(height), indexing="ij")
    scale = min(width, height) / 50.0

    def hill(cx, cy):
        d = np.sqrt((x - cx * width) ** 2 + (y - cy * height) ** 2) / scale
        return np.select([d < 5, d < 10, d < 15, d < 21], [4, 3, 2, 1], 0)

    return np.maximum(hill(0.3, 0.7), hill(0.7, 0.3)).astype(np.float64)


@batch_operator(Trader, "step")
def _trader_step(agents: DeviceAgentSet, epoch: int, **_):
    agents.model._run_activation(epoch)


@batch_operator(Trader, "trade_with_neighbors")
def _trade(agents: DeviceAgentSet, epoch: int, **_):
    agents.model._run_trade(epoch)


def geometric_mean(prices) -> float:
    """model.py:20-27 (the reference's own expression, on the host: same values in the same order, same result)."""
    if len(prices) == 0:
        return -1
    return np.exp(np.log(prices).mean())


def _report_traders(m):
    return len(m.agents)


def _report_volume(m):
    return len(m._trades[2])


def _report_price(m):
    return geometric_mean(list(m._trades[2]))


def _report_network(agents):
    """agent reporter "Trade Network" (model.py:70): the partner lists as a lazy ragged column (built when the DataFrame is)."""
    from ..datacollection import RaggedColumn

    m = agents.model
    counts = np.bincount(m._trade_rows, minlength=len(agents)) if len(m._trade_rows) else np.zeros(len(agents), dtype=np.int64)
    return RaggedColumn(np.concatenate([[0], np.cumsum(counts)]), m._trades[1])


class SugarscapeG1mt(DeviceModel):
    """Sugarscape with traders (model.py:43-134), movement phase."""

    def __init__(self, scenario: SugarScapeScenario = SugarScapeScenario, device="cuda", sugar_map=None, width: int = 50,
                 height: int = 50, initial_state: dict | None = None, pow_table_amounts: int = 16384):
        if isinstance(scenario, type):
            scenario = scenario()
        super().__init__(scenario=scenario)
        sc = scenario
        self.enable_trade = bool(sc.enable_trade)
        if isinstance(sugar_map, str):
            sugar_map = np.genfromtxt(sugar_map)                       # model.py:74
        if sugar_map is not None:
            sugar_map = np.asarray(sugar_map, dtype=np.float64)
            width, height = sugar_map.shape
        else:
            sugar_map = synthetic_sugar_map(width, height)
        self.width, self.height = int(width), int(height)
        self.grid = DeviceVonNeumannGrid((self.width, self.height), torus=False, random=self.random, device=device)   # model.py:58-60
        dev = self.grid.device
        self.sugar_distribution = torch.as_tensor(sugar_map).to(dev).contiguous()
        self.spice_distribution = torch.as_tensor(np.flip(sugar_map, 1).copy()).to(dev).contiguous()              # model.py:75
        self.grid.add_property_layer("sugar", self.sugar_distribution.clone())
        self.grid.add_property_layer("spice", self.spice_distribution.clone())
        if initial_state is None:
            initial_state = self._reference_initial_state(sc)
        st = initial_state
        n = len(st["cell"])
        cols = {"cell": torch.as_tensor(np.asarray(st["cell"])).to(torch.int32).to(dev).contiguous(),
                "sugar": torch.as_tensor(np.asarray(st["sugar"], dtype=np.float64)).to(dev).contiguous(),
                "spice": torch.as_tensor(np.asarray(st["spice"], dtype=np.float64)).to(dev).contiguous(),
                "metabolism_sugar": torch.as_tensor(np.asarray(st["metabolism_sugar"])).to(torch.int32).to(dev).contiguous(),
                "metabolism_spice": torch.as_tensor(np.asarray(st["metabolism_spice"])).to(torch.int32).to(dev).contiguous(),
                "vision": torch.as_tensor(np.asarray(st["vision"])).to(torch.int32).to(dev).contiguous()}
        vmax = int(cols["vision"].max().item()) if n else 1
        if not 1 <= vmax <= 7 or (n and int(cols["vision"].min().item()) < 1):
            raise NotImplementedError("the device model takes visions of 1 to 7 cells")
        agents = DeviceAgentSet(self, Trader, n, cols, self.random)
        ids = st.get("unique_id")
        agents._pop.ids = (torch.as_tensor(np.asarray(ids)).to(torch.int64).to(dev) if ids is not None else
                           torch.arange(1, n + 1, dtype=torch.int64, device=dev))
        self.agents = agents
        self.grid.attach_population(agents)
        self._vmax = vmax
        self._voff = torch.as_tensor(von_neumann_bfs_offsets(vmax)).to(dev).contiguous()
        reach = 2 * vmax
        woff = [(dx, dy) for dx in range(-reach, reach + 1) for dy in range(-reach + abs(dx), reach - abs(dx) + 1)]
        self._woff = torch.as_tensor(np.asarray(woff, dtype=np.int16)).to(dev).contiguous()
        self._count = torch.zeros(self.width * self.height, dtype=torch.int32, device=dev)
        self._status = torch.zeros(3, dtype=torch.int32, device=dev)
        mmax = int(max(cols["metabolism_sugar"].max().item(), cols["metabolism_spice"].max().item())) if n else 1
        # (the table is mmax^2 x amounts doubles from mmax^2 x amounts libm calls: metabolisms beyond 8 use CUDA's pow and count)
        self._pow_m, self._pow_x = min(mmax, 8), int(pow_table_amounts)
        self._pow = torch.as_tensor(welfare_power_table(self._pow_m, self._pow_x)).to(dev).contiguous()
        self.inexact_welfares = 0     # pow calls that missed the table so far (amounts >= pow_table_amounts): CUDA pow, <= 2 ulp
        self._ws = None
        self._regrow = False
        self.last_deaths = 0
        self._rebuild_count()
        self.grid.property_layers.hooks["empty"] = self._sync_empty
        self._trades = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))   # this step's trades
        self._trade_rows = np.zeros(0, np.int64)
        self._rec = None
        self.datacollector = DeviceDataCollector(                           # model.py:62-71 (module-level functions: models pickle)
            model_reporters={"#Traders": _report_traders, "Trade Volume": _report_volume, "Price": _report_price},
            agent_reporters={"Trade Network": _report_network})
        self.running = True

    def _reference_initial_state(self, sc) -> dict:
        """model.py:80-116 on the model's own two generators (reference_init.sugarscape_initial_state): the same seed gives
        the reference's traders."""
        from ..reference_init import sugarscape_initial_state

        return sugarscape_initial_state(self.random, self.rng, self.width * self.height, sc.initial_population, sc.endowment_min,
                                        sc.endowment_max, sc.metabolism_min, sc.metabolism_max, sc.vision_min, sc.vision_max)

    def _rebuild_count(self):
        dev = self.grid.device
        with torch.cuda.device(dev):
            N.check(N.lib().mb_sugarscape_build_count(N.ptr(self.agents.columns["cell"]), self.agents.n, N.ptr(self._count),
                                                      self.width * self.height, N.stream_ptr(dev)))

    def _sync_empty(self):
        dict.__getitem__(self.grid.property_layers, "empty").copy_((self._count == 0).reshape(self.width, self.height))

    def _run_activation(self, epoch: int):
        a, dev = self.agents, self.grid.device
        n, ncells = a.n, self.width * self.height
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_sugarscape_workspace_bytes(n, ncells, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        c = a.columns
        dead = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        uid = a.unique_ids().contiguous()
        layers = self.grid.property_layers
        regrow, self._regrow = self._regrow, False
        with torch.cuda.device(dev):
            N.check(N.lib().mb_sugarscape_step(
                N.ptr(c["cell"]), N.ptr(c["sugar"]), N.ptr(c["spice"]), N.ptr(c["metabolism_sugar"]), N.ptr(c["metabolism_spice"]),
                N.ptr(c["vision"]), N.ptr(uid), N.ptr(dead), n, N.ptr(layers["sugar"]), N.ptr(self.sugar_distribution),
                N.ptr(layers["spice"]), N.ptr(self.spice_distribution), N.ptr(self._count), self.width, self.height,
                N.ptr(self._voff), self._vmax, N.ptr(self._woff), int(self._woff.shape[0]), int(regrow), N.ptr(self._pow),
                self._pow_m, self._pow_x, self.philox_seed, epoch, N.ptr(self._status), N.ptr(self._ws), self._ws.numel(),
                N.stream_ptr(dev)))
        status, deaths, missed = self._status.tolist()
        self.inexact_welfares += int(missed)
        if status == 2:
            raise NotImplementedError("sugarscape: more than 4 traders started the step in one cell")
        if status:
            raise RuntimeError("sugarscape: a dependency wait ran into its spin limit; the step is not valid")
        self.last_deaths = int(deaths)
        if deaths:
            a.remove_agents(dead[:n].bool())          # agents.py:273-281 `self.remove()` (stable: the survivors keep their order)

    def _run_trade(self, epoch: int):
        a, dev = self.agents, self.grid.device
        n, ncells = a.n, self.width * self.height
        nbytes = ctypes.c_size_t(0)
        N.check(N.lib().mb_sugarscape_workspace_bytes(n, ncells, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
        cap = max(64 * n, 1024)
        if self._rec is None or self._rec[0].numel() < cap:
            self._rec = (torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int64, device=dev),
                         torch.empty(cap, dtype=torch.float64, device=dev))
        c = a.columns
        uid = a.unique_ids().contiguous()
        with torch.cuda.device(dev):
            N.check(N.lib().mb_sugarscape_trade(
                N.ptr(c["cell"]), N.ptr(c["sugar"]), N.ptr(c["spice"]), N.ptr(c["metabolism_sugar"]), N.ptr(c["metabolism_spice"]),
                N.ptr(c["vision"]), N.ptr(uid), n, self.width, self.height, N.ptr(self._voff), self._vmax, N.ptr(self._woff),
                int(self._woff.shape[0]), N.ptr(self._pow), self._pow_m, self._pow_x, self.philox_seed, epoch, N.ptr(self._rec[0]),
                N.ptr(self._rec[1]), N.ptr(self._rec[2]), self._rec[0].numel(), N.ptr(self._status), N.ptr(self._ws),
                self._ws.numel(), N.stream_ptr(dev)))
        status, ntrades, missed = self._status.tolist()
        self.inexact_welfares += int(missed)
        if status == 2:
            raise NotImplementedError("sugarscape: more than 4 traders stand in one cell")
        if status:
            raise RuntimeError("sugarscape: a dependency wait ran into its spin limit; the step is not valid")
        if ntrades > self._rec[0].numel():
            raise RuntimeError("sugarscape: more trades than the record buffer holds (the amounts are right, the record is not)")
        rows = self._rec[0][:ntrades].cpu().numpy().astype(np.int64)
        order = np.argsort(rows, kind="stable")          # model.agents order; a trader's trades stay in the order it made them
        ids = uid.cpu().numpy()
        self._trade_rows = rows[order]
        self._trades = (ids[self._trade_rows], self._rec[1][:ntrades].cpu().numpy()[order], self._rec[2][:ntrades].cpu().numpy()[order])

    def trade_records(self):
        """This step's trades as (trader unique ids, partner unique ids, prices), flattened in `model.agents` order -- what
        `[a.trade_partners for a in model.agents]` / `[a.prices ...]` hold in the reference (agents.py:196-201)."""
        return self._trades

    def trade_partners(self) -> np.ndarray:
        """Per trader (rows of `model.agents`) the list of partner ids of this step (agents.py:201 `trade_partners`); the
        traders that did not trade share one empty list."""
        n = len(self.agents)
        out = np.empty(n, dtype=object)
        out[:] = [[]] * n
        rows = self._trade_rows
        if len(rows):
            starts = np.flatnonzero(np.r_[True, rows[1:] != rows[:-1]])
            ends = np.r_[starts[1:], len(rows)]
            partners = self._trades[1].tolist()
            for r, lo, hi in zip(rows[starts].tolist(), starts.tolist(), ends.tolist()):
                out[r] = partners[lo:hi]
        return out

    def step(self):
        """model.py:118-131."""
        self._trades = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64))     # agents.py:284-285
        self._trade_rows = np.zeros(0, np.int64)
        self._regrow = True                                # model.py:123-124, fused into the activation's launch sequence
        self.agents.shuffle_do("step")
        if self.enable_trade:
            self.agents.shuffle_do("trade_with_neighbors")
        self.datacollector.collect(self)

    def run_model(self, step_count: int = 1000):
        for _ in range(step_count):
            self.step()
