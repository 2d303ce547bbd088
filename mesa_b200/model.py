"""DeviceModel: the part of mesa.Model the hot path needs (mesa/model.py:85-150, 296-422).

Seeds: like the reference, `rng` / `seed` seed both a stdlib `random.Random` (`model.random`) and a
numpy Generator (`model.rng`), with the reference's own derivation (mesa_b200/reference_init.py), so the same `rng`
gives the same two streams -- and the example models the same initial state -- as in the reference; in addition `model.philox_seed` keys the device draws and
`model._next_epoch()` counts activations (`shuffle_do` / `shuffle` calls) so that every shuffle is a
fresh, replayable permutation (csrc/philox.cuh).
"""
from __future__ import annotations

import random

import numpy as np


class DeviceModel:
    def __init__(self, *args, seed=None, rng=None, scenario=None, **kwargs):
        if scenario is not None and rng is None and seed is None:
            rng = getattr(scenario, "rng", None)
        if rng is None:
            rng = seed
        self.scenario = scenario
        # scenario.py:156-158, 214-221 + model.py:125-129: `rng` seeds the numpy Generator, and the first word of its seed
        # sequence seeds the stdlib generator -- the same two streams the reference model gets from the same `rng`
        from .reference_init import seed_streams

        self.rng, stdlib_seed = seed_streams(rng)
        self.random = random.Random(stdlib_seed)
        self._seed = stdlib_seed
        # the key of the device's Philox streams: the user's integer seed itself when there is one
        user_int = isinstance(rng, (int, np.integer)) and not isinstance(rng, bool)
        self.philox_seed = (int(rng) if user_int else stdlib_seed) & 0xFFFFFFFFFFFFFFFF
        self._epoch = 0
        self.running = True
        self.steps = 0
        self.time = 0.0
        self._agents_by_type: dict = {}
        self._agents = None  # set by the concrete model (a DeviceAgentSet) through the `agents` property
        # model.py:133-140: the user's step() is wrapped so that every call advances time by one unit
        self._user_step = self.step
        self.step = self._wrapped_step

    # model.py:143-148, 221-247: `agents` is the set of all agents, `agents_by_type` maps an agent class to the set of
    # its instances (typed activation, e.g. wolf_sheep/model.py:140-141: agents_by_type[Sheep].shuffle_do("step")).
    # A device model registers one DeviceAgentSet per agent class; assigning `model.agents` registers that set.
    @property
    def agents(self):
        return self._agents

    @agents.setter
    def agents(self, agentset) -> None:
        self._agents = agentset
        if agentset is not None:
            self.register_agentset(agentset)

    def register_agentset(self, agentset) -> None:
        """Make `agentset` (one population of one agent class) reachable through `agents_by_type`."""
        self._agents_by_type[agentset.agent_type] = agentset

    @property
    def agent_types(self) -> list[type]:
        return list(self._agents_by_type.keys())

    @property
    def agents_by_type(self) -> dict:
        return self._agents_by_type

    def remove_all_agents(self) -> None:
        """model.py:307-320: remove every agent of every registered population."""
        for agentset in list(self._agents_by_type.values()):
            agentset.clear()

    def _next_epoch(self) -> int:
        e = self._epoch
        self._epoch += 1
        return e

    def step(self) -> None:  # pragma: no cover - user models override
        """A single step. Fill in here."""

    def _wrapped_step(self) -> None:
        """model.py:152-154: advance time by one unit and run the user's step."""
        self._step_at(self.time + 1.0)

    def _step_at(self, time: float) -> None:
        self.steps += 1
        self.time = time
        self._user_step()

    def _advance_time(self, until: float) -> None:
        """model.py:156-190 for the default schedule (one step per unit of time, at the integer times): run the steps
        that fall in (time, until], then move the clock to `until`."""
        import math

        while True:
            nxt = float(math.floor(self.time) + 1)
            if nxt > until:
                break
            self._step_at(nxt)
        if until > self.time:
            self.time = float(until)

    def run_for(self, duration: float | int) -> None:
        """model.py:395-402: advance the clock by `duration`, running the steps that fall due."""
        self._advance_time(self.time + duration)

    def run_until(self, end_time: float | int) -> None:
        """model.py:404-422: run up to the absolute time `end_time`; a time in the past only warns."""
        if self.time > end_time:
            import warnings

            warnings.warn(f"end_time {end_time} is not larger than current time {self.time}", RuntimeWarning, stacklevel=2)
            return
        self._advance_time(end_time)

    def run_model(self) -> None:
        while self.running:
            self.step()
