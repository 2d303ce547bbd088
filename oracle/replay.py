"""ReplayRandom: drive the UNMODIFIED reference with the GPU's Philox draws -- TEST INFRASTRUCTURE.

The reference draws from one sequential MT19937 stream (mesa/model.py:129,
`self.random = random.Random(seed)`) that the grid, every cell and every AgentSet share.
The device path instead keys every number on (seed, epoch, agent, draw#) (oracle/philox.py).
To compare the two on identical inputs the reference is run with a `random.Random` subclass
that returns those keyed numbers:

* `shuffle(agents)`      -> order by priority64(seed, epoch, unique_id - 1); epoch += 1
                            (replaces Fisher-Yates of mesa/agentset.py:777; any permutation is a
                            valid activation order, this one is the order the GPU uses);
* `getrandbits(k)`       -> the next keyed draw of the agent currently being activated
                            (`random.choice` -> `_randbelow` -> `getrandbits`, CPython
                            Lib/random.py:242-250, so rejection sampling is reproduced by the
                            unmodified stdlib code);
* outside a replayed step it behaves as the stock MT19937, so model construction is untouched.

The "current agent" is tracked by wrapping the agent class's step method (class attribute
patch, `activation_context`); the model, grid, cell and agent-set code run unmodified.
SURVEY.md A.4 describes the injection point (`mesa.model.random`).
"""
from __future__ import annotations

import contextlib
import os
import random
import sys
import types

from . import philox

REFERENCE_ROOT = "/root/reference"


def import_reference():
    """Import the live reference (build container only; the GPU box has no /root/reference)."""
    if "mesa" not in sys.modules:
        if not os.path.isdir(os.path.join(REFERENCE_ROOT, "mesa")):
            raise ImportError("reference not available at /root/reference")
        sys.path.insert(0, REFERENCE_ROOT)
    import mesa  # noqa: F401

    return sys.modules["mesa"]


class ReplayRandom(random.Random):
    """random.Random whose draws are the keyed Philox numbers once `replay` is switched on."""

    def __new__(cls, *args, **kwargs):
        return super().__new__(cls)

    def __init__(self, mt_seed=None, philox_seed: int = 0):
        super().__init__(mt_seed)
        self.philox_seed = int(philox_seed)
        self.replay = False
        self.epoch = 0          # number of shuffles replayed so far
        self.draw_epoch = 0     # epoch the current activations belong to
        self.agent = None       # 0-based agent index being activated
        self.counter = 0        # draw# within the activation
        self.log = []           # (epoch, agent, counter, k, value) for debugging

    # -- activation order -------------------------------------------------
    def shuffle(self, x):
        if not self.replay:
            return super().shuffle(x)
        ids = [a.unique_id - 1 for a in x]
        pri = philox.priority64(self.philox_seed, self.epoch, ids)
        order = sorted(range(len(x)), key=lambda i: int(pri[i]))
        x[:] = [x[i] for i in order]
        self.draw_epoch = self.epoch
        self.epoch += 1

    # -- per-agent draws ---------------------------------------------------
    def getrandbits(self, k):
        if not self.replay or self.agent is None:
            return super().getrandbits(k)
        if not 0 < k <= 64:
            raise ValueError("replayed getrandbits supports 1..64 bits")
        v = philox.getrandbits(self.philox_seed, self.draw_epoch, self.agent, self.counter, k)
        self.counter += 1
        return v

    def random(self):
        if not self.replay or self.agent is None:
            return super().random()
        v = float(philox.uniform01(self.philox_seed, self.draw_epoch, self.agent, self.counter))
        self.counter += 1
        return v

    def enter_agent(self, index: int):
        self.agent = int(index)
        self.counter = 0

    def leave_agent(self):
        self.agent = None


@contextlib.contextmanager
def patched_model_random(philox_seed: int):
    """Make `Model.__init__` build a ReplayRandom instead of random.Random."""
    import_reference()
    import mesa.model as mm

    orig = mm.random
    mm.random = types.SimpleNamespace(
        Random=lambda seed=None: ReplayRandom(seed, philox_seed=philox_seed)
    )
    try:
        yield
    finally:
        mm.random = orig


@contextlib.contextmanager
def activation_context(agent_cls, method: str = "step"):
    """Wrap `agent_cls.method` so the model's ReplayRandom knows which agent is drawing."""
    orig = getattr(agent_cls, method)

    def wrapped(self, *args, **kwargs):
        rng = self.model.random
        rng.enter_agent(self.unique_id - 1)
        try:
            return orig(self, *args, **kwargs)
        finally:
            rng.leave_agent()

    setattr(agent_cls, method, wrapped)
    try:
        yield
    finally:
        setattr(agent_cls, method, orig)
