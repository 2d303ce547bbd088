"""Minimal stand-in for mesa.experimental.scenarios.Scenario (scenario.py:112-177): annotated class
attributes are the defaults, keyword arguments override them, `rng` is the seed."""
from __future__ import annotations


class Scenario:
    rng = None

    def __init__(self, **kwargs):
        names = {}
        for klass in reversed(type(self).__mro__):
            names.update(getattr(klass, "__annotations__", {}))
        for k, v in kwargs.items():
            if k != "rng" and k not in names:
                raise TypeError(f"{type(self).__name__} has no parameter {k!r}")
            object.__setattr__(self, k, v)

    def __setattr__(self, key, value):  # frozen after init like the reference (scenario.py:164-171)
        raise TypeError(f"Scenario is frozen; cannot set '{key}'. Create a new Scenario instance instead.")

    def __delattr__(self, key):  # scenario.py:173-177
        raise TypeError(f"Scenario is frozen; cannot delete '{key}'.")
