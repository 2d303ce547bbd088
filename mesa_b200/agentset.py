"""DeviceAgentSet: structure-of-arrays agent populations behind the AbstractAgentSet surface.

Mirrors mesa/agentset.py:27-355 (AbstractAgentSet): `do` / `shuffle_do` / `map` / `get` / `set` /
`agg` / `select` / `shuffle` / `sort` / `__len__` / `__iter__` / `__contains__`, with `do` and
`shuffle_do` returning `self` (agentset.py:756-787).  Agents are rows of device columns
(`agents.columns[name]` is a [n] tensor); `agent.attr` on a proxy reads one element (debug speed).

`do("name")` / `shuffle_do("name")` do not call Python per agent: they launch the batch operator
registered for (agent type, "name").  Unregistered names raise -- there is no per-agent CPU
fallback.  `shuffle_do` consumes one activation epoch: the order is the ascending Philox
priority64(model seed, epoch, agent) (csrc/philox.cuh), which replaces `random.shuffle`
(agentset.py:777) and can be replayed into the reference (oracle/replay.py).
"""
from __future__ import annotations

from typing import Callable

import torch

_OPERATORS: dict[tuple[type, str], Callable] = {}


def batch_operator(agent_type: type, name: str):
    """Register `fn(agentset, *, epoch=None, **kwargs)` as the device implementation of `agent_type.name`."""

    def deco(fn):
        _OPERATORS[(agent_type, name)] = fn
        return fn

    return deco


def _lookup(agent_type: type, name: str) -> Callable:
    for klass in agent_type.__mro__:
        fn = _OPERATORS.get((klass, name))
        if fn is not None:
            return fn
    raise NotImplementedError(
        f"no batch operator registered for {agent_type.__name__}.{name}: mesa_b200 runs agent methods as "
        "device kernels (mesa_b200.agentset.batch_operator) and has no per-agent CPU fallback"
    )


class DeviceAgent:
    """Base class of device agent types; instances are row proxies handed out on demand."""

    __slots__ = ("_set", "_index")

    def __init__(self, agentset: "DeviceAgentSet", index: int):
        object.__setattr__(self, "_set", agentset)
        object.__setattr__(self, "_index", int(index))

    @property
    def unique_id(self) -> int:  # agent.py:51-71: ids count from 1 in registration order
        return self._index + 1

    @property
    def model(self):
        return self._set.model

    def __getattr__(self, name):
        cols = object.__getattribute__(self, "_set").columns
        if name in cols:
            v = cols[name][self._index]
            return v.item() if v.dim() == 0 else v.cpu().numpy()
        raise AttributeError(name)

    def __setattr__(self, name, value):
        cols = self._set.columns
        if name in cols:
            cols[name][self._index] = torch.as_tensor(value, device=cols[name].device, dtype=cols[name].dtype)
            return
        raise AttributeError(name)

    def __eq__(self, other):
        return isinstance(other, DeviceAgent) and other._set is self._set and other._index == self._index

    def __hash__(self):
        return hash((id(self._set), self._index))

    def __repr__(self):
        return f"<{type(self).__name__} {self.unique_id}>"


class DeviceAgentSet:
    """An ordered set of `n` agents of one type stored column-wise on the device."""

    def __init__(self, model, agent_type: type, n: int, columns: dict[str, torch.Tensor] | None = None,
                 random=None, index: torch.Tensor | None = None):
        self.model = model
        self.agent_type = agent_type
        self.n = int(n)
        # views share the parent's column mapping (same object), so membership is an identity test
        self.columns: dict[str, torch.Tensor] = columns if columns is not None else {}
        self.random = random if random is not None else getattr(model, "random", None)
        # `index` selects / orders rows for views produced by select / shuffle / sort (None = all, in order)
        self.index = index

    # -- collection protocol ------------------------------------------------
    def __len__(self) -> int:
        return self.n if self.index is None else int(self.index.numel())

    def __iter__(self):
        if self.index is None:
            return (self.agent_type(self, i) for i in range(self.n))
        return (self.agent_type(self, i) for i in self.index.cpu().tolist())

    def __contains__(self, agent) -> bool:
        if not isinstance(agent, DeviceAgent) or agent._set.columns is not self.columns:
            return False
        if self.index is None:
            return 0 <= agent._index < self.n
        return bool((self.index == agent._index).any().item())

    def __getitem__(self, item):
        if isinstance(item, slice):
            rng = range(len(self))[item]
            return [self[i] for i in rng]
        if item < 0:
            item += len(self)
        if not 0 <= item < len(self):
            raise IndexError(item)
        row = item if self.index is None else int(self.index[item].item())
        return self.agent_type(self, row)

    def __repr__(self) -> str:
        return f"<{type(self).__name__} ({len(self)} agents)>"

    def add(self, agent):
        raise NotImplementedError("device populations are fixed-size in this release (SURVEY §8f)")

    discard = remove = add

    # -- batch activation ---------------------------------------------------
    def _require_full(self, what: str):
        if self.index is not None:
            raise NotImplementedError(f"{what} on a selected / reordered view is not supported on the device path")

    def do(self, method, *args, **kwargs) -> "DeviceAgentSet":
        """agentset.py:756-770: call `method` on every agent (registration order) -- one kernel launch."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.do takes the name of a registered batch operator, not a callable")
        self._require_full("do")
        _lookup(self.agent_type, method)(self, *args, **kwargs)
        return self

    def shuffle_do(self, method, *args, **kwargs) -> "DeviceAgentSet":
        """agentset.py:772-787: activate every agent once in a random order (Philox priority order)."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.shuffle_do takes the name of a registered batch operator")
        self._require_full("shuffle_do")
        epoch = self.model._next_epoch()
        _lookup(self.agent_type, method)(self, *args, epoch=epoch, **kwargs)
        return self

    def map(self, method, *args, **kwargs):
        """agentset.py:789-804: like do, returning the operator's per-agent result tensor."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.map takes the name of a registered batch operator")
        self._require_full("map")
        return _lookup(self.agent_type, method)(self, *args, **kwargs)

    # -- columns --------------------------------------------------------------
    def _col(self, name: str) -> torch.Tensor:
        try:
            col = self.columns[name]
        except KeyError:
            raise AttributeError(f"agents have no attribute {name!r}") from None
        return col if self.index is None else col[self.index]

    def get(self, attr_names, handle_missing: str = "error", default_value=None):
        """agentset.py:171-227: attribute column(s) as device tensors (bulk copy instead of a Python loop)."""
        if isinstance(attr_names, str):
            if attr_names not in self.columns and handle_missing == "default":
                return torch.full((len(self),), default_value)
            return self._col(attr_names)
        return [self.get(a, handle_missing, default_value) for a in attr_names]

    def set(self, attr_name: str, value) -> "DeviceAgentSet":
        col = self.columns[attr_name]
        if self.index is None:
            col[:] = torch.as_tensor(value, device=col.device, dtype=col.dtype)
        else:
            col[self.index] = torch.as_tensor(value, device=col.device, dtype=col.dtype)
        return self

    def agg(self, attribute: str, func):
        """agentset.py:137-169: aggregate a column; `func` gets the device tensor (e.g. torch.sum) or a name."""
        col = self._col(attribute)
        if isinstance(func, str):
            from . import ops

            return ops.layer_reduce(col.contiguous(), func).item()
        if isinstance(func, (list, tuple)):
            return [self.agg(attribute, f) for f in func]
        return func(col)

    def select(self, filter_func=None, at_most=float("inf"), inplace: bool = False, agent_type=None):
        """agentset.py:66-135 with a column predicate: `filter_func(columns) -> bool mask [n]`."""
        idx = torch.arange(self.n, device=self._device()) if self.index is None else self.index
        if agent_type is not None and not issubclass(self.agent_type, agent_type):
            idx = idx[:0]
        if filter_func is not None:
            mask = filter_func(self.columns)
            idx = idx[mask[idx]]
        if at_most != float("inf"):
            k = int(at_most) if at_most >= 1 else int(len(idx) * at_most)
            idx = idx[:k]
        if inplace:
            self.index = idx
            return self
        return DeviceAgentSet(self.model, self.agent_type, self.n, self.columns, self.random, idx)

    def shuffle(self, inplace: bool = False):
        """agentset.py:415-437 with the Philox activation order of a fresh epoch."""
        from . import rng

        epoch = self.model._next_epoch()
        if self.index is None and self._device().type == "cuda":
            from . import ops

            idx = ops.activation_order(self.model.philox_seed, epoch, self.n, self._device()).to(torch.int64)
        else:
            base = torch.arange(self.n, device=self._device()) if self.index is None else self.index
            pri = rng.priority64(self.model.philox_seed, epoch, base)
            idx = base[torch.argsort(pri.view(torch.int64) ^ (-2**63), stable=True)]
        if inplace:
            self.index = idx
            return self
        return DeviceAgentSet(self.model, self.agent_type, self.n, self.columns, self.random, idx)

    def sort(self, key, ascending: bool = False, inplace: bool = False):
        if not isinstance(key, str):
            raise NotImplementedError("DeviceAgentSet.sort takes a column name")
        base = torch.arange(self.n, device=self._device()) if self.index is None else self.index
        order = torch.argsort(self.columns[key][base], stable=True, descending=not ascending)
        idx = base[order]
        if inplace:
            self.index = idx
            return self
        return DeviceAgentSet(self.model, self.agent_type, self.n, self.columns, self.random, idx)

    def to_list(self):
        return list(self)

    def _device(self):
        for c in self.columns.values():
            return c.device
        return torch.device("cuda")


try:  # make isinstance(x, mesa.agentset.AbstractAgentSet) hold when mesa is importable
    from mesa.agentset import AbstractAgentSet as _Abstract  # type: ignore

    _Abstract.register(DeviceAgentSet)
except Exception:  # pragma: no cover
    pass
