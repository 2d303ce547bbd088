"""DeviceAgentSet: structure-of-arrays agent populations behind the AbstractAgentSet surface.

Mirrors mesa/agentset.py:27-355 (AbstractAgentSet): `do` / `shuffle_do` / `map` / `get` / `set` /
`agg` / `select` / `groupby` / `shuffle` / `sort` / `__len__` / `__iter__` / `__contains__`, with `do` and
`shuffle_do` returning `self` (agentset.py:756-787).  Agents are rows of device columns
(`agents.columns[name]` is a [n] tensor); `agent.attr` on a proxy reads one element (debug speed).

`do("name")` / `shuffle_do("name")` do not call Python per agent: they launch the batch operator
registered for (agent type, "name").  Unregistered names raise -- there is no per-agent CPU
fallback.  `shuffle_do` consumes one activation epoch: the order is the ascending Philox
priority64(model seed, epoch, agent) (csrc/philox.cuh), which replaces `random.shuffle`
(agentset.py:777) and can be replayed into the reference (oracle/replay.py).

Populations are dynamic (agentset.py:700-711 `add / discard / remove`, agent.py:73-108 `Agent.remove`,
agent.py:115-160 `Agent.create_agents`): `add_agents` appends rows, `remove_agents` / `remove` /
`discard` take rows out with ONE stable compaction shared by every column (csrc/agents.cu), so the
survivors keep their registration order -- the base order of `do` / `shuffle_do` -- and their
`unique_id`.  Row proxies and `select / shuffle / sort` views taken before a removal refer to row
numbers that no longer exist; using them afterwards raises `ReferenceError`.
"""
from __future__ import annotations

from typing import Callable

import torch

_OPERATORS: dict[tuple[type, str], Callable] = {}


def batch_operator(agent_type: type, name: str, views: bool = False):
    """Register `fn(agentset, *, epoch=None, **kwargs)` as the device implementation of `agent_type.name`.

    views=True declares that the operator honours `agentset.index` (the row tensor of a `select / shuffle / sort` view,
    None for the whole population), so `do / shuffle_do / map` may be called on such views like on any reference
    AgentSet (agentset.py:66-135 `select`, :474-518); operators without it only accept whole populations."""

    def deco(fn):
        fn._mb_views = bool(views)
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


def _reduction_name(func) -> str | None:
    """The device reduction a well-known aggregation function stands for (None: call it on the column)."""
    import builtins

    import numpy as np

    table = {builtins.min: "min", builtins.max: "max", builtins.sum: "sum", builtins.len: "len",
             np.min: "min", np.max: "max", np.sum: "sum", np.mean: "mean", np.count_nonzero: "count_nonzero"}
    try:
        return table.get(func)
    except TypeError:  # unhashable callable
        return None


class DeviceAgent:
    """Base class of device agent types; instances are row proxies handed out on demand."""

    __slots__ = ("_set", "_index", "_gen")

    def __init__(self, agentset: "DeviceAgentSet", index: int):
        object.__setattr__(self, "_set", agentset)
        object.__setattr__(self, "_index", int(index))
        object.__setattr__(self, "_gen", agentset._pop.generation)

    def _row(self) -> int:
        """The row this proxy stands for; rows are renumbered by a removal, which invalidates older proxies."""
        if self._gen != self._set._pop.generation:
            raise ReferenceError("agent proxy taken before agents were removed from the population; "
                                 "look the agent up again (rows are renumbered by the compaction)")
        return self._index

    @property
    def unique_id(self) -> int:  # agent.py:51-71, model.py:250-263: ids count from 1 in registration order
        return self._set._unique_id(self._row())

    @property
    def model(self):
        return self._set.model

    def remove(self) -> None:
        """agent.py:73-108: take this agent out of the model's population (KeyError-free like the reference)."""
        self._set.discard(self)

    def __getattr__(self, name):
        cols = object.__getattribute__(self, "_set").columns
        if name in cols:
            v = cols[name][self._row()]
            return v.item() if v.dim() == 0 else v.cpu().numpy()
        raise AttributeError(name)

    def __setattr__(self, name, value):
        cols = self._set.columns
        if name in cols:
            cols[name][self._row()] = torch.as_tensor(value, device=cols[name].device, dtype=cols[name].dtype)
            return
        raise AttributeError(name)

    def __eq__(self, other):
        return (isinstance(other, DeviceAgent) and other._set.columns is self._set.columns
                and other._gen == self._gen and other._index == self._index)

    def __hash__(self):
        return hash((id(self._set.columns), self._gen, self._index))

    def __repr__(self):
        return f"<{type(self).__name__} {self.unique_id}>"

    def __reduce__(self):  # slots + the column-routing __setattr__ rule out the default protocol
        return (type(self), (self._set, self._row()))


class _Population:
    """State shared by a population and its views: the removal generation (bumped whenever rows are renumbered)
    and the unique ids (`ids` stays None while unique_id == row + 1, i.e. until the first removal)."""

    def __init__(self, n: int):
        self.generation = 0
        self.ids: torch.Tensor | None = None
        self.next_id = n + 1  # model.py:250-263: ids are handed out in registration order and never reused


class DeviceAgentSet:
    """An ordered set of `n` agents of one type stored column-wise on the device."""

    def __init__(self, model, agent_type: type, n: int, columns: dict[str, torch.Tensor] | None = None,
                 random=None, index: torch.Tensor | None = None, resizable: bool = True, _pop: _Population | None = None):
        self.model = model
        self.agent_type = agent_type
        self.n = int(n)
        # views share the parent's column mapping (same object), so membership is an identity test
        self.columns: dict[str, torch.Tensor] = columns if columns is not None else {}
        self.random = random if random is not None else getattr(model, "random", None)
        # `index` selects / orders rows for views produced by select / shuffle / sort (None = all, in order)
        self.index = index
        # False for populations whose rows are also keys of other state (grid layers holding agent rows, a
        # continuous space's storage order): those models add / remove through their own space, not here
        self.resizable = bool(resizable)
        self._pop = _pop if _pop is not None else _Population(self.n)
        self._view_gen = self._pop.generation
        # called as hook(agentset, plan) after rows were appended (plan None) or compacted (plan = the
        # ops.CompactionPlan): lets a model keep row-keyed side state (workspaces, layers) in step
        self.resize_hooks: list[Callable] = []

    def _view(self, idx: torch.Tensor) -> "DeviceAgentSet":
        return DeviceAgentSet(self.model, self.agent_type, self.n, self.columns, self.random, idx, self.resizable,
                              self._pop)

    def _check_view(self) -> None:
        if self.index is not None and self._view_gen != self._pop.generation:
            raise ReferenceError("this view was taken before agents were removed from the population; "
                                 "select / shuffle / sort again (rows are renumbered by the compaction)")

    # -- collection protocol ------------------------------------------------
    def __len__(self) -> int:
        return self.n if self.index is None else int(self.index.numel())

    def __iter__(self):
        self._check_view()
        if self.index is None:
            return (self.agent_type(self, i) for i in range(self.n))
        return (self.agent_type(self, i) for i in self.index.cpu().tolist())

    def __contains__(self, agent) -> bool:
        if not isinstance(agent, DeviceAgent) or agent._set.columns is not self.columns:
            return False
        if agent._gen != self._pop.generation:  # a proxy from before a removal names no current agent
            return False
        self._check_view()
        if self.index is None:
            return 0 <= agent._index < self.n
        return bool((self.index == agent._index).any().item())

    def __getitem__(self, item):
        self._check_view()
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

    # -- unique ids -----------------------------------------------------------
    def _unique_id(self, row: int) -> int:
        ids = self._pop.ids
        return row + 1 if ids is None else int(ids[row].item())

    def unique_ids(self) -> torch.Tensor:
        """`unique_id` of every agent of this set, in the set's order (int64, on the device)."""
        self._check_view()
        ids = self._pop.ids
        if ids is None:
            if self.index is None:
                return torch.arange(1, self.n + 1, dtype=torch.int64, device=self._device())
            return self.index.to(torch.int64) + 1
        return ids if self.index is None else ids[self.index]

    def unique_ids_host(self):
        """The same as a numpy array; a population that never shrank answers without touching the device."""
        import numpy as np

        if self._pop.ids is None and self.index is None:
            return np.arange(1, self.n + 1, dtype=np.int64)
        return self.unique_ids().cpu().numpy()

    # -- creation and removal (agentset.py:700-711, agent.py:73-160) --------------------
    def _require_resizable(self, what: str) -> None:
        self._require_full(what)
        if not self.resizable:
            raise NotImplementedError(f"{what}: the rows of this population key other model state (grid layers / "
                                      "space storage); the model owns its size")
        if not isinstance(self.columns, dict) or not all(isinstance(c, torch.Tensor) for c in self.columns.values()):
            raise NotImplementedError(f"{what} needs plain tensor columns")
        for name, col in self.columns.items():
            if col.shape[0] != self.n:
                raise ValueError(f"column {name!r} has {col.shape[0]} rows for {self.n} agents")

    def add_agents(self, k: int, unique_ids=None, **values) -> "DeviceAgentSet":
        """Create `k` agents at the end of the registration order (agent.py:115-160 `create_agents`): each
        keyword names a column and gives one value for all new agents or one per agent (torch broadcasting
        against [k, ...]); unnamed columns start at zero.  Returns the set of the new agents.  `unique_ids` ([k] int64) gives
        the new agents' ids explicitly -- for models whose populations share ONE id counter, as all agent types of a
        reference model do (model.py:250-263: `agent_id_counter`); by default a population counts its own ids."""
        k = int(k)
        if k < 0:
            raise ValueError("cannot create a negative number of agents")
        self._require_resizable("add_agents")
        for name in values:
            if name not in self.columns:
                raise AttributeError(f"agents have no attribute {name!r}")
        old_n, new_n = self.n, self.n + k
        dev = self._device()
        grown = {}
        for name, col in self.columns.items():  # build everything first: a bad value must not leave half a resize
            g = torch.empty((new_n, *col.shape[1:]), dtype=col.dtype, device=col.device)
            g[:old_n].copy_(col)
            if name in values:
                g[old_n:] = torch.as_tensor(values[name], device=col.device).to(col.dtype)
            else:
                g[old_n:].zero_()
            grown[name] = g
        pop = self._pop
        if unique_ids is not None:
            fresh = torch.as_tensor(unique_ids, device=dev).to(torch.int64).reshape(-1)
            if fresh.numel() != k:
                raise ValueError(f"{fresh.numel()} unique ids for {k} new agents")
            if pop.ids is None:
                pop.ids = torch.arange(1, old_n + 1, dtype=torch.int64, device=dev)
            pop.ids = torch.cat([pop.ids, fresh])
            if k:
                pop.next_id = max(pop.next_id, int(fresh.max().item()) + 1)
        else:
            if pop.ids is not None:
                fresh = torch.arange(pop.next_id, pop.next_id + k, dtype=torch.int64, device=pop.ids.device)
                pop.ids = torch.cat([pop.ids, fresh])
            pop.next_id += k
        self.columns.update(grown)
        self.n = new_n
        for hook in self.resize_hooks:
            hook(self, None)
        return self._view(torch.arange(old_n, new_n, dtype=torch.int64, device=dev))

    def remove_agents(self, which) -> "DeviceAgentSet":
        """Remove a batch of agents: `which` is a bool mask [n] (True = remove), a tensor / sequence of row
        numbers, or an iterable of agent proxies.  One stable compaction of every column (csrc/agents.cu): the
        other agents keep their order and their unique_id."""
        from . import ops

        self._require_resizable("remove_agents")
        dev = self._device()
        if isinstance(which, DeviceAgentSet):
            which._check_view()
            which = torch.arange(which.n, device=dev) if which.index is None else which.index
        if hasattr(which, "__array__") and not isinstance(which, torch.Tensor):   # numpy masks / index arrays
            import numpy as np

            which = torch.from_numpy(np.ascontiguousarray(which))
        if isinstance(which, torch.Tensor) and which.dtype == torch.bool:
            if tuple(which.shape) != (self.n,):
                raise ValueError(f"removal mask has shape {tuple(which.shape)}, expected ({self.n},)")
            keep = ~which.to(dev)
        else:
            if not isinstance(which, torch.Tensor):
                rows = []
                for a in which:
                    if isinstance(a, DeviceAgent):
                        if a not in self:
                            raise KeyError(a)
                        a = a._index
                    rows.append(int(a))
                which = torch.as_tensor(rows, dtype=torch.int64)
            rows = which.to(dev, torch.int64).reshape(-1)
            if rows.numel() and (int(rows.min().item()) < 0 or int(rows.max().item()) >= self.n):
                raise IndexError("agent row out of range")
            keep = torch.ones(self.n, dtype=torch.bool, device=dev)
            keep[rows] = False
        plan = ops.CompactionPlan(keep)
        if plan.kept == self.n:
            return self
        pop = self._pop
        ids = pop.ids if pop.ids is not None else torch.arange(1, self.n + 1, dtype=torch.int64, device=dev)
        compacted = {name: plan.apply(col) for name, col in self.columns.items()}
        pop.ids = plan.apply(ids)
        self.columns.update(compacted)
        self.n = plan.kept
        pop.generation += 1
        for hook in self.resize_hooks:
            hook(self, plan)
        return self

    def add(self, agent):
        """agentset.py:700-702.  Device agents are created in batches (`add_agents`); adding an agent that is
        already a member is the no-op it is in the reference."""
        if agent in self:
            return
        raise TypeError("DeviceAgentSet.add takes a current member (no-op); create agents with add_agents(k, **columns)")

    def discard(self, agent):
        """agentset.py:704-707: remove the agent if it is a member."""
        if agent in self:
            self.remove_agents([agent._index])

    def remove(self, agent):
        """agentset.py:709-711: remove the agent, KeyError if it is not a member."""
        if agent not in self:
            raise KeyError(agent)
        self.remove_agents([agent._index])

    def clear(self):
        """MutableSet.clear: remove every agent (model.py:307-320 `remove_all_agents`)."""
        if self.n:
            self.remove_agents(torch.ones(self.n, dtype=torch.bool, device=self._device()))

    # -- batch activation ---------------------------------------------------
    def _require_full(self, what: str):
        if self.index is not None:
            raise NotImplementedError(f"{what} on a selected / reordered view is not supported on the device path")

    def _operator(self, what: str, method: str) -> Callable:
        fn = _lookup(self.agent_type, method)
        if self.index is not None:
            self._check_view()
            if not getattr(fn, "_mb_views", False):
                raise NotImplementedError(f"{what}({method!r}) on a selected / reordered view: the operator registered for "
                                          f"{self.agent_type.__name__}.{method} works on whole populations only "
                                          "(register it with batch_operator(..., views=True) to take agentset.index)")
        return fn

    def do(self, method, *args, **kwargs) -> "DeviceAgentSet":
        """agentset.py:756-770: call `method` on every agent (registration order) -- one kernel launch."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.do takes the name of a registered batch operator, not a callable")
        self._operator("do", method)(self, *args, **kwargs)
        return self

    def shuffle_do(self, method, *args, **kwargs) -> "DeviceAgentSet":
        """agentset.py:772-787: activate every agent once in a random order (Philox priority order)."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.shuffle_do takes the name of a registered batch operator")
        fn = self._operator("shuffle_do", method)
        epoch = self.model._next_epoch()
        fn(self, *args, epoch=epoch, **kwargs)
        return self

    def map(self, method, *args, **kwargs):
        """agentset.py:789-804: like do, returning the operator's per-agent result tensor."""
        if not isinstance(method, str):
            raise NotImplementedError("DeviceAgentSet.map takes the name of a registered batch operator")
        return self._operator("map", method)(self, *args, **kwargs)

    # -- columns --------------------------------------------------------------
    def _col(self, name: str) -> torch.Tensor:
        self._check_view()
        try:
            col = self.columns[name]
        except KeyError:
            raise AttributeError(f"agents have no attribute {name!r}") from None
        return col if self.index is None else col[self.index]

    def get(self, attr_names, handle_missing: str = "error", default_value=None):
        """agentset.py:171-213 as bulk column reads instead of a Python loop: one name -> the column ([n, ...] device
        tensor); a list of names -> ONE [n, k] tensor, row i = the values of agent i in the order of the names (the
        reference's list of per-agent lists; scalar columns only, promoted to their common dtype).
        `handle_missing="default"` fills unknown attributes with `default_value`."""
        if handle_missing not in ("error", "default"):
            raise ValueError(f"Unknown handle_missing option: {handle_missing}, should be one of 'error' or 'default'")
        if isinstance(attr_names, str):
            if attr_names not in self.columns and handle_missing == "default":
                return torch.full((len(self),), default_value, device=self._device())
            return self._col(attr_names)
        cols = [self.get(a, handle_missing, default_value) for a in attr_names]
        if any(c.dim() != 1 for c in cols):
            raise ValueError("get() of several attributes stacks scalar columns; read vector attributes one by one")
        if not cols:
            return torch.empty((len(self), 0), device=self._device())
        common = cols[0].dtype
        for c in cols[1:]:
            common = torch.promote_types(common, c.dtype)
        return torch.stack([c.to(common) for c in cols], dim=1)

    def set(self, attr_name: str, value) -> "DeviceAgentSet":
        """agentset.py:215-227: set an attribute of every agent of this set; a name the agents do not have yet becomes a
        new column (the reference's setattr creates the attribute), zero for agents outside a view."""
        if attr_name not in self.columns:
            if not isinstance(self.columns, dict):
                raise AttributeError(f"agents have no attribute {attr_name!r}")
            v = torch.as_tensor(value)
            shape = (self.n, *v.shape[1:]) if v.dim() and v.shape[0] == len(self) else (self.n, *v.shape)
            self.columns[attr_name] = torch.zeros(shape, dtype=v.dtype, device=self._device())
        col = self.columns[attr_name]
        if self.index is None:
            col[:] = torch.as_tensor(value, device=col.device, dtype=col.dtype)
        else:
            col[self.index] = torch.as_tensor(value, device=col.device, dtype=col.dtype)
        return self

    def agg(self, attribute: str, func):
        """agentset.py:137-169: aggregate a column.  `func` names a device reduction ("sum", "min", "max",
        "count_nonzero", "mean", "len") or IS one of the functions the reference's users pass -- the builtins `min`,
        `max`, `sum`, `len` and numpy's `min / max / sum / mean` run as the same device reductions (one launch, one
        scalar read-back) instead of iterating a device tensor; any other callable gets the device column.  A list or
        tuple of functions gives a list of results."""
        if isinstance(func, (list, tuple)):
            return [self.agg(attribute, f) for f in func]
        col = self._col(attribute)
        name = func if isinstance(func, str) else _reduction_name(func)
        if name is None:
            return func(col)
        n = int(col.shape[0])
        if name == "len":
            return n
        if col.dim() != 1:
            raise ValueError(f"agg({attribute!r}, {name!r}) needs a scalar attribute")
        from . import ops

        if name == "mean":
            if n == 0:
                return float("nan")
            return ops.layer_reduce(col.contiguous(), "sum").item() / n
        if n == 0:
            if name in ("min", "max"):
                raise ValueError(f"{name}() of an empty agent set")     # like the builtins on an empty list
            return 0
        return ops.layer_reduce(col.contiguous(), name).item()

    def select(self, filter_func=None, at_most=float("inf"), inplace: bool = False, agent_type=None):
        """agentset.py:66-135 with a column predicate: `filter_func(columns) -> bool mask [n]`."""
        self._check_view()
        idx = torch.arange(self.n, device=self._device()) if self.index is None else self.index
        if agent_type is not None and not issubclass(self.agent_type, agent_type):
            idx = idx[:0]
        if filter_func is not None:
            mask = filter_func(self.columns)
            idx = idx[mask[idx]]
        if at_most != float("inf"):
            # agentset.py:100-102: a float <= 1.0 is a fraction of the ORIGINAL set (rounded down), else a count
            k = int(len(self) * at_most) if (at_most <= 1.0 and isinstance(at_most, float)) else int(at_most)
            idx = idx[:k]
        if inplace:
            self.index, self._view_gen = idx, self._pop.generation
            return self
        return self._view(idx)

    def shuffle(self, inplace: bool = False):
        """agentset.py:415-437 with the Philox activation order of a fresh epoch."""
        from . import rng

        self._check_view()
        epoch = self.model._next_epoch()
        if self.index is None and self._device().type == "cuda":
            from . import ops

            idx = ops.activation_order(self.model.philox_seed, epoch, self.n, self._device()).to(torch.int64)
        else:
            from . import ops

            base = torch.arange(self.n, device=self._device()) if self.index is None else self.index
            pri = rng.priority64(self.model.philox_seed, epoch, base)        # uint64 bit patterns in int64
            idx = base[ops.argsort_stable(pri.view(torch.int64) ^ (-2**63))]
        if inplace:
            self.index, self._view_gen = idx, self._pop.generation
            return self
        return self._view(idx)

    def sort(self, key, ascending: bool = False, inplace: bool = False):
        if not isinstance(key, str):
            raise NotImplementedError("DeviceAgentSet.sort takes a column name")
        self._check_view()
        base = torch.arange(self.n, device=self._device()) if self.index is None else self.index
        from . import ops

        order = ops.argsort_stable(self.columns[key][base], descending=not ascending)
        idx = base[order]
        if inplace:
            self.index, self._view_gen = idx, self._pop.generation
            return self
        return self._view(idx)

    def groupby(self, by, result_type: str = "agentset") -> "DeviceGroupBy":
        """agentset.py:282-320: group the agents by a column (`by` = its name) or by `by(columns) -> key tensor [n]`.
        Groups appear in the order of their first member and keep the set's order inside, like the reference's dict of
        lists.  One stable radix sort of (key bits, position) on the device (csrc/sort.cu) and one read-back of the
        group boundaries; float keys are grouped by bit pattern."""
        from . import ops

        self._check_view()
        if result_type not in ("agentset", "list"):
            raise ValueError(f"Unknown result_type {result_type!r}, expected 'agentset' or 'list'")
        if isinstance(by, str) and by not in self.columns:
            raise AttributeError(f"agents have no attribute {by!r}")
        keys = self.columns[by] if isinstance(by, str) else by(self.columns)
        if keys.dim() != 1:
            raise ValueError("groupby needs one scalar key per agent")
        rows = torch.arange(self.n, device=keys.device) if self.index is None else self.index
        keys = keys[rows] if self.index is not None else keys
        m = int(rows.numel())
        if m == 0:
            return DeviceGroupBy({})
        if keys.dtype.is_floating_point:
            wide = keys.element_size() > 4
            canon = keys.to(torch.float64 if wide else torch.float32) + 0.0   # -0.0 and 0.0 are one key, as in a dict
            bits = canon.contiguous().view(torch.int64 if wide else torch.int32)
            keys = canon
        elif keys.dtype in (torch.int32, torch.int64):
            bits = keys.contiguous()
        else:                                                              # bool and narrow integers
            bits = keys.to(torch.int32)
        order_in = torch.arange(m, dtype=torch.int32, device=keys.device)
        sorted_bits, order = ops.sort_pairs(bits, order_in)             # stable: set order survives inside a group
        order = order.to(torch.int64)
        starts = ops.nonzero_indices(sorted_bits[1:] != sorted_bits[:-1]) + 1
        starts = torch.cat([starts.new_zeros(1), starts])
        first = order[starts]                                            # position of each group's first member
        by_first = ops.argsort_stable(first)                             # groups in order of first appearance
        starts_h = starts[by_first].cpu().tolist()
        ends_h = torch.cat([starts[1:], starts.new_full((1,), m)])[by_first].cpu().tolist()
        names = keys[first[by_first]].cpu().tolist()
        groups = {}
        for name, lo, hi in zip(names, starts_h, ends_h):
            members = self._view(rows[order[lo:hi]])
            groups[name] = members if result_type == "agentset" else list(members)
        return DeviceGroupBy(groups)

    def to_list(self):
        return list(self)

    def _device(self):
        for c in self.columns.values():
            return c.device
        return torch.device("cuda")


_MISSING = object()


class DeviceGroupBy:
    """agentset.py:837-970 (GroupBy): `groups` maps each key to the DeviceAgentSet (or list) of its agents."""

    def __init__(self, groups: dict):
        self.groups = dict(groups)

    def get_group(self, name, default=_MISSING):
        try:
            return self.groups[name]
        except KeyError as err:
            if default is not _MISSING:
                return default
            raise KeyError(f"No group found with name: {name}") from err

    def map(self, method, *args, **kwargs) -> dict:
        if isinstance(method, str):
            return {k: getattr(v, method)(*args, **kwargs) for k, v in self.groups.items()}
        return {k: method(v, *args, **kwargs) for k, v in self.groups.items()}

    def do(self, method, *args, **kwargs) -> "DeviceGroupBy":
        self.map(method, *args, **kwargs)
        return self

    def count(self) -> dict:
        return {k: len(v) for k, v in self.groups.items()}

    def agg(self, attr_name: str, func) -> dict:
        """Aggregate a column per group; `func` gets the group's device column (or names a reduction), like
        DeviceAgentSet.agg."""
        out = {}
        for name, group in self.groups.items():
            if isinstance(group, DeviceAgentSet):
                out[name] = group.agg(attr_name, func)
            else:
                out[name] = func([getattr(a, attr_name) for a in group])
        return out

    def __iter__(self):
        return iter(self.groups.items())

    def __len__(self):
        return len(self.groups)

    def __repr__(self) -> str:
        sizes = {name: len(group) for name, group in self.groups.items()}
        return f"{type(self).__name__}({len(self)} groups: {sizes})"


try:  # make isinstance(x, mesa.agentset.AbstractAgentSet) hold when mesa is importable
    from mesa.agentset import AbstractAgentSet as _Abstract  # type: ignore

    _Abstract.register(DeviceAgentSet)
except Exception:  # pragma: no cover
    pass
