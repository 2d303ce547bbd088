"""DeviceDataCollector: mesa.datacollection.DataCollector served from the device-resident columns.

Mirrors the public surface of mesa/datacollection.py:60-520 -- `model_reporters`, `agent_reporters`,
`agenttype_reporters`, `tables`, `collect(model)`, `get_model_vars_dataframe()`, `get_agent_vars_dataframe()`,
`get_agenttype_vars_dataframe(type)`, `add_table_row`, `get_table_dataframe` -- with the same reporter forms for the
model (attribute name, lambda / function of the model, bound method, `[function, [params]]`).

The reference records agents with one Python call per agent and reporter (`map(get_reports, model.agents)`,
datacollection.py:314-333: ~28 % of a profiled Schelling step).  Here an agent reporter is COLUMNAR:
  * an attribute name  -> the agents' column of that name: ONE device-to-host copy per reporter and collection;
  * a callable         -> called once with the DeviceAgentSet, returns a length-n tensor / array (vectorised form of the
                          reference's `lambda a: ...`); `[function, [params]]` -> `function(agentset, *params)`.
Records are kept as arrays per collection and only turned into the reference's (Step, AgentID)-indexed DataFrame when
it is asked for.  AgentID is `unique_id` = registration index + 1 (agent.py:51-71)."""
from __future__ import annotations

import types
import warnings
from copy import deepcopy
from functools import partial

import numpy as np
import torch


def _to_host(value):
    """Model-reporter values: 0-d tensors become Python scalars, tensors numpy copies, the rest is deep-copied
    (datacollection.py:383-397 stores copies so later steps cannot change recorded values)."""
    if isinstance(value, torch.Tensor):
        return value.item() if value.dim() == 0 else value.detach().cpu().numpy().copy()
    return deepcopy(value)


class RaggedColumn:
    """A per-agent column of LISTS (e.g. the reference's `trade_partners`, sugarscape_g1mt/agents.py:201) kept as a CSR pair
    -- `offsets` [n + 1] and the flat `values` -- until somebody asks for the DataFrame: an agent reporter may return one
    instead of n Python lists per collection."""

    def __init__(self, offsets: np.ndarray, values: np.ndarray):
        self.offsets = np.asarray(offsets, dtype=np.int64)
        self.values = np.asarray(values)
        if self.offsets.ndim != 1 or len(self.offsets) < 1 or int(self.offsets[-1]) != len(self.values):
            raise ValueError("RaggedColumn: offsets must be [n + 1] with offsets[-1] == len(values)")
        self.shape = (len(self.offsets) - 1,)
        self.ndim = 1

    def __len__(self) -> int:
        return self.shape[0]

    def copy(self) -> "RaggedColumn":
        return RaggedColumn(self.offsets.copy(), self.values.copy())

    def materialize(self) -> np.ndarray:
        """The object array of per-agent lists the reference would have recorded."""
        out = np.empty(len(self), dtype=object)
        flat = self.values.tolist()
        out[:] = [flat[lo:hi] for lo, hi in zip(self.offsets[:-1].tolist(), self.offsets[1:].tolist())]
        return out


class TableMissingException(KeyError):
    """mesa/datacollection.py: raised when a table name is not part of the collector."""


class DeviceDataCollector:
    def __init__(self, model_reporters=None, agent_reporters=None, agenttype_reporters=None, tables=None):
        self.model_reporters, self.agent_reporters, self.agenttype_reporters = {}, {}, {}
        self.model_vars: dict[str, list] = {}
        self._collection_steps: list[float] = []
        self._agent_records: dict[float, tuple] = {}
        self._agenttype_records: dict[float, dict] = {}
        self.tables: dict[str, dict[str, list]] = {}
        for name, rep in (model_reporters or {}).items():
            self._new_model_reporter(name, rep)
        for name, rep in (agent_reporters or {}).items():
            self.agent_reporters[name] = self._check_agent_reporter(name, rep)
        for agent_type, reps in (agenttype_reporters or {}).items():
            self.agenttype_reporters[agent_type] = {n: self._check_agent_reporter(n, r) for n, r in reps.items()}
        for name, columns in (tables or {}).items():
            self.tables[name] = {column: [] for column in columns}

    # ------------------------------------------------------------------ reporters
    @staticmethod
    def _check_list_reporter(name, reporter):
        # datacollection.py:143-155
        if len(reporter) != 2 or not callable(reporter[0]):
            raise ValueError(f"Reporter '{name}' must use the format [function, [param1, param2]]. Got: {reporter!r}")
        if not isinstance(reporter[1], (list, tuple)):
            raise ValueError(f"Reporter '{name}' must use the format [function, [param1, param2]]. "
                             f"The second element must be a list or tuple of parameters, got: {reporter[1]!r}")

    def _new_model_reporter(self, name, reporter):
        if reporter is None:
            raise ValueError(f"Model reporter '{name}' is None")
        if isinstance(reporter, list):
            self._check_list_reporter(name, reporter)
        elif not (isinstance(reporter, str) or callable(reporter)):
            raise TypeError(f"Model reporter '{name}' must be an attribute name, a callable or [function, [params]]")
        self.model_reporters[name] = reporter
        self.model_vars[name] = []

    def _check_agent_reporter(self, name, reporter):
        if isinstance(reporter, list):
            self._check_list_reporter(name, reporter)
            func, params = reporter

            def with_params(agents, func=func, params=tuple(params)):
                return func(agents, *params)

            return with_params
        if isinstance(reporter, str) or callable(reporter):
            return reporter
        raise TypeError(f"Agent reporter '{name}' must be a column name, a vectorised callable or [function, [params]]")

    @staticmethod
    def _column(agents, name, reporter) -> np.ndarray:
        if isinstance(reporter, str):
            value = agents.get(reporter)   # raises AttributeError for an unknown column, like the reference's getattr
        else:
            value = reporter(agents)
        if not isinstance(value, RaggedColumn):
            value = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
        if value.shape[:1] != (len(agents),):
            raise ValueError(f"Agent reporter '{name}' must return one value per agent ({len(agents)}), got shape {value.shape}")
        return value.copy()

    # ------------------------------------------------------------------ collection
    def collect(self, model) -> None:
        """datacollection.py:370-408."""
        if self.model_reporters:
            self._collection_steps.append(model.time)
            for var, reporter in self.model_reporters.items():
                if isinstance(reporter, (types.LambdaType, partial)):
                    value = reporter(model)
                elif isinstance(reporter, str):
                    value = getattr(model, reporter, None)
                elif isinstance(reporter, list):
                    value = reporter[0](*reporter[1])
                else:
                    value = reporter()
                self.model_vars[var].append(_to_host(value))
        if self.agent_reporters:
            # datacollection.py:314-333 iterates model.agents = EVERY agent of every type: concatenate the registered
            # populations (registration order), one bulk column copy per population and reporter
            sets = self._populations(model)
            ids = np.concatenate([a.unique_ids_host() for a in sets]) if sets else np.zeros(0, np.int64)
            cols = [self._concat([self._column(a, n, r) for a in sets]) if sets else np.zeros(0)
                    for n, r in self.agent_reporters.items()]
            self._agent_records[model.time] = (ids, cols)
        if self.agenttype_reporters:
            self._agenttype_records[model.time] = {}
            for agent_type, reps in self.agenttype_reporters.items():
                # datacollection.py:335-365: model.agents_by_type[agent_type], else every agent that is an instance of it
                by_type = getattr(model, "agents_by_type", {}) or {}
                if agent_type in by_type:
                    sets = [by_type[agent_type]]
                else:
                    sets = [a for a in self._populations(model) if issubclass(a.agent_type, agent_type)]
                if not sets:
                    raise ValueError(f"Agent type {agent_type} is not recognized as an Agent type in the model or Agent "
                                     f"subclass. Use an Agent (sub)class, like {[a.agent_type for a in self._populations(model)]}.")
                ids = np.concatenate([a.unique_ids_host() for a in sets])
                cols = [self._concat([self._column(a, n, r) for a in sets]) for n, r in reps.items()]
                self._agenttype_records[model.time][agent_type] = (ids, cols)

    @staticmethod
    def _concat(parts: list):
        """The columns of several populations in a row; a single ragged column stays lazy."""
        if len(parts) == 1:
            return parts[0]
        return np.concatenate([p.materialize() if isinstance(p, RaggedColumn) else p for p in parts])

    @staticmethod
    def _populations(model) -> list:
        """Every registered DeviceAgentSet of the model (agents_by_type values, registration order); a model that only
        assigned `model.agents` has exactly that one."""
        sets = list((getattr(model, "agents_by_type", None) or {}).values())
        if not sets and getattr(model, "agents", None) is not None:
            sets = [model.agents]
        return sets

    # ------------------------------------------------------------------ tables
    def add_table_row(self, table_name, row, ignore_missing=False):
        if table_name not in self.tables:
            raise TableMissingException(table_name)
        for column in self.tables[table_name]:
            if column in row:
                self.tables[table_name][column].append(row[column])
            elif ignore_missing:
                self.tables[table_name][column].append(None)
            else:
                raise ValueError(f"Could not insert row with missing column '{column}'")

    # ------------------------------------------------------------------ data frames
    def get_model_vars_dataframe(self):
        import pandas as pd

        if not self.model_reporters:
            warnings.warn("No model reporters have been defined in the DataCollector, returning empty DataFrame.",
                          UserWarning, stacklevel=2)
            return pd.DataFrame()
        return pd.DataFrame(self.model_vars)

    @staticmethod
    def _frame(records, names):
        import pandas as pd

        steps, ids, cols = [], [], [[] for _ in names]
        for time, (agent_ids, values) in records:
            steps.append(np.full(len(agent_ids), time))
            ids.append(agent_ids)
            for k, v in enumerate(values):
                cols[k].append(v.materialize() if isinstance(v, RaggedColumn) else v)
        data = {"Step": np.concatenate(steps) if steps else [], "AgentID": np.concatenate(ids) if ids else []}
        for n, c in zip(names, cols):
            data[n] = list(np.concatenate(c)) if c and c[0].ndim > 1 else (np.concatenate(c) if c else [])
        return pd.DataFrame(data).set_index(["Step", "AgentID"])

    def get_agent_vars_dataframe(self):
        import pandas as pd

        if not self.agent_reporters:
            warnings.warn("No agent reporters have been defined in the DataCollector, returning empty DataFrame.",
                          UserWarning, stacklevel=2)
            return pd.DataFrame()
        return self._frame(self._agent_records.items(), list(self.agent_reporters))

    def get_agenttype_vars_dataframe(self, agent_type):
        import pandas as pd

        if agent_type not in self.agenttype_reporters:
            warnings.warn(f"No agent-type reporters have been defined for {agent_type} in the DataCollector, "
                          "returning empty DataFrame.", UserWarning, stacklevel=2)
            return pd.DataFrame()
        records = ((t, r[agent_type]) for t, r in self._agenttype_records.items() if agent_type in r)
        return self._frame(records, list(self.agenttype_reporters[agent_type]))

    def get_table_dataframe(self, table_name):
        import pandas as pd

        if table_name not in self.tables:
            raise TableMissingException(table_name)
        return pd.DataFrame(self.tables[table_name])
