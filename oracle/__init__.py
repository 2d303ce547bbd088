"""CPU oracle for the mesa_b200 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; nothing under mesa_b200/ does.  It holds

* mesa_oracle.c  -- plain-C restatement of the reference's algorithms (ctypes, below);
* philox.py      -- numpy Philox4x32-10 + the keyed draw protocol;
* replay.py      -- ReplayRandom: drives the UNMODIFIED reference with those draws;
* make_golden.py -- runs the live reference (build container only) and writes tests/golden/.

Parity pin: tests/test_oracle_golden.py checks every function here against fixtures that
make_golden.py produced from the live reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_int64, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mesa_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def set_threads(n: int) -> None:
    lib().orc_set_threads(c_int(int(n)))


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr), _p(key), _p(out))
    return out


def gol_step(state: np.ndarray, torus: bool = True) -> np.ndarray:
    """One synchronous Game of Life generation of a uint8 [rows, cols] layer."""
    state = np.ascontiguousarray(state, dtype=np.uint8)
    out = np.empty_like(state)
    rc = lib().orc_gol_step_u8(_p(state), _p(out), c_int64(state.shape[0]), c_int64(state.shape[1]),
                               c_int(bool(torus)))
    if rc != 0:
        raise ValueError("gol_step: torus needs both dimensions >= 3")
    return out
