"""ctypes binding of libmesa_b200.so — the C-ABI boundary (include/mesa_b200.h).

There is no CPU fallback: if the shared library is missing or a symbol cannot be
resolved the import of any compute entry point raises.  Argument errors reported by
the library (MB_ERR_INVALID) surface as ``ValueError`` like the reference's own
parameter validation (mesa/discrete_space/grid.py:280-286).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmesa_b200.so")

MB_OK = 0
MB_ERR_INVALID = -1
MB_ERR_CUDA = -2
MB_ERR_WORKSPACE = -3
MB_ERR_UNSUPPORTED = -4


class NativeLibraryMissing(ImportError):
    """libmesa_b200.so has not been built (python -m mesa_b200.build)."""


MB_MAX_PEERS = 16


class Shard(ctypes.Structure):
    """mb_shard_t: row-block partition + per-rank (peer-mapped) base pointers of the per-cell arrays."""

    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("rows_total", c_int64), ("cols", c_int64),
                ("slots", c_void_p * MB_MAX_PEERS), ("type", c_void_p * MB_MAX_PEERS),
                ("happy", c_void_p * MB_MAX_PEERS), ("agent_id", c_void_p * MB_MAX_PEERS),
                ("static_bits", c_void_p), ("bits_stride", c_int64)]


class MesaB200Error(RuntimeError):
    """A CUDA/runtime failure reported through the C ABI."""


_P = c_void_p  # device or host pointer
_S = c_void_p  # cudaStream_t

# name -> (restype, argtypes).  Kept in the same order as include/mesa_b200.h.
SIGNATURES: dict[str, tuple] = {
    "mb_version": (c_int, []),
    "mb_last_error": (c_char_p, []),
    "mb_sm_count": (c_int, []),
    # Game of Life
    "mb_gol_step_u8": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, _S]),
    "mb_gol_step_u8_fused": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, _P, _P, _P, _P, _P, _P, c_uint32, _S]),
    "mb_gol_pack_u8": (c_int, [_P, _P, c_int64, c_int64, _S]),
    "mb_gol_unpack_u8": (c_int, [_P, _P, c_int64, c_int64, _S]),
    "mb_gol_step_bits": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _S]),
    # Schelling
    "mb_schelling_happy_i8": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, _P, c_int, _P, _S]),
    "mb_reloc_table_bytes": (c_int, [c_int64, c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_schelling_relocate": (c_int, [_P, _P, _P, _P, _P, ctypes.c_size_t, c_int64, c_int64, c_uint64, c_uint32, _P, ctypes.c_size_t,
                                      c_int64, ctypes.POINTER(c_int64), ctypes.POINTER(c_int), _S]),
    "mb_reloc_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_reloc_slots_init": (c_int, [_P, _S]),
    "mb_reloc_collect": (c_int, [_P, c_uint64, c_uint32, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_counters": (c_int, [_P, ctypes.c_size_t, c_int64, ctypes.POINTER(c_uint64), _S]),
    "mb_reloc_pass": (c_int, [_P, c_int64, c_uint64, c_uint32, c_uint64, c_uint64, c_int, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_bucket_lists": (c_int, [c_int64, c_int, _P, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_pass_list": (c_int, [_P, c_int64, c_int64, c_uint64, c_uint32, c_uint64, c_uint64, c_int, _P, ctypes.c_size_t, c_int64,
                                   _S]),
    "mb_reloc_release_take": (c_int, [_P, ctypes.c_size_t, c_int64, _P, _P, c_int64, _S]),
    "mb_reloc_release_scan": (c_int, [_P, c_int64, c_uint64, c_uint32, c_uint64, c_uint64, _P, _P, c_int64, _P,
                                      ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_fb_export": (c_int, [_P, c_int64, c_int, c_int64, c_uint64, c_uint32, _P, _P, _P, _P, ctypes.c_size_t,
                                   c_int64, _S]),
    "mb_reloc_fb_count": (c_int, [_P, _P, c_int, _P, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_fb_pick": (c_int, [_P, _P, _P, c_int, _P, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_fb_commit": (c_int, [_P, _P, _P, _P, c_int, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_fb_count_into": (c_int, [_P, _P, c_int, _P, _P, _S]),
    "mb_reloc_event_loop": (c_int, [_P, ctypes.POINTER(c_void_p), _P, _P, _P, c_int, _P, _P, _P, c_uint64, c_uint32, _S]),
    "mb_reloc_apply_overrides": (c_int, [c_int64, c_uint64, c_uint64, _P, _P, c_int, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_reloc_apply": (c_int, [_P, c_int, c_int64, _P, _P, ctypes.c_size_t, c_int64, _S]),
    "mb_schelling_small_max_cells": (c_int, []),
    "mb_schelling_small_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_schelling_step_small": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, _P, c_uint64, c_uint32, c_int,
                                        c_int, c_int, _P, ctypes.c_size_t, c_int64, _P, _S]),
    # radix sort
    "mb_sort_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_sort_pairs_u32": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, ctypes.c_size_t,
                                  ctypes.POINTER(c_int), _S]),
    "mb_sort_pairs_u64": (c_int, [_P, _P, _P, _P, c_int64, c_int, c_int, _P, ctypes.c_size_t,
                                  ctypes.POINTER(c_int), _S]),
    # continuous space
    "mb_celllist_workspace_bytes": (c_int, [c_int64, c_double, c_double, c_double, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_boids_step_f32": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int,
                                  c_double, c_double, c_double, c_double, c_double, c_double, _P, _P,
                                  ctypes.c_size_t, _S]),
    "mb_activation_order_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_activation_order": (c_int, [c_uint64, c_uint32, c_int64, _P, _P, ctypes.c_size_t, _S]),
    "mb_boids_seq_workspace_bytes": (c_int, [c_int64, c_double, c_double, c_double, c_double,
                                             ctypes.POINTER(ctypes.c_size_t)]),
    "mb_boids_step_seq_f32": (c_int, [_P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int,
                                      c_double, c_double, c_double, c_double, c_double, c_double, c_double,
                                      c_uint64, c_uint32, _P, _P, ctypes.c_size_t, ctypes.POINTER(c_int), _S]),
    "mb_radius_query_f32": (c_int, [_P, c_int64, c_double, c_double, c_double, c_double, c_int, _P, _P, c_int64,
                                    c_double, _P, _P, _P, _P, _P, ctypes.c_size_t, c_int, _S]),
    "mb_knn_query_f32": (c_int, [_P, c_int64, c_double, c_double, c_double, c_double, c_int, _P, _P, c_int64, c_int,
                                 c_double, _P, _P, _P, ctypes.c_size_t, c_int, _S]),
    "mb_point_distances_f32": (c_int, [_P, _P, c_int64, c_double, c_double, c_int, c_double, c_double, _P, _P, _S]),
    "mb_point_distances_nd_f32": (c_int, [_P, c_int, _P, c_int64, ctypes.POINTER(c_double), c_int,
                                          ctypes.POINTER(c_double), _P, _P, _S]),
    "mb_boids_send_ghosts": (c_int, [_P, _P, c_int64, c_double, c_double, c_double, _P, _P, _P, _P, _P, _P, c_int64, _P, _S]),
    "mb_boids_route": (c_int, [_P, _P, _P, c_int64, c_double, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               c_int64, _P, _S]),
    "mb_boids_compact": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P, _P, _S]),
    "mb_boids_resort_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_double, c_double, c_double, c_double, c_int, c_double,
                                    _P, ctypes.c_size_t, _S]),
    "mb_scan_scratch_words": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_offsets_from_counts": (c_int, [_P, c_int64, _P, _P, _S]),
    "mb_ws_count_cells": (c_int, [_P, c_int64, _P, c_int64, _S]),
    "mb_ws_segment_starts": (c_int, [_P, c_int64, c_int64, _P, _S]),
    "mb_ws_sheep_pass": (c_int, [_P, _P, c_int64, c_int, c_int, c_uint64, c_uint32, _P, _P, _P, _P, _S]),
    "mb_ws_compare_reset": (c_int, [_P, _P, c_int64, _P, _S]),
    "mb_ws_sheep_apply": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_uint64, c_uint32, _P, _P, _P, _P, _P, c_double, c_double,
                                  ctypes.c_int32, _P, _P, _P, _P, _S]),
    "mb_ws_wolf_pass": (c_int, [_P, _P, c_int64, c_int, c_int, c_uint64, c_uint32, _P, _P, _P, _P, _S]),
    "mb_ws_wolf_apply": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_uint64, c_uint32, _P, _P, _P, c_double, c_double, _P, _P, _P,
                                 _S]),
    "mb_ws_regrow": (c_int, [_P, _P, c_int64, ctypes.c_int32, _S]),
    # Epstein civil violence
    "mb_epstein_workspace_bytes": (c_int, [c_int64, c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_epstein_build_occ": (c_int, [_P, _P, _P, c_int64, _P, c_int64, _S]),
    "mb_epstein_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int, _P, c_int, _P, c_int,
                                c_double, c_int, c_int, c_uint64, c_uint32, _P, _P, ctypes.c_size_t, _S]),
    # Sugarscape G1MT (movement phase)
    "mb_sugarscape_workspace_bytes": (c_int, [c_int64, c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_sugarscape_build_count": (c_int, [_P, c_int64, _P, c_int64, _S]),
    "mb_sugarscape_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int64, _P, c_int, _P,
                                   c_int, c_int, _P, c_int, c_int, c_uint64, c_uint32, _P, _P, ctypes.c_size_t, _S]),
    "mb_sugarscape_trade": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int64, _P, c_int, _P, c_int, _P, c_int, c_int,
                                    c_uint64, c_uint32, _P, _P, _P, c_int64, _P, _P, ctypes.c_size_t, _S]),
    # dynamic populations
    "mb_compact_plan": (c_int, [_P, c_int64, _P, _P, _S]),
    "mb_compact_rows": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _S]),
    # Boltzmann wealth
    "mb_boltzmann_workspace_bytes": (c_int, [c_int64, ctypes.POINTER(ctypes.c_size_t)]),
    "mb_boltzmann_init": (c_int, [_P, c_int64, c_int64, c_int64, _P, ctypes.c_size_t, _S]),
    "mb_boltzmann_step": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int, c_uint64, c_uint32, _P,
                                  ctypes.c_size_t, _P, _S]),
    "mb_boltzmann_status": (c_int, [_P, ctypes.c_size_t, c_int64, ctypes.POINTER(c_uint64), _S]),
    # property layers / neighbourhood sums
    "mb_layer_fill": (c_int, [_P, c_int64, c_int, c_double, _S]),
    "mb_layer_affine_clamp": (c_int, [_P, _P, c_int64, c_int, c_double, c_double, c_double, c_double, _S]),
    "mb_layer_binary": (c_int, [_P, _P, _P, c_int64, c_int, c_int, _S]),
    "mb_layer_reduce": (c_int, [_P, c_int64, c_int, c_int, _P, _P, ctypes.c_size_t, _S]),
    "mb_neighborhood_sum": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, _S]),
    "mb_layer_diffuse": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_double, c_int, _S]),
    "mb_neighborhood_sum_nd": (c_int, [_P, _P, c_int, ctypes.POINTER(c_int64), c_int, c_int, c_int, c_int, c_int, _S]),
    # deep-ghost refresh with a flag hand-shake (no barrier)
    "mb_ghost_exchange": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, _S]),
    "mb_ghost_wait_copied": (c_int, [_P, c_int, c_int, _S]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m mesa_b200.build` "
            "(mesa_b200 has no CPU fallback)"
        )
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().mb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Map a C-ABI return code to the exception the reference would raise."""
    if rc == MB_OK:
        return
    msg = last_error()
    if rc == MB_ERR_INVALID:
        raise ValueError(msg)
    if rc == MB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise MesaB200Error(f"[{rc}] {msg}")


def ptr(t) -> int | None:
    """Device (or pinned host) address of a torch tensor; None passes NULL."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    """cudaStream_t of torch's current stream on `device`."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream
