"""mesa_b200 -- B200-native (sm_100a) execution backend for Mesa's data-parallel hot path.

Neighbourhood queries and per-step agent updates over Moore / von-Neumann grids and their
property layers, continuous-space radius queries, and the AgentSet.do / shuffle_do batch
step that drives them.  State lives in structure-of-arrays device
buffers owned by torch tensors; a thin ctypes layer (``_native``) calls hand-written CUDA
kernels through the C ABI in include/mesa_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
