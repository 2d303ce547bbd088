"""Host-side view of the keyed Philox protocol (csrc/philox.cuh) for the Python shim.

Only what the host logic needs: priorities of a few agents for `AgentSet.shuffle` views.  The
bulk numbers are generated inside the kernels; this module is NOT used on any hot path.
"""
from __future__ import annotations

import torch

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def _philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """int64 tensors holding uint32 values; Random123 philox4x32-10."""
    for _ in range(10):
        # 32x32 -> 64-bit products without overflowing int64: split the counter word into 16-bit halves
        def mulhilo(m, x):
            xl, xh = x & 0xFFFF, x >> 16
            lo = (m * xl) & 0xFFFFFFFFFFFF  # < 2^48
            mid = m * xh                    # < 2^48
            full_lo = (lo + ((mid & 0xFFFF) << 16))
            hi = (mid >> 16) + (full_lo >> 32)
            return hi & _MASK, full_lo & _MASK

        hi0, lo0 = mulhilo(_M0, c0)
        hi1, lo1 = mulhilo(_M1, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def priority64(seed: int, epoch: int, agent: torch.Tensor) -> torch.Tensor:
    """priority64(seed, epoch, agent) for a tensor of agent indices; returned as the uint64 bit pattern in int64...
    compare with `x ^ (-2**63)` for unsigned order."""
    a = agent.to(torch.int64)
    seed &= 0xFFFFFFFFFFFFFFFF
    zeros = torch.zeros_like(a)
    c0, _, _, _ = _philox4x32_10(zeros, a & _MASK, (a >> 32) & 0xFFFF, zeros + (epoch & _MASK), seed & _MASK, seed >> 32)
    return (c0 << 32) | (a & _MASK)
