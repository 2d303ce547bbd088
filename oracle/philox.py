"""numpy Philox4x32-10 and the mesa_b200 keyed draw protocol -- TEST INFRASTRUCTURE.

Restates Random123's philox4x32-10 (Salmon et al., SC'11) and the stream layout that
mesa_b200/csrc/philox.cuh uses on the device, so the reference can be driven with exactly
the numbers the GPU draws (oracle/replay.py).  Product code never imports this module.

Stream layout (all words uint32):
    key     = (seed & 0xffffffff, seed >> 32)
    counter = (block, agent & 0xffffffff, ((agent >> 32) & 0xffff) | (stream << 16), epoch)
    streams : 0 PRIORITY (activation order), 1 DRAW (per-agent getrandbits sequence),
              2 INIT (synthetic initial states)
    priority64(seed, epoch, agent) = philox(block=0, PRIORITY)[0] << 32 | (agent & 0xffffffff)
    draw64(seed, epoch, agent, c)  = o = philox(block=c >> 1, DRAW); i = 2*(c & 1);
                                     o[i] << 32 | o[i+1]
    getrandbits(k) = draw64 >> (64 - k)                (CPython random.getrandbits contract)
    _randbelow(n)  : k = n.bit_length(); r = getrandbits(k) until r < n
                     (CPython 3.12 Lib/random.py:242-250)
    random()       = (draw64 >> 11) * 2**-53
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_PRIORITY = 0
STREAM_DRAW = 1
STREAM_INIT = 2


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & MASK32
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & MASK32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)


def _counter(block, agent, stream: int, epoch: int):
    agent = np.asarray(agent, dtype=np.uint64)
    block = np.broadcast_to(np.asarray(block, dtype=np.uint64), agent.shape)
    ctr = np.empty(agent.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (block & MASK32).astype(np.uint32)
    ctr[..., 1] = (agent & MASK32).astype(np.uint32)
    ctr[..., 2] = (((agent >> np.uint64(32)) & np.uint64(0xFFFF)) | np.uint64(stream << 16)).astype(np.uint32)
    ctr[..., 3] = np.uint32(int(epoch) & 0xFFFFFFFF)
    return ctr


def priority64(seed: int, epoch: int, agent):
    """Activation priority of `agent` (array ok) in shuffle number `epoch`; lower acts first."""
    agent = np.asarray(agent, dtype=np.uint64)
    o = philox4x32_10(_counter(0, agent, STREAM_PRIORITY, epoch), _key(seed))
    return (o[..., 0].astype(np.uint64) << np.uint64(32)) | (agent & MASK32)


def draw64(seed: int, epoch: int, agent, c, stream: int = STREAM_DRAW):
    """c-th 64-bit draw of `agent` in `epoch` (arrays broadcast)."""
    agent = np.asarray(agent, dtype=np.uint64)
    c = np.broadcast_to(np.asarray(c, dtype=np.uint64), agent.shape)
    o = philox4x32_10(_counter(c >> np.uint64(1), agent, stream, epoch), _key(seed))
    odd = (c & np.uint64(1)).astype(bool)
    hi = np.where(odd, o[..., 2], o[..., 0]).astype(np.uint64)
    lo = np.where(odd, o[..., 3], o[..., 1]).astype(np.uint64)
    return (hi << np.uint64(32)) | lo


def getrandbits(seed: int, epoch: int, agent: int, c: int, k: int, stream: int = STREAM_DRAW) -> int:
    return int(draw64(seed, epoch, agent, c, stream)) >> (64 - k)


def uniform01(seed: int, epoch: int, agent, c, stream: int = STREAM_DRAW):
    """53-bit double in [0,1) like random.random()."""
    return (draw64(seed, epoch, agent, c, stream) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def permutation(seed: int, epoch: int, n: int) -> np.ndarray:
    """Activation order of agents 0..n-1 in `epoch`: ascending priority64."""
    return np.argsort(priority64(seed, epoch, np.arange(n, dtype=np.uint64)), kind="stable")
