// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the keyed draw protocol (K5 of
// SURVEY.md §2.4).
//
// The reference draws every number from ONE sequential MT19937 stream shared by the grid,
// the cells and the agent sets (mesa/model.py:129; call sites agentset.py:777,
// grid.py:304/310, cell_collection.py:119/141).  A sequential stream cannot be consumed in
// parallel, so the device path keys each number on (seed, epoch, agent, draw#):
//
//   key     = (seed lo32, seed hi32)
//   counter = (block, agent lo32, (agent hi16) | stream << 16, epoch)
//   stream 0 PRIORITY: activation order of shuffle number `epoch` = ascending
//            priority64 = philox(block 0)[0] << 32 | agent lo32   (replaces random.shuffle,
//            agentset.py:772-787; any permutation is a valid activation order)
//   stream 1 DRAW: the getrandbits sequence an agent consumes while it is activated:
//            draw64(c) = words (2*(c&1), 2*(c&1)+1) of philox(block c>>1);
//            getrandbits(k) = draw64 >> (64-k); _randbelow(n) rejects like CPython 3.12
//            Lib/random.py:242-250 (k = n.bit_length(); redraw while r >= n).
//   stream 2 INIT: synthetic initial states.
//
// oracle/philox.py restates the same layout in numpy and oracle/replay.py feeds it to the
// unmodified reference, so both sides consume identical numbers.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif

namespace mb {

enum : uint32_t { kStreamPriority = 0, kStreamDraw = 1, kStreamInit = 2 };

struct U4 {
    uint32_t x, y, z, w;
};

MB_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

MB_HD U4 philox_keyed(uint64_t seed, uint32_t epoch, uint64_t agent, uint32_t stream, uint32_t block) {
    return philox4x32_10(block, (uint32_t)agent, (uint32_t)((agent >> 32) & 0xffffu) | (stream << 16), epoch,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Activation priority: lower acts first.  Unique per agent (low word is the agent index).
MB_HD uint64_t priority64(uint64_t seed, uint32_t epoch, uint64_t agent) {
    const U4 o = philox_keyed(seed, epoch, agent, kStreamPriority, 0u);
    return ((uint64_t)o.x << 32) | (uint32_t)agent;
}

// The per-activation getrandbits sequence of one agent.
struct DrawStream {
    uint64_t seed;
    uint64_t agent;
    uint32_t epoch;
    uint32_t stream;
    uint32_t c;  // next draw index
    U4 cache;    // philox block (c >> 1) when (c & 1) == 1

    MB_HD DrawStream(uint64_t seed_, uint32_t epoch_, uint64_t agent_, uint32_t stream_ = kStreamDraw)
        : seed(seed_), agent(agent_), epoch(epoch_), stream(stream_), c(0u), cache{0u, 0u, 0u, 0u} {}

    MB_HD uint64_t next64() {
        uint64_t v;
        if ((c & 1u) == 0u) {
            cache = philox_keyed(seed, epoch, agent, stream, c >> 1);
            v = ((uint64_t)cache.x << 32) | cache.y;
        } else {
            v = ((uint64_t)cache.z << 32) | cache.w;
        }
        ++c;
        return v;
    }
    // Position the stream at draw index `c0` (used to resume after cached probes).
    MB_HD void seek(uint32_t c0) {
        c = c0;
        if (c & 1u) cache = philox_keyed(seed, epoch, agent, stream, c >> 1);
    }
    MB_HD uint64_t getrandbits(int k) { return next64() >> (64 - k); }
    // CPython Random._randbelow_with_getrandbits (Lib/random.py:242-250); n >= 1, n < 2^63.
    MB_HD uint64_t randbelow(uint64_t n) {
        const int k = 64 - clz64(n);  // n.bit_length()
        uint64_t r = getrandbits(k);
        while (r >= n) r = getrandbits(k);
        return r;
    }
    // random.random(): 53-bit mantissa in [0, 1).
    MB_HD double uniform01() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }

    static MB_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
        return __clzll((long long)x);
#else
        return x ? __builtin_clzll(x) : 64;
#endif
    }
};

}  // namespace mb
