// Epstein's civil-violence model: one shuffle_do("step") over citizens AND cops in ONE launch, in the reference's
// sequential semantics (radius-`vision` neighbourhoods with is_empty filtering).
//
//   mesa/examples/advanced/epstein_civil_violence/model.py:100-117   self.agents.shuffle_do("step")
//   agents.py:14-26    update_neighbors: neighborhood = cell.get_neighborhood(radius = cop_vision) (BFS order, centre
//                      excluded), neighbors = its agents, empty_neighbors = its empty cells; move = random.choice of those
//   agents.py:66-103   Citizen.step: jailed -> count down; else cops / actives in vision -> arrest probability
//                      1 - exp(-k round(cops / actives)) -> ACTIVE iff grievance - risk_aversion p > threshold; move
//   agents.py:124-140  Cop.step: arrest a random ACTIVE citizen in vision (jail = randint(0, max_jail_term)); move
//
// Sequential semantics, in parallel, exactly.  Agent a only reads and writes cells within `vision` of its own cell, and
// nobody moves further than `vision`, so a depends on an earlier agent b only if b STARTED the step within 2 vision of a
// (b moved into or out of a's view, changed its state there, or arrested somebody there); influences from further away
// reach a only through an agent in between, which itself waits.  The step is therefore a dependency graph, not a chain:
// a WARP owns the agent of activation rank r (ascending Philox priority), warps take their ranks from a ticket counter
// (every lower rank is held by a running warp or done), and the warp acts once every lower-ranked agent that started
// within 2 vision of it has finished -- rank_at[cell] holds the rank of the agent that started the step in a cell until
// that agent has acted.  When it acts it sees exactly the state the sequential run would show it.
//
// One 4-byte word per cell carries all a neighbour scan needs: -1 = empty, else agent << 2 | cop << 1 | active, kept up to
// date by whoever changes it (the mover, the citizen that changes its mind, the cop that arrests).  The arrest probability
// takes the value ptab[round(cops / actives)], a host-made table of the reference's own expression (math.exp on the
// integer multiples of the constant: bit-identical by construction); round() is round-half-even on the float64 quotient
// like Python's; the comparison is the reference's float64 expression (no fused operations).
#include "common.cuh"
#include "philox.cuh"
#include <stdlib.h>

extern "C" int mb_activation_order_workspace_bytes(int64_t n, size_t* out);
extern "C" int mb_activation_order(uint64_t seed, uint32_t epoch, int64_t n, uint32_t* order_out, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream);

namespace {

constexpr uint32_t kNoRank = 0xffffffffu;
constexpr int kThreads = 256;
constexpr unsigned kPollSleepNs = 100;
constexpr unsigned kSpinLimit = 1u << 22;   // polls of one block of the dependency range before a warp gives up (status word, no hang)
constexpr uint8_t kActive = 1, kQuiet = 2, kArrested = 3;   // CitizenState (agents.py:7-10)

struct EpsWs {
    uint32_t* order;     // [n] agent of rank r
    uint32_t* rank_at;   // [ncells] rank of the agent that starts the step in the cell until it has acted, kNoRank if empty / done
    uint32_t* counters;  // [0] ticket, [1] status (non-zero: a wait ran into the spin limit), [2..4] active / quiet / arrested
    void* order_ws;
    size_t order_ws_bytes;
};

size_t carve(EpsWs* ws, char* base, int64_t n, int64_t ncells) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    EpsWs t;
    t.order = (uint32_t*)take(4 * n);
    t.rank_at = (uint32_t*)take(4 * ncells);
    t.counters = (uint32_t*)take(4 * 16);
    mb_activation_order_workspace_bytes(n, &t.order_ws_bytes);
    t.order_ws = take(t.order_ws_bytes);
    if (ws) *ws = t;
    return off;
}

struct EpsArgs {
    int32_t* occ;                 // [ncells] -1 | agent << 2 | cop << 1 | active
    int32_t* cell;                // [n]
    const uint8_t* kind;          // [n] 0 citizen, 1 cop
    uint8_t* state;               // [n] CitizenState value (cops: 0)
    int32_t* jail;                // [n]
    const double* risk_aversion;  // [n]
    const double* grievance;      // [n]
    double* arrest_probability;   // [n]
    const short2* voff;           // [nv] neighbourhood offsets (dx, dy) in the reference's BFS order
    const short2* woff;           // [nw] offsets of the 2 vision ball (any order)
    const double* ptab;           // [nv + 1]
    int64_t n;
    int width, height, nv, nw, max_jail_term, movement;
    double threshold;
    uint64_t seed;
    uint32_t epoch;
};

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// (__threadfence() is the sequentially consistent fence; acquire / release is all the hand-over needs)
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__global__ void eps_prepare(EpsArgs a, EpsWs ws, int sequential) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    if (sequential) ws.order[r] = (uint32_t)r;          // agents.do("step"): registration order
    ws.rank_at[a.cell[ws.order[r]]] = (uint32_t)r;
}

__device__ __forceinline__ int wrap_cell(int x, int y, short2 d, int W, int H) {
    int nx = x + d.x, ny = y + d.y;
    nx += nx < 0 ? W : 0; nx -= nx >= W ? W : 0;
    ny += ny < 0 ? H : 0; ny -= ny >= H ? H : 0;
    return nx * H + ny;
}

// ---- one agent = one WARP ----------------------------------------------------------------------------------------------
// (The first version gave every agent a thread: its 421-cell dependency scan and its two 112-cell neighbourhood scans were
// chains of dependent L2 round trips, ~48 us per agent on the critical path -- 59 ms for the 1231 agents of the 40 x 40
// default grid, slower than the reference.)  Here the 32 lanes share the scans: every lane polls 8 cells of the dependency
// range per round trip, the neighbourhood (<= 256 cells: vision 7 has 112 / 224) is ONE round trip with the words kept in
// registers, the counts are ballots and the k-th empty cell / active citizen is found with ballots + fns in BFS order.
constexpr int kPer = 8;                  // cells per lane and block
constexpr int kBlock = 32 * kPer;
// kWaitPer (template): cells of the dependency range per lane and poll -- 16 when the range has more than 256 cells (vision 7,
// von Neumann: 421 = one poll; measured 3.3 -> 2.8 ms on 1024^2), 8 otherwise (fewer registers, no padding loads)

template <bool kEmpty>
__device__ __forceinline__ bool hit_word(int32_t w) {
    return kEmpty ? w < 0 : (w >= 0 && (w & 3) == 1);
}

// the k-th cell (BFS order: block, q, lane) whose word satisfies the predicate; block 0 comes from the registers
template <bool kEmpty>
__device__ __forceinline__ int select_kth(const EpsArgs& a, int x, int y, int k, int lane, const int32_t (&w0)[kPer],
                                          const int (&c0)[kPer], int32_t* word_out) {
    for (int blk = 0; blk * kBlock < a.nv; ++blk) {
        int32_t w[kPer];
        int c[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if (blk == 0) { w[q] = w0[q]; c[q] = c0[q]; continue; }
            const int i = blk * kBlock + q * 32 + lane;
            c[q] = i < a.nv ? wrap_cell(x, y, __ldg(a.voff + i), a.width, a.height) : -1;
            w[q] = c[q] >= 0 ? __ldcg(a.occ + c[q]) : 0;         // padding word 0 satisfies no predicate
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const unsigned b = __ballot_sync(0xffffffffu, c[q] >= 0 && hit_word<kEmpty>(w[q]));
            const int nhit = __popc(b);
            if (k < nhit) {
                const int src = (int)__fns(b, 0u, (unsigned)k + 1u);
                *word_out = __shfl_sync(0xffffffffu, w[q], src);
                return __shfl_sync(0xffffffffu, c[q], src);
            }
            k -= nhit;
        }
    }
    return -1;
}

__device__ void act_warp(const EpsArgs& a, uint32_t agent, int p, int x, int y, int lane) {
    const bool cop = __ldg(a.kind + agent) != 0;
    const int j = cop ? 0 : __ldcg(a.jail + agent);     // (issued together with the neighbourhood loads: one round trip)
    // the neighbourhood: empty cells, cops, active citizens
    int nempty = 0, ncops = 0, nactive = 0;
    int32_t w0[kPer];
    int c0[kPer];
    for (int blk = 0; blk * kBlock < a.nv; ++blk) {
        int32_t w[kPer];
        int c[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int i = blk * kBlock + q * 32 + lane;
            c[q] = i < a.nv ? wrap_cell(x, y, __ldg(a.voff + i), a.width, a.height) : -1;
            w[q] = c[q] >= 0 ? __ldcg(a.occ + c[q]) : 0;
        }
        if (blk == 0 && j != 0) {                       // agents.py:70-72: in jail, nothing else happens
            if (lane == 0) a.jail[agent] = j - 1;
            return;
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const bool v = c[q] >= 0;
            nempty += __popc(__ballot_sync(0xffffffffu, v && w[q] < 0));
            ncops += __popc(__ballot_sync(0xffffffffu, v && w[q] >= 0 && (w[q] & 2)));
            nactive += __popc(__ballot_sync(0xffffffffu, v && w[q] >= 0 && (w[q] & 3) == 1));
            if (blk == 0) { w0[q] = w[q]; c0[q] = c[q]; }
        }
    }
    mb::DrawStream ds(a.seed, a.epoch, (uint64_t)agent);
    int32_t me = (int32_t)(agent << 2) | (cop ? 2 : 0);
    if (!cop) {
        // agents.py:86-103 (the citizen counts herself among the actives)
        const double q = (double)ncops / (double)(nactive + 1);
        const int m = __double2int_rn(q);               // round-half-even, like Python's round()
        const double prob = __ldg(a.ptab + m);
        const double net_risk = __dmul_rn(__ldg(a.risk_aversion + agent), prob);
        const bool active = __dsub_rn(__ldg(a.grievance + agent), net_risk) > a.threshold;
        me |= active ? 1 : 0;
        if (lane == 0) {
            a.arrest_probability[agent] = prob;
            a.state[agent] = active ? kActive : kQuiet;
            a.occ[p] = me;                              // (rewritten below if she moves)
        }
    } else if (nactive > 0) {
        // agents.py:130-138
        const int pick = (int)ds.randbelow((uint64_t)nactive);
        int32_t w = 0;
        const int c = select_kth<false>(a, x, y, pick, lane, w0, c0, &w);
        const int32_t sentence = (int32_t)ds.randbelow((uint64_t)a.max_jail_term + 1ull);   // randint(0, max_jail_term)
        if (lane == 0) {
            const uint32_t arrestee = (uint32_t)w >> 2;
            a.jail[arrestee] = sentence;
            a.state[arrestee] = kArrested;
            a.occ[c] = w & ~1;                          // no longer active
        }
    }
    if (a.movement && nempty > 0) {                     // agents.py:22-26
        const int pick = (int)ds.randbelow((uint64_t)nempty);
        int32_t w = 0;
        const int dest = select_kth<true>(a, x, y, pick, lane, w0, c0, &w);
        if (lane == 0) {
            a.occ[p] = -1;
            a.occ[dest] = me;
            a.cell[agent] = dest;
        }
    }
}

template <int kWaitPer>
__global__ void __launch_bounds__(kThreads)
eps_chain(EpsArgs a, EpsWs ws, int batch) {
    const int lane = threadIdx.x & 31;
    while (true) {
        unsigned first = 0;
        if (lane == 0) first = atomicAdd(&ws.counters[0], (unsigned)batch);   // ranks are handed out in order: every lower rank
        first = __shfl_sync(0xffffffffu, first, 0);                           // is held by a running warp or done
        if ((int64_t)first >= a.n) break;
        const int64_t end = (int64_t)first + batch < a.n ? (int64_t)first + batch : a.n;
        for (int64_t r = first; r < end; ++r) {
            const uint32_t agent = ws.order[r];
            const int p = __ldcg(a.cell + agent);       // its own cell: nobody else moves it
            const int x = p / a.height, y = p - x * a.height;
            // every agent of lower rank that STARTED within 2 vision must have finished: pending[cell] = the rank of the agent
            // that started the step there until it has acted, kNoRank afterwards (and for cells that were empty)
            bool bad = false;
            for (int i0 = 0; i0 < a.nw && !bad; i0 += 32 * kWaitPer) {
                int wc[kWaitPer];
                unsigned pend = 0u;
#pragma unroll
                for (int q = 0; q < kWaitPer; ++q) {
                    const int i = i0 + q * 32 + lane;
                    wc[q] = 0;
                    if (i < a.nw) {
                        wc[q] = wrap_cell(x, y, __ldg(a.woff + i), a.width, a.height);
                        pend |= 1u << q;
                    }
                }
                unsigned spins = 0u;
                while (true) {
                    uint32_t v[kWaitPer];
#pragma unroll
                    for (int q = 0; q < kWaitPer; ++q) v[q] = (pend >> q) & 1u ? ld_relaxed(ws.rank_at + wc[q]) : kNoRank;
#pragma unroll
                    for (int q = 0; q < kWaitPer; ++q)
                        if (v[q] >= (uint32_t)r) pend &= ~(1u << q);
                    if (!__any_sync(0xffffffffu, pend != 0u)) break;
                    if (++spins > kSpinLimit) { bad = true; break; }
                }
            }
            fence_acq_rel();                            // their writes, before this warp's reads
            __syncwarp();
            if (!bad) act_warp(a, agent, p, x, y, lane);
            __syncwarp();
            if (lane == 0) {
                if (bad) ws.counters[1] = 1u;           // gave up WITHOUT acting: the status word reports it
                fence_acq_rel();                        // my writes, before my flag
                st_relaxed(ws.rank_at + p, kNoRank);
            }
        }
    }
}

// ======================================================================================================================
// The single-CTA variant for small grids: its own copy of the act code (shared-memory accessors, work moved before the
// dependency wait) -- the multi-CTA kernel above is throughput-bound at the sizes it serves and keeps its registers low.
// kShared: the occupancy words and the jail sentences live in SHARED memory (the single-CTA kernel for small grids below);
// the pointers are generic, the accesses volatile instead of L2-coherent loads
template <bool kShared>
__device__ __forceinline__ int32_t ld_live(const int32_t* p) {
    return kShared ? *reinterpret_cast<const volatile int32_t*>(p) : __ldcg(p);
}
template <bool kShared>
__device__ __forceinline__ void st_live(int32_t* p, int32_t v) {
    if (kShared) *reinterpret_cast<volatile int32_t*>(p) = v;
    else *p = v;
}

template <bool kEmpty, bool kShared>
__device__ __forceinline__ int select_kth_s(const EpsArgs& a, int x, int y, int k, int lane, const int32_t (&w0)[kPer],
                                          const int (&c0)[kPer], int32_t* word_out) {
    for (int blk = 0; blk * kBlock < a.nv; ++blk) {
        int32_t w[kPer];
        int c[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if (blk == 0) { w[q] = w0[q]; c[q] = c0[q]; continue; }
            const int i = blk * kBlock + q * 32 + lane;
            c[q] = i < a.nv ? wrap_cell(x, y, kShared ? a.voff[i] : __ldg(a.voff + i), a.width, a.height) : -1;
            w[q] = c[q] >= 0 ? ld_live<kShared>(a.occ + c[q]) : 0;   // padding word 0 satisfies no predicate
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const unsigned b = __ballot_sync(0xffffffffu, c[q] >= 0 && hit_word<kEmpty>(w[q]));
            const int nhit = __popc(b);
            if (k < nhit) {
                const int src = (int)__fns(b, 0u, (unsigned)k + 1u);
                *word_out = __shfl_sync(0xffffffffu, w[q], src);
                return __shfl_sync(0xffffffffu, c[q], src);
            }
            k -= nhit;
        }
    }
    return -1;
}

// What does not depend on the other agents is computed BEFORE the dependency wait (the agent's constants, the cells of its
// neighbourhood, the first draws of its stream): the ~900 dependency levels of a step are a chain of act_warp calls, and
// every instruction between "the last dependency is done" and "my flag is set" is paid once per level.
struct PreAct {
    int c0[kPer];          // cells of the first neighbourhood block (this lane's)
    uint64_t draw[4];      // the first four 64-bit draws of the agent's stream
    bool cop;
    double risk_aversion, grievance;
};

template <bool kShared>
__device__ __forceinline__ void pre_act(const EpsArgs& a, uint32_t agent, int x, int y, int lane, PreAct& pre) {
    pre.cop = __ldg(a.kind + agent) != 0;
    pre.risk_aversion = __ldg(a.risk_aversion + agent);
    pre.grievance = __ldg(a.grievance + agent);
    if (kShared) {
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int i = q * 32 + lane;
            pre.c0[q] = i < a.nv ? wrap_cell(x, y, a.voff[i], a.width, a.height) : -1;
        }       // (the multi-CTA kernel is throughput-bound at the sizes it serves: four draws per agent up front, most
                         // of them never used, cost it 25 %; the single-CTA kernel is latency-bound and gains 15 %)
        mb::DrawStream ds(a.seed, a.epoch, (uint64_t)agent);
#pragma unroll
        for (int k = 0; k < 4; ++k) pre.draw[k] = ds.next64();
    }
}

// _randbelow on the prefetched draws, going on with the stream when they run out (rejections: < 1 in 10^3 activations)
struct PreDraws {
    const uint64_t* pre;
    mb::DrawStream ds;
    int used;
    __device__ PreDraws(const EpsArgs& a, uint32_t agent, const uint64_t* p, bool prefetched)
        : pre(p), ds(a.seed, a.epoch, (uint64_t)agent), used(prefetched ? 0 : 5) {}
    __device__ __forceinline__ uint64_t next64() {
        if (used < 4) return pre[used++];
        if (used == 4) { ds.seek(4u); ++used; }
        return ds.next64();
    }
    __device__ __forceinline__ uint64_t randbelow(uint64_t n) {
        const int k = 64 - __clzll((long long)n);
        uint64_t r = next64() >> (64 - k);
        while (r >= n) r = next64() >> (64 - k);
        return r;
    }
};

template <bool kShared>
__device__ __forceinline__ void act_warp_s(const EpsArgs& a, uint32_t agent, int p, int x, int y, int lane, const PreAct& pre) {
    const bool cop = pre.cop;
    const int j = cop ? 0 : ld_live<kShared>(a.jail + agent);   // (issued together with the neighbourhood loads: one round trip)
    // the neighbourhood: empty cells, cops, active citizens -- counted per lane (three 10-bit fields) and summed by shuffles
    unsigned packed = 0u;
    int32_t w0[kPer];
    int c0[kPer];
    for (int blk = 0; blk * kBlock < a.nv; ++blk) {
        int32_t w[kPer];
        int c[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if (kShared && blk == 0) {
                c[q] = pre.c0[q];
            } else {
                const int i = blk * kBlock + q * 32 + lane;
                c[q] = i < a.nv ? wrap_cell(x, y, kShared ? a.voff[i] : __ldg(a.voff + i), a.width, a.height) : -1;
            }
            w[q] = c[q] >= 0 ? ld_live<kShared>(a.occ + c[q]) : 0;
        }
        if (blk == 0 && j != 0) {                       // agents.py:70-72: in jail, nothing else happens
            if (lane == 0) st_live<kShared>(a.jail + agent, j - 1);
            return;
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const bool v = c[q] >= 0;
            packed += (v && w[q] < 0 ? 1u : 0u) + (v && w[q] >= 0 && (w[q] & 2) ? 1u << 10 : 0u) +
                      (v && w[q] >= 0 && (w[q] & 3) == 1 ? 1u << 20 : 0u);
            if (blk == 0) { w0[q] = w[q]; c0[q] = c[q]; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);   // (<= 1024 cells per field)
    const int nempty = (int)(packed & 0x3ffu), ncops = (int)((packed >> 10) & 0x3ffu), nactive = (int)(packed >> 20);
    PreDraws ds(a, agent, pre.draw, kShared);
    int32_t me = (int32_t)(agent << 2) | (cop ? 2 : 0);
    if (!cop) {
        // agents.py:86-103 (the citizen counts herself among the actives)
        const double q = (double)ncops / (double)(nactive + 1);
        const int m = __double2int_rn(q);               // round-half-even, like Python's round()
        const double prob = kShared ? a.ptab[m] : __ldg(a.ptab + m);
        const double net_risk = __dmul_rn(pre.risk_aversion, prob);
        const bool active = __dsub_rn(pre.grievance, net_risk) > a.threshold;
        me |= active ? 1 : 0;
        if (lane == 0) {
            a.arrest_probability[agent] = prob;
            a.state[agent] = active ? kActive : kQuiet;
            st_live<kShared>(a.occ + p, me);            // (rewritten below if she moves)
        }
    } else if (nactive > 0) {
        // agents.py:130-138
        const int pick = (int)ds.randbelow((uint64_t)nactive);
        int32_t w = 0;
        const int c = select_kth_s<false, kShared>(a, x, y, pick, lane, w0, c0, &w);
        const int32_t sentence = (int32_t)ds.randbelow((uint64_t)a.max_jail_term + 1ull);   // randint(0, max_jail_term)
        if (lane == 0) {
            const uint32_t arrestee = (uint32_t)w >> 2;
            st_live<kShared>(a.jail + arrestee, sentence);
            a.state[arrestee] = kArrested;
            st_live<kShared>(a.occ + c, w & ~1);        // no longer active
        }
    }
    if (a.movement && nempty > 0) {                     // agents.py:22-26
        const int pick = (int)ds.randbelow((uint64_t)nempty);
        int32_t w = 0;
        const int dest = select_kth_s<true, kShared>(a, x, y, pick, lane, w0, c0, &w);
        if (lane == 0) {
            st_live<kShared>(a.occ + p, -1);
            st_live<kShared>(a.occ + dest, me);
            a.cell[agent] = dest;
        }
    }
}

// ---- small grids: ONE CTA, the live words in shared memory -----------------------------------------------------------------
// A 40 x 40 grid (the reference's default) has ~700 dependency levels whatever the machine; with the occupancy words, the
// pending ranks and the jail sentences in shared memory a level costs shared-memory latencies instead of L2 round trips and
// device-scope fences (1.5 ms -> see DESIGN.md section 4).  32 warps work on 32 consecutive ranks: more would only wait.
constexpr int kSmallThreads = 1024;
constexpr int kSmallWords = 28000;        // 2 ncells + n words of dynamic shared memory at most (a 100 x 100 grid; beyond, the
                                          // 32 warps of one CTA are fewer than the agents that could act at once)
constexpr int kSmallOffsets = 1024, kSmallWait = 4096;   // static tables: neighbourhood, dependency range

template <int kWaitPer>
__global__ void __launch_bounds__(kSmallThreads, 1)
eps_chain_small(EpsArgs a, EpsWs ws) {
    extern __shared__ int32_t sm_words[];
    __shared__ unsigned s_ticket, s_status;
    __shared__ short2 s_voff[kSmallOffsets], s_woff[kSmallWait];
    __shared__ double s_ptab[kSmallOffsets + 1];
    const int ncells = a.width * a.height;
    int32_t* occ_s = sm_words;
    uint32_t* pend_s = reinterpret_cast<uint32_t*>(sm_words + ncells);
    int32_t* jail_s = sm_words + 2 * ncells;
    for (int c = threadIdx.x; c < ncells; c += kSmallThreads) { occ_s[c] = a.occ[c]; pend_s[c] = ws.rank_at[c]; }
    for (int64_t i = threadIdx.x; i < a.n; i += kSmallThreads) jail_s[i] = a.jail[i];
    for (int i = threadIdx.x; i < a.nv; i += kSmallThreads) s_voff[i] = a.voff[i];
    for (int i = threadIdx.x; i <= a.nv; i += kSmallThreads) s_ptab[i] = a.ptab[i];
    for (int i = threadIdx.x; i < a.nw; i += kSmallThreads) s_woff[i] = a.woff[i];
    if (threadIdx.x == 0) { s_ticket = 0u; s_status = 0u; }
    __syncthreads();
    EpsArgs b = a;
    b.occ = occ_s;
    b.jail = jail_s;
    b.voff = s_voff;      // (generic pointers: act_warp reads the tables with plain loads when kShared)
    b.ptab = s_ptab;
    const int lane = threadIdx.x & 31;
    while (true) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&s_ticket, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if ((int64_t)r >= a.n) break;
        const uint32_t agent = ws.order[r];
        const int p = a.cell[agent];                    // its own cell: nobody else moves it
        const int x = p / a.height, y = p - x * a.height;
        PreAct pre;
        pre_act<true>(b, agent, x, y, lane, pre);
        bool bad = false;
        for (int i0 = 0; i0 < a.nw && !bad; i0 += 32 * kWaitPer) {
            int wc[kWaitPer];
            unsigned pend = 0u;
#pragma unroll
            for (int q = 0; q < kWaitPer; ++q) {
                const int i = i0 + q * 32 + lane;
                wc[q] = 0;
                if (i < a.nw) {
                    wc[q] = wrap_cell(x, y, s_woff[i], a.width, a.height);
                    pend |= 1u << q;
                }
            }
            unsigned spins = 0u;
            while (true) {
#pragma unroll
                for (int q = 0; q < kWaitPer; ++q)
                    if (((pend >> q) & 1u) && *reinterpret_cast<volatile uint32_t*>(pend_s + wc[q]) >= r) pend &= ~(1u << q);
                if (!__any_sync(0xffffffffu, pend != 0u)) break;
                if (++spins > kSpinLimit) { bad = true; break; }
                __nanosleep(kPollSleepNs);
            }
        }
        __threadfence_block();                          // their writes, before this warp's reads
        __syncwarp();
        if (!bad) act_warp_s<true>(b, agent, p, x, y, lane, pre);
        __syncwarp();
        if (lane == 0) {
            if (bad) s_status = 1u;
            __threadfence_block();                      // my writes, before my flag
            *reinterpret_cast<volatile uint32_t*>(pend_s + p) = kNoRank;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncells; c += kSmallThreads) a.occ[c] = occ_s[c];
    for (int64_t i = threadIdx.x; i < a.n; i += kSmallThreads) a.jail[i] = jail_s[i];
    if (threadIdx.x == 0 && s_status) ws.counters[1] = s_status;
}

__global__ void eps_counts(const uint8_t* __restrict__ kind, const uint8_t* __restrict__ state, int64_t n, uint32_t* counters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = (i < n && kind[i] == 0) ? state[i] : 0;
    const unsigned na = __popc(__ballot_sync(0xffffffffu, s == kActive));
    const unsigned nq = __popc(__ballot_sync(0xffffffffu, s == kQuiet));
    const unsigned nj = __popc(__ballot_sync(0xffffffffu, s == kArrested));
    if ((threadIdx.x & 31) == 0) {
        if (na) atomicAdd(&counters[2], na);
        if (nq) atomicAdd(&counters[3], nq);
        if (nj) atomicAdd(&counters[4], nj);
    }
}

__global__ void eps_build_occ(const int32_t* __restrict__ cell, const uint8_t* __restrict__ kind, const uint8_t* __restrict__ state,
                              int64_t n, int32_t* __restrict__ occ) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occ[cell[i]] = (int32_t)((uint32_t)i << 2) | (kind[i] ? 2 : 0) | (kind[i] == 0 && state[i] == kActive ? 1 : 0);
}

}  // namespace

MB_API int mb_epstein_workspace_bytes(int64_t n, int64_t ncells, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0 && ncells >= 0, "mb_epstein_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, n, ncells);
    return MB_OK;
}

// occ[cell] = -1 | agent << 2 | cop << 1 | active from the agent table (after placing the agents / restoring a model)
MB_API int mb_epstein_build_occ(const int32_t* cell, const uint8_t* kind, const uint8_t* state, int64_t n, int32_t* occ,
                                int64_t ncells, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 29) && ncells >= 0 && ncells < (1ll << 31), "mb_epstein_build_occ: size out of range");
    MB_REQUIRE((n == 0 || (cell && kind && state)) && (ncells == 0 || occ), "mb_epstein_build_occ: null buffer");
    int rc = mb::check_cuda(cudaMemsetAsync(occ, 0xff, 4 * (size_t)ncells, stream), "memset");
    if (rc) return rc;
    if (n) eps_build_occ<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(cell, kind, state, n, occ);
    MB_LAUNCH_CHECK("mb_epstein_build_occ");
    return MB_OK;
}

// One activation of all agents (shuffle number `epoch`; sequential != 0: registration order, `epoch` only keys the draws).  Enqueues on `stream` and returns; counts_dev (optional, device,
// 4 x uint32): status (non-zero = a wait hit the spin limit, the step is NOT valid), active, quiet, arrested citizens.
MB_API int mb_epstein_step(int32_t* occ, int32_t* cell, const uint8_t* kind, uint8_t* state, int32_t* jail,
                           const double* risk_aversion, const double* grievance, double* arrest_probability, int64_t n,
                           int64_t width, int64_t height, const int16_t* voff, int nv, const int16_t* woff, int nw,
                           const double* ptab, int max_jail_term, double threshold, int movement, int sequential,
                           uint64_t seed, uint32_t epoch, uint32_t* counts_dev, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 29), "mb_epstein_step: n out of range");
    MB_REQUIRE(width > 0 && height > 0 && width * height < (1ll << 31), "mb_epstein_step: grid out of range");
    MB_REQUIRE(nv > 0 && nw > 0 && max_jail_term >= 0, "mb_epstein_step: bad neighbourhood / jail term");
    if (n == 0) return MB_OK;
    MB_REQUIRE(occ && cell && kind && state && jail && risk_aversion && grievance && arrest_probability && voff && woff && ptab && workspace,
               "mb_epstein_step: null buffer");
    EpsWs ws;
    if (carve(&ws, (char*)workspace, n, width * height) > workspace_bytes) {
        mb::set_error("mb_epstein_step: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    EpsArgs a;
    a.occ = occ; a.cell = cell; a.kind = kind; a.state = state; a.jail = jail; a.risk_aversion = risk_aversion;
    a.grievance = grievance; a.arrest_probability = arrest_probability;
    a.voff = reinterpret_cast<const short2*>(voff); a.woff = reinterpret_cast<const short2*>(woff); a.ptab = ptab;
    a.n = n; a.width = (int)width; a.height = (int)height; a.nv = nv; a.nw = nw; a.max_jail_term = max_jail_term;
    a.movement = movement; a.threshold = threshold; a.seed = seed; a.epoch = epoch;
    int rc = sequential ? MB_OK : mb_activation_order(seed, epoch, n, ws.order, ws.order_ws, ws.order_ws_bytes, stream);
    if (rc) return rc;
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.rank_at, 0xff, 4 * (size_t)(width * height), stream), "memset"))) return rc;
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 4 * 16, stream), "memset"))) return rc;
    eps_prepare<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(a, ws, sequential);
    // one warp per agent, one ticket per agent.  (Handing a warp 8 consecutive ranks per ticket saves atomics on one
    // address but chains the 8 agents of a batch behind each other ON TOP of their real dependencies: 90 ms instead of a
    // few for 811 k agents on 1024 x 1024.  MB_EPSTEIN_BATCH=n keeps the experiment reachable.)
    const int warps_per_cta = kThreads / 32;
    int64_t ctas = mb::ceil_div(n, warps_per_cta);
    const int64_t cap = (int64_t)mb::sm_count() * 8;
    if (ctas > cap) ctas = cap;
    static int batch = 0;
    if (batch == 0) {
        const char* e = getenv("MB_EPSTEIN_BATCH");
        batch = e ? atoi(e) : 1;
        if (batch < 1 || batch > 64) batch = 1;
    }
    static int small_ok = -1;   // MB_EPSTEIN_SMALL=0 keeps every grid on the multi-CTA kernel (cross-checks)
    if (small_ok < 0) {
        const char* e = getenv("MB_EPSTEIN_SMALL");
        small_ok = e ? atoi(e) : 1;
    }
    const int64_t words = 2 * width * height + n;
    if (small_ok && words <= kSmallWords && nv <= kSmallOffsets && nw <= kSmallWait) {
        const int smem = (int)(words * 4);
        static bool attr_set = false;
        if (!attr_set) {
            if ((rc = mb::check_cuda(cudaFuncSetAttribute(eps_chain_small<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallWords * 4), "cudaFuncSetAttribute"))) return rc;
            if ((rc = mb::check_cuda(cudaFuncSetAttribute(eps_chain_small<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallWords * 4), "cudaFuncSetAttribute"))) return rc;
            attr_set = true;
        }
        if (nw > 256) eps_chain_small<16><<<1, kSmallThreads, smem, stream>>>(a, ws);
        else eps_chain_small<8><<<1, kSmallThreads, smem, stream>>>(a, ws);
    } else if (nw > 256) {
        eps_chain<16><<<(unsigned)ctas, kThreads, 0, stream>>>(a, ws, batch);
    } else {
        eps_chain<8><<<(unsigned)ctas, kThreads, 0, stream>>>(a, ws, batch);
    }
    if (counts_dev != nullptr) {
        eps_counts<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(kind, state, n, ws.counters);
        if ((rc = mb::check_cuda(cudaMemcpyAsync(counts_dev, ws.counters + 1, 4 * 4, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    }
    MB_LAUNCH_CHECK("mb_epstein_step");
    return MB_OK;
}
