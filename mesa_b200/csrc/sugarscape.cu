// Sugarscape G1MT: regrowth + the traders' shuffle_do("step") (move / eat / die) in ONE launch, and their
// shuffle_do("trade_with_neighbors") in another, both in the reference's sequential semantics.
//
//   mesa/examples/advanced/sugarscape_g1mt/model.py:118-127   sugar / spice regrow by 1 up to the landscape's maximum,
//                                                             then self.agents.shuffle_do("step")
//   agents.py:209-262   move: the EMPTY cells of get_neighborhood(vision, include_center=True) (von Neumann, clamped, BFS
//                       order), Cobb-Douglas welfare of (own + cell) sugar and spice, the cells math.isclose to the best
//                       welfare, of those the ones within 1 % of the smallest Euclidean distance, random.choice of these
//   agents.py:264-288   eat: take the cell's sugar and spice, pay the metabolism; die: sugar <= 0 or spice <= 0 -> remove()
//   agents.py:84-207, 290-301   trade_with_neighbors (second half of this file)
//
// The same rank-ordered dependency graph as csrc/epstein.cu: a trader reads (emptiness, sugar, spice) within its own vision
// and changes its old and new cell, so it depends on an earlier trader only if that one STARTED within vision_a + vision_b
// <= 2 max-vision of it.  A warp owns the trader of the next activation rank (ticket counter), polls the cells of that range
// until no lower rank is pending there, then acts on the live layers.  Cells are not capacity-limited (the initial
// placement may stack traders, model.py:84): a cell keeps up to 4 pending ranks (one 16-byte word; more: status 2).
// Keys: priority and draws are keyed on unique_id - 1 (traders die, rows are not ids).
//
// Float64: sugar ** (ms / mt) * spice ** (mp / mt) decides trades through STRICT comparisons, and exact mathematical ties do
// occur (28 x 54 = 27 x 56 with equal metabolisms: C's pow happens to give equal products there, a pow that is off by one
// ulp does not).  The amounts are integer-valued (integer endowments, landscape, metabolisms and exchange amounts) and the
// metabolisms small, so the caller passes a TABLE of the host libm's own x ** (ms / (ms + mp)) for every metabolism pair
// and amount (pow_table[ms - 1][mp - 1][x], like the exp table of csrc/epstein.cu): bit-identical by construction.  Amounts
// beyond the table fall back to CUDA's pow (<= 2 ulp) and are counted in the status words.
#include "common.cuh"
#include "philox.cuh"
#include "sort.cuh"
#include <string.h>

namespace {

constexpr uint32_t kNoRank = 0xffffffffu;
constexpr int kThreads = 256;
constexpr int kSlots = 4;                 // pending ranks per cell
constexpr int kMaxVision = 7;             // neighbourhood of 112 cells = 4 per lane
constexpr int kPer = 4;
constexpr unsigned kSpinLimit = 1u << 22;

struct SugWs {
    uint64_t* keys0;
    uint64_t* keys1;
    uint32_t* vals0;
    uint32_t* vals1;
    uint4* pending;      // [ncells] ranks of the traders that start the step in the cell and have not acted yet
    uint32_t* nstart;    // [ncells]
    uint4* occupants;    // [ncells] rows of the traders standing in the cell (trade phase)
    uint32_t* counters;  // [0] ticket, [1] status (1: spin limit, 2: more than 4 traders started in one cell), [2] deaths, [3] trades
    void* sort_ws;
    size_t sort_ws_bytes;
};

size_t carve(SugWs* ws, char* base, int64_t n, int64_t ncells) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    SugWs t;
    t.keys0 = (uint64_t*)take(8 * n);
    t.keys1 = (uint64_t*)take(8 * n);
    t.vals0 = (uint32_t*)take(4 * n);
    t.vals1 = (uint32_t*)take(4 * n);
    t.pending = (uint4*)take(16 * ncells);
    t.nstart = (uint32_t*)take(4 * ncells);
    t.occupants = (uint4*)take(16 * ncells);
    t.counters = (uint32_t*)take(4 * 16);
    t.sort_ws_bytes = mb::sort_workspace_bytes(n);
    t.sort_ws = take(t.sort_ws_bytes);
    if (ws) *ws = t;
    return off;
}

struct SugArgs {
    int32_t* cell;                // [n]
    double* sugar;                // [n]
    double* spice;                // [n]
    const int32_t* ms;            // [n] metabolism_sugar
    const int32_t* mp;            // [n] metabolism_spice
    const int32_t* vision;        // [n]
    const int64_t* uid;           // [n] unique ids
    uint8_t* dead;                // [n] out: starved in this step
    double* sugar_layer;          // [ncells]
    double* spice_layer;          // [ncells]
    int32_t* count;               // [ncells] traders in the cell
    const short2* voff;           // neighbourhood offsets of the largest vision in BFS order (centre excluded): vision v = the
                                  // first 2 v (v + 1) entries
    const short2* woff;           // [nw] offsets within 2 max-vision (dependency range)
    const double* pow_table;      // [pow_m][pow_m][pow_x] x ** (ms / (ms + mp)) from the host's libm (NULL: CUDA pow)
    uint32_t* approx;             // counter of pow calls that missed the table
    int64_t n;
    int width, height, nw, pow_m, pow_x;
    uint64_t seed;
    uint32_t epoch;
};

// x ** (ms / (ms + mp)): the host's value when (ms, mp, x) is in the table
__device__ __forceinline__ double pow_ref(const SugArgs& a, double x, int ms, int mp) {
    if (a.pow_table != nullptr && ms >= 1 && mp >= 1 && ms <= a.pow_m && mp <= a.pow_m && x >= 0.0 && x < (double)a.pow_x) {
        const int xi = (int)x;
        if ((double)xi == x) return __ldg(a.pow_table + ((size_t)((ms - 1) * a.pow_m + (mp - 1)) * (size_t)a.pow_x + (size_t)xi));
    }
    if (a.approx != nullptr) atomicAdd(a.approx, 1u);
    return pow(x, (double)ms / (double)(ms + mp));
}

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__global__ void sug_keys(const int64_t* __restrict__ uid, int64_t n, uint64_t seed, uint32_t epoch, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = mb::priority64(seed, epoch, (uint64_t)(uid[i] - 1));
    vals[i] = (uint32_t)i;
}

// model.py:123-124
__global__ void sug_regrow(double* __restrict__ sugar, const double* __restrict__ sugar_max, double* __restrict__ spice,
                           const double* __restrict__ spice_max, int64_t ncells) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    sugar[c] = fmin(sugar[c] + 1.0, sugar_max[c]);
    spice[c] = fmin(spice[c] + 1.0, spice_max[c]);
}

__global__ void sug_prepare(const int32_t* __restrict__ cell, const uint32_t* __restrict__ order, int64_t n, SugWs ws,
                            uint8_t* __restrict__ dead) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t a = order[r];
    dead[a] = 0;
    const int c = cell[a];
    const uint32_t slot = atomicAdd(&ws.nstart[c], 1u);
    if (slot < (uint32_t)kSlots) reinterpret_cast<uint32_t*>(ws.pending + c)[slot] = (uint32_t)r;
    else ws.counters[1] = 2u;
}

__global__ void sug_count(const int32_t* __restrict__ cell, int64_t n, int32_t* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[cell[i]], 1);
}

// math.isclose(a, b, rel_tol) with abs_tol = 0 (CPython Modules/mathmodule.c)
__device__ __forceinline__ bool is_close(double a, double b, double rel_tol) {
    if (a == b) return true;
    const double diff = fabs(b - a);
    return (diff <= fabs(rel_tol * b)) || (diff <= fabs(rel_tol * a));
}

__device__ void act_warp(const SugArgs& a, uint32_t agent, int p, int x, int y, int lane) {
    const int v = __ldg(a.vision + agent);
    const int nv = 2 * v * (v + 1);
    const double own_sugar = __ldcg(a.sugar + agent), own_spice = __ldcg(a.spice + agent);
    const int ms = __ldg(a.ms + agent), mp = __ldg(a.mp + agent);
    // 1. the empty cells of the neighbourhood (the centre holds the trader itself: never empty) and their welfare
    int c[kPer];
    double w[kPer], dist[kPer];
    bool cand[kPer];
    double best = -1.0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int i = q * 32 + lane;
        c[q] = -1;
        cand[q] = false;
        w[q] = -1.0;
        dist[q] = 0.0;
        if (i < nv) {
            const short2 d = __ldg(a.voff + i);
            const int nx = x + d.x, ny = y + d.y;
            if (nx >= 0 && nx < a.width && ny >= 0 && ny < a.height) {
                c[q] = nx * a.height + ny;
                dist[q] = sqrt((double)(d.x * d.x + d.y * d.y));     // agents.py:7-17
            }
        }
    }
    int cnt[kPer];
    double cs[kPer], cp[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        cnt[q] = c[q] >= 0 ? __ldcg(a.count + c[q]) : 1;
        cs[q] = c[q] >= 0 ? __ldcg(a.sugar_layer + c[q]) : 0.0;
        cp[q] = c[q] >= 0 ? __ldcg(a.spice_layer + c[q]) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        cand[q] = c[q] >= 0 && cnt[q] == 0;
        if (cand[q]) {
            w[q] = pow_ref(a, own_sugar + cs[q], ms, mp) * pow_ref(a, own_spice + cp[q], mp, ms);       // agents.py:57-70
            best = fmax(best, w[q]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    int dest = p;
    if (best >= 0.0) {     // some cell is empty (a welfare is never negative)
        // 2./3. the cells close to the best welfare, of those the closest ones (agents.py:236-256)
        double dmin = 1e300;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            cand[q] = cand[q] && is_close(w[q], best, 1e-9);
            if (cand[q]) dmin = fmin(dmin, dist[q]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        unsigned masks[kPer];
        int nfinal = 0;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            masks[q] = __ballot_sync(0xffffffffu, cand[q] && is_close(dist[q], dmin, 1e-2));
            nfinal += __popc(masks[q]);
        }
        // 4. random.choice(final_candidates), BFS order = (q, lane)
        mb::DrawStream ds(a.seed, a.epoch, (uint64_t)(__ldg(a.uid + agent) - 1));
        int k = (int)ds.randbelow((uint64_t)nfinal);
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int nq = __popc(masks[q]);
            if (k >= 0 && k < nq) {
                const int src = (int)__fns(masks[q], 0u, (unsigned)k + 1u);
                dest = __shfl_sync(0xffffffffu, c[q], src);
                k = -1;
            } else if (k >= 0) {
                k -= nq;
            }
        }
    }
    if (lane == 0) {
        if (dest != p) {
            a.count[p] -= 1;
            a.count[dest] += 1;
            a.cell[agent] = dest;
        }
        // agents.py:264-271
        double s = own_sugar + __ldcg(a.sugar_layer + dest);
        a.sugar_layer[dest] = 0.0;
        s -= (double)ms;
        double sp = own_spice + __ldcg(a.spice_layer + dest);
        a.spice_layer[dest] = 0.0;
        sp -= (double)mp;
        a.sugar[agent] = s;
        a.spice[agent] = sp;
        if (s <= 0.0 || sp <= 0.0) {     // agents.py:273-281: remove()
            a.dead[agent] = 1;
            a.count[dest] -= 1;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
sug_chain(SugArgs a, SugWs ws, const uint32_t* __restrict__ order) {
    const int lane = threadIdx.x & 31;
    while (true) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&ws.counters[0], 1u);       // ranks are handed out in order: every lower rank is held by a
        r = __shfl_sync(0xffffffffu, r, 0);                      // running warp or done
        if ((int64_t)r >= a.n) break;
        const uint32_t agent = order[r];
        const int p = __ldcg(a.cell + agent);
        const int x = p / a.height, y = p - x * a.height;
        bool bad = false;
        for (int i0 = 0; i0 < a.nw && !bad; i0 += 32 * 8) {
            int wc[8];
            unsigned pend = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = i0 + q * 32 + lane;
                wc[q] = 0;
                if (i < a.nw) {
                    const short2 d = __ldg(a.woff + i);
                    const int nx = x + d.x, ny = y + d.y;
                    if (nx >= 0 && nx < a.width && ny >= 0 && ny < a.height) {
                        wc[q] = nx * a.height + ny;
                        pend |= 1u << q;
                    }
                }
            }
            unsigned spins = 0u;
            while (true) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!((pend >> q) & 1u)) continue;
                    const uint4 v = ld_relaxed_v4(ws.pending + wc[q]);
                    if (min(min(v.x, v.y), min(v.z, v.w)) >= r) pend &= ~(1u << q);
                }
                if (!__any_sync(0xffffffffu, pend != 0u)) break;
                if (++spins > kSpinLimit) { bad = true; break; }
            }
        }
        fence_acq_rel();
        __syncwarp();
        if (!bad) act_warp(a, agent, p, x, y, lane);
        __syncwarp();
        if (lane == 0) {
            if (bad) ws.counters[1] = 1u;
            fence_acq_rel();
            uint32_t* slots = reinterpret_cast<uint32_t*>(ws.pending + p);
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
                if (slots[s] == r) st_relaxed(slots + s, kNoRank);
        }
    }
}

// ---- trade_with_neighbors (agents.py:290-301, 156-207, 94-153) -----------------------------------------------------------
// A second shuffled activation: trader a trades with every trader in its radius-vision neighbourhood (cells in BFS order,
// centre excluded; a cell's traders in arrival order = ascending unique id, because traders only ever move to EMPTY
// cells) until no trade happens.  Nobody moves, so the same dependency graph applies with static positions: a and an
// earlier b interact only through a trader both can reach, i.e. if they stand within vision_a + vision_b.  The partners of
// a trader are a chain (each trade changes its own sugar and spice), so lane 0 walks them; the lanes only share the wait
// and the gathering of the partner list.  No randomness: only the order matters.
struct TradeOut {
    uint32_t* rec_trader;   // [cap] row of the trader whose turn it was
    int64_t* rec_partner;   // [cap] unique id of the partner
    double* rec_price;      // [cap]
    uint32_t cap;
};

__global__ void sug_trade_prepare(const int32_t* __restrict__ cell, const uint32_t* __restrict__ order, int64_t n, SugWs ws,
                                  uint4* __restrict__ occupants) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t a = order[r];
    const int c = cell[a];
    const uint32_t slot = atomicAdd(&ws.nstart[c], 1u);
    if (slot < (uint32_t)kSlots) {
        reinterpret_cast<uint32_t*>(ws.pending + c)[slot] = (uint32_t)r;
        reinterpret_cast<uint32_t*>(occupants + c)[slot] = a;
    } else {
        ws.counters[1] = 2u;
    }
}

__device__ __forceinline__ double welfare_of(const SugArgs& a, double sugar, double spice, int ms, int mp) {
    return pow_ref(a, sugar, ms, mp) * pow_ref(a, spice, mp, ms);             // agents.py:57-70
}

// agents.py:113-153 with `s` the spice seller and `b` the other side; returns whether the trade was executed
__device__ __forceinline__ bool maybe_sell_spice(const SugArgs& a, double& s_sugar, double& s_spice, int s_ms, int s_mp, double w_s,
                                                 double& b_sugar, double& b_spice, int b_ms, int b_mp, double w_b, double price) {
    double sugar_ex, spice_ex;
    if (price >= 1.0) { sugar_ex = 1.0; spice_ex = trunc(price); }            // agents.py:94-106 (int() truncates)
    else { sugar_ex = trunc(1.0 / price); spice_ex = 1.0; }
    const double ns_sugar = s_sugar + sugar_ex, nb_sugar = b_sugar - sugar_ex;
    const double ns_spice = s_spice - spice_ex, nb_spice = b_spice + spice_ex;
    if (ns_sugar <= 0.0 || nb_sugar <= 0.0 || ns_spice <= 0.0 || nb_spice <= 0.0) return false;
    const bool better = (w_s < welfare_of(a, ns_sugar, ns_spice, s_ms, s_mp)) && (w_b < welfare_of(a, nb_sugar, nb_spice, b_ms, b_mp));
    const bool not_crossing = (ns_spice / (double)s_mp) / (ns_sugar / (double)s_ms) > (nb_spice / (double)b_mp) / (nb_sugar / (double)b_ms);
    if (!(better && not_crossing)) return false;
    s_sugar = ns_sugar; s_spice = ns_spice; b_sugar = nb_sugar; b_spice = nb_spice;
    return true;
}

__global__ void __launch_bounds__(kThreads)
sug_trade_chain(SugArgs a, SugWs ws, const uint32_t* __restrict__ order, const uint4* __restrict__ occupants, TradeOut out) {
    __shared__ uint32_t s_partners[kThreads / 32][kPer * 32 * kSlots];      // partner rows of the warp's trader, BFS order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* plist = s_partners[warp];
    while (true) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(&ws.counters[0], 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if ((int64_t)r >= a.n) break;
        const uint32_t agent = order[r];
        const int p = __ldg(a.cell + agent);                                 // nobody moves in this phase
        const int x = p / a.height, y = p - x * a.height;
        bool bad = false;
        for (int i0 = 0; i0 < a.nw && !bad; i0 += 32 * 8) {
            int wc[8];
            unsigned pend = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = i0 + q * 32 + lane;
                wc[q] = 0;
                if (i < a.nw) {
                    const short2 d = __ldg(a.woff + i);
                    const int nx = x + d.x, ny = y + d.y;
                    if (nx >= 0 && nx < a.width && ny >= 0 && ny < a.height) {
                        wc[q] = nx * a.height + ny;
                        pend |= 1u << q;
                    }
                }
            }
            unsigned spins = 0u;
            while (true) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (!((pend >> q) & 1u)) continue;
                    const uint4 v = ld_relaxed_v4(ws.pending + wc[q]);
                    if (min(min(v.x, v.y), min(v.z, v.w)) >= r) pend &= ~(1u << q);
                }
                if (!__any_sync(0xffffffffu, pend != 0u)) break;
                if (++spins > kSpinLimit) { bad = true; break; }
            }
        }
        fence_acq_rel();
        __syncwarp();
        if (!bad) {
            // the partner list: occupants of the neighbourhood cells in BFS order, a cell's traders by ascending unique id
            const int v = __ldg(a.vision + agent);
            const int nv = 2 * v * (v + 1);
            int total = 0;
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const int i = q * 32 + lane;
                uint32_t rows[kSlots] = {kNoRank, kNoRank, kNoRank, kNoRank};
                int cnt = 0;
                if (i < nv) {
                    const short2 d = __ldg(a.voff + i);
                    const int nx = x + d.x, ny = y + d.y;
                    if (nx >= 0 && nx < a.width && ny >= 0 && ny < a.height) {
                        const int c = nx * a.height + ny;
                        cnt = min((int)__ldg(ws.nstart + c), kSlots);
                        if (cnt > 0) {
                            const uint4 o = __ldg(occupants + c);
                            rows[0] = o.x; rows[1] = o.y; rows[2] = o.z; rows[3] = o.w;
                            // arrival order = ascending unique id (at most 4: insertion sort)
                            for (int s1 = 1; s1 < cnt; ++s1)
                                for (int s2 = s1; s2 > 0 && __ldg(a.uid + rows[s2 - 1]) > __ldg(a.uid + rows[s2]); --s2) {
                                    const uint32_t t = rows[s2]; rows[s2] = rows[s2 - 1]; rows[s2 - 1] = t;
                                }
                        }
                    }
                }
                // exclusive prefix of the counts over the lanes of this q
                int incl = cnt;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o2);
                    if (lane >= o2) incl += t;
                }
                const int base = total + incl - cnt;
                for (int s1 = 0; s1 < cnt; ++s1) plist[base + s1] = rows[s1];
                total += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            if (lane == 0 && total > 0) {
                double my_sugar = __ldcg(a.sugar + agent), my_spice = __ldcg(a.spice + agent);
                const int my_ms = __ldg(a.ms + agent), my_mp = __ldg(a.mp + agent);
                for (int k = 0; k < total; ++k) {
                    const uint32_t b = plist[k];
                    double o_sugar = __ldcg(a.sugar + b), o_spice = __ldcg(a.spice + b);
                    const int o_ms = __ldg(a.ms + b), o_mp = __ldg(a.mp + b);
                    const int64_t o_uid = __ldg(a.uid + b);
                    bool changed = false;
                    while (true) {                                          // agents.py:156-207 (the recursion as a loop)
                        const double mrs_a = (my_spice / (double)my_mp) / (my_sugar / (double)my_ms);       // agents.py:84-92
                        const double mrs_b = (o_spice / (double)o_mp) / (o_sugar / (double)o_ms);
                        if (is_close(mrs_a, mrs_b, 1e-9)) break;
                        const double w_a = welfare_of(a, my_sugar, my_spice, my_ms, my_mp);
                        const double w_b = welfare_of(a, o_sugar, o_spice, o_ms, o_mp);
                        const double price = sqrt(mrs_a * mrs_b);
                        const bool sold = mrs_a > mrs_b
                            ? maybe_sell_spice(a, my_sugar, my_spice, my_ms, my_mp, w_a, o_sugar, o_spice, o_ms, o_mp, w_b, price)
                            : maybe_sell_spice(a, o_sugar, o_spice, o_ms, o_mp, w_b, my_sugar, my_spice, my_ms, my_mp, w_a, price);
                        if (!sold) break;
                        changed = true;
                        const uint32_t slot = atomicAdd(&ws.counters[3], 1u);
                        if (slot < out.cap) {
                            out.rec_trader[slot] = agent;
                            out.rec_partner[slot] = o_uid;
                            out.rec_price[slot] = price;
                        }
                    }
                    if (changed) { a.sugar[b] = o_sugar; a.spice[b] = o_spice; }
                }
                a.sugar[agent] = my_sugar;
                a.spice[agent] = my_spice;
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (bad) ws.counters[1] = 1u;
            fence_acq_rel();
            uint32_t* slots = reinterpret_cast<uint32_t*>(ws.pending + p);
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
                if (slots[s] == r) st_relaxed(slots + s, kNoRank);
        }
    }
}

__global__ void sug_deaths(const uint8_t* __restrict__ dead, int64_t n, uint32_t* counters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, i < n && dead[i] != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&counters[2], (unsigned)__popc(b));
}

}  // namespace

MB_API int mb_sugarscape_workspace_bytes(int64_t n, int64_t ncells, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0 && ncells >= 0, "mb_sugarscape_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, n, ncells);
    return MB_OK;
}

// count[cell] = number of traders in the cell (after placing traders / removing the starved)
MB_API int mb_sugarscape_build_count(const int32_t* cell, int64_t n, int32_t* count, int64_t ncells, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && ncells >= 0 && (ncells == 0 || count) && (n == 0 || cell), "mb_sugarscape_build_count: bad argument");
    int rc = mb::check_cuda(cudaMemsetAsync(count, 0, 4 * (size_t)ncells, stream), "memset");
    if (rc) return rc;
    if (n) sug_count<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(cell, n, count);
    MB_LAUNCH_CHECK("mb_sugarscape_build_count");
    return MB_OK;
}

// One model step without trade: regrowth (regrow != 0), then the traders' shuffle_do("step") with shuffle number `epoch`.
// dead [n] (uint8, out) marks the traders that starved: the caller compacts them away.  status_dev (device, 3 x uint32,
// optional): status (0 valid; 1 a wait hit its spin limit; 2 more than 4 traders started the step in one cell), deaths, pow
// calls that missed pow_table (0 = every welfare is the host libm's value).
MB_API int mb_sugarscape_step(int32_t* cell, double* sugar, double* spice, const int32_t* metabolism_sugar,
                              const int32_t* metabolism_spice, const int32_t* vision, const int64_t* unique_id, uint8_t* dead,
                              int64_t n, double* sugar_layer, const double* sugar_max, double* spice_layer,
                              const double* spice_max, int32_t* count, int64_t width, int64_t height, const int16_t* voff,
                              int max_vision, const int16_t* woff, int nw, int regrow, const double* pow_table, int pow_m,
                              int pow_x, uint64_t seed, uint32_t epoch, uint32_t* status_dev, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_sugarscape_step: n out of range");
    MB_REQUIRE(width > 0 && height > 0 && width * height < (1ll << 31), "mb_sugarscape_step: grid out of range");
    MB_REQUIRE(max_vision >= 1 && max_vision <= kMaxVision && nw > 0, "mb_sugarscape_step: vision must be 1..7");
    MB_REQUIRE(sugar_layer && spice_layer && count && workspace && voff && woff, "mb_sugarscape_step: null buffer");
    MB_REQUIRE(!regrow || (sugar_max && spice_max), "mb_sugarscape_step: regrowth needs the landscape maxima");
    const int64_t ncells = width * height;
    SugWs ws;
    if (carve(&ws, (char*)workspace, n, ncells) > workspace_bytes) {
        mb::set_error("mb_sugarscape_step: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    if (regrow) sug_regrow<<<(unsigned)mb::ceil_div(ncells, 256), 256, 0, stream>>>(sugar_layer, sugar_max, spice_layer, spice_max, ncells);
    int rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 4 * 16, stream), "memset");
    if (rc) return rc;
    if (n > 0) {
        MB_REQUIRE(cell && sugar && spice && metabolism_sugar && metabolism_spice && vision && unique_id && dead, "mb_sugarscape_step: null agent column");
        SugArgs a;
        a.cell = cell; a.sugar = sugar; a.spice = spice; a.ms = metabolism_sugar; a.mp = metabolism_spice; a.vision = vision;
        a.uid = unique_id; a.dead = dead; a.sugar_layer = sugar_layer; a.spice_layer = spice_layer; a.count = count;
        a.voff = reinterpret_cast<const short2*>(voff); a.woff = reinterpret_cast<const short2*>(woff);
        a.n = n; a.width = (int)width; a.height = (int)height; a.nw = nw; a.seed = seed; a.epoch = epoch;
        a.pow_table = pow_table; a.pow_m = pow_m; a.pow_x = pow_x; a.approx = ws.counters + 4;
        const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
        sug_keys<<<blocks, 256, 0, stream>>>(unique_id, n, seed, epoch, ws.keys0, ws.vals0);
        const int where = mb::radix_sort_pairs<uint64_t>(ws.keys0, ws.vals0, ws.keys1, ws.vals1, n, 0, 64, ws.sort_ws, ws.sort_ws_bytes, stream);
        if (where < 0) return where;
        const uint32_t* order = where ? ws.vals1 : ws.vals0;
        if ((rc = mb::check_cuda(cudaMemsetAsync(ws.pending, 0xff, 16 * (size_t)ncells, stream), "memset"))) return rc;
        if ((rc = mb::check_cuda(cudaMemsetAsync(ws.nstart, 0, 4 * (size_t)ncells, stream), "memset"))) return rc;
        sug_prepare<<<blocks, 256, 0, stream>>>(cell, order, n, ws, dead);
        int64_t ctas = mb::ceil_div(n, kThreads / 32);
        const int64_t cap = (int64_t)mb::sm_count() * 8;
        if (ctas > cap) ctas = cap;
        sug_chain<<<(unsigned)ctas, kThreads, 0, stream>>>(a, ws, order);
        sug_deaths<<<blocks, 256, 0, stream>>>(dead, n, ws.counters);
    }
    if (status_dev != nullptr) {
        if ((rc = mb::check_cuda(cudaMemcpyAsync(status_dev, ws.counters + 1, 4 * 2, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
        if ((rc = mb::check_cuda(cudaMemcpyAsync(status_dev + 2, ws.counters + 4, 4, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    }
    MB_LAUNCH_CHECK("mb_sugarscape_step");
    return MB_OK;
}

// The trade activation of one model step (model.py:128-129: `self.agents.shuffle_do("trade_with_neighbors")`, shuffle number
// `epoch`).  Every executed trade is appended to (rec_trader = row of the trader whose turn it was, rec_partner = unique id of
// the other side, rec_price), a trader's trades in the order it made them; status_dev (device, 3 x uint32): status as in
// mb_sugarscape_step, number of trades, pow calls that missed pow_table; the number of trades (may exceed rec_cap: the caller retries with larger buffers -- nothing else is lost,
// but the step has been applied).
MB_API int mb_sugarscape_trade(const int32_t* cell, double* sugar, double* spice, const int32_t* metabolism_sugar,
                               const int32_t* metabolism_spice, const int32_t* vision, const int64_t* unique_id, int64_t n,
                               int64_t width, int64_t height, const int16_t* voff, int max_vision, const int16_t* woff, int nw,
                               const double* pow_table, int pow_m, int pow_x, uint64_t seed, uint32_t epoch, uint32_t* rec_trader,
                               int64_t* rec_partner, double* rec_price, int64_t rec_cap, uint32_t* status_dev, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_sugarscape_trade: n out of range");
    MB_REQUIRE(width > 0 && height > 0 && width * height < (1ll << 31), "mb_sugarscape_trade: grid out of range");
    MB_REQUIRE(max_vision >= 1 && max_vision <= kMaxVision && nw > 0, "mb_sugarscape_trade: vision must be 1..7");
    MB_REQUIRE(workspace && voff && woff && rec_cap >= 0 && rec_cap < (1ll << 32), "mb_sugarscape_trade: bad argument");
    MB_REQUIRE(rec_cap == 0 || (rec_trader && rec_partner && rec_price), "mb_sugarscape_trade: null record buffer");
    const int64_t ncells = width * height;
    SugWs ws;
    if (carve(&ws, (char*)workspace, n, ncells) > workspace_bytes) {
        mb::set_error("mb_sugarscape_trade: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    int rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 4 * 16, stream), "memset");
    if (rc) return rc;
    if (n > 0) {
        MB_REQUIRE(cell && sugar && spice && metabolism_sugar && metabolism_spice && vision && unique_id, "mb_sugarscape_trade: null agent column");
        SugArgs a;
        memset(&a, 0, sizeof(a));
        a.cell = const_cast<int32_t*>(cell); a.sugar = sugar; a.spice = spice; a.ms = metabolism_sugar; a.mp = metabolism_spice;
        a.vision = vision; a.uid = unique_id;
        a.voff = reinterpret_cast<const short2*>(voff); a.woff = reinterpret_cast<const short2*>(woff);
        a.n = n; a.width = (int)width; a.height = (int)height; a.nw = nw; a.seed = seed; a.epoch = epoch;
        a.pow_table = pow_table; a.pow_m = pow_m; a.pow_x = pow_x; a.approx = ws.counters + 4;
        const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
        sug_keys<<<blocks, 256, 0, stream>>>(unique_id, n, seed, epoch, ws.keys0, ws.vals0);
        const int where = mb::radix_sort_pairs<uint64_t>(ws.keys0, ws.vals0, ws.keys1, ws.vals1, n, 0, 64, ws.sort_ws, ws.sort_ws_bytes, stream);
        if (where < 0) return where;
        const uint32_t* order = where ? ws.vals1 : ws.vals0;
        if ((rc = mb::check_cuda(cudaMemsetAsync(ws.pending, 0xff, 16 * (size_t)ncells, stream), "memset"))) return rc;
        if ((rc = mb::check_cuda(cudaMemsetAsync(ws.nstart, 0, 4 * (size_t)ncells, stream), "memset"))) return rc;
        sug_trade_prepare<<<blocks, 256, 0, stream>>>(cell, order, n, ws, ws.occupants);
        int64_t ctas = mb::ceil_div(n, kThreads / 32);
        const int64_t cap = (int64_t)mb::sm_count() * 8;
        if (ctas > cap) ctas = cap;
        TradeOut out{rec_trader, rec_partner, rec_price, (uint32_t)rec_cap};
        sug_trade_chain<<<(unsigned)ctas, kThreads, 0, stream>>>(a, ws, order, ws.occupants, out);
    }
    if (status_dev != nullptr) {
        if ((rc = mb::check_cuda(cudaMemcpyAsync(status_dev, ws.counters + 1, 4, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
        if ((rc = mb::check_cuda(cudaMemcpyAsync(status_dev + 1, ws.counters + 3, 4 * 2, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    }
    MB_LAUNCH_CHECK("mb_sugarscape_trade");
    return MB_OK;
}
