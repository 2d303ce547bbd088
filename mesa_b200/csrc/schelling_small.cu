// The WHOLE Schelling model step of a small grid in ONE launch of ONE CTA (BASELINE config 1: 100 x 100).
//
//   mesa/examples/basic/schelling/model.py:87-93   happy = 0; agents.shuffle_do("step"); agents.do("assign_state")
//   agents.py:44-47 / grid.py:288-316 / cell_agent.py:39-54   the relocation of the unhappy agents (csrc/relocate.cu)
//   agents.py:24-42                                            the happiness of every agent (csrc/schelling.cu)
//
// A 100 x 100 grid is 10 KB of state: the multi-launch path (collect, passes to the fixed point with a read-back of
// the "changed" flag after each, apply, stencil) costs ~90 us per step of launch and host latency for ~2 us of work.
// Here the 16-byte relocation slots of all cells live in SHARED memory (<= 12800 cells), the same fixed-point
// iteration as csrc/relocate.cu runs between __syncthreads() barriers -- every mover re-evaluates its probe
// sequence, claims with a shared-memory atomicMin and releases with a compare-and-swap, the 50-probe miss takes the
// k-th empty cell in row-major order at its time, and a pass that changes nothing ends it --, then origins are vacated,
// arrivals written, and the happiness stencil runs from a shared-memory copy of the type layer (the slot table's
// memory, free by then).  Nothing returns to the host: model.happy and the mover count are device words.
//
// The probe cells of a mover are state-independent, so they are drawn once (Philox, ~100 instructions per block) and
// cached as 16-bit cell numbers; later passes only walk the cache.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr unsigned long long kNever = ~0ull;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kMaxProbes = 50;          // grid.py:303
constexpr int kCache = 12;              // cached probe cells per mover (a mover needs ~5; 0.8^12 = 7 % need more)
constexpr int kThreads = 1024;
constexpr int kMaxCells = 12800;        // 16 B of shared memory per cell
constexpr int kFbCap = 1024;            // fallback movers per pass (more: error, the multi-launch path serves them)

struct Slot {
    unsigned long long vac, arr;
};

struct SmallRule {
    uint8_t min_similar[232];           // min_similar[valid]: smallest `similar` that is happy (> valid: never)
};

// per-mover state (global memory, L2-resident; caller-owned workspace)
struct SmallWs {
    unsigned long long* prio;           // [cap]
    uint32_t* origin;                   // [cap]
    uint32_t* agent;                    // [cap]
    uint32_t* choice;                   // [cap]
    uint32_t* fbk;                      // [cap] cached k = _randbelow(nempty) of a fallback-state mover (kNone: not drawn)
    uint16_t* cache;                    // [cap][kCache] probe cells
    uint16_t* after;                    // [cap] draws consumed by the cached probes
    uint8_t* ncached;                   // [cap]
    int8_t* mtype;                      // [cap]
};

size_t carve(SmallWs* ws, char* base, int64_t cap) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    SmallWs t;
    t.prio = (unsigned long long*)take(8 * cap);
    t.origin = (uint32_t*)take(4 * cap);
    t.agent = (uint32_t*)take(4 * cap);
    t.choice = (uint32_t*)take(4 * cap);
    t.fbk = (uint32_t*)take(4 * cap);
    t.cache = (uint16_t*)take(2 * kCache * cap);
    t.after = (uint16_t*)take(2 * cap);
    t.ncached = (uint8_t*)take(cap);
    t.mtype = (int8_t*)take(cap);
    if (ws) *ws = t;
    return off;
}

__device__ __forceinline__ bool empty_at(const Slot& s, unsigned long long t) { return s.vac <= t && t <= s.arr; }

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_sums, unsigned& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    __syncthreads();
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    unsigned before = 0, tot = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
        const unsigned s = warp_sums[w];
        before += w < warp ? s : 0u;
        tot += s;
    }
    total = tot;
    return before + incl - v;
}

// do("assign_state") on a PADDED shared-memory copy of the type layer: radius R cells of border around the grid -- empty
// (-1) outside a clamped grid, the wrapped cells on a torus -- so the (2R+1)^2 - 1 neighbour reads need no bounds checks
// and the loops unroll (the generic loop below costs ~365 instructions per cell, 80 % of a step without movers:
// profiles/r2_schelling_small_ncu.csv).  Returns this thread's number of happy agents.
template <int R>
__device__ __forceinline__ unsigned assign_state_padded(int8_t* __restrict__ tl, const int8_t* __restrict__ type,
                                                        uint8_t* __restrict__ happy, uint8_t* __restrict__ empty_layer, int rows,
                                                        int cols, int torus, const SmallRule& rule) {
    const int tid = threadIdx.x;
    const int pc = cols + 2 * R, pr = rows + 2 * R;
    {   // fill: padded index p = pi * pc + pj walks in steps of kThreads without divisions
        int pi = tid / pc, pj = tid - pi * pc;
        const int di = kThreads / pc, dj = kThreads - di * pc;
        for (int p = tid; p < pr * pc; p += kThreads) {
            int i = pi - R, j = pj - R;
            int8_t v = -1;
            if (torus) {
                i += i < 0 ? rows : 0; i -= i >= rows ? rows : 0;
                j += j < 0 ? cols : 0; j -= j >= cols ? cols : 0;
                v = __ldcg(type + i * cols + j);
            } else if (i >= 0 && i < rows && j >= 0 && j < cols) {
                v = __ldcg(type + i * cols + j);
            }
            tl[p] = v;
            pi += di; pj += dj;
            if (pj >= pc) { pj -= pc; ++pi; }
        }
    }
    __syncthreads();
    unsigned nh = 0;
    int i = tid / cols, j = tid - i * cols;
    const int di = kThreads / cols, dj = kThreads - di * cols;
    for (int c = tid; c < rows * cols; c += kThreads) {
        const int8_t* centre = tl + (i + R) * pc + (j + R);
        const int t = *centre;
        unsigned h = 0;
        if (t >= 0) {
            int similar = 0, valid = 0;
#pragma unroll
            for (int a = -R; a <= R; ++a)
#pragma unroll
                for (int b = -R; b <= R; ++b) {
                    if (a == 0 && b == 0) continue;
                    const int u = centre[a * pc + b];
                    valid += u >= 0;
                    similar += u == t;
                }
            h = similar >= (int)rule.min_similar[valid];
        }
        happy[c] = (uint8_t)h;
        if (empty_layer != nullptr) empty_layer[c] = t < 0;
        nh += h;
        i += di; j += dj;
        if (j >= cols) { j -= cols; ++i; }
    }
    return nh;
}

// status word: 0 ok, 1 grid completely full with an unhappy agent (ValueError, grid.py:311-315), 2 more movers than
// the workspace holds, 3 more than kFbCap fallback movers in a pass, 4 no convergence
__global__ void __launch_bounds__(kThreads, 1)
schelling_small_step(int8_t* __restrict__ type, uint8_t* __restrict__ happy, uint32_t* __restrict__ agent_id,
                     uint32_t* __restrict__ agent_cell, uint8_t* __restrict__ empty_layer, int rows, int cols, int radius,
                     int torus, SmallRule rule, uint64_t seed, uint32_t epoch0, int relocate, int nsteps, int happy_current, SmallWs ws, int cap,
                     unsigned long long* __restrict__ out /* [0] model.happy [1] movers [2] status [3] passes */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Slot* slots = reinterpret_cast<Slot*>(smem_raw);
    __shared__ unsigned warp_sums[kThreads / 32];
    __shared__ unsigned s_nmovers, s_nocc, s_changed, s_nfb, s_pick, s_status, s_happy;
    __shared__ uint32_t fb_list[kFbCap];
    const int tid = threadIdx.x;
    const int ncells = rows * cols;
    // nsteps > 1: model.run_for(n) as ONE launch -- the steps of a 100 x 100 grid are shorter than the host's launch path
    // (~10 us of kernel against ~20 us to issue a launch from Python), shuffle number epoch0 + s for step s
    for (int step = 0; step < nsteps; ++step) {
    const uint32_t epoch = epoch0 + (uint32_t)step;
    __syncthreads();                       // the previous step's last readers of the shared words
    if (tid == 0) { s_nmovers = 0; s_nocc = 0; s_status = 0; s_happy = 0; s_changed = 0; s_nfb = 0; }
    __syncthreads();
    unsigned passes = 0;
    if (relocate) {
        // ---- slots + movers (movers = occupied && !happy at the start of the step, agents.py:46).  All of a thread's cells
        // are loaded first (independent loads: one L2 round trip instead of one per 1024-cell slice)
        constexpr int kSlices = (kMaxCells + kThreads - 1) / kThreads;
        int8_t t_reg[kSlices];
        uint8_t h_reg[kSlices];
        uint32_t a_reg[kSlices];
#pragma unroll
        for (int k = 0; k < kSlices; ++k) {
            const int c = k * kThreads + tid;
            const bool in = c < ncells;
            t_reg[k] = in ? type[c] : (int8_t)-1;
            h_reg[k] = in ? happy[c] : (uint8_t)1;
            a_reg[k] = in ? agent_id[c] : 0u;
        }
        // a warp reserves the slots of ALL its movers with one atomic (one per slice and warp was 832 atomics on two words)
        unsigned wmovers = 0, wocc = 0, whappy = 0;
#pragma unroll
        for (int k = 0; k < kSlices; ++k) {
            const int t = t_reg[k];
            wmovers += __popc(__ballot_sync(0xffffffffu, t >= 0 && h_reg[k] == 0));
            wocc += __popc(__ballot_sync(0xffffffffu, t >= 0));
            whappy += __popc(__ballot_sync(0xffffffffu, t >= 0 && h_reg[k] != 0));
        }
        unsigned next = 0;
        if ((tid & 31) == 0) {
            if (wmovers) next = atomicAdd(&s_nmovers, wmovers);
            if (wocc) atomicAdd(&s_nocc, wocc);
            if (whappy) atomicAdd(&s_happy, whappy);
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        // A step without movers changes nothing: if the happy layer on entry IS the evaluation of the type layer (always
        // inside a model run: the previous step's do("assign_state") wrote it), the step is over -- a settled model then
        // costs the load of its layers per step instead of a slot table and a stencil (BASELINE config 1 settles after ~30
        // of its 100 steps).
        __syncthreads();
        if (s_nmovers == 0u && (step > 0 || happy_current)) {
            if (tid == 0) { out[0] = s_happy; out[1] = 0ull; out[2] = 0ull; out[3] = 0ull; }
            continue;
        }
        __syncthreads();
        if (tid == 0) s_happy = 0u;     // (recounted by the stencil below)
#pragma unroll
        for (int k = 0; k < kSlices; ++k) {
            const int c = k * kThreads + tid;
            if (k * kThreads >= ncells) break;
            const int t = t_reg[k];
            const bool mover = t >= 0 && h_reg[k] == 0;
            const unsigned ballot = __ballot_sync(0xffffffffu, mover);
            unsigned long long vac = t >= 0 ? kNever : 0ull;
            if (mover) {
                const unsigned m = next + __popc(ballot & ((1u << (tid & 31)) - 1u));
                const uint32_t agent = a_reg[k];
                const unsigned long long pr = mb::priority64(seed, epoch, agent);
                vac = pr + 1ull;
                if (m < (unsigned)cap) {
                    ws.prio[m] = pr; ws.origin[m] = (uint32_t)c; ws.agent[m] = agent; ws.choice[m] = kNone;
                    ws.fbk[m] = kNone; ws.ncached[m] = 0; ws.after[m] = 0; ws.mtype[m] = (int8_t)t;
                }
            }
            next += __popc(ballot);
            if (c < ncells) slots[c] = Slot{vac, kNever};
        }
        __syncthreads();
        const unsigned nmovers = s_nmovers;
        const unsigned nempty = (unsigned)ncells - s_nocc;
        if (nmovers > (unsigned)cap) { if (tid == 0) s_status = 2; }
        else if (nmovers > 0 && nempty == 0) { if (tid == 0) s_status = 1; }
        __syncthreads();
        if (s_status == 0 && nmovers > 0) {
            // ---- fixed-point passes
            for (;;) {
                ++passes;
                if (passes > 100000u) { if (tid == 0) s_status = 4; break; }
                for (unsigned m = tid; m < nmovers; m += kThreads) {
                    // everything the pass needs of this mover in ONE round trip (independent loads)
                    const unsigned long long t = ws.prio[m];
                    const int nc = ws.ncached[m];
                    const uint32_t held = ws.choice[m];
                    const uint2* crow = reinterpret_cast<const uint2*>(ws.cache + (size_t)m * kCache);   // 24-byte rows
                    const uint2 c01 = crow[0], c23 = crow[1], c45 = crow[2];
                    const uint32_t cw[6] = {c01.x, c01.y, c23.x, c23.y, c45.x, c45.y};
                    uint32_t pick = kNone;
#pragma unroll
                    for (int j = 0; j < kCache; ++j) {
                        if (j >= nc || pick != kNone) continue;
                        const uint32_t c = (cw[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                        if (empty_at(slots[c], t)) pick = c;
                    }
                    if (pick == kNone) {
                        // beyond the cached probes: draw on from the saved stream position, extending the cache while
                        // there is room (probe cells are state-independent)
                        mb::DrawStream ds(seed, epoch, ws.agent[m]);
                        ds.seek(ws.after[m]);
                        for (int j = nc; j < kMaxProbes; ++j) {
                            const uint32_t c = (uint32_t)ds.randbelow((uint64_t)ncells);
                            if (j < kCache) {
                                ws.cache[(size_t)m * kCache + j] = (uint16_t)c;
                                ws.ncached[m] = (uint8_t)(j + 1);
                                ws.after[m] = (uint16_t)ds.c;
                            }
                            if (empty_at(slots[c], t)) { pick = c; break; }
                        }
                        // all 50 probes occupied: k = _randbelow(nempty) is the next draw (grid.py:308-310)
                        if (pick == kNone && ws.fbk[m] == kNone) ws.fbk[m] = (uint32_t)ds.randbelow((uint64_t)nempty);
                    }
                    if (pick == kNone) {
                        const unsigned f = atomicAdd(&s_nfb, 1u);
                        if (f < (unsigned)kFbCap) fb_list[f] = m;
                        continue;
                    }
                    if (held != pick) {
                        if (held != kNone) atomicCAS(&slots[held].arr, t, kNever);
                        ws.choice[m] = pick;
                        atomicMin(&slots[pick].arr, t);
                        s_changed = 1u;
                    } else if (atomicMin(&slots[pick].arr, t) > t) {
                        s_changed = 1u;
                    }
                }
                __syncthreads();
                // ---- argwhere fallback (grid.py:308-316): the k-th empty cell in row-major order at the mover's time
                const unsigned nfb = s_nfb;
                if (nfb > (unsigned)kFbCap) { if (tid == 0) s_status = 3; break; }
                const int per = (ncells + kThreads - 1) / kThreads;
                for (unsigned f = 0; f < nfb; ++f) {
                    const uint32_t m = fb_list[f];
                    const unsigned long long t = ws.prio[m];
                    const uint32_t k = ws.fbk[m];
                    const int lo = tid * per, hi = min(lo + per, ncells);
                    unsigned n = 0;
                    for (int c = lo; c < hi; ++c) n += empty_at(slots[c], t);
                    unsigned total;
                    const unsigned excl = block_exclusive_scan(n, warp_sums, total);
                    if (tid == 0) s_pick = kNone;
                    __syncthreads();
                    if (k >= excl && k < excl + n) {
                        unsigned need = k - excl;
                        for (int c = lo; c < hi; ++c)
                            if (empty_at(slots[c], t)) {
                                if (need == 0) { s_pick = (uint32_t)c; break; }
                                --need;
                            }
                    }
                    __syncthreads();
                    if (tid == 0) {
                        const uint32_t pick = s_pick;
                        const uint32_t held = ws.choice[m];
                        if (pick == kNone) {
                            s_changed = 1u;   // cannot happen (the number of empty cells is constant in time)
                        } else if (held != pick) {
                            if (held != kNone) atomicCAS(&slots[held].arr, t, kNever);
                            ws.choice[m] = pick;
                            atomicMin(&slots[pick].arr, t);
                            s_changed = 1u;
                        } else if (atomicMin(&slots[pick].arr, t) > t) {
                            s_changed = 1u;
                        }
                    }
                    __syncthreads();
                }
                __syncthreads();
                const unsigned changed = s_changed;
                __syncthreads();
                if (tid == 0) { s_changed = 0; s_nfb = 0; }
                __syncthreads();
                if (!changed) break;
            }
            __syncthreads();
            if (s_status == 0) {
                // ---- cell_agent.py:39-54: every origin becomes empty, then every arrival is written
                for (unsigned m = tid; m < nmovers; m += kThreads) {
                    const uint32_t o = ws.origin[m];
                    type[o] = -1;
                    happy[o] = 0;
                }
                __syncthreads();
                for (unsigned m = tid; m < nmovers; m += kThreads) {
                    const uint32_t d = ws.choice[m];
                    const uint32_t agent = ws.agent[m];
                    type[d] = ws.mtype[m];
                    happy[d] = 0;
                    agent_id[d] = agent;
                    if (agent_cell != nullptr) agent_cell[agent] = d;
                }
            }
        }
        __syncthreads();
    }
    if (s_status != 0) {
        if (tid == 0) { out[1] = s_nmovers; out[2] = s_status; out[3] = passes; }
        return;
    }
    // ---- do("assign_state"), agents.py:24-42, from a shared-memory copy of the type layer
    int8_t* tl = reinterpret_cast<int8_t*>(smem_raw);
    __syncthreads();
    unsigned nh = 0;
    if (radius == 1) {
        nh = assign_state_padded<1>(tl, type, happy, empty_layer, rows, cols, torus, rule);
    } else if (radius == 2) {
        nh = assign_state_padded<2>(tl, type, happy, empty_layer, rows, cols, torus, rule);
    } else {
    for (int c = tid; c < ncells; c += kThreads) tl[c] = __ldcg(type + c);
    __syncthreads();
    for (int c = tid; c < ncells; c += kThreads) {
        const int t = tl[c];
        unsigned h = 0;
        if (t >= 0) {
            const int i = c / cols, j = c - i * cols;
            int similar = 0, valid = 0;
            for (int di = -radius; di <= radius; ++di) {
                int ii = i + di;
                if (torus) ii = (ii % rows + rows) % rows; else if (ii < 0 || ii >= rows) continue;
                for (int dj = -radius; dj <= radius; ++dj) {
                    if (di == 0 && dj == 0) continue;
                    int jj = j + dj;
                    if (torus) jj = (jj % cols + cols) % cols; else if (jj < 0 || jj >= cols) continue;
                    const int u = tl[ii * cols + jj];
                    valid += u >= 0;
                    similar += u == t;
                }
            }
            h = similar >= (int)rule.min_similar[valid];
        }
        happy[c] = (uint8_t)h;
        if (empty_layer != nullptr) empty_layer[c] = t < 0;
        nh += h;
    }
    }
    nh = warp_reduce_add(nh);
    if ((tid & 31) == 0 && nh) atomicAdd(&s_happy, nh);
    __syncthreads();
    if (tid == 0) { out[0] = s_happy; out[1] = s_nmovers; out[2] = 0ull; out[3] = passes; }
    }
}

}  // namespace

MB_API int mb_schelling_small_max_cells(void) { return kMaxCells; }

MB_API int mb_schelling_small_workspace_bytes(int64_t max_movers, size_t* out) {
    MB_REQUIRE(out != nullptr && max_movers >= 0, "mb_schelling_small_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, max_movers);
    return MB_OK;
}

// One whole model step of a grid of at most mb_schelling_small_max_cells() cells in ONE launch, nothing read back:
// relocate != 0: shuffle_do("step") of the agents that are unhappy according to the `happy` layer, then (always)
// do("assign_state"); nsteps > 1 repeats that with shuffle numbers epoch, epoch + 1, ... inside the same launch
// (model.run_for(n): out4 then describes the LAST step; a failed step ends the run).  happy_current != 0: the caller
// guarantees that the `happy` layer on entry is the evaluation of the `type` layer (true inside a model run) -- a step
// without movers then changes nothing and costs only the reading of the layers.  min_similar_host as in mb_schelling_happy_i8.  agent_cell / empty_layer optional.  out: device
// uint64[4] = {model.happy, movers, status (0 ok, 1 grid completely full -> ValueError, 2 workspace too small,
// 3 too many fallback movers, 4 no convergence), passes}.  A torus with a dimension < 2*radius+1 is rejected.
MB_API int mb_schelling_step_small(int8_t* type, uint8_t* happy, uint32_t* agent_id, uint32_t* agent_cell, uint8_t* empty_layer,
                                   int64_t rows, int64_t cols, int radius, int torus, const uint8_t* min_similar_host,
                                   uint64_t seed, uint32_t epoch, int relocate, int nsteps, int happy_current, void* workspace,
                                   size_t workspace_bytes, int64_t max_movers, uint64_t* out4, cudaStream_t stream) {
    MB_REQUIRE(rows > 0 && cols > 0 && rows * cols <= kMaxCells, "mb_schelling_step_small: the grid must have 1..12800 cells");
    MB_REQUIRE(radius >= 1 && radius <= 7, "mb_schelling_step_small: radius must be 1..7");
    MB_REQUIRE(type && happy && agent_id && min_similar_host && workspace && out4, "mb_schelling_step_small: null buffer");
    MB_REQUIRE(!(torus && (rows < 2 * radius + 1 || cols < 2 * radius + 1)),
               "mb_schelling_step_small: torus needs rows, cols >= 2*radius+1 (SURVEY A.2)");
    MB_REQUIRE(max_movers >= 0 && max_movers < (1ll << 31), "mb_schelling_step_small: bad max_movers");
    MB_REQUIRE(nsteps >= 1 && (nsteps == 1 || relocate), "mb_schelling_step_small: nsteps >= 1 (several steps only with relocate)");
    SmallWs ws;
    if (carve(&ws, (char*)workspace, max_movers) > workspace_bytes) {
        mb::set_error("mb_schelling_step_small: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    SmallRule rule;
    const int nmax = (2 * radius + 1) * (2 * radius + 1) - 1;
    for (int v = 0; v < 232; ++v) rule.min_similar[v] = v <= nmax ? min_similar_host[v] : 255;
    int smem = (int)(rows * cols) * (int)sizeof(Slot);
    const int padded = (int)((rows + 4) * (cols + 4));      // the stencil's padded copy of the type layer (radius <= 2)
    if (smem < padded) smem = padded;
    static bool attr_set = false;
    if (!attr_set) {
        int rc = mb::check_cuda(cudaFuncSetAttribute(schelling_small_step, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kMaxCells * (int)sizeof(Slot)), "cudaFuncSetAttribute");
        if (rc) return rc;
        attr_set = true;
    }
    schelling_small_step<<<1, kThreads, smem, stream>>>(type, happy, agent_id, agent_cell, empty_layer, (int)rows, (int)cols,
                                                       radius, torus, rule, seed, epoch, relocate, nsteps, happy_current, ws, (int)max_movers,
                                                       (unsigned long long*)out4);
    MB_LAUNCH_CHECK("mb_schelling_step_small");
    return MB_OK;
}
