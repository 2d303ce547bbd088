// Batched relocation of unhappy Schelling agents with deterministic, order-exact conflict
// resolution (K4+K5 of SURVEY.md §2.4).
//
// Replaces AgentSet.shuffle_do("step") (mesa/agentset.py:772-787) over
//   mesa/examples/basic/schelling/agents.py:44-47   if not self.happy: self.cell = grid.select_random_empty_cell()
//   mesa/discrete_space/grid.py:288-316             <=50 x random.choice(_celllist), first cell that is empty
//                                                   NOW wins; else argwhere(empty) + one more choice
//   mesa/discrete_space/cell_agent.py:39-54         add to the new cell, then remove from the old one
//
// The reference is sequential: agent #rho in the shuffled order sees the moves of agents
// < rho.  Here the activation order is the ascending Philox priority64 (philox.cuh) and the
// sequential outcome is computed as the unique fixed point of a parallel iteration:
//
//   * every mover (occupied && !happy at step start) has a state-independent probe sequence
//     p[0..49] = _randbelow(ncells) draws of its own DRAW stream;
//   * every cell has a 16-byte slot {vac, arr}: the cell is empty at time t (t = priority of
//     the acting mover) iff  vac <= t <= arr, where
//         vac = 0 (empty since the start of the step) | priority+1 of its occupant if the
//               occupant moves this step (it is gone once its turn is over) | NEVER
//         arr = min priority of the movers that currently choose the cell, NEVER if none
//     (a cell receives at most one arrival per step: an arrived agent has already acted);
//     one probe = ONE 16-byte load;
//   * a pass lets every mover take its first probe that is empty at its own time and keep `arr`
//     up to date in place (atomicMin to claim, compare-and-swap to release, re-asserted every
//     pass).  The mover of lowest rank is right after one pass and its claim is never disturbed;
//     by induction rank k is right one pass after ranks < k are, so the passes converge to
//     exactly the sequential result (in practice a handful of passes: conflicts are sparse);
//   * the 50-probe miss (probability density^50) falls back to the k-th empty cell in row-major
//     order at the mover's time, k = _randbelow(#empty), #empty = ncells - nagents at every
//     instant (grid.py:308-316).  All fallback movers of a pass are served by ONE pass over the
//     slot table: their times are sorted, a cell's emptiness interval is a contiguous range of
//     that order (two binary searches -> +1/-1 in a per-block difference array), per-block
//     counts locate the block of the k-th empty cell and one warp scans that block.
//
// All state is caller-owned: the int8 type layer, the uint8 happy layer, the uint32
// agent-id layer (who sits in each cell), optionally the per-agent cell array, and the
// workspace holding the slot table (persistent: initialise it from the type layer once).
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr unsigned long long kNever = ~0ull;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kMaxProbes = 50;   // grid.py:303 `for _ in range(50)`
constexpr int kMaxPasses = 100000;
constexpr int kFbBatch = 2048;   // fallback movers served per table pass
constexpr int kFbBlocksMax = 4096;

struct Slot {
    unsigned long long vac, arr;
};

struct RelocWs {
    Slot* slots;                    // [ncells]
    uint32_t* mv_cell;              // [cap] origin cell of mover m
    uint32_t* mv_agent;             // [cap] agent index
    unsigned long long* mv_prio;    // [cap] priority64
    int8_t* mv_type;                // [cap]
    uint32_t* choice;               // [cap] cell currently chosen (claimed) by mover m
    uint32_t* fb_list;              // [cap] movers that missed all 50 probes in this pass
    unsigned long long* fb_t;       // [kFbBatch] sorted times of the current fallback batch
    uint32_t* fb_m;                 // [kFbBatch] mover of each sorted entry
    unsigned long long* fb_k;       // [kFbBatch] rank of the empty cell to take
    uint32_t* fb_block;             // [kFbBatch] table block holding it
    uint32_t* fb_rem;               // [kFbBatch] rank inside that block
    uint32_t* cnt;                  // [kFbBatch][kFbBlocksMax] empties per table block at each fallback time
    unsigned long long* counters;   // [0] movers [1] occupied [2] fallback movers of the pass [3] changes of the pass
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(RelocWs* ws, char* base, int64_t ncells, int64_t cap) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    RelocWs t;
    t.slots = (Slot*)take(sizeof(Slot) * ncells);
    t.mv_cell = (uint32_t*)take(4 * cap);
    t.mv_agent = (uint32_t*)take(4 * cap);
    t.mv_prio = (unsigned long long*)take(8 * cap);
    t.mv_type = (int8_t*)take(cap);
    t.choice = (uint32_t*)take(4 * cap);
    t.fb_list = (uint32_t*)take(4 * cap);
    t.fb_t = (unsigned long long*)take(8 * kFbBatch);
    t.fb_m = (uint32_t*)take(4 * kFbBatch);
    t.fb_k = (unsigned long long*)take(8 * kFbBatch);
    t.fb_block = (uint32_t*)take(4 * kFbBatch);
    t.fb_rem = (uint32_t*)take(4 * kFbBatch);
    t.cnt = (uint32_t*)take(4ull * kFbBlocksMax * kFbBatch);
    t.counters = (unsigned long long*)take(8 * 8);
    if (ws) *ws = t;
    return off;
}

__device__ __forceinline__ Slot load_slot(const Slot* p) {
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));  // L2: sees other SMs' atomics
    return Slot{v.x, v.y};
}

__device__ __forceinline__ bool empty_at(const Slot& s, unsigned long long t) { return s.vac <= t && t <= s.arr; }

__global__ void init_slots(const int8_t* __restrict__ type, int64_t ncells, Slot* __restrict__ slots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += stride) {
        const bool occupied = type != nullptr && type[c] >= 0;
        reinterpret_cast<ulonglong2*>(slots)[c] = make_ulonglong2(occupied ? kNever : 0ull, kNever);
    }
}

// movers = occupied && !happy; appended in arbitrary order (the result does not depend on it).
__global__ void collect_movers(const int8_t* __restrict__ type, const uint8_t* __restrict__ happy,
                               const uint32_t* __restrict__ agent_id, int64_t ncells, RelocWs ws, int64_t cap,
                               uint64_t seed, uint32_t epoch) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned occupied = 0;
    for (int64_t c0 = (int64_t)blockIdx.x * blockDim.x; c0 < ncells; c0 += stride) {
        const int64_t c = c0 + threadIdx.x;
        const int t = c < ncells ? type[c] : -1;
        const bool mover = t >= 0 && happy[c] == 0;
        occupied += t >= 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, mover);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&ws.counters[0], (unsigned long long)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (mover) {
                const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
                if ((int64_t)slot < cap) {
                    const uint32_t agent = agent_id[c];
                    const unsigned long long pr = mb::priority64(seed, epoch, agent);
                    ws.mv_cell[slot] = (uint32_t)c;
                    ws.mv_agent[slot] = agent;
                    ws.mv_prio[slot] = pr;
                    ws.mv_type[slot] = (int8_t)t;
                    ws.choice[slot] = kNone;
                    ws.slots[c].vac = pr + 1;  // empty for every mover that acts after this one
                }
            }
        }
    }
    occupied = warp_reduce_add(occupied);
    if ((threadIdx.x & 31) == 0 && occupied) atomicAdd(&ws.counters[1], (unsigned long long)occupied);
}

// "the table changed in this pass": a plain store of a constant (millions of movers change in the
// first pass; counting them with one atomic address would serialise the whole pass)
__device__ __forceinline__ void mark_changed(RelocWs& ws) { *(volatile unsigned long long*)&ws.counters[3] = 1ull; }

// claim `pick` for mover m (and release what it held before)
__device__ __forceinline__ void commit_choice(RelocWs& ws, int64_t m, unsigned long long t, uint32_t pick) {
    const uint32_t held = ws.choice[m];
    if (held != pick) {
        if (held != kNone) atomicCAS(&ws.slots[held].arr, t, kNever);  // release only if I am the registered arrival
        ws.choice[m] = pick;
        atomicMin(&ws.slots[pick].arr, t);
        mark_changed(ws);
    } else {
        // re-assert: my registration may have been shadowed by an earlier arrival that has left since
        if (atomicMin(&ws.slots[pick].arr, t) > t) mark_changed(ws);
    }
}

__global__ void __launch_bounds__(128)
relocate_pass(RelocWs ws, int64_t ncells, int64_t nmovers, uint64_t seed, uint32_t epoch) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const unsigned long long t = ws.mv_prio[m];
    mb::DrawStream ds(seed, epoch, ws.mv_agent[m]);
    for (int j = 0; j < kMaxProbes; ++j) {
        const int64_t c = (int64_t)ds.randbelow((uint64_t)ncells);
        if (empty_at(load_slot(ws.slots + c), t)) {
            commit_choice(ws, m, t, (uint32_t)c);
            return;
        }
    }
    ws.fb_list[atomicAdd(&ws.counters[2], 1ull)] = (uint32_t)m;  // served by the fallback kernels of this pass
}

// ---- argwhere fallback, batched -----------------------------------------------------------------
// sort the batch by time (bitonic, one CTA) and draw each mover's rank k = _randbelow(nempty)
__global__ void __launch_bounds__(1024)
fb_sort(RelocWs ws, int64_t first, int count, int64_t ncells, int64_t nempty, uint64_t seed, uint32_t epoch) {
    __shared__ unsigned long long st[kFbBatch];
    __shared__ uint32_t sm[kFbBatch];
    for (int i = threadIdx.x; i < kFbBatch; i += blockDim.x) {
        if (i < count) {
            const uint32_t m = ws.fb_list[first + i];
            st[i] = ws.mv_prio[m];
            sm[i] = m;
        } else {
            st[i] = kNever;
            sm[i] = kNone;
        }
    }
    __syncthreads();
    for (int k = 2; k <= kFbBatch; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < kFbBatch; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    if ((st[i] > st[p]) == up) {
                        const unsigned long long a = st[i]; st[i] = st[p]; st[p] = a;
                        const uint32_t b = sm[i]; sm[i] = sm[p]; sm[p] = b;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const uint32_t m = sm[i];
        ws.fb_t[i] = st[i];
        ws.fb_m[i] = m;
        // replay the 50 accepted probes to position the stream, then draw the rank (grid.py:308-310)
        mb::DrawStream ds(seed, epoch, ws.mv_agent[m]);
        for (int j = 0; j < kMaxProbes; ++j) (void)ds.randbelow((uint64_t)ncells);
        ws.fb_k[i] = ds.randbelow((uint64_t)nempty);
    }
}

__device__ __forceinline__ int first_ge(const unsigned long long* __restrict__ a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// cnt[b][f] = number of cells of table block b that are empty at the time of fallback f
__global__ void __launch_bounds__(256)
fb_count(RelocWs ws, int64_t ncells, int64_t block_cells, int count) {
    __shared__ int diff[kFbBatch + 1];
    __shared__ unsigned long long times[kFbBatch];
    __shared__ int warp_base[8];
    for (int i = threadIdx.x; i <= kFbBatch; i += blockDim.x) diff[i] = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) times[i] = ws.fb_t[i];
    __syncthreads();
    const int64_t lo_c = (int64_t)blockIdx.x * block_cells;
    const int64_t hi_c = min(lo_c + block_cells, ncells);
    int base = 0;
    for (int64_t c = lo_c + threadIdx.x; c < hi_c; c += blockDim.x) {
        const Slot s = load_slot(ws.slots + c);
        if (s.vac == kNever) continue;                       // occupied during the whole step
        if (s.vac == 0 && s.arr == kNever) { ++base; continue; }  // empty during the whole step
        const int lo = first_ge(times, count, s.vac);        // first fallback with t >= vac
        const int hi = s.arr == kNever ? count : first_ge(times, count, s.arr + 1);  // first with t > arr
        if (lo < hi) {
            atomicAdd(&diff[lo], 1);
            atomicSub(&diff[hi], 1);
        }
    }
    base = (int)warp_reduce_add((unsigned)base);
    if ((threadIdx.x & 31) == 0) warp_base[threadIdx.x >> 5] = base;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0;
        for (int w = 0; w < 8; ++w) b += warp_base[w];
        int run = b;
        for (int f = 0; f < count; ++f) {  // prefix sum of the difference array (count <= 2048)
            run += diff[f];
            diff[f] = run;
        }
    }
    __syncthreads();
    for (int f = threadIdx.x; f < count; f += blockDim.x) ws.cnt[(size_t)f * kFbBlocksMax + blockIdx.x] = (uint32_t)diff[f];
}

// one warp per fallback mover: walk the per-block counts 32 blocks at a time (coalesced, warp scan)
__global__ void __launch_bounds__(256) fb_locate(RelocWs ws, int nblocks, int count) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= count) return;
    unsigned long long k = ws.fb_k[f];
    const uint32_t* __restrict__ row = ws.cnt + (size_t)f * kFbBlocksMax;
    int block = nblocks - 1;
    unsigned rem = 0;
    bool done = false;
    for (int b0 = 0; b0 < nblocks && !done; b0 += 32) {
        const int b = b0 + lane;
        const unsigned n = b < nblocks ? row[b] : 0u;
        unsigned incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        if (k < total) {
            // first lane whose inclusive count exceeds k holds the block
            const unsigned hit = __ballot_sync(0xffffffffu, k < incl);
            const int l = __ffs(hit) - 1;
            const unsigned before = __shfl_sync(0xffffffffu, incl - n, l);
            block = b0 + l;
            rem = (unsigned)(k - before);
            done = true;
        } else {
            k -= total;
        }
    }
    if (lane == 0) {
        ws.fb_block[f] = (uint32_t)block;
        ws.fb_rem[f] = done ? rem : 0xffffffffu;  // not found: fb_pick reports "changed" and the pass repeats
    }
}

// one warp per fallback mover: the fb_rem-th empty cell of block fb_block at the mover's time
__global__ void __launch_bounds__(256)
fb_pick(RelocWs ws, int64_t ncells, int64_t block_cells, int count) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= count) return;
    const unsigned long long t = ws.fb_t[f];
    const int64_t lo_c = (int64_t)ws.fb_block[f] * block_cells;
    const int64_t hi_c = min(lo_c + block_cells, ncells);
    unsigned need = ws.fb_rem[f];
    int64_t found = -1;
    for (int64_t c0 = lo_c; c0 < hi_c; c0 += 32) {
        const int64_t c = c0 + lane;
        const bool e = c < hi_c && empty_at(load_slot(ws.slots + c), t);
        const unsigned ballot = __ballot_sync(0xffffffffu, e);
        const unsigned n = __popc(ballot);
        if (need < n) {
            unsigned b = ballot;
            for (unsigned i = 0; i < need; ++i) b &= b - 1;  // drop the `need` lowest set bits
            found = c0 + (__ffs(b) - 1);
            break;
        }
        need -= n;
    }
    if (lane == 0) {
        if (found >= 0) commit_choice(ws, ws.fb_m[f], t, (uint32_t)found);
        else mark_changed(ws);  // table moved under the scan (concurrent commits): force another pass
    }
}

__global__ void relocate_vacate(RelocWs ws, int8_t* type, uint8_t* happy, int64_t nmovers) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const uint32_t o = ws.mv_cell[m];
    type[o] = -1;
    happy[o] = 0;
    reinterpret_cast<ulonglong2*>(ws.slots)[o] = make_ulonglong2(0ull, kNever);
}

__global__ void relocate_arrive(RelocWs ws, int8_t* type, uint8_t* happy, uint32_t* agent_id,
                                uint32_t* agent_cell, int64_t nmovers) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const uint32_t d = ws.choice[m];
    const uint32_t agent = ws.mv_agent[m];
    type[d] = ws.mv_type[m];
    happy[d] = 0;  // the flag travels with the agent: movers are unhappy until the next assign_state
    agent_id[d] = agent;
    if (agent_cell != nullptr) agent_cell[agent] = d;
    reinterpret_cast<ulonglong2*>(ws.slots)[d] = make_ulonglong2(kNever, kNever);
}

int read_counters(const RelocWs& ws, unsigned long long* host4, cudaStream_t stream) {
    int rc = mb::check_cuda(cudaMemcpyAsync(host4, ws.counters, 4 * sizeof(unsigned long long),
                                            cudaMemcpyDeviceToHost, stream), "read counters");
    if (rc) return rc;
    return mb::check_cuda(cudaStreamSynchronize(stream), "sync");
}

}  // namespace

MB_API int mb_relocate_workspace_bytes(int64_t ncells, int64_t max_movers, size_t* out) {
    MB_REQUIRE(out != nullptr && ncells >= 0 && max_movers >= 0, "mb_relocate_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, ncells, max_movers);
    return MB_OK;
}

// Builds the slot table from the occupancy (type[c] >= 0 = occupied; type == NULL = all empty).
// Call once, and again whenever the type layer was modified outside mb_schelling_relocate.
MB_API int mb_relocate_workspace_init(void* workspace, size_t workspace_bytes, int64_t ncells,
                                      int64_t max_movers, const int8_t* type, cudaStream_t stream) {
    RelocWs ws;
    MB_REQUIRE(workspace != nullptr, "mb_relocate_workspace_init: null workspace");
    if (carve(&ws, (char*)workspace, ncells, max_movers) > workspace_bytes) {
        mb::set_error("mb_relocate_workspace_init: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    if (ncells > 0) init_slots<<<mb::sm_count() * 8, 256, 0, stream>>>(type, ncells, ws.slots);
    MB_LAUNCH_CHECK("mb_relocate_workspace_init");
    return MB_OK;
}

// Synchronises `stream` (needs the mover count and the per-pass convergence flags on the host).
MB_API int mb_schelling_relocate(int8_t* type, uint8_t* happy, uint32_t* agent_id, uint32_t* agent_cell,
                                 int64_t ncells, uint64_t seed, uint32_t epoch, void* workspace,
                                 size_t workspace_bytes, int64_t max_movers, int64_t* out_movers,
                                 int* out_iterations, cudaStream_t stream) {
    MB_REQUIRE(ncells >= 0 && ncells < 0xffffffffll, "mb_schelling_relocate: ncells must be < 2^32 - 1");
    if (out_movers) *out_movers = 0;
    if (out_iterations) *out_iterations = 0;
    if (ncells == 0) return MB_OK;
    MB_REQUIRE(type && happy && agent_id && workspace, "mb_schelling_relocate: null buffer");
    RelocWs ws;
    if (carve(&ws, (char*)workspace, ncells, max_movers) > workspace_bytes) {
        mb::set_error("mb_schelling_relocate: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    int rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 8 * sizeof(unsigned long long), stream), "memset");
    if (rc) return rc;
    const int sms = mb::sm_count();
    {
        const int64_t want = mb::ceil_div(ncells, 256);
        const unsigned blocks = (unsigned)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
        collect_movers<<<blocks, 256, 0, stream>>>(type, happy, agent_id, ncells, ws, max_movers, seed, epoch);
        MB_LAUNCH_CHECK("collect_movers");
    }
    unsigned long long hc[4];
    rc = read_counters(ws, hc, stream);
    if (rc) return rc;
    const int64_t nmovers = (int64_t)hc[0], noccupied = (int64_t)hc[1];
    if (out_movers) *out_movers = nmovers;
    if (nmovers == 0) return MB_OK;
    if (nmovers > max_movers) {
        mb::set_error("mb_schelling_relocate: %lld movers exceed workspace capacity %lld", (long long)nmovers,
                      (long long)max_movers);
        return MB_ERR_WORKSPACE;
    }
    const int64_t nempty = ncells - noccupied;
    // grid.py:311-315: a completely full grid raises ValueError at the first unhappy agent
    MB_REQUIRE(nempty > 0, "Grid is completely full. No empty cells available. Cannot select a random empty cell.");

    const unsigned mblocks = (unsigned)mb::ceil_div(nmovers, 128);
    int64_t fb_blocks = mb::ceil_div(ncells, 1024);
    if (fb_blocks > kFbBlocksMax) fb_blocks = kFbBlocksMax;
    const int64_t block_cells = mb::ceil_div(ncells, fb_blocks);
    int pass = 0;
    for (;; ++pass) {
        if (pass >= kMaxPasses) {
            mb::set_error("mb_schelling_relocate: no convergence after %d passes", pass);
            return MB_ERR_CUDA;
        }
        rc = mb::check_cuda(cudaMemsetAsync(ws.counters + 2, 0, 2 * sizeof(unsigned long long), stream), "memset");
        if (rc) return rc;
        relocate_pass<<<mblocks, 128, 0, stream>>>(ws, ncells, nmovers, seed, epoch);
        MB_LAUNCH_CHECK("relocate_pass");
        rc = read_counters(ws, hc, stream);
        if (rc) return rc;
        const int64_t nfb = (int64_t)hc[2];
        if (nfb > 0) {
            for (int64_t first = 0; first < nfb; first += kFbBatch) {
                const int count = (int)(nfb - first < kFbBatch ? nfb - first : kFbBatch);
                fb_sort<<<1, 1024, 0, stream>>>(ws, first, count, ncells, nempty, seed, epoch);
                fb_count<<<(unsigned)fb_blocks, 256, 0, stream>>>(ws, ncells, block_cells, count);
                fb_locate<<<(unsigned)mb::ceil_div((int64_t)count * 32, 256), 256, 0, stream>>>(ws, (int)fb_blocks, count);
                fb_pick<<<(unsigned)mb::ceil_div((int64_t)count * 32, 256), 256, 0, stream>>>(ws, ncells, block_cells, count);
            }
            MB_LAUNCH_CHECK("relocate_fallback");
            rc = read_counters(ws, hc, stream);
            if (rc) return rc;
        }
        if (hc[3] == 0) break;  // a pass without any change of the table: fixed point
    }
    relocate_vacate<<<mblocks, 128, 0, stream>>>(ws, type, happy, nmovers);
    relocate_arrive<<<mblocks, 128, 0, stream>>>(ws, type, happy, agent_id, agent_cell, nmovers);
    MB_LAUNCH_CHECK("relocate_apply");
    if (out_iterations) *out_iterations = pass + 1;
    return MB_OK;
}
