// Boltzmann wealth model: Moore random walk + same-cell exchange, order-exact (K9 of SURVEY.md §2.4).
//
// Replaces AgentSet.shuffle_do("step") (mesa/agentset.py:772-787) over
//   mesa/examples/basic/boltzmann_wealth_model/agents.py:24-26  move:
//        self.cell = self.cell.neighborhood.select_random_cell()
//        (cell_collection.py:101-119: random.choice over the connection-order cell list,
//         grid.py:447-451 Moore offsets filtered by bounds on a clamped grid)
//   agents.py:28-35  give_money: cellmates = cell.agents minus self (list = arrival order,
//        cell.py:143-170); if any: other = random.choice(cellmates); other.wealth += 1; self.wealth -= 1
//   agents.py:37-44  step: move(); if self.wealth > 0: give_money()
//
// The reference is sequential.  The same result is obtained in parallel because
//   * every agent moves exactly once per step and its destination depends only on its own cell
//     and its own keyed draws (philox.cuh), so all moves are independent;
//   * at the time t_a of agent a, its new cell c holds, in list order, the agents whose OLD cell is c
//     and act later (t_b > t_a; they have not left yet; kept in last step's arrival order) followed
//     by the agents whose NEW cell is c and acted earlier (t_b < t_a; in arrival = priority order).
//     Both groups are contiguous runs of two sorted arrays: last step's agents sorted by
//     (cell, arrival priority) and this step's agents sorted by (new cell, priority);
//   * whom a would give to is therefore state-independent; WHETHER it gives depends on
//     wealth(a) at t_a = w0 + gifts received earlier, resolved by a monotone fixed point
//     (an agent with w0 == 0 gives iff some earlier giver targets it).
#include "common.cuh"
#include "philox.cuh"
#include "sort.cuh"

namespace {

constexpr unsigned long long kNever = ~0ull;
constexpr uint32_t kNoTarget = 0xffffffffu;
constexpr int kCounterWords = 2 + 2 * 4096 + 3;   // status, queue length, (start, length) x kLongCap, candidate count, 4 sync words

struct BoltzWs {
    uint64_t* keysA;      // [n] sorted (cell << 32 | arrival priority hi32) of the PREVIOUS step (persistent)
    uint32_t* valsA;      // [n] agent at each sorted slot (persistent)
    uint64_t* keys0;      // [n] sort ping
    uint32_t* vals0;
    uint64_t* keys1;      // [n] sort pong
    uint32_t* vals1;
    uint64_t* prio;       // [n] priority64 of this epoch
    uint32_t* new_cell;   // [n]
    uint32_t* rankB;      // [n] slot of agent in this step's sorted order
    uint32_t* target;     // [n] agent that would receive the gift, or kNoTarget
    unsigned long long* first_gift;  // [n] earliest time a gift reaches the agent
    uint8_t* gives;       // [n]
    uint8_t* draws;       // [n] draws consumed by move()
    unsigned long long* counters;    // [kCounterWords] status | crowded-cell queue | barrier / convergence words
    uint32_t* cand;       // [n] agents that have a cell mate to give to (the only ones the give decisions involve)
    uint32_t* occA;       // [occ_words] hashed bitmap of the cells that hold an agent at the start of the step
    uint32_t occ_mask;    // occ_words * 32 - 1 (a power of two)
    void* sort_ws;
    size_t sort_ws_bytes;
};

size_t carve(BoltzWs* ws, char* base, int64_t n) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    BoltzWs t;
    t.keysA = (uint64_t*)take(8 * n);
    t.valsA = (uint32_t*)take(4 * n);
    t.keys0 = (uint64_t*)take(8 * n);
    t.vals0 = (uint32_t*)take(4 * n);
    t.keys1 = (uint64_t*)take(8 * n);
    t.vals1 = (uint32_t*)take(4 * n);
    t.prio = (uint64_t*)take(8 * n);
    t.new_cell = (uint32_t*)take(4 * n);
    t.rankB = (uint32_t*)take(4 * n);
    t.target = (uint32_t*)take(4 * n);
    t.first_gift = (unsigned long long*)take(8 * n);
    t.gives = (uint8_t*)take(n);
    t.draws = (uint8_t*)take(n);
    t.counters = (unsigned long long*)take(8 * kCounterWords);
    t.cand = (uint32_t*)take(4 * n);
    {   // >= 32 bits per agent, a power of two: a false positive only costs the binary search it would have skipped
        uint64_t bits = 1024;
        while (bits < 32ull * (uint64_t)(n > 0 ? n : 1) && bits < (1ull << 31)) bits <<= 1;
        t.occ_mask = (uint32_t)(bits - 1);
        t.occA = (uint32_t*)take((size_t)(bits / 8));
    }
    t.sort_ws_bytes = mb::sort_workspace_bytes(n);
    t.sort_ws = take(t.sort_ws_bytes);
    if (ws) *ws = t;
    return off;
}

__global__ void init_keys(const uint32_t* __restrict__ cell, int64_t n, uint64_t* __restrict__ keys,
                          uint32_t* __restrict__ vals) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    keys[a] = (uint64_t)cell[a] << 32;  // initial list order = registration order (stable sort)
    vals[a] = (uint32_t)a;
}

// move(): destination = connection-order neighbour list [k], k = _randbelow(len)
__global__ void boltz_move(const uint32_t* __restrict__ cell, int64_t n, int rows, int cols, int torus,
                           uint64_t seed, uint32_t epoch, BoltzWs ws) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint32_t c = cell[a];
    const int i = (int)(c / (uint32_t)cols), j = (int)(c % (uint32_t)cols);
    {   // the cell holds somebody at the start of the step (boltz_target skips the search of last step's order otherwise)
        const uint32_t h = (c * 0x9E3779B1u) & ws.occ_mask;
        atomicOr(&ws.occA[h >> 5], 1u << (h & 31u));
    }
    // grid.py:447-451 offsets in order; clamped grids drop out-of-range cells (grid.py:398): count the existing
    // neighbours, draw k = _randbelow(count), take the k-th existing one
    int cnt = 8;
    if (!torus) {
        const int ni = (i > 0) + 1 + (i < rows - 1), nj = (j > 0) + 1 + (j < cols - 1);
        cnt = ni * nj - 1;
    }
    mb::DrawStream ds(seed, epoch, (uint64_t)a);
    int k = (int)ds.randbelow((uint64_t)cnt);
    uint32_t dest = c;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        if (q == 4) continue;
        int ni = i + q / 3 - 1, nj = j + q % 3 - 1;
        bool exists = true;
        if (torus) {
            ni = (ni + rows) % rows;
            nj = (nj + cols) % cols;
        } else {
            exists = ni >= 0 && ni < rows && nj >= 0 && nj < cols;
        }
        if (exists && k-- == 0) dest = (uint32_t)ni * (uint32_t)cols + (uint32_t)nj;
    }
    const uint64_t t = mb::priority64(seed, epoch, (uint64_t)a);
    ws.new_cell[a] = dest;
    ws.prio[a] = t;
    ws.draws[a] = (uint8_t)ds.c;
    ws.keys0[a] = ((uint64_t)dest << 32) | (t >> 32);
    ws.vals0[a] = (uint32_t)a;
}

__global__ void boltz_rank(const uint32_t* __restrict__ valsB, int64_t n, uint32_t* __restrict__ rankB) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rankB[valsB[i]] = (uint32_t)i;
}

// Sparse grids: the pairs are radix-sorted on the CELL bits only (3 passes instead of 7 for 4096^2); a run of equal
// cells is short (mean occupancy 0.06 at BASELINE config 3) and arrives in agent order, so ordering it by the full key
// (cell, priority high word; ties keep ascending agent index = the tie-break of priority64) is an insertion sort by
// the thread that owns the run's first slot.  Runs longer than kMaxFixRun raise a flag and the step falls back to the
// full-width sort.
constexpr int kMaxFixRun = 32;
constexpr int kLongCap = 4096;       // crowded cells queued per step (more: error status)
constexpr int kLongSmem = 4096;      // run length sorted in shared memory by one CTA (longer: in place, global memory)

// longs: [0] number of queued runs, then (start, length) pairs
__global__ void boltz_fix_runs(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int64_t n,
                               unsigned long long* __restrict__ longs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t cell = keys[i] >> 32;
    if (i > 0 && (keys[i - 1] >> 32) == cell) return;      // not the start of a run
    if (i + 1 >= n || (keys[i + 1] >> 32) != cell) return;  // a run of one
    int64_t e = i + 2;
    while (e < n && (keys[e] >> 32) == cell) ++e;
    if (e - i > kMaxFixRun) {
        // a crowded cell: queued for the cooperative sort (boltz_fix_long) -- no host round trip
        const unsigned long long slot = atomicAdd(&longs[0], 1ull);
        if (slot < (unsigned long long)kLongCap) {
            longs[1 + 2 * slot] = (unsigned long long)i;
            longs[2 + 2 * slot] = (unsigned long long)(e - i);
        }
        return;
    }
    for (int64_t a = i + 1; a < e; ++a) {
        const uint64_t k = keys[a];
        const uint32_t v = vals[a];
        int64_t b = a;
        while (b > i && keys[b - 1] > k) {   // strict: equal keys keep their (ascending agent index) order
            keys[b] = keys[b - 1];
            vals[b] = vals[b - 1];
            --b;
        }
        keys[b] = k;
        vals[b] = v;
    }
}

// Crowded cells (runs longer than kMaxFixRun): one CTA per queued run, bitonic sort on (key, agent) -- the radix sort
// on the cell bits is stable, so a run arrives in ascending agent index and ordering ties by the agent index equals the
// stable order.  Runs of up to kLongSmem pairs are sorted in shared memory, longer ones in place.
__global__ void __launch_bounds__(1024)
boltz_fix_long(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, const unsigned long long* __restrict__ longs,
               unsigned long long* __restrict__ status) {
    __shared__ uint64_t sk[kLongSmem];
    __shared__ uint32_t sv[kLongSmem];
    unsigned long long nruns = longs[0];
    if (nruns > (unsigned long long)kLongCap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *status = 1ull;   // more crowded cells than the queue holds
        nruns = kLongCap;
    }
    for (unsigned long long r = blockIdx.x; r < nruns; r += gridDim.x) {
        const int64_t start = (int64_t)longs[1 + 2 * r];
        const int64_t len = (int64_t)longs[2 + 2 * r];
        int64_t pow2 = 1;
        while (pow2 < len) pow2 <<= 1;
        const bool in_smem = len <= kLongSmem;
        uint64_t* K = in_smem ? sk : keys + start;
        uint32_t* V = in_smem ? sv : vals + start;
        if (in_smem) {
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                sk[i] = keys[start + i];
                sv[i] = vals[start + i];
            }
        }
        __syncthreads();
        // bitonic network in its all-ascending form (first stage of every merge pairs i with i ^ (k - 1), the others
        // i with i ^ j): the smaller element always goes to the lower index, so the virtual +infinity elements that pad
        // the run to a power of two never move and pairs with a partner beyond `len` are simply skipped
        for (int64_t k = 2; k <= pow2; k <<= 1)
            for (int64_t j = k >> 1; j > 0; j >>= 1) {
                const int64_t mask = j == (k >> 1) ? k - 1 : j;
                for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                    const int64_t p = i ^ mask;
                    if (p > i && p < len) {
                        const uint64_t ki = K[i], kp = K[p];
                        const uint32_t vi = V[i], vp = V[p];
                        if (ki > kp || (ki == kp && vi > vp)) {
                            K[i] = kp; V[i] = vp;
                            K[p] = ki; V[p] = vi;
                        }
                    }
                }
                __syncthreads();
            }
        if (in_smem) {
            for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
                keys[start + i] = sk[i];
                vals[start + i] = sv[i];
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* __restrict__ a, int64_t n, uint64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// give_money() target of every agent (state independent) + initial give decision
__global__ void boltz_target(const int32_t* __restrict__ wealth, int64_t n, uint64_t seed, uint32_t epoch,
                             const uint64_t* __restrict__ keysB, const uint32_t* __restrict__ valsB, BoltzWs ws,
                             unsigned long long* __restrict__ cand_count) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint32_t c = ws.new_cell[a];
    const uint64_t t = ws.prio[a];
    // agents still sitting in c (old cell == c, acting later), in last step's arrival order -- nobody if the cell was
    // empty at the start of the step (94 % of the cells at BASELINE config 3: one bit instead of a 20-step search)
    const uint32_t hc = (c * 0x9E3779B1u) & ws.occ_mask;
    int64_t a0 = 0, cntA = 0;
    if ((ws.occA[hc >> 5] >> (hc & 31u)) & 1u) {
        a0 = lower_bound_u64(ws.keysA, n, (uint64_t)c << 32);
        for (int64_t s = a0; s < n && (uint32_t)(ws.keysA[s] >> 32) == c; ++s) cntA += ws.prio[ws.valsA[s]] > t;
    }
    // agents that already arrived in c, in arrival order: the run of this step's order before me
    const int64_t me = ws.rankB[a];
    int64_t b0 = me;
    while (b0 > 0 && (uint32_t)(keysB[b0 - 1] >> 32) == c) --b0;
    const int64_t cntB = me - b0;
    const int64_t mates = cntA + cntB;
    uint32_t tgt = kNoTarget;
    if (mates > 0) {
        mb::DrawStream ds(seed, epoch, (uint64_t)a);
        ds.seek(ws.draws[a]);
        int64_t m = (int64_t)ds.randbelow((uint64_t)mates);
        if (m < cntA) {
            for (int64_t s = a0;; ++s) {
                const uint32_t b = ws.valsA[s];
                if (ws.prio[b] > t && m-- == 0) { tgt = b; break; }
            }
        } else {
            tgt = valsB[b0 + (m - cntA)];
        }
    }
    ws.target[a] = tgt;
    ws.gives[a] = (tgt != kNoTarget && wealth[a] > 0) ? 1 : 0;
    ws.first_gift[a] = kNever;
    if (tgt != kNoTarget) {   // warp-aggregated append to the candidate list
        const unsigned peers = __activemask();
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(cand_count, (unsigned long long)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        ws.cand[base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)a;
    }
}

// The give decisions as ONE persistent kernel: a giver announces the time of its gift to the receiver (atomicMin), an
// agent that started the step broke gives iff a gift reached it before its turn, repeated until nothing changes --
// a monotone fixed point, ~4 rounds.  The rounds are separated by a grid-wide barrier (all CTAs are resident: the
// grid is sized from the occupancy), so the convergence test never goes to the host.
//   sync[0]: barrier arrivals (monotone), sync[1 + (round & 1)]: "somebody changed in this round", sync[3]: rounds done
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < target) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
boltz_iterate(BoltzWs ws, const unsigned long long* __restrict__ cand_count, unsigned* sync, uint32_t* iterations_out) {
    const int64_t n = (int64_t)*cand_count;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned target = 0, round = 0;
    for (;; ++round) {
        // push: givers that have not announced yet (gives == 1 -> 2)
        for (int64_t i = first; i < n; i += stride) {
            const uint32_t a = ws.cand[i];
            if (ws.gives[a] == 1) {
                atomicMin(&ws.first_gift[ws.target[a]], (unsigned long long)ws.prio[a]);
                ws.gives[a] = 2;
            }
        }
        grid_barrier(&sync[0], target);
        unsigned changed = 0;
        for (int64_t i = first; i < n; i += stride) {
            const uint32_t a = ws.cand[i];
            if (ws.gives[a] == 0 && __ldcg(ws.first_gift + a) < ws.prio[a]) {
                ws.gives[a] = 1;
                changed = 1;
            }
        }
        if (__syncthreads_or(changed) && threadIdx.x == 0) *(volatile unsigned*)&sync[1 + (round & 1u)] = 1u;
        if (blockIdx.x == 0 && threadIdx.x == 0) sync[1 + ((round + 1u) & 1u)] = 0u;   // the other round's flag: everybody has read it
        grid_barrier(&sync[0], target);
        if (*(volatile unsigned*)&sync[1 + (round & 1u)] == 0u) break;
        if (round > 1000000u) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && iterations_out != nullptr) *iterations_out = round + 1u;
}

// + the bookkeeping that used to be five more stream operations per step: this step's arrival order becomes next step's
// "still sitting here" order (keysA / valsA), and the words the NEXT step expects cleared (crowded-cell queue length,
// candidate count, barrier words, the occupancy filter) are cleared here, after their last reader
__global__ void boltz_apply(int32_t* __restrict__ wealth, uint32_t* __restrict__ cell, int64_t n, BoltzWs ws,
                            const uint64_t* __restrict__ keysB, const uint32_t* __restrict__ valsB) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a == 0) {
        ws.counters[1] = 0ull;
        ws.counters[kCounterWords - 3] = 0ull; ws.counters[kCounterWords - 2] = 0ull; ws.counters[kCounterWords - 1] = 0ull;
    }
    const int64_t total = (int64_t)gridDim.x * blockDim.x;
    const int64_t occ_words = ((int64_t)ws.occ_mask + 1) / 32;
    for (int64_t i = a; i < occ_words; i += total) ws.occA[i] = 0u;
    if (a >= n) return;
    ws.keysA[a] = keysB[a];
    ws.valsA[a] = valsB[a];
    cell[a] = ws.new_cell[a];
    if (ws.gives[a]) {
        atomicAdd(&wealth[ws.target[a]], 1);
        atomicSub(&wealth[a], 1);
    }
    // (gives is 0 / 2 here: every giver has announced its gift before the last round found nothing to change)
}

int sort_bits(int64_t ncells) {
    int b = 1;
    while ((1ll << b) < ncells) ++b;
    return 32 + b;
}

}  // namespace

MB_API int mb_boltzmann_workspace_bytes(int64_t n, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0, "mb_boltzmann_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, n);
    return MB_OK;
}

// (Re)builds the per-cell arrival order from the current cells: list order = agent index
// (registration order, boltzmann_wealth_model/model.py:70-74).  Call once after placing agents.
MB_API int mb_boltzmann_init(const uint32_t* cell, int64_t n, int64_t rows, int64_t cols, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_boltzmann_init: n out of range");
    MB_REQUIRE(rows > 0 && cols > 0 && rows * cols < (1ll << 32), "mb_boltzmann_init: grid must have < 2^32 cells");
    if (n == 0) return MB_OK;
    MB_REQUIRE(cell && workspace, "mb_boltzmann_init: null buffer");
    BoltzWs ws;
    if (carve(&ws, (char*)workspace, n) > workspace_bytes) {
        mb::set_error("mb_boltzmann_init: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    int rc0 = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 8 * kCounterWords, stream), "memset");   // status word, queue
    if (rc0) return rc0;
    if ((rc0 = mb::check_cuda(cudaMemsetAsync(ws.occA, 0, ((size_t)ws.occ_mask + 1) / 8, stream), "memset"))) return rc0;
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    init_keys<<<blocks, 256, 0, stream>>>(cell, n, ws.keys0, ws.vals0);
    const int where = mb::radix_sort_pairs<uint64_t>(ws.keys0, ws.vals0, ws.keys1, ws.vals1, n, 32, sort_bits(rows * cols),
                                                     ws.sort_ws, ws.sort_ws_bytes, stream);
    if (where < 0) return where;
    int rc = mb::check_cuda(cudaMemcpyAsync(ws.keysA, where ? ws.keys1 : ws.keys0, 8 * n, cudaMemcpyDeviceToDevice, stream), "copy");
    if (rc) return rc;
    return mb::check_cuda(cudaMemcpyAsync(ws.valsA, where ? ws.vals1 : ws.vals0, 4 * n, cudaMemcpyDeviceToDevice, stream), "copy");
}

// One model step (shuffle number `epoch`).  Enqueues ~20 launches on `stream` and returns: no host synchronisation
// (crowded cells and the convergence of the give decisions are handled on the device).  iterations_dev (optional,
// device uint32): rounds of the give-decision fixed point.  status word (device, ws.counters[0], read it with
// mb_boltzmann_status): non-zero if more than 4096 cells held more than 32 agents in this step (not supported).
MB_API int mb_boltzmann_step(uint32_t* cell, int32_t* wealth, int64_t n, int64_t rows, int64_t cols, int torus,
                             uint64_t seed, uint32_t epoch, void* workspace, size_t workspace_bytes,
                             uint32_t* iterations_dev, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_boltzmann_step: n out of range");
    MB_REQUIRE(rows > 0 && cols > 0 && rows * cols < (1ll << 32), "mb_boltzmann_step: grid must have < 2^32 cells");
    // a 1x1 grid has an empty neighbourhood: select_random_cell raises LookupError (cell_collection.py:114-117)
    MB_REQUIRE(rows * cols >= 2, "Cannot select random cell from empty collection");
    MB_REQUIRE(!torus || (rows >= 3 && cols >= 3), "mb_boltzmann_step: torus needs dims >= 3 (SURVEY A.2)");
    if (n == 0) return MB_OK;
    MB_REQUIRE(cell && wealth && workspace, "mb_boltzmann_step: null buffer");
    BoltzWs ws;
    if (carve(&ws, (char*)workspace, n) > workspace_bytes) {
        mb::set_error("mb_boltzmann_step: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    // counters: [0] status (sticky), [1] queued crowded cells, [2 ..] their (start, length) pairs; sync words at the end
    unsigned long long* status = ws.counters;
    unsigned long long* longs = ws.counters + 1;
    unsigned* sync = reinterpret_cast<unsigned*>(ws.counters + kCounterWords - 2);
    unsigned long long* cand_count = ws.counters + kCounterWords - 3;
    // (the queue length, the candidate count, the barrier words and the occupancy filter are clear: mb_boltzmann_init /
    // the previous step's boltz_apply)
    int rc = MB_OK;
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    boltz_move<<<blocks, 256, 0, stream>>>(cell, n, (int)rows, (int)cols, torus, seed, epoch, ws);
    // sparse grids sort on the cell bits only and order the (short) runs of equal cells afterwards; crowded grids
    // sort on the full (cell, priority) key
    int where = 0;
    if (n <= rows * cols) {
        where = mb::radix_sort_pairs<uint64_t>(ws.keys0, ws.vals0, ws.keys1, ws.vals1, n, 32, sort_bits(rows * cols),
                                               ws.sort_ws, ws.sort_ws_bytes, stream);
        if (where < 0) return where;
        uint64_t* k = where ? ws.keys1 : ws.keys0;
        uint32_t* v = where ? ws.vals1 : ws.vals0;
        boltz_fix_runs<<<blocks, 256, 0, stream>>>(k, v, n, longs);
        boltz_fix_long<<<32, 1024, 0, stream>>>(k, v, longs, status);
    } else {
        where = mb::radix_sort_pairs<uint64_t>(ws.keys0, ws.vals0, ws.keys1, ws.vals1, n, 0, sort_bits(rows * cols),
                                               ws.sort_ws, ws.sort_ws_bytes, stream);
        if (where < 0) return where;
    }
    const uint64_t* keysB = where ? ws.keys1 : ws.keys0;
    const uint32_t* valsB = where ? ws.vals1 : ws.vals0;
    boltz_rank<<<blocks, 256, 0, stream>>>(valsB, n, ws.rankB);
    boltz_target<<<blocks, 256, 0, stream>>>(wealth, n, seed, epoch, keysB, valsB, ws, cand_count);
    MB_LAUNCH_CHECK("boltz_target");
    // persistent give-decision kernel over the candidate list: one CTA per SM at most, all resident (grid barrier)
    int64_t grid = mb::ceil_div(n, 1024 * 4);
    if (grid > mb::sm_count()) grid = mb::sm_count();
    boltz_iterate<<<(unsigned)grid, 1024, 0, stream>>>(ws, cand_count, sync, iterations_dev);
    boltz_apply<<<blocks, 256, 0, stream>>>(wealth, cell, n, ws, keysB, valsB);
    (void)rc;
    MB_LAUNCH_CHECK("mb_boltzmann_step");
    return MB_OK;
}

// Sticky status word of the workspace (0 = every step so far was exact).  Synchronises `stream`.
MB_API int mb_boltzmann_status(void* workspace, size_t workspace_bytes, int64_t n, uint64_t* status_host, cudaStream_t stream) {
    MB_REQUIRE(workspace && status_host && n >= 0, "mb_boltzmann_status: bad argument");
    *status_host = 0;
    if (n == 0) return MB_OK;
    BoltzWs ws;
    if (carve(&ws, (char*)workspace, n) > workspace_bytes) {
        mb::set_error("mb_boltzmann_status: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    int rc = mb::check_cuda(cudaMemcpyAsync(status_host, ws.counters, 8, cudaMemcpyDeviceToHost, stream), "read");
    if (rc) return rc;
    return mb::check_cuda(cudaStreamSynchronize(stream), "sync");
}
