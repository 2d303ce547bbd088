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
//   * a cell c is empty at time t (the priority of the acting mover) iff
//         vacate(c) < t  and  not (arrive(c) < t)
//     where vacate(c) = 0 for an initially empty cell, the occupant's priority if the
//     occupant is a mover (it has left once its turn is over), +inf otherwise, and
//     arrive(c) = min priority of the movers that choose c (+inf if none).  A cell receives
//     at most one arrival per step (the arriving agent has already been activated);
//   * iteration k recomputes every mover's choice (first probe that is empty at its time)
//     from the arrive table of iteration k-1 and rebuilds the table with atomicMin.
//     By induction on the rank, after k iterations the first k movers are final, so the
//     iteration converges to exactly the sequential result; in practice it takes O(log M)
//     rounds because conflicts are sparse.
//   * the 50-probe miss (probability density^50) falls back to the k-th empty cell in
//     row-major order at the mover's time, k = _randbelow(#empty), #empty = ncells - nagents
//     at every instant (grid.py:308-316).
//
// All state is caller-owned: the int8 type layer, the uint8 happy layer, the uint32
// agent-id layer (who sits in each cell) and optionally the per-agent cell array.
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr unsigned long long kNever = ~0ull;
constexpr uint32_t kNone = 0xffffffffu;
constexpr int kMaxProbes = 50;  // grid.py:303 `for _ in range(50)`
constexpr int kIterBatch = 4;   // iterations enqueued between convergence checks
constexpr int kMaxIterSlots = 1024;

struct RelocWs {
    unsigned long long* arrive[3];  // [ncells] each, all kNever between calls
    uint32_t* mv_cell;              // [cap] origin cell of mover m
    uint32_t* mv_agent;             // [cap] agent index of mover m
    int8_t* mv_type;                // [cap]
    uint32_t* choice[3];            // [cap] destination chosen in iteration k % 3
    uint32_t* fb_list;              // [cap] movers that missed all 50 probes this iteration
    unsigned long long* counters;   // [0] movers, [1] occupied, [2] fallback count, [3] unused,
                                    // [8 + k] choices changed in iteration k
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(RelocWs* ws, char* base, int64_t ncells, int64_t cap) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align256(bytes);
        return p;
    };
    RelocWs tmp;
    for (int i = 0; i < 3; ++i) tmp.arrive[i] = (unsigned long long*)take(sizeof(unsigned long long) * ncells);
    tmp.mv_cell = (uint32_t*)take(sizeof(uint32_t) * cap);
    tmp.mv_agent = (uint32_t*)take(sizeof(uint32_t) * cap);
    tmp.mv_type = (int8_t*)take(cap);
    for (int i = 0; i < 3; ++i) tmp.choice[i] = (uint32_t*)take(sizeof(uint32_t) * cap);
    tmp.fb_list = (uint32_t*)take(sizeof(uint32_t) * cap);
    tmp.counters = (unsigned long long*)take(sizeof(unsigned long long) * (8 + kMaxIterSlots));
    if (ws) *ws = tmp;
    return off;
}

__global__ void fill_u64(unsigned long long* p, int64_t n, unsigned long long v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// movers = occupied && !happy; appended in arbitrary order (the result does not depend on it).
__global__ void collect_movers(const int8_t* __restrict__ type, const uint8_t* __restrict__ happy,
                               const uint32_t* __restrict__ agent_id, int64_t ncells, RelocWs ws, int64_t cap) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long occupied = 0;
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
                    ws.mv_cell[slot] = (uint32_t)c;
                    ws.mv_agent[slot] = agent_id[c];
                    ws.mv_type[slot] = (int8_t)t;
                    ws.choice[0][slot] = kNone;
                    ws.choice[1][slot] = kNone;
                    ws.choice[2][slot] = kNone;
                }
            }
        }
    }
    occupied = warp_reduce_add((unsigned)occupied);  // per-thread counts are small
    if ((threadIdx.x & 31) == 0 && occupied) atomicAdd(&ws.counters[1], occupied);
}

struct GridView {
    const int8_t* type;
    const uint8_t* happy;
    const uint32_t* agent_id;
    uint64_t seed;
    uint32_t epoch;
};

// Is cell c empty at time t, given the arrivals in `arrive`?
__device__ __forceinline__ bool empty_at(const GridView& g, const unsigned long long* __restrict__ arrive,
                                         int64_t c, unsigned long long t) {
    if (arrive[c] < t) return false;  // somebody moved in earlier
    if (g.type[c] < 0) return true;   // empty since the start of the step
    if (g.happy[c]) return false;     // occupant does not move this step
    return mb::priority64(g.seed, g.epoch, g.agent_id[c]) < t;  // occupant already left
}

__global__ void relocate_iterate(GridView g, RelocWs ws, int64_t ncells, int64_t nmovers, int iter) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const unsigned long long* rd = ws.arrive[iter % 3];
    unsigned long long* claim = ws.arrive[(iter + 1) % 3];
    unsigned long long* reset = ws.arrive[(iter + 2) % 3];
    const uint32_t agent = ws.mv_agent[m];
    const unsigned long long t = mb::priority64(g.seed, g.epoch, agent);

    const uint32_t two_ago = ws.choice[(iter + 1) % 3][m];
    if (two_ago != kNone) reset[two_ago] = kNever;
    const uint32_t prev = ws.choice[(iter + 2) % 3][m];

    mb::DrawStream ds(g.seed, g.epoch, agent);
    uint32_t pick = kNone;
    for (int j = 0; j < kMaxProbes; ++j) {
        const int64_t c = (int64_t)ds.randbelow((uint64_t)ncells);
        if (empty_at(g, rd, c, t)) {
            pick = (uint32_t)c;
            break;
        }
    }
    if (pick == kNone) {
        // resolved by relocate_fallback (same iteration); it also counts the change
        const unsigned long long slot = atomicAdd(&ws.counters[2], 1ull);
        ws.fb_list[slot] = (uint32_t)m;
        ws.choice[iter % 3][m] = kNone;
        return;
    }
    ws.choice[iter % 3][m] = pick;
    atomicMin(&claim[pick], t);
    if (pick != prev) atomicAdd(&ws.counters[8 + iter], 1ull);
}

// One CTA per fallback mover: k-th empty cell (row-major) at the mover's time.
__global__ void __launch_bounds__(256)
relocate_fallback(GridView g, RelocWs ws, int64_t ncells, int64_t nempty, int iter) {
    __shared__ unsigned warp_tot[8];
    __shared__ long long found;
    const unsigned long long nfb = ws.counters[2];
    const unsigned long long* rd = ws.arrive[iter % 3];
    unsigned long long* claim = ws.arrive[(iter + 1) % 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long f = blockIdx.x; f < nfb; f += gridDim.x) {
        const uint32_t m = ws.fb_list[f];
        const uint32_t agent = ws.mv_agent[m];
        const unsigned long long t = mb::priority64(g.seed, g.epoch, agent);
        // replay the 50 accepted probes to position the stream, then draw the rank
        mb::DrawStream ds(g.seed, g.epoch, agent);
        for (int j = 0; j < kMaxProbes; ++j) (void)ds.randbelow((uint64_t)ncells);
        const long long k = (long long)ds.randbelow((uint64_t)nempty);
        if (threadIdx.x == 0) found = -1;
        __syncthreads();
        long long seen = 0;
        constexpr int kPer = 8;
        for (int64_t base = 0; base < ncells; base += 256 * kPer) {
            const int64_t c0 = base + (int64_t)threadIdx.x * kPer;
            unsigned bits = 0;
#pragma unroll
            for (int i = 0; i < kPer; ++i)
                if (c0 + i < ncells && empty_at(g, rd, c0 + i, t)) bits |= 1u << i;
            const unsigned cnt = __popc(bits);
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            unsigned before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w < warp) before += warp_tot[w];
                total += warp_tot[w];
            }
            const long long lo = seen + before + incl - cnt;  // empties before this thread's cells
            if (k >= lo && k < lo + cnt) {
                int need = (int)(k - lo);
                for (int i = 0; i < kPer; ++i)
                    if (bits & (1u << i)) {
                        if (need == 0) { found = c0 + i; break; }
                        --need;
                    }
            }
            seen += total;
            __syncthreads();
            if (found >= 0) break;
        }
        if (threadIdx.x == 0) {
            const uint32_t pick = (uint32_t)found;  // always found: nempty counts exactly these cells
            const uint32_t prev = ws.choice[(iter + 2) % 3][m];
            ws.choice[iter % 3][m] = pick;
            atomicMin(&claim[pick], t);
            if (pick != prev) atomicAdd(&ws.counters[8 + iter], 1ull);
        }
        __syncthreads();
    }
}

__global__ void reset_fallback_count(RelocWs ws) { ws.counters[2] = 0; }

__global__ void relocate_vacate(RelocWs ws, int8_t* type, uint8_t* happy, int64_t nmovers) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const uint32_t o = ws.mv_cell[m];
    type[o] = -1;
    happy[o] = 0;
}

__global__ void relocate_arrive(RelocWs ws, int8_t* type, uint8_t* happy, uint32_t* agent_id,
                                uint32_t* agent_cell, int64_t nmovers, int last_iter) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const uint32_t d = ws.choice[last_iter % 3][m];
    const uint32_t agent = ws.mv_agent[m];
    type[d] = ws.mv_type[m];
    happy[d] = 0;  // the flag travels with the agent: movers are unhappy until the next assign_state
    agent_id[d] = agent;
    if (agent_cell != nullptr) agent_cell[agent] = d;
    // restore the all-kNever invariant of the arrive tables
    ws.arrive[(last_iter + 1) % 3][d] = kNever;
    const uint32_t d1 = ws.choice[(last_iter + 2) % 3][m];  // iteration last_iter - 1
    if (d1 != kNone) ws.arrive[last_iter % 3][d1] = kNever;
    const uint32_t d2 = ws.choice[(last_iter + 1) % 3][m];  // iteration last_iter - 2 (normally reset already)
    if (d2 != kNone) ws.arrive[(last_iter + 2) % 3][d2] = kNever;
}

}  // namespace

MB_API int mb_relocate_workspace_bytes(int64_t ncells, int64_t max_movers, size_t* out) {
    MB_REQUIRE(out != nullptr && ncells >= 0 && max_movers >= 0, "mb_relocate_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, ncells, max_movers);
    return MB_OK;
}

MB_API int mb_relocate_workspace_init(void* workspace, size_t workspace_bytes, int64_t ncells,
                                      int64_t max_movers, cudaStream_t stream) {
    RelocWs ws;
    MB_REQUIRE(workspace != nullptr, "mb_relocate_workspace_init: null workspace");
    if (carve(&ws, (char*)workspace, ncells, max_movers) > workspace_bytes) {
        mb::set_error("mb_relocate_workspace_init: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    const int blocks = mb::sm_count() * 8;
    for (int i = 0; i < 3; ++i)
        if (ncells > 0) fill_u64<<<blocks, 256, 0, stream>>>(ws.arrive[i], ncells, kNever);
    MB_LAUNCH_CHECK("mb_relocate_workspace_init");
    return MB_OK;
}

// Synchronises `stream` (needs the mover count and the convergence flags on the host).
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
    int rc;
    rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, sizeof(unsigned long long) * (8 + kMaxIterSlots), stream),
                        "memset counters");
    if (rc) return rc;
    const int sms = mb::sm_count();
    {
        const int64_t want = mb::ceil_div(ncells, 256);
        const unsigned blocks = (unsigned)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
        collect_movers<<<blocks, 256, 0, stream>>>(type, happy, agent_id, ncells, ws, max_movers);
        MB_LAUNCH_CHECK("collect_movers");
    }
    unsigned long long hc[2];
    rc = mb::check_cuda(cudaMemcpyAsync(hc, ws.counters, sizeof(hc), cudaMemcpyDeviceToHost, stream), "read counters");
    if (rc) return rc;
    rc = mb::check_cuda(cudaStreamSynchronize(stream), "sync");
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

    const GridView g{type, happy, agent_id, seed, epoch};
    const unsigned mblocks = (unsigned)mb::ceil_div(nmovers, 128);
    int iter = 0;
    bool converged = false;
    while (!converged) {
        if (iter + kIterBatch >= kMaxIterSlots) {
            mb::set_error("mb_schelling_relocate: no convergence after %d iterations", iter);
            return MB_ERR_CUDA;
        }
        const int first = iter;
        for (int b = 0; b < kIterBatch; ++b, ++iter) {
            reset_fallback_count<<<1, 1, 0, stream>>>(ws);
            relocate_iterate<<<mblocks, 128, 0, stream>>>(g, ws, ncells, nmovers, iter);
            relocate_fallback<<<sms, 256, 0, stream>>>(g, ws, ncells, nempty, iter);
        }
        MB_LAUNCH_CHECK("relocate_iterate");
        unsigned long long changed[kIterBatch];
        rc = mb::check_cuda(cudaMemcpyAsync(changed, ws.counters + 8 + first, sizeof(changed),
                                            cudaMemcpyDeviceToHost, stream), "read changed");
        if (rc) return rc;
        rc = mb::check_cuda(cudaStreamSynchronize(stream), "sync");
        if (rc) return rc;
        converged = changed[kIterBatch - 1] == 0;
    }
    const int last = iter - 1;
    relocate_vacate<<<mblocks, 128, 0, stream>>>(ws, type, happy, nmovers);
    relocate_arrive<<<mblocks, 128, 0, stream>>>(ws, type, happy, agent_id, agent_cell, nmovers, last);
    MB_LAUNCH_CHECK("relocate_apply");
    if (out_iterations) *out_iterations = iter;
    return MB_OK;
}
