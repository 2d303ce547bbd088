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
// thread r owns the agent of activation rank r (ascending Philox priority), CTAs take their rank ranges from a ticket
// counter (every lower rank is running or done), and a thread acts once every lower-ranked agent that started within
// 2 vision of it has finished -- found through rank_at[cell] (the rank of the agent that started the step in a cell) and
// done[rank].  When it acts it sees exactly the state the sequential run would show it.  Inside a warp nobody blocks: the
// lanes loop, each scans on from where it last had to wait, and acts when it can.
//
// One 4-byte word per cell carries all a neighbour scan needs: -1 = empty, else agent << 2 | cop << 1 | active, kept up to
// date by whoever changes it (the mover, the citizen that changes its mind, the cop that arrests).  The arrest probability
// takes the value ptab[round(cops / actives)], a host-made table of the reference's own expression (math.exp on the
// integer multiples of the constant: bit-identical by construction); round() is round-half-even on the float64 quotient
// like Python's; the comparison is the reference's float64 expression (no fused operations).
#include "common.cuh"
#include "philox.cuh"

extern "C" int mb_activation_order_workspace_bytes(int64_t n, size_t* out);
extern "C" int mb_activation_order(uint64_t seed, uint32_t epoch, int64_t n, uint32_t* order_out, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream);

namespace {

constexpr uint32_t kNoRank = 0xffffffffu;
constexpr int kThreads = 128;
constexpr unsigned kSpinLimit = 1u << 24;   // warp loop iterations before a thread gives up (status word, no hang)
constexpr uint8_t kActive = 1, kQuiet = 2, kArrested = 3;   // CitizenState (agents.py:7-10)

struct EpsWs {
    uint32_t* order;     // [n] agent of rank r
    uint32_t* rank_at;   // [ncells] rank of the agent that starts the step in the cell, kNoRank if empty
    uint32_t* done;      // [n] by rank
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
    t.done = (uint32_t*)take(4 * n);
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

__global__ void eps_prepare(EpsArgs a, EpsWs ws, int sequential) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    if (sequential) ws.order[r] = (uint32_t)r;          // agents.do("step"): registration order
    ws.rank_at[a.cell[ws.order[r]]] = (uint32_t)r;
    ws.done[r] = 0u;
}

__device__ __forceinline__ int wrap_cell(int x, int y, short2 d, int W, int H) {
    int nx = x + d.x, ny = y + d.y;
    nx += nx < 0 ? W : 0; nx -= nx >= W ? W : 0;
    ny += ny < 0 ? H : 0; ny -= ny >= H ? H : 0;
    return nx * H + ny;
}

// the k-th cell of the neighbourhood (BFS order) whose word satisfies `want` (empty: word < 0; active citizen: (word & 3) == 1)
template <bool kEmpty>
__device__ __forceinline__ int kth_cell(const EpsArgs& a, int x, int y, int k, int32_t* word_out) {
    for (int i0 = 0; i0 < a.nv; i0 += 8) {
        int c[8];
        int32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            c[q] = i0 + q < a.nv ? wrap_cell(x, y, __ldg(a.voff + i0 + q), a.width, a.height) : -1;
            w[q] = c[q] >= 0 ? __ldcg(a.occ + c[q]) : 0;    // 0 = agent 0, citizen, not active: matches neither predicate... (guarded below)
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (c[q] < 0) continue;
            const bool hit = kEmpty ? w[q] < 0 : (w[q] >= 0 && (w[q] & 3) == 1);
            if (hit && k-- == 0) { *word_out = w[q]; return c[q]; }
        }
    }
    return -1;
}

__device__ void act(const EpsArgs& a, uint32_t agent) {
    const int p = __ldcg(a.cell + agent);
    const int x = p / a.height, y = p - x * a.height;
    const bool cop = __ldg(a.kind + agent) != 0;
    mb::DrawStream ds(a.seed, a.epoch, (uint64_t)agent);
    if (!cop) {
        const int j = __ldcg(a.jail + agent);
        if (j != 0) {                                   // agents.py:70-72: in jail, nothing else happens
            a.jail[agent] = j - 1;
            return;
        }
    }
    // one scan of the neighbourhood: empty cells, cops, active citizens
    int nempty = 0, ncops = 0, nactive = 0;
    for (int i0 = 0; i0 < a.nv; i0 += 8) {
        int32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
            w[q] = i0 + q < a.nv ? __ldcg(a.occ + wrap_cell(x, y, __ldg(a.voff + i0 + q), a.width, a.height)) : 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (i0 + q >= a.nv) continue;
            nempty += w[q] < 0;
            ncops += w[q] >= 0 && (w[q] & 2);
            nactive += w[q] >= 0 && (w[q] & 3) == 1;
        }
    }
    int32_t me = (int32_t)(agent << 2) | (cop ? 2 : 0);
    if (!cop) {
        // agents.py:86-103 (the citizen counts herself among the actives)
        const double q = (double)ncops / (double)(nactive + 1);
        const int m = __double2int_rn(q);               // round-half-even, like Python's round()
        const double prob = a.ptab[m];
        a.arrest_probability[agent] = prob;
        const double net_risk = __dmul_rn(__ldg(a.risk_aversion + agent), prob);
        const bool active = __dsub_rn(__ldg(a.grievance + agent), net_risk) > a.threshold;
        a.state[agent] = active ? kActive : kQuiet;
        me |= active ? 1 : 0;
        a.occ[p] = me;                                  // (rewritten below if she moves)
    } else if (nactive > 0) {
        // agents.py:130-138
        const int pick = (int)ds.randbelow((uint64_t)nactive);
        int32_t w = 0;
        const int c = kth_cell<false>(a, x, y, pick, &w);
        const uint32_t arrestee = (uint32_t)w >> 2;
        a.jail[arrestee] = (int32_t)ds.randbelow((uint64_t)a.max_jail_term + 1ull);   // randint(0, max_jail_term)
        a.state[arrestee] = kArrested;
        a.occ[c] = w & ~1;                              // no longer active
    }
    if (a.movement && nempty > 0) {                     // agents.py:22-26
        const int pick = (int)ds.randbelow((uint64_t)nempty);
        int32_t w = 0;
        const int dest = kth_cell<true>(a, x, y, pick, &w);
        a.occ[p] = -1;
        a.occ[dest] = me;
        a.cell[agent] = dest;
    }
}

__global__ void __launch_bounds__(kThreads)
eps_chain(EpsArgs a, EpsWs ws) {
    __shared__ unsigned s_base;
    if (threadIdx.x == 0) s_base = atomicAdd(&ws.counters[0], (unsigned)kThreads);
    __syncthreads();
    const int64_t r = (int64_t)s_base + threadIdx.x;
    bool finished = r >= a.n;
    uint32_t agent = 0;
    int x = 0, y = 0, i = 0;
    if (!finished) {
        agent = ws.order[r];
        const int p = a.cell[agent];                   // its own cell: nobody else moves it
        x = p / a.height; y = p - x * a.height;
    }
    unsigned spins = 0;
    while (true) {
        if (!finished) {
            // every agent of lower rank that STARTED within 2 vision must have finished (done only ever goes 0 -> 1, so the
            // scan resumes where it last had to wait)
            while (i < a.nw) {
                const uint32_t rk = __ldg(ws.rank_at + wrap_cell(x, y, __ldg(a.woff + i), a.width, a.height));
                if (rk < (uint32_t)r && ld_relaxed(ws.done + rk) == 0u) break;
                ++i;
            }
            if (i == a.nw) {
                __threadfence();                        // their writes, before my reads
                act(a, agent);
                __threadfence();                        // my writes, before my flag
                st_relaxed(ws.done + r, 1u);
                finished = true;
            } else if (++spins > kSpinLimit) {
                ws.counters[1] = 1u;                    // gives up WITHOUT acting: the status word reports it
                st_relaxed(ws.done + r, 1u);
                finished = true;
            }
        }
        if (__all_sync(0xffffffffu, finished)) break;
    }
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
    const unsigned blocks = (unsigned)mb::ceil_div(n, kThreads);
    eps_prepare<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(a, ws, sequential);
    eps_chain<<<blocks, kThreads, 0, stream>>>(a, ws);
    if (counts_dev != nullptr) {
        eps_counts<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(kind, state, n, ws.counters);
        if ((rc = mb::check_cuda(cudaMemcpyAsync(counts_dev, ws.counters + 1, 4 * 4, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    }
    MB_LAUNCH_CHECK("mb_epstein_step");
    return MB_OK;
}
