// Sugarscape G1MT, the movement phase: regrowth + one shuffle_do("step") of the traders (move / eat / die) in ONE launch,
// in the reference's sequential semantics.
//
//   mesa/examples/advanced/sugarscape_g1mt/model.py:118-127   sugar / spice regrow by 1 up to the landscape's maximum,
//                                                             then self.agents.shuffle_do("step")
//   agents.py:209-262   move: the EMPTY cells of get_neighborhood(vision, include_center=True) (von Neumann, clamped, BFS
//                       order), Cobb-Douglas welfare of (own + cell) sugar and spice, the cells math.isclose to the best
//                       welfare, of those the ones within 1 % of the smallest Euclidean distance, random.choice of these
//   agents.py:264-288   eat: take the cell's sugar and spice, pay the metabolism; die: sugar <= 0 or spice <= 0 -> remove()
//
// The same rank-ordered dependency graph as csrc/epstein.cu: a trader reads (emptiness, sugar, spice) within its own vision
// and changes its old and new cell, so it depends on an earlier trader only if that one STARTED within vision_a + vision_b
// <= 2 max-vision of it.  A warp owns the trader of the next activation rank (ticket counter), polls the cells of that range
// until no lower rank is pending there, then acts on the live layers.  Cells are not capacity-limited (the initial
// placement may stack traders, model.py:84): a cell keeps up to 4 pending ranks (one 16-byte word; more: status 2).
// Keys: priority and draws are keyed on unique_id - 1 (traders die, rows are not ids).
//
// Float64: the welfare is the reference's expression with CUDA's pow (<= 2 ulp from C's); the DECISIONS only use it through
// max + isclose(rel_tol 1e-9) and are therefore the reference's unless two different welfares lie within 1e-9 of each other
// but not within 1e-9 +- 1e-16 -- the stored amounts (sugar, spice: sums of small integers) are exact.
#include "common.cuh"
#include "philox.cuh"
#include "sort.cuh"

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
    uint32_t* counters;  // [0] ticket, [1] status (1: spin limit, 2: more than 4 traders started in one cell), [2] deaths
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
    int64_t n;
    int width, height, nw;
    uint64_t seed;
    uint32_t epoch;
};

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
    const double mt = (double)(ms + mp);
    const double es = (double)ms / mt, ep = (double)mp / mt;       // agents.py:66-70
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
            w[q] = pow(own_sugar + cs[q], es) * pow(own_spice + cp[q], ep);
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
// dead [n] (uint8, out) marks the traders that starved: the caller compacts them away.  status_dev (device, 2 x uint32,
// optional): status (0 valid; 1 a wait hit its spin limit; 2 more than 4 traders started the step in one cell), deaths.
MB_API int mb_sugarscape_step(int32_t* cell, double* sugar, double* spice, const int32_t* metabolism_sugar,
                              const int32_t* metabolism_spice, const int32_t* vision, const int64_t* unique_id, uint8_t* dead,
                              int64_t n, double* sugar_layer, const double* sugar_max, double* spice_layer,
                              const double* spice_max, int32_t* count, int64_t width, int64_t height, const int16_t* voff,
                              int max_vision, const int16_t* woff, int nw, int regrow, uint64_t seed, uint32_t epoch,
                              uint32_t* status_dev, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
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
    if (status_dev != nullptr &&
        (rc = mb::check_cuda(cudaMemcpyAsync(status_dev, ws.counters + 1, 4 * 2, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    MB_LAUNCH_CHECK("mb_sugarscape_step");
    return MB_OK;
}
