// Schelling happiness for every occupied cell at once (K1+K3 of SURVEY.md §2.4).
//
// Replaces AgentSet.do("assign_state") over
//   mesa/examples/basic/schelling/agents.py:24-42
//       neighbors = cell.get_neighborhood(radius).agents      (cell.py:202-276: Moore radius r
//                                                              = Chebyshev square minus centre)
//       similar = #{n.type == self.type}; valid = len(neighbors)
//       frac = similar / valid if valid > 0 else 0.0
//       happy = not (frac < homophily); model.happy += happy
//
// State: int8 type layer [rows, cols]: -1 empty cell, 0/1 agent type (capacity 1 grid,
// schelling/model.py:45-47).  Output: uint8 happy layer (1 = happy agent, 0 = unhappy or
// empty) and the global happy count.  Rows above/below the block come from halo_top /
// halo_bot ([radius, cols] each, NULL = outside the grid = empty) so the same kernel serves
// a row-sharded grid.
//
// The float64 divide + compare is made exact by a host-built table
// min_similar[valid] = smallest `similar` for which `not (similar / valid < homophily)`
// evaluated with Python floats (the reference's arithmetic); the vector kernel evaluates it
// on packed bytes as base + sum_k [valid >= step_k].
//
// Roofline: HBM-bound, 2 B per cell (1 B type read + 1 B happy write).
#include "common.cuh"

#include <string.h>

namespace {

constexpr int kWarps = 8;
constexpr int kRowsPerWarp = 64;
constexpr int kMaxRadius = 7;  // (2*7+1)^2 - 1 = 224 neighbours fit a byte

struct HappyRule {
    int mode;      // 0: happy iff similar >= thr(valid); 1: happy iff (valid - similar) <= thr(valid)
    int base;      // thr(valid) = base + #{k : valid >= steps[k]}  for valid >= 1
    int nsteps;
    int happy_if_alone;  // valid == 0 -> happy?
    uint8_t steps[24];
    uint8_t min_similar[232];  // full table for the scalar kernel (index = valid)
};

__device__ __forceinline__ unsigned splat(unsigned b) { return b * 0x01010101u; }

// E[0..5] = word left of the lane, 4 words of the lane, word right of the lane.
struct RowQ {
    unsigned occ[6];
    unsigned one[6];
};

template <int R>
__device__ __forceinline__ const int8_t* happy_row_ptr(const int8_t* __restrict__ in,
                                                       const int8_t* __restrict__ halo_top,
                                                       const int8_t* __restrict__ halo_bot, int64_t r,
                                                       int64_t rows, int64_t cols) {
    if (r < 0) return (halo_top != nullptr && r >= -R) ? halo_top + (R + r) * cols : nullptr;
    if (r >= rows) return (halo_bot != nullptr && r - rows < R) ? halo_bot + (r - rows) * cols : nullptr;
    return in + r * cols;
}

__device__ __forceinline__ RowQ load_rowq(const int8_t* __restrict__ p, int64_t col, bool active, bool edge_l,
                                          bool edge_r, bool ld_left, bool ld_right, int64_t lcol,
                                          int64_t rcol) {
    uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);  // -1 = empty outside the grid
    unsigned lw = ~0u, rw = ~0u;
    if (p != nullptr) {
        if (active) v = __ldg(reinterpret_cast<const uint4*>(p + col));
        if (ld_left) lw = __ldg(reinterpret_cast<const unsigned*>(p + lcol));
        if (ld_right) rw = __ldg(reinterpret_cast<const unsigned*>(p + rcol));
    }
    const unsigned sl = __shfl_up_sync(0xffffffffu, v.w, 1);
    const unsigned sr = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (!edge_l) lw = sl;
    if (!edge_r) rw = sr;
    const unsigned raw[6] = {lw, v.x, v.y, v.z, v.w, rw};
    RowQ q;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned occ = ~(raw[k] >> 7) & 0x01010101u;  // type >= 0
        q.occ[k] = occ;
        q.one[k] = raw[k] & occ;  // type == 1
    }
    return q;
}

template <int R>
__global__ void __launch_bounds__(kWarps * 32)
schelling_happy_vec(const int8_t* __restrict__ type, const int8_t* __restrict__ halo_top,
                    const int8_t* __restrict__ halo_bot, uint8_t* __restrict__ happy, int64_t rows,
                    int64_t cols, int col_torus, int nstrips, const HappyRule rule,
                    unsigned long long* __restrict__ happy_count) {
    const int lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t strip = task % nstrips;
    const int64_t row0 = (task / nstrips) * kRowsPerWarp;
    if (row0 >= rows) return;

    const int64_t col = strip * 512 + lane * 16;
    const bool active = col < cols;
    const bool edge_l = lane == 0;
    const bool edge_r = (lane == 31) || (col + 16 >= cols);
    bool ld_left = active && edge_l, ld_right = active && edge_r;
    int64_t lcol = col - 4, rcol = col + 16;
    if (ld_left && lcol < 0) {
        if (col_torus) lcol = cols - 4; else ld_left = false;
    }
    if (ld_right && rcol >= cols) {
        if (col_torus) rcol = 0; else ld_right = false;
    }

    constexpr int W = 2 * R + 1;
    RowQ ring[W];
    unsigned vocc[6], vone[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) vocc[k] = vone[k] = 0u;
    // ring[0..W-2] = rows row0-R .. row0+R-1 ; the loop loads row r+R into ring[W-1]
#pragma unroll
    for (int i = 0; i < W - 1; ++i) {
        ring[i] = load_rowq(happy_row_ptr<R>(type, halo_top, halo_bot, row0 - R + i, rows, cols), col, active,
                            edge_l, edge_r, ld_left, ld_right, lcol, rcol);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vocc[k] += ring[i].occ[k];
            vone[k] += ring[i].one[k];
        }
    }

    unsigned nhappy = 0;
    const int64_t rend = min(row0 + (int64_t)kRowsPerWarp, rows);
#pragma unroll 2
    for (int64_t r = row0; r < rend; ++r) {
        ring[W - 1] = load_rowq(happy_row_ptr<R>(type, halo_top, halo_bot, r + R, rows, cols), col, active,
                                edge_l, edge_r, ld_left, ld_right, lcol, rcol);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vocc[k] += ring[W - 1].occ[k];
            vone[k] += ring[W - 1].one[k];
        }
        uint4 o;
        unsigned* op = reinterpret_cast<unsigned*>(&o);
#pragma unroll
        for (int k = 1; k <= 4; ++k) {
            unsigned bocc = vocc[k], bone = vone[k];
#pragma unroll
            for (int d = 1; d <= R; ++d) {
                bocc += __funnelshift_r(vocc[k], vocc[k + 1], 8 * d) + __funnelshift_l(vocc[k - 1], vocc[k], 8 * d);
                bone += __funnelshift_r(vone[k], vone[k + 1], 8 * d) + __funnelshift_l(vone[k - 1], vone[k], 8 * d);
            }
            const unsigned occ_c = ring[R].occ[k], one_c = ring[R].one[k];
            const unsigned valid = bocc - occ_c;          // occupied neighbours (for occupied centres)
            const unsigned m1 = one_c * 0xffu;            // 0xff where the centre is type 1
            const unsigned sim1 = bone - one_c;           // type-1 neighbours
            const unsigned similar = (m1 & sim1) | (~m1 & (valid - sim1));
            unsigned thr = splat((unsigned)rule.base);
            for (int s = 0; s < rule.nsteps; ++s) thr += __vsetgeu4(valid, splat(rule.steps[s]));
            const unsigned cmp = rule.mode == 0 ? __vsetgeu4(similar, thr) : __vsetleu4(valid - similar, thr);
            const unsigned alone = __vseteq4(valid, 0u);
            const unsigned res = occ_c & ((cmp & ~alone) | (alone & splat((unsigned)rule.happy_if_alone)));
            op[k - 1] = res;
            nhappy += __popc(res);
        }
        if (active) __stcs(reinterpret_cast<uint4*>(happy + r * cols + col), o);
        else nhappy = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            vocc[k] -= ring[0].occ[k];
            vone[k] -= ring[0].one[k];
        }
#pragma unroll
        for (int i = 0; i < W - 1; ++i) ring[i] = ring[i + 1];
    }
    nhappy = warp_reduce_add(active ? nhappy : 0u);
    if (lane == 0 && nhappy != 0 && happy_count != nullptr) atomicAdd(happy_count, (unsigned long long)nhappy);
}

// Any radius <= 7, any type values 0..127, any shape: one thread per cell.
__global__ void schelling_happy_scalar(const int8_t* __restrict__ type, const int8_t* __restrict__ halo_top,
                                       const int8_t* __restrict__ halo_bot, uint8_t* __restrict__ happy,
                                       int64_t rows, int64_t cols, int radius, int col_torus,
                                       const HappyRule rule, unsigned long long* __restrict__ happy_count) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned res = 0;
    if (idx < rows * cols) {
        const int64_t i = idx / cols, j = idx % cols;
        const int t = type[idx];
        if (t >= 0) {
            int valid = 0, similar = 0;
            for (int di = -radius; di <= radius; ++di) {
                const int64_t r = i + di;
                const int8_t* p;
                if (r < 0) p = (halo_top != nullptr && r >= -radius) ? halo_top + (radius + r) * cols : nullptr;
                else if (r >= rows) p = (halo_bot != nullptr && r - rows < radius) ? halo_bot + (r - rows) * cols : nullptr;
                else p = type + r * cols;
                if (p == nullptr) continue;
                for (int dj = -radius; dj <= radius; ++dj) {
                    if (di == 0 && dj == 0) continue;
                    int64_t jj = j + dj;
                    if (jj < 0 || jj >= cols) {
                        if (!col_torus) continue;
                        jj = (jj % cols + cols) % cols;
                    }
                    const int n = p[jj];
                    valid += n >= 0;
                    similar += n == t;
                }
            }
            res = similar >= (int)rule.min_similar[valid];
        }
        happy[idx] = (uint8_t)res;
    }
    const unsigned n = warp_reduce_add(res);
    if ((threadIdx.x & 31) == 0 && n != 0 && happy_count != nullptr) atomicAdd(happy_count, (unsigned long long)n);
}

// Decompose the threshold table into base + unit steps (see file header); false if the
// table is not of that shape (then the scalar kernel evaluates it directly).
bool make_rule(const uint8_t* tbl, int nmax, HappyRule* rule) {
    memset(rule, 0, sizeof(*rule));
    for (int v = 0; v <= nmax; ++v) rule->min_similar[v] = tbl[v];
    rule->happy_if_alone = tbl[0] == 0;
    if (nmax < 1) return false;
    // representation T: similar >= T[v]
    bool okT = true, okU = true;
    int stepsT[256], nT = 0, stepsU[256], nU = 0;
    for (int v = 2; v <= nmax; ++v) {
        const int d = (int)tbl[v] - (int)tbl[v - 1];
        if (d == 1) stepsT[nT++] = v; else if (d != 0) okT = false;
        // U[v] = v - T[v]; increment = 1 - d
        if (d == 0) stepsU[nU++] = v; else if (d != 1) okU = false;
    }
    const int baseU = 1 - (int)tbl[1];
    if (baseU < 0) okU = false;
    if (nT > 24) okT = false;
    if (nU > 24) okU = false;
    if (okU && (!okT || nU < nT)) {
        rule->mode = 1; rule->base = baseU; rule->nsteps = nU;
        for (int k = 0; k < nU; ++k) rule->steps[k] = (uint8_t)stepsU[k];
        return true;
    }
    if (okT) {
        rule->mode = 0; rule->base = tbl[1]; rule->nsteps = nT;
        for (int k = 0; k < nT; ++k) rule->steps[k] = (uint8_t)stepsT[k];
        return true;
    }
    return false;
}

}  // namespace

MB_API int mb_schelling_happy_i8(const int8_t* type, const int8_t* halo_top, const int8_t* halo_bot,
                                 uint8_t* happy, int64_t rows, int64_t cols, int radius, int col_torus,
                                 const uint8_t* min_similar_host, int binary_types,
                                 unsigned long long* happy_count, cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0, "mb_schelling_happy_i8: negative shape");
    // radius < 1 is a ValueError in the reference too (cell.py:234-235)
    MB_REQUIRE(radius >= 1, "mb_schelling_happy_i8: radius must be at least 1");
    if (radius > kMaxRadius) {
        mb::set_error("mb_schelling_happy_i8: radius %d > %d not supported", radius, kMaxRadius);
        return MB_ERR_UNSUPPORTED;
    }
    MB_REQUIRE(min_similar_host != nullptr, "mb_schelling_happy_i8: null threshold table");
    MB_REQUIRE(!(col_torus && cols > 0 && cols < 2 * radius + 1),
               "mb_schelling_happy_i8: torus needs cols >= 2*radius+1 (SURVEY A.2)");
    if (happy_count != nullptr) {
        int rc = mb::check_cuda(cudaMemsetAsync(happy_count, 0, sizeof(unsigned long long), stream), "memset");
        if (rc != MB_OK) return rc;
    }
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(type != nullptr && happy != nullptr, "mb_schelling_happy_i8: null buffer");
    const int nmax = (2 * radius + 1) * (2 * radius + 1) - 1;
    HappyRule rule;
    const bool packed = make_rule(min_similar_host, nmax, &rule);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = packed && binary_types && radius <= 2 && cols % 16 == 0 && al16(type) && al16(happy) &&
                     (halo_top == nullptr || al16(halo_top)) && (halo_bot == nullptr || al16(halo_bot));
    if (vec) {
        const int nstrips = (int)mb::ceil_div(cols, 512);
        const int64_t tasks = (int64_t)nstrips * mb::ceil_div(rows, kRowsPerWarp);
        const unsigned blocks = (unsigned)mb::ceil_div(tasks, kWarps);
        if (radius == 1)
            schelling_happy_vec<1><<<blocks, kWarps * 32, 0, stream>>>(type, halo_top, halo_bot, happy, rows, cols,
                                                                       col_torus, nstrips, rule, happy_count);
        else
            schelling_happy_vec<2><<<blocks, kWarps * 32, 0, stream>>>(type, halo_top, halo_bot, happy, rows, cols,
                                                                       col_torus, nstrips, rule, happy_count);
    } else {
        const int64_t n = rows * cols;
        schelling_happy_scalar<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(
            type, halo_top, halo_bot, happy, rows, cols, radius, col_torus, rule, happy_count);
    }
    MB_LAUNCH_CHECK("mb_schelling_happy_i8");
    return MB_OK;
}
