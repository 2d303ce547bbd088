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

#include <stdlib.h>
#include <string.h>

namespace mb {
// csrc/schelling_bits.cu: the bit-plane kernel (radius <= 2, binary types, cols % 32 == 0, rules with a small integer form)
int happy_bits_launch(const int8_t* type, const int8_t* halo_top, const int8_t* halo_bot, uint8_t* happy, int64_t rows,
                      int64_t cols, int radius, int col_torus, const uint8_t* min_similar_host,
                      unsigned long long* happy_count, cudaStream_t stream, bool* handled);
}

namespace {

constexpr int kWarps = 8;
constexpr int kMaxRadius = 7;  // (2*7+1)^2 - 1 = 224 neighbours fit a byte

enum { MODE_LINEAR = 0, MODE_T = 1, MODE_U = 2 };

// How the vector kernel evaluates `similar >= min_similar[valid]` on packed bytes (all < 128):
//   MODE_LINEAR  A*similar + C >= B*valid        (host-verified against the exact table, e.g. h=0.4 -> 5s >= 2v)
//   MODE_T       similar >= base + #{k : valid >= steps[k]}
//   MODE_U       valid - similar <= base + #{k : valid >= steps[k]}
// valid == 0 is decided by happy_if_alone.  The scalar kernel reads min_similar[] directly.
struct HappyRule {
    int mode;
    int A, B, C;
    int base, nsteps;
    int happy_if_alone;
    uint8_t steps[24];
    uint8_t min_similar[232];
};

__device__ __forceinline__ unsigned splat(unsigned b) { return b * 0x01010101u; }

// One input row of a lane after the horizontal pass.
//   R == 1: h[k] holds both box-row sums per byte: low nibble = #occupied, high nibble = #type-1 (each <= 3)
template <int R>
struct RowH {
    unsigned h[4];
    unsigned occ[4];  // centre cells: occupied (0/1 per byte)
    unsigned one[4];  // centre cells: type 1
    // (packing the centre cells like the radius-2 ring saves 15 registers but costs 12 unpacking operations per row:
    //  121 us at 4 CTAs per SM, 119 us at 5 against 115 us -- radius 1 is instruction-bound, not occupancy-bound)
};
//   R == 2, packed: hg[k] holds both 5-column sums per byte (low nibble = #occupied, high nibble = #type-1, each <= 5)
//   and cen[k] the centre cells (bit 0 = occupied, bit 4 = type 1): 8 registers per row instead of 16, so the 5-row
//   ring fits the register budget of 3 CTAs per SM (the unpacked ring: 128 registers, 2 CTAs, occupancy-bound).
template <>
struct RowH<2> {
    unsigned hg[4];
    unsigned cen[4];
};

struct LaneAddr {
    unsigned col, lcol, rcol;
    bool active, edge_l, edge_r, ld_left, ld_right;
};

struct RawRow {
    uint4 v;
    unsigned lw, rw;  // aligned word left / right of the lane's 16 cells (strip-edge lanes only)
};

template <bool kChecked>
__device__ __forceinline__ RawRow load_raw(const int8_t* __restrict__ row, const LaneAddr& la) {
    RawRow x;
    x.v = make_uint4(~0u, ~0u, ~0u, ~0u);  // -1 = empty outside the grid
    x.lw = ~0u;
    x.rw = ~0u;
    if (!kChecked || row != nullptr) {
        if (la.active) x.v = __ldg(reinterpret_cast<const uint4*>(row + la.col));
        if (la.ld_left) x.lw = __ldg(reinterpret_cast<const unsigned*>(row + la.lcol));
        if (la.ld_right) x.rw = __ldg(reinterpret_cast<const unsigned*>(row + la.rcol));
    }
    return x;
}

template <int R>
__device__ __forceinline__ RowH<R> make_rowh(const RawRow& x, const LaneAddr& la) {
    const uint4 v = x.v;
    unsigned lw = x.lw, rw = x.rw;
    const unsigned sl = __shfl_up_sync(0xffffffffu, v.w, 1);
    const unsigned sr = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (!la.edge_l) lw = sl;
    if (!la.edge_r) rw = sr;
    const unsigned raw[6] = {lw, v.x, v.y, v.z, v.w, rw};
    RowH<R> q;
    if constexpr (R == 1) {
        unsigned e[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const unsigned occ = ~(raw[k] >> 7) & 0x01010101u;  // type >= 0
            const unsigned one = raw[k] & occ;                  // type == 1
            e[k] = occ + (one << 4);
            if (k >= 1 && k <= 4) { q.occ[k - 1] = occ; q.one[k - 1] = one; }
        }
#pragma unroll
        for (int k = 1; k <= 4; ++k)
            q.h[k - 1] = e[k] + __funnelshift_r(e[k], e[k + 1], 8) + __funnelshift_l(e[k - 1], e[k], 8);
    } else {
        unsigned e[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const unsigned occ = ~(raw[k] >> 7) & 0x01010101u;
            const unsigned one = raw[k] & occ;
            e[k] = occ + (one << 4);
            if (k >= 1 && k <= 4) q.cen[k - 1] = e[k];   // bit 0 = occupied, bit 4 = type 1
        }
#pragma unroll
        for (int k = 1; k <= 4; ++k) {
            unsigned a = e[k];
#pragma unroll
            for (int d = 1; d <= R; ++d)
                a += __funnelshift_r(e[k], e[k + 1], 8 * d) + __funnelshift_l(e[k - 1], e[k], 8 * d);
            q.hg[k - 1] = a;   // each nibble <= 5
        }
    }
    return q;
}

template <int R>
__device__ __forceinline__ const int8_t* happy_row_ptr(const int8_t* __restrict__ in,
                                                       const int8_t* __restrict__ halo_top,
                                                       const int8_t* __restrict__ halo_bot, int r, int rows,
                                                       size_t pitch) {
    if (r < 0) return (halo_top != nullptr && r >= -R) ? halo_top + (size_t)(R + r) * pitch : nullptr;
    if (r >= rows) return (halo_bot != nullptr && r - rows < R) ? halo_bot + (size_t)(r - rows) * pitch : nullptr;
    return in + (size_t)r * pitch;
}

// happiness of 4 packed centre cells from the box sums (incl. the centre) of occupied / type-1 cells
template <int MODE>
__device__ __forceinline__ unsigned decide4(unsigned bocc, unsigned bone, unsigned occ_c, unsigned one_c,
                                            const HappyRule& rule) {
    const unsigned valid = bocc - occ_c;                       // occupied neighbours
    const unsigned m1 = one_c * 0xffu;                         // 0xff where the centre is type 1
    const unsigned same = (m1 & bone) | (~m1 & (bocc - bone)); // cells of the centre's type incl. itself
    const unsigned similar = same - occ_c;
    unsigned flag;  // bit 7 of each byte = happy
    if (MODE == MODE_LINEAR) {
        flag = (similar * (unsigned)rule.A + splat((unsigned)rule.C + 128u)) - valid * (unsigned)rule.B;
    } else {
        const unsigned vo = valid | 0x80808080u;
        unsigned thr = splat((unsigned)rule.base);
        for (int s = 0; s < rule.nsteps; ++s) thr += ((vo - splat(rule.steps[s])) >> 7) & 0x01010101u;
        flag = MODE == MODE_T ? (similar | 0x80808080u) - thr : (thr | 0x80808080u) - (valid - similar);
    }
    const unsigned nonalone = valid + 0x7f7f7f7fu;             // bit 7 = valid != 0
    flag = rule.happy_if_alone ? (flag | ~nonalone) : (flag & nonalone);
    return (flag >> 7) & occ_c;
}

template <int R, int MODE, int kRowsPerWarp, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
schelling_happy_vec(const int8_t* __restrict__ type, const int8_t* __restrict__ halo_top,
                    const int8_t* __restrict__ halo_bot, uint8_t* __restrict__ happy, int rows, int cols,
                    int col_torus, int nstrips, const HappyRule rule,
                    unsigned long long* __restrict__ happy_count) {
    const int lane = threadIdx.x & 31;
    const unsigned task = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const unsigned strip = task % (unsigned)nstrips;
    const int row0 = (int)(task / (unsigned)nstrips) * kRowsPerWarp;
    if (row0 >= rows) return;

    LaneAddr la;
    la.col = strip * 512u + lane * 16u;
    la.active = la.col < (unsigned)cols;
    la.edge_l = lane == 0;
    la.edge_r = (lane == 31) || (la.col + 16u >= (unsigned)cols);
    la.ld_left = la.active && la.edge_l;
    la.ld_right = la.active && la.edge_r;
    la.lcol = la.col - 4u;
    la.rcol = la.col + 16u;
    if (la.ld_left && la.col == 0u) {
        if (col_torus) la.lcol = cols - 4; else la.ld_left = false;
    }
    if (la.ld_right && la.rcol >= (unsigned)cols) {
        if (col_torus) la.rcol = 0u; else la.ld_right = false;
    }
    const size_t pitch = (size_t)cols;

    constexpr int W = 2 * R + 1;
    constexpr int kAhead = 2;  // rows fetched ahead of the compute
    RowH<R> ring[W];  // ring[i] = row r - R + i while row r is being produced
#pragma unroll
    for (int i = 0; i < W - 1; ++i)
        ring[i + 1] = make_rowh<R>(load_raw<true>(happy_row_ptr<R>(type, halo_top, halo_bot, row0 - R + i, rows, pitch), la), la);

    unsigned nhappy = 0;
    const int rend = min(row0 + kRowsPerWarp, rows);
    uint8_t* op = happy + (size_t)row0 * pitch + la.col;

    auto produce = [&](const RawRow& fresh) {
#pragma unroll
        for (int i = 0; i < W - 1; ++i) ring[i] = ring[i + 1];
        ring[W - 1] = make_rowh<R>(fresh, la);
        uint4 o;
        unsigned* ow = reinterpret_cast<unsigned*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned bocc, bone, occ_c, one_c;
            if constexpr (R == 1) {
                const unsigned b = ring[0].h[k] + ring[1].h[k] + ring[2].h[k];
                bocc = b & 0x0f0f0f0fu;
                bone = (b >> 4) & 0x0f0f0f0fu;
                occ_c = ring[R].occ[k];
                one_c = ring[R].one[k];
            } else {
                // three rows (nibbles <= 15) + two rows (<= 10): no nibble overflows before the unpacking
                const unsigned sa = ring[0].hg[k] + ring[1].hg[k] + ring[2].hg[k];
                const unsigned sb = ring[3].hg[k] + ring[4].hg[k];
                bocc = (sa & 0x0f0f0f0fu) + (sb & 0x0f0f0f0fu);
                bone = ((sa >> 4) & 0x0f0f0f0fu) + ((sb >> 4) & 0x0f0f0f0fu);
                occ_c = ring[R].cen[k] & 0x01010101u;
                one_c = (ring[R].cen[k] >> 4) & 0x01010101u;
            }
            const unsigned res = decide4<MODE>(bocc, bone, occ_c, one_c, rule);
            ow[k] = res;
            nhappy += __popc(res);
        }
        if (la.active) __stcs(reinterpret_cast<uint4*>(op), o);
        op += pitch;
    };

    int r = row0;
    const int8_t* fp = type + (size_t)(row0 + R) * pitch;  // row r + R
    // groups whose look-ahead rows r+R .. r+R+kAhead-1 all lie inside this block: unchecked loads
    for (; r + kAhead <= rend && r + R + kAhead <= rows; r += kAhead) {
        RawRow nx[kAhead];
#pragma unroll
        for (int k = 0; k < kAhead; ++k) nx[k] = load_raw<false>(fp + (size_t)k * pitch, la);
        fp += (size_t)kAhead * pitch;
#pragma unroll
        for (int k = 0; k < kAhead; ++k) produce(nx[k]);
    }
    for (; r < rend; ++r) produce(load_raw<true>(happy_row_ptr<R>(type, halo_top, halo_bot, r + R, rows, pitch), la));
    nhappy = warp_reduce_add(la.active ? nhappy : 0u);
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
    // linear form A*s + C >= B*v, all byte quantities < 128, verified on every (s, v) with v >= 1
    const int lim = 127 / nmax;
    for (int A = 1; A <= lim; ++A)
        for (int B = 0; B <= lim; ++B)
            for (int C = 0; C <= (A < 8 ? A : 8) && A * nmax + C <= 127; ++C) {
                bool ok = true;
                for (int v = 1; v <= nmax && ok; ++v)
                    for (int sim = 0; sim <= v; ++sim)
                        if ((A * sim + C >= B * v) != (sim >= (int)tbl[v])) { ok = false; break; }
                if (ok) {
                    rule->mode = MODE_LINEAR; rule->A = A; rule->B = B; rule->C = C;
                    return true;
                }
            }
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
        rule->mode = MODE_U; rule->base = baseU; rule->nsteps = nU;
        for (int k = 0; k < nU; ++k) rule->steps[k] = (uint8_t)stepsU[k];
        return true;
    }
    if (okT) {
        rule->mode = MODE_T; rule->base = tbl[1]; rule->nsteps = nT;
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
    if (binary_types) {
        bool handled = false;
        const int rc = mb::happy_bits_launch(type, halo_top, halo_bot, happy, rows, cols, radius, col_torus, min_similar_host,
                                             happy_count, stream, &handled);
        if (rc != MB_OK || handled) return rc;
    }
    HappyRule rule;
    const bool packed = make_rule(min_similar_host, nmax, &rule);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool vec = packed && binary_types && radius <= 2 && cols % 16 == 0 && rows < (1ll << 30) &&
                     cols < (1ll << 30) && al16(type) && al16(happy) &&
                     (halo_top == nullptr || al16(halo_top)) && (halo_bot == nullptr || al16(halo_bot));
    if (vec) {
        const int nstrips = (int)mb::ceil_div(cols, 512);
#define MB_HAPPY_LAUNCH3(RR, MM, RPW, MINB)                                                                        \
    schelling_happy_vec<RR, MM, RPW, MINB><<<(unsigned)mb::ceil_div((int64_t)nstrips * mb::ceil_div(rows, RPW), kWarps), \
                                             kWarps * 32, 0, stream>>>(type, halo_top, halo_bot, happy, (int)rows,  \
                                                                       (int)cols, col_torus, nstrips, rule, happy_count)
        // rows per warp / CTAs per SM, best of the sweeps on 16384^2: r = 1: 32 rows, 4 CTAs (114 us; 64 rows / 3 CTAs
        // 134 us); r = 2 with the nibble-packed ring: 64 rows, 4 CTAs 143 us (64 / 3: 150, 32 / 3: 151, 32 / 4: 143;
        // the unpacked 16-registers-per-row ring: 128 registers, 2 CTAs, 171 us)
#define MB_HAPPY_LAUNCH(RR, MM)                          \
    do {                                                 \
        if (RR == 1) MB_HAPPY_LAUNCH3(1, MM, 32, 4);     \
        else MB_HAPPY_LAUNCH3(2, MM, 64, 4);             \
    } while (0)
        if (radius == 1) {
            if (rule.mode == MODE_LINEAR) MB_HAPPY_LAUNCH(1, MODE_LINEAR);
            else if (rule.mode == MODE_T) MB_HAPPY_LAUNCH(1, MODE_T);
            else MB_HAPPY_LAUNCH(1, MODE_U);
        } else {
            if (rule.mode == MODE_LINEAR) MB_HAPPY_LAUNCH(2, MODE_LINEAR);
            else if (rule.mode == MODE_T) MB_HAPPY_LAUNCH(2, MODE_T);
            else MB_HAPPY_LAUNCH(2, MODE_U);
        }
#undef MB_HAPPY_LAUNCH3
#undef MB_HAPPY_LAUNCH
    } else {
        const int64_t n = rows * cols;
        schelling_happy_scalar<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(
            type, halo_top, halo_bot, happy, rows, cols, radius, col_torus, rule, happy_count);
    }
    MB_LAUNCH_CHECK("mb_schelling_happy_i8");
    return MB_OK;
}
