// Schelling happiness on BIT PLANES (the fast path of mb_schelling_happy_i8, csrc/schelling.cu).
//
// Same reference lines as csrc/schelling.cu (mesa/examples/basic/schelling/agents.py:24-42; neighbourhood =
// mesa/discrete_space/cell.py:202-276, Moore radius r = Chebyshev square minus the centre):
//     similar = #{n.type == self.type};  valid = len(neighbors)
//     happy   = not ((similar / valid if valid > 0 else 0.0) < homophily)
//
// The byte-packed kernel (schelling_happy_vec) sits at the integer-ALU roofline (ncu: ALU pipe 75-80 %, ~156 / ~208
// instructions per lane and row of 16 cells for r = 1 / 2), not at the HBM one.  Here a lane owns 32 adjacent cells
// of a row as ONE bit per cell in two planes (type 1 / type 0), so every logic instruction advances 32 cells:
//   * int8 -> bits: 8 cells per multiply -- y = (w0 & 0x01010101) + 16 (w1 & 0x01010101) has the bit of cell (4u + i)
//     at 8i + 4u; y * 0x01020408 moves all eight to bits 24..31 in column order (every cross term lands on a
//     distinct lower bit, no carries); the multiplies run on the FMA pipe, beside the ALU pipe that binds the rest;
//   * horizontal sums of the 2r+1 (and, for the centre row, 2r) cells by funnel shifts with the neighbour lanes'
//     words (SHFL) and carry-save adders on the planes; vertical sum of the 2r+1 rows by carry-save adders:
//     the number of type-1 and of type-0 NEIGHBOURS of every cell as 4 (r = 1) or 5 (r = 2) bit planes;
//   * same = centre type 1 ? n1 : n0, other = the opposite; the float64 test `not (similar / valid < homophily)`
//     is a host-verified integer form  a * same + c >= b * other  (checked against the exact Python-float table
//     for every (similar, valid); e.g. homophily 0.4 -> 3 same >= 2 other), evaluated bit-sliced: shift-and-add
//     multiplications by the (warp-uniform, < 16) constants and one ripple comparator;
//   * bits -> uint8: 4 cells per multiply ((nibble * 0x00204081) & 0x01010101).
// Rules without such a form, layers whose width is not a multiple of 32, radius > 2 and non-binary types stay on
// the byte-packed / scalar kernels.
//
// Roofline: HBM, 2 B per cell (1 B type read + 1 B happy written).
#include "common.cuh"

#include <stdlib.h>

namespace mb {
int happy_bits_launch(const int8_t* type, const int8_t* halo_top, const int8_t* halo_bot, uint8_t* happy, int64_t rows,
                      int64_t cols, int radius, int col_torus, const uint8_t* min_similar_host,
                      unsigned long long* happy_count, cudaStream_t stream, bool* handled);
}

namespace {

constexpr int kWarpsPerCta = 4;
#ifndef MB_HAPPY_MINB1
#define MB_HAPPY_MINB1 5   // CTAs per SM the radius-1 kernel is compiled for
#endif

struct BitRule {
    unsigned a, b, c;        // happy (valid > 0)  <=>  a * same + c >= b * other
    int happy_if_alone;      // valid == 0
};

struct LaneCols {
    unsigned col, lcol, rcol;
    bool active, edge_l, edge_r, ld_left, ld_right;
};

struct RawRow32 {
    uint4 lo, hi;       // the lane's 32 cells
    unsigned lw, rw;    // aligned word left / right of them (strip-edge lanes only)
};

// The lane's 32 cells with ONE 256-bit load (sm_100: LDG.E.256; a warp then asks for 1 KiB of consecutive bytes in 8
// L1 wavefronts -- two 128-bit loads per lane stride the lanes by 32 bytes and take 16), 32-byte aligned rows only
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256_cs(void* p, const uint4& lo, const uint4& hi) {
    asm volatile("st.global.cs.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z),
                 "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}

template <bool kChecked>
__device__ __forceinline__ RawRow32 load_raw32(const int8_t* __restrict__ row, const LaneCols& la) {
    RawRow32 x;
    x.lo = make_uint4(~0u, ~0u, ~0u, ~0u);  // -1 = empty outside the grid
    x.hi = x.lo;
    x.lw = ~0u;
    x.rw = ~0u;
    if (!kChecked || row != nullptr) {
        if (la.active) ldg256(row + la.col, x.lo, x.hi);
        if (la.ld_left) x.lw = __ldg(reinterpret_cast<const unsigned*>(row + la.lcol));
        if (la.ld_right) x.rw = __ldg(reinterpret_cast<const unsigned*>(row + la.rcol));
    }
    return x;
}

__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c) { return (a & b) | (c & (a | b)); }

// One row of a lane after the horizontal pass, per plane p (0: type-1 cells, 1: type-0 cells):
//   cen[p]   the cells themselves
//   full[p]  2r+1-cell horizontal sum (the row above / below a centre), NB bits
//   nbr[p]   the same without the centre cell (the centre row)
template <int R>
struct RowBits {
    static constexpr int NB = R == 1 ? 2 : 3;
    unsigned cen[2];
    unsigned full[2][NB];
    unsigned nbr[2][NB];
};

// int8 cells {-1, 0, 1} -> (type-1 plane, type-0 plane) of the lane's 32 cells, bit k = column col + k
__device__ __forceinline__ void to_planes(const uint4& lo, const uint4& hi, unsigned& one, unsigned& zero) {
    const unsigned w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    unsigned pb[4], pe[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // bit 0 of a byte: 1 for type 1 and for empty (0xff); bit 1: empty only.  The second word of a pair is moved up by
        // four bits, so that cell 4u + i of the pair sits at bit 8i + 4u (+1 for the `empty` bit)
        const unsigned up = w[2 * q + 1] << 4;
        const unsigned yb = (w[2 * q] & 0x01010101u) | (up & 0x10101010u);
        const unsigned ye = (w[2 * q] & 0x02020202u) | (up & 0x20202020u);
        pb[q] = yb * 0x01020408u;   // bit 8i + 4u     -> bit 24 + 4u + i
        pe[q] = ye * 0x00810204u;   // bit 8i + 4u + 1 -> bit 24 + 4u + i
    }
    const unsigned b = __byte_perm(__byte_perm(pb[0], pb[1], 0x0073), __byte_perm(pb[2], pb[3], 0x0073), 0x5410);
    const unsigned e = __byte_perm(__byte_perm(pe[0], pe[1], 0x0073), __byte_perm(pe[2], pe[3], 0x0073), 0x5410);
    one = b & ~e;
    zero = ~(b | e);
}

template <int R>
__device__ __forceinline__ RowBits<R> make_row(const RawRow32& x, const LaneCols& la) {
    RowBits<R> q;
    unsigned pl[2];
    to_planes(x.lo, x.hi, pl[0], pl[1]);
    // the two cells left / right of the lane's word: bits 31, 30 of L / bits 0, 1 of Rr
    unsigned Lp[2], Rp[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        Lp[p] = __shfl_up_sync(0xffffffffu, pl[p], 1);
        Rp[p] = __shfl_down_sync(0xffffffffu, pl[p], 1);
    }
    {   // strip-edge lanes take the R cells beyond either end of the strip from the extra words (branch-free: selects).
        // A cell is type 1 iff bit 0 and not bit 1 of its byte, type 0 iff neither: both tests for all four bytes at once
        const unsigned l1 = x.lw & ~(x.lw >> 1), l0 = ~(x.lw | (x.lw >> 1));   // bit 8i: cell i of the word is type 1 / type 0
        const unsigned r1 = x.rw & ~(x.rw >> 1), r0 = ~(x.rw | (x.rw >> 1));
        unsigned lo1, lo0, ro1, ro0;
        if constexpr (R == 1) {
            lo1 = (l1 & 0x01000000u) << 7;   // cell col - 1 (byte 3) -> bit 31
            lo0 = (l0 & 0x01000000u) << 7;
            ro1 = r1 & 1u;                   // cell col + 32 (byte 0) -> bit 0
            ro0 = r0 & 1u;
        } else {
            // cells col - 1, col - 2 (bytes 3, 2) -> bits 31, 30: one multiply (bit 24 + 7, bit 16 + 14; the stray product
            // bit 23 is masked off); cells col + 32, col + 33 (bytes 0, 1) -> bits 0, 1
            lo1 = ((l1 & 0x01010000u) * 0x4080u) & 0xc0000000u;
            lo0 = ((l0 & 0x01010000u) * 0x4080u) & 0xc0000000u;
            ro1 = (r1 & 1u) | ((r1 >> 7) & 2u);
            ro0 = (r0 & 1u) | ((r0 >> 7) & 2u);
        }
        Lp[0] = la.edge_l ? lo1 : Lp[0];
        Lp[1] = la.edge_l ? lo0 : Lp[1];
        Rp[0] = la.edge_r ? ro1 : Rp[0];
        Rp[1] = la.edge_r ? ro0 : Rp[1];
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const unsigned c = pl[p];
        q.cen[p] = c;
        const unsigned l1 = __funnelshift_l(Lp[p], c, 1), r1 = __funnelshift_r(c, Rp[p], 1);
        if constexpr (R == 1) {
            q.nbr[p][0] = l1 ^ r1;
            q.nbr[p][1] = l1 & r1;
            q.full[p][0] = l1 ^ r1 ^ c;
            q.full[p][1] = maj3(l1, r1, c);
        } else {
            const unsigned l2 = __funnelshift_l(Lp[p], c, 2), r2 = __funnelshift_r(c, Rp[p], 2);
            // l2 + l1 + r1 + r2 (<= 4)
            const unsigned s = l2 ^ l1 ^ r1, k = maj3(l2, l1, r1);
            const unsigned n0 = s ^ r2, k2 = s & r2;
            const unsigned n1 = k ^ k2, n2 = k & k2;
            q.nbr[p][0] = n0; q.nbr[p][1] = n1; q.nbr[p][2] = n2;
            // + centre (<= 5)
            const unsigned t = n0 & c;
            q.full[p][0] = n0 ^ c;
            q.full[p][1] = n1 ^ t;
            q.full[p][2] = n2 | (n1 & t);
        }
    }
    return q;
}

// number of neighbours in plane p around the centre row `m` as NS bit planes
template <int R>
struct NeighbourSum;

template <>
struct NeighbourSum<1> {
    static constexpr int NS = 4;
    // kIncl: the centre cell counts too (3 x 3 box sum)
    template <bool kIncl>
    __device__ __forceinline__ static void run(const RowBits<1>* const* ring, int p, unsigned (&n)[NS]) {
        const unsigned* t = ring[0]->full[p];
        const unsigned* m = kIncl ? ring[1]->full[p] : ring[1]->nbr[p];
        const unsigned* b = ring[2]->full[p];
        const unsigned a0 = t[0] ^ b[0], k0 = t[0] & b[0];
        const unsigned a1 = t[1] ^ b[1] ^ k0, a2 = maj3(t[1], b[1], k0);
        n[0] = a0 ^ m[0];
        const unsigned k = a0 & m[0];
        n[1] = a1 ^ m[1] ^ k;
        const unsigned k1 = maj3(a1, m[1], k);
        n[2] = a2 ^ k1;
        n[3] = a2 & k1;
    }
};

template <>
struct NeighbourSum<2> {
    static constexpr int NS = 5;
    template <bool kIncl>
    __device__ __forceinline__ static void run(const RowBits<2>* const* ring, int p, unsigned (&n)[NS]) {
        const unsigned* t2 = ring[0]->full[p];
        const unsigned* t1 = ring[1]->full[p];
        const unsigned* m = kIncl ? ring[2]->full[p] : ring[2]->nbr[p];
        const unsigned* b1 = ring[3]->full[p];
        const unsigned* b2 = ring[4]->full[p];
        // weight 1
        const unsigned sa = t2[0] ^ t1[0] ^ m[0], ca = maj3(t2[0], t1[0], m[0]);
        n[0] = sa ^ b1[0] ^ b2[0];
        const unsigned cb = maj3(sa, b1[0], b2[0]);
        // weight 2: five bits + two carries
        const unsigned s1 = t2[1] ^ t1[1] ^ m[1], d1 = maj3(t2[1], t1[1], m[1]);
        const unsigned s2 = b1[1] ^ b2[1] ^ ca, d2 = maj3(b1[1], b2[1], ca);
        n[1] = s1 ^ s2 ^ cb;
        const unsigned d3 = maj3(s1, s2, cb);
        // weight 4: five bits + three carries
        const unsigned u1 = t2[2] ^ t1[2] ^ m[2], e1 = maj3(t2[2], t1[2], m[2]);
        const unsigned u2 = b1[2] ^ b2[2] ^ d1, e2 = maj3(b1[2], b2[2], d1);
        const unsigned u3 = u1 ^ u2 ^ d2, e3 = maj3(u1, u2, d2);
        n[2] = u3 ^ d3;
        const unsigned e4 = u3 & d3;
        // weight 8: four carries; weight 16 (the sum is <= 24)
        const unsigned v = e1 ^ e2 ^ e3, f1 = maj3(e1, e2, e3);
        n[3] = v ^ e4;
        n[4] = f1 ^ (v & e4);
    }
};

// acc (NX bit planes) = K * x (NI bit planes), K a compile-time constant: shift-and-add, fully unrolled
template <int K, int NI, int NX>
__device__ __forceinline__ void mul_ct(const unsigned (&x)[NI], unsigned (&acc)[NX]) {
#pragma unroll
    for (int i = 0; i < NX; ++i) acc[i] = 0u;
    bool first = true;   // compile-time after unrolling
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if ((K >> j) & 1) {
            if (first) {
                first = false;
#pragma unroll
                for (int i = 0; i < NI; ++i)
                    if (j + i < NX) acc[j + i] = x[i];
            } else {
                // acc < 2^(NI + j): the add ripples from bit j to bit j + NI
                unsigned carry = 0u;
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    if (j + i < NX) {
                        const unsigned t = acc[j + i];
                        acc[j + i] = t ^ x[i] ^ carry;
                        carry = maj3(t, x[i], carry);
                    }
                }
                if (j + NI < NX) acc[j + NI] = carry;
            }
        }
    }
}

// the same for a warp-uniform runtime constant k < 16: one jump into the specialised straight-line code
template <int NI, int NX>
__device__ __forceinline__ void mul_const(const unsigned (&x)[NI], unsigned k, unsigned (&acc)[NX]) {
    switch (k) {
        case 0: mul_ct<0, NI, NX>(x, acc); break;
        case 1: mul_ct<1, NI, NX>(x, acc); break;
        case 2: mul_ct<2, NI, NX>(x, acc); break;
        case 3: mul_ct<3, NI, NX>(x, acc); break;
        case 4: mul_ct<4, NI, NX>(x, acc); break;
        case 5: mul_ct<5, NI, NX>(x, acc); break;
        case 6: mul_ct<6, NI, NX>(x, acc); break;
        case 7: mul_ct<7, NI, NX>(x, acc); break;
        case 8: mul_ct<8, NI, NX>(x, acc); break;
        case 9: mul_ct<9, NI, NX>(x, acc); break;
        case 10: mul_ct<10, NI, NX>(x, acc); break;
        case 11: mul_ct<11, NI, NX>(x, acc); break;
        case 12: mul_ct<12, NI, NX>(x, acc); break;
        case 13: mul_ct<13, NI, NX>(x, acc); break;
        case 14: mul_ct<14, NI, NX>(x, acc); break;
        default: mul_ct<15, NI, NX>(x, acc); break;
    }
}

// kA, kB >= 0: the rule's constants at compile time (the common homophily values; no jump, no moves); kA < 0: warp-uniform
// runtime constants (any a, b < 16 and the additive constant c)
template <int R, int kA, int kB, int kMinBlocks>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kMinBlocks)
schelling_happy_bits(const int8_t* __restrict__ type, const int8_t* __restrict__ halo_top,
                     const int8_t* __restrict__ halo_bot, uint8_t* __restrict__ happy, int rows, int cols, int col_torus,
                     int nstrips, int rows_per_warp, const BitRule rule, unsigned long long* __restrict__ happy_count) {
    constexpr int W = 2 * R + 1;
    constexpr int NS = NeighbourSum<R>::NS;
    constexpr int NX = NS + 4;   // a * same + c < 16 * 25 + 9 <= 2^(NS + 4)
#ifndef MB_HAPPY_LUT_NIBBLE
    // bits -> uint8: 8 cells per shared-memory lookup (byte of the happy word -> two words of 0 / 1 bytes); the lookups
    // run on the load / store pipe, the arithmetic form (shift, mask, multiply, mask per 4 cells) on the ALU pipe that
    // binds this kernel.  (A per-lane copy of a 16-entry nibble table is conflict-free but needs twice the index
    // arithmetic: measured equal, profiles/r2_happy_bits_*.)
    __shared__ uint2 spread[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        spread[i] = make_uint2((((unsigned)i & 0xfu) * 0x00204081u) & 0x01010101u, (((unsigned)i >> 4) * 0x00204081u) & 0x01010101u);
#else
    __shared__ unsigned spread[16 * 32];
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) spread[i] = (((unsigned)i >> 5) * 0x00204081u) & 0x01010101u;
#endif
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned task = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const unsigned strip = task % (unsigned)nstrips;
    const int row0 = (int)(task / (unsigned)nstrips) * rows_per_warp;
    if (row0 >= rows) return;

    LaneCols la;
    la.col = strip * 1024u + lane * 32u;
    la.active = la.col < (unsigned)cols;
    la.edge_l = lane == 0;
    la.edge_r = (lane == 31) || (la.col + 32u >= (unsigned)cols);
    la.ld_left = la.active && la.edge_l;
    la.ld_right = la.active && la.edge_r;
    la.lcol = la.col - 4u;
    la.rcol = la.col + 32u;
    if (la.ld_left && la.col == 0u) {
        if (col_torus) la.lcol = cols - 4; else la.ld_left = false;
    }
    if (la.ld_right && la.rcol >= (unsigned)cols) {
        if (col_torus) la.rcol = 0u; else la.ld_right = false;
    }
    const size_t pitch = (size_t)cols;
    const int rend = min(row0 + rows_per_warp, rows);

    auto row_ptr = [&](int r) -> const int8_t* {
        if (r < 0) return (halo_top != nullptr && r >= -R) ? halo_top + (size_t)(R + r) * pitch : nullptr;
        if (r >= rows) return (halo_bot != nullptr && r - rows < R) ? halo_bot + (size_t)(r - rows) * pitch : nullptr;
        return type + (size_t)r * pitch;
    };

    RowBits<R> ring[W];   // slot (r - (row0 - R)) % W holds row r
#pragma unroll
    for (int i = 0; i < W - 1; ++i) ring[i] = make_row<R>(load_raw32<true>(row_ptr(row0 - R + i), la), la);

    unsigned nhappy = 0;
    uint8_t* op = happy + (size_t)row0 * pitch + la.col;

    // one output row: the window is ring[(k + i) % W], i = 0 .. 2R (k: position in the group, compile-time)
    auto produce = [&](const RowBits<R>* const* win) {
        // kA == 0 (happy <=> no neighbour of the other type, e.g. homophily 1): `same` is not needed, so the box sums may
        // include the centre cell (it adds to the count of its own type only) -- the centre row then needs no separate
        // centre-less horizontal sum (dead code below), and "has a neighbour" becomes "the two counts add up to >= 2"
        constexpr bool kIncl = kA == 0;
        unsigned n1[NS], n0[NS];
        NeighbourSum<R>::template run<kIncl>(win, 0, n1);
        NeighbourSum<R>::template run<kIncl>(win, 1, n0);
        const unsigned c1 = win[R]->cen[0], c0 = win[R]->cen[1];
        unsigned same[NS], other[NS], any = kIncl ? (n1[0] & n0[0]) : (n1[0] | n0[0]);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            same[i] = (c1 & n1[i]) | (~c1 & n0[i]);
            other[i] = (c1 & n0[i]) | (~c1 & n1[i]);
            if (i > 0) any |= n1[i] | n0[i];
        }
        unsigned X[NX], Y[NX];
        if constexpr (kA >= 0) {
            mul_ct<kA, NS, NX>(same, X);
            mul_ct<kB, NS, NX>(other, Y);
        } else {
            mul_const<NS, NX>(same, rule.a, X);
            mul_const<NS, NX>(other, rule.b, Y);
            if (rule.c != 0u) {   // uniform
                unsigned carry = 0u;
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const unsigned m = ((rule.c >> i) & 1u) ? ~0u : 0u;
                    const unsigned t = X[i];
                    X[i] = t ^ m ^ carry;
                    carry = maj3(t, m, carry);
                }
            }
        }
        unsigned ge = ~0u;
#pragma unroll
        for (int i = 0; i < NX; ++i) ge = (X[i] & ~Y[i]) | (~(X[i] ^ Y[i]) & ge);
        const unsigned alone = rule.happy_if_alone ? ~0u : 0u;
        const unsigned H = (c1 | c0) & ((any & ge) | (~any & alone));
        nhappy += __popc(H);
        uint4 o0, o1;
#ifndef MB_HAPPY_LUT_NIBBLE
        const uint2 q0 = spread[H & 0xffu], q1 = spread[__byte_perm(H, 0u, 0x4441)];
        const uint2 q2 = spread[__byte_perm(H, 0u, 0x4442)], q3 = spread[H >> 24];
        o0 = make_uint4(q0.x, q0.y, q1.x, q1.y);
        o1 = make_uint4(q2.x, q2.y, q3.x, q3.y);
#else
        // nibble 2j of H is byte j of Ha, nibble 2j + 1 byte j of Hb; address = lane + 32 * nibble
        const unsigned Ha = H & 0x0f0f0f0fu, Hb = (H >> 4) & 0x0f0f0f0fu;
        const unsigned* tab = spread + lane;
        o0.x = tab[__byte_perm(Ha, 0u, 0x4440) * 32u]; o0.y = tab[__byte_perm(Hb, 0u, 0x4440) * 32u];
        o0.z = tab[__byte_perm(Ha, 0u, 0x4441) * 32u]; o0.w = tab[__byte_perm(Hb, 0u, 0x4441) * 32u];
        o1.x = tab[__byte_perm(Ha, 0u, 0x4442) * 32u]; o1.y = tab[__byte_perm(Hb, 0u, 0x4442) * 32u];
        o1.z = tab[__byte_perm(Ha, 0u, 0x4443) * 32u]; o1.w = tab[__byte_perm(Hb, 0u, 0x4443) * 32u];
#endif
        if (la.active) stg256_cs(op, o0, o1);
        op += pitch;
    };

    int r = row0;
    // Whole groups of W rows whose look-ahead rows all exist: unchecked loads, no row tests.  The raw rows are PREFETCHED one
    // group ahead, slot by slot: as soon as raw[k] has been turned into bit planes, the same registers receive row k of the
    // next group, so every load has a whole group of arithmetic (~150 instructions per row) to arrive and W rows
    // (2W 16-byte loads per lane) are in flight per warp all the time.
    {
        RawRow32 raw[W];
        bool have = r + W <= rend && r + R + W <= rows;
        if (have) {
            const int8_t* fp = type + (size_t)(r + R) * pitch;
#pragma unroll
            for (int k = 0; k < W; ++k) raw[k] = load_raw32<false>(fp + (size_t)k * pitch, la);
        }
        while (have) {
            const bool next = r + 2 * W <= rend && r + R + 2 * W <= rows;   // warp-uniform
            const int8_t* fp = type + (size_t)(r + R + W) * pitch;
#pragma unroll
            for (int k = 0; k < W; ++k) {
                ring[(k + W - 1) % W] = make_row<R>(raw[k], la);
                if (next) raw[k] = load_raw32<false>(fp + (size_t)k * pitch, la);
                const RowBits<R>* win[W];
#pragma unroll
                for (int i = 0; i < W; ++i) win[i] = &ring[(k + i) % W];
                produce(win);
            }
            r += W;
            have = next;
        }
    }
    // the last rows of the block / of the grid: checked row pointers, row tests
    for (; r < rend; r += W) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            if (r + k < rend) {
                ring[(k + W - 1) % W] = make_row<R>(load_raw32<true>(row_ptr(r + R + k), la), la);
                const RowBits<R>* win[W];
#pragma unroll
                for (int i = 0; i < W; ++i) win[i] = &ring[(k + i) % W];
                produce(win);
            }
        }
    }
    nhappy = warp_reduce_add(la.active ? nhappy : 0u);
    if (lane == 0 && nhappy != 0 && happy_count != nullptr) atomicAdd(happy_count, (unsigned long long)nhappy);
}

// happy (valid >= 1)  <=>  similar >= tbl[valid]  <=>  a * similar + c >= b * (valid - similar), every (similar, valid) checked
bool make_bit_rule(const uint8_t* tbl, int nmax, BitRule* rule) {
    rule->happy_if_alone = tbl[0] == 0;
    for (int total = 0; total <= 30; ++total)          // small coefficients first (fewer shift-and-add terms)
        for (int a = 0; a <= 15 && a <= total; ++a) {
            const int b = total - a;
            if (b > 15) continue;
            for (int c = 0; c <= 8; ++c) {
                bool ok = true;
                for (int v = 1; v <= nmax && ok; ++v)
                    for (int s = 0; s <= v; ++s)
                        if ((a * s + c >= b * (v - s)) != (s >= (int)tbl[v])) { ok = false; break; }
                if (ok) {
                    rule->a = (unsigned)a; rule->b = (unsigned)b; rule->c = (unsigned)c;
                    return true;
                }
            }
        }
    return false;
}

}  // namespace

namespace mb {

// *handled = false: the caller falls back to the byte-packed / scalar kernels
int happy_bits_launch(const int8_t* type, const int8_t* halo_top, const int8_t* halo_bot, uint8_t* happy, int64_t rows,
                      int64_t cols, int radius, int col_torus, const uint8_t* min_similar_host,
                      unsigned long long* happy_count, cudaStream_t stream, bool* handled) {
    *handled = false;
    static int enabled = -1, rpw_env = 0;
    if (enabled < 0) {
        const char* e = getenv("MB_HAPPY_BITS");
        enabled = e ? atoi(e) : 1;
        const char* r = getenv("MB_HAPPY_BITS_ROWS");
        rpw_env = r ? atoi(r) : 0;
    }
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };   // 256-bit loads / stores
    if (!enabled || radius > 2 || cols % 32 != 0 || rows >= (1ll << 30) || cols >= (1ll << 30) || !al32(type) || !al32(happy) ||
        (halo_top != nullptr && !al32(halo_top)) || (halo_bot != nullptr && !al32(halo_bot)))
        return MB_OK;
    const int nmax = (2 * radius + 1) * (2 * radius + 1) - 1;
    BitRule rule;
    if (!make_bit_rule(min_similar_host, nmax, &rule)) return MB_OK;
    const int nstrips = (int)ceil_div(cols, 1024);
    const int W = 2 * radius + 1;
    // rows per warp: a multiple of the group size; ~8 warps per SM and wave at the least, at most 32 (r = 1) / 60 rows
    int rpw = rpw_env > 0 ? rpw_env : (radius == 1 ? 33 : 60);
    const int64_t want = (int64_t)sm_count() * 64;
    while (rpw_env <= 0 && rpw > 2 * W && (int64_t)nstrips * ceil_div(rows, rpw) < want) rpw -= W;
    rpw = (int)ceil_div(rpw, W) * W;
    const unsigned blocks = (unsigned)ceil_div((int64_t)nstrips * ceil_div(rows, rpw), kWarpsPerCta);
#define MB_BITS_LAUNCH(RR, AA, BB, MINB)                                                                                  \
    schelling_happy_bits<RR, AA, BB, MINB><<<blocks, kWarpsPerCta * 32, 0, stream>>>(                                   \
        type, halo_top, halo_bot, happy, (int)rows, (int)cols, col_torus, nstrips, rpw, rule, happy_count)
#define MB_BITS_CASE(AA, BB)                                        \
    if (!done && rule.c == 0u && rule.a == AA && rule.b == BB) {    \
        if (radius == 1) MB_BITS_LAUNCH(1, AA, BB, MB_HAPPY_MINB1); \
        else MB_BITS_LAUNCH(2, AA, BB, (AA == 0 ? 4 : 3));           \
        done = true;                                                \
    }
    bool done = false;
    // homophily 1 (0 >= other), 1/2, 2/5 (3 same >= 2 other), 1/3, 2/3, 1/4, 3/4, 3/5, 3/10, 7/10: compiled-in constants
    MB_BITS_CASE(0, 1) MB_BITS_CASE(1, 1) MB_BITS_CASE(3, 2) MB_BITS_CASE(2, 1) MB_BITS_CASE(1, 2)
    MB_BITS_CASE(3, 1) MB_BITS_CASE(1, 3) MB_BITS_CASE(2, 3) MB_BITS_CASE(7, 3) MB_BITS_CASE(3, 7)
    if (!done) {
        if (radius == 1) MB_BITS_LAUNCH(1, -1, -1, MB_HAPPY_MINB1);
        else MB_BITS_LAUNCH(2, -1, -1, 3);
    }
#undef MB_BITS_CASE
#undef MB_BITS_LAUNCH
    MB_LAUNCH_CHECK("mb_schelling_happy_i8 (bit planes)");
    *handled = true;
    return MB_OK;
}

}  // namespace mb
