// Game of Life on a BIT-PACKED layer, several generations per launch (temporal blocking).
//
// Same rule and the same reference lines as gol.cu (conways_game_of_life/agents.py:33-55, torus wrap
// grid.py:390-399), different storage: while a model only steps, the uint8 `state` layer is kept as one
// bit per cell -- word w of row i holds cells (i, 32w .. 32w+31), bit k = column 32w+k -- and is unpacked
// only when somebody reads the layer.  A 16384^2 layer is 32 MiB instead of 256 MiB: both ping-pong
// buffers live in the 126 MB L2, so a generation no longer costs 2 B/cell of HBM traffic and the kernel is
// bound by integer issue rate instead (about 20 LOP3/SHF per 32 cells and generation).
//
// One warp owns a strip of 30 words (960 cells; lanes 0 and 31 carry one halo word each) and walks down a
// segment of rows.  The row sums are bit-sliced: per row the horizontal 3-sum of every cell is two bit planes
// (h0, h1), the 9-cell total is a carry-save add of three such rows, and
//     alive' = (total == 3) | (alive & total == 4)            (total counts the cell itself)
// is a handful of 3-input logic ops for 32 cells at once.
//
// Temporal blocking: T generations are chained through registers as a wavefront -- when row j of generation 0
// arrives, stage g (0 <= g < T) turns row j-g of generation g into row j-g-1 of generation g+1 -- so one pass
// over the layer (one load and one store per word) advances T generations.  Each stage keeps the (h0, h1)
// planes of its two previous rows and the centre word of the previous row: 6 registers per stage and lane.
// A segment re-computes T rows above and below itself, and the halo lanes absorb the (at most T <= 8 bits
// of) garbage creeping in from outside the strip.
#include "common.cuh"

#include <stdlib.h>

namespace {

constexpr int kBitWarps = 4;          // warps per CTA (independent of each other)
constexpr int kStripWords = 30;       // useful words per warp; lanes 0 / 31 are halo
constexpr int kMaxGenerations = 8;    // generations per launch (template instances 1..8)

struct BitRows {
    const uint32_t* in;
    const uint32_t* halo_top;   // [T, wpr] rows -T .. -1 (NULL: dead)
    const uint32_t* halo_bot;   // [T, wpr] rows rows .. rows+T-1 (NULL: dead)
    int rows, wpr;
};

// word `col` of (virtual) row j of the input generation
template <int T>
__device__ __forceinline__ uint32_t fetch_word(const BitRows& g, int j, int col, bool colvalid) {
    const uint32_t* p;
    if (j >= 0 && j < g.rows) p = g.in + (size_t)j * g.wpr;
    else if (j < 0) p = (g.halo_top != nullptr && j >= -T) ? g.halo_top + (size_t)(T + j) * g.wpr : nullptr;
    else p = (g.halo_bot != nullptr && j - g.rows < T) ? g.halo_bot + (size_t)(j - g.rows) * g.wpr : nullptr;
    if (p == nullptr || !colvalid) return 0u;
    return __ldg(p + col);
}

// one 3-input logic op (LOP3.LUT); kLut = the function evaluated on a = 0xF0, b = 0xCC, c = 0xAA
template <int kLut>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(kLut));
    return d;
}
constexpr int kXor3 = 0x96, kMaj = 0xE8;

// horizontal 3-sums of one packed row: bit k of (h0, h1) = cells k-1, k, k+1 summed (two bit planes)
__device__ __forceinline__ void row_sums(uint32_t x, uint32_t& h0, uint32_t& h1) {
    const uint32_t lw = __shfl_up_sync(0xffffffffu, x, 1), rw = __shfl_down_sync(0xffffffffu, x, 1);
    const uint32_t l = __funnelshift_l(lw, x, 1);   // (x << 1) | (lw >> 31): the cell to the left of each cell
    const uint32_t r = __funnelshift_r(x, rw, 1);   // (x >> 1) | (rw << 31): the cell to the right
    h0 = lop3<kXor3>(l, x, r);
    h1 = lop3<kMaj>(l, x, r);
}

// next state of the centre row from the (h0, h1) planes of the rows above / centre / below and the centre word:
// total = t0 + 2 (k0 + u0 + 2 u1) counts the cell itself; alive' = total == 3 | (alive & total == 4).  8 LOP3.
__device__ __forceinline__ uint32_t life_rule(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t c0,
                                              uint32_t c1, uint32_t centre) {
    const uint32_t t0 = lop3<kXor3>(a0, b0, c0);   // total bit 0
    const uint32_t k0 = lop3<kMaj>(a0, b0, c0);    // carry of the low planes   (weight 2)
    const uint32_t u0 = lop3<kXor3>(a1, b1, c1);   // sum of the high planes    (weight 2)
    const uint32_t u1 = lop3<kMaj>(a1, b1, c1);    // carry of the high planes  (weight 4)
    const uint32_t is1 = lop3<0x14>(k0, u0, u1);   // k0 + u0 + 2 u1 == 1: (k0 ^ u0) & ~u1          -> total 2 or 3
    const uint32_t is2 = lop3<0x42>(k0, u0, u1);   // == 2: (k0 & u0 & ~u1) | (~k0 & ~u0 & u1)      -> total 4 or 5
    const uint32_t four = lop3<0x08>(t0, centre, is2);  // ~t0 & centre & is2: alive and total == 4
    return lop3<0xEA>(t0, is1, four);                   // (t0 & is1) | four
}

// One input row (row j of generation 0) through all T stages: stage s turns row j-s of generation s into row
// j-s-1 of generation s+1.  Slot kOlder of every stage holds the planes of row-2, the other slot those of row-1;
// the caller alternates kOlder, so the window slides without register moves.  Returns row j-T of generation T.
// kRowCheck: rows outside a clamped grid are forced dead at every stage (segments that touch the grid edge);
// kColMask: lanes outside a clamped grid are forced dead (edge strips).
template <int T, bool kRowCheck, bool kColMask, int kOlder>
__device__ __forceinline__ uint32_t advance_row(uint32_t x, int j, uint32_t (&h0)[T][2], uint32_t (&h1)[T][2],
                                                uint32_t (&cw)[T][2], int rows, bool top_dead, bool bot_dead,
                                                uint32_t cmask) {
    constexpr int kNewer = 1 - kOlder;
#pragma unroll
    for (int s = 0; s < T; ++s) {
        uint32_t c0, c1;
        row_sums(x, c0, c1);
        uint32_t y = life_rule(h0[s][kOlder], h1[s][kOlder], h0[s][kNewer], h1[s][kNewer], c0, c1, cw[s][kNewer]);
        if (kRowCheck) {
            const int jj = j - s - 1;  // row index of y
            if ((jj < 0 && top_dead) || (jj >= rows && bot_dead)) y = 0u;
        }
        if (kColMask) y &= cmask;
        h0[s][kOlder] = c0;  // the slot of row-2 is free now: it becomes the new row-1
        h1[s][kOlder] = c1;
        cw[s][kOlder] = x;
        x = y;
    }
    return x;
}

// The T-generation wavefront over the input rows jstart .. jstart+nsteps-1 of one strip (nsteps = segment rows
// + 2T); the first 2T steps only fill the pipeline.  kFast: every row touched (including the two prefetched past
// the end) lies inside the block, so rows are addressed by pointer increments; otherwise fetch_word resolves
// halos / dead rows per row.
template <int T, bool kFast, bool kRowCheck, bool kColMask>
__device__ __forceinline__ void run_segment(const BitRows& g, uint32_t* __restrict__ op, int jstart, int nsteps,
                                            int col, bool colvalid, bool stores, uint32_t cmask) {
    const bool top_dead = g.halo_top == nullptr, bot_dead = g.halo_bot == nullptr;
    uint32_t h0[T][2], h1[T][2], cw[T][2];
#pragma unroll
    for (int s = 0; s < T; ++s) {
        h0[s][0] = h0[s][1] = h1[s][0] = h1[s][1] = cw[s][0] = cw[s][1] = 0u;
    }
    const size_t pitch = (size_t)g.wpr;
    const uint32_t* ip = g.in + (ptrdiff_t)jstart * (ptrdiff_t)pitch + col;  // only dereferenced when kFast
    auto fetch = [&](int j) -> uint32_t {
        if (kFast) return (!kColMask || colvalid) ? __ldg(ip + (size_t)(j - jstart) * pitch) : 0u;
        return fetch_word<T>(g, j, col, colvalid);
    };
    // two rows per iteration, loaded one iteration before they are used
    uint32_t n0 = fetch(jstart), n1 = fetch(jstart + 1);
    int j = jstart;
    auto next_rows = [&](uint32_t& x0, uint32_t& x1) {
        x0 = n0;
        x1 = n1;
        n0 = fetch(j + 2);   // rows past the last needed one are never used for a stored word
        n1 = fetch(j + 3);
    };
    // fill: 2T rows, nothing to store yet
#pragma unroll 1
    for (int s = 0; s < 2 * T; s += 2, j += 2) {
        uint32_t x0, x1;
        next_rows(x0, x1);
        advance_row<T, kRowCheck, kColMask, 0>(x0, j, h0, h1, cw, g.rows, top_dead, bot_dead, cmask);
        advance_row<T, kRowCheck, kColMask, 1>(x1, j + 1, h0, h1, cw, g.rows, top_dead, bot_dead, cmask);
    }
    // steady state: every input row completes one output row
    const int jend = jstart + nsteps;
#pragma unroll 1
    for (; j + 1 < jend; j += 2) {
        uint32_t x0, x1;
        next_rows(x0, x1);
        const uint32_t y0 = advance_row<T, kRowCheck, kColMask, 0>(x0, j, h0, h1, cw, g.rows, top_dead, bot_dead, cmask);
        if (stores) op[0] = y0;
        const uint32_t y1 = advance_row<T, kRowCheck, kColMask, 1>(x1, j + 1, h0, h1, cw, g.rows, top_dead, bot_dead, cmask);
        if (stores) op[pitch] = y1;
        op += 2 * pitch;
    }
    if (j < jend) {  // odd number of rows in the segment
        const uint32_t y0 = advance_row<T, kRowCheck, kColMask, 0>(n0, j, h0, h1, cw, g.rows, top_dead, bot_dead, cmask);
        if (stores) op[0] = y0;
    }
}

// T generations of the rows [seg*rps, seg*rps + rps) of one 30-word strip.
template <int T>
__global__ void __launch_bounds__(kBitWarps * 32)
gol_bits_kernel(const BitRows g, uint32_t* __restrict__ out, int col_torus, int nstrips, int rps) {
    const int lane = threadIdx.x & 31;
    const unsigned task = blockIdx.x * kBitWarps + (threadIdx.x >> 5);
    const unsigned strip = task % (unsigned)nstrips, seg = task / (unsigned)nstrips;
    const int r0 = (int)seg * rps;
    if (r0 >= g.rows) return;  // whole warp
    const int r1 = min(r0 + rps, g.rows);

    const int rawcol = (int)strip * kStripWords + lane - 1;
    int col = rawcol;
    bool colvalid = rawcol >= 0 && rawcol < g.wpr;
    if (col_torus) {
        col = rawcol % g.wpr;
        if (col < 0) col += g.wpr;
        colvalid = true;
    }
    if (!colvalid) col = 0;
    const bool stores = lane >= 1 && lane <= kStripWords && rawcol < g.wpr;
    const uint32_t cmask = colvalid ? 0xffffffffu : 0u;
    const bool masked = __any_sync(0xffffffffu, !colvalid);  // an edge strip of a column-clamped grid

    const int jstart = r0 - T;
    const int nsteps = (r1 - r0) + 2 * T;  // input rows r0-T .. r1+T-1
    uint32_t* op = out + (size_t)r0 * g.wpr + (stores ? rawcol : 0);
    // interior segment: all rows it reads (and the two it prefetches past its end) are rows of this block; rows
    // outside the grid then lie outside the cone of the rows it stores, so no row needs to be forced dead
    const bool interior = jstart >= 0 && jstart + nsteps + 3 <= g.rows;
    if (interior) {
        if (masked) run_segment<T, true, false, true>(g, op, jstart, nsteps, col, colvalid, stores, cmask);
        else run_segment<T, true, false, false>(g, op, jstart, nsteps, col, colvalid, stores, cmask);
    } else {
        run_segment<T, false, true, true>(g, op, jstart, nsteps, col, colvalid, stores, cmask);
    }
}

// uint8 0/1 layer -> bit planes, 32 cells per thread
__global__ void gol_pack_kernel(const uint8_t* __restrict__ state, uint32_t* __restrict__ bits, int64_t nwords) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    const uint4* p = reinterpret_cast<const uint4*>(state + w * 32);
    const uint4 a = ldg_nc_v4(p), b = ldg_nc_v4(p + 1);
    // 4 bytes (0/1) -> 4 bits: the multiplier moves byte k's bit 0 to bit 24+k without carries
    auto nib = [](uint32_t v) { return ((v & 0x01010101u) * 0x01020408u) >> 24; };
    bits[w] = nib(a.x) | (nib(a.y) << 4) | (nib(a.z) << 8) | (nib(a.w) << 12) | (nib(b.x) << 16) |
              (nib(b.y) << 20) | (nib(b.z) << 24) | (nib(b.w) << 28);
}

__global__ void gol_unpack_kernel(const uint32_t* __restrict__ bits, uint8_t* __restrict__ state, int64_t nwords) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    const uint32_t v = bits[w];
    // 4 bits -> 4 bytes: bit k lands on bit 8k (copies at distance 7 never collide)
    auto spread = [](uint32_t n) { return ((n & 0xfu) * 0x00204081u) & 0x01010101u; };
    uint4 a, b;
    a.x = spread(v); a.y = spread(v >> 4); a.z = spread(v >> 8); a.w = spread(v >> 12);
    b.x = spread(v >> 16); b.y = spread(v >> 20); b.z = spread(v >> 24); b.w = spread(v >> 28);
    uint4* p = reinterpret_cast<uint4*>(state + w * 32);
    stg_na_v4(p, a);
    stg_na_v4(p + 1, b);
}

// One template instance per number of fused generations.  (Developer variants that were swept and dropped: deeper
// row prefetch, 2 / 8 warps per CTA, register caps, shifts on the FMA pipe -- profiles/r1_gol_bits_sweep*.jsonl.)
int launch_bits(int T, const BitRows& g, uint32_t* out, int col_torus, int nstrips, int rps, int64_t tasks,
                cudaStream_t stream) {
    const unsigned blocks = (unsigned)mb::ceil_div(tasks, kBitWarps);
#define MB_BITS_CASE(TT)                                                                                  \
    case TT:                                                                                              \
        gol_bits_kernel<TT><<<blocks, kBitWarps * 32, 0, stream>>>(g, out, col_torus, nstrips, rps);     \
        break;
    switch (T) {
        MB_BITS_CASE(1) MB_BITS_CASE(2) MB_BITS_CASE(3) MB_BITS_CASE(4)
        MB_BITS_CASE(5) MB_BITS_CASE(6) MB_BITS_CASE(7) MB_BITS_CASE(8)
        default: mb::set_error("mb_gol_step_bits: generations must be 1..%d", kMaxGenerations); return MB_ERR_INVALID;
    }
#undef MB_BITS_CASE
    return MB_OK;
}

}  // namespace

MB_API int mb_gol_pack_u8(const uint8_t* state, uint32_t* bits, int64_t rows, int64_t cols, cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0 && cols % 32 == 0, "mb_gol_pack_u8: cols must be a multiple of 32");
    const int64_t nwords = rows * (cols / 32);
    if (nwords == 0) return MB_OK;
    MB_REQUIRE(state != nullptr && bits != nullptr, "mb_gol_pack_u8: null buffer");
    MB_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0, "mb_gol_pack_u8: state must be 16-byte aligned");
    gol_pack_kernel<<<(unsigned)mb::ceil_div(nwords, 256), 256, 0, stream>>>(state, bits, nwords);
    MB_LAUNCH_CHECK("mb_gol_pack_u8");
    return MB_OK;
}

MB_API int mb_gol_unpack_u8(const uint32_t* bits, uint8_t* state, int64_t rows, int64_t cols, cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0 && cols % 32 == 0, "mb_gol_unpack_u8: cols must be a multiple of 32");
    const int64_t nwords = rows * (cols / 32);
    if (nwords == 0) return MB_OK;
    MB_REQUIRE(state != nullptr && bits != nullptr, "mb_gol_unpack_u8: null buffer");
    MB_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0, "mb_gol_unpack_u8: state must be 16-byte aligned");
    gol_unpack_kernel<<<(unsigned)mb::ceil_div(nwords, 256), 256, 0, stream>>>(bits, state, nwords);
    MB_LAUNCH_CHECK("mb_gol_unpack_u8");
    return MB_OK;
}

MB_API int mb_gol_step_bits(const uint32_t* in, const uint32_t* halo_top, const uint32_t* halo_bot, uint32_t* out,
                            int64_t rows, int64_t cols, int col_torus, int generations, int rows_per_segment,
                            cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0 && cols % 32 == 0, "mb_gol_step_bits: cols must be a multiple of 32");
    MB_REQUIRE(generations >= 1 && generations <= kMaxGenerations, "mb_gol_step_bits: generations must be 1..%d",
               kMaxGenerations);
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(in != nullptr && out != nullptr && in != out, "mb_gol_step_bits: in / out must be distinct buffers");
    MB_REQUIRE(rows < (1ll << 30) && cols < (1ll << 35), "mb_gol_step_bits: layer too large");
    // same restriction as mb_gol_step_u8: a torus narrower than 3 cells aliases neighbours (cell.py:239-241)
    MB_REQUIRE(!(col_torus && cols < 3), "mb_gol_step_bits: torus needs cols >= 3");
    const int wpr = (int)(cols / 32);
    const int nstrips = (int)mb::ceil_div(wpr, kStripWords);
    if (rows_per_segment <= 0) {
        static const int env_rps = getenv("MB_GOL_BITS_RPS") ? atoi(getenv("MB_GOL_BITS_RPS")) : 0;
        rows_per_segment = env_rps;
    }
    if (rows_per_segment <= 0) {
        // Segment height by a small cost model (fitted to the sweeps in profiles/): every SM executes k = ceil(CTAs / SMs)
        // CTAs one after the other (the kernel is bound by the SM's ALU pipe) and a CTA costs rows + 2T row steps, so the
        // launch costs ~ (rows_per_segment + 2T) * k.  The load-balance quantisation matters more than the 2T recomputed
        // rows: 16384 rows at 72 rows per segment are 1026 CTAs = 6.93 per SM (9.6 us per generation), 16640 rows would be
        // 1044 = 7.05 per SM, i.e. 8 on some SMs (11.9 us).  Few CTAs per SM leave no slack for dynamic balance (measured: 7 per SM
        // beats 6 and 5 at equal model cost), hence the small penalties.
        const int sms = mb::sm_count();
        double best = 1e30;
        for (int rps = 32; rps <= 160; rps += 2) {
            const int64_t ctas = mb::ceil_div((int64_t)nstrips * mb::ceil_div(rows, rps), kBitWarps);
            const int64_t k = mb::ceil_div(ctas, sms);
            // an SM with fewer than 5 CTAs (its residency at T = 8) does not run them any faster: count at least 5
            const double cost = (double)(rps + 2 * generations) * (double)(k < 5 ? 5 : k) * (k < 6 ? 1.15 : (k == 6 ? 1.04 : 1.0));
            if (cost < best) {
                best = cost;
                rows_per_segment = rps;
            }
        }
    }
    rows_per_segment = (rows_per_segment + 1) & ~1;
    const int64_t tasks = (int64_t)nstrips * mb::ceil_div(rows, rows_per_segment);
    const BitRows g{in, halo_top, halo_bot, (int)rows, wpr};
    const int rc = launch_bits(generations, g, out, col_torus, nstrips, rows_per_segment, tasks, stream);
    if (rc != MB_OK) return rc;
    MB_LAUNCH_CHECK("mb_gol_step_bits");
    return MB_OK;
}
