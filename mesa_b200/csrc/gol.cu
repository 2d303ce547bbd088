// Conway's Game of Life generation on a uint8 [rows, cols] layer (K1+K2 of SURVEY.md §2.4).
//
// Reproduces, for every cell at once, what the reference does agent by agent:
//   mesa/examples/basic/conways_game_of_life/agents.py:33-51  determine_state
//       live = sum(n.is_alive for n in cell.neighborhood.agents)   (8 Moore neighbours)
//       alive and (live < 2 or live > 3) -> dead ; dead and live == 3 -> alive
//   agents.py:53-55  assume_state   (synchronous two-phase => ping-pong in/out buffers)
//   mesa/discrete_space/grid.py:390-399  torus wrap (`% dim`) or clamped edges.
//
// Layout: C-order [rows, cols], `cols` contiguous (grid.py:146 np.full(dimensions)).
// The row above row 0 and the row below row rows-1 come from `halo_top` / `halo_bot`
// (NULL = dead boundary).  A single-GPU torus passes the opposite edge rows of `in`;
// a row-sharded grid passes the rows received from the neighbouring ranks.
//
// Roofline: HBM-bound, 2 B per cell-update (1 B read + 1 B write).  The fast kernel
// keeps a 3-row rolling window in registers (each row is loaded from HBM once per
// 64-row segment), moves 16 cells per lane per LDG.128/STG.128 and does the rule on
// packed bytes (4 cells per 32-bit ALU op).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;       // warps per CTA
constexpr int kRowsPerWarp = 64;  // rows walked by one warp (2 halo rows re-read per segment)
constexpr int kUnroll = 4;      // rows loaded ahead of the compute (16 B x 4 in flight per lane)

struct Row {
    uint4 v;        // 16 cells of this lane
    unsigned l, r;  // cell left of v / right of v (only meaningful on strip-edge lanes)
};

__device__ __forceinline__ const uint8_t* row_ptr(const uint8_t* __restrict__ in,
                                                  const uint8_t* __restrict__ halo_top,
                                                  const uint8_t* __restrict__ halo_bot, int64_t r,
                                                  int64_t rows, int64_t cols) {
    if (r < 0) return halo_top;
    if (r >= rows) return (r == rows) ? halo_bot : nullptr;
    return in + r * cols;
}

__device__ __forceinline__ Row load_row(const uint8_t* __restrict__ p, int64_t col, bool active,
                                        bool ld_left, bool ld_right, int64_t lcol, int64_t rcol) {
    Row x;
    x.v = make_uint4(0u, 0u, 0u, 0u);
    x.l = 0u;
    x.r = 0u;
    if (p != nullptr) {
        if (active) x.v = __ldg(reinterpret_cast<const uint4*>(p + col));
        if (ld_left) x.l = __ldg(p + lcol);
        if (ld_right) x.r = __ldg(p + rcol);
    }
    return x;
}

// next state of 4 packed cells: n = neighbour count (0..8) per byte, c = current (0/1) per byte.
// alive' = (n == 3) || (c && n == 2)  <=>  (n | c) == 3.
__device__ __forceinline__ unsigned rule4(unsigned n, unsigned c) {
    return __vcmpeq4(n | c, 0x03030303u) & 0x01010101u;
}

__global__ void __launch_bounds__(kWarps * 32)
gol_step_u8_vec(const uint8_t* __restrict__ in, const uint8_t* __restrict__ halo_top,
                const uint8_t* __restrict__ halo_bot, uint8_t* __restrict__ out, int64_t rows,
                int64_t cols, int col_torus, int nstrips) {
    const int lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t strip = task % nstrips;
    const int64_t seg = task / nstrips;
    const int64_t row0 = seg * kRowsPerWarp;
    if (row0 >= rows) return;  // whole warp exits together

    const int64_t col = strip * 512 + lane * 16;
    const bool active = col < cols;
    // Neighbours across the lane boundary come from the adjacent lane (shuffle) except at
    // the strip edges, where one extra byte per row is loaded.
    bool ld_left = active && lane == 0;
    bool ld_right = active && (lane == 31 || col + 16 >= cols);
    int64_t lcol = col - 1, rcol = col + 16;
    if (ld_left && lcol < 0) {
        if (col_torus) lcol = cols - 1; else ld_left = false;
    }
    if (ld_right && rcol >= cols) {
        if (col_torus) rcol = 0; else ld_right = false;
    }
    const bool edge_l = lane == 0;
    const bool edge_r = (lane == 31) || (col + 16 >= cols);

    Row a = load_row(row_ptr(in, halo_top, halo_bot, row0 - 1, rows, cols), col, active, ld_left,
                     ld_right, lcol, rcol);
    Row b = load_row(row_ptr(in, halo_top, halo_bot, row0, rows, cols), col, active, ld_left,
                     ld_right, lcol, rcol);

    const int64_t rend = min(row0 + (int64_t)kRowsPerWarp, rows);
    for (int64_t r = row0; r < rend; r += kUnroll) {
        Row nx[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k)
            nx[k] = load_row(row_ptr(in, halo_top, halo_bot, r + k + 1, rows, cols), col, active,
                             ld_left, ld_right, lcol, rcol);
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            const Row c = nx[k];
            // vertical 3-sums per byte (<= 3)
            uint4 s;
            s.x = a.v.x + b.v.x + c.v.x;
            s.y = a.v.y + b.v.y + c.v.y;
            s.z = a.v.z + b.v.z + c.v.z;
            s.w = a.v.w + b.v.w + c.v.w;
            unsigned sl = __shfl_up_sync(0xffffffffu, s.w >> 24, 1);
            unsigned sr = __shfl_down_sync(0xffffffffu, s.x & 0xffu, 1);
            if (edge_l) sl = a.l + b.l + c.l;
            if (edge_r) sr = a.r + b.r + c.r;
            // horizontal 3-sums: left neighbour of byte k is byte k-1, crossing words
            uint4 h;
            h.x = s.x + ((s.x << 8) | sl) + __funnelshift_r(s.x, s.y, 8);
            h.y = s.y + __funnelshift_l(s.x, s.y, 8) + __funnelshift_r(s.y, s.z, 8);
            h.z = s.z + __funnelshift_l(s.y, s.z, 8) + __funnelshift_r(s.z, s.w, 8);
            h.w = s.w + __funnelshift_l(s.z, s.w, 8) + ((s.w >> 8) | (sr << 24));
            uint4 o;
            o.x = rule4(h.x - b.v.x, b.v.x);
            o.y = rule4(h.y - b.v.y, b.v.y);
            o.z = rule4(h.z - b.v.z, b.v.z);
            o.w = rule4(h.w - b.v.w, b.v.w);
            if (active && r + k < rend) __stcs(reinterpret_cast<uint4*>(out + (r + k) * cols + col), o);
            a = b;
            b = c;
        }
    }
}

// Any shape (cols not a multiple of 16, tiny grids): one thread per cell.
__global__ void gol_step_u8_scalar(const uint8_t* __restrict__ in, const uint8_t* __restrict__ halo_top,
                                   const uint8_t* __restrict__ halo_bot, uint8_t* __restrict__ out,
                                   int64_t rows, int64_t cols, int col_torus) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t i = idx / cols, j = idx % cols;
    int n = 0;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
        const uint8_t* p = row_ptr(in, halo_top, halo_bot, i + di, rows, cols);
        if (p == nullptr) continue;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            if (di == 0 && dj == 0) continue;
            int64_t jj = j + dj;
            if (jj < 0 || jj >= cols) {
                if (!col_torus) continue;
                jj = (jj + cols) % cols;
            }
            n += p[jj];
        }
    }
    const int c = in[idx];
    out[idx] = ((n | c) == 3) ? 1 : 0;
}

}  // namespace

MB_API int mb_gol_step_u8(const uint8_t* in, const uint8_t* halo_top, const uint8_t* halo_bot,
                              uint8_t* out, int64_t rows, int64_t cols, int col_torus,
                              cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0, "mb_gol_step_u8: negative shape");
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(in != nullptr && out != nullptr, "mb_gol_step_u8: null buffer");
    MB_REQUIRE(in != out, "mb_gol_step_u8: in and out must not alias (two-phase update)");
    // A torus narrower than 3 cells aliases neighbours (cell.py:239-241 dedups them); the
    // stencil would double count, so reject instead of silently differing.
    MB_REQUIRE(!(col_torus && cols < 3), "mb_gol_step_u8: torus needs cols >= 3");
    const bool aligned = (cols % 16 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                         (halo_top == nullptr || (reinterpret_cast<uintptr_t>(halo_top) & 15) == 0) &&
                         (halo_bot == nullptr || (reinterpret_cast<uintptr_t>(halo_bot) & 15) == 0);
    if (aligned) {
        const int nstrips = (int)mb::ceil_div(cols, 512);
        const int64_t tasks = (int64_t)nstrips * mb::ceil_div(rows, kRowsPerWarp);
        const int64_t blocks = mb::ceil_div(tasks, kWarps);
        gol_step_u8_vec<<<(unsigned)blocks, kWarps * 32, 0, stream>>>(in, halo_top, halo_bot, out, rows,
                                                                      cols, col_torus, nstrips);
    } else {
        const int64_t n = rows * cols;
        gol_step_u8_scalar<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(in, halo_top, halo_bot, out,
                                                                               rows, cols, col_torus);
    }
    MB_LAUNCH_CHECK("mb_gol_step_u8");
    return MB_OK;
}
