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
// 16-row segment; the two re-read halo rows hit L2), moves 16 cells per lane per LDG.128/STG.128
// and does the rule on packed bytes (4 cells per 32-bit ALU op).
#include "common.cuh"

#include <stdlib.h>

namespace {

constexpr int kWarps = 8;       // warps per CTA
// Tunables (template parameters of the vector kernel):
//   kRowsPerWarp  rows walked by one warp (2 halo rows re-read per segment)
//   kUnroll       rows loaded ahead of the compute (16 B x kUnroll in flight per lane)
//   kMinBlocks    CTAs per SM the register budget is capped for

__device__ __forceinline__ const uint8_t* row_ptr(const uint8_t* __restrict__ in,
                                                  const uint8_t* __restrict__ halo_top,
                                                  const uint8_t* __restrict__ halo_bot, int64_t r,
                                                  int64_t rows, int64_t cols) {
    if (r < 0) return halo_top;
    if (r >= rows) return (r == rows) ? halo_bot : nullptr;
    return in + r * cols;
}

// 16 cells of one lane plus the cell left / right of them (edge lanes only).
struct Row {
    uint4 v;
    unsigned l, r;
};

// is-equal-to-3 on packed bytes whose values are < 0x80: 1 per byte where (n | c) == 3.
// alive' = (n == 3) || (c && n == 2)  <=>  (n | c) == 3   (n = 0..8 neighbours, c = 0/1).
__device__ __forceinline__ unsigned rule4(unsigned n, unsigned c) {
    const unsigned x = (n | c) ^ 0x03030303u;             // 0 where equal
    return (~((x + 0x7f7f7f7fu) >> 7)) & 0x01010101u;     // bit7 of x+0x7f is set iff x != 0
}

// Per-lane addressing of one strip: the lane's 16 cells and, on strip-edge lanes, the byte left /
// right of them (torus wrap or absent).
struct LaneAddr {
    unsigned col, lcol, rcol;
    bool active, ld_left, ld_right;
};

__device__ __forceinline__ Row load_row(const uint8_t* __restrict__ row, const LaneAddr& la) {
    Row x;
    x.v = make_uint4(0u, 0u, 0u, 0u);
    x.l = 0u;
    x.r = 0u;
    if (la.active) x.v = __ldg(reinterpret_cast<const uint4*>(row + la.col));
    if (la.ld_left) x.l = __ldg(row + la.lcol);
    if (la.ld_right) x.r = __ldg(row + la.rcol);
    return x;
}

__device__ __forceinline__ Row load_row_checked(const uint8_t* __restrict__ row, const LaneAddr& la) {
    if (row == nullptr) {
        Row x;
        x.v = make_uint4(0u, 0u, 0u, 0u);
        x.l = 0u;
        x.r = 0u;
        return x;
    }
    return load_row(row, la);
}

// next generation of the 16 cells of row b given the rows above (a) and below (c)
__device__ __forceinline__ uint4 gol_rows(const Row& a, const Row& b, const Row& c, bool edge_l, bool edge_r) {
    // vertical 3-sums per byte (<= 3)
    uint4 s;
    s.x = a.v.x + b.v.x + c.v.x;
    s.y = a.v.y + b.v.y + c.v.y;
    s.z = a.v.z + b.v.z + c.v.z;
    s.w = a.v.w + b.v.w + c.v.w;
    unsigned sl = __shfl_up_sync(0xffffffffu, s.w, 1) >> 24;
    unsigned sr = __shfl_down_sync(0xffffffffu, s.x, 1) << 24;
    if (edge_l) sl = a.l + b.l + c.l;
    if (edge_r) sr = (a.r + b.r + c.r) << 24;
    // horizontal 3-sums: left neighbour of byte k is byte k-1, crossing words
    uint4 h;
    h.x = s.x + ((s.x << 8) | sl) + __funnelshift_r(s.x, s.y, 8);
    h.y = s.y + __funnelshift_l(s.x, s.y, 8) + __funnelshift_r(s.y, s.z, 8);
    h.z = s.z + __funnelshift_l(s.y, s.z, 8) + __funnelshift_r(s.z, s.w, 8);
    h.w = s.w + __funnelshift_l(s.z, s.w, 8) + ((s.w >> 8) | sr);
    uint4 o;
    o.x = rule4(h.x - b.v.x, b.v.x);
    o.y = rule4(h.y - b.v.y, b.v.y);
    o.z = rule4(h.z - b.v.z, b.v.z);
    o.w = rule4(h.w - b.v.w, b.v.w);
    return o;
}

// ---- fused halo synchronisation for row-sharded grids (one process per GPU) ------------------------
// The neighbours' edge rows are read IN PLACE from their HBM over NVLink (halo_top / halo_bot are
// peer-mapped pointers) and the hand-shake that a separate exchange + barrier would do is done by
// the stencil kernel itself with generation counters in peer-mapped memory:
//   * the warps that own the first / last row segment wait (ld.acquire.sys) until the neighbour has
//     published generation gen-1 of its edge rows, and only then touch the halo;
//   * when all edge strips of the new generation are written, the last one publishes `gen` to the
//     neighbour (threadfence_system + st.release.sys into ITS memory).
// Edge segments are mapped to the lowest CTA indices, so the flags are published at the very start
// of a launch and in steady state nobody waits: one launch per generation, no NCCL call, no barrier.
// The same counters order the write-after-read on the ping-pong buffers (see PeerSync::gen).
struct PeerSync {
    unsigned* wait_top;     // local: generation of the upper neighbour's edge rows that is complete
    unsigned* wait_bot;     // local: same for the lower neighbour
    unsigned* signal_up;    // peer-mapped: the upper neighbour's wait_bot
    unsigned* signal_down;  // peer-mapped: the lower neighbour's wait_top
    unsigned* edge_done;    // local [2]: edge strips finished in this launch (top, bottom)
    unsigned* error;        // local: set to 1 if a wait timed out
    unsigned gen;           // generation being written (1, 2, ...): reads generation gen-1 everywhere
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_generation(const unsigned* flag, unsigned want, unsigned* error, int lane) {
    if (flag != nullptr && lane == 0) {
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(flag) - want) < 0) {
            if (++spins > (1u << 26)) {  // ~ seconds: a peer died; do not hang the GPU
                *error = 1u;
                break;
            }
            __nanosleep(32);
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void publish_edge(unsigned* done, unsigned* peer_flag, unsigned nstrips, unsigned gen, int lane) {
    if (peer_flag == nullptr) return;
    __syncwarp();
    if (lane == 0) {
        __threadfence_system();  // this strip's edge row is visible system-wide before it is counted
        if (atomicAdd(done, 1u) == nstrips - 1u) {
            *done = 0u;
            __threadfence_system();
            st_release_sys(peer_flag, gen);
        }
    }
}

template <int kRowsPerWarp, int kUnroll, int kMinBlocks, bool kSync>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
gol_step_u8_vec(const uint8_t* __restrict__ in, const uint8_t* __restrict__ halo_top,
                const uint8_t* __restrict__ halo_bot, uint8_t* __restrict__ out, int rows, int cols,
                int col_torus, int nstrips, const PeerSync sync) {
    const int lane = threadIdx.x & 31;
    const unsigned task = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const unsigned strip = task % (unsigned)nstrips;
    unsigned seg = task / (unsigned)nstrips;
    const unsigned nseg = (unsigned)((rows + kRowsPerWarp - 1) / kRowsPerWarp);
    if (seg >= nseg) return;  // whole warp exits together
    if (kSync && nseg >= 2u) seg = seg == 0u ? 0u : (seg == 1u ? nseg - 1u : seg - 1u);  // edge segments first
    const int row0 = (int)seg * kRowsPerWarp;
    const bool top_edge = kSync && seg == 0u, bot_edge = kSync && seg == nseg - 1u;
    if (top_edge) wait_generation(sync.wait_top, sync.gen - 1u, sync.error, lane);

    LaneAddr la;
    la.col = strip * 512u + lane * 16u;
    la.active = la.col < (unsigned)cols;
    // Neighbours across the lane boundary come from the adjacent lane (shuffle) except at
    // the strip edges, where one extra byte per row is loaded.
    const bool edge_l = lane == 0;
    const bool edge_r = (lane == 31) || (la.col + 16u >= (unsigned)cols);
    la.ld_left = la.active && edge_l;
    la.ld_right = la.active && edge_r;
    la.lcol = la.col - 1u;
    la.rcol = la.col + 16u;
    if (la.ld_left && la.col == 0u) {
        if (col_torus) la.lcol = cols - 1; else la.ld_left = false;
    }
    if (la.ld_right && la.rcol >= (unsigned)cols) {
        if (col_torus) la.rcol = 0u; else la.ld_right = false;
    }

    const int rend = min(row0 + kRowsPerWarp, rows);
    const size_t pitch = (size_t)cols;
    const uint8_t* rp = in + (size_t)row0 * pitch;                      // row r
    Row a = load_row_checked(row0 > 0 ? rp - pitch : halo_top, la);     // row r-1
    Row b = load_row(rp, la);
    uint8_t* op = out + (size_t)row0 * pitch + la.col;

    int r = row0;
    // full groups whose look-ahead rows r+1 .. r+kUnroll all lie inside this block: no checks
    for (; r + kUnroll <= rend && r + kUnroll < rows; r += kUnroll) {
        Row nx[kUnroll];
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) nx[k] = load_row(rp + (size_t)(k + 1) * pitch, la);
        rp += (size_t)kUnroll * pitch;
#pragma unroll
        for (int k = 0; k < kUnroll; ++k) {
            const uint4 o = gol_rows(a, b, nx[k], edge_l, edge_r);
            if (la.active) __stcs(reinterpret_cast<uint4*>(op), o);
            op += pitch;
            a = b;
            b = nx[k];
        }
    }
    // remainder (at most kUnroll rows; touches the bottom halo when the block ends here)
    for (; r < rend; ++r) {
        const int rr = r + 1;
        if (bot_edge && rr == rows) wait_generation(sync.wait_bot, sync.gen - 1u, sync.error, lane);
        const Row c = load_row_checked(rr < rows ? rp + pitch : (rr == rows ? halo_bot : nullptr), la);
        rp += pitch;
        const uint4 o = gol_rows(a, b, c, edge_l, edge_r);
        if (la.active) __stcs(reinterpret_cast<uint4*>(op), o);
        op += pitch;
        a = b;
        b = c;
    }
    if (top_edge) publish_edge(sync.edge_done, sync.signal_up, (unsigned)nstrips, sync.gen, lane);
    if (bot_edge) publish_edge(sync.edge_done + 1, sync.signal_down, (unsigned)nstrips, sync.gen, lane);
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
    const bool aligned = (cols % 16 == 0) && rows < (1ll << 30) && cols < (1ll << 30) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                         (halo_top == nullptr || (reinterpret_cast<uintptr_t>(halo_top) & 15) == 0) &&
                         (halo_bot == nullptr || (reinterpret_cast<uintptr_t>(halo_bot) & 15) == 0);
    if (aligned) {
        const int nstrips = (int)mb::ceil_div(cols, 512);
        // 16 rows per warp, 4 rows loaded ahead, 4 CTAs per SM: the best of the sweep recorded in profiles/README.md
        // (rows-per-warp / rows-ahead / min-CTAs: 64/4/3 95.7 us ... 16/4/4 90.1 us on 16384^2)
        constexpr int kRpw = 16;
        const int64_t tasks = (int64_t)nstrips * mb::ceil_div(rows, kRpw);
        gol_step_u8_vec<kRpw, 4, 4, false><<<(unsigned)mb::ceil_div(tasks, kWarps), kWarps * 32, 0, stream>>>(
            in, halo_top, halo_bot, out, (int)rows, (int)cols, col_torus, nstrips, PeerSync{});
    } else {
        const int64_t n = rows * cols;
        gol_step_u8_scalar<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(in, halo_top, halo_bot, out,
                                                                               rows, cols, col_torus);
    }
    MB_LAUNCH_CHECK("mb_gol_step_u8");
    return MB_OK;
}

// One generation of a row block whose neighbours' edge rows (halo_top / halo_bot, peer-mapped) belong to other
// GPUs; the hand-shake with the neighbours is fused into the kernel (see PeerSync).  wait_top / wait_bot are this
// rank's flags, signal_up / signal_down the peer-mapped addresses of the neighbours' flags (NULL on a clamped
// outer edge), edge_done two zero-initialised local counters, error a local word, gen = 1, 2, 3, ...
MB_API int mb_gol_step_u8_fused(const uint8_t* in, const uint8_t* halo_top, const uint8_t* halo_bot, uint8_t* out,
                                int64_t rows, int64_t cols, int col_torus, unsigned* wait_top, unsigned* wait_bot,
                                unsigned* signal_up, unsigned* signal_down, unsigned* edge_done, unsigned* error,
                                unsigned gen, cudaStream_t stream) {
    MB_REQUIRE(rows > 0 && cols > 0 && rows < (1ll << 30) && cols < (1ll << 30), "mb_gol_step_u8_fused: bad shape");
    MB_REQUIRE(cols % 16 == 0, "mb_gol_step_u8_fused: cols must be a multiple of 16");
    MB_REQUIRE(in && out && in != out && edge_done && error, "mb_gol_step_u8_fused: null buffer");
    MB_REQUIRE(!(col_torus && cols < 3), "mb_gol_step_u8_fused: torus needs cols >= 3");
    const PeerSync sync{halo_top ? wait_top : nullptr, halo_bot ? wait_bot : nullptr, halo_top ? signal_up : nullptr,
                        halo_bot ? signal_down : nullptr, edge_done, error, gen};
    const int nstrips = (int)mb::ceil_div(cols, 512);
    constexpr int kRpw = 16;
    const int64_t tasks = (int64_t)nstrips * mb::ceil_div(rows, kRpw);
    gol_step_u8_vec<kRpw, 4, 4, true><<<(unsigned)mb::ceil_div(tasks, kWarps), kWarps * 32, 0, stream>>>(
        in, halo_top, halo_bot, out, (int)rows, (int)cols, col_torus, nstrips, sync);
    MB_LAUNCH_CHECK("mb_gol_step_u8_fused");
    return MB_OK;
}
