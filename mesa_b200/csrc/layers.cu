// Property-layer kernels (K6 of SURVEY.md §2.4) and the generic neighbourhood sum (K1).
//
// The reference keeps property layers as raw numpy arrays (mesa/discrete_space/grid.py:112,
// 130-220) that users mutate in place with numpy expressions, e.g.
//   grid.sugar[:] = np.minimum(grid.sugar + 1, cap)      examples/advanced/sugarscape_g1mt/model.py:123-124
//   np.argwhere(layers["empty"])                          grid.py:308
//   np.full(dimensions, default, dtype)                   grid.py:146
// and computes neighbourhoods cell by cell (cell.py:225-276).  Here the layers are device
// buffers and these are the elementwise / aggregate / stencil kernels over them:
//   fill, affine+clamp, binary (add sub mul min max), reduce (sum min max count_nonzero),
//   Moore / von-Neumann neighbourhood sums of any radius, and 8-neighbour diffusion.
// Diffusion has NO reference implementation (parity unpinned: the oracle is a numpy
// restatement of the NetLogo `diffuse` rule stated below).
//
// All kernels are HBM-bound streaming kernels: 16-byte vector accesses in the elementwise
// paths (several in flight per thread), grid-stride over a grid sized to the SM count.
// Reductions are two-pass with a fixed partial count and a fixed element-to-thread assignment, so
// floating-point sums are reproducible run to run.  uint8 layers never go through per-element
// float64 arithmetic (conversions run at 16 / clk / SM and would bind before HBM does): the affine
// map of a uint8 layer is a 256-entry table built per CTA, the uint8 reductions work on packed words.
// The radius-1 / radius-2 neighbourhood sums and the diffusion step use a COLUMN WALKER: a thread owns
// 4 bytes' worth of adjacent columns and walks down a segment of rows with the (2r+1)-row window in
// registers, so every input row is fetched from L1/L2 once per (2r+1) outputs and from HBM once in all;
// other radii take the generic one-thread-per-cell kernel.
#include "common.cuh"

#include <float.h>

namespace {

enum { DT_U8 = 0, DT_I32 = 1, DT_F32 = 2, DT_F64 = 3 };
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_MIN = 3, OP_MAX = 4 };
enum { RED_SUM = 0, RED_MIN = 1, RED_MAX = 2, RED_COUNT_NONZERO = 3 };
constexpr int kPartials = 1024;

template <typename T> struct VecOf { static constexpr int n = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ T from_double(double v) { return (T)v; }
template <> __device__ __forceinline__ uint8_t from_double<uint8_t>(double v) {
    return (uint8_t)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));
}

inline unsigned stream_blocks(int64_t work_items) {
    const int64_t want = mb::ceil_div(work_items, 256);
    const int64_t cap = (int64_t)mb::sm_count() * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <typename T>
__global__ void k_fill(T* __restrict__ p, int64_t n, double value) {
    constexpr int V = VecOf<T>::n;
    const T v = from_double<T>(value);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // scalar head up to the first 16-byte boundary, 16-byte stores, scalar tail
    int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / sizeof(T));
    if (head > n) head = n;
    const int64_t nvec = (n - head) / V;
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int k = 0; k < V; ++k) e[k] = v;
    uint4* pv = reinterpret_cast<uint4*>(p + head);
    for (int64_t i = tid; i < nvec; i += stride) pv[i] = raw;
    if (tid < head) p[tid] = v;
    for (int64_t i = head + nvec * V + tid; i < n; i += stride) p[i] = v;
}

// out = clamp(alpha * in + beta, lo, hi), evaluated in double (float in float for f32)
template <typename T>
__global__ void k_affine(const T* __restrict__ in, T* __restrict__ out, int64_t n, double alpha, double beta,
                         double lo, double hi) {
    constexpr int V = VecOf<T>::n;
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto f = [&](T x) {
        double y = alpha * (double)x + beta;
        y = y < lo ? lo : (y > hi ? hi : y);
        return from_double<T>(y);
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned) {
        int64_t i = tid;
        for (; i + stride < nvec; i += 2 * stride) {   // two independent vectors in flight per thread
            uint4 raw0 = __ldg(reinterpret_cast<const uint4*>(in) + i);
            uint4 raw1 = __ldg(reinterpret_cast<const uint4*>(in) + i + stride);
            T* e0 = reinterpret_cast<T*>(&raw0);
            T* e1 = reinterpret_cast<T*>(&raw1);
#pragma unroll
            for (int k = 0; k < V; ++k) e0[k] = f(e0[k]);
#pragma unroll
            for (int k = 0; k < V; ++k) e1[k] = f(e1[k]);
            reinterpret_cast<uint4*>(out)[i] = raw0;
            reinterpret_cast<uint4*>(out)[i + stride] = raw1;
        }
        for (; i < nvec; i += stride) {
            uint4 raw = __ldg(reinterpret_cast<const uint4*>(in) + i);
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int k = 0; k < V; ++k) e[k] = f(e[k]);
            reinterpret_cast<uint4*>(out)[i] = raw;
        }
        for (int64_t i = nvec * V + tid; i < n; i += stride) out[i] = f(in[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = f(in[i]);
    }
}

// uint8 layers: the map has 256 possible arguments -- every CTA evaluates it once per value in float64 (the
// same expression as k_affine) into shared memory and the stream becomes byte look-ups.
__global__ void __launch_bounds__(256) k_affine_u8(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t n,
                                                   double alpha, double beta, double lo, double hi) {
    __shared__ uint8_t lut[256];
    {
        double y = alpha * (double)threadIdx.x + beta;
        y = y < lo ? lo : (y > hi ? hi : y);
        lut[threadIdx.x] = from_double<uint8_t>(y);
    }
    __syncthreads();
    const int64_t nvec = n / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned) {
        for (int64_t i = tid; i < nvec; i += stride) {
            uint4 raw = __ldg(reinterpret_cast<const uint4*>(in) + i);
            uint32_t* w = reinterpret_cast<uint32_t*>(&raw);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = w[k];
                w[k] = (uint32_t)lut[x & 255u] | ((uint32_t)lut[(x >> 8) & 255u] << 8) |
                       ((uint32_t)lut[(x >> 16) & 255u] << 16) | ((uint32_t)lut[x >> 24] << 24);
            }
            reinterpret_cast<uint4*>(out)[i] = raw;
        }
        for (int64_t i = nvec * 16 + tid; i < n; i += stride) out[i] = lut[in[i]];
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = lut[in[i]];
    }
}

template <typename T, int OP> __device__ __forceinline__ T apply_op(T a, T b) {
    if (OP == OP_ADD) return (T)(a + b);
    if (OP == OP_SUB) return (T)(a - b);
    if (OP == OP_MUL) return (T)(a * b);
    if (OP == OP_MIN) return a < b ? a : b;
    return a > b ? a : b;
}

// 16 uint8 cells at a time: packed-byte add / sub / min / max (the per-byte path costs ~4 ALU ops per cell and
// binds before HBM does: ncu showed the ALU pipe at 72 % with DRAM at 76 %)
template <int OP> __device__ __forceinline__ unsigned apply_op_u8x4(unsigned a, unsigned b) {
    if (OP == OP_ADD) return __vadd4(a, b);
    if (OP == OP_SUB) return __vsub4(a, b);
    if (OP == OP_MIN) return __vminu4(a, b);
    if (OP == OP_MAX) return __vmaxu4(a, b);
    unsigned r = 0;   // OP_MUL: per byte, modulo 256 like numpy uint8
#pragma unroll
    for (int k = 0; k < 4; ++k) r |= ((((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u)) & 255u) << (8 * k);
    return r;
}

template <typename T, int OP>
__device__ __forceinline__ void apply_op_vec(uint4& ra, const uint4& rb) {
    if (sizeof(T) == 1) {
        ra.x = apply_op_u8x4<OP>(ra.x, rb.x);
        ra.y = apply_op_u8x4<OP>(ra.y, rb.y);
        ra.z = apply_op_u8x4<OP>(ra.z, rb.z);
        ra.w = apply_op_u8x4<OP>(ra.w, rb.w);
    } else {
        T* ea = reinterpret_cast<T*>(&ra);
        const T* eb = reinterpret_cast<const T*>(&rb);
#pragma unroll
        for (int k = 0; k < VecOf<T>::n; ++k) ea[k] = apply_op<T, OP>(ea[k], eb[k]);
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) k_binary(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                int64_t n) {
    constexpr int V = VecOf<T>::n;
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                           reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned) {
        for (int64_t i = tid; i < nvec; i += stride) {   // (two vectors per iteration measured slower: 0.527 vs 0.494 ms)
            uint4 ra = __ldg(reinterpret_cast<const uint4*>(a) + i);
            const uint4 rb = __ldg(reinterpret_cast<const uint4*>(b) + i);
            apply_op_vec<T, OP>(ra, rb);
            reinterpret_cast<uint4*>(out)[i] = ra;
        }
        for (int64_t i = nvec * V + tid; i < n; i += stride) out[i] = apply_op<T, OP>(a[i], b[i]);
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = apply_op<T, OP>(a[i], b[i]);
    }
}

template <typename T>
void launch_binary(const T* a, const T* b, T* out, int64_t n, int op, unsigned blocks, cudaStream_t stream) {
    switch (op) {
        case OP_ADD: k_binary<T, OP_ADD><<<blocks, 256, 0, stream>>>(a, b, out, n); break;
        case OP_SUB: k_binary<T, OP_SUB><<<blocks, 256, 0, stream>>>(a, b, out, n); break;
        case OP_MUL: k_binary<T, OP_MUL><<<blocks, 256, 0, stream>>>(a, b, out, n); break;
        case OP_MIN: k_binary<T, OP_MIN><<<blocks, 256, 0, stream>>>(a, b, out, n); break;
        default: k_binary<T, OP_MAX><<<blocks, 256, 0, stream>>>(a, b, out, n); break;
    }
}

// ---- reductions: pass 1 -> kPartials partials (fixed assignment), pass 2 -> scalar ----------
template <typename A> __device__ __forceinline__ A red_identity(int op);
template <> __device__ __forceinline__ double red_identity<double>(int op) {
    return op == RED_MIN ? DBL_MAX : (op == RED_MAX ? -DBL_MAX : 0.0);
}
template <> __device__ __forceinline__ long long red_identity<long long>(int op) {
    return op == RED_MIN ? LLONG_MAX : (op == RED_MAX ? LLONG_MIN : 0ll);
}
template <typename A> __device__ __forceinline__ A red_combine(A a, A b, int op) {
    if (op == RED_MIN) return a < b ? a : b;
    if (op == RED_MAX) return a > b ? a : b;
    return a + b;
}

template <typename A>
__device__ __forceinline__ A block_reduce(A v, int op) {
    __shared__ A sh[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine<A>(v, __shfl_down_sync(0xffffffffu, v, o), op);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? sh[lane] : red_identity<A>(op);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = red_combine<A>(v, __shfl_down_sync(0xffffffffu, v, o), op);
    }
    return v;  // valid in thread 0
}

// element -> accumulator for one reduction (count_nonzero counts, the others take the value)
template <typename T, typename A, int OP> __device__ __forceinline__ A red_value(T x) {
    return OP == RED_COUNT_NONZERO ? (A)(x != (T)0) : (A)x;
}

// One 16-byte vector folded into the accumulator.  uint8: sum by IDP.4A against 0x01010101, count_nonzero by an
// OR-fold of each byte onto its lowest bit + POPC; min / max per byte.
template <typename T, typename A, int OP>
__device__ __forceinline__ A red_vec(A acc, const uint4& raw) {
    constexpr int ROP = OP == RED_COUNT_NONZERO ? RED_SUM : OP;
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int k = 0; k < VecOf<T>::n; ++k) acc = red_combine<A>(acc, red_value<T, A, OP>(e[k]), ROP);
    return acc;
}
template <> __device__ __forceinline__ long long red_vec<uint8_t, long long, RED_SUM>(long long acc, const uint4& raw) {
    unsigned s = __dp4a(raw.x, 0x01010101u, 0u);
    s = __dp4a(raw.y, 0x01010101u, s);
    s = __dp4a(raw.z, 0x01010101u, s);
    s = __dp4a(raw.w, 0x01010101u, s);
    return acc + (long long)s;
}
__device__ __forceinline__ unsigned nonzero_bytes(unsigned x) {
    x |= x >> 4;
    x |= x >> 2;
    x |= x >> 1;
    return __popc(x & 0x01010101u);
}
template <>
__device__ __forceinline__ long long red_vec<uint8_t, long long, RED_COUNT_NONZERO>(long long acc, const uint4& raw) {
    return acc + (long long)(nonzero_bytes(raw.x) + nonzero_bytes(raw.y) + nonzero_bytes(raw.z) + nonzero_bytes(raw.w));
}

// uint8 min / max: bytes widened to two 16-bit lanes per word, VIMNMX.U16x2 tree, one scalar fold per vector
template <bool IS_MIN> __device__ __forceinline__ unsigned mnmx16x2(unsigned a, unsigned b) {
    return IS_MIN ? __vminu2(a, b) : __vmaxu2(a, b);
}
template <bool IS_MIN> __device__ __forceinline__ long long red_vec_u8_mnmx(long long acc, const uint4& raw) {
    const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
    unsigned m = mnmx16x2<IS_MIN>(w[0] & 0x00ff00ffu, (w[0] >> 8) & 0x00ff00ffu);
#pragma unroll
    for (int k = 1; k < 4; ++k)
        m = mnmx16x2<IS_MIN>(m, mnmx16x2<IS_MIN>(w[k] & 0x00ff00ffu, (w[k] >> 8) & 0x00ff00ffu));
    const unsigned lo = m & 0xffffu, hi = m >> 16;
    const long long r = (long long)(IS_MIN ? (lo < hi ? lo : hi) : (lo > hi ? lo : hi));
    return IS_MIN ? (acc < r ? acc : r) : (acc > r ? acc : r);
}
template <> __device__ __forceinline__ long long red_vec<uint8_t, long long, RED_MIN>(long long acc, const uint4& raw) {
    return red_vec_u8_mnmx<true>(acc, raw);
}
template <> __device__ __forceinline__ long long red_vec<uint8_t, long long, RED_MAX>(long long acc, const uint4& raw) {
    return red_vec_u8_mnmx<false>(acc, raw);
}

template <typename T, typename A, int OP>
__global__ void __launch_bounds__(256) k_reduce1(const T* __restrict__ in, int64_t n, A* __restrict__ partial) {
    // block b owns the contiguous chunk [b*chunk, (b+1)*chunk), chunk a multiple of one vector, and inside it
    // thread t owns vectors t, t+256, ...: a fixed assignment and a fixed combination order => reproducible.
    constexpr int V = VecOf<T>::n;
    constexpr int ROP = OP == RED_COUNT_NONZERO ? RED_SUM : OP;
    int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + V - 1) / V * V;
    const int64_t lo = (int64_t)blockIdx.x * chunk < n ? (int64_t)blockIdx.x * chunk : n;
    const int64_t hi = lo + chunk < n ? lo + chunk : n;
    A acc = red_identity<A>(ROP);
    if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
        const uint4* pv = reinterpret_cast<const uint4*>(in + lo);   // lo is a multiple of V
        const int64_t nvec = (hi - lo) / V;
        int64_t i = threadIdx.x;
        for (; i + 3 * 256 < nvec; i += 4 * 256) {   // four independent 16-byte loads in flight per thread (eight: slower)
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = ldg_nc_v4(pv + i + u * 256);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = red_vec<T, A, OP>(acc, r[u]);
        }
        for (; i < nvec; i += 256) acc = red_vec<T, A, OP>(acc, ldg_nc_v4(pv + i));
        for (int64_t j = lo + nvec * V + threadIdx.x; j < hi; j += 256)
            acc = red_combine<A>(acc, red_value<T, A, OP>(in[j]), ROP);
    } else {
        for (int64_t j = lo + threadIdx.x; j < hi; j += 256) acc = red_combine<A>(acc, red_value<T, A, OP>(in[j]), ROP);
    }
    acc = block_reduce<A>(acc, ROP);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <typename T, typename A>
void launch_reduce1(const T* in, int64_t n, int op, A* partial, int parts, cudaStream_t stream) {
    switch (op) {
        case RED_SUM: k_reduce1<T, A, RED_SUM><<<parts, 256, 0, stream>>>(in, n, partial); break;
        case RED_MIN: k_reduce1<T, A, RED_MIN><<<parts, 256, 0, stream>>>(in, n, partial); break;
        case RED_MAX: k_reduce1<T, A, RED_MAX><<<parts, 256, 0, stream>>>(in, n, partial); break;
        default: k_reduce1<T, A, RED_COUNT_NONZERO><<<parts, 256, 0, stream>>>(in, n, partial); break;
    }
}

template <typename A>
__global__ void __launch_bounds__(1024) k_reduce2(const A* __restrict__ partial, int nparts, int op, A* __restrict__ out) {
    A acc = (int)threadIdx.x < nparts ? partial[threadIdx.x] : red_identity<A>(op);
    acc = block_reduce<A>(acc, op);
    if (threadIdx.x == 0) *out = acc;
}

// ---- neighbourhood sums (Moore = Chebyshev square, von Neumann = Manhattan diamond; cell.py:225-276,
//      tests/discrete_space/test_discrete_space.py:380-398) ------------------------------------
template <typename T, typename A>
__global__ void k_neighborhood_sum(const T* __restrict__ in, const T* __restrict__ halo_top,
                                   const T* __restrict__ halo_bot, A* __restrict__ out, int64_t rows, int64_t cols,
                                   int radius, int von_neumann, int include_center, int col_torus) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t i = idx / cols, j = idx % cols;
    A acc = 0;
    for (int di = -radius; di <= radius; ++di) {
        const int64_t r = i + di;
        const T* p;
        if (r < 0) p = (halo_top != nullptr && r >= -radius) ? halo_top + (radius + r) * cols : nullptr;
        else if (r >= rows) p = (halo_bot != nullptr && r - rows < radius) ? halo_bot + (r - rows) * cols : nullptr;
        else p = in + r * cols;
        if (p == nullptr) continue;
        const int w = von_neumann ? radius - (di < 0 ? -di : di) : radius;
        for (int dj = -w; dj <= w; ++dj) {
            if (di == 0 && dj == 0 && !include_center) continue;
            int64_t jj = j + dj;
            if (jj < 0 || jj >= cols) {
                if (!col_torus) continue;
                jj = (jj % cols + cols) % cols;
            }
            acc += (A)p[jj];
        }
    }
    out[idx] = acc;
}

// ---- diffusion (NetLogo `diffuse`): every cell gives rate/8 of its value to each of its existing
// Moore neighbours and keeps the rest (cells on a clamped edge keep the shares that have nowhere to go):
//   out[c] = in[c] * (1 - rate * k_c / 8) + rate/8 * sum_{nb} in[nb],  k_c = #existing neighbours
template <typename T>
__global__ void k_diffuse(const T* __restrict__ in, const T* __restrict__ halo_top, const T* __restrict__ halo_bot,
                          T* __restrict__ out, int64_t rows, int64_t cols, double rate, int col_torus) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t i = idx / cols, j = idx % cols;
    double acc = 0.0;
    int k = 0;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
        const int64_t r = i + di;
        const T* p;
        if (r < 0) p = halo_top;
        else if (r >= rows) p = halo_bot;
        else p = in + r * cols;
        if (p == nullptr) continue;
#pragma unroll
        for (int dj = -1; dj <= 1; ++dj) {
            if (di == 0 && dj == 0) continue;
            int64_t jj = j + dj;
            if (jj < 0 || jj >= cols) {
                if (!col_torus) continue;
                jj = (jj + cols) % cols;
            }
            acc += (double)p[jj];
            ++k;
        }
    }
    const double self = (double)in[idx];
    out[idx] = (T)__fma_rn(rate / 8.0, acc, __dmul_rn(self, 1.0 - rate * (double)k / 8.0));
}

// ---- column walker: radius-R neighbourhood sums and diffusion -----------------------------------------
// A thread owns V adjacent columns (V * sizeof(T) >= 4 bytes: V = 4 for uint8, 1 otherwise; cols % V == 0 and
// V*sizeof(T)-aligned rows are checked on the host) and walks down SEG rows of the layer.  Per input row it
// loads the 2*NL+1 vectors around its columns (the side ones are L1 hits of the neighbour threads' loads),
// converts them once to the accumulator type and keeps the last 2R+1 rows in registers; the loop is fully
// unrolled, so the window never moves and the loads of later rows are hoisted above the arithmetic.  Rows
// outside the block come from halo_top / halo_bot (R rows each; nullptr = no cells there), columns outside wrap
// (col_torus) or contribute nothing.  Sums run in the (di, dj) order of the generic kernel and of the oracle.
template <typename T, int V> struct WalkVec;
template <> struct WalkVec<uint8_t, 4> { using type = uint32_t; };
template <typename T> struct WalkVec<T, 1> { using type = T; };

// The loads of a row are UNCONDITIONAL (absent rows / columns read a valid dummy address) and stay raw; the
// conversion to the accumulator type and the zeroing of absent cells happen when the row enters the window, PF
// rows later.  (A predicated load followed by its conversion under the same predicate makes the compiler emit
// "@P LDG; @P F2F" back to back -- every row then waits a full memory latency; seen in the SASS of the first version.)
template <typename T, int V, int NL>
__device__ __forceinline__ void walk_fetch_row(const T* __restrict__ p, const int64_t (&cs)[2 * NL + 1],
                                               typename WalkVec<T, V>::type (&dst)[2 * NL + 1]) {
    using VT = typename WalkVec<T, V>::type;
#pragma unroll
    for (int k = 0; k < 2 * NL + 1; ++k) dst[k] = __ldg(reinterpret_cast<const VT*>(p) + cs[k]);
}

template <typename T, typename A, int V, int NL, bool ALL_OK = false>
__device__ __forceinline__ void walk_widen_row(const typename WalkVec<T, V>::type (&src)[2 * NL + 1], bool row_ok,
                                               const int64_t (&cv)[2 * NL + 1], A (&dst)[(2 * NL + 1) * V]) {
#pragma unroll
    for (int k = 0; k < 2 * NL + 1; ++k) {
        const T* e = reinterpret_cast<const T*>(&src[k]);
        const bool ok = ALL_OK || (row_ok && cv[k] >= 0);
#pragma unroll
        for (int v = 0; v < V; ++v) dst[k * V + v] = ok ? (A)e[v] : (A)0;
    }
}

template <typename T, int R>
__device__ __forceinline__ const T* walk_row_ptr(const T* in, const T* halo_top, const T* halo_bot, int64_t r,
                                                 int64_t rows, int64_t cols) {
    if (r < 0) return (halo_top != nullptr && r >= -R) ? halo_top + (R + r) * cols : nullptr;
    if (r >= rows) return (halo_bot != nullptr && r - rows < R) ? halo_bot + (r - rows) * cols : nullptr;
    return in + r * cols;
}

template <int V, int NL>
__device__ __forceinline__ void walk_columns(int64_t jv, int64_t nvec, int col_torus, int64_t (&cv)[2 * NL + 1]) {
#pragma unroll
    for (int k = 0; k < 2 * NL + 1; ++k) {
        int64_t c = jv + k - NL;
        if (c < 0 || c >= nvec) c = col_torus ? ((c % nvec) + nvec) % nvec : -1;
        cv[k] = c;
    }
}

// INTERIOR = every row i0-R .. i0+SEG-1+R lies inside this block and every column vector of the CTA exists: plain pointer
// arithmetic, no row / column selects (the generic body spends ~50 of its ~85 instructions per warp and row on them and is
// ALU-bound: ncu 86 % ALU pipe on the uint8 layer).  The condition is uniform over the CTA.
template <typename T, typename A, int R, bool VN, int V, int SEG, bool INTERIOR>
__device__ __forceinline__ void nsum_walk_body(const T* __restrict__ in, const T* __restrict__ halo_top,
                                               const T* __restrict__ halo_bot, A* __restrict__ out, int64_t rows,
                                               int64_t cols, int include_center, int col_torus, int64_t jv, int64_t nvec,
                                               int64_t i0) {
    constexpr int NL = (R + V - 1) / V, W = (2 * NL + 1) * V;
    using VT = typename WalkVec<T, V>::type;
    int64_t cv[2 * NL + 1], cs[2 * NL + 1];
    if (INTERIOR) {
#pragma unroll
        for (int k = 0; k < 2 * NL + 1; ++k) cv[k] = cs[k] = jv + k - NL;
    } else {
        walk_columns<V, NL>(jv, nvec, col_torus, cv);
#pragma unroll
        for (int k = 0; k < 2 * NL + 1; ++k) cs[k] = cv[k] >= 0 ? cv[k] : jv;
    }
    // rows i0-R .. i0+SEG-1+R of the walk; statically indexed (only 2R+1+PF of them are live at a time):
    // the raw loads run PF rows ahead of the conversion + arithmetic
    constexpr int PF = 3;
    VT raw[SEG + 2 * R][2 * NL + 1];
    bool ok[SEG + 2 * R];
    A win[SEG + 2 * R][W];
    const T* const first = in + (i0 - R) * cols;   // INTERIOR only
    auto fetch = [&](int k) {
        if (INTERIOR) {
            ok[k] = true;
            walk_fetch_row<T, V, NL>(first + (int64_t)k * cols, cs, raw[k]);
        } else {
            const T* p = walk_row_ptr<T, R>(in, halo_top, halo_bot, i0 - R + k, rows, cols);
            ok[k] = p != nullptr;
            walk_fetch_row<T, V, NL>(ok[k] ? p : in, cs, raw[k]);
        }
    };
#pragma unroll
    for (int k = 0; k < 2 * R + PF; ++k) fetch(k);
#pragma unroll
    for (int k = 0; k < 2 * R; ++k) walk_widen_row<T, A, V, NL, INTERIOR>(raw[k], ok[k], cv, win[k]);
    A* o = out + i0 * cols + jv * V;
#pragma unroll
    for (int s = 0; s < SEG; ++s) {
        if (s + 2 * R + PF < SEG + 2 * R) fetch((s + 2 * R + PF) % (SEG + 2 * R));
        walk_widen_row<T, A, V, NL, INTERIOR>(raw[s + 2 * R], ok[s + 2 * R], cv, win[s + 2 * R]);
        A res[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            A acc = 0;
#pragma unroll
            for (int di = -R; di <= R; ++di)
#pragma unroll
                for (int dj = -R; dj <= R; ++dj) {
                    if (VN && (di < 0 ? -di : di) + (dj < 0 ? -dj : dj) > R) continue;
                    const A x = win[s + di + R][NL * V + v + dj];
                    if (di == 0 && dj == 0) {
                        if (include_center) acc += x;
                    } else {
                        acc += x;
                    }
                }
            res[v] = acc;
        }
        if (INTERIOR || i0 + s < rows) {
            if (V == 4) {
                __stcs(reinterpret_cast<int4*>(o), make_int4((int)res[0], (int)res[V > 1 ? 1 : 0], (int)res[V > 2 ? 2 : 0], (int)res[V > 3 ? 3 : 0]));
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) __stcs(o + v, res[v]);
            }
        }
        o += cols;
    }
}

template <typename T, typename A, int R, bool VN, int V, int SEG, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_nsum_walk(const T* __restrict__ in, const T* __restrict__ halo_top, const T* __restrict__ halo_bot,
            A* __restrict__ out, int64_t rows, int64_t cols, int include_center, int col_torus) {
    constexpr int NL = (R + V - 1) / V;
    const int64_t nvec = cols / V;
    const int64_t jv = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (jv >= nvec) return;
    const int64_t i0 = (int64_t)blockIdx.y * SEG;
    const int64_t c0 = (int64_t)blockIdx.x * 128;
    const bool interior = i0 >= R && i0 + SEG + R <= rows && c0 >= NL && c0 + 127 + NL < nvec;
    if (interior)
        nsum_walk_body<T, A, R, VN, V, SEG, true>(in, halo_top, halo_bot, out, rows, cols, include_center, col_torus, jv, nvec, i0);
    else
        nsum_walk_body<T, A, R, VN, V, SEG, false>(in, halo_top, halo_bot, out, rows, cols, include_center, col_torus, jv, nvec, i0);
}

// diffusion on the same walker (radius 1, Moore, one column per thread); k_c from the position
template <typename T, int SEG>
__global__ void __launch_bounds__(128, 8)
k_diffuse_walk(const T* __restrict__ in, const T* __restrict__ halo_top, const T* __restrict__ halo_bot,
               T* __restrict__ out, int64_t rows, int64_t cols, double rate, int col_torus) {
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= cols) return;
    int64_t cv[3], cs[3];
    walk_columns<1, 1>(j, cols, col_torus, cv);
#pragma unroll
    for (int k = 0; k < 3; ++k) cs[k] = cv[k] >= 0 ? cv[k] : j;
    const int kcols = (cv[0] >= 0) + 1 + (cv[2] >= 0);
    const int64_t i0 = (int64_t)blockIdx.y * SEG;
    constexpr int PF = 3;
    T raw[SEG + 2][3];
    bool ok[SEG + 2];
    double win[SEG + 2][3];
    auto fetch = [&](int k) {
        const T* p = walk_row_ptr<T, 1>(in, halo_top, halo_bot, i0 - 1 + k, rows, cols);
        ok[k] = p != nullptr;
        walk_fetch_row<T, 1, 1>(ok[k] ? p : in, cs, raw[k]);
    };
#pragma unroll
    for (int k = 0; k < 2 + PF; ++k) fetch(k);
    walk_widen_row<T, double, 1, 1>(raw[0], ok[0], cv, win[0]);
    walk_widen_row<T, double, 1, 1>(raw[1], ok[1], cv, win[1]);
    const double share = rate / 8.0;
#pragma unroll
    for (int s = 0; s < SEG; ++s) {
        const int64_t i = i0 + s;
        if (s + 2 + PF < SEG + 2) fetch((s + 2 + PF) % (SEG + 2));
        walk_widen_row<T, double, 1, 1>(raw[s + 2], ok[s + 2], cv, win[s + 2]);
        const int krows = ((i > 0) || halo_top != nullptr) + 1 + ((i + 1 < rows) || halo_bot != nullptr);
        const int k = krows * kcols - 1;
        double acc = 0.0;
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int dj = 0; dj < 3; ++dj)
                if (!(di == 1 && dj == 1)) acc += win[s + di][dj];
        const double self = win[s + 1][1];
        if (i < rows) __stcs(out + i * cols + j, (T)__fma_rn(share, acc, __dmul_rn(self, 1.0 - rate * (double)k / 8.0)));
    }
}

// ---- float32 layers, radius-1 Moore box: 4 columns per thread, separable sums ---------------------------------
// The generic walker spends ~130 issue slots per warp and row on a float32 layer (3 F2F conversions and 8 DADDs per
// cell, address arithmetic per 4-byte load) and ends up pipe-bound at < 50 % of HBM.  Here a thread owns FOUR
// adjacent columns: one 16-byte load per row plus the two side cells (L1 hits), 6 conversions and 8 DADDs for the four
// horizontal 3-sums, and per output row (H0 + H1) + H2 minus the centre.  Everything is float64, so for float32
// inputs the result is the exact sum of the 8 (or 9) neighbours -- identical to the (di, dj)-ordered sum of the
// generic kernels and of the oracle -- as long as the magnitudes inside one neighbourhood span less than 2^25
// (53 - 24 - 4 bits of headroom); beyond that the last bit may differ.  DIFFUSE selects the diffusion epilogue.
template <typename T> struct Box3Row {
    uint4 own;      // 16 bytes = V adjacent cells (4 float32 / 2 float64)
    T left, right;
};

template <typename T>
__device__ __forceinline__ Box3Row<T> box3_load(const T* __restrict__ p, int64_t j0, int64_t jls, int64_t jrs) {
    Box3Row<T> r;   // unconditional loads (see walk_fetch_row); absent cells are zeroed when the row is widened
    r.own = __ldg(reinterpret_cast<const uint4*>(p + j0));
    r.left = __ldg(p + jls);
    r.right = __ldg(p + jrs);
    return r;
}

// T = float: sums (DIFFUSE = false, float64 out) and diffusion (float out); T = double: diffusion only -- the separable
// order is not the oracle's (di, dj) order, which float64 neighbourhood SUMS keep by using the generic walker.
template <typename T, bool DIFFUSE, int SEG, bool INTERIOR>
__device__ __forceinline__ void box3_body(const T* __restrict__ in, const T* __restrict__ halo_top,
                                          const T* __restrict__ halo_bot, void* __restrict__ out_, int64_t rows,
                                          int64_t cols, int include_center, int col_torus, double rate, int64_t j0,
                                          int64_t i0) {
    constexpr int V = 16 / sizeof(T);
    // INTERIOR: all rows i0-1 .. i0+SEG exist in this block and the CTA touches neither column edge
    const int64_t jl = INTERIOR ? j0 - 1 : (j0 > 0 ? j0 - 1 : (col_torus ? cols - 1 : -1));
    const int64_t jr = INTERIOR ? j0 + V : (j0 + V < cols ? j0 + V : (col_torus ? 0 : -1));
    constexpr int PF = 4;   // rows of loads in flight ahead of the arithmetic
    Box3Row<T> raw[SEG + 2];
    bool ok[SEG + 2];
    double H[SEG + 2][V], C[SEG + 2][V];
    const int64_t jls = jl >= 0 ? jl : j0, jrs = jr >= 0 ? jr : j0;
    const T* const first = in + (i0 - 1) * cols;   // INTERIOR only
    auto fetch = [&](int k) {
        if (INTERIOR) {
            ok[k] = true;
            raw[k] = box3_load<T>(first + (int64_t)k * cols, j0, jls, jrs);
        } else {
            const T* p = walk_row_ptr<T, 1>(in, halo_top, halo_bot, i0 - 1 + k, rows, cols);
            ok[k] = p != nullptr;
            raw[k] = box3_load<T>(ok[k] ? p : in, j0, jls, jrs);
        }
    };
    auto widen = [&](int k) {
        const bool rk = INTERIOR || ok[k];
        const bool lk = INTERIOR || (rk && jl >= 0), rrk = INTERIOR || (rk && jr >= 0);
        const T* own = reinterpret_cast<const T*>(&raw[k].own);
        double e[V + 2];
        e[0] = lk ? (double)raw[k].left : 0.0;
#pragma unroll
        for (int v = 0; v < V; ++v) e[v + 1] = rk ? (double)own[v] : 0.0;
        e[V + 1] = rrk ? (double)raw[k].right : 0.0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            C[k][v] = e[v + 1];
            H[k][v] = (e[v] + e[v + 1]) + e[v + 2];
        }
    };
#pragma unroll
    for (int k = 0; k < 2 + PF; ++k) fetch(k);
    widen(0);
    widen(1);
    const double share = rate / 8.0;
    const int kcf = (jl >= 0) + 2, kcl = (jr >= 0) + 2;   // existing columns around the first / last owned column
    // 1 - rate * k / 8 for the first / inner / last owned column of a row with both vertical neighbours
    // (k = 3 * kcols - 1); rows on a clamped edge (krows = 2, or 1 on a one-row grid) compute theirs on the spot
    const double keepf = 1.0 - rate * (double)(3 * kcf - 1) / 8.0, keepi = 1.0 - rate, keepl = 1.0 - rate * (double)(3 * kcl - 1) / 8.0;
    T* ot = reinterpret_cast<T*>(out_) + i0 * cols + j0;
    double* od = reinterpret_cast<double*>(out_) + i0 * cols + j0;
#pragma unroll
    for (int s = 0; s < SEG; ++s) {
        const int64_t i = i0 + s;
        if (s + 2 + PF < SEG + 2) fetch((s + 2 + PF) % (SEG + 2));
        widen(s + 2);
        if (INTERIOR || i < rows) {
            double res[V];
            double kf = keepf, ki = keepi, kl = keepl;
            if (DIFFUSE && !INTERIOR) {
                const int krows = ((i > 0) || halo_top != nullptr) + 1 + ((i + 1 < rows) || halo_bot != nullptr);
                if (krows != 3) {
                    kf = 1.0 - rate * (double)(krows * kcf - 1) / 8.0;
                    ki = 1.0 - rate * (double)(krows * 3 - 1) / 8.0;
                    kl = 1.0 - rate * (double)(krows * kcl - 1) / 8.0;
                }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const double box = (H[s][v] + H[s + 1][v]) + H[s + 2][v];
                if (DIFFUSE) {
                    // explicit rounding points: both bodies (and every cell) evaluate the identical expression
                    res[v] = __fma_rn(share, box - C[s + 1][v], __dmul_rn(C[s + 1][v], v == 0 ? kf : (v == V - 1 ? kl : ki)));
                } else {
                    res[v] = include_center ? box : box - C[s + 1][v];
                }
            }
            if (DIFFUSE) {
                uint4 o;
                T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
                for (int v = 0; v < V; ++v) oe[v] = (T)res[v];
                __stcs(reinterpret_cast<uint4*>(ot), o);
            } else {
#pragma unroll
                for (int v = 0; v < V; v += 2) __stcs(reinterpret_cast<double2*>(od + v), make_double2(res[v], res[v + 1]));
            }
        }
        ot += cols;
        od += cols;
    }
}

template <typename T, bool DIFFUSE, int SEG>
__global__ void __launch_bounds__(128, 5)
k_box3(const T* __restrict__ in, const T* __restrict__ halo_top, const T* __restrict__ halo_bot,
       void* __restrict__ out_, int64_t rows, int64_t cols, int include_center, int col_torus, double rate) {
    constexpr int V = 16 / sizeof(T);
    const int64_t jv = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t j0 = jv * V;
    if (j0 >= cols) return;
    const int64_t i0 = (int64_t)blockIdx.y * SEG;
    const int64_t c0 = (int64_t)blockIdx.x * (128 * V);   // first column of the CTA
    const bool interior = i0 >= 1 && i0 + SEG + 1 <= rows && c0 >= 1 && c0 + 128 * V < cols;
    if (interior)
        box3_body<T, DIFFUSE, SEG, true>(in, halo_top, halo_bot, out_, rows, cols, include_center, col_torus, rate, j0, i0);
    else
        box3_body<T, DIFFUSE, SEG, false>(in, halo_top, halo_bot, out_, rows, cols, include_center, col_torus, rate, j0, i0);
}

template <typename T, typename A, int V, int MINB1, int MINB2>
bool launch_nsum_walk(const T* in, const T* ht, const T* hb, A* out, int64_t rows, int64_t cols, int radius,
                      int von_neumann, int include_center, int col_torus, cudaStream_t stream) {
    constexpr int SEG = 16;
    const dim3 grid((unsigned)mb::ceil_div(cols / V, 128), (unsigned)mb::ceil_div(rows, SEG));
    if (grid.y > 65535u) return false;
    if (radius == 1) {
        if (von_neumann) k_nsum_walk<T, A, 1, true, V, SEG, MINB1><<<grid, 128, 0, stream>>>(in, ht, hb, out, rows, cols, include_center, col_torus);
        else k_nsum_walk<T, A, 1, false, V, SEG, MINB1><<<grid, 128, 0, stream>>>(in, ht, hb, out, rows, cols, include_center, col_torus);
    } else {
        if (von_neumann) k_nsum_walk<T, A, 2, true, V, SEG, MINB2><<<grid, 128, 0, stream>>>(in, ht, hb, out, rows, cols, include_center, col_torus);
        else k_nsum_walk<T, A, 2, false, V, SEG, MINB2><<<grid, 128, 0, stream>>>(in, ht, hb, out, rows, cols, include_center, col_torus);
    }
    return true;
}

// ---- n-dimensional grids (grid.py:457-462 Moore = product([-1,0,1]^n), :488-501 von Neumann = +-1 per axis;
//      radius r = Chebyshev cube / Manhattan ball by the same BFS, cell.py:225-276) -------------------------------
// One thread per cell: decode the C-order coordinate, run an odometer over the (2r+1)^n box.  3-d and 4-d layers are
// small next to the 2-d ones of the benchmark configs; the loads are served by L1 / L2.
constexpr int kMaxDims = 4;
struct NdShape {
    int ndim;
    int64_t dim[kMaxDims];
    int64_t stride[kMaxDims];
};

template <typename T, typename A>
__global__ void k_neighborhood_sum_nd(const T* __restrict__ in, A* __restrict__ out, int64_t n, NdShape sh, int radius,
                                      int von_neumann, int include_center, int torus) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int64_t c[kMaxDims];
    {
        int64_t rem = idx;
        for (int d = 0; d < sh.ndim; ++d) {
            c[d] = rem / sh.stride[d];
            rem -= c[d] * sh.stride[d];
        }
    }
    int o[kMaxDims];
    for (int d = 0; d < sh.ndim; ++d) o[d] = -radius;
    A acc = 0;
    while (true) {
        int manhattan = 0;
        bool centre = true, inside = true;
        int64_t at = 0;
        for (int d = 0; d < sh.ndim; ++d) {
            manhattan += o[d] < 0 ? -o[d] : o[d];
            centre = centre && o[d] == 0;
            int64_t x = c[d] + o[d];
            if (x < 0 || x >= sh.dim[d]) {
                if (torus) x = (x % sh.dim[d] + sh.dim[d]) % sh.dim[d];
                else inside = false;
            }
            at += x * sh.stride[d];
        }
        if (inside && !(von_neumann && manhattan > radius) && !(centre && !include_center)) acc += (A)in[at];
        int d = sh.ndim - 1;      // next offset, last axis fastest (the oracle's order)
        while (d >= 0 && o[d] == radius) o[d--] = -radius;
        if (d < 0) break;
        ++o[d];
    }
    out[idx] = acc;
}

#define MB_DISPATCH_DTYPE(dtype, CALL)                                   \
    switch (dtype) {                                                     \
        case DT_U8: { using T = uint8_t; CALL; } break;                  \
        case DT_I32: { using T = int32_t; CALL; } break;                 \
        case DT_F32: { using T = float; CALL; } break;                   \
        case DT_F64: { using T = double; CALL; } break;                  \
        default: mb::set_error("unknown dtype %d", dtype); return MB_ERR_INVALID; \
    }

}  // namespace

MB_API int mb_layer_fill(void* data, int64_t n, int dtype, double value, cudaStream_t stream) {
    MB_REQUIRE(n >= 0, "mb_layer_fill: negative size");
    if (n == 0) return MB_OK;
    MB_REQUIRE(data != nullptr, "mb_layer_fill: null buffer");
    MB_DISPATCH_DTYPE(dtype, (k_fill<T><<<stream_blocks(n), 256, 0, stream>>>((T*)data, n, value)));
    MB_LAUNCH_CHECK("mb_layer_fill");
    return MB_OK;
}

MB_API int mb_layer_affine_clamp(const void* in, void* out, int64_t n, int dtype, double alpha, double beta,
                                 double lo, double hi, cudaStream_t stream) {
    MB_REQUIRE(n >= 0, "mb_layer_affine_clamp: negative size");
    if (n == 0) return MB_OK;
    MB_REQUIRE(in && out, "mb_layer_affine_clamp: null buffer");
    if (dtype == DT_U8) {
        k_affine_u8<<<stream_blocks(n / 16 + 1), 256, 0, stream>>>((const uint8_t*)in, (uint8_t*)out, n, alpha, beta, lo, hi);
    } else {
        MB_DISPATCH_DTYPE(dtype, (k_affine<T><<<stream_blocks(n / VecOf<T>::n + 1), 256, 0, stream>>>(
                                     (const T*)in, (T*)out, n, alpha, beta, lo, hi)));
    }
    MB_LAUNCH_CHECK("mb_layer_affine_clamp");
    return MB_OK;
}

MB_API int mb_layer_binary(const void* a, const void* b, void* out, int64_t n, int dtype, int op,
                           cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && op >= OP_ADD && op <= OP_MAX, "mb_layer_binary: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(a && b && out, "mb_layer_binary: null buffer");
    MB_DISPATCH_DTYPE(dtype, (launch_binary<T>((const T*)a, (const T*)b, (T*)out, n, op,
                                               stream_blocks(n / VecOf<T>::n + 1), stream)));
    MB_LAUNCH_CHECK("mb_layer_binary");
    return MB_OK;
}

// result: int64 for u8/i32 layers, float64 for f32/f64 layers (count_nonzero is always int64).
// workspace: 1024 * 8 bytes.
MB_API int mb_layer_reduce(const void* in, int64_t n, int dtype, int op, void* result, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && op >= RED_SUM && op <= RED_COUNT_NONZERO, "mb_layer_reduce: bad argument");
    MB_REQUIRE(result && workspace && (n == 0 || in), "mb_layer_reduce: null buffer");
    MB_REQUIRE(n > 0 || op == RED_SUM || op == RED_COUNT_NONZERO, "mb_layer_reduce: min/max of an empty layer");
    if (workspace_bytes < kPartials * 8) {
        mb::set_error("mb_layer_reduce: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    int parts = (int)(mb::ceil_div(n > 0 ? n : 1, 4096) < kPartials ? mb::ceil_div(n > 0 ? n : 1, 4096) : kPartials);
    const bool as_int = dtype == DT_U8 || dtype == DT_I32 || op == RED_COUNT_NONZERO;
    if (as_int) {
        using A = long long;
        MB_DISPATCH_DTYPE(dtype, (launch_reduce1<T, A>((const T*)in, n, op, (A*)workspace, parts, stream)));
        k_reduce2<A><<<1, 1024, 0, stream>>>((const A*)workspace, parts, op == RED_COUNT_NONZERO ? RED_SUM : op, (A*)result);
    } else {
        using A = double;
        MB_DISPATCH_DTYPE(dtype, (launch_reduce1<T, A>((const T*)in, n, op, (A*)workspace, parts, stream)));
        k_reduce2<A><<<1, 1024, 0, stream>>>((const A*)workspace, parts, op, (A*)result);
    }
    MB_LAUNCH_CHECK("mb_layer_reduce");
    return MB_OK;
}

// out dtype: int32 for u8/i32 inputs, float64 for f32/f64 inputs.
MB_API int mb_neighborhood_sum(const void* in, const void* halo_top, const void* halo_bot, void* out, int64_t rows,
                               int64_t cols, int dtype, int radius, int von_neumann, int include_center,
                               int col_torus, cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0, "mb_neighborhood_sum: negative shape");
    MB_REQUIRE(radius >= 1, "radius must be at least 1");  // cell.py:234-235
    MB_REQUIRE(!(col_torus && cols > 0 && cols < 2 * radius + 1),
               "mb_neighborhood_sum: torus needs cols >= 2*radius+1 (SURVEY A.2)");
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(in && out, "mb_neighborhood_sum: null buffer");
    if (radius <= 2 && dtype >= DT_U8 && dtype <= DT_F64) {
        // column walker (see above); uint8 rows must be 4-byte aligned for the 4-column version
        const uintptr_t align = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(halo_top) |
                                reinterpret_cast<uintptr_t>(halo_bot);
        bool done = false;
        if (dtype == DT_U8 && cols % 4 == 0 && (align & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
            done = launch_nsum_walk<uint8_t, int32_t, 4, 8, 4>((const uint8_t*)in, (const uint8_t*)halo_top, (const uint8_t*)halo_bot, (int32_t*)out, rows, cols, radius, von_neumann, include_center, col_torus, stream);
        else if (dtype == DT_U8)
            done = launch_nsum_walk<uint8_t, int32_t, 1, 8, 8>((const uint8_t*)in, (const uint8_t*)halo_top, (const uint8_t*)halo_bot, (int32_t*)out, rows, cols, radius, von_neumann, include_center, col_torus, stream);
        else if (dtype == DT_I32)
            done = launch_nsum_walk<int32_t, int32_t, 1, 8, 8>((const int32_t*)in, (const int32_t*)halo_top, (const int32_t*)halo_bot, (int32_t*)out, rows, cols, radius, von_neumann, include_center, col_torus, stream);
        else if (dtype == DT_F32 && radius == 1 && !von_neumann && cols % 4 == 0 && (align & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0 && mb::ceil_div(rows, 16) <= 65535) {
            const dim3 grid((unsigned)mb::ceil_div(cols / 4, 128), (unsigned)mb::ceil_div(rows, 16));
            k_box3<float, false, 16><<<grid, 128, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, out, rows, cols, include_center, col_torus, 0.0);
            done = true;
        } else if (dtype == DT_F32)
            done = launch_nsum_walk<float, double, 1, 8, 4>((const float*)in, (const float*)halo_top, (const float*)halo_bot, (double*)out, rows, cols, radius, von_neumann, include_center, col_torus, stream);
        else
            done = launch_nsum_walk<double, double, 1, 8, 4>((const double*)in, (const double*)halo_top, (const double*)halo_bot, (double*)out, rows, cols, radius, von_neumann, include_center, col_torus, stream);
        if (done) {
            MB_LAUNCH_CHECK("mb_neighborhood_sum");
            return MB_OK;
        }
    }
    const unsigned blocks = (unsigned)mb::ceil_div(rows * cols, 256);
    switch (dtype) {
        case DT_U8: k_neighborhood_sum<uint8_t, int32_t><<<blocks, 256, 0, stream>>>((const uint8_t*)in, (const uint8_t*)halo_top, (const uint8_t*)halo_bot, (int32_t*)out, rows, cols, radius, von_neumann, include_center, col_torus); break;
        case DT_I32: k_neighborhood_sum<int32_t, int32_t><<<blocks, 256, 0, stream>>>((const int32_t*)in, (const int32_t*)halo_top, (const int32_t*)halo_bot, (int32_t*)out, rows, cols, radius, von_neumann, include_center, col_torus); break;
        case DT_F32: k_neighborhood_sum<float, double><<<blocks, 256, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, (double*)out, rows, cols, radius, von_neumann, include_center, col_torus); break;
        case DT_F64: k_neighborhood_sum<double, double><<<blocks, 256, 0, stream>>>((const double*)in, (const double*)halo_top, (const double*)halo_bot, (double*)out, rows, cols, radius, von_neumann, include_center, col_torus); break;
        default: mb::set_error("unknown dtype %d", dtype); return MB_ERR_INVALID;
    }
    MB_LAUNCH_CHECK("mb_neighborhood_sum");
    return MB_OK;
}

MB_API int mb_layer_diffuse(const void* in, const void* halo_top, const void* halo_bot, void* out, int64_t rows,
                            int64_t cols, int dtype, double rate, int col_torus, cudaStream_t stream) {
    MB_REQUIRE(rows >= 0 && cols >= 0, "mb_layer_diffuse: negative shape");
    MB_REQUIRE(rate >= 0.0 && rate <= 1.0, "mb_layer_diffuse: rate must be in [0, 1]");
    MB_REQUIRE(dtype == DT_F32 || dtype == DT_F64, "mb_layer_diffuse: float32 / float64 layers only");
    MB_REQUIRE(!(col_torus && cols > 0 && cols < 3), "mb_layer_diffuse: torus needs cols >= 3");
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(in && out && in != out, "mb_layer_diffuse: null or aliased buffer");
    {
        constexpr int SEG = 16;
        const dim3 grid((unsigned)mb::ceil_div(cols, 128), (unsigned)mb::ceil_div(rows, SEG));
        const uintptr_t align = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(halo_top) |
                                reinterpret_cast<uintptr_t>(halo_bot) | reinterpret_cast<uintptr_t>(out);
        const int vcols = dtype == DT_F32 ? 4 : 2;   // columns per thread of the box kernel
        if (grid.y <= 65535u && cols % vcols == 0 && (align & 15) == 0) {
            const dim3 gridv((unsigned)mb::ceil_div(cols / vcols, 128), grid.y);
            if (dtype == DT_F32)
                k_box3<float, true, SEG><<<gridv, 128, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, out, rows, cols, 0, col_torus, rate);
            else
                k_box3<double, true, SEG><<<gridv, 128, 0, stream>>>((const double*)in, (const double*)halo_top, (const double*)halo_bot, out, rows, cols, 0, col_torus, rate);
            MB_LAUNCH_CHECK("mb_layer_diffuse");
            return MB_OK;
        }
        if (grid.y <= 65535u) {
            if (dtype == DT_F32)
                k_diffuse_walk<float, SEG><<<grid, 128, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, (float*)out, rows, cols, rate, col_torus);
            else
                k_diffuse_walk<double, SEG><<<grid, 128, 0, stream>>>((const double*)in, (const double*)halo_top, (const double*)halo_bot, (double*)out, rows, cols, rate, col_torus);
            MB_LAUNCH_CHECK("mb_layer_diffuse");
            return MB_OK;
        }
    }
    const unsigned blocks = (unsigned)mb::ceil_div(rows * cols, 256);
    if (dtype == DT_F32)
        k_diffuse<float><<<blocks, 256, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, (float*)out, rows, cols, rate, col_torus);
    else
        k_diffuse<double><<<blocks, 256, 0, stream>>>((const double*)in, (const double*)halo_top, (const double*)halo_bot, (double*)out, rows, cols, rate, col_torus);
    MB_LAUNCH_CHECK("mb_layer_diffuse");
    return MB_OK;
}

// Neighbourhood sums on an n-dimensional layer (1 <= ndim <= 4, C order).  out dtype as mb_neighborhood_sum.
MB_API int mb_neighborhood_sum_nd(const void* in, void* out, int ndim, const int64_t* dims, int dtype, int radius,
                                  int von_neumann, int include_center, int torus, cudaStream_t stream) {
    MB_REQUIRE(ndim >= 1 && ndim <= kMaxDims && dims != nullptr, "mb_neighborhood_sum_nd: 1 to 4 dimensions");
    MB_REQUIRE(radius >= 1, "radius must be at least 1");  // cell.py:234-235
    NdShape sh;
    sh.ndim = ndim;
    int64_t n = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        MB_REQUIRE(dims[d] >= 0, "mb_neighborhood_sum_nd: negative shape");
        MB_REQUIRE(!(torus && dims[d] > 0 && dims[d] < 2 * radius + 1),
                   "mb_neighborhood_sum_nd: torus needs every dimension >= 2*radius+1 (SURVEY A.2)");
        sh.dim[d] = dims[d];
        sh.stride[d] = n;
        n *= dims[d];
    }
    for (int d = ndim; d < kMaxDims; ++d) { sh.dim[d] = 1; sh.stride[d] = 0; }
    if (n == 0) return MB_OK;
    MB_REQUIRE(in && out, "mb_neighborhood_sum_nd: null buffer");
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    switch (dtype) {
        case DT_U8: k_neighborhood_sum_nd<uint8_t, int32_t><<<blocks, 256, 0, stream>>>((const uint8_t*)in, (int32_t*)out, n, sh, radius, von_neumann, include_center, torus); break;
        case DT_I32: k_neighborhood_sum_nd<int32_t, int32_t><<<blocks, 256, 0, stream>>>((const int32_t*)in, (int32_t*)out, n, sh, radius, von_neumann, include_center, torus); break;
        case DT_F32: k_neighborhood_sum_nd<float, double><<<blocks, 256, 0, stream>>>((const float*)in, (double*)out, n, sh, radius, von_neumann, include_center, torus); break;
        case DT_F64: k_neighborhood_sum_nd<double, double><<<blocks, 256, 0, stream>>>((const double*)in, (double*)out, n, sh, radius, von_neumann, include_center, torus); break;
        default: mb::set_error("unknown dtype %d", dtype); return MB_ERR_INVALID;
    }
    MB_LAUNCH_CHECK("mb_neighborhood_sum_nd");
    return MB_OK;
}
