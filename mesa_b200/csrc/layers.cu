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
// paths, grid-stride over a grid sized to the SM count.  Reductions are two-pass with a fixed
// partial count so floating-point sums are reproducible run to run.
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
    const T v = from_double<T>(value);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
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
        for (int64_t i = tid; i < nvec; i += stride) {
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

template <typename T> __device__ __forceinline__ T apply_op(T a, T b, int op) {
    switch (op) {
        case OP_ADD: return (T)(a + b);
        case OP_SUB: return (T)(a - b);
        case OP_MUL: return (T)(a * b);
        case OP_MIN: return a < b ? a : b;
        default: return a > b ? a : b;
    }
}

template <typename T>
__global__ void k_binary(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n, int op) {
    constexpr int V = VecOf<T>::n;
    const int64_t nvec = n / V;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                           reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned) {
        for (int64_t i = tid; i < nvec; i += stride) {
            uint4 ra = __ldg(reinterpret_cast<const uint4*>(a) + i);
            uint4 rb = __ldg(reinterpret_cast<const uint4*>(b) + i);
            T* ea = reinterpret_cast<T*>(&ra);
            const T* eb = reinterpret_cast<const T*>(&rb);
#pragma unroll
            for (int k = 0; k < V; ++k) ea[k] = apply_op<T>(ea[k], eb[k], op);
            reinterpret_cast<uint4*>(out)[i] = ra;
        }
        for (int64_t i = nvec * V + tid; i < n; i += stride) out[i] = apply_op<T>(a[i], b[i], op);
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] = apply_op<T>(a[i], b[i], op);
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

template <typename T, typename A>
__global__ void __launch_bounds__(256) k_reduce1(const T* __restrict__ in, int64_t n, int op, A* __restrict__ partial) {
    // block b owns the contiguous chunk [b*chunk, (b+1)*chunk): fixed order => reproducible
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = lo + chunk < n ? lo + chunk : n;
    A acc = red_identity<A>(op);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const T x = in[i];
        const A v = op == RED_COUNT_NONZERO ? (A)(x != (T)0) : (A)x;
        acc = red_combine<A>(acc, v, op);
    }
    acc = block_reduce<A>(acc, op);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
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
    out[idx] = (T)(self * (1.0 - rate * (double)k / 8.0) + rate / 8.0 * acc);
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
    MB_DISPATCH_DTYPE(dtype, (k_affine<T><<<stream_blocks(n / VecOf<T>::n + 1), 256, 0, stream>>>(
                                 (const T*)in, (T*)out, n, alpha, beta, lo, hi)));
    MB_LAUNCH_CHECK("mb_layer_affine_clamp");
    return MB_OK;
}

MB_API int mb_layer_binary(const void* a, const void* b, void* out, int64_t n, int dtype, int op,
                           cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && op >= OP_ADD && op <= OP_MAX, "mb_layer_binary: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(a && b && out, "mb_layer_binary: null buffer");
    MB_DISPATCH_DTYPE(dtype, (k_binary<T><<<stream_blocks(n / VecOf<T>::n + 1), 256, 0, stream>>>(
                                 (const T*)a, (const T*)b, (T*)out, n, op)));
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
        MB_DISPATCH_DTYPE(dtype, (k_reduce1<T, A><<<parts, 256, 0, stream>>>((const T*)in, n, op, (A*)workspace)));
        k_reduce2<A><<<1, 1024, 0, stream>>>((const A*)workspace, parts, op == RED_COUNT_NONZERO ? RED_SUM : op, (A*)result);
    } else {
        using A = double;
        MB_DISPATCH_DTYPE(dtype, (k_reduce1<T, A><<<parts, 256, 0, stream>>>((const T*)in, n, op, (A*)workspace)));
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
    const unsigned blocks = (unsigned)mb::ceil_div(rows * cols, 256);
    if (dtype == DT_F32)
        k_diffuse<float><<<blocks, 256, 0, stream>>>((const float*)in, (const float*)halo_top, (const float*)halo_bot, (float*)out, rows, cols, rate, col_torus);
    else
        k_diffuse<double><<<blocks, 256, 0, stream>>>((const double*)in, (const double*)halo_top, (const double*)halo_bot, (double*)out, rows, cols, rate, col_torus);
    MB_LAUNCH_CHECK("mb_layer_diffuse");
    return MB_OK;
}
