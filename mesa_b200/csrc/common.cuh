// Shared helpers for the mesa_b200 CUDA kernels (sm_100a only).
//
// Conventions of the C ABI (include/mesa_b200.h):
//   * every entry point returns 0 on success, a negative MB_ERR_* code otherwise;
//   * the message for the last failure on the calling thread is mb_last_error();
//   * the library never allocates, frees or retains device memory: every buffer
//     (state, scratch, halos) is owned by the caller (a torch.Tensor);
//   * kernels are enqueued on the caller's stream, no implicit synchronisation.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define MB_API extern "C" __attribute__((visibility("default")))

#define MB_OK 0
#define MB_ERR_INVALID (-1)   // bad argument (maps to ValueError in the Python shim)
#define MB_ERR_CUDA (-2)      // CUDA runtime error
#define MB_ERR_WORKSPACE (-3) // workspace too small
#define MB_ERR_UNSUPPORTED (-4)

namespace mb {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MB_ERR_CUDA;
    }
    return MB_OK;
}

// Number of SMs of the current device (148 on B200); cached per process.
int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mb

#define MB_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            mb::set_error(__VA_ARGS__);  \
            return MB_ERR_INVALID;       \
        }                                \
    } while (0)

#define MB_LAUNCH_CHECK(what)                                  \
    do {                                                       \
        int _rc = mb::check_cuda(cudaGetLastError(), what);    \
        if (_rc != MB_OK) return _rc;                          \
    } while (0)

// ---- device helpers -------------------------------------------------------

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_na_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ unsigned warp_reduce_add(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
