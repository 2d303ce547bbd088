// Deep-ghost-zone refresh of a row-sharded layer WITHOUT a barrier or a collective (grid
// stencils; the reference is single-process and has no counterpart).
//
// Every rank keeps `g` rows of each ring neighbour around its row block and steps the extended block; once per `g`
// generations the ghost rows are refreshed from the neighbours' HBM over NVLink (peer-mapped pointers).  The first
// version bracketed that copy with two device-wide symmetric-memory barriers (~45 us each at 8 GPUs).  Here the
// hand-shake is two monotone counters per neighbour, living in the RECEIVER's memory so that every poll is a local load:
//
//   ready[side]   written by the neighbour on that side: "my plane holds exchange #s, come and get it"
//   copied[side]  written by the neighbour on that side: "I have copied my ghost rows of exchange #s out of your plane"
//
// mb_ghost_exchange (one launch, a few CTAs):  s = seq + 1;  publish ready = s to both neighbours;  wait until both
// neighbours' ready >= s;  copy their edge rows into my ghost rows (16-byte vectors over NVLink);  the last CTA sets
// seq = s and publishes copied = s to both neighbours.
// mb_ghost_wait_copied (one warp): before the launch that OVERWRITES the plane the neighbours copied from, wait
// until both neighbours' copied >= seq.  That launch comes one whole stencil launch after the exchange, so the wait is
// normally already satisfied; no rank ever waits for a rank that is not already past its own publish, so the protocol
// cannot deadlock, and a wait that exceeds ~seconds sets an error word instead of hanging the GPU.
//
// The exchange number lives on the device (`seq`), not in a kernel argument: the launches are identical every
// time and can be captured in a CUDA graph and replayed.
#include "common.cuh"

namespace {

// flag block layout (uint32 words), one block per rank in peer-mapped memory
enum { kReadyUp = 0, kReadyDown = 1, kCopiedUp = 2, kCopiedDown = 3, kSeq = 4, kDone = 5, kError = 6, kFlagWords = 8 };

__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {
    uint4 r;   // relaxed system-scope load: never served from a stale L1 line
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}

__device__ __forceinline__ void wait_at_least(const unsigned* flag, unsigned want, unsigned* error) {
    unsigned spins = 0;
    while ((int)(ld_acquire_sys_u32(flag) - want) < 0) {
        if (++spins > (1u << 25)) {   // ~ seconds: a neighbour died; do not hang the GPU
            *error = 1u;
            break;
        }
        __nanosleep(64);
    }
}

__global__ void __launch_bounds__(256)
ghost_exchange(uint4* __restrict__ dst_top, const uint4* __restrict__ src_up, uint4* __restrict__ dst_bot,
               const uint4* __restrict__ src_down, int64_t vecs, unsigned* flags, unsigned* up_flags,
               unsigned* down_flags) {
    __shared__ unsigned s_seq;
    if (threadIdx.x == 0) {
        const unsigned s = *(volatile unsigned*)&flags[kSeq] + 1u;
        if (blockIdx.x == 0) {
            // the stencil launch before this one is complete (stream order); make its rows visible system-wide
            __threadfence_system();
            if (up_flags) st_release_sys_u32(&up_flags[kReadyDown], s);      // I am the upper neighbour's DOWN side
            if (down_flags) st_release_sys_u32(&down_flags[kReadyUp], s);
        }
        if (up_flags) wait_at_least(&flags[kReadyUp], s, &flags[kError]);
        if (down_flags) wait_at_least(&flags[kReadyDown], s, &flags[kError]);
        s_seq = s;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += stride) {
        if (src_up) dst_top[i] = ld_peer_v4(src_up + i);
        if (src_down) dst_bot[i] = ld_peer_v4(src_down + i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&flags[kDone], 1u) == gridDim.x - 1u) {
            flags[kDone] = 0u;
            *(volatile unsigned*)&flags[kSeq] = s_seq;
            __threadfence_system();
            if (up_flags) st_release_sys_u32(&up_flags[kCopiedDown], s_seq);
            if (down_flags) st_release_sys_u32(&down_flags[kCopiedUp], s_seq);
        }
    }
}

__global__ void ghost_wait_copied(unsigned* flags, int has_up, int has_down) {
    if (threadIdx.x == 0) {
        const unsigned s = *(volatile unsigned*)&flags[kSeq];
        if (has_up) wait_at_least(&flags[kCopiedUp], s, &flags[kError]);
        if (has_down) wait_at_least(&flags[kCopiedDown], s, &flags[kError]);
    }
}

}  // namespace

// Refresh the ghost rows of one plane.  dst_top / dst_bot: this rank's ghost rows above / below its block; src_up /
// src_down: PEER-MAPPED pointers to the rows to copy there (the upper neighbour's last owned rows, the lower
// neighbour's first owned rows; NULL on a clamped outer edge); `bytes` per side, a multiple of 16 with 16-byte
// aligned pointers.  flags: this rank's 8-word flag block (zero-initialised, peer-mapped memory); up_flags /
// down_flags: the peer-mapped addresses of the neighbours' flag blocks (NULL where src_* is NULL).
MB_API int mb_ghost_exchange(void* dst_top, const void* src_up, void* dst_bot, const void* src_down, int64_t bytes,
                             unsigned* flags, unsigned* up_flags, unsigned* down_flags, cudaStream_t stream) {
    MB_REQUIRE(flags != nullptr && bytes >= 0 && bytes % 16 == 0, "mb_ghost_exchange: bytes must be a multiple of 16");
    MB_REQUIRE((src_up == nullptr) == (up_flags == nullptr) && (src_down == nullptr) == (down_flags == nullptr),
               "mb_ghost_exchange: a neighbour needs both its rows and its flag block");
    MB_REQUIRE((!src_up || dst_top) && (!src_down || dst_bot), "mb_ghost_exchange: null ghost rows");
    MB_REQUIRE(((uintptr_t)dst_top | (uintptr_t)src_up | (uintptr_t)dst_bot | (uintptr_t)src_down) % 16 == 0,
               "mb_ghost_exchange: pointers must be 16-byte aligned");
    if (!src_up && !src_down) return MB_OK;
    const int64_t vecs = bytes / 16;
    int64_t blocks = mb::ceil_div(vecs > 0 ? vecs : 1, 256 * 4);
    if (blocks > 64) blocks = 64;   // 64 CTAs x 256 threads x 16 B in flight saturate one NVLink direction
    ghost_exchange<<<(unsigned)blocks, 256, 0, stream>>>((uint4*)dst_top, (const uint4*)src_up, (uint4*)dst_bot,
                                                         (const uint4*)src_down, vecs, flags, up_flags, down_flags);
    MB_LAUNCH_CHECK("mb_ghost_exchange");
    return MB_OK;
}

// Wait until the neighbours have copied their ghost rows of the latest exchange out of this rank's plane (call
// before the launch that overwrites that plane).
MB_API int mb_ghost_wait_copied(unsigned* flags, int has_up, int has_down, cudaStream_t stream) {
    MB_REQUIRE(flags != nullptr, "mb_ghost_wait_copied: null flag block");
    if (!has_up && !has_down) return MB_OK;
    ghost_wait_copied<<<1, 32, 0, stream>>>(flags, has_up, has_down);
    MB_LAUNCH_CHECK("mb_ghost_wait_copied");
    return MB_OK;
}
