// Domain-decomposed continuous space: the exchange steps of the sharded Boids model as
// kernels that write straight into the neighbouring ranks' memory (peer-mapped, NVLink), instead of
// pack -> NCCL send/recv -> unpack.
//
//   mb_boids_send_ghosts   boids within `reach` of a slab edge are appended to the ghost inbox of the rank
//                          that owns the other side of the edge (atomic slot counter in ITS memory);
//   mb_boids_route         after the step: boids whose y left the slab are appended to the migrant inbox of
//                          their new owner (with their global id); the others get stay = 1;
//   mb_boids_compact       stable compaction of the stayers (flags -> exclusive scan -> scatter).
//
// The reference has no counterpart: one process, one positions array (continuous_space.py:85-101).
// Inbox order depends on atomics; it only changes the order in which a boid's neighbours are summed in
// float64 (results agree to the last float32 bit except for rare ties) -- the driver sorts the small
// inboxes by id where run-to-run reproducibility is wanted.
#include "common.cuh"
#include "sort.cuh"

namespace {

struct Inbox {
    float2* pos;
    float2* dir;
    long long* ids;          // may be NULL (ghosts carry no identity)
    unsigned* count;
};

__device__ __forceinline__ void inbox_push(const Inbox& in, long long cap, float2 p, float2 d, long long id,
                                           unsigned* overflow) {
    const unsigned slot = atomicAdd(in.count, 1u);
    if ((long long)slot >= cap) {
        *overflow = 1u;
        return;
    }
    in.pos[slot] = p;
    in.dir[slot] = d;
    if (in.ids != nullptr) in.ids[slot] = id;
}

__global__ void send_ghosts(const float2* __restrict__ pos, const float2* __restrict__ dir, int64_t n, double y_lo,
                            double y_hi, double reach, Inbox down, Inbox up, long long cap, unsigned* overflow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = pos[i];
    const double y = (double)p.y;
    if (down.count != nullptr && y < y_lo + reach) inbox_push(down, cap, p, dir[i], 0, overflow);
    if (up.count != nullptr && y >= y_hi - reach) inbox_push(up, cap, p, dir[i], 0, overflow);
}

// slab index of y (the top edge y == height belongs to the last slab), as ShardedBoids.owner_of
__device__ __forceinline__ int slab_of(double y, double height, int world) {
    int s = (int)floor(y * (double)world / height);
    return s < 0 ? 0 : (s >= world ? world - 1 : s);
}

__global__ void route(const float2* __restrict__ pos, const float2* __restrict__ dir, const long long* __restrict__ ids,
                      int64_t n, double height, int world, int rank, unsigned* __restrict__ stay, Inbox down, Inbox up,
                      long long cap, unsigned* flags /* [0] overflow, [1] jumped over a slab */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = pos[i];
    const int owner = slab_of((double)p.y, height, world);
    unsigned keep = 1u;
    if (owner != rank) {
        keep = 0u;
        const int dn = (rank + world - 1) % world, upr = (rank + 1) % world;
        if (owner == dn && down.count != nullptr) inbox_push(down, cap, p, dir[i], ids[i], flags);
        else if (owner == upr && up.count != nullptr) inbox_push(up, cap, p, dir[i], ids[i], flags);
        else flags[1] = 1u;
    }
    stay[i] = keep;
}

__global__ void compact(const float2* __restrict__ pos, const float2* __restrict__ dir, const long long* __restrict__ ids,
                        int64_t n, const unsigned* __restrict__ stay, const unsigned* __restrict__ offs,
                        float2* __restrict__ pos_out, float2* __restrict__ dir_out, long long* __restrict__ ids_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !stay[i]) return;
    const unsigned o = offs[i];
    pos_out[o] = pos[i];
    dir_out[o] = dir[i];
    ids_out[o] = ids[i];
}

__global__ void copy_flags(const unsigned* __restrict__ a, unsigned* __restrict__ b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}

}  // namespace

MB_API int mb_boids_send_ghosts(const float* pos, const float* dir, int64_t n, double y_lo, double y_hi, double reach,
                                float* down_pos, float* down_dir, unsigned* down_count, float* up_pos, float* up_dir,
                                unsigned* up_count, int64_t cap, unsigned* overflow, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && cap >= 0 && overflow != nullptr, "mb_boids_send_ghosts: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos && dir, "mb_boids_send_ghosts: null buffer");
    const Inbox down{(float2*)down_pos, (float2*)down_dir, nullptr, down_count};
    const Inbox up{(float2*)up_pos, (float2*)up_dir, nullptr, up_count};
    send_ghosts<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>((const float2*)pos, (const float2*)dir, n, y_lo, y_hi,
                                                                  reach, down, up, cap, overflow);
    MB_LAUNCH_CHECK("mb_boids_send_ghosts");
    return MB_OK;
}

// stay_offsets[i] = rank of boid i among the stayers (exclusive scan of the stay flags); stay_flags[i] = 0/1.
// scratch: mb_scan_scratch_words(n) uint32 words.  The number of stayers is stay_offsets[n-1] + stay_flags[n-1].
MB_API int mb_boids_route(const float* pos, const float* dir, const int64_t* ids, int64_t n, double height, int world,
                          int rank, unsigned* stay_flags, unsigned* stay_offsets, unsigned* scratch, float* down_pos,
                          float* down_dir, int64_t* down_ids, unsigned* down_count, float* up_pos, float* up_dir,
                          int64_t* up_ids, unsigned* up_count, int64_t cap, unsigned* flags, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && world >= 1 && rank >= 0 && rank < world && flags != nullptr, "mb_boids_route: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos && dir && ids && stay_flags && stay_offsets && scratch, "mb_boids_route: null buffer");
    const Inbox down{(float2*)down_pos, (float2*)down_dir, (long long*)down_ids, down_count};
    const Inbox up{(float2*)up_pos, (float2*)up_dir, (long long*)up_ids, up_count};
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    route<<<blocks, 256, 0, stream>>>((const float2*)pos, (const float2*)dir, (const long long*)ids, n, height, world, rank,
                                    stay_flags, down, up, cap, flags);
    copy_flags<<<blocks, 256, 0, stream>>>(stay_flags, stay_offsets, n);
    MB_LAUNCH_CHECK("mb_boids_route");
    return mb::exclusive_scan_u32(stay_offsets, n, scratch, stream);
}

MB_API int mb_boids_compact(const float* pos, const float* dir, const int64_t* ids, int64_t n, const unsigned* stay_flags,
                            const unsigned* stay_offsets, float* pos_out, float* dir_out, int64_t* ids_out,
                            cudaStream_t stream) {
    MB_REQUIRE(n >= 0, "mb_boids_compact: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos && dir && ids && stay_flags && stay_offsets && pos_out && dir_out && ids_out, "mb_boids_compact: null buffer");
    compact<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>((const float2*)pos, (const float2*)dir, (const long long*)ids, n,
                                                              stay_flags, stay_offsets, (float2*)pos_out, (float2*)dir_out,
                                                              (long long*)ids_out);
    MB_LAUNCH_CHECK("mb_boids_compact");
    return MB_OK;
}

MB_API int mb_scan_scratch_words(int64_t n, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0, "mb_scan_scratch_words: bad argument");
    *out = mb::scan_scratch_words(n);
    return MB_OK;
}
