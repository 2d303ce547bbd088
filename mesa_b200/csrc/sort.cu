// Stable LSD radix sort of (key, uint32) pairs -- see sort.cuh.
#include "sort.cuh"
#include <stdlib.h>

namespace mb {
namespace {

constexpr int kWarpsPerCta = kSortThreads / 32;

template <typename K>
__global__ void __launch_bounds__(kSortThreads)
sort_hist(const K* __restrict__ keys, int64_t n, int bit, uint32_t* __restrict__ hist, int64_t tiles) {
    __shared__ unsigned sh[kRadix];
    sh[threadIdx.x] = 0u;  // kSortThreads == kRadix
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = base + (int64_t)i * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(unsigned)(keys[idx] >> bit) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan of `total` uint32 counters in place, three small kernels:
//   scan_chunks  each CTA scans its 4096-counter chunk in place and writes the chunk sum
//   scan_top     one CTA turns the chunk sums into exclusive offsets (<= 4096*... chunks, looped)
//   scan_add     each CTA adds its chunk offset
constexpr int kScanChunk = 4096;

__device__ __forceinline__ unsigned block_excl_scan_1024(unsigned v, unsigned* warp_sum, unsigned& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warp_sum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    block_total = warp_sum[31];
    return (warp > 0 ? warp_sum[warp - 1] : 0u) + incl - v;
}

__global__ void __launch_bounds__(1024) scan_chunks(uint32_t* __restrict__ data, int64_t total, uint32_t* __restrict__ sums) {
    __shared__ unsigned warp_sum[32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < total) ? data[i0 + k] : 0u;
    unsigned block_total;
    unsigned excl = block_excl_scan_1024(v[0] + v[1] + v[2] + v[3], warp_sum, block_total);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < total) data[i0 + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = block_total;
}

__global__ void __launch_bounds__(1024) scan_top(uint32_t* __restrict__ sums, int64_t nchunks) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int64_t c0 = 0; c0 < nchunks; c0 += 1024) {
        const int64_t i = c0 + threadIdx.x;
        const unsigned v = i < nchunks ? sums[i] : 0u;
        unsigned block_total;
        const unsigned excl = block_excl_scan_1024(v, warp_sum, block_total);
        if (i < nchunks) sums[i] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry += block_total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) scan_add(uint32_t* __restrict__ data, int64_t total, const uint32_t* __restrict__ sums) {
    const unsigned off = sums[blockIdx.x];
    const int64_t i0 = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k < total) data[i0 + k] += off;
}

// Scatter pass.  Every element's slot inside its digit is found with warp multi-split (match_any) + per-warp digit
// counts, as in the histogram order (sub-chunk, warp, lane) = input order, so the sort is stable.  The tile is then
// REORDERED IN SHARED MEMORY by digit before it is written: consecutive threads store consecutive elements of one
// digit, i.e. runs of ~tile/256 elements go to consecutive addresses instead of one 4-byte store per 32-byte sector
// (the direct scatter reached 0.86 TB/s on 10^7 pairs).
template <typename K>
__global__ void __launch_bounds__(kSortThreads)
sort_scatter(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out, int64_t n, int bit, const uint32_t* __restrict__ hist,
             int64_t tiles) {
    __shared__ unsigned gbase[kRadix];                  // global slot of this tile's first element of each digit
    __shared__ unsigned lbase[kRadix];                  // position of that element in the reordered tile
    __shared__ unsigned run[kRadix];                    // elements of each digit seen in earlier sub-chunks
    __shared__ unsigned wcount[kWarpsPerCta][kRadix];   // per-warp digit counts of the current sub-chunk
    __shared__ unsigned wsum[kWarpsPerCta];
    extern __shared__ __align__(16) unsigned char tile_raw[];
    K* skey = reinterpret_cast<K*>(tile_raw);
    uint32_t* sval = reinterpret_cast<uint32_t*>(tile_raw + sizeof(K) * kSortTile);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    gbase[tid] = hist[(int64_t)tid * tiles + blockIdx.x];
    run[tid] = 0u;
    const int64_t tile0 = (int64_t)blockIdx.x * kSortTile;
    const int tile_n = (int)(n - tile0 < kSortTile ? n - tile0 : kSortTile);
    K key[kSortItems];
    uint32_t val[kSortItems];
    unsigned where[kSortItems];  // digit | slot within the digit << 8
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int s = i * kSortThreads + tid;
        key[i] = s < tile_n ? keys_in[tile0 + s] : (K)0;
        val[i] = s < tile_n ? vals_in[tile0 + s] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) wcount[w][tid] = 0u;
        __syncthreads();
        const bool valid = i * kSortThreads + tid < tile_n;
        const unsigned digit = valid ? (unsigned)(key[i] >> bit) & 0xffu : 0xffffffffu;  // invalid lanes match nobody real
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcount[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned slot = run[digit] + rank;
            for (int w = 0; w < warp; ++w) slot += wcount[w][digit];
            where[i] = digit | (slot << 8);
        }
        __syncthreads();
        unsigned add = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) add += wcount[w][tid];
        run[tid] += add;
    }
    __syncthreads();
    {   // lbase = exclusive scan of the tile's digit counts (256 counters, one per thread)
        const unsigned c = run[tid];
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        lbase[tid] = off + incl - c;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        if (i * kSortThreads + tid < tile_n) {
            const unsigned pos = lbase[where[i] & 0xffu] + (where[i] >> 8);
            skey[pos] = key[i];
            sval[pos] = val[i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int s = i * kSortThreads + tid;
        if (s < tile_n) {
            const K k = skey[s];
            const unsigned digit = (unsigned)(k >> bit) & 0xffu;
            const unsigned dst = gbase[digit] + ((unsigned)s - lbase[digit]);
            keys_out[dst] = k;
            vals_out[dst] = sval[s];
        }
    }
}


// ---- ONE SWEEP per digit (decoupled look-back) ----------------------------------------------------------------------------
// The three launches per digit above (tile histograms, their scan -- itself three launches --, scatter) read the keys twice
// and cost five launch gaps per digit: at 10^6 keys (Boltzmann: one sort per step) the gaps ARE the sort.  Here the digit
// histograms of ALL passes come from one read of the keys (sort_hist_all), and a pass is ONE kernel: a CTA takes its tile
// from a ticket counter (so every tile with a lower number is already running or done), ranks its keys exactly as
// sort_scatter does, publishes its 256 digit counts (flag AGGREGATE), and thread d sums the counts of the preceding tiles
// backwards until it meets a tile that has published its INCLUSIVE prefix, then publishes its own -- flag and 30-bit count
// share one word, so plain relaxed loads / stores are enough.  The walk reads kLookWin predecessors per round trip (the
// first wave starts in lock-step: one predecessor per L2 latency would cost ~sqrt(2 x tiles) latencies).
// Tiles are numbered in input order and the walk adds lower tiles only: the sort stays stable.
constexpr uint32_t kFlagAgg = 1u << 30, kFlagPre = 2u << 30, kCntMask = (1u << 30) - 1u;
constexpr int kLookWin = 8;
constexpr int kMaxPasses = 8;
constexpr unsigned kSpinLimit = 1u << 26;   // polls of one status word before the kernel gives up (error word, no hang)

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// head of the one-sweep workspace: [kMaxPasses][kRadix] digit totals, then kMaxPasses tickets, then one error word
constexpr int kHeadWords = kMaxPasses * kRadix + kMaxPasses + 8;

template <typename K>
__global__ void __launch_bounds__(kSortThreads)
sort_hist_all(const K* __restrict__ keys, int64_t n, int begin_bit, int passes, uint32_t* __restrict__ ghist,
              uint32_t* __restrict__ status0, int64_t status_words) {
    __shared__ unsigned sh[kMaxPasses][kRadix];
    for (int p = 0; p < passes; ++p) sh[p][threadIdx.x] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kSortThreads;
    const int64_t first = (int64_t)blockIdx.x * kSortThreads + threadIdx.x;
    for (int64_t i = first; i < n; i += stride) {
        const K k = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p][(unsigned)(k >> (begin_bit + 8 * p)) & 0xffu], 1u);
    }
    for (int64_t i = first; i < status_words; i += stride) status0[i] = 0u;   // the first pass's tile states
    __syncthreads();
    for (int p = 0; p < passes; ++p) {
        const unsigned c = sh[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * kRadix + threadIdx.x], c);
    }
}

template <typename K, int kItems>   // kItems elements per thread: 16 (tiles of 4096), 8 for small inputs (more, shorter CTAs)
__global__ void __launch_bounds__(kSortThreads, sizeof(K) == 4 ? 4 : 3)
sort_onesweep(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
              uint32_t* __restrict__ vals_out, int64_t n, int bit, const uint32_t* __restrict__ ghist,
              uint32_t* __restrict__ ticket, uint32_t* __restrict__ status, uint32_t* __restrict__ status_next,
              uint32_t* __restrict__ error_word) {
    // Ranking: warp w owns the contiguous 512 elements [512 w, 512 (w + 1)) of the tile (item i of lane l = element
    // 512 w + 32 i + l) and keeps PRIVATE digit counters wcount[w][.]: the lanes holding one digit find each other with
    // match_any, their leader advances the warp's counter and hands the old value out -- no block barrier inside the
    // item loop (the three-launch kernel above pays three per item).  One pass over the 8 x 256 counters afterwards
    // turns them into exclusive offsets across warps; (warp, item, lane) order = input order, so the sort is stable.
    __shared__ unsigned gbase[kRadix];                  // global slot of this tile's first element of each digit
    __shared__ unsigned lbase[kRadix];                  // position of that element in the reordered tile
    __shared__ unsigned wcount[kWarpsPerCta][kRadix];   // per-warp digit counts, then exclusive offsets across warps
    __shared__ unsigned wsum[kWarpsPerCta];
    __shared__ unsigned s_tile;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    K* skey = reinterpret_cast<K*>(tile_raw);
    uint32_t* sval = reinterpret_cast<uint32_t*>(tile_raw + sizeof(K) * kSortThreads * kItems);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) wcount[w][tid] = 0u;
    unsigned dbase;                                     // first global slot of digit `tid`: exclusive scan of the digit totals
    {
        const unsigned c = ghist[tid];
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();                                // (also publishes s_tile and the cleared counters)
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        dbase = off + incl - c;
    }
    const int64_t tile = (int64_t)s_tile;
    if (status_next != nullptr) status_next[tile * kRadix + tid] = 0u;   // the next pass's tile states (other buffer)
    constexpr int kTile = kSortThreads * kItems;
    const int64_t tile0 = tile * kTile;
    const int tile_n = (int)(n - tile0 < kTile ? n - tile0 : kTile);
    constexpr int kWarpItems = kItems * 32;
    const int wbase = warp * kWarpItems + lane;
    K key[kItems];
    uint32_t val[kItems];
    unsigned where[kItems];  // digit | slot within (digit, warp) << 8
    if (tile_n == kTile) {                             // (CTA-uniform) a full tile: no bound checks
        const K* kp = keys_in + tile0 + wbase;
        const uint32_t* vp = vals_in + tile0 + wbase;
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            key[i] = kp[i * 32];
            val[i] = vp[i * 32];
        }
    } else {
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int s = wbase + i * 32;
            key[i] = s < tile_n ? keys_in[tile0 + s] : (K)0;
            val[i] = s < tile_n ? vals_in[tile0 + s] : 0u;
        }
    }
    unsigned* mine = wcount[warp];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const bool valid = wbase + i * 32 < tile_n;
        const unsigned digit = valid ? (unsigned)(key[i] >> bit) & 0xffu : 0xffffffffu;  // invalid lanes match nobody real
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const unsigned rank = __popc(peers & lt_mask);
        unsigned old = 0u;
        if (valid && rank == 0) {
            old = mine[digit];
            mine[digit] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        where[i] = (digit & 0xffu) | ((old + rank) << 8);
        __syncwarp();
    }
    __syncthreads();
    // ---- counts of digit `tid` per warp -> exclusive offsets across warps, the tile's count
    unsigned count = 0u;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) {
        const unsigned c = wcount[w][tid];
        wcount[w][tid] = count;
        count += c;
    }
    // ---- publish the tile's digit counts at once (the tiles behind this one can go on with them) ...
    st_relaxed(status + tile * kRadix + tid, (tile > 0 ? kFlagAgg : kFlagPre) | count);
    {   // lbase = exclusive scan of the tile's digit counts (256 counters, one per thread)
        unsigned incl = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();                                // wsum of the digit-total scan has been read by everyone
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned off = 0;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        lbase[tid] = off + incl - count;
    }
    __syncthreads();
    // ... reorder the tile in shared memory (needs local offsets only; frees the registers of the elements) ...
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        if (wbase + i * 32 < tile_n) {
            const unsigned d = where[i] & 0xffu;
            const unsigned pos = lbase[d] + mine[d] + (where[i] >> 8);
            skey[pos] = key[i];
            sval[pos] = val[i];
        }
    }
    // ... and only then sum the preceding tiles (look-back) and publish the inclusive prefix.  (A window of 32 tiles per
    // round trip instead of 8 changed nothing at 10^6 pairs and cost 30 registers: the walk is not what bounds the kernel.)
    unsigned excl = 0u;
    if (tile > 0) {
        int64_t s = tile - 1;
        bool found = false;
        unsigned spins = 0u;
        while (!found) {
            uint32_t v[kLookWin];
#pragma unroll
            for (int q = 0; q < kLookWin; ++q)
                v[q] = s - q >= 0 ? ld_relaxed(status + (s - q) * kRadix + tid) : kFlagPre;   // before tile 0: prefix 0
            int advanced = 0;
#pragma unroll
            for (int q = 0; q < kLookWin; ++q) {
                if (found || advanced != q) continue;   // stop at the first tile that has not published yet
                const uint32_t f = v[q] >> 30;
                if (f == 0u) continue;
                excl += v[q] & kCntMask;
                found = f == 2u;
                ++advanced;
            }
            s -= advanced;
            if (advanced == 0 && ++spins > kSpinLimit) { *error_word = 1u; break; }
        }
        st_relaxed(status + tile * kRadix + tid, kFlagPre | (excl + count));
    }
    gbase[tid] = dbase + excl;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int s = i * kSortThreads + tid;
        if (s < tile_n) {
            const K k = skey[s];
            const unsigned digit = (unsigned)(k >> bit) & 0xffu;
            const unsigned dst = gbase[digit] + ((unsigned)s - lbase[digit]);
            keys_out[dst] = k;
            vals_out[dst] = sval[s];
        }
    }
}

}  // namespace

// small inputs (<= 32 chunks): ONE CTA walks the chunks with a running carry -- one launch instead of three
__global__ void __launch_bounds__(1024) scan_single(uint32_t* __restrict__ data, int64_t total) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int64_t c0 = 0; c0 < total; c0 += kScanChunk) {
        const int64_t i0 = c0 + (int64_t)threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < total) ? data[i0 + k] : 0u;
        unsigned block_total;
        unsigned excl = carry + block_excl_scan_1024(v[0] + v[1] + v[2] + v[3], warp_sum, block_total);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < total) data[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += block_total;
        __syncthreads();
    }
}

int exclusive_scan_u32(uint32_t* data, int64_t n, uint32_t* sums, cudaStream_t stream) {
    if (n <= 0) return MB_OK;
    const int64_t nchunks = ceil_div(n, kScanChunk);
    if (nchunks <= 2) {
        scan_single<<<1, 1024, 0, stream>>>(data, n);
        return check_cuda(cudaGetLastError(), "exclusive_scan_u32");
    }
    scan_chunks<<<(unsigned)nchunks, 1024, 0, stream>>>(data, n, sums);
    scan_top<<<1, 1024, 0, stream>>>(sums, nchunks);
    scan_add<<<(unsigned)nchunks, 1024, 0, stream>>>(data, n, sums);
    return check_cuda(cudaGetLastError(), "exclusive_scan_u32");
}

template <typename K>
int radix_sort_pairs(K* k0, uint32_t* v0, K* k1, uint32_t* v1, int64_t n, int begin_bit, int end_bit,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n < 0 || begin_bit < 0 || end_bit > (int)(8 * sizeof(K)) || begin_bit > end_bit) {
        set_error("radix_sort_pairs: bad argument");
        return MB_ERR_INVALID;
    }
    if (n >= (1ll << 32)) {
        set_error("radix_sort_pairs: n must be < 2^32");
        return MB_ERR_UNSUPPORTED;
    }
    if (n == 0 || begin_bit == end_bit) return 0;
    if (workspace == nullptr || workspace_bytes < sort_workspace_bytes(n)) {
        set_error("radix_sort_pairs: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    const int64_t tiles = sort_num_tiles(n);
    constexpr size_t tile_smem = (sizeof(K) + sizeof(uint32_t)) * kSortTile;  // the reordered tile (32 / 48 KiB)
    static bool attr_set = false;  // one flag per instantiation (K)
    if (!attr_set) {
        if (check_cuda(cudaFuncSetAttribute(sort_scatter<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem),
                       "radix_sort_pairs") != MB_OK)
            return MB_ERR_CUDA;
        if (check_cuda(cudaFuncSetAttribute(sort_onesweep<K, kSortItems>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem),
                       "radix_sort_pairs") != MB_OK)
            return MB_ERR_CUDA;
        attr_set = true;
    }
    K* ks[2] = {k0, k1};
    uint32_t* vs[2] = {v0, v1};
    int cur = 0;
    const int passes = (end_bit - begin_bit + 7) / 8;
    static int legacy = -1;   // MB_SORT_LEGACY=1: the three-launch passes (cross-checks)
    if (legacy < 0) {
        const char* e = getenv("MB_SORT_LEGACY");
        legacy = e ? atoi(e) : 0;
    }
    if (!legacy && n < (1ll << 30) && passes <= kMaxPasses) {
        // one sweep per digit: [head: digit totals of every pass, tickets, error word][tile states A][tile states B]
        // small inputs take tiles of 2048: one wave of 4096-element tiles leaves SMs idle and is as long as its slowest CTA
        const bool small = false;   // (tiles of 2048 for n <= 1.2 M measured SLOWER: twice the tiles to look back over)
        const int64_t stiles = small ? ceil_div(n, (int64_t)kSortTile / 2) : tiles;
        uint32_t* head = reinterpret_cast<uint32_t*>(workspace);
        uint32_t* st[2] = {head + kHeadWords, head + kHeadWords + (int64_t)kRadix * stiles};
        if (check_cuda(cudaMemsetAsync(head, 0, sizeof(uint32_t) * kHeadWords, stream), "radix_sort_pairs") != MB_OK) return MB_ERR_CUDA;
        const int64_t want = ceil_div(n, (int64_t)kSortThreads * 8);
        const int64_t cap = (int64_t)sm_count() * 8;
        sort_hist_all<K><<<(unsigned)(want < cap ? want : cap), kSortThreads, 0, stream>>>(ks[cur], n, begin_bit, passes, head, st[0],
                                                                                       (int64_t)kRadix * stiles);
        for (int p = 0; p < passes; ++p) {
            uint32_t* nxt = p + 1 < passes ? st[(p + 1) & 1] : nullptr;
            if (small)
                sort_onesweep<K, kSortItems / 2><<<(unsigned)stiles, kSortThreads, tile_smem / 2, stream>>>(
                    ks[cur], vs[cur], ks[cur ^ 1], vs[cur ^ 1], n, begin_bit + 8 * p, head + p * kRadix, head + kMaxPasses * kRadix + p,
                    st[p & 1], nxt, head + kMaxPasses * kRadix + kMaxPasses);
            else
                sort_onesweep<K, kSortItems><<<(unsigned)stiles, kSortThreads, tile_smem, stream>>>(
                    ks[cur], vs[cur], ks[cur ^ 1], vs[cur ^ 1], n, begin_bit + 8 * p, head + p * kRadix, head + kMaxPasses * kRadix + p,
                    st[p & 1], nxt, head + kMaxPasses * kRadix + kMaxPasses);
            cur ^= 1;
        }
        if (check_cuda(cudaGetLastError(), "radix_sort_pairs") != MB_OK) return MB_ERR_CUDA;
        return cur;
    }
    uint32_t* hist = reinterpret_cast<uint32_t*>(workspace);
    for (int bit = begin_bit; bit < end_bit; bit += 8) {
        sort_hist<K><<<(unsigned)tiles, kSortThreads, 0, stream>>>(ks[cur], n, bit, hist, tiles);
        if (exclusive_scan_u32(hist, (int64_t)kRadix * tiles, hist + (int64_t)kRadix * tiles, stream) != MB_OK) return MB_ERR_CUDA;
        sort_scatter<K><<<(unsigned)tiles, kSortThreads, tile_smem, stream>>>(ks[cur], vs[cur], ks[cur ^ 1], vs[cur ^ 1],
                                                                             n, bit, hist, tiles);
        cur ^= 1;
    }
    if (check_cuda(cudaGetLastError(), "radix_sort_pairs") != MB_OK) return MB_ERR_CUDA;
    return cur;
}

template int radix_sort_pairs<uint32_t>(uint32_t*, uint32_t*, uint32_t*, uint32_t*, int64_t, int, int, void*, size_t,
                                        cudaStream_t);
template int radix_sort_pairs<uint64_t>(uint64_t*, uint32_t*, uint64_t*, uint32_t*, int64_t, int, int, void*, size_t,
                                        cudaStream_t);

}  // namespace mb

// ---- C ABI ---------------------------------------------------------------------------------
MB_API int mb_sort_workspace_bytes(int64_t n, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0, "mb_sort_workspace_bytes: bad argument");
    *out = mb::sort_workspace_bytes(n);
    return MB_OK;
}

MB_API int mb_sort_pairs_u32(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, int64_t n,
                             int begin_bit, int end_bit, void* workspace, size_t workspace_bytes,
                             int* result_buffer, cudaStream_t stream) {
    const int rc = mb::radix_sort_pairs<uint32_t>(keys0, vals0, keys1, vals1, n, begin_bit, end_bit, workspace,
                                                  workspace_bytes, stream);
    if (rc < 0) return rc;
    if (result_buffer) *result_buffer = rc;
    return MB_OK;
}

MB_API int mb_sort_pairs_u64(uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1, int64_t n,
                             int begin_bit, int end_bit, void* workspace, size_t workspace_bytes,
                             int* result_buffer, cudaStream_t stream) {
    const int rc = mb::radix_sort_pairs<uint64_t>(keys0, vals0, keys1, vals1, n, begin_bit, end_bit, workspace,
                                                  workspace_bytes, stream);
    if (rc < 0) return rc;
    if (result_buffer) *result_buffer = rc;
    return MB_OK;
}

// ---- offsets of a CSR result from its per-row counts (radius queries: csrc/continuous.cu counts first, the host sizes
// the index / distance arrays, the second launch fills them): offsets[i] = sum of counts[0 .. i), offsets[n] = the total.
// scratch: (n + 1) + mb_scan_scratch_words(n + 1) uint32 words.  The running sum is 32-bit: the caller checks the total
// (a 64-bit reduction of the counts) against 2^32 first.
namespace {
__global__ void counts_to_words(const int32_t* __restrict__ counts, uint32_t* __restrict__ words, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) words[i] = i < n ? (uint32_t)counts[i] : 0u;
}
__global__ void words_to_offsets(const uint32_t* __restrict__ words, int64_t* __restrict__ offsets, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) offsets[i] = (int64_t)words[i];
}
}  // namespace

MB_API int mb_offsets_from_counts(const int32_t* counts, int64_t n, int64_t* offsets, uint32_t* scratch, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && offsets != nullptr && scratch != nullptr && (n == 0 || counts != nullptr), "mb_offsets_from_counts: bad argument");
    const unsigned blocks = (unsigned)mb::ceil_div(n + 1, 256);
    counts_to_words<<<blocks, 256, 0, stream>>>(counts, scratch, n);
    const int rc = mb::exclusive_scan_u32(scratch, n + 1, scratch + (n + 1), stream);
    if (rc != MB_OK) return rc;
    words_to_offsets<<<blocks, 256, 0, stream>>>(scratch, offsets, n);
    MB_LAUNCH_CHECK("mb_offsets_from_counts");
    return MB_OK;
}
