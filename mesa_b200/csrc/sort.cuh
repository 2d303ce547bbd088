// Stable LSD radix sort of (key, uint32 value) pairs, 8 bits per pass: warp-level multi-split
// (match_any ranks) + per-CTA digit histograms + one prefix scan per pass (K7 of SURVEY.md §2.4).
//
// Used for the cell list of the continuous space (key = cell id) and for the per-cell arrival
// order of the Boltzmann model (key = cell id << 32 | Philox priority).  The reference has no
// counterpart: it scans all agents per query (continuous_space.py:202-235) and keeps Python
// lists per cell (cell.py:143-170).
//
// Stability matters: pairs enter in agent-index order, so equal keys keep ascending agent
// index -- exactly the tie-break of priority64 (philox.cuh).
#pragma once

#include "common.cuh"

namespace mb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                         // items per thread
constexpr int kSortTile = kSortThreads * kSortItems;   // items per CTA
constexpr int kRadix = 256;

inline int64_t sort_num_tiles(int64_t n) { return ceil_div(n > 0 ? n : 1, kSortTile); }

// workspace: the larger of
//   one-sweep passes (sort.cu): digit totals of up to 8 passes + tickets (2120 words), two [tiles][kRadix] tile-state arrays
//   three-launch passes (n >= 2^30, MB_SORT_LEGACY): histogram [kRadix][tiles] uint32 (digit-major so one scan orders digits
//   first) + one sum per 4096-counter scan chunk
inline size_t sort_workspace_bytes(int64_t n) {
    const size_t total = (size_t)kRadix * (size_t)sort_num_tiles(n);
    const size_t legacy = total + (total + 4095) / 4096 + 64;
    const size_t sweep = 2 * (2 * total + (size_t)kRadix) + 2120;   // (small inputs: tiles of half the size)
    return sizeof(uint32_t) * (legacy > sweep ? legacy : sweep) + 256;
}

// Sort n pairs on key bits [begin_bit, end_bit).  Ping-pongs between (k0,v0) and (k1,v1);
// returns 0 if the result is in (k0,v0), 1 if in (k1,v1), negative on error.
template <typename K>
int radix_sort_pairs(K* k0, uint32_t* v0, K* k1, uint32_t* v1, int64_t n, int begin_bit, int end_bit,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

// In-place exclusive prefix sum of n uint32 counters; `sums` needs ceil(n / 4096) + 1 words of scratch.
inline size_t scan_scratch_words(int64_t n) { return (size_t)((n + 4095) / 4096) + 64; }
int exclusive_scan_u32(uint32_t* data, int64_t n, uint32_t* sums, cudaStream_t stream);

}  // namespace mb
