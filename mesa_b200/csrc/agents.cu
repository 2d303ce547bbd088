// Dynamic agent populations: stable compaction of structure-of-arrays columns.
//
// Reference: removing an agent deletes its key from the ordered dict behind model.agents
// (mesa/agentset.py:700-711 `_HardKeyAgentSet.add / discard / remove`, mesa/agent.py:73-108 `Agent.remove` ->
// `Model.deregister_agent`, mesa/model.py:281-294) -- the remaining agents keep their relative
// (registration) order, which is the base order of every later `do / shuffle_do`.  On the device a
// population is one tensor per attribute, row = position in that order, so removing a batch of agents
// is ONE stable compaction shared by all columns: keep flags -> exclusive prefix sum (the scan of
// csrc/sort.cu) -> one scatter per column.  Creation appends rows (host side: a grown tensor and a
// device-to-device copy), so `unique_id` order == row order is preserved.
//
// HBM-bound byte work: per column row_bytes read + row_bytes written per surviving agent, 1 B flag +
// 4 B offset read per agent.  Rows are moved in the widest unit (16 / 8 / 4 / 2 / 1 bytes) that divides
// row_bytes and both base addresses; consecutive threads move consecutive units, so loads are fully
// coalesced and stores are coalesced within every run of survivors.
#include "common.cuh"
#include "sort.cuh"

namespace {

__global__ void flags_to_words(const uint8_t* __restrict__ keep, uint32_t* __restrict__ offs, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) offs[i] = (i < n && keep[i] != 0) ? 1u : 0u;  // offs[n] = 0: after the scan it holds the total
}

template <typename U>
__global__ void compact_rows(const U* __restrict__ in, U* __restrict__ out, int64_t units, unsigned upr,
                             const uint8_t* __restrict__ keep, const uint32_t* __restrict__ offs) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < units; t += stride) {
        const int64_t row = t / upr;
        if (!keep[row]) continue;
        const unsigned k = (unsigned)(t - row * upr);
        out[(int64_t)offs[row] * upr + k] = in[t];
    }
}

template <typename U>
int launch_rows(const void* in, void* out, int64_t n, int64_t row_bytes, const uint8_t* keep, const uint32_t* offs,
                cudaStream_t stream) {
    const int64_t upr = row_bytes / (int64_t)sizeof(U);
    const int64_t units = n * upr;
    int64_t blocks = mb::ceil_div(units, 256);
    const int64_t cap = (int64_t)mb::sm_count() * 16;  // grid-stride beyond 16 CTAs per SM
    if (blocks > cap) blocks = cap;
    compact_rows<U><<<(unsigned)blocks, 256, 0, stream>>>((const U*)in, (U*)out, units, (unsigned)upr, keep, offs);
    return mb::check_cuda(cudaGetLastError(), "mb_compact_rows");
}

}  // namespace

// offsets[i] (i < n) = number of kept rows before row i; offsets[n] = number of kept rows.
// keep: n bytes (0 = drop, anything else = keep); offsets: n + 1 words; scratch: mb_scan_scratch_words(n + 1) words.
MB_API int mb_compact_plan(const uint8_t* keep, int64_t n, uint32_t* offsets, uint32_t* scratch, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && offsets != nullptr, "mb_compact_plan: bad argument");
    MB_REQUIRE(n < (1ll << 32) - 1, "mb_compact_plan: n must be < 2^32 - 1");
    MB_REQUIRE((n == 0 || keep != nullptr) && scratch != nullptr, "mb_compact_plan: null buffer");
    flags_to_words<<<(unsigned)mb::ceil_div(n + 1, 256), 256, 0, stream>>>(keep, offsets, n);
    MB_LAUNCH_CHECK("mb_compact_plan");
    return mb::exclusive_scan_u32(offsets, n + 1, scratch, stream);
}

// out[offsets[i]] = in[i] for every kept row i; rows are row_bytes wide, in and out must not overlap.
MB_API int mb_compact_rows(const void* in, void* out, int64_t n, int64_t row_bytes, const uint8_t* keep,
                           const uint32_t* offsets, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && row_bytes > 0 && row_bytes <= (1 << 20), "mb_compact_rows: bad argument");
    if (n == 0) return MB_OK;
    MB_REQUIRE(in && out && keep && offsets, "mb_compact_rows: null buffer");
    const uintptr_t bits = (uintptr_t)in | (uintptr_t)out | (uintptr_t)row_bytes;
    if ((bits & 15) == 0) return launch_rows<uint4>(in, out, n, row_bytes, keep, offsets, stream);
    if ((bits & 7) == 0) return launch_rows<uint2>(in, out, n, row_bytes, keep, offsets, stream);
    if ((bits & 3) == 0) return launch_rows<uint32_t>(in, out, n, row_bytes, keep, offsets, stream);
    if ((bits & 1) == 0) return launch_rows<uint16_t>(in, out, n, row_bytes, keep, offsets, stream);
    return launch_rows<uint8_t>(in, out, n, row_bytes, keep, offsets, stream);
}
