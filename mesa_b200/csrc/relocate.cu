// Batched relocation of unhappy Schelling agents with deterministic, order-exact conflict
// resolution (K4+K5 of SURVEY.md §2.4).
//
// Replaces AgentSet.shuffle_do("step") (mesa/agentset.py:772-787) over
//   mesa/examples/basic/schelling/agents.py:44-47   if not self.happy: self.cell = grid.select_random_empty_cell()
//   mesa/discrete_space/grid.py:288-316             <=50 x random.choice(_celllist), first cell that is empty
//                                                   NOW wins; else argwhere(empty) + one more choice
//   mesa/discrete_space/cell_agent.py:39-54         add to the new cell, then remove from the old one
//
// The reference is sequential: agent #rho in the shuffled order sees the moves of agents
// < rho.  Here the activation order is the ascending Philox priority64 (philox.cuh) and the
// sequential outcome is computed as the unique fixed point of a parallel iteration:
//
//   * every mover (occupied && !happy at step start) has a state-independent probe sequence
//     p[0..49] = _randbelow(ncells) draws of its own DRAW stream;
//   * every cell has a 16-byte slot {vac, arr}: the cell is empty at time t (t = priority of
//     the acting mover) iff  vac <= t <= arr, where
//         vac = 0 (empty since the start of the step) | priority+1 of its occupant if the
//               occupant moves this step (it is gone once its turn is over) | NEVER
//         arr = min priority of the movers that currently choose the cell, NEVER if none
//     (a cell receives at most one arrival per step: an arrived agent has already acted);
//     one probe = ONE 16-byte load;
//   * a pass lets every mover take its first probe that is empty at its own time and keep `arr`
//     up to date in place (atomicMin to claim, compare-and-swap to release, re-asserted every
//     pass).  The mover of lowest rank is right after one pass and its claim is never disturbed;
//     by induction rank k is right one pass after ranks < k are, so the passes converge to
//     exactly the sequential result (in practice a handful of passes: conflicts are sparse);
//   * the 50-probe miss (probability density^50) falls back to the k-th empty cell in row-major
//     order at the mover's time, k = _randbelow(#empty), #empty = ncells - nagents at every
//     instant (grid.py:308-316).  All fallback movers of a pass are served by ONE pass over the
//     slot table: their times are sorted, a cell's emptiness interval is a contiguous range of
//     that order (two binary searches -> +1/-1 in a per-block difference array), per-block
//     counts locate the block of the k-th empty cell and one warp scans that block.
//
// All state is caller-owned: the int8 type layer, the uint8 happy layer, the uint32
// agent-id layer (who sits in each cell), the slot table (16 B per cell, persistent: built
// from the type layer once), optionally the per-agent cell array, and a private workspace.
//
// ROW-SHARDED GRIDS (one process per GPU): a probe is a uniform draw over the
// WHOLE grid, so most probes land on another rank's rows.  The per-cell arrays of every rank
// are peer-mapped (torch symmetric memory over NVLink / NVSwitch) and described by mb_shard_t;
// the same kernels then load remote slots, claim them with NVLink atomics and write arrivals
// into the owner's layers directly -- no all-to-all of claims.  The host side
// (mesa_b200/sharded.py) only all-reduces the per-pass flags and runs the small collectives of
// the fallback (gather the fallback movers, all-gather per-rank empty counts).  Priorities and
// draws are keyed on global agent ids, so the result is identical for every GPU count.
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include "philox.cuh"

#include <string.h>

namespace {

constexpr unsigned long long kNever = ~0ull;
constexpr unsigned long long kNoCell = ~0ull;
constexpr int kMaxProbes = 50;   // grid.py:303 `for _ in range(50)`
constexpr int kMaxPasses = 100000;
constexpr int kFbBatch = 2048;   // fallback movers served per table pass
constexpr int kRelCap = 4096;    // released cells recorded between two scans (more: answered with a full pass)
constexpr int kFbBlocksMax = 65536;   // blocks of a rank's slot table with per-fallback empty counts (fine blocks: short scans)
constexpr int kFbSuper = 64;          // blocks per super-block: the count walk is two short scans (supers, then 64 blocks)
constexpr int kFbSupersMax = kFbBlocksMax / kFbSuper;
constexpr int kMaxPeers = 16;
constexpr int kOvCapWs = 4096;   // = kOvCap (event loop overrides)

struct Slot {
    unsigned long long vac, arr;
};

// mirror of mb_shard_t (include/mesa_b200.h)
struct ShardDesc {
    int32_t world, rank;
    int64_t rows_total, cols;
    void* slots[kMaxPeers];
    void* type[kMaxPeers];
    void* happy[kMaxPeers];
    void* agent_id[kMaxPeers];
    void* static_bits;      // optional: [world][bits_stride] uint32, bit l of section r = local cell l of rank r is occupied by
    int64_t bits_stride;    //           an agent that does NOT move this step (this rank's copy of every rank's section)
};

// Row-block partition of the global grid (same rule as mesa_b200/sharding.py block_bounds): the first
// `extra` ranks own base+1 rows.  Per-cell arrays of rank r are reached through peer-mapped base
// pointers; with world == 1 everything is local.
struct ShardView {
    int world, rank;
    long long cols, base, extra, ncells, local_lo, local_cells;  // local_lo / local_cells in cells
    Slot* slots[kMaxPeers];
    int8_t* type[kMaxPeers];
    uint8_t* happy[kMaxPeers];
    uint32_t* agent_id[kMaxPeers];
    uint32_t* static_bits;    // see ShardDesc (nullptr: no filter)
    long long bits_stride;

    __host__ __device__ long long row_lo(int r) const { return (long long)r * base + (r < extra ? r : extra); }
    __device__ __forceinline__ void locate(unsigned long long c, int& owner, unsigned long long& local) const {
        if (world == 1) { owner = 0; local = c; return; }
        const long long row = (long long)(c / (unsigned long long)cols);
        const long long cut = (base + 1) * extra;
        owner = row < cut ? (int)(row / (base + 1)) : (int)(extra + (row - cut) / base);
        local = c - (unsigned long long)(row_lo(owner) * cols);
    }
    __device__ __forceinline__ Slot* slot(unsigned long long c) const {
        int o; unsigned long long l;
        locate(c, o, l);
        return slots[o] + l;
    }
    // slot pointer + "the cell holds an agent that stays put this step" (one L2-resident bit instead of a 16-byte
    // random DRAM access: more than half of all probes on a 0.8-dense grid end here)
    __device__ __forceinline__ Slot* slot_and_static(unsigned long long c, bool& is_static) const {
        int o; unsigned long long l;
        locate(c, o, l);
        is_static = static_bits != nullptr && ((__ldg(static_bits + (size_t)o * bits_stride + (l >> 5)) >> (l & 31u)) & 1u);
        return slots[o] + l;
    }
};

int make_view(const ShardDesc* d, ShardView* v) {
    MB_REQUIRE(d != nullptr, "null shard descriptor");
    MB_REQUIRE(d->world >= 1 && d->world <= kMaxPeers && d->rank >= 0 && d->rank < d->world, "bad world / rank");
    MB_REQUIRE(d->rows_total >= d->world && d->cols > 0, "every rank needs at least one row");
    v->world = d->world; v->rank = d->rank; v->cols = d->cols;
    v->base = d->rows_total / d->world; v->extra = d->rows_total % d->world;
    v->ncells = d->rows_total * d->cols;
    v->local_lo = v->row_lo(d->rank) * d->cols;
    v->local_cells = (v->row_lo(d->rank + 1) - v->row_lo(d->rank)) * d->cols;
    MB_REQUIRE(v->local_cells < 0xffffffffll, "a rank's block must hold < 2^32 - 1 cells");
    MB_REQUIRE(v->ncells <= (1ll << 32), "the grid must hold at most 2^32 cells");
    for (int r = 0; r < d->world; ++r) {
        MB_REQUIRE(d->slots[r] && d->type[r] && d->happy[r] && d->agent_id[r], "null peer pointer in shard descriptor");
        v->slots[r] = (Slot*)d->slots[r]; v->type[r] = (int8_t*)d->type[r];
        v->happy[r] = (uint8_t*)d->happy[r]; v->agent_id[r] = (uint32_t*)d->agent_id[r];
    }
    v->static_bits = (uint32_t*)d->static_bits;
    v->bits_stride = d->bits_stride;
    MB_REQUIRE(v->static_bits == nullptr || v->bits_stride * 32 >= (v->base + (v->extra ? 1 : 0)) * v->cols,
               "static bitmap sections are shorter than a rank's block");
    return MB_OK;
}

struct RelocWs {
    uint32_t* mv_cell;              // [cap] LOCAL origin cell of mover m
    uint32_t* mv_agent;             // [cap] global agent index
    unsigned long long* mv_prio;    // [cap] priority64
    int8_t* mv_type;                // [cap]
    unsigned long long* choice;     // [cap] GLOBAL cell currently chosen (claimed) by mover m, kNoCell if none
    uint32_t* progress;             // [cap] probe index of `choice` (kMaxProbes: fallback state) | draws consumed << 8
    unsigned long long* softmask;   // [cap] bit j: probe j was skipped only because an EARLIER ARRIVAL holds the cell
                                    //       (vac <= t but arr < t) -- the probes a release can re-open
    uint32_t* softcell;             // [cap] the cell of the LOWEST such probe (most movers have at most one: the release
                                    //       scan then needs no Philox replay for them)
    uint32_t* fb_list;              // [cap] movers that missed all 50 probes in this pass
    uint32_t* bucket_list;          // [cap] mover indices grouped by rank-ordered bucket (mb_reloc_bucket_lists)
    unsigned long long* bucket_cur; // [256] per-bucket counts, then write cursors
    unsigned long long* fb_t;       // [kFbBatch] sorted times of the current fallback batch
    uint32_t* fb_m;                 // [kFbBatch] local mover of each sorted entry
    unsigned long long* fb_k;       // [kFbBatch] rank of the empty cell to take (within this rank's table)
    unsigned long long* fb_res;     // [kFbBatch] chosen global cell + 1 (0 = not found here)
    uint32_t* fb_tot;               // [kFbBatch] empties of the local table at each fallback time
    uint32_t* fb_block;             // [kFbBatch]
    uint32_t* fb_rem;               // [kFbBatch]
    uint32_t* cnt;                  // [kFbBlocksMax][kFbBatch] empties per local table block, then [kFbSupersMax][kFbBatch]
                                    // the same per super-block (fb_sup)
    unsigned long long* rel_w_cell;  // [kRelCap] cells released since the last take (write set)
    unsigned long long* rel_w_time;  // [kRelCap] time of the mover that left each of them
    unsigned long long* rel_r_cell;  // [kRelCap] read set of the single-GPU driver's scans
    unsigned long long* rel_r_time;
    unsigned long long* counters;   // [0] movers [1] occupied [2] fallback movers listed [3] table changed [4] (unused)
                                    // [5] entries of the release write set [6] a full pass is required
    unsigned long long* ov_time;    // [kOvCap] event loop overrides: time of the mover ...
    unsigned long long* ov_cell;    // [kOvCap] ... the global cell it holds now
    unsigned long long* ev_out;     // [4] overrides, bail reason, fallback events, displaced events
};

size_t carve(RelocWs* ws, char* base, int64_t cap) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    RelocWs t;
    t.mv_cell = (uint32_t*)take(4 * cap);
    t.mv_agent = (uint32_t*)take(4 * cap);
    t.mv_prio = (unsigned long long*)take(8 * cap);
    t.mv_type = (int8_t*)take(cap);
    t.choice = (unsigned long long*)take(8 * cap);
    t.progress = (uint32_t*)take(4 * cap);
    t.softmask = (unsigned long long*)take(8 * cap);
    t.softcell = (uint32_t*)take(4 * cap);
    t.fb_list = (uint32_t*)take(4 * cap);
    t.bucket_list = (uint32_t*)take(4 * cap);
    t.bucket_cur = (unsigned long long*)take(8 * 256);
    t.fb_t = (unsigned long long*)take(8 * kFbBatch);
    t.fb_m = (uint32_t*)take(4 * kFbBatch);
    t.fb_k = (unsigned long long*)take(8 * kFbBatch);
    t.fb_res = (unsigned long long*)take(8 * kFbBatch);
    t.fb_tot = (uint32_t*)take(4 * kFbBatch);
    t.fb_block = (uint32_t*)take(4 * kFbBatch);
    t.fb_rem = (uint32_t*)take(4 * kFbBatch);
    t.cnt = (uint32_t*)take(4ull * (kFbBlocksMax + kFbSupersMax) * kFbBatch);
    t.rel_w_cell = (unsigned long long*)take(8 * kRelCap);
    t.rel_w_time = (unsigned long long*)take(8 * kRelCap);
    t.rel_r_cell = (unsigned long long*)take(8 * kRelCap);
    t.rel_r_time = (unsigned long long*)take(8 * kRelCap);
    t.counters = (unsigned long long*)take(8 * 8);
    t.ov_time = (unsigned long long*)take(8 * kOvCapWs);
    t.ov_cell = (unsigned long long*)take(8 * kOvCapWs);
    t.ev_out = (unsigned long long*)take(8 * 8);
    if (ws) *ws = t;
    return off;
}

int get_ws(RelocWs* ws, void* workspace, size_t workspace_bytes, int64_t cap) {
    MB_REQUIRE(workspace != nullptr && cap >= 0, "null relocation workspace");
    if (carve(ws, (char*)workspace, cap) > workspace_bytes) {
        mb::set_error("relocation workspace too small");
        return MB_ERR_WORKSPACE;
    }
    return MB_OK;
}

__device__ __forceinline__ Slot load_slot(const Slot* p) {
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2*>(p));  // L2: sees other SMs' atomics
    return Slot{v.x, v.y};
}

__device__ __forceinline__ bool empty_at(const Slot& s, unsigned long long t) { return s.vac <= t && t <= s.arr; }

// the super-block counts follow the block counts in a count table
__host__ __device__ __forceinline__ uint32_t* fb_sup(uint32_t* cnt) { return cnt + (size_t)kFbBlocksMax * kFbBatch; }
__host__ __device__ __forceinline__ const uint32_t* fb_sup(const uint32_t* cnt) { return cnt + (size_t)kFbBlocksMax * kFbBatch; }

__global__ void init_slots(ShardView v) {
    const int8_t* __restrict__ type = v.type[v.rank];
    ulonglong2* __restrict__ slots = reinterpret_cast<ulonglong2*>(v.slots[v.rank]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < v.local_cells; c += stride)
        slots[c] = make_ulonglong2(type[c] >= 0 ? kNever : 0ull, kNever);
}

// movers = occupied && !happy; appended in arbitrary order (the result does not depend on it).
// Every CTA owns one contiguous chunk of cells: it counts its movers first, reserves their slots with ONE atomic and
// then writes them (the first version took a slot range per warp ballot: 2 M atomics on one address at 8192^2 made the
// kernel 1.6 ms long; the chunk is re-read from L2 for the second sweep).
__global__ void __launch_bounds__(256)
collect_movers(ShardView v, RelocWs ws, int64_t cap, uint64_t seed, uint32_t epoch, int64_t chunk) {
    const int8_t* __restrict__ type = v.type[v.rank];
    const uint8_t* __restrict__ happy = v.happy[v.rank];
    const uint32_t* __restrict__ agent_id = v.agent_id[v.rank];
    Slot* __restrict__ slots = v.slots[v.rank];
    __shared__ unsigned warp_cnt[8], warp_occ[8];
    __shared__ unsigned long long cta_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t lo = (int64_t)blockIdx.x * chunk;
    const int64_t hi = min(lo + chunk, (int64_t)v.local_cells);
    unsigned movers = 0, occupied = 0;
    for (int64_t c = lo + threadIdx.x; c < hi; c += blockDim.x) {
        const int t = type[c];
        occupied += t >= 0;
        movers += t >= 0 && happy[c] == 0;
    }
    movers = warp_reduce_add(movers);
    occupied = warp_reduce_add(occupied);
    if (lane == 0) { warp_cnt[warp] = movers; warp_occ[warp] = occupied; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m = 0, o = 0;
        for (int w = 0; w < 8; ++w) { m += warp_cnt[w]; o += warp_occ[w]; }
        cta_base = m ? atomicAdd(&ws.counters[0], (unsigned long long)m) : 0ull;
        if (o) atomicAdd(&ws.counters[1], (unsigned long long)o);
    }
    __syncthreads();
    unsigned long long next = cta_base;   // uniform across the CTA
    uint32_t* __restrict__ bits = v.static_bits ? v.static_bits + (size_t)v.rank * v.bits_stride : nullptr;
    for (int64_t c0 = lo; c0 < hi; c0 += blockDim.x) {
        const int64_t c = c0 + threadIdx.x;
        const int t = c < hi ? type[c] : -1;
        const bool mover = t >= 0 && happy[c] == 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, mover);
        const unsigned stat = __ballot_sync(0xffffffffu, t >= 0 && !mover);
        if (bits != nullptr && lane == 0 && c < hi) bits[c >> 5] = stat;   // chunks are multiples of 256 cells: whole words
        __syncthreads();                   // warp_cnt of the previous round has been read
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned n = warp_cnt[w];
            before += w < warp ? n : 0u;
            total += n;
        }
        // the slot of EVERY cell is rebuilt here (sequential 16-byte stores): vac = 0 for an empty cell, priority + 1 of a
        // mover (it is gone once its turn is over), never for an agent that stays; nobody has arrived yet
        unsigned long long vac = t >= 0 ? kNever : 0ull;
        if (mover) {
            const unsigned long long slot = next + before + __popc(ballot & ((1u << lane) - 1u));
            const uint32_t agent = agent_id[c];
            const unsigned long long pr = mb::priority64(seed, epoch, agent);
            vac = pr + 1ull;
            if ((int64_t)slot < cap) {
                ws.mv_cell[slot] = (uint32_t)c;
                ws.mv_agent[slot] = agent;
                ws.mv_prio[slot] = pr;
                ws.mv_type[slot] = (int8_t)t;
                ws.choice[slot] = kNoCell;
                ws.progress[slot] = 0u;
                ws.softmask[slot] = 0ull;
            }
        }
        if (c < hi) reinterpret_cast<ulonglong2*>(slots)[c] = make_ulonglong2(vac, kNever);
        next += total;
    }
}

// "the table changed in this pass": a plain store of a constant (millions of movers change in the
// first pass; counting them with one atomic address would serialise the whole pass)
__device__ __forceinline__ void mark_changed(RelocWs& ws) { *(volatile unsigned long long*)&ws.counters[3] = 1ull; }

// give up `held` if I am its registered arrival.  A successful release can re-open the cell for later movers that skipped
// it on the way to their pick: it is recorded for the next release scan (relocate_release_scan)
__device__ __forceinline__ void release_held(const ShardView& v, RelocWs& ws, unsigned long long held, unsigned long long t) {
    if (held != kNoCell && atomicCAS(&v.slot(held)->arr, t, kNever) == t) {
        const unsigned long long slot = atomicAdd(&ws.counters[5], 1ull);
        if (slot < (unsigned long long)kRelCap) {
            ws.rel_w_cell[slot] = held;
            ws.rel_w_time[slot] = t;
        } else {
            *(volatile unsigned long long*)&ws.counters[6] = 1ull;  // too many to scan for: full pass
        }
    }
}

// claim `pick` for mover m (and release what it held before)
__device__ __forceinline__ void commit_choice(const ShardView& v, RelocWs& ws, int64_t m, unsigned long long t,
                                              unsigned long long pick) {
    const unsigned long long held = ws.choice[m];
    if (held != pick) {
        // release only if I am the registered arrival.  A successful release can re-open the cell for later movers that
        // skipped it on the way to their pick: it is recorded for the next release scan (relocate_release_scan)
        if (held != kNoCell && atomicCAS(&v.slot(held)->arr, t, kNever) == t) {
            const unsigned long long slot = atomicAdd(&ws.counters[5], 1ull);
            if (slot < (unsigned long long)kRelCap) {
                ws.rel_w_cell[slot] = held;
                ws.rel_w_time[slot] = t;
            } else {
                *(volatile unsigned long long*)&ws.counters[6] = 1ull;  // too many to scan for: full pass
            }
        }
        ws.choice[m] = pick;
        atomicMin(&v.slot(pick)->arr, t);
        mark_changed(ws);
    } else {
        // re-assert: my registration may have been shadowed by an earlier arrival that has left since
        if (atomicMin(&v.slot(pick)->arr, t) > t) mark_changed(ws);
    }
}

// One pass over the movers whose time lies in [t_lo, t_hi].  Movers are processed in rank-ordered
// buckets (earliest first, each to convergence): a mover only depends on earlier movers, so a
// converged bucket is final, and conflicts inside a bucket are 1/K as dense as in the whole set --
// far fewer passes touch far fewer movers (978 M movers on 65536^2: 88 passes over everyone without).
//
// Three kinds of pass.  FULL (mode 0): every mover re-evaluates its probe sequence from the first probe.
// RESUME (mode 1): as long as every release is answered by a release scan, the table is MONOTONE for a mover -- claims
// only lower `arr` (atomicMin), so a probe that was not empty at the mover's time stays so and its pick can only move
// FORWARD in its probe sequence.  A mover therefore only checks that it still holds its pick (one 16-byte load:
// arr == its time) and, if an earlier mover took it, continues probing after it (the draw counter is kept in
// `progress`).  RELEASE SCAN (relocate_release_scan): finds the movers for which a released cell re-opens an EARLIER
// probe.  8192^2, 15.3 M movers: a full pass 3.1-4.8 ms, a resume pass 0.35 ms once the movers have settled.
template <int kBatch>   // probes whose slot loads are issued together
__global__ void __launch_bounds__(128)
relocate_pass(ShardView v, RelocWs ws, int64_t nmovers, uint64_t seed, uint32_t epoch, unsigned long long t_lo,
              unsigned long long t_hi, int mode) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const unsigned long long t = ws.mv_prio[m];
    if (t < t_lo || t > t_hi) return;
    mb::DrawStream ds(seed, epoch, ws.mv_agent[m]);
    if (mode == 1) {
        const uint32_t pr = ws.progress[m];
        if ((pr & 0xffu) >= (uint32_t)kMaxProbes) return;  // fallback state: served by the fallback phase
        const unsigned long long held = ws.choice[m];
        const unsigned long long arr = load_slot(v.slot(held)).arr;
        if (arr == t) return;                              // still the registered arrival of my pick
        if (arr > t) {                                     // my registration was shadowed and the shadow has left
            atomicMin(&v.slot(held)->arr, t);
            mark_changed(ws);
            return;
        }
        ds.seek(pr >> 8);                                  // an earlier mover took it: go on with the next probe
        unsigned long long soft = ws.softmask[m];
        if (soft == 0ull) ws.softcell[m] = (uint32_t)held;   // the lost pick becomes the lowest soft-skipped probe
        soft |= 1ull << (pr & 0xffu);                        // (it was skipped because of an earlier arrival, too)
        // probes are issued kBatch at a time: the draws of a batch are state-independent, so its slot loads go out
        // together (memory-level parallelism; a lone dependent load per probe left the pass latency-bound)
        for (int j0 = (int)(pr & 0xffu) + 1; j0 < kMaxProbes; j0 += kBatch) {
            unsigned long long cells[kBatch];
            uint32_t after[kBatch];
            Slot sl[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                cells[b] = ds.randbelow((uint64_t)v.ncells);
                after[b] = ds.c;
            }
            Slot* sp[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                bool is_static;
                sp[b] = v.slot_and_static(cells[b], is_static);
                sl[b] = Slot{kNever, kNever};                    // an agent that stays: never empty, no slot load
                if (!is_static) sl[b] = load_slot(sp[b]);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int j = j0 + b;
                if (j >= kMaxProbes) continue;
                if (!empty_at(sl[b], t) || atomicMin(&sp[b]->arr, t) < t) {   // (the atomicMin: lost a race, not empty after all)
                    if (sl[b].vac <= t) soft |= 1ull << j;
                    continue;
                }
                ws.choice[m] = cells[b];
                ws.progress[m] = (uint32_t)j | (after[b] << 8);
                ws.softmask[m] = soft;
                mark_changed(ws);
                return;
            }
        }
        ws.choice[m] = kNoCell;
        ws.softmask[m] = soft;
        ws.progress[m] = (uint32_t)kMaxProbes;
        mark_changed(ws);
        ws.fb_list[atomicAdd(&ws.counters[2], 1ull)] = (uint32_t)m;
        return;
    }
    unsigned long long soft = 0ull;
    const unsigned long long held = ws.choice[m];
    for (int j0 = 0; j0 < kMaxProbes; j0 += kBatch) {
        unsigned long long cells[kBatch];
        uint32_t after[kBatch];
        Slot sl[kBatch];
        Slot* sp[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            cells[b] = ds.randbelow((uint64_t)v.ncells);
            after[b] = ds.c;
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            bool is_static;
            sp[b] = v.slot_and_static(cells[b], is_static);
            sl[b] = Slot{kNever, kNever};                        // an agent that stays: never empty, no slot load
            if (!is_static) sl[b] = load_slot(sp[b]);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const int j = j0 + b;
            if (j >= kMaxProbes) continue;
            if (empty_at(sl[b], t)) {
                if (cells[b] == held) {
                    // re-assert: my registration may have been shadowed by an earlier arrival that has left since
                    if (atomicMin(&sp[b]->arr, t) > t) mark_changed(ws);
                } else {
                    // claim; an earlier mover that registered between my load and my claim wins the cell: go on probing
                    // at once instead of finding out in the next pass
                    if (atomicMin(&sp[b]->arr, t) < t) {
                        if (soft == 0ull) ws.softcell[m] = (uint32_t)cells[b];
                        soft |= 1ull << j;
                        continue;
                    }
                    release_held(v, ws, held, t);
                    ws.choice[m] = cells[b];
                    mark_changed(ws);
                }
                ws.progress[m] = (uint32_t)j | (after[b] << 8);
                ws.softmask[m] = soft;
                return;
            }
            if (sl[b].vac <= t) {
                if (soft == 0ull) ws.softcell[m] = (uint32_t)cells[b];
                soft |= 1ull << j;
            }
        }
    }
    // (a pick held from an earlier pass is compared and, if need be, released by the fallback phase: commit_choice)
    ws.softmask[m] = soft;
    ws.progress[m] = (uint32_t)kMaxProbes;
    ws.fb_list[atomicAdd(&ws.counters[2], 1ull)] = (uint32_t)m;  // served by the fallback phase of this pass
}

// ---- the same pass with LANE REFILL -------------------------------------------------------------------------------------
// ncu of the one-thread-per-mover kernel above (profiles/r2_reloc_pass_ncu.txt): 5.1 of 32 lanes active in a FULL pass,
// 2.8 in a RESUME pass -- a mover needs a geometric number of probes (or, in a RESUME pass, none at all), so a warp
// spends most of its instructions on its slowest lane; 2.3 G warp instructions for 15 M movers, issue-bound, DRAM at
// 1.2 TB/s.  Here a warp owns a contiguous range of movers and keeps ALL lanes busy: it loads the movers 32 at a time
// (one coalesced tile in registers), every lane that finishes its mover takes the next one of the tile by shuffle, and
// one loop iteration advances every lane by one step of its own mover -- the verification of the held pick (RESUME) or one
// batch of kBatch probes.  Results are identical (same draws, same claims); only the mapping of movers to lanes changes.
// (Tried and dropped: one Philox block per lane and loop iteration with a queue of accepted draws, to take _randbelow's
// rejection loop out of the divergent part -- 15.8 lanes active instead of 11.9, but 4.9 ms instead of 2.8 ms for the
// FULL pass at 8192^2: more loop iterations, each paying the refill ballots / shuffles, and 80 registers.)
template <int kBatch>
__global__ void __launch_bounds__(128)
relocate_pass_packed(ShardView v, RelocWs ws, const uint32_t* __restrict__ mlist, int64_t nmovers, uint64_t seed, uint32_t epoch,
                     unsigned long long t_lo, unsigned long long t_hi, int mode, int64_t movers_per_warp) {
    // mlist != NULL: the movers of this pass are mlist[0 .. nmovers) (one rank-ordered bucket), otherwise 0 .. nmovers-1
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t next = warp_id * movers_per_warp;                       // first mover of the next tile (warp-uniform)
    const int64_t end = min(next + movers_per_warp, nmovers);
    if (next >= end) return;
    // the current tile: lane l holds mover tile0 + l; tile_mask = entries not handed out yet (in the bucket's time range)
    int64_t tile0 = 0;
    unsigned tile_mask = 0u;
    unsigned long long tile_t = 0ull;
    uint32_t tile_agent = 0u, tile_m = 0u;
    // this lane's mover
    int64_t m = -1;
    unsigned long long t = 0ull, held = kNoCell, soft = 0ull;
    int j = 0;                 // next probe
    bool verify = false;       // RESUME: the held pick has not been checked yet
    mb::DrawStream ds(seed, epoch, 0u);
    for (;;) {
        // ---- refill idle lanes from the tile (loading the next tile when it runs out)
        for (int round = 0; round < 2; ++round) {
            const unsigned idle = __ballot_sync(0xffffffffu, m < 0);
            if (idle == 0u) break;
            if (tile_mask == 0u) {
                if (next >= end) break;
                tile0 = next;
                next = min(next + 32, end);
                const int64_t mine = tile0 + lane;
                bool in_range = false;
                if (mine < next) {
                    tile_m = mlist != nullptr ? mlist[mine] : (uint32_t)mine;
                    tile_t = ws.mv_prio[tile_m];
                    tile_agent = ws.mv_agent[tile_m];
                    in_range = tile_t >= t_lo && tile_t <= t_hi;
                }
                tile_mask = __ballot_sync(0xffffffffu, in_range);
                if (tile_mask == 0u) { round = -1; continue; }   // nothing of this bucket in the tile: take the next one
            }
            const int rank = __popc(idle & lt_mask);
            const int avail = __popc(tile_mask);
            const int src = (m < 0 && rank < avail) ? (int)__fns(tile_mask, 0, rank + 1) : 0;
            const unsigned long long nt = __shfl_sync(0xffffffffu, tile_t, src);
            const uint32_t na = __shfl_sync(0xffffffffu, tile_agent, src);
            const uint32_t nm = __shfl_sync(0xffffffffu, tile_m, src);
            if (m < 0 && rank < avail) {
                m = (int64_t)nm;
                t = nt;
                ds = mb::DrawStream(seed, epoch, na);
                held = ws.choice[m];
                soft = 0ull;
                j = 0;
                verify = mode == 1;
            }
            const int taken = min(__popc(idle), avail);
            for (int q = 0; q < taken; ++q) tile_mask &= tile_mask - 1u;   // drop the entries handed out (lowest first)
        }
        if (__ballot_sync(0xffffffffu, m >= 0) == 0u) break;
        if (m < 0) continue;
        // ---- one step of this lane's mover
        if (verify) {
            verify = false;
            const uint32_t pr = ws.progress[m];
            if ((pr & 0xffu) >= (uint32_t)kMaxProbes) { m = -1; continue; }   // fallback state: served by the fallback phase
            Slot* hp = v.slot(held);
            const unsigned long long arr = load_slot(hp).arr;
            if (arr == t) { m = -1; continue; }                               // still the registered arrival of my pick
            if (arr > t) {                                                    // shadowed, and the shadow has left
                atomicMin(&hp->arr, t);
                mark_changed(ws);
                m = -1;
                continue;
            }
            ds.seek(pr >> 8);                                                 // an earlier mover took it: go on after it
            soft = ws.softmask[m];
            if (soft == 0ull) ws.softcell[m] = (uint32_t)held;
            soft |= 1ull << (pr & 0xffu);
            j = (int)(pr & 0xffu) + 1;
            held = kNoCell;                                                   // nothing of mine is registered any more
            continue;
        }
        unsigned long long cells[kBatch];
        uint32_t after[kBatch];
        Slot sl[kBatch];
        Slot* sp[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            cells[b] = ds.randbelow((uint64_t)v.ncells);
            after[b] = ds.c;
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            bool is_static;
            sp[b] = v.slot_and_static(cells[b], is_static);
            sl[b] = Slot{kNever, kNever};
            if (!is_static) sl[b] = load_slot(sp[b]);
        }
        bool done = false;
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const int jj = j + b;
            if (done || jj >= kMaxProbes) continue;
            if (empty_at(sl[b], t)) {
                if (cells[b] == held) {
                    if (atomicMin(&sp[b]->arr, t) > t) mark_changed(ws);      // re-assert (see relocate_pass)
                } else {
                    if (atomicMin(&sp[b]->arr, t) < t) {                      // lost the race for it: go on probing
                        if (soft == 0ull) ws.softcell[m] = (uint32_t)cells[b];
                        soft |= 1ull << jj;
                        continue;
                    }
                    if (mode == 0) release_held(v, ws, held, t);
                    ws.choice[m] = cells[b];
                    mark_changed(ws);
                }
                ws.progress[m] = (uint32_t)jj | (after[b] << 8);
                ws.softmask[m] = soft;
                done = true;
                continue;
            }
            if (sl[b].vac <= t) {
                if (soft == 0ull) ws.softcell[m] = (uint32_t)cells[b];
                soft |= 1ull << jj;
            }
        }
        j += kBatch;
        if (done) { m = -1; continue; }
        if (j >= kMaxProbes) {
            // all 50 probes occupied: fallback state
            ws.softmask[m] = soft;
            ws.progress[m] = (uint32_t)kMaxProbes;
            if (mode == 1) { ws.choice[m] = kNoCell; mark_changed(ws); }
            ws.fb_list[atomicAdd(&ws.counters[2], 1ull)] = (uint32_t)m;
            m = -1;
        }
    }
}

// ---- rank-ordered buckets as index lists ---------------------------------------------------------------------------------
// With many movers the passes work bucket by bucket (earliest priorities first).  Scanning ALL movers for the 1 / B that
// belong to the current bucket made every pass cost at least a sweep over the mover arrays (122 M movers per GPU on the
// 65536^2 grid: 3.1 ms per RESUME pass that did nothing else, profiles/r2_c5_shard_n1_launches.csv); here the mover
// indices are grouped by bucket once per step (histogram, host prefix, scatter) and a pass walks only its bucket's list.
__global__ void __launch_bounds__(256)
bucket_hist(RelocWs ws, int64_t nmovers, int shift, unsigned long long* __restrict__ counts) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nmovers; m += stride)
        atomicAdd(&h[(unsigned)(ws.mv_prio[m] >> shift)], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
bucket_scatter(RelocWs ws, int64_t nmovers, int shift, unsigned long long* __restrict__ cursors, int64_t per_cta) {
    __shared__ unsigned h[256];
    __shared__ unsigned long long base[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * per_cta, hi = min(lo + per_cta, nmovers);
    for (int64_t m = lo + threadIdx.x; m < hi; m += blockDim.x) atomicAdd(&h[(unsigned)(ws.mv_prio[m] >> shift)], 1u);
    __syncthreads();
    base[threadIdx.x] = h[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)h[threadIdx.x]) : 0ull;
    h[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t m = lo + threadIdx.x; m < hi; m += blockDim.x) {
        const unsigned b = (unsigned)(ws.mv_prio[m] >> shift);
        ws.bucket_list[base[b] + atomicAdd(&h[b], 1u)] = (uint32_t)m;    // order inside a bucket is irrelevant
    }
}

// RELEASE SCAN.  When the registered arrival of a cell leaves it (time t_r), the cell may be empty again for the
// movers that act later than t_r and had skipped it on the way to their pick.  Finding them needs no table traffic:
// every mover replays the draws of the probes BEFORE its pick (Philox, registers only) and looks each cell up in a
// shared-memory hash of the released cells; only a hit loads the slot, and if the cell is empty at the mover's time
// the mover moves back to it (releasing its own pick in turn, which lands in the next release set).  A fallback-state
// mover that finds one of its probes open again asks for a full pass (rare).
__global__ void __launch_bounds__(512)
relocate_release_scan(ShardView v, RelocWs ws, int64_t nmovers, uint64_t seed, uint32_t epoch, unsigned long long t_lo,
                      unsigned long long t_hi, const unsigned long long* __restrict__ rel_cell,
                      const unsigned long long* __restrict__ rel_time, int nrel, int nslots) {
    extern __shared__ __align__(16) unsigned long long hash_raw[];
    unsigned long long* hkey = hash_raw;            // cell + 1, 0 = free slot; nslots (a power of two >= 2 nrel) entries
    unsigned long long* htime = hash_raw + nslots;  // earliest release time of that cell in this set
    const unsigned hmask = (unsigned)nslots - 1u;
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) { hkey[i] = 0ull; htime[i] = kNever; }
    __syncthreads();
    for (int i = threadIdx.x; i < nrel; i += blockDim.x) {
        const unsigned long long key = rel_cell[i] + 1ull;
        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & hmask;
        while (true) {
            const unsigned long long prev = atomicCAS(&hkey[h], 0ull, key);
            if (prev == 0ull || prev == key) { atomicMin(&htime[h], rel_time[i]); break; }
            h = (h + 1) & hmask;
        }
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nmovers; m += stride) {
        const unsigned long long t = ws.mv_prio[m];
        if (t < t_lo || t > t_hi) continue;
        const uint32_t pr = ws.progress[m];
        const int upto = (int)(pr & 0xffu);  // probes before the pick (all kMaxProbes of a fallback-state mover)
        // only probes that were skipped because of an earlier ARRIVAL can re-open; most movers have none
        const unsigned long long soft = ws.softmask[m] & (upto >= 64 ? ~0ull : ((1ull << upto) - 1ull));
        if (soft == 0ull) continue;
        if ((soft & (soft - 1ull)) == 0ull) {   // exactly one candidate probe: its cell is cached, no replay unless it hits
            const unsigned long long key = (unsigned long long)ws.softcell[m] + 1ull;
            unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & hmask;
            bool hit = false;
            while (hkey[h] != 0ull) {
                if (hkey[h] == key) { hit = htime[h] < t; break; }
                h = (h + 1) & hmask;
            }
            if (!hit) continue;
        }
        const int last = 63 - __clzll((long long)soft);
        mb::DrawStream ds(seed, epoch, ws.mv_agent[m]);
        for (int j = 0; j <= last; ++j) {
            const unsigned long long c = ds.randbelow((uint64_t)v.ncells);
            if (((soft >> j) & 1ull) == 0ull) continue;
            const unsigned long long key = c + 1ull;
            unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & hmask;
            bool hit = false;
            while (hkey[h] != 0ull) {
                if (hkey[h] == key) { hit = htime[h] < t; break; }
                h = (h + 1) & hmask;
            }
            if (!hit || !empty_at(load_slot(v.slot(c)), t)) continue;
            if (upto >= kMaxProbes) {
                *(volatile unsigned long long*)&ws.counters[6] = 1ull;
            } else {
                commit_choice(v, ws, m, t, c);
                ws.progress[m] = (uint32_t)j | (ds.c << 8);
                ws.softmask[m] = soft & ((1ull << j) - 1ull);
            }
            break;
        }
    }
}

// ---- argwhere fallback, batched -----------------------------------------------------------------
// (time, rank k = _randbelow(nempty)) of the local fallback movers [first, first+count) of this pass
__global__ void fb_export(ShardView v, RelocWs ws, int64_t first, int count, int64_t nempty, uint64_t seed,
                          uint32_t epoch, unsigned long long* __restrict__ out_t,
                          unsigned long long* __restrict__ out_k, uint32_t* __restrict__ out_m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t m = ws.fb_list[first + i];
    // replay the 50 accepted probes to position the stream, then draw the rank (grid.py:308-310)
    mb::DrawStream ds(seed, epoch, ws.mv_agent[m]);
    for (int j = 0; j < kMaxProbes; ++j) (void)ds.randbelow((uint64_t)v.ncells);
    out_t[i] = ws.mv_prio[m];
    out_k[i] = ds.randbelow((uint64_t)nempty);
    out_m[i] = m;
}

// sort a batch of (time, k, mover) by time (bitonic, one CTA) into fb_t / fb_k / fb_m
__global__ void __launch_bounds__(1024)
fb_sort(RelocWs ws, const unsigned long long* __restrict__ in_t, const unsigned long long* __restrict__ in_k,
        const uint32_t* __restrict__ in_m, int count) {
    __shared__ unsigned long long st[kFbBatch];
    __shared__ uint32_t si[kFbBatch];
    for (int i = threadIdx.x; i < kFbBatch; i += blockDim.x) {
        st[i] = i < count ? in_t[i] : kNever;
        si[i] = (uint32_t)i;
    }
    __syncthreads();
    for (int k = 2; k <= kFbBatch; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < kFbBatch; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    if ((st[i] > st[p]) == up) {
                        const unsigned long long a = st[i]; st[i] = st[p]; st[p] = a;
                        const uint32_t b = si[i]; si[i] = si[p]; si[p] = b;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        ws.fb_t[i] = st[i];
        ws.fb_k[i] = in_k[si[i]];
        ws.fb_m[i] = in_m[si[i]];
    }
}

__device__ __forceinline__ int first_ge(const unsigned long long* __restrict__ a, int n, unsigned long long x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// cnt[b][f] = number of cells of LOCAL table block b that are empty at the time of fallback f (+ the sums per super-block,
// which must be zero on entry)
__global__ void __launch_bounds__(256)
fb_count(ShardView v, uint32_t* __restrict__ cnt, uint32_t* __restrict__ fb_tot, const unsigned long long* __restrict__ fb_t,
         int64_t block_cells, int count) {
    __shared__ int diff[kFbBatch + 1];
    __shared__ unsigned long long times[kFbBatch];
    __shared__ int warp_base[8];
    for (int i = threadIdx.x; i <= kFbBatch; i += blockDim.x) diff[i] = 0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) times[i] = fb_t[i];
    __syncthreads();
    const Slot* __restrict__ slots = v.slots[v.rank];
    const int64_t lo_c = (int64_t)blockIdx.x * block_cells;
    const int64_t hi_c = min(lo_c + block_cells, (int64_t)v.local_cells);
    int base = 0;
    for (int64_t c = lo_c + threadIdx.x; c < hi_c; c += blockDim.x) {
        const Slot s = load_slot(slots + c);
        if (s.vac == kNever) continue;                            // occupied during the whole step
        if (s.vac == 0 && s.arr == kNever) { ++base; continue; }  // empty during the whole step
        const int lo = first_ge(times, count, s.vac);             // first fallback with t >= vac
        const int hi = s.arr == kNever ? count : first_ge(times, count, s.arr + 1);  // first with t > arr
        if (lo < hi) {
            atomicAdd(&diff[lo], 1);
            atomicSub(&diff[hi], 1);
        }
    }
    base = (int)warp_reduce_add((unsigned)base);
    if ((threadIdx.x & 31) == 0) warp_base[threadIdx.x >> 5] = base;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < 8; ++w) run += warp_base[w];
        for (int f = 0; f < count; ++f) {  // prefix sum of the difference array (count <= 2048)
            run += diff[f];
            diff[f] = run;
        }
    }
    __syncthreads();
    uint32_t* __restrict__ sup = fb_sup(cnt) + (size_t)(blockIdx.x / kFbSuper) * kFbBatch;
    for (int f = threadIdx.x; f < count; f += blockDim.x) {
        cnt[(size_t)blockIdx.x * kFbBatch + f] = (uint32_t)diff[f];     // block-major: coalesced here and in the updates
        if (diff[f]) {
            atomicAdd(&sup[f], (uint32_t)diff[f]);
            atomicAdd(&fb_tot[f], (uint32_t)diff[f]);
        }
    }
}

// one warp per fallback mover whose k-th empty cell lies in the LOCAL table (k_local != kNoCell):
// walk the per-block counts 32 blocks at a time, then scan the block; result = global cell + 1
__global__ void __launch_bounds__(256)
fb_pick(ShardView v, RelocWs ws, const unsigned long long* __restrict__ fb_t,
        const unsigned long long* __restrict__ k_local, int nblocks, int64_t block_cells, int count,
        unsigned long long* __restrict__ result) {
    const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= count) return;
    unsigned long long k = k_local[f];
    if (k == kNoCell) return;
    // two-level walk: the super-block that holds the k-th empty cell, then the block inside it
    auto walk = [&](const uint32_t* __restrict__ base, int first, int n_entries, int& hit, unsigned long long& kk) {
        hit = -1;
        for (int b0 = 0; b0 < n_entries && hit < 0; b0 += 32) {
            const int b = b0 + lane;
            const unsigned n = b < n_entries ? base[(size_t)(first + b) * kFbBatch + f] : 0u;
            unsigned incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
            if (kk < total) {
                const int l = __ffs(__ballot_sync(0xffffffffu, kk < incl)) - 1;
                kk -= __shfl_sync(0xffffffffu, incl - n, l);
                hit = b0 + l;
            } else {
                kk -= total;
            }
        }
    };
    int block = -1, super = -1;
    unsigned need = 0;
    const int nsupers = (nblocks + kFbSuper - 1) / kFbSuper;
    walk(fb_sup(ws.cnt), 0, nsupers, super, k);
    if (super >= 0) {
        const int first = super * kFbSuper;
        int inner = -1;
        walk(ws.cnt, first, min(kFbSuper, nblocks - first), inner, k);
        if (inner >= 0) { block = first + inner; need = (unsigned)k; }
    }
    long long found = -1;
    if (block >= 0) {
        const unsigned long long t = fb_t[f];
        const Slot* __restrict__ slots = v.slots[v.rank];
        const int64_t lo_c = (int64_t)block * block_cells;
        const int64_t hi_c = min(lo_c + block_cells, (int64_t)v.local_cells);
        for (int64_t c0 = lo_c; c0 < hi_c; c0 += 32) {
            const int64_t c = c0 + lane;
            const bool e = c < hi_c && empty_at(load_slot(slots + c), t);
            const unsigned ballot = __ballot_sync(0xffffffffu, e);
            const unsigned n = __popc(ballot);
            if (need < n) {
                unsigned b = ballot;
                for (unsigned i = 0; i < need; ++i) b &= b - 1;  // drop the `need` lowest set bits
                found = c0 + (__ffs(b) - 1);
                break;
            }
            need -= n;
        }
    }
    if (lane == 0) result[f] = found >= 0 ? (unsigned long long)(v.local_lo + found) + 1ull : 0ull;
}

// commit the fallback picks of the movers that live on this rank (mover[f] != 0xffffffff)
__global__ void fb_commit(ShardView v, RelocWs ws, const unsigned long long* __restrict__ fb_t,
                          const uint32_t* __restrict__ mover, const unsigned long long* __restrict__ result, int count) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= count) return;
    const uint32_t m = mover[f];
    if (m == 0xffffffffu) return;
    if (result[f] != 0ull) commit_choice(v, ws, m, fb_t[f], result[f] - 1ull);
    else mark_changed(ws);  // the table moved under the scan (concurrent commits): force another pass
}

// ---- the fallback cascade as an EVENT LOOP ---------------------------------------------------------------------------
// What made the first version slow (60 ms for the first 8192^2 step) was not the movers but the ~200 argwhere-fallback
// movers among them: the k-th empty cell in row-major order at a mover's time depends on EVERY earlier move, so each
// generic iteration fixed roughly one more of them and shifted the picks of the later ones -- ~16 full-size iterations
// (a release scan over all movers + a pass over the whole slot table each) for a few hundred events.
//
// Once the ordinary movers have converged WITHOUT any fallback arrival (a fixed point of claims that only ever lower
// `arr`), the rest is a strictly time-ordered chain of small events, and processing them in time order needs neither
// iteration nor release:
//   * fallback mover f (time t_f): every mover earlier than t_f is final, so its pick -- the k-th cell with
//     vac <= t_f <= arr in row-major order -- is final too.  It is located through the per-block counts cnt[f][b]
//     (computed for all fallback times in ONE table pass, fb_count) and a scan of one block, and claimed (arr = t_f);
//   * a claim at time t on a cell whose registered arrival was a later mover (arr = a > t) DISPLACES that mover: it is
//     pushed as an event at its own time a and, when its turn comes, continues probing AFTER the lost pick (its earlier
//     probes were not empty at a and stay so: claims only lower `arr`); its new claim may displace an even later mover;
//   * every claim lowers `arr` of one cell from a_old to t: the cell stops being empty for the fallback times in
//     (t, a_old] -- their cnt / totals are decremented, so later fallback picks see the exact table.
// Events are taken in ascending time (the earliest pending displaced mover or the next fallback mover), all earlier
// movers are final at that point by induction, nothing is ever released, and the result is the sequential one.  The
// picks that changed are returned as an override list (time of the mover, new cell) which `apply_overrides` writes
// into the movers' `choice`; the termination criterion stays a confirming FULL iteration that changes nothing.
// A displaced mover that runs out of probes (it becomes a fallback mover itself, ~1e-5 of them) or any inconsistency
// ends the loop early (bail != 0) and the generic iteration takes over from the consistent state reached so far.
//
// One CTA: the chain is sequential by nature (each event is a handful of dependent L2 / NVLink accesses); the CTA's
// 1024 threads cooperate on the parts that are wide (block-count walk, block scan, count updates).  Row-sharded grids
// run it on one rank: slot tables and count tables of all ranks are peer-mapped.
constexpr int kPendCap = 1024;   // displaced movers waiting for their turn
constexpr int kOvCap = kOvCapWs;  // overrides returned per event loop

struct EventArgs {
    uint32_t* cnt[kMaxPeers];          // per-rank count tables: [kFbBlocksMax][kFbBatch] blocks + [kFbSupersMax][kFbBatch] super-blocks
    uint32_t* tot;                     // [world][kFbBatch] empties of each rank's table at each fallback time (updated)
    const unsigned long long* fb_t;    // [count] ascending fallback times
    const unsigned long long* fb_k;    // [count] rank of the empty cell to take over the WHOLE grid (row-major)
    unsigned long long* ov_time;       // [kOvCap] overrides: time of the mover ...
    unsigned long long* ov_cell;       //          ... and the global cell it holds now
    unsigned long long* out;           // [0] overrides, [1] bail reason (0 = completed), [2] fallback events, [3] displaced events
    int count;
};

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* warp_sums, unsigned& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    __syncthreads();                       // warp_sums of the previous use has been read
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    unsigned before = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {
        const unsigned s = warp_sums[w];
        before += w < warp ? s : 0u;
        tot += s;
    }
    total = tot;
    return before + incl - v;
}

// kBatched: the fallback events are taken in BATCHES of up to 32 (csrc comment "batched fallback events" below)
constexpr int kEvBatch = 32;      // events per batch = warps of the CTA
constexpr int kEvBmWords = 256;   // bitmap words per event: blocks of at most 8192 cells

template <bool kBatched>
__global__ void __launch_bounds__(1024, 1)
reloc_event_loop(ShardView v, EventArgs a, uint64_t seed, uint32_t epoch) {
    extern __shared__ __align__(16) unsigned ev_bitmaps[];   // [kEvBatch][kEvBmWords] (kBatched only)
    __shared__ unsigned long long times[kFbBatch];
    __shared__ unsigned long long pend_t[kPendCap], pend_c[kPendCap];
    __shared__ unsigned warp_sums[32];
    __shared__ int s_npend, s_kind, s_bail, s_owner, s_block, s_found;
    __shared__ unsigned s_need;
    __shared__ unsigned long long s_t, s_lost, s_cell, s_aold, s_vac, s_klocal, s_nov, s_nfb, s_ndis;
    const int tid = threadIdx.x;
    for (int f = tid; f < a.count; f += blockDim.x) times[f] = a.fb_t[f];
    if (tid == 0) { s_npend = 0; s_bail = 0; s_nov = 0ull; s_nfb = 0ull; s_ndis = 0ull; }
    __syncthreads();
    int i = 0;   // next fallback mover (uniform)
    while (true) {
        if (tid == 0) {
            int best = -1;
            unsigned long long bt = kNever;
            for (int p = 0; p < s_npend; ++p)
                if (pend_t[p] < bt) { bt = pend_t[p]; best = p; }
            const unsigned long long ft = i < a.count ? times[i] : kNever;
            if (best < 0 && i >= a.count) {
                s_kind = 0;
            } else if (best >= 0 && bt < ft) {
                s_kind = 1; s_t = bt; s_lost = pend_c[best];
                --s_npend;
                pend_t[best] = pend_t[s_npend]; pend_c[best] = pend_c[s_npend];
            } else {
                s_kind = 2; s_t = ft;
            }
        }
        __syncthreads();
        if (s_kind == 0 || s_bail) break;
        const unsigned long long t = s_t;
        if (s_kind == 1) {
            // ---- a displaced mover continues after the pick it lost.  Its probe cells are state-independent: the first
            // kDraws draws of its stream are computed by kDraws threads at once (draw d = half of Philox block d / 2),
            // the accepted ones (r < ncells, random.py:242-250) are numbered by a block scan = the probe sequence, and
            // the slots of the probes after the lost one are loaded by one thread each.
            constexpr int kDraws = 512;
            __shared__ unsigned long long probes[kMaxProbes];
            __shared__ int s_lostidx, s_pickidx;
            if (tid == 0) { s_lostidx = kMaxProbes; s_pickidx = kMaxProbes; }
            unsigned long long r = 0;
            unsigned ok = 0;
            if (tid < kDraws) {
                const mb::U4 blk = mb::philox_keyed(seed, epoch, (uint32_t)t, mb::kStreamDraw, (uint32_t)tid >> 1);
                const unsigned long long d = (tid & 1) ? (((unsigned long long)blk.z << 32) | blk.w) : (((unsigned long long)blk.x << 32) | blk.y);
                const int kbits = 64 - __clzll((long long)v.ncells);
                r = d >> (64 - kbits);
                ok = r < (unsigned long long)v.ncells;
            }
            unsigned total;
            const unsigned idx = block_exclusive_scan(ok, warp_sums, total);
            if (ok && idx < (unsigned)kMaxProbes) probes[idx] = r;
            __syncthreads();
            if (total < (unsigned)kMaxProbes) {
                if (tid == 0) s_bail = 9;        // fewer than 50 probes in 512 draws: practically impossible
            } else {
                if (tid < kMaxProbes && probes[tid] == s_lost) atomicMin(&s_lostidx, tid);
                __syncthreads();
                Slot sl{kNever, 0ull};
                if (tid < kMaxProbes && tid > s_lostidx) {
                    sl = load_slot(v.slot(probes[tid]));
                    if (empty_at(sl, t)) atomicMin(&s_pickidx, tid);
                }
                __syncthreads();
                if (tid == 0) {
                    if (s_lostidx >= kMaxProbes) s_bail = 3;           // the lost cell is not one of its probes
                    else if (s_pickidx >= kMaxProbes) s_bail = 2;      // out of probes: a new fallback mover
                    ++s_ndis;
                }
                if (tid == s_pickidx && tid < kMaxProbes) {
                    atomicMin(&v.slot(probes[tid])->arr, t);
                    s_cell = probes[tid]; s_aold = sl.arr; s_vac = sl.vac;
                }
            }
        } else if (kBatched) {
            // ---- batched fallback events.  One event costs a chain of dependent (on a row-sharded grid: remote) accesses --
            // walk the owner's super-block and block counts, scan the block, claim, fix the counts -- ~28 us over NVLink,
            // and 14 000 events of a 65536^2 step were 60 % of it.  The events are sequential only through the cells the
            // earlier ones took, so up to 32 consecutive ones (all earlier than the first pending displaced mover) are
            // LOCATED in parallel, one warp each, on the tables as they are (phase A: owner, super-block, block, and the
            // block's emptiness bitmap at the event's time), and then COMMITTED in time order by warp 0 (phase B): the
            // cells the earlier events of the batch took (at most 31, lane-resident) that the stale tables still count as
            // empty at this event's time either lie before the block -- the pick moves that many empties further -- or
            // inside it -- their bits are cleared.  If the corrected pick leaves the block, or a claim displaces a mover
            // (which must act before later events), the batch ends there and the rest is located again.  The first event
            // of a batch needs no correction, so every batch makes progress; the result is the sequential one.
            __shared__ int b_owner[kEvBatch], b_ok[kEvBatch], s_nbatch, s_done;
            __shared__ unsigned b_need[kEvBatch], b_ncell[kEvBatch], b_blk[kEvBatch];
            __shared__ unsigned long long b_lo[kEvBatch];
            const int w = tid >> 5, lane = tid & 31;
            if (tid == 0) {
                int nbm = a.count - i < kEvBatch ? a.count - i : kEvBatch;
                unsigned long long bt = kNever;
                for (int p = 0; p < s_npend; ++p) bt = pend_t[p] < bt ? pend_t[p] : bt;
                while (nbm > 1 && times[i + nbm - 1] > bt) --nbm;
                s_nbatch = nbm;
            }
            __syncthreads();
            const int nbatch = s_nbatch;
            // ---- phase A: warp w locates event i + w
            if (w < nbatch) {
                const int e = i + w;
                const unsigned long long te = times[e];
                unsigned long long k = a.fb_k[e];
                int ok = 1, owner = -1;
                {   // owner rank: first rank whose cumulative number of empties at te exceeds k
                    unsigned long long n = lane < v.world ? (unsigned long long)a.tot[(size_t)lane * kFbBatch + e] : 0ull, incl = n;
#pragma unroll
                    for (int o2 = 1; o2 < 32; o2 <<= 1) {
                        const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o2);
                        if (lane >= o2) incl += x;
                    }
                    const unsigned hit = __ballot_sync(0xffffffffu, k < incl);
                    if (hit == 0u) ok = 0;
                    else {
                        owner = __ffs(hit) - 1;
                        k -= __shfl_sync(0xffffffffu, incl - n, owner);
                    }
                }
                unsigned need = 0, ncell = 0, blk = 0;
                unsigned long long glo = 0;
                if (ok) {
                    const long long lc = (v.row_lo(owner + 1) - v.row_lo(owner)) * v.cols;
                    long long nb = (lc + 4095) / 4096;
                    if (nb > kFbBlocksMax) nb = kFbBlocksMax;
                    if (nb < 1) nb = 1;
                    const long long bc = (lc + nb - 1) / nb;
                    const int nsup = (int)((nb + kFbSuper - 1) / kFbSuper);
                    // one level of the walk: `cnt_n` entries base[(first + j) * kFbBatch + e], j < cnt_n <= 1024; lane l sums the
                    // chunk [l * per, (l + 1) * per) (all loads in flight at once), the warp finds the chunk, then its entry
                    auto level = [&](const uint32_t* base, int first, int cnt_n, int& found, unsigned long long& kk) {
                        const int per = (cnt_n + 31) / 32;
                        unsigned long long mine = 0;
                        for (int j = 0; j < per; ++j) {
                            const int idx = lane * per + j;
                            if (idx < cnt_n) mine += __ldcg(base + (size_t)(first + idx) * kFbBatch + e);
                        }
                        unsigned long long incl = mine;
#pragma unroll
                        for (int o2 = 1; o2 < 32; o2 <<= 1) {
                            const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o2);
                            if (lane >= o2) incl += x;
                        }
                        const unsigned hit = __ballot_sync(0xffffffffu, kk < incl);
                        if (hit == 0u) { found = -1; return; }
                        const int L = __ffs(hit) - 1;
                        kk -= __shfl_sync(0xffffffffu, incl - mine, L);
                        // the chunk of lane L, one entry per lane (per <= 32)
                        const int idx = L * per + lane;
                        const unsigned long long n = (lane < per && idx < cnt_n) ? (unsigned long long)__ldcg(base + (size_t)(first + idx) * kFbBatch + e) : 0ull;
                        unsigned long long inc2 = n;
#pragma unroll
                        for (int o2 = 1; o2 < 32; o2 <<= 1) {
                            const unsigned long long x = __shfl_up_sync(0xffffffffu, inc2, o2);
                            if (lane >= o2) inc2 += x;
                        }
                        const unsigned hit2 = __ballot_sync(0xffffffffu, kk < inc2);
                        if (hit2 == 0u) { found = -1; return; }
                        const int l2 = __ffs(hit2) - 1;
                        kk -= __shfl_sync(0xffffffffu, inc2 - n, l2);
                        found = L * per + l2;
                    };
                    int super = -1, inner = -1;
                    level(fb_sup(a.cnt[owner]), 0, nsup, super, k);
                    if (super >= 0) {
                        const int first = super * kFbSuper;
                        const int in_super = (int)(nb - first < kFbSuper ? nb - first : kFbSuper);
                        level(a.cnt[owner], first, in_super, inner, k);
                    }
                    if (super < 0 || inner < 0 || bc > 32 * kEvBmWords) ok = 0;
                    else {
                        blk = (unsigned)(super * kFbSuper + inner);
                        need = (unsigned)k;
                        const long long lo_c = (long long)blk * bc;
                        const long long hi_c = lo_c + bc < lc ? lo_c + bc : lc;
                        ncell = (unsigned)(hi_c - lo_c);
                        glo = (unsigned long long)(v.row_lo(owner) * v.cols + lo_c);
                        // the block's emptiness bitmap at te: 8 independent 16-byte loads per lane and round
                        const Slot* slots = v.slots[owner] + lo_c;
                        unsigned* bm = ev_bitmaps + (size_t)w * kEvBmWords;
                        for (unsigned w0 = 0; w0 < (unsigned)kEvBmWords; w0 += 8) {
                            Slot sl[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const unsigned c = (w0 + q) * 32u + lane;
                                sl[q] = Slot{kNever, 0ull};
                                if (c < ncell) sl[q] = load_slot(slots + c);
                            }
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const unsigned bits = __ballot_sync(0xffffffffu, empty_at(sl[q], te));
                                if (lane == q) bm[w0 + q] = bits;
                            }
                        }
                    }
                }
                if (lane == 0) { b_owner[w] = owner; b_ok[w] = ok; b_need[w] = need; b_ncell[w] = ncell; b_blk[w] = blk; b_lo[w] = glo; }
            }
            __syncthreads();
            // ---- phase B: warp 0 commits the events in time order
            if (w == 0) {
                unsigned long long S_cell = 0, S_vac = kNever, S_aold = 0;   // lane q: the cell event q of this batch took
                int done = 0;
                for (int q = 0; q < nbatch; ++q) {
                    const unsigned long long te = times[i + q];
                    if (!b_ok[q]) {
                        if (q == 0 && lane == 0) s_bail = 5;
                        break;
                    }
                    unsigned* bm = ev_bitmaps + (size_t)q * kEvBmWords;
                    const unsigned long long glo = b_lo[q];
                    const bool covered = lane < q && S_vac <= te && te <= S_aold;   // still counted as empty at te by the tables
                    const unsigned before = __popc(__ballot_sync(0xffffffffu, covered && S_cell < glo));
                    if (covered && S_cell >= glo && S_cell < glo + b_ncell[q]) {
                        const unsigned pos = (unsigned)(S_cell - glo);
                        atomicAnd(&bm[pos >> 5], ~(1u << (pos & 31u)));
                    }
                    __syncwarp();
                    const unsigned need = b_need[q] + before;
                    // the need-th set bit of the bitmap: lane l owns words 8 l .. 8 l + 7
                    unsigned wsum = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) wsum += __popc(bm[lane * 8 + j]);
                    unsigned incl = wsum;
#pragma unroll
                    for (int o2 = 1; o2 < 32; o2 <<= 1) {
                        const unsigned x = __shfl_up_sync(0xffffffffu, incl, o2);
                        if (lane >= o2) incl += x;
                    }
                    const unsigned hit = __ballot_sync(0xffffffffu, need < incl);
                    if (hit == 0u) {            // the pick left the block: locate this event again in the next batch
                        if (q == 0 && lane == 0) s_bail = 6;
                        break;
                    }
                    const int L = __ffs(hit) - 1;
                    unsigned left = need - __shfl_sync(0xffffffffu, incl - wsum, L);
                    int pos = -1;
                    if (lane == L) {
                        for (int j = 0; j < 8 && pos < 0; ++j) {
                            const unsigned word = bm[lane * 8 + j];
                            const unsigned n = __popc(word);
                            if (left < n) pos = (lane * 8 + j) * 32 + (int)__fns(word, 0u, left + 1u);
                            else left -= n;
                        }
                    }
                    pos = __shfl_sync(0xffffffffu, pos, L);
                    const int o = b_owner[q];
                    const unsigned long long cell = glo + (unsigned long long)pos;
                    Slot* sp = v.slots[o] + (cell - (unsigned long long)(v.row_lo(o) * v.cols));
                    unsigned long long aold = 0ull, vac = 0ull;
                    if (lane == 0) aold = atomicMin(&sp->arr, te);
                    if (lane == 1) vac = load_slot(sp).vac;
                    aold = __shfl_sync(0xffffffffu, aold, 0);
                    vac = __shfl_sync(0xffffffffu, vac, 1);
                    if (aold < te || vac > te) {            // not empty at te after all: the tables and the slots disagree
                        if (lane == 0) s_bail = 6;
                        break;
                    }
                    if (lane == q) { S_cell = cell; S_vac = vac; S_aold = aold; }
                    bool cut = false;
                    if (lane == 0) {
                        if (s_nov < (unsigned long long)kOvCap) { a.ov_time[s_nov] = te; a.ov_cell[s_nov] = cell; ++s_nov; }
                        else s_bail = 7;
                        ++s_nfb;
                        if (aold != kNever) {
                            if (s_npend < kPendCap) { pend_t[s_npend] = aold; pend_c[s_npend] = cell; ++s_npend; }
                            else s_bail = 8;
                        }
                    }
                    cut = aold != kNever;     // the displaced mover acts before the later events of the batch may
                    {   // the cell stops being empty for the fallback times in (te, aold]: fix the owner's counts
                        const unsigned blk = b_blk[q];
                        const unsigned long long from = vac > te + 1ull ? vac : te + 1ull;
                        const int lo = first_ge(times, a.count, from);
                        const int hi = aold == kNever ? a.count : first_ge(times, a.count, aold + 1ull);
                        uint32_t* cb = a.cnt[o] + (size_t)blk * kFbBatch;
                        uint32_t* cs = fb_sup(a.cnt[o]) + (size_t)(blk / kFbSuper) * kFbBatch;
                        for (int f = lo + lane; f < hi; f += 32) {
                            atomicSub(cb + f, 1u);
                            atomicSub(cs + f, 1u);
                            a.tot[(size_t)o * kFbBatch + f] -= 1u;
                        }
                    }
                    done = q + 1;
                    __syncwarp();
                    const int bail_now = *(volatile int*)&s_bail;    // uniform: every lane reads the shared word
                    if (cut || bail_now) break;
                }
                done = __shfl_sync(0xffffffffu, done, 0);
                if (lane == 0) s_done = done;
            }
            if (v.world > 1) __threadfence_system(); else __threadfence();
            __syncthreads();
            if (s_bail) break;
            i += s_done;
            continue;
        } else {
            // ---- fallback mover i: the k-th empty cell of the whole grid at its time
            if (tid == 0) {
                unsigned long long k = a.fb_k[i];
                int owner = -1;
                for (int r = 0; r < v.world; ++r) {
                    const unsigned long long n = a.tot[(size_t)r * kFbBatch + i];
                    if (k < n) { owner = r; break; }
                    k -= n;
                }
                s_owner = owner; s_klocal = k; s_found = 0;
                if (owner < 0) s_bail = 4;
            }
            __syncthreads();
            if (s_bail) break;
            const int o = s_owner;
            const long long lc = (v.row_lo(o + 1) - v.row_lo(o)) * v.cols;
            long long nb = (lc + 4095) / 4096;
            if (nb > kFbBlocksMax) nb = kFbBlocksMax;
            if (nb < 1) nb = 1;
            const long long bc = (lc + nb - 1) / nb;
            const int nsup = (int)((nb + kFbSuper - 1) / kFbSuper);
            {   // two-level walk of the owner's counts at fallback time i: super-block (one entry per thread), then its blocks
                const uint32_t* sup = fb_sup(a.cnt[o]);
                const unsigned n = tid < nsup ? __ldcg(sup + (size_t)tid * kFbBatch + i) : 0u;
                unsigned total;
                const unsigned excl = block_exclusive_scan(n, warp_sums, total);
                const unsigned long long k = s_klocal;
                if (k >= excl && k < (unsigned long long)excl + n) { s_block = tid; s_need = (unsigned)(k - excl); s_found = 1; }
            }
            __syncthreads();
            if (!s_found) { if (tid == 0) s_bail = 5; __syncthreads(); break; }
            {
                const int first = s_block * kFbSuper;
                const unsigned long long k = s_need;
                __syncthreads();
                if (tid == 0) s_found = 0;
                const int b = first + tid;
                const unsigned n = (tid < kFbSuper && b < nb) ? __ldcg(a.cnt[o] + (size_t)b * kFbBatch + i) : 0u;
                unsigned total;
                const unsigned excl = block_exclusive_scan(n, warp_sums, total);   // (the scan's first barrier orders the reset)
                if (k >= excl && k < (unsigned long long)excl + n) { s_block = b; s_need = (unsigned)(k - excl); s_found = 1; }
            }
            __syncthreads();
            if (!s_found) { if (tid == 0) s_bail = 5; __syncthreads(); break; }
            const Slot* slots = v.slots[o];
            const long long lo_c = (long long)s_block * bc;
            const long long hi_c = min(lo_c + bc, lc);
            unsigned need = s_need;
            __syncthreads();
            if (tid == 0) s_found = 0;
            // every thread takes kPer consecutive cells (independent 16-byte loads, all in flight at once), one block
            // scan per 1024 * kPer cells
            constexpr int kPer = 8;
            for (long long c0 = lo_c; c0 < hi_c; c0 += (long long)blockDim.x * kPer) {
                const long long mine = c0 + (long long)tid * kPer;
                Slot sl[kPer];
#pragma unroll
                for (int q = 0; q < kPer; ++q) {
                    sl[q] = Slot{kNever, 0ull};
                    if (mine + q < hi_c) sl[q] = load_slot(slots + mine + q);
                }
                unsigned e = 0;
#pragma unroll
                for (int q = 0; q < kPer; ++q) e += empty_at(sl[q], t);
                unsigned total;
                const unsigned excl = block_exclusive_scan(e, warp_sums, total);
                if (need < total) {
                    if (need >= excl && need < excl + e) {
                        unsigned left = need - excl;
#pragma unroll
                        for (int q = 0; q < kPer; ++q) {
                            if (empty_at(sl[q], t)) {
                                if (left == 0u) {
                                    const long long c = mine + q;
                                    s_cell = (unsigned long long)(v.row_lo(o) * v.cols + c); s_aold = sl[q].arr; s_vac = sl[q].vac;
                                    s_found = 1;
                                    atomicMin(&v.slots[o][c].arr, t);
                                }
                                --left;
                            }
                        }
                    }
                    break;
                }
                need -= total;
            }
            __syncthreads();
            if (!s_found) { if (tid == 0) s_bail = 6; __syncthreads(); break; }
            if (tid == 0) ++s_nfb;
        }
        __syncthreads();
        if (s_bail) break;
        // ---- the claim lowered arr of s_cell from s_aold to t: record it, wake the mover it displaced, fix the counts
        int o; unsigned long long l;
        v.locate(s_cell, o, l);
        if (tid == 0) {
            if (s_nov < (unsigned long long)kOvCap) { a.ov_time[s_nov] = t; a.ov_cell[s_nov] = s_cell; ++s_nov; }
            else s_bail = 7;
            if (s_aold != kNever) {
                if (s_npend < kPendCap) { pend_t[s_npend] = s_aold; pend_c[s_npend] = s_cell; ++s_npend; }
                else s_bail = 8;
            }
        }
        {
            const long long lc = (v.row_lo(o + 1) - v.row_lo(o)) * v.cols;
            long long nb = (lc + 4095) / 4096;
            if (nb > kFbBlocksMax) nb = kFbBlocksMax;
            if (nb < 1) nb = 1;
            const long long bc = (lc + nb - 1) / nb;
            const unsigned blk = (unsigned)(l / (unsigned long long)bc);
            const unsigned long long from = s_vac > t + 1ull ? s_vac : t + 1ull;
            const int lo = first_ge(times, a.count, from);
            const int hi = s_aold == kNever ? a.count : first_ge(times, a.count, s_aold + 1ull);
            uint32_t* cb = a.cnt[o] + (size_t)blk * kFbBatch;
            uint32_t* cs = fb_sup(a.cnt[o]) + (size_t)(blk / kFbSuper) * kFbBatch;
            for (int f = lo + tid; f < hi; f += blockDim.x) {     // consecutive words of one block row / super row
                atomicSub(cb + f, 1u);
                atomicSub(cs + f, 1u);
                a.tot[(size_t)o * kFbBatch + f] -= 1u;
            }
        }
        if (v.world > 1) __threadfence_system(); else __threadfence();
        if (s_kind == 2) ++i;
        __syncthreads();
    }
    if (tid == 0) {
        a.out[0] = s_nov; a.out[1] = (unsigned long long)s_bail; a.out[2] = s_nfb; a.out[3] = s_ndis;
    }
}

// choice[m] = the cell the event loop gave mover m (the LAST override recorded for its time); every rank applies the
// whole list to its own movers of the bucket.
__global__ void __launch_bounds__(512)
apply_overrides(RelocWs ws, int64_t nmovers, unsigned long long t_lo, unsigned long long t_hi,
                const unsigned long long* __restrict__ ov_time, const unsigned long long* __restrict__ ov_cell, int nov,
                int nslots) {
    extern __shared__ __align__(16) unsigned long long ov_hash[];
    unsigned long long* hkey = ov_hash;                     // time + 1, 0 = free
    int* hidx = reinterpret_cast<int*>(ov_hash + nslots);   // largest override index of that time
    const unsigned hmask = (unsigned)nslots - 1u;
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) { hkey[i] = 0ull; hidx[i] = -1; }
    __syncthreads();
    for (int i = threadIdx.x; i < nov; i += blockDim.x) {
        const unsigned long long key = ov_time[i] + 1ull;
        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & hmask;
        while (true) {
            const unsigned long long prev = atomicCAS(&hkey[h], 0ull, key);
            if (prev == 0ull || prev == key) { atomicMax(&hidx[h], i); break; }
            h = (h + 1) & hmask;
        }
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < nmovers; m += stride) {
        const unsigned long long t = ws.mv_prio[m];
        if (t < t_lo || t > t_hi) continue;
        const unsigned long long key = t + 1ull;
        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40) & hmask;
        while (hkey[h] != 0ull) {
            if (hkey[h] == key) { ws.choice[m] = ov_cell[hidx[h]]; break; }
            h = (h + 1) & hmask;
        }
    }
}

__global__ void relocate_vacate(ShardView v, RelocWs ws, int64_t nmovers) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    const uint32_t o = ws.mv_cell[m];
    v.type[v.rank][o] = -1;
    v.happy[v.rank][o] = 0;
}

__global__ void relocate_arrive(ShardView v, RelocWs ws, uint32_t* agent_cell, int64_t nmovers) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nmovers) return;
    int o; unsigned long long d;
    v.locate(ws.choice[m], o, d);
    const uint32_t agent = ws.mv_agent[m];
    v.type[o][d] = ws.mv_type[m];
    v.happy[o][d] = 0;  // the flag travels with the agent: movers are unhappy until the next assign_state
    v.agent_id[o][d] = agent;
    if (agent_cell != nullptr) agent_cell[agent] = (uint32_t)ws.choice[m];
}

int read_counters(const RelocWs& ws, unsigned long long* host7, cudaStream_t stream) {
    int rc = mb::check_cuda(cudaMemcpyAsync(host7, ws.counters, 7 * sizeof(unsigned long long),
                                            cudaMemcpyDeviceToHost, stream), "read counters");
    if (rc) return rc;
    return mb::check_cuda(cudaStreamSynchronize(stream), "sync");
}

inline int64_t fb_num_blocks(int64_t local_cells) {
    int64_t b = mb::ceil_div(local_cells > 0 ? local_cells : 1, 4096);
    return b > kFbBlocksMax ? kFbBlocksMax : b;
}

// the super-block sums accumulate with atomics: clear them before a count pass
inline int fb_clear_supers(uint32_t* cnt_table, int64_t nblocks, cudaStream_t stream) {
    return mb::check_cuda(cudaMemsetAsync(fb_sup(cnt_table), 0, sizeof(uint32_t) * (size_t)mb::ceil_div(nblocks, kFbSuper) * kFbBatch, stream),
                          "memset");
}

// serve `count` (<= kFbBatch) fallback entries whose (t, k, m) sit in device arrays, single rank
int serve_fallback_local(const ShardView& v, RelocWs& ws, const unsigned long long* in_t, const unsigned long long* in_k,
                         const uint32_t* in_m, int count, cudaStream_t stream) {
    const int64_t nblocks = fb_num_blocks(v.local_cells);
    const int64_t block_cells = mb::ceil_div(v.local_cells, nblocks);
    fb_sort<<<1, 1024, 0, stream>>>(ws, in_t, in_k, in_m, count);
    int rc = mb::check_cuda(cudaMemsetAsync(ws.fb_tot, 0, 4 * kFbBatch, stream), "memset");
    if (rc) return rc;
    if ((rc = fb_clear_supers(ws.cnt, nblocks, stream))) return rc;
    fb_count<<<(unsigned)nblocks, 256, 0, stream>>>(v, ws.cnt, ws.fb_tot, ws.fb_t, block_cells, count);
    const unsigned wblocks = (unsigned)mb::ceil_div((int64_t)count * 32, 256);
    fb_pick<<<wblocks, 256, 0, stream>>>(v, ws, ws.fb_t, ws.fb_k, (int)nblocks, block_cells, count, ws.fb_res);
    fb_commit<<<(unsigned)mb::ceil_div(count, 128), 128, 0, stream>>>(v, ws, ws.fb_t, ws.fb_m, ws.fb_res, count);
    MB_LAUNCH_CHECK("relocate fallback");
    return MB_OK;
}

}  // namespace

// ================================================================================================
// C ABI: phase-level entry points (used one by one by the row-sharded driver, mesa_b200/sharded.py)
// ================================================================================================
MB_API int mb_reloc_workspace_bytes(int64_t max_movers, size_t* out) {
    MB_REQUIRE(out != nullptr && max_movers >= 0, "mb_reloc_workspace_bytes: bad argument");
    *out = carve(nullptr, nullptr, max_movers);
    return MB_OK;
}

MB_API int mb_reloc_slots_init(const void* shard, cudaStream_t stream) {
    ShardView v;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if (v.local_cells > 0) init_slots<<<mb::sm_count() * 8, 256, 0, stream>>>(v);
    MB_LAUNCH_CHECK("mb_reloc_slots_init");
    return MB_OK;
}

MB_API int mb_reloc_collect(const void* shard, uint64_t seed, uint32_t epoch, void* workspace, size_t workspace_bytes,
                            int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.counters, 0, 8 * sizeof(unsigned long long), stream), "memset"))) return rc;
    if (v.local_cells > 0) {
        // one contiguous chunk of cells per CTA, a multiple of the CTA size, ~16 CTAs per SM
        int64_t chunk = mb::ceil_div(v.local_cells, (int64_t)mb::sm_count() * 16);
        chunk = mb::ceil_div(chunk, 256) * 256;
        collect_movers<<<(unsigned)mb::ceil_div(v.local_cells, chunk), 256, 0, stream>>>(v, ws, max_movers, seed, epoch, chunk);
    }
    MB_LAUNCH_CHECK("mb_reloc_collect");
    return MB_OK;
}

// copies counters [movers, occupied, fallback movers listed, table changed since the last take / full pass, -, entries of
// the release write set, a full pass is required] to the host (7 words); synchronises
MB_API int mb_reloc_counters(void* workspace, size_t workspace_bytes, int64_t max_movers, uint64_t* host7,
                             cudaStream_t stream) {
    RelocWs ws;
    int rc = get_ws(&ws, workspace, workspace_bytes, max_movers);
    if (rc) return rc;
    MB_REQUIRE(host7 != nullptr, "mb_reloc_counters: null output");
    return read_counters(ws, (unsigned long long*)host7, stream);
}

// start of an incremental iteration: moves the first `count` (<= 4096) entries of the release write set (cells whose
// registered arrival left, and the time of the mover that left) into caller buffers, clears the set and the changed flag
MB_API int mb_reloc_release_take(void* workspace, size_t workspace_bytes, int64_t max_movers, uint64_t* cells_out,
                                 uint64_t* times_out, int64_t count, cudaStream_t stream) {
    RelocWs ws;
    int rc = get_ws(&ws, workspace, workspace_bytes, max_movers);
    if (rc) return rc;
    MB_REQUIRE(count >= 0 && count <= kRelCap && (count == 0 || (cells_out && times_out)), "mb_reloc_release_take: bad argument");
    if (count > 0) {
        if ((rc = mb::check_cuda(cudaMemcpyAsync(cells_out, ws.rel_w_cell, 8 * count, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
        if ((rc = mb::check_cuda(cudaMemcpyAsync(times_out, ws.rel_w_time, 8 * count, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    }
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.counters + 3, 0, sizeof(unsigned long long), stream), "memset"))) return rc;
    return mb::check_cuda(cudaMemsetAsync(ws.counters + 5, 0, sizeof(unsigned long long), stream), "memset");
}

// release scan (see relocate_release_scan): the local movers in [t_lo, t_hi] whose EARLIER probes include one of the
// `nrel` (<= 4096) released cells move back to it if it is empty at their time.  Flags accumulate.
MB_API int mb_reloc_release_scan(const void* shard, int64_t nmovers, uint64_t seed, uint32_t epoch, uint64_t t_lo,
                                 uint64_t t_hi, const uint64_t* rel_cells, const uint64_t* rel_times, int64_t nrel,
                                 void* workspace, size_t workspace_bytes, int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(nmovers >= 0 && nmovers <= max_movers, "mb_reloc_release_scan: mover count exceeds the workspace");
    MB_REQUIRE(nrel >= 0 && nrel <= kRelCap && (nrel == 0 || (rel_cells && rel_times)), "mb_reloc_release_scan: bad release set");
    if (nmovers > 0 && nrel > 0) {
        int slots = 256;  // small sets (the usual case) get a small hash so that several CTAs fit on an SM
        while (slots < 2 * nrel) slots <<= 1;
        const int smem = 2 * slots * (int)sizeof(unsigned long long);
        static bool attr_set = false;
        if (!attr_set) {
            if ((rc = mb::check_cuda(cudaFuncSetAttribute(relocate_release_scan, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                          2 * 2 * kRelCap * (int)sizeof(unsigned long long)),
                                     "cudaFuncSetAttribute"))) return rc;
            attr_set = true;
        }
        const int64_t want = mb::ceil_div(nmovers, 512);
        const int64_t cap = (int64_t)mb::sm_count() * (smem <= 32768 ? 4 : 1);
        relocate_release_scan<<<(unsigned)(want < cap ? want : cap), 512, smem, stream>>>(
            v, ws, nmovers, seed, epoch, t_lo, t_hi, (const unsigned long long*)rel_cells, (const unsigned long long*)rel_times,
            (int)nrel, slots);
    }
    MB_LAUNCH_CHECK("mb_reloc_release_scan");
    return MB_OK;
}

// mode 0: FULL pass (every mover from its first probe; clears the fallback list, the flags and the release set first);
// mode 1: RESUME pass (movers only verify their pick and continue after it; flags accumulate).  RESUME passes are valid
// as long as every recorded release has been answered by a release scan (or a FULL pass).
namespace {
int launch_pass(const ShardView& v, RelocWs& ws, const uint32_t* mlist, int64_t count, uint64_t seed, uint32_t epoch,
                uint64_t t_lo, uint64_t t_hi, int mode, cudaStream_t stream) {
    int rc;
    if (mode == 0 && (rc = mb::check_cuda(cudaMemsetAsync(ws.counters + 2, 0, 5 * sizeof(unsigned long long), stream), "memset"))) return rc;
    if (count > 0) {
        static int batch = -1;
        if (batch < 0) {
            const char* e = getenv("MB_RELOC_BATCH");
            batch = e ? atoi(e) : 2;
        }
        static int packed = -1;
        if (packed < 0) {
            const char* e = getenv("MB_RELOC_PACKED");
            packed = e ? atoi(e) : 1;
        }
        if (packed || mlist != nullptr) {
            // one warp per `per_warp` consecutive movers: enough warps to fill the machine several times over, long
            // enough ranges for the lane refill to pay
            int64_t per_warp = 256;
            const int64_t want_warps = (int64_t)mb::sm_count() * 64 * 4;
            if (mb::ceil_div(count, per_warp) < want_warps) per_warp = mb::ceil_div(mb::ceil_div(count, want_warps), 32) * 32;
            if (per_warp < 32) per_warp = 32;
            const unsigned blocks = (unsigned)mb::ceil_div(mb::ceil_div(count, per_warp), 4);
            if (batch <= 1) relocate_pass_packed<1><<<blocks, 128, 0, stream>>>(v, ws, mlist, count, seed, epoch, t_lo, t_hi, mode, per_warp);
            else if (batch == 2) relocate_pass_packed<2><<<blocks, 128, 0, stream>>>(v, ws, mlist, count, seed, epoch, t_lo, t_hi, mode, per_warp);
            else relocate_pass_packed<4><<<blocks, 128, 0, stream>>>(v, ws, mlist, count, seed, epoch, t_lo, t_hi, mode, per_warp);
        } else {
            const unsigned blocks = (unsigned)mb::ceil_div(count, 128);
            if (batch <= 1) relocate_pass<1><<<blocks, 128, 0, stream>>>(v, ws, count, seed, epoch, t_lo, t_hi, mode);
            else if (batch == 2) relocate_pass<2><<<blocks, 128, 0, stream>>>(v, ws, count, seed, epoch, t_lo, t_hi, mode);
            else relocate_pass<4><<<blocks, 128, 0, stream>>>(v, ws, count, seed, epoch, t_lo, t_hi, mode);
        }
    }
    MB_LAUNCH_CHECK("mb_reloc_pass");
    return MB_OK;
}
}  // namespace

MB_API int mb_reloc_pass(const void* shard, int64_t nmovers, uint64_t seed, uint32_t epoch, uint64_t t_lo, uint64_t t_hi,
                         int mode, void* workspace, size_t workspace_bytes, int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(nmovers >= 0 && nmovers <= max_movers, "mb_reloc_pass: mover count exceeds the workspace");
    MB_REQUIRE(mode == 0 || mode == 1, "mb_reloc_pass: mode must be 0 (full) or 1 (resume)");
    return launch_pass(v, ws, nullptr, nmovers, seed, epoch, t_lo, t_hi, mode, stream);
}

// Group the mover indices by rank-ordered bucket (nbuckets: a power of two <= 256; bucket = top bits of the priority)
// into the workspace's bucket list and return the per-bucket counts (host array of nbuckets words).  Synchronises.
MB_API int mb_reloc_bucket_lists(int64_t nmovers, int nbuckets, uint64_t* counts_host, void* workspace, size_t workspace_bytes,
                                 int64_t max_movers, cudaStream_t stream) {
    RelocWs ws;
    int rc = get_ws(&ws, workspace, workspace_bytes, max_movers);
    if (rc) return rc;
    MB_REQUIRE(nmovers >= 0 && nmovers <= max_movers && counts_host != nullptr, "mb_reloc_bucket_lists: bad argument");
    MB_REQUIRE(nbuckets >= 2 && nbuckets <= 256 && (nbuckets & (nbuckets - 1)) == 0, "mb_reloc_bucket_lists: nbuckets must be a power of two in 2..256");
    int log2b = 0;
    while ((1 << log2b) < nbuckets) ++log2b;
    const int shift = 64 - log2b;
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.bucket_cur, 0, 8 * 256, stream), "memset"))) return rc;
    if (nmovers > 0) bucket_hist<<<mb::sm_count() * 8, 256, 0, stream>>>(ws, nmovers, shift, ws.bucket_cur);
    unsigned long long host[256];
    if ((rc = mb::check_cuda(cudaMemcpyAsync(host, ws.bucket_cur, 8 * 256, cudaMemcpyDeviceToHost, stream), "read bucket counts"))) return rc;
    if ((rc = mb::check_cuda(cudaStreamSynchronize(stream), "sync"))) return rc;
    unsigned long long cursors[256], run = 0;
    for (int b = 0; b < 256; ++b) {
        cursors[b] = run;
        run += host[b];
        if (b < nbuckets) counts_host[b] = host[b];
    }
    if ((rc = mb::check_cuda(cudaMemcpyAsync(ws.bucket_cur, cursors, 8 * 256, cudaMemcpyHostToDevice, stream), "write bucket cursors"))) return rc;
    if ((rc = mb::check_cuda(cudaStreamSynchronize(stream), "sync"))) return rc;     // `cursors` lives on this stack frame
    if (nmovers > 0) {
        const int64_t per_cta = mb::ceil_div(mb::ceil_div(nmovers, (int64_t)mb::sm_count() * 16), 256) * 256;
        bucket_scatter<<<(unsigned)mb::ceil_div(nmovers, per_cta), 256, 0, stream>>>(ws, nmovers, shift, ws.bucket_cur, per_cta);
    }
    MB_LAUNCH_CHECK("mb_reloc_bucket_lists");
    return MB_OK;
}

// mb_reloc_pass over ONE bucket's list: the movers bucket_list[list_first .. list_first + list_count)
MB_API int mb_reloc_pass_list(const void* shard, int64_t list_first, int64_t list_count, uint64_t seed, uint32_t epoch, uint64_t t_lo,
                              uint64_t t_hi, int mode, void* workspace, size_t workspace_bytes, int64_t max_movers,
                              cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(list_first >= 0 && list_count >= 0 && list_first + list_count <= max_movers, "mb_reloc_pass_list: list range exceeds the workspace");
    MB_REQUIRE(mode == 0 || mode == 1, "mb_reloc_pass_list: mode must be 0 (full) or 1 (resume)");
    return launch_pass(v, ws, ws.bucket_list + list_first, list_count, seed, epoch, t_lo, t_hi, mode, stream);
}

// (time, k = _randbelow(nempty), local mover) of this rank's fallback movers [first, first+count) of the pass
MB_API int mb_reloc_fb_export(const void* shard, int64_t first, int count, int64_t nempty, uint64_t seed, uint32_t epoch,
                              uint64_t* out_t, uint64_t* out_k, uint32_t* out_m, void* workspace, size_t workspace_bytes,
                              int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(count >= 0 && nempty > 0 && out_t && out_k && out_m, "mb_reloc_fb_export: bad argument");
    if (count > 0)
        fb_export<<<(unsigned)mb::ceil_div(count, 128), 128, 0, stream>>>(v, ws, first, count, nempty, seed, epoch,
                                                                           (unsigned long long*)out_t, (unsigned long long*)out_k, out_m);
    MB_LAUNCH_CHECK("mb_reloc_fb_export");
    return MB_OK;
}

// empties of the LOCAL table at each of `count` (<= 2048) ascending times -> totals[count]
MB_API int mb_reloc_fb_count(const void* shard, const uint64_t* times_sorted, int count, uint32_t* totals, void* workspace,
                             size_t workspace_bytes, int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(count >= 0 && count <= kFbBatch && times_sorted && totals, "mb_reloc_fb_count: bad argument");
    if (count == 0) return MB_OK;
    const int64_t nblocks = fb_num_blocks(v.local_cells);
    if ((rc = mb::check_cuda(cudaMemsetAsync(ws.fb_tot, 0, 4 * kFbBatch, stream), "memset"))) return rc;
    if ((rc = fb_clear_supers(ws.cnt, nblocks, stream))) return rc;
    fb_count<<<(unsigned)nblocks, 256, 0, stream>>>(v, ws.cnt, ws.fb_tot, (const unsigned long long*)times_sorted,
                                                   mb::ceil_div(v.local_cells, nblocks), count);
    if ((rc = mb::check_cuda(cudaMemcpyAsync(totals, ws.fb_tot, 4 * (size_t)count, cudaMemcpyDeviceToDevice, stream), "copy"))) return rc;
    MB_LAUNCH_CHECK("mb_reloc_fb_count");
    return MB_OK;
}

// after mb_reloc_fb_count with the same times: result[f] = global cell + 1 of the k_local[f]-th empty cell of the
// LOCAL table at times_sorted[f] (k_local[f] == ~0: not on this rank -> untouched)
MB_API int mb_reloc_fb_pick(const void* shard, const uint64_t* times_sorted, const uint64_t* k_local, int count,
                            uint64_t* result, void* workspace, size_t workspace_bytes, int64_t max_movers,
                            cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(count >= 0 && count <= kFbBatch && times_sorted && k_local && result, "mb_reloc_fb_pick: bad argument");
    if (count == 0) return MB_OK;
    const int64_t nblocks = fb_num_blocks(v.local_cells);
    fb_pick<<<(unsigned)mb::ceil_div((int64_t)count * 32, 256), 256, 0, stream>>>(
        v, ws, (const unsigned long long*)times_sorted, (const unsigned long long*)k_local, (int)nblocks,
        mb::ceil_div(v.local_cells, nblocks), count, (unsigned long long*)result);
    MB_LAUNCH_CHECK("mb_reloc_fb_pick");
    return MB_OK;
}

// claim result[f]-1 for the local mover movers[f] (0xffffffff = not mine) at times_sorted[f]
MB_API int mb_reloc_fb_commit(const void* shard, const uint64_t* times_sorted, const uint32_t* movers,
                              const uint64_t* result, int count, void* workspace, size_t workspace_bytes,
                              int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    MB_REQUIRE(count >= 0 && times_sorted && movers && result, "mb_reloc_fb_commit: bad argument");
    if (count > 0)
        fb_commit<<<(unsigned)mb::ceil_div(count, 128), 128, 0, stream>>>(v, ws, (const unsigned long long*)times_sorted,
                                                                           movers, (const unsigned long long*)result, count);
    MB_LAUNCH_CHECK("mb_reloc_fb_commit");
    return MB_OK;
}

// phase 0: every origin becomes empty; phase 1 (after ALL ranks finished phase 0): every arrival is written
MB_API int mb_reloc_apply(const void* shard, int phase, int64_t nmovers, uint32_t* agent_cell, void* workspace,
                          size_t workspace_bytes, int64_t max_movers, cudaStream_t stream) {
    ShardView v;
    RelocWs ws;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    if (nmovers > 0) {
        const unsigned blocks = (unsigned)mb::ceil_div(nmovers, 128);
        if (phase == 0) relocate_vacate<<<blocks, 128, 0, stream>>>(v, ws, nmovers);
        else relocate_arrive<<<blocks, 128, 0, stream>>>(v, ws, agent_cell, nmovers);
    }
    MB_LAUNCH_CHECK("mb_reloc_apply");
    return MB_OK;
}

// Count the empties of the LOCAL table at `count` (<= 2048) ascending times into a caller-owned block-count table
// cnt_table [65536 + 1024][2048] uint32 (blocks, then super-blocks of 64; peer-mapped memory on a row-sharded grid: the event loop of another rank reads and
// updates it) and totals[count].
MB_API int mb_reloc_fb_count_into(const void* shard, const uint64_t* times_sorted, int count, uint32_t* totals,
                                  uint32_t* cnt_table, cudaStream_t stream) {
    ShardView v;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    MB_REQUIRE(count >= 0 && count <= kFbBatch && times_sorted && totals && cnt_table, "mb_reloc_fb_count_into: bad argument");
    if (count == 0) return MB_OK;
    const int64_t nblocks = fb_num_blocks(v.local_cells);
    if ((rc = mb::check_cuda(cudaMemsetAsync(totals, 0, 4 * (size_t)count, stream), "memset"))) return rc;
    if ((rc = fb_clear_supers(cnt_table, nblocks, stream))) return rc;
    fb_count<<<(unsigned)nblocks, 256, 0, stream>>>(v, cnt_table, totals, (const unsigned long long*)times_sorted,
                                                   mb::ceil_div(v.local_cells, nblocks), count);
    MB_LAUNCH_CHECK("mb_reloc_fb_count_into");
    return MB_OK;
}

// The fallback cascade of one converged bucket as a time-ordered event loop (see reloc_event_loop).  cnt_tables_host:
// `world` device pointers (host array) to the ranks' block-count tables filled by mb_reloc_fb_count_into with the same
// times; totals [world][2048] uint32 (this rank's memory; row r = rank r's totals, updated in place); fb_times /
// fb_ranks [count]: ascending times and k = _randbelow(nempty) of the fallback movers of ALL ranks; overrides
// (time, cell) [4096] and out4 = {overrides, bail reason (0 = completed), fallback events, displaced events} are
// written.  Runs on ONE rank of a row-sharded grid (every other rank must be idle on the slot tables meanwhile).
MB_API int mb_reloc_event_loop(const void* shard, void* const* cnt_tables_host, uint32_t* totals, const uint64_t* fb_times,
                               const uint64_t* fb_ranks, int count, uint64_t* ov_times, uint64_t* ov_cells, uint64_t* out4,
                               uint64_t seed, uint32_t epoch, cudaStream_t stream) {
    ShardView v;
    int rc = make_view((const ShardDesc*)shard, &v);
    if (rc) return rc;
    MB_REQUIRE(count >= 0 && count <= kFbBatch && cnt_tables_host && totals && fb_times && fb_ranks && ov_times && ov_cells && out4,
               "mb_reloc_event_loop: bad argument");
    EventArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < v.world; ++r) {
        MB_REQUIRE(cnt_tables_host[r] != nullptr, "mb_reloc_event_loop: null count table");
        a.cnt[r] = (uint32_t*)cnt_tables_host[r];
    }
    a.tot = totals; a.fb_t = (const unsigned long long*)fb_times; a.fb_k = (const unsigned long long*)fb_ranks;
    a.ov_time = (unsigned long long*)ov_times; a.ov_cell = (unsigned long long*)ov_cells; a.out = (unsigned long long*)out4;
    a.count = count;
    static int batched = -1;
    if (batched < 0) {
        const char* e = getenv("MB_RELOC_EVENT_BATCH");     // 0: one event at a time (the first version, kept for cross-checks)
        batched = e ? atoi(e) : 1;
        if (batched) {
            rc = mb::check_cuda(cudaFuncSetAttribute(reloc_event_loop<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     kEvBatch * kEvBmWords * (int)sizeof(unsigned)), "cudaFuncSetAttribute");
            if (rc) return rc;
        }
    }
    // Batches pay on one GPU (8192 x 65536 cells, 1780 events: 24 -> 14 ms) and lose on a row-sharded grid, where the 32
    // speculative block scans of a batch are remote reads of 128 KiB each through ONE SM (65536^2 on 8 GPUs, 14 926
    // events: 419 ms one at a time, 899 ms batched) -- MB_RELOC_EVENT_BATCH=2 forces them there (the multi-GPU parity
    // tests pass either way)
    if (batched >= 2 || (batched == 1 && v.world == 1))
        reloc_event_loop<true><<<1, 1024, kEvBatch * kEvBmWords * sizeof(unsigned), stream>>>(v, a, seed, epoch);
    else reloc_event_loop<false><<<1, 1024, 0, stream>>>(v, a, seed, epoch);
    MB_LAUNCH_CHECK("mb_reloc_event_loop");
    return MB_OK;
}

// choice of the local movers in [t_lo, t_hi] <- the event loop's overrides (the last one recorded for a mover wins)
MB_API int mb_reloc_apply_overrides(int64_t nmovers, uint64_t t_lo, uint64_t t_hi, const uint64_t* ov_times,
                                    const uint64_t* ov_cells, int count, void* workspace, size_t workspace_bytes,
                                    int64_t max_movers, cudaStream_t stream) {
    RelocWs ws;
    int rc = get_ws(&ws, workspace, workspace_bytes, max_movers);
    if (rc) return rc;
    MB_REQUIRE(count >= 0 && count <= kOvCap && (count == 0 || (ov_times && ov_cells)), "mb_reloc_apply_overrides: bad argument");
    MB_REQUIRE(nmovers >= 0 && nmovers <= max_movers, "mb_reloc_apply_overrides: mover count exceeds the workspace");
    if (count == 0 || nmovers == 0) return MB_OK;
    int slots = 256;
    while (slots < 2 * count) slots <<= 1;
    const int smem = slots * (int)(sizeof(unsigned long long) + sizeof(int));
    static bool attr_set = false;
    if (!attr_set) {
        if ((rc = mb::check_cuda(cudaFuncSetAttribute(apply_overrides, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      2 * kOvCap * (int)(sizeof(unsigned long long) + sizeof(int))),
                                 "cudaFuncSetAttribute"))) return rc;
        attr_set = true;
    }
    const int64_t want = mb::ceil_div(nmovers, 512);
    const int64_t cap = (int64_t)mb::sm_count() * 2;
    apply_overrides<<<(unsigned)(want < cap ? want : cap), 512, smem, stream>>>(
        ws, nmovers, t_lo, t_hi, (const unsigned long long*)ov_times, (const unsigned long long*)ov_cells, count, slots);
    MB_LAUNCH_CHECK("mb_reloc_apply_overrides");
    return MB_OK;
}

// ================================================================================================
// Single-GPU driver: the whole shuffle_do("step") in one call.  Synchronises `stream` (needs the mover
// count and the per-pass flags on the host).
// ================================================================================================
MB_API int mb_reloc_table_bytes(int64_t rows, int64_t cols, size_t* out) {
    MB_REQUIRE(out != nullptr && rows >= 0 && cols >= 0, "mb_reloc_table_bytes: bad argument");
    const int64_t n = rows * cols;
    *out = (size_t)n * sizeof(Slot) + (size_t)mb::ceil_div(n, 32) * sizeof(uint32_t);
    return MB_OK;
}

MB_API int mb_schelling_relocate(int8_t* type, uint8_t* happy, uint32_t* agent_id, uint32_t* agent_cell, void* slots,
                                 size_t slots_bytes, int64_t rows, int64_t cols, uint64_t seed, uint32_t epoch, void* workspace,
                                 size_t workspace_bytes, int64_t max_movers, int64_t* out_movers,
                                 int* out_iterations, cudaStream_t stream) {
    if (out_movers) *out_movers = 0;
    if (out_iterations) *out_iterations = 0;
    MB_REQUIRE(rows >= 0 && cols >= 0, "mb_schelling_relocate: negative shape");
    if (rows == 0 || cols == 0) return MB_OK;
    MB_REQUIRE(type && happy && agent_id && slots && workspace, "mb_schelling_relocate: null buffer");
    ShardDesc d;
    memset(&d, 0, sizeof(d));
    d.world = 1; d.rank = 0; d.rows_total = rows; d.cols = cols;
    d.slots[0] = slots; d.type[0] = type; d.happy[0] = happy; d.agent_id[0] = agent_id;
    // the table buffer = 16-byte slots of all cells [+ the bitmap of the agents that stay put (mb_reloc_table_bytes)]
    const size_t slot_bytes = (size_t)(rows * cols) * sizeof(Slot);
    MB_REQUIRE(slots_bytes >= slot_bytes, "mb_schelling_relocate: the slot table is smaller than 16 bytes per cell");
    if (slots_bytes >= slot_bytes + (size_t)mb::ceil_div(rows * cols, 32) * sizeof(uint32_t) && getenv("MB_RELOC_NOBITS") == nullptr) {
        d.static_bits = (char*)slots + slot_bytes;
        d.bits_stride = mb::ceil_div(rows * cols, 32);
    }
    ShardView v;
    RelocWs ws;
    int rc = make_view(&d, &v);
    if (rc) return rc;
    if ((rc = get_ws(&ws, workspace, workspace_bytes, max_movers))) return rc;
    if ((rc = mb_reloc_collect(&d, seed, epoch, workspace, workspace_bytes, max_movers, stream))) return rc;
    unsigned long long hc[7];
    if ((rc = read_counters(ws, hc, stream))) return rc;
    const int64_t nmovers = (int64_t)hc[0], noccupied = (int64_t)hc[1];
    if (out_movers) *out_movers = nmovers;
    if (nmovers == 0) return MB_OK;
    if (nmovers > max_movers) {
        mb::set_error("mb_schelling_relocate: %lld movers exceed workspace capacity %lld", (long long)nmovers,
                      (long long)max_movers);
        return MB_ERR_WORKSPACE;
    }
    const int64_t nempty = v.ncells - noccupied;
    // grid.py:311-315: a completely full grid raises ValueError at the first unhappy agent
    MB_REQUIRE(nempty > 0, "Grid is completely full. No empty cells available. Cannot select a random empty cell.");
    // rank-ordered buckets: a power of two, ~16 M movers each (smaller buckets lose to the fixed cost of a pass:
    // measured on 8192^2 / 15 M movers, 1 bucket 79 ms, 8 buckets 96 ms)
    int buckets = 1;
    while (buckets < 256 && nmovers / buckets > (16ll << 20)) buckets <<= 1;
    {   // MB_RELOC_BUCKETS=n (a power of two <= 256) forces the bucket count: how the parity suite covers the bucketed path
        static int forced = -1;
        if (forced < 0) {
            const char* e = getenv("MB_RELOC_BUCKETS");
            forced = e ? atoi(e) : 0;
        }
        if (forced >= 1 && forced <= 256 && (forced & (forced - 1)) == 0) buckets = forced;
    }
    int pass = 0;
    // several buckets: group the mover indices by bucket once; a pass then walks only its bucket's list
    unsigned long long bucket_off[257] = {0};
    if (buckets > 1) {
        uint64_t counts[256];
        if ((rc = mb_reloc_bucket_lists(nmovers, buckets, counts, workspace, workspace_bytes, max_movers, stream))) return rc;
        for (int b = 0; b < buckets; ++b) bucket_off[b + 1] = bucket_off[b] + counts[b];
    }
    auto run_pass = [&](int b, uint64_t t_lo, uint64_t t_hi, int mode) -> int {
        if (buckets > 1)
            return mb_reloc_pass_list(&d, (int64_t)bucket_off[b], (int64_t)(bucket_off[b + 1] - bucket_off[b]), seed, epoch, t_lo, t_hi,
                                      mode, workspace, workspace_bytes, max_movers, stream);
        return mb_reloc_pass(&d, nmovers, seed, epoch, t_lo, t_hi, mode, workspace, workspace_bytes, max_movers, stream);
    };
    for (int b = 0; b < buckets; ++b) {
        const unsigned long long width = buckets > 1 ? (1ull << 63) / (buckets / 2) : 0ull;  // 2^64 / buckets
        const unsigned long long t_lo = buckets > 1 ? width * (unsigned long long)b : 0ull;
        const unsigned long long t_hi = (buckets > 1 && b + 1 < buckets) ? t_lo + (width - 1) : kNever;
        // 1. the ordinary movers first: a FULL pass and then RESUME passes until nothing changes -- the fallback movers
        //    (all 50 probes occupied) are only listed, they hold nothing yet, so claims only ever lower `arr` and no
        //    release can occur: this converges to the exact sequential outcome of a world without fallback arrivals;
        // 2. the fallback movers and everything they displace, in time order, by the event loop (reloc_event_loop);
        // 3. the generic iteration below, starting with a FULL one, CONFIRMS: it ends at once when nothing changes
        //    and otherwise (the event loop bailed out, > 2048 fallback movers, ...) iterates to the fixed point as before.
        bool prefix_ok = getenv("MB_RELOC_GENERIC") == nullptr;
        bool settled = false;   // the bucket reached its fixed point by construction (no confirming iteration needed)
        if (prefix_ok) {
            if ((rc = run_pass(b, t_lo, t_hi, 0))) return rc;
            ++pass;
            for (int guard = 0; guard < kMaxPasses; ++guard) {
                if ((rc = read_counters(ws, hc, stream))) return rc;
                if (hc[5] != 0 || hc[6] != 0) { prefix_ok = false; break; }   // cannot happen (no releases); be safe
                if (guard > 0 && hc[3] == 0) break;                            // a RESUME pass changed nothing
                if ((rc = mb_reloc_release_take(workspace, workspace_bytes, max_movers, nullptr, nullptr, 0, stream))) return rc;
                if ((rc = run_pass(b, t_lo, t_hi, 1))) return rc;
                ++pass;
            }
            const int64_t nfb = (int64_t)hc[2];
            settled = prefix_ok && nfb == 0;   // claims only ever lowered `arr`: the RESUME fixed point is the sequential result
            if (prefix_ok && nfb > 0 && nfb <= kFbBatch) {
                unsigned long long* tmp_t = ws.fb_res;
                unsigned long long* tmp_k = reinterpret_cast<unsigned long long*>(ws.cnt);
                uint32_t* tmp_m = ws.fb_block;
                fb_export<<<(unsigned)mb::ceil_div(nfb, 128), 128, 0, stream>>>(v, ws, 0, (int)nfb, nempty, seed, epoch, tmp_t, tmp_k, tmp_m);
                fb_sort<<<1, 1024, 0, stream>>>(ws, tmp_t, tmp_k, tmp_m, (int)nfb);
                MB_LAUNCH_CHECK("relocate fallback export");
                if ((rc = mb_reloc_fb_count_into(&d, (const uint64_t*)ws.fb_t, (int)nfb, ws.fb_tot, ws.cnt, stream))) return rc;
                void* tables[1] = {ws.cnt};
                if ((rc = mb_reloc_event_loop(&d, tables, ws.fb_tot, (const uint64_t*)ws.fb_t, (const uint64_t*)ws.fb_k, (int)nfb,
                                              (uint64_t*)ws.ov_time, (uint64_t*)ws.ov_cell, (uint64_t*)ws.ev_out, seed, epoch, stream))) return rc;
                unsigned long long ev[4];
                if ((rc = mb::check_cuda(cudaMemcpyAsync(ev, ws.ev_out, sizeof(ev), cudaMemcpyDeviceToHost, stream), "read event loop result"))) return rc;
                if ((rc = mb::check_cuda(cudaStreamSynchronize(stream), "sync"))) return rc;
                if ((rc = mb_reloc_apply_overrides(nmovers, t_lo, t_hi, (const uint64_t*)ws.ov_time, (const uint64_t*)ws.ov_cell,
                                                   (int)ev[0], workspace, workspace_bytes, max_movers, stream))) return rc;
                if (getenv("MB_RELOC_DEBUG") != nullptr)
                    fprintf(stderr, "[mb reloc] bucket %d: %lld fallback movers, event loop: %llu overrides, bail %llu, %llu fallback + %llu displaced events, %d passes so far\n",
                            b, (long long)nfb, ev[0], ev[1], ev[2], ev[3], pass);
                settled = ev[1] == 0ull;       // the event loop processed every event in time order and released nothing
            }
        }
        // The confirming FULL iteration (one pass over every mover from its first probe + the fallback phase over the whole
        // slot table: 3.8 of the 20 ms of a first 8192^2 step) never found work in any test or benchmark run; it is kept
        // behind MB_RELOC_CONFIRM=1 (the parity suite runs both ways) and whenever the event loop bailed out or was not run.
        static int confirm = -1;
        if (confirm < 0) {
            const char* e = getenv("MB_RELOC_CONFIRM");
            confirm = e ? atoi(e) : 0;
        }
        if (settled && !confirm) continue;
        // An ITERATION = [movers re-evaluate] + [fallback phase].
        bool full = true, confirmed = prefix_ok;
        int64_t nrel = 0;
        for (;; ++pass) {
            if (pass >= kMaxPasses) {
                mb::set_error("mb_schelling_relocate: no convergence after %d passes", pass);
                return MB_ERR_CUDA;
            }
            if (full) {
                if ((rc = run_pass(b, t_lo, t_hi, 0))) return rc;
            } else {
                if ((rc = mb_reloc_release_take(workspace, workspace_bytes, max_movers, (uint64_t*)ws.rel_r_cell,
                                                (uint64_t*)ws.rel_r_time, nrel, stream))) return rc;
                if (nrel > 0 && (rc = mb_reloc_release_scan(&d, nmovers, seed, epoch, t_lo, t_hi, (const uint64_t*)ws.rel_r_cell,
                                                            (const uint64_t*)ws.rel_r_time, nrel, workspace, workspace_bytes,
                                                            max_movers, stream))) return rc;
                if ((rc = run_pass(b, t_lo, t_hi, 1))) return rc;
            }
            if ((rc = read_counters(ws, hc, stream))) return rc;
            const int64_t nfb = (int64_t)hc[2];
            if (nfb > 0) {
                for (int64_t first = 0; first < nfb; first += kFbBatch) {
                    const int count = (int)(nfb - first < kFbBatch ? nfb - first : kFbBatch);
                    // fb_res / cnt / fb_block double as the export buffers: the sort consumes them before
                    // fb_count / fb_pick overwrite them (same stream)
                    unsigned long long* tmp_t = ws.fb_res;
                    unsigned long long* tmp_k = reinterpret_cast<unsigned long long*>(ws.cnt);
                    uint32_t* tmp_m = ws.fb_block;
                    fb_export<<<(unsigned)mb::ceil_div(count, 128), 128, 0, stream>>>(v, ws, first, count, nempty, seed, epoch,
                                                                                       tmp_t, tmp_k, tmp_m);
                    if ((rc = serve_fallback_local(v, ws, tmp_t, tmp_k, tmp_m, count, stream))) return rc;
                }
                if ((rc = read_counters(ws, hc, stream))) return rc;
            }
            const bool changed = hc[3] != 0 || hc[5] != 0 || hc[6] != 0;
            const bool was_full = full;
            nrel = (int64_t)hc[5];
            if (!changed) {
                if (was_full) { ++pass; break; }  // a full iteration without any change: this bucket is final
                full = confirmed = true;           // confirm the incremental result with one full iteration
                continue;
            }
            if (confirmed && was_full && getenv("MB_RELOC_DEBUG") != nullptr)
                fprintf(stderr, "[mb reloc] the confirming full iteration %d changed the table (changed=%llu released=%llu)\n",
                        pass, hc[3], hc[5]);
            confirmed = false;
            full = hc[6] != 0 || nrel > kRelCap;   // the release set overflowed, or a fallback-state mover found a probe open
        }
    }
    if ((rc = mb_reloc_apply(&d, 0, nmovers, agent_cell, workspace, workspace_bytes, max_movers, stream))) return rc;
    if ((rc = mb_reloc_apply(&d, 1, nmovers, agent_cell, workspace, workspace_bytes, max_movers, stream))) return rc;
    if (out_iterations) *out_iterations = pass;
    return MB_OK;
}
