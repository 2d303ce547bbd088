/* mesa_b200 -- C ABI of the B200-native execution backend for Mesa's data-parallel hot path.
 *
 * The reference (mesa/mesa 4.0.0a0) is pure Python and has NO plugin / FFI interface
 * (SURVEY.md section 8b); its seams are Python classes.  This header is therefore the ABI a
 * maintainer binds with ctypes (INTEGRATION.md shows the stub) from inside those seams:
 *   AbstractAgentSet.do / shuffle_do      mesa/agentset.py:27-56, 330-349
 *   Grid / OrthogonalMooreGrid / ...      mesa/discrete_space/grid.py:84-128, 434-501
 *   Cell.get_neighborhood                 mesa/discrete_space/cell.py:202-276
 *   property layers                       mesa/discrete_space/grid.py:130-234
 *   ContinuousSpace radius queries        mesa/experimental/continuous_space/continuous_space.py:169-298
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross the boundary;
 *   - every function returns 0 (MB_OK) or a negative MB_ERR_* code; the message of the last
 *     failure on the calling thread is mb_last_error() (valid until the next call);
 *   - `const T*` / `T*` buffer arguments are DEVICE pointers unless the name ends in `_host`;
 *   - the library never allocates, frees or retains device memory and never calls NCCL;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns without waiting for it.  The
 *     exceptions say "synchronises" in their comment -- they hand a data-dependent size or flag back to the host:
 *     mb_schelling_relocate and the mb_reloc_counters / mb_reloc_bucket_lists phases (mover counts, per-pass flags of
 *     the fixed point on large grids), mb_boids_step_seq_f32 (reads back its error word: a boid that moved farther
 *     than `reach` invalidates the schedule), mb_boltzmann_status.  The single-launch step of small Schelling grids
 *     (mb_schelling_step_small), the Boltzmann step, every stencil / layer / query kernel never synchronise;
 *   - grids are C-order [rows, cols] = [d0, d1] of Grid((d0, d1)) (grid.py:74-82, :146), so the
 *     flat cell id is i*cols + j = index into Grid._celllist (grid.py:120-126).
 */
#ifndef MESA_B200_H
#define MESA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_OK 0
#define MB_ERR_INVALID (-1)     /* bad argument -> ValueError in the Python shim */
#define MB_ERR_CUDA (-2)        /* CUDA runtime failure */
#define MB_ERR_WORKSPACE (-3)   /* workspace too small */
#define MB_ERR_UNSUPPORTED (-4) /* valid request the device path does not implement */

typedef void* mb_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------- */
int mb_version(void);
const char* mb_last_error(void);
int mb_sm_count(void);

/* ---- Game of Life (replaces examples/basic/conways_game_of_life/agents.py:33-55 driven by
 * AgentSet.do("determine_state") + do("assume_state"), model.py:30-36) ----------------------
 * One synchronous generation of a uint8 0/1 layer.  `halo_top` / `halo_bot` are the row above
 * row 0 and the row below row rows-1 ([cols] each; NULL = dead boundary): a single-GPU torus
 * passes in+(rows-1)*cols and in; a row-sharded grid passes the rows received from its
 * neighbour ranks.  col_torus wraps the column direction (grid.py:396-397). */
int mb_gol_step_u8(const uint8_t* in, const uint8_t* halo_top, const uint8_t* halo_bot, uint8_t* out,
                   int64_t rows, int64_t cols, int col_torus, mb_stream_t stream);

/* Row-sharded variant with the neighbour hand-shake fused into the stencil kernel: halo_top / halo_bot are
 * PEER-MAPPED pointers to the neighbouring ranks' edge rows (read in place over NVLink); wait_top / wait_bot are
 * this rank's generation flags, signal_up / signal_down the peer-mapped addresses of the neighbours' flags
 * (NULL where halo_* is NULL), edge_done two zero-initialised local counters, error a local word that is set
 * if a neighbour never shows up, gen = 1, 2, 3, ... the generation being written.  One launch per generation:
 * no exchange, no barrier (csrc/gol.cu, PeerSync). */
int mb_gol_step_u8_fused(const uint8_t* in, const uint8_t* halo_top, const uint8_t* halo_bot, uint8_t* out,
                         int64_t rows, int64_t cols, int col_torus, unsigned* wait_top, unsigned* wait_bot,
                         unsigned* signal_up, unsigned* signal_down, unsigned* edge_done, unsigned* error,
                         unsigned gen, mb_stream_t stream);

/* Bit-packed Game of Life (same rule, same reference lines; csrc/gol_bits.cu).  While a model only steps, the
 * uint8 layer is held as one bit per cell: word w of row i = cells (i, 32w .. 32w+31), bit k = column 32w+k
 * (cols % 32 == 0).  mb_gol_pack_u8 / mb_gol_unpack_u8 convert a 0/1 uint8 layer to / from that form (what a
 * read of the property layer costs).  mb_gol_step_bits advances `generations` (1..8) synchronous generations in
 * ONE launch by chaining them through registers; halo_top / halo_bot are [generations, cols/32] packed rows
 * above row 0 (last one adjacent) / below row rows-1 (first one adjacent), NULL = dead boundary: a single-GPU
 * torus passes in + (rows-generations)*cols/32 and in (needs rows >= generations), a row-sharded grid the
 * neighbours' edge rows.  rows_per_segment <= 0 picks the default tiling. */
int mb_gol_pack_u8(const uint8_t* state, uint32_t* bits, int64_t rows, int64_t cols, mb_stream_t stream);
int mb_gol_unpack_u8(const uint32_t* bits, uint8_t* state, int64_t rows, int64_t cols, mb_stream_t stream);
int mb_gol_step_bits(const uint32_t* in, const uint32_t* halo_top, const uint32_t* halo_bot, uint32_t* out,
                     int64_t rows, int64_t cols, int col_torus, int generations, int rows_per_segment,
                     mb_stream_t stream);

/* ---- Schelling happiness (replaces AgentSet.do("assign_state") over
 * examples/basic/schelling/agents.py:24-42; neighbourhood = cell.py:225-276, Moore radius r) ----
 * type: int8 layer, -1 = empty cell, >= 0 = agent type (capacity-1 grid).  happy: uint8 layer out
 * (1 = happy agent; 0 = unhappy agent or empty cell).  halo_top / halo_bot: [radius, cols] rows
 * above / below the block (NULL = outside the grid).  min_similar_host: HOST table of
 * (2r+1)^2 entries, entry v = smallest `similar` with `not (similar / v < homophily)` in the
 * reference's float arithmetic (> v = never happy).  binary_types != 0 promises types in {0,1}
 * (enables the packed-byte kernel).  happy_count: device uint64, receives model.happy. */
int mb_schelling_happy_i8(const int8_t* type, const int8_t* halo_top, const int8_t* halo_bot, uint8_t* happy,
                          int64_t rows, int64_t cols, int radius, int col_torus,
                          const uint8_t* min_similar_host, int binary_types,
                          unsigned long long* happy_count, mb_stream_t stream);

/* ---- Schelling relocation (replaces AgentSet.shuffle_do("step"), agentset.py:772-787, over
 * agents.py:44-47 -> Grid.select_random_empty_cell grid.py:288-316 -> HasCell.cell setter
 * cell_agent.py:39-54).  Activation order = ascending Philox priority64(seed, epoch, agent);
 * probes = the agent's keyed _randbelow(ncells) draws; result identical to the sequential
 * reference driven with the same numbers (csrc/relocate.cu explains the fixed point).
 * Layers (row-major [rows, cols]): type int8 (-1 empty), happy uint8, agent_id uint32 (agent
 * index sitting in each occupied cell), slots = 16 bytes per cell (when the cell empties / who
 * arrives; scratch: rebuilt from the type / happy layers at the start of every step by the mover collection, followed --
 * when the buffer is mb_reloc_table_bytes() long -- by the bitmap of the agents that stay put).  A full grid with an unhappy agent returns MB_ERR_INVALID
 * (ValueError, grid.py:311-315).
 *
 * mb_schelling_relocate: the whole step on ONE GPU (synchronises `stream`); agent_cell is an
 * optional [nagents] uint32 inverse map (NULL to skip).
 *
 * Row-sharded grids use the phase-level entry points with a shard descriptor whose per-rank
 * base pointers are peer-mapped (symmetric memory): rank r owns the rows
 * block_bounds(rows_total, world, r) of every layer; kernels read / claim remote slots and
 * write arrivals over NVLink.  mesa_b200/sharded.py holds the (small) host protocol. */
#define MB_MAX_PEERS 16
typedef struct mb_shard {
    int32_t world, rank;
    int64_t rows_total, cols;
    void* slots[MB_MAX_PEERS];    /* 16 B per cell */
    void* type[MB_MAX_PEERS];     /* int8 */
    void* happy[MB_MAX_PEERS];    /* uint8 */
    void* agent_id[MB_MAX_PEERS]; /* uint32, global agent index */
    void* static_bits;            /* optional (NULL: off): uint32 [world][bits_stride], THIS rank's copy of every rank's
                                   * "cell holds an agent that stays put this step" bitmap (bit l of section r = local cell l
                                   * of rank r).  mb_reloc_collect writes this rank's section; the caller all-gathers the
                                   * sections before the first pass.  Probes of such cells cost one L2-resident bit instead
                                   * of a 16-byte random (NVLink) access */
    int64_t bits_stride;          /* words per section, >= ceil(cells of the largest block / 32) */
} mb_shard_t;

int mb_reloc_table_bytes(int64_t rows, int64_t cols, size_t* out); /* slots + bitmap of one GPU's table buffer */
int mb_schelling_relocate(int8_t* type, uint8_t* happy, uint32_t* agent_id, uint32_t* agent_cell, void* slots,
                          size_t slots_bytes, int64_t rows, int64_t cols, uint64_t seed, uint32_t epoch, void* workspace,
                          size_t workspace_bytes, int64_t max_movers, int64_t* out_movers, int* out_iterations,
                          mb_stream_t stream);

int mb_reloc_workspace_bytes(int64_t max_movers, size_t* out);
int mb_reloc_slots_init(const void* shard /* mb_shard_t* */, mb_stream_t stream);
/* movers = occupied && !happy of the local block; counters[0..1] = movers, occupied */
int mb_reloc_collect(const void* shard, uint64_t seed, uint32_t epoch, void* workspace, size_t workspace_bytes,
                     int64_t max_movers, mb_stream_t stream);
/* host7 = {movers, occupied, fallback movers listed, table changed since the last take / FULL pass, (unused), entries of
 * the release write set, a FULL pass is required}; synchronises */
int mb_reloc_counters(void* workspace, size_t workspace_bytes, int64_t max_movers, uint64_t* host7,
                      mb_stream_t stream);
/* one pass over the local movers whose priority lies in [t_lo, t_hi] (rank-ordered buckets, earliest first).
 * mode 0 = FULL: every mover re-evaluates from its first probe (clears the fallback list, the flags and the release set);
 * mode 1 = RESUME: a mover only checks that it still holds its pick and otherwise continues probing after it (claims only
 * lower `arr`, so between releases a pick can only move forward); flags accumulate. */
int mb_reloc_pass(const void* shard, int64_t nmovers, uint64_t seed, uint32_t epoch, uint64_t t_lo, uint64_t t_hi,
                  int mode, void* workspace, size_t workspace_bytes, int64_t max_movers, mb_stream_t stream);
/* Rank-ordered buckets as index lists: mb_reloc_bucket_lists groups the mover indices by bucket (top bits of the
 * priority; nbuckets a power of two, 2..256) inside the workspace and writes the per-bucket counts to a HOST array
 * (synchronises); mb_reloc_pass_list is mb_reloc_pass over one bucket's list [list_first, list_first + list_count). */
int mb_reloc_bucket_lists(int64_t nmovers, int nbuckets, uint64_t* counts_host, void* workspace, size_t workspace_bytes,
                          int64_t max_movers, mb_stream_t stream);
int mb_reloc_pass_list(const void* shard, int64_t list_first, int64_t list_count, uint64_t seed, uint32_t epoch, uint64_t t_lo,
                       uint64_t t_hi, int mode, void* workspace, size_t workspace_bytes, int64_t max_movers,
                       mb_stream_t stream);
/* release sets: a cell whose registered arrival leaves (a fallback pick that shifts, a mover that moves back) may be empty
 * again for later movers that skipped it.  mb_reloc_release_take starts an incremental iteration: it moves the first
 * `count` (<= 4096) recorded (cell, time of the leaver) pairs into caller buffers and clears the set and the changed flag;
 * mb_reloc_release_scan lets the local movers whose EARLIER probes contain one of `nrel` (<= 4096) released cells move
 * back to it if it is empty at their time (the probes are replayed from the Philox stream: no table traffic except for
 * hits).  Row-sharded grids scan the union of all ranks' sets.  Protocol per bucket: a FULL iteration (pass + fallback
 * phase), then incremental iterations (take, scan, RESUME pass, fallback phase) until one changes nothing, confirmed by a
 * FULL iteration (csrc/relocate.cu, mb_schelling_relocate). */
int mb_reloc_release_take(void* workspace, size_t workspace_bytes, int64_t max_movers, uint64_t* cells_out,
                          uint64_t* times_out, int64_t count, mb_stream_t stream);
int mb_reloc_release_scan(const void* shard, int64_t nmovers, uint64_t seed, uint32_t epoch, uint64_t t_lo, uint64_t t_hi,
                          const uint64_t* rel_cells, const uint64_t* rel_times, int64_t nrel, void* workspace,
                          size_t workspace_bytes, int64_t max_movers, mb_stream_t stream);
/* argwhere fallback (grid.py:308-316), batched: export (time, k = _randbelow(nempty), mover) of the local
 * fallback movers; count the empties of the local table at <= 2048 ascending times; pick the k_local-th one;
 * commit the picks of the local movers. */
int mb_reloc_fb_export(const void* shard, int64_t first, int count, int64_t nempty, uint64_t seed, uint32_t epoch,
                       uint64_t* out_t, uint64_t* out_k, uint32_t* out_m, void* workspace, size_t workspace_bytes,
                       int64_t max_movers, mb_stream_t stream);
int mb_reloc_fb_count(const void* shard, const uint64_t* times_sorted, int count, uint32_t* totals, void* workspace,
                      size_t workspace_bytes, int64_t max_movers, mb_stream_t stream);
int mb_reloc_fb_pick(const void* shard, const uint64_t* times_sorted, const uint64_t* k_local, int count,
                     uint64_t* result, void* workspace, size_t workspace_bytes, int64_t max_movers,
                     mb_stream_t stream);
int mb_reloc_fb_commit(const void* shard, const uint64_t* times_sorted, const uint32_t* movers,
                       const uint64_t* result, int count, void* workspace, size_t workspace_bytes,
                       int64_t max_movers, mb_stream_t stream);
/* The fallback cascade as a time-ordered EVENT LOOP (csrc/relocate.cu, reloc_event_loop): once the ordinary movers of a
 * bucket have converged without any fallback arrival (FULL pass + RESUME passes until nothing changes), the fallback
 * movers -- and the later movers their picks displace -- are processed in ascending time by ONE CTA: no iteration, no
 * release.  mb_reloc_fb_count_into fills a caller-owned count table [65536 + 1024][2048] uint32 (per block of the slot table, then per super-block of 64 blocks; block-major) (peer-mapped on a
 * row-sharded grid) and totals[count] for `count` ascending fallback times; mb_reloc_event_loop (one rank) takes the
 * `world` tables (host array of device pointers), totals [world][2048], the times and ranks k = _randbelow(nempty) of
 * the fallback movers of all ranks, and writes the picks that changed as overrides (time of the mover, global cell;
 * <= 4096) plus out4 = {overrides, bail reason (0 = completed), fallback events, displaced events};
 * mb_reloc_apply_overrides writes them into the local movers' choice.  A confirming FULL iteration follows. */
int mb_reloc_fb_count_into(const void* shard, const uint64_t* times_sorted, int count, uint32_t* totals,
                           uint32_t* cnt_table, mb_stream_t stream);
int mb_reloc_event_loop(const void* shard, void* const* cnt_tables_host, uint32_t* totals, const uint64_t* fb_times,
                        const uint64_t* fb_ranks, int count, uint64_t* ov_times, uint64_t* ov_cells, uint64_t* out4,
                        uint64_t seed, uint32_t epoch, mb_stream_t stream);
int mb_reloc_apply_overrides(int64_t nmovers, uint64_t t_lo, uint64_t t_hi, const uint64_t* ov_times,
                             const uint64_t* ov_cells, int count, void* workspace, size_t workspace_bytes,
                             int64_t max_movers, mb_stream_t stream);
/* phase 0: origins become empty; phase 1 (after ALL ranks did phase 0): arrivals are written */
int mb_reloc_apply(const void* shard, int phase, int64_t nmovers, uint32_t* agent_cell, void* workspace,
                   size_t workspace_bytes, int64_t max_movers, mb_stream_t stream);

/* ---- the WHOLE Schelling model step of a small grid in ONE launch (csrc/schelling_small.cu; schelling/model.py:87-93:
 * shuffle_do("step") then do("assign_state")).  Grids of at most mb_schelling_small_max_cells() (12800) cells keep the
 * relocation slots in shared memory; one CTA runs the same fixed-point iteration as mb_schelling_relocate between
 * barriers, applies the moves and evaluates every agent's happiness.  Nothing is read back: out4 = device uint64[4] =
 * {model.happy, movers, status (0 ok; 1 = grid completely full with an unhappy agent -> ValueError, grid.py:311-315;
 * 2 = workspace too small; 3 = too many fallback movers; 4 = no convergence), passes}.  relocate == 0 only evaluates
 * happiness (the model's initial do("assign_state")).  agent_cell [nagents] and empty_layer [rows, cols] (the grid's
 * implicit "empty" layer, grid.py:128) are optional.  min_similar_host as in mb_schelling_happy_i8.  nsteps > 1 (with
 * relocate != 0) runs that many model steps inside the one launch -- `model.run_for(n)`, mesa/model.py:395-402 -- with the
 * shuffle numbers epoch, epoch + 1, ...; out4 then describes the last step, a failed step ends the run.  happy_current != 0:
 * the caller guarantees that `happy` on entry is the evaluation of `type` (true inside a model run); a step without movers
 * then changes nothing and only reads the layers (a settled model steps at the cost of that read). */
int mb_schelling_small_max_cells(void);
int mb_schelling_small_workspace_bytes(int64_t max_movers, size_t* out);
int mb_schelling_step_small(int8_t* type, uint8_t* happy, uint32_t* agent_id, uint32_t* agent_cell, uint8_t* empty_layer,
                            int64_t rows, int64_t cols, int radius, int torus, const uint8_t* min_similar_host,
                            uint64_t seed, uint32_t epoch, int relocate, int nsteps, int happy_current, void* workspace,
                            size_t workspace_bytes, int64_t max_movers, uint64_t* out4, mb_stream_t stream);

/* ---- stable LSD radix sort of (key, uint32) pairs (cell lists; no reference counterpart: the
 * reference scans all agents per query, continuous_space.py:202-235) --------------------------
 * Sorts on key bits [begin_bit, end_bit), 8 bits per pass, ping-ponging between buffer 0 and 1;
 * *result_buffer tells which pair of buffers holds the result.  Equal keys keep input order. */
int mb_sort_workspace_bytes(int64_t n, size_t* out);
int mb_sort_pairs_u32(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, int64_t n,
                      int begin_bit, int end_bit, void* workspace, size_t workspace_bytes, int* result_buffer,
                      mb_stream_t stream);
int mb_sort_pairs_u64(uint64_t* keys0, uint32_t* vals0, uint64_t* keys1, uint32_t* vals1, int64_t n,
                      int begin_bit, int end_bit, void* workspace, size_t workspace_bytes, int* result_buffer,
                      mb_stream_t stream);

/* offsets[i] = counts[0] + ... + counts[i-1] (i = 0 .. n): the row offsets of a CSR result from its per-row counts (the
 * batched radius queries count first, then fill).  scratch: (n + 1) + mb_scan_scratch_words(n + 1) uint32 words.  The
 * running sum is 32 bits wide: totals >= 2^32 are the caller's to refuse. */
int mb_offsets_from_counts(const int32_t* counts, int64_t n, int64_t* offsets, uint32_t* scratch, mb_stream_t stream);

/* The activation order of shuffle / shuffle_do number `epoch` as an explicit permutation (agentset.py:415-437,
 * 772-787): order_out[r] = index of the agent activated r-th = agents 0..n-1 sorted by priority64(seed, epoch, agent). */
int mb_activation_order_workspace_bytes(int64_t n, size_t* out);
int mb_activation_order(uint64_t seed, uint32_t epoch, int64_t n, uint32_t* order_out, void* workspace,
                        size_t workspace_bytes, mb_stream_t stream);

/* Boids in the reference's SEQUENTIAL activation order (AgentSet.shuffle_do("step"), agentset.py:772-787, over
 * boid_flockers/agents.py:63-92): boid a sees the new position / direction of every boid activated before it and the
 * old state of the others.  Order = ascending Philox priority64(seed, epoch, agent) (the order oracle/replay.py feeds to
 * the unmodified reference).  Computed in dependency passes (csrc/continuous.cu, boids_seq_pass): identical to the
 * sequential loop, float64 inside the step, float32 state at the step boundary like mb_boids_step_f32.  `reach` bounds
 * the distance a boid moves in one step (speed * max(1, max |direction|)); a longer move returns MB_ERR_INVALID.
 * Synchronises `stream`; *out_passes receives the number of dependency passes. */
int mb_boids_seq_workspace_bytes(int64_t n, double size_x, double size_y, double vision, double reach, size_t* out);
int mb_boids_step_seq_f32(const float* pos_in, const float* dir_in, float* pos_out, float* dir_out, int64_t n,
                          double x0, double y0, double size_x, double size_y, int torus, double vision,
                          double separation, double cohere, double separate, double match, double speed,
                          double reach, uint64_t seed, uint32_t epoch, uint32_t* neighbor_count, void* workspace,
                          size_t workspace_bytes, int* out_passes, mb_stream_t stream);

/* ---- continuous space (replaces experimental/continuous_space/continuous_space.py:169-298 and
 * examples/basic/boid_flockers/agents.py:63-92 for all agents at once) -------------------------
 * Positions / directions are float32 [n, 2] in storage-index order (ContinuousSpace._agent_positions
 * rows, continuous_space.py:85-101).  The space is [x0, x0+size_x] x [y0, y0+size_y].
 * mb_boids_step_f32 is one SYNCHRONOUS Boid.step of every agent (activation modes);
 * neighbor_count (optional, [n]) receives len(neighbors).
 * mb_radius_query_f32 is a batched get_agents_in_radius (dist <= radius, inclusive): first call
 * with out_idx == NULL fills counts[nq]; the caller prefix-sums them into offsets[nq] and calls
 * again to fill out_idx / out_dist (float64, the reference's distances) in ascending agent index.
 * exclude[q] >= 0 drops that agent (ContinuousSpaceAgent.get_neighbors_in_radius). */
int mb_celllist_workspace_bytes(int64_t n, double size_x, double size_y, double radius, size_t* out);
int mb_boids_step_f32(const float* pos_in, const float* dir_in, float* pos_out, float* dir_out, int64_t n,
                      double x0, double y0, double size_x, double size_y, int torus, double vision,
                      double separation, double cohere, double separate, double match, double speed,
                      uint32_t* neighbor_count, void* workspace, size_t workspace_bytes, mb_stream_t stream);
int mb_radius_query_f32(const float* pos, int64_t n, double x0, double y0, double size_x, double size_y,
                        int torus, const float* queries, const int32_t* exclude, int64_t nq, double radius,
                        uint32_t* counts, const int64_t* offsets, int32_t* out_idx, double* out_dist,
                        void* workspace, size_t workspace_bytes, int reuse_cell_list, mb_stream_t stream);

/* mb_knn_query_f32: batched ContinuousSpace.get_k_nearest_agents (continuous_space.py:249-283) on the cell list:
 * out_idx / out_dist are [nq, k], each row ascending in (distance, agent index) -- "in case of ties, the earlier an
 * agent is inserted the higher it will rank" (:256-257) -- padded with -1 / +inf when fewer than k agents exist;
 * exclude[q] >= 0 drops that agent (ContinuousSpaceAgent.get_nearest_neighbors, continuous_space_agents.py:80-91).
 * cell_edge sizes the cells; workspace = mb_celllist_workspace_bytes(n, size_x, size_y, cell_edge).
 * mb_point_distances_f32: calculate_distances (:202-235; out_dist [m]) and / or calculate_difference_vector
 * (:169-200; out_delta [m, 2]) of one point against agents sel[0..m) (sel NULL: agents 0..m), float64, in the
 * reference's operation order. */
int mb_knn_query_f32(const float* pos, int64_t n, double x0, double y0, double size_x, double size_y, int torus,
                     const float* queries, const int32_t* exclude, int64_t nq, int k, double cell_edge,
                     int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes,
                     int reuse_cell_list, mb_stream_t stream);
int mb_point_distances_f32(const float* pos, const int32_t* sel, int64_t m, double size_x, double size_y, int torus,
                           double px, double py, double* out_dist, double* out_delta, mb_stream_t stream);

/* The same two queries in a space of any dimension (ContinuousSpace keeps a (capacity, ndims) array,
 * continuous_space.py:85-101; distances add the squares axis by axis, :223-234): pos [n, ndim] float32, size_host and
 * point_host [ndim] doubles in HOST memory, out_dist [m], out_delta [m, ndim] float64 (either may be NULL). */
int mb_point_distances_nd_f32(const float* pos, int ndim, const int32_t* sel, int64_t m, const double* size_host,
                              int torus, const double* point_host, double* out_dist, double* out_delta,
                              mb_stream_t stream);

/* Domain-decomposed continuous space (one rank per slab of y, mesa_b200/sharded.py ShardedBoids): the exchange
 * steps as kernels that append straight into the neighbouring ranks' peer-mapped inboxes (NVLink stores + an
 * atomic slot counter in the receiver's memory) -- no pack / send / recv / unpack.  *_count are peer-mapped
 * uint32 counters (NULL: no neighbour on that side), cap the inbox capacity, overflow / flags local words
 * (flags[0] inbox overflow, flags[1] a boid jumped over a whole slab).  mb_boids_route also produces the stay
 * flags and their exclusive scan for the stable compaction done by mb_boids_compact. */
int mb_boids_send_ghosts(const float* pos, const float* dir, int64_t n, double y_lo, double y_hi, double reach,
                         float* down_pos, float* down_dir, unsigned* down_count, float* up_pos, float* up_dir,
                         unsigned* up_count, int64_t cap, unsigned* overflow, mb_stream_t stream);
int mb_boids_route(const float* pos, const float* dir, const int64_t* ids, int64_t n, double height, int world,
                   int rank, unsigned* stay_flags, unsigned* stay_offsets, unsigned* scratch, float* down_pos,
                   float* down_dir, int64_t* down_ids, unsigned* down_count, float* up_pos, float* up_dir,
                   int64_t* up_ids, unsigned* up_count, int64_t cap, unsigned* flags, mb_stream_t stream);
int mb_boids_compact(const float* pos, const float* dir, const int64_t* ids, int64_t n, const unsigned* stay_flags,
                     const unsigned* stay_offsets, float* pos_out, float* dir_out, int64_t* ids_out,
                     mb_stream_t stream);
/* Rewrite a boid population in cell-list order (row i of the outputs = the boid at sorted slot i, perm_out[i] = its
 * previous row; ids_in NULL: the previous rows are the ids).  A model that calls this every few steps keeps its storage
 * order spatially sorted, which turns the gathers / scatters of mb_boids_step_f32 into coalesced accesses.  Same
 * workspace as the step (mb_celllist_workspace_bytes with radius = vision). */
int mb_boids_resort_f32(const float* pos_in, const float* dir_in, const int64_t* ids_in, float* pos_out, float* dir_out,
                        int64_t* ids_out, uint32_t* perm_out, int64_t n, double x0, double y0, double size_x,
                        double size_y, int torus, double vision, void* workspace, size_t workspace_bytes,
                        mb_stream_t stream);
int mb_scan_scratch_words(int64_t n, size_t* out);

/* ---- Dynamic populations (replaces the per-agent dict deletions of _HardKeyAgentSet.discard / remove,
 * mesa/agentset.py:700-711, reached through Agent.remove -> Model.deregister_agent, mesa/agent.py:73-108,
 * mesa/model.py:281-294): removing a batch of agents is one STABLE compaction of every structure-of-arrays
 * column -- survivors keep their registration order, the base order of do / shuffle_do (agentset.py:776).
 * mb_compact_plan: keep [n] bytes (0 = remove) -> offsets [n + 1] uint32, offsets[i] = survivors before row i,
 * offsets[n] = number of survivors; scratch = mb_scan_scratch_words(n + 1) uint32 words.
 * mb_compact_rows: out[offsets[i]] = in[i] for every kept row (row_bytes wide; in / out must not overlap);
 * called once per column with the same plan. */
int mb_compact_plan(const uint8_t* keep, int64_t n, uint32_t* offsets, uint32_t* scratch, mb_stream_t stream);
int mb_compact_rows(const void* in, void* out, int64_t n, int64_t row_bytes, const uint8_t* keep,
                    const uint32_t* offsets, mb_stream_t stream);

/* ---- Boltzmann wealth (replaces AgentSet.shuffle_do("step") over
 * examples/basic/boltzmann_wealth_model/agents.py:24-44: Moore random walk, then give one unit to
 * a random cell mate) ---------------------------------------------------------------------------
 * cell: [n] uint32 flat cell id per agent (agent index = registration order), wealth: [n] int32.
 * The workspace carries the per-cell list order (arrival order, cell.py:143-170) between steps:
 * call mb_boltzmann_init after placing agents, then mb_boltzmann_step once per model step with
 * epoch = number of shuffles so far.  Result identical to the sequential reference driven with the
 * same keyed draws (csrc/boltzmann.cu).  mb_boltzmann_step only ENQUEUES (no host synchronisation: crowded cells
 * and the convergence of the give decisions are resolved on the device); iterations_dev: optional device uint32. */
int mb_boltzmann_workspace_bytes(int64_t n, size_t* out);
int mb_boltzmann_init(const uint32_t* cell, int64_t n, int64_t rows, int64_t cols, void* workspace,
                      size_t workspace_bytes, mb_stream_t stream);
int mb_boltzmann_step(uint32_t* cell, int32_t* wealth, int64_t n, int64_t rows, int64_t cols, int torus,
                      uint64_t seed, uint32_t epoch, void* workspace, size_t workspace_bytes,
                      uint32_t* iterations_dev, mb_stream_t stream);
/* sticky status word of the workspace (0 = every step was exact; 1 = a step met more than 4096 cells holding more
 * than 32 agents each, which the device path does not order).  Synchronises `stream`. */
int mb_boltzmann_status(void* workspace, size_t workspace_bytes, int64_t n, uint64_t* status_host, mb_stream_t stream);

/* ---- property layers (replace the numpy expressions users run on Grid.property_layers arrays,
 * grid.py:112,130-220; e.g. sugarscape_g1mt/model.py:123-124, grid.py:146, grid.py:308) and the
 * generic neighbourhood sum (cell.py:225-276: Moore = Chebyshev square, von Neumann = Manhattan
 * diamond, centre excluded unless include_center) --------------------------------------------------
 * dtype: 0 uint8, 1 int32, 2 float32, 3 float64.  op (binary): 0 add 1 sub 2 mul 3 min 4 max.
 * op (reduce): 0 sum 1 min 2 max 3 count_nonzero; result is int64 for integer layers and
 * count_nonzero, float64 otherwise; reductions are reproducible (fixed two-pass order).
 * mb_neighborhood_sum writes int32 sums for integer layers, float64 for float layers.
 * mb_layer_diffuse: out = in*(1 - rate*k/8) + rate/8 * sum(existing Moore neighbours), k = their
 * number (NetLogo `diffuse`; the reference has no diffusion -- parity unpinned). */
int mb_layer_fill(void* data, int64_t n, int dtype, double value, mb_stream_t stream);
int mb_layer_affine_clamp(const void* in, void* out, int64_t n, int dtype, double alpha, double beta, double lo,
                          double hi, mb_stream_t stream);
int mb_layer_binary(const void* a, const void* b, void* out, int64_t n, int dtype, int op, mb_stream_t stream);
int mb_layer_reduce(const void* in, int64_t n, int dtype, int op, void* result, void* workspace,
                    size_t workspace_bytes, mb_stream_t stream);
int mb_neighborhood_sum(const void* in, const void* halo_top, const void* halo_bot, void* out, int64_t rows,
                        int64_t cols, int dtype, int radius, int von_neumann, int include_center, int col_torus,
                        mb_stream_t stream);
int mb_layer_diffuse(const void* in, const void* halo_top, const void* halo_bot, void* out, int64_t rows,
                     int64_t cols, int dtype, double rate, int col_torus, mb_stream_t stream);
/* Neighbourhood sums on an n-dimensional layer (1 <= ndim <= 4, C order; dims[ndim]): the n-d grids of
 * grid.py:457-462 (Moore: product([-1,0,1]^n) connections) and :488-501 (von Neumann: +-1 per axis), radius r =
 * Chebyshev cube / Manhattan ball (cell.py:225-276).  out dtype as mb_neighborhood_sum. */
int mb_neighborhood_sum_nd(const void* in, void* out, int ndim, const int64_t* dims_host, int dtype, int radius,
                           int von_neumann, int include_center, int torus, mb_stream_t stream);

/* ---- deep-ghost-zone refresh of a row-sharded layer without barrier or collective (csrc/ghost.cu; SURVEY 8e,
 * no reference counterpart: the reference is single-process).  flags = this rank's 8-word uint32 flag block in
 * peer-mapped memory (zero-initialised); up_flags / down_flags = peer-mapped addresses of the neighbours' blocks.
 * mb_ghost_exchange: publish "my plane is ready", wait for both neighbours, copy `bytes` (multiple of 16) from
 * src_up -> dst_top and src_down -> dst_bot over NVLink, publish "copied".  mb_ghost_wait_copied: device-side wait
 * until the neighbours have copied out of this rank's plane (before the launch that overwrites it).  A wait that
 * exceeds ~seconds sets flags[6] instead of hanging. */
int mb_ghost_exchange(void* dst_top, const void* src_up, void* dst_bot, const void* src_down, int64_t bytes,
                      unsigned* flags, unsigned* up_flags, unsigned* down_flags, mb_stream_t stream);
int mb_ghost_wait_copied(unsigned* flags, int has_up, int has_down, mb_stream_t stream);

/* ---- Wolf-Sheep predation (replaces agents_by_type[Sheep].shuffle_do("step") / agents_by_type[Wolf].shuffle_do("step"),
 * mesa/examples/advanced/wolf_sheep/model.py:136-141 over agents.py:36-158; csrc/wolf_sheep.cu).  Sequential semantics as
 * the fixed point of Jacobi passes: the caller repeats *_pass + mb_ws_compare_reset until the table no longer changes,
 * then calls *_apply.  cell: int32 flat cell (row-major [height, width], von Neumann torus), energy: float64, uid: the
 * agents' unique ids (the Philox key is uid - 1), tables: uint64 priorities, all bits set = never.  Sheep: wolves_in
 * [ncells] int32 (mb_ws_count_cells of the wolves), grown [ncells] uint8 or NULL (no grass).  Wolves: slist = sheep rows
 * sorted by (cell, position in the cell's list), cstart [ncells + 1] (mb_ws_segment_starts). */
int mb_ws_count_cells(const int32_t* cell, int64_t n, int32_t* counts, int64_t ncells, mb_stream_t stream);
int mb_ws_segment_starts(const int32_t* sorted_cell, int64_t n, int64_t ncells, int32_t* cstart, mb_stream_t stream);
int mb_ws_sheep_pass(const int32_t* cell, const int64_t* uid, int64_t n, int height, int width, uint64_t seed, uint32_t epoch,
                     const int32_t* wolves_in, const uint8_t* grown, const uint64_t* E_prev, uint64_t* E_next,
                     mb_stream_t stream);
int mb_ws_compare_reset(uint64_t* a, const uint64_t* b, int64_t n, unsigned* changed, mb_stream_t stream);
int mb_ws_sheep_apply(int32_t* cell, double* energy, const int64_t* uid, int64_t n, int height, int width, uint64_t seed,
                      uint32_t epoch, const int32_t* wolves_in, const uint8_t* grown, uint8_t* grown_out, int32_t* regrow_at,
                      const uint64_t* E, double gain, double p_reproduce, int32_t regrow_time, uint8_t* moved, uint8_t* dead,
                      uint8_t* repro, uint64_t* prio, mb_stream_t stream);
int mb_ws_wolf_pass(const int32_t* cell, const int64_t* uid, int64_t n, int height, int width, uint64_t seed, uint32_t epoch,
                    const int32_t* slist, const int32_t* cstart, const uint64_t* D_prev, uint64_t* D_next, mb_stream_t stream);
int mb_ws_wolf_apply(int32_t* cell, double* energy, const int64_t* uid, int64_t n, int height, int width, uint64_t seed,
                     uint32_t epoch, const int32_t* slist, const int32_t* cstart, const uint64_t* D, double gain,
                     double p_reproduce, uint8_t* dead, uint8_t* repro, uint64_t* prio, mb_stream_t stream);
int mb_ws_regrow(uint8_t* grown, int32_t* regrow_at, int64_t ncells, int32_t now, mb_stream_t stream);

/* ---- Epstein civil violence (csrc/epstein.cu) -----------------------------------------------------------------------
 * replaces, for one model step, mesa/examples/advanced/epstein_civil_violence/model.py:100-117
 * `self.agents.shuffle_do("step")` over citizens and cops with agents.py:14-26 (update_neighbors / move: radius-`vision`
 * neighbourhood in BFS order, is_empty filtering), agents.py:66-103 (Citizen.step) and agents.py:124-140 (Cop.step), in the
 * reference's sequential semantics (rank-ordered dependency graph, one launch).
 *   occ[width * height] int32: -1 = empty cell, else agent << 2 | cop << 1 | active (flat cell = x * height + y);
 *   kind 0 citizen / 1 cop; state = CitizenState value (1 ACTIVE, 2 QUIET, 3 ARRESTED; cops 0);
 *   voff [nv][2] int16: neighbourhood offsets (dx, dy) in the order of Cell._neighborhood (cell.py:240-276);
 *   woff [nw][2] int16: every offset within 2 * vision (dependency range); ptab [nv + 1]: 1 - exp(-k * m), m = 0 .. nv;
 *   sequential != 0: registration order (`agents.do("step")`, activation_order "Sequential") instead of the Philox order;
 *   counts_dev (device, 4 x uint32, optional): status (0 = valid), ACTIVE, QUIET, ARRESTED citizens after the step.
 * Both dimensions must be >= 2 * vision + 1 (a torus on which the neighbourhood does not wrap onto itself; the caller checks). */
int mb_epstein_workspace_bytes(int64_t n, int64_t ncells, size_t* out);
int mb_epstein_build_occ(const int32_t* cell, const uint8_t* kind, const uint8_t* state, int64_t n, int32_t* occ,
                         int64_t ncells, mb_stream_t stream);
int mb_epstein_step(int32_t* occ, int32_t* cell, const uint8_t* kind, uint8_t* state, int32_t* jail,
                    const double* risk_aversion, const double* grievance, double* arrest_probability, int64_t n,
                    int64_t width, int64_t height, const int16_t* voff, int nv, const int16_t* woff, int nw,
                    const double* ptab, int max_jail_term, double threshold, int movement, int sequential, uint64_t seed,
                    uint32_t epoch, uint32_t* counts_dev, void* workspace, size_t workspace_bytes, mb_stream_t stream);

/* ---- Sugarscape G1MT (csrc/sugarscape.cu): movement phase --------------------------------------------------------------
 * replaces, for one model step without trade, mesa/examples/advanced/sugarscape_g1mt/model.py:118-127 (regrowth of both
 * layers, `self.agents.shuffle_do("step")`) with agents.py:209-262 (move: empty cells of the radius-`vision` von Neumann
 * neighbourhood in BFS order, Cobb-Douglas welfare, isclose to the best, closest, random.choice), :264-271 (eat), :273-288
 * (starve -> remove), in the reference's sequential semantics (rank-ordered dependency graph, one launch).
 *   flat cell = x * height + y, clamped grid; layers float64 [width * height]; count[cell] = traders in the cell;
 *   voff [2 V (V + 1)][2] int16: the neighbourhood of the LARGEST vision V in BFS order, centre excluded (vision v = its first
 *   2 v (v + 1) entries); woff [nw][2] int16: every offset within 2 V (dependency range); keys: unique_id - 1;
 *   dead [n] uint8 (out): starved in this step -- the caller removes them (mb_compact_plan / mb_compact_rows);
 *   pow_table [pow_m][pow_m][pow_x] float64 (optional): x ** (ms / (ms + mp)) at [ms - 1][mp - 1][x] as the HOST's libm
 *   computes it -- the welfare sugar ** (ms / mt) * spice ** (mp / mt) (agents.py:57-70) then is the reference's bit for bit
 *   (amounts are integer-valued); amounts or metabolisms beyond the table use CUDA's pow and are counted;
 *   status_dev (device, 3 x uint32, optional): status (0 valid, 1 spin limit, 2 more than 4 traders started in one cell),
 *   deaths, pow calls that missed the table. */
int mb_sugarscape_workspace_bytes(int64_t n, int64_t ncells, size_t* out);
int mb_sugarscape_build_count(const int32_t* cell, int64_t n, int32_t* count, int64_t ncells, mb_stream_t stream);
int mb_sugarscape_step(int32_t* cell, double* sugar, double* spice, const int32_t* metabolism_sugar,
                       const int32_t* metabolism_spice, const int32_t* vision, const int64_t* unique_id, uint8_t* dead, int64_t n,
                       double* sugar_layer, const double* sugar_max, double* spice_layer, const double* spice_max,
                       int32_t* count, int64_t width, int64_t height, const int16_t* voff, int max_vision, const int16_t* woff,
                       int nw, int regrow, const double* pow_table, int pow_m, int pow_x, uint64_t seed, uint32_t epoch,
                       uint32_t* status_dev, void* workspace, size_t workspace_bytes, mb_stream_t stream);
/* the trade activation of a step (model.py:128-129 `shuffle_do("trade_with_neighbors")`, agents.py:290-301, 156-207, 94-153):
 * every executed trade is appended to (rec_trader = ROW of the trader whose turn it was, rec_partner = unique id of the other
 * side, rec_price), a trader's trades in the order it made them; status_dev = {status, number of trades (may exceed
 * rec_cap), pow calls that missed pow_table}. */
int mb_sugarscape_trade(const int32_t* cell, double* sugar, double* spice, const int32_t* metabolism_sugar,
                        const int32_t* metabolism_spice, const int32_t* vision, const int64_t* unique_id, int64_t n,
                        int64_t width, int64_t height, const int16_t* voff, int max_vision, const int16_t* woff, int nw,
                        const double* pow_table, int pow_m, int pow_x, uint64_t seed, uint32_t epoch, uint32_t* rec_trader,
                        int64_t* rec_partner, double* rec_price, int64_t rec_cap, uint32_t* status_dev, void* workspace,
                        size_t workspace_bytes, mb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MESA_B200_H */
