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
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit sync;
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

#ifdef __cplusplus
}
#endif
#endif /* MESA_B200_H */
