/* CPU oracle for the mesa_b200 hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of what the reference (mesa/mesa 4.0.0a0, pure Python) computes on
 * the path named in BASELINE.json.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file; mesa_b200/ never does.
 *
 * Parity pin: every function here is checked against the live reference (imported from
 * /root/reference in the build container by oracle/make_golden.py) through the committed
 * fixtures under tests/golden/, see tests/test_oracle_golden.py.
 *
 * File:line citations are relative to the reference root.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ tiny parallel-for
 * (libgomp is not in the image).  Row-parallel loops of the synchronous kernels only;
 * every sequential-semantics function below is single-threaded like the reference. */
static int g_threads = 1;
ORC_API void orc_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
ORC_API int orc_get_threads(void) { return g_threads; }

typedef void (*range_fn)(int64_t lo, int64_t hi, void* arg);
typedef struct { range_fn fn; int64_t lo, hi; void* arg; } range_job;
static void* range_tramp(void* p) { range_job* j = (range_job*)p; j->fn(j->lo, j->hi, j->arg); return NULL; }
static void parallel_for(int64_t n, range_fn fn, void* arg) {
    int t = g_threads;
    if (t > n) t = (int)(n > 0 ? n : 1);
    if (t <= 1) { fn(0, n, arg); return; }
    pthread_t th[256]; range_job jobs[256];
    for (int k = 0; k < t; ++k) {
        jobs[k].fn = fn; jobs[k].arg = arg;
        jobs[k].lo = n * k / t; jobs[k].hi = n * (k + 1) / t;
        pthread_create(&th[k], NULL, range_tramp, &jobs[k]);
    }
    for (int k = 0; k < t; ++k) pthread_join(th[k], NULL);
}

/* ------------------------------------------------------------------ Philox4x32-10
 * Counter-based generator (Salmon et al., SC'11; Random123 v1.14 philox.h).  The
 * reference draws from CPython's MT19937 (mesa/model.py:129); the build replaces that
 * stream by Philox keyed on (seed, step, agent, draw#) and REPLAYS it into the unmodified
 * reference through oracle/replay.py, so both sides consume identical numbers. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

/* ------------------------------------------------------------------ Game of Life
 * examples/basic/conways_game_of_life/agents.py:33-55 (determine_state / assume_state),
 * model.py:30-36 (two-phase synchronous step), grid.py:390-399 (torus wrap `% dim`,
 * clamped edges drop the neighbour).  Requires dims >= 3 on a torus (smaller tori dedup
 * wrapped neighbours, cell.py:239-241 -- not a stencil; the device path rejects them too). */
typedef struct { const uint8_t* in; uint8_t* out; int64_t rows, cols; int torus; } gol_args;
static void gol_rows(int64_t lo, int64_t hi, void* p) {
    const gol_args* a = (const gol_args*)p;
    const uint8_t* in = a->in; uint8_t* out = a->out;
    const int64_t rows = a->rows, cols = a->cols; const int torus = a->torus;
    for (int64_t i = lo; i < hi; ++i) {
        for (int64_t j = 0; j < cols; ++j) {
            int live = 0;
            for (int di = -1; di <= 1; ++di) {
                for (int dj = -1; dj <= 1; ++dj) {
                    if (di == 0 && dj == 0) continue;
                    int64_t ni = i + di, nj = j + dj;
                    if (torus) {
                        ni = (ni + rows) % rows;
                        nj = (nj + cols) % cols;
                    } else if (ni < 0 || ni >= rows || nj < 0 || nj >= cols) {
                        continue;
                    }
                    live += in[ni * cols + nj] != 0;
                }
            }
            uint8_t next = in[i * cols + j];
            if (next) {
                if (live < 2 || live > 3) next = 0;
            } else {
                if (live == 3) next = 1;
            }
            out[i * cols + j] = next;
        }
    }
}
ORC_API int orc_gol_step_u8(const uint8_t* in, uint8_t* out, int64_t rows, int64_t cols, int torus) {
    if (torus && (rows < 3 || cols < 3)) return -1;
    gol_args a = {in, out, rows, cols, torus};
    parallel_for(rows, gol_rows, &a);
    return 0;
}
