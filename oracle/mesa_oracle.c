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

/* ------------------------------------------------------------------ keyed draw protocol
 * Same layout as oracle/philox.py and mesa_b200/csrc/philox.cuh. */
enum { STREAM_PRIORITY = 0, STREAM_DRAW = 1, STREAM_INIT = 2 };

static void philox_keyed(uint64_t seed, uint32_t epoch, uint64_t agent, uint32_t stream, uint32_t block,
                         uint32_t out[4]) {
    const uint32_t ctr[4] = {block, (uint32_t)agent, (uint32_t)((agent >> 32) & 0xffffu) | (stream << 16), epoch};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

ORC_API uint64_t orc_priority64(uint64_t seed, uint32_t epoch, uint64_t agent) {
    uint32_t o[4];
    philox_keyed(seed, epoch, agent, STREAM_PRIORITY, 0, o);
    return ((uint64_t)o[0] << 32) | (uint32_t)agent;
}

typedef struct {
    uint64_t seed, agent;
    uint32_t epoch, stream, c;
    uint32_t cache[4];
} draw_stream;

static void ds_init(draw_stream* s, uint64_t seed, uint32_t epoch, uint64_t agent, uint32_t stream) {
    s->seed = seed; s->agent = agent; s->epoch = epoch; s->stream = stream; s->c = 0;
}
static uint64_t ds_next64(draw_stream* s) {
    uint64_t v;
    if ((s->c & 1u) == 0) {
        philox_keyed(s->seed, s->epoch, s->agent, s->stream, s->c >> 1, s->cache);
        v = ((uint64_t)s->cache[0] << 32) | s->cache[1];
    } else {
        v = ((uint64_t)s->cache[2] << 32) | s->cache[3];
    }
    s->c++;
    return v;
}
/* CPython 3.12 Lib/random.py:242-250 _randbelow_with_getrandbits */
static uint64_t ds_randbelow(draw_stream* s, uint64_t n) {
    int k = 0;
    for (uint64_t t = n; t; t >>= 1) ++k; /* n.bit_length() */
    uint64_t r = ds_next64(s) >> (64 - k);
    while (r >= n) r = ds_next64(s) >> (64 - k);
    return r;
}

typedef struct { uint64_t pri; int64_t idx; } pri_item;
static int pri_cmp(const void* a, const void* b) {
    const uint64_t x = ((const pri_item*)a)->pri, y = ((const pri_item*)b)->pri;
    return (x > y) - (x < y);
}
/* activation order of agents 0..n-1 for shuffle number `epoch` (ascending priority64);
 * replaces random.shuffle in mesa/agentset.py:772-787. */
ORC_API int orc_activation_order(uint64_t seed, uint32_t epoch, int64_t n, int64_t* order) {
    pri_item* it = (pri_item*)malloc(sizeof(pri_item) * (size_t)(n > 0 ? n : 1));
    if (!it) return -1;
    for (int64_t a = 0; a < n; ++a) { it[a].pri = orc_priority64(seed, epoch, (uint64_t)a); it[a].idx = a; }
    qsort(it, (size_t)n, sizeof(pri_item), pri_cmp);
    for (int64_t a = 0; a < n; ++a) order[a] = it[a].idx;
    free(it);
    return 0;
}

/* ------------------------------------------------------------------ Schelling
 * State mirrors the reference's objects: per agent (registration order = index) its flat
 * cell id, type and happy flag; per cell the occupying agent index or -1 (capacity 1,
 * examples/basic/schelling/model.py:45-47). */

/* examples/basic/schelling/agents.py:24-42 assign_state for every agent
 * (AgentSet.do, agentset.py:756-770); neighbourhood = cell.py:225-276 Moore radius r
 * (Chebyshev square minus centre; clamped edges drop cells, torus wraps, grid.py:390-399).
 * Returns model.happy. */
ORC_API int64_t orc_schelling_assign(int64_t rows, int64_t cols, int radius, int torus, double homophily,
                                     int64_t n, const int64_t* cell, const int8_t* atype, uint8_t* happy,
                                     const int64_t* occ) {
    int64_t total = 0;
    for (int64_t a = 0; a < n; ++a) {
        const int64_t i = cell[a] / cols, j = cell[a] % cols;
        int64_t similar = 0, valid = 0;
        for (int di = -radius; di <= radius; ++di) {
            for (int dj = -radius; dj <= radius; ++dj) {
                if (di == 0 && dj == 0) continue;
                int64_t ni = i + di, nj = j + dj;
                if (torus) {
                    ni = ((ni % rows) + rows) % rows;
                    nj = ((nj % cols) + cols) % cols;
                } else if (ni < 0 || ni >= rows || nj < 0 || nj >= cols) {
                    continue;
                }
                const int64_t b = occ[ni * cols + nj];
                if (b < 0) continue;
                ++valid;
                similar += atype[b] == atype[a];
            }
        }
        double frac = 0.0;
        if (valid > 0) frac = (double)similar / (double)valid;
        if (frac < homophily) {
            happy[a] = 0;
        } else {
            happy[a] = 1;
            ++total;
        }
    }
    return total;
}

/* AgentSet.shuffle_do("step") (agentset.py:772-787) over agents.py:44-47:
 * unhappy agents, in activation order, call Grid.select_random_empty_cell (grid.py:288-316:
 * <= 50 probes random.choice(_celllist), first empty wins; else argwhere(empty) + choice)
 * and move (cell_agent.py:39-54).  Returns the number of movers, or -1 on a full grid
 * (ValueError in the reference).  *nfallback counts argwhere fallbacks. */
ORC_API int64_t orc_schelling_move(int64_t ncells, int64_t n, int64_t* cell, const uint8_t* happy, int64_t* occ,
                                   uint64_t seed, uint32_t epoch, int64_t* nfallback) {
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (!order || orc_activation_order(seed, epoch, n, order) != 0) { free(order); return -2; }
    int64_t movers = 0, fb = 0;
    for (int64_t q = 0; q < n; ++q) {
        const int64_t a = order[q];
        if (happy[a]) continue;
        draw_stream ds;
        ds_init(&ds, seed, epoch, (uint64_t)a, STREAM_DRAW);
        int64_t dest = -1;
        for (int p = 0; p < 50; ++p) {
            const int64_t c = (int64_t)ds_randbelow(&ds, (uint64_t)ncells);
            if (occ[c] < 0) { dest = c; break; }
        }
        if (dest < 0) {
            int64_t nempty = 0;
            for (int64_t c = 0; c < ncells; ++c) nempty += occ[c] < 0;
            if (nempty == 0) { free(order); return -1; }
            int64_t k = (int64_t)ds_randbelow(&ds, (uint64_t)nempty);
            for (int64_t c = 0; c < ncells; ++c)
                if (occ[c] < 0 && k-- == 0) { dest = c; break; }
            ++fb;
        }
        occ[dest] = a;          /* add to the new cell first ... */
        occ[cell[a]] = -1;      /* ... then leave the old one */
        cell[a] = dest;
        ++movers;
    }
    free(order);
    if (nfallback) *nfallback = fb;
    return movers;
}

/* ------------------------------------------------------------------ continuous space / Boids
 * experimental/continuous_space/continuous_space.py:202-235 calculate_distances (torus branch),
 * :237-247 get_agents_in_radius (dist <= radius, inclusive), continuous_space_agents.py:66-78
 * (drop self), continuous_space.py:169-200 calculate_difference_vector,
 * examples/basic/boid_flockers/agents.py:63-92 Boid.step, position setter
 * continuous_space_agents.py:36-44 + continuous_space.py:285-298. */
static double boid_dist(const double* p, const double* q, const double* size, int torus) {
    double dx = fabs(p[0] - q[0]), dy = fabs(p[1] - q[1]);
    if (torus) {
        dx = fmin(dx, size[0] - dx);
        dy = fmin(dy, size[1] - dy);
    }
    const double a = dx * dx, b = dy * dy;
    return sqrt(a + b);
}
static double boid_delta(double q, double p, double size, int torus) {
    const double d = q - p;
    if (!torus) return d;
    const double sgn = d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0);
    const double inv = d - sgn * size;
    return fabs(d) < fabs(inv) ? d : inv;
}
static double np_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0 && ((b < 0) != (m < 0))) m += b;
    return m;
}

typedef struct {
    int64_t n; const double *pos, *dir; double *pos_out, *dir_out;
    double lo[2], size[2]; int torus;
    double vision, separation, cohere, separate, match, speed;
    int64_t* nbr_count;
} boid_args;

/* one Boid.step for agent a reading (pos, dir), writing (po, dout) */
static void boid_one(const boid_args* g, int64_t a, const double* pos, const double* dir, double* po, double* dout) {
    const double* p = pos + 2 * a;
    double coh[2] = {0, 0}, sep[2] = {0, 0}, mat[2] = {0, 0};
    int64_t k = 0;
    for (int64_t b = 0; b < g->n; ++b) {  /* storage-index order, like the reference */
        if (b == a) continue;
        const double d = boid_dist(p, pos + 2 * b, g->size, g->torus);
        if (d <= g->vision) {
            const double dx = boid_delta(pos[2 * b], p[0], g->size[0], g->torus);
            const double dy = boid_delta(pos[2 * b + 1], p[1], g->size[1], g->torus);
            coh[0] += dx; coh[1] += dy;
            if (d < g->separation) { sep[0] += dx; sep[1] += dy; }
            mat[0] += dir[2 * b]; mat[1] += dir[2 * b + 1];
            ++k;
        }
    }
    double d0 = dir[2 * a], d1 = dir[2 * a + 1];
    if (k > 0) {
        d0 += (coh[0] * g->cohere + (-1 * sep[0] * g->separate) + mat[0] * g->match) / (double)k;
        d1 += (coh[1] * g->cohere + (-1 * sep[1] * g->separate) + mat[1] * g->match) / (double)k;
        const double q0 = d0 * d0, q1 = d1 * d1;
        const double norm = sqrt(q0 + q1);
        d0 /= norm; d1 /= norm;
    }
    double x = p[0] + d0 * g->speed, y = p[1] + d1 * g->speed;
    const int inb = x >= g->lo[0] && x <= g->lo[0] + g->size[0] && y >= g->lo[1] && y <= g->lo[1] + g->size[1];
    if (!inb && g->torus) {
        x = g->lo[0] + np_mod(x - g->lo[0], g->size[0]);
        y = g->lo[1] + np_mod(y - g->lo[1], g->size[1]);
    }
    po[0] = x; po[1] = y; dout[0] = d0; dout[1] = d1;
    if (g->nbr_count) g->nbr_count[a] = k;
}

static void boid_range(int64_t lo, int64_t hi, void* p) {
    const boid_args* g = (const boid_args*)p;
    for (int64_t a = lo; a < hi; ++a) boid_one(g, a, g->pos, g->dir, g->pos_out + 2 * a, g->dir_out + 2 * a);
}

/* order == NULL: synchronous (every boid reads the pre-step snapshot; what the device computes).
 * order != NULL: sequential Gauss-Seidel in that activation order, in place semantics
 *                (AgentSet.shuffle_do, agentset.py:772-787) -- the reference's own mode. */
ORC_API int orc_boids_step(int64_t n, const double* pos, const double* dir, double* pos_out, double* dir_out,
                           double x0, double y0, double size_x, double size_y, int torus, double vision,
                           double separation, double cohere, double separate, double match, double speed,
                           const int64_t* order, int64_t* nbr_count) {
    boid_args g = {n, pos, dir, pos_out, dir_out, {x0, y0}, {size_x, size_y}, torus,
                   vision, separation, cohere, separate, match, speed, nbr_count};
    if (order == NULL) {
        parallel_for(n, boid_range, &g);
        return 0;
    }
    memcpy(pos_out, pos, sizeof(double) * 2 * (size_t)n);
    memcpy(dir_out, dir, sizeof(double) * 2 * (size_t)n);
    for (int64_t q = 0; q < n; ++q) {
        const int64_t a = order[q];
        double np_[2], nd[2];
        boid_one(&g, a, pos_out, dir_out, np_, nd);
        pos_out[2 * a] = np_[0]; pos_out[2 * a + 1] = np_[1];
        dir_out[2 * a] = nd[0]; dir_out[2 * a + 1] = nd[1];
    }
    return 0;
}

/* get_agents_in_radius / get_neighbors_in_radius for one point: indices (ascending) and
 * distances of agents with dist <= radius, skipping `exclude` (-1: none).  Returns the count. */
ORC_API int64_t orc_radius_query(int64_t n, const double* pos, double px, double py, double size_x, double size_y,
                                 int torus, double radius, int64_t exclude, int64_t* out_idx, double* out_dist) {
    const double p[2] = {px, py}, size[2] = {size_x, size_y};
    int64_t k = 0;
    for (int64_t b = 0; b < n; ++b) {
        if (b == exclude) continue;
        const double d = boid_dist(p, pos + 2 * b, size, torus);
        if (d <= radius) {
            if (out_idx) out_idx[k] = b;
            if (out_dist) out_dist[k] = d;
            ++k;
        }
    }
    return k;
}

/* ------------------------------------------------------------------ Boltzmann wealth
 * examples/basic/boltzmann_wealth_model/agents.py:24-44 (move / give_money / step) under
 * AgentSet.shuffle_do (agentset.py:772-787).  Each cell keeps its agents in a doubly linked
 * list in list order (cell.py:143-170: append on arrival, list.remove on departure).
 * State: cell[n], wealth[n], and the list order: `next`/`prev` per agent, `head`/`tail` per cell
 * (all -1 terminated), owned by the caller so it persists between steps. */
ORC_API void orc_boltzmann_init_lists(int64_t ncells, int64_t n, const int64_t* cell, int64_t* next, int64_t* prev,
                                      int64_t* head, int64_t* tail) {
    for (int64_t c = 0; c < ncells; ++c) head[c] = tail[c] = -1;
    for (int64_t a = 0; a < n; ++a) { /* registration order = initial list order (model.py:70-74) */
        const int64_t c = cell[a];
        next[a] = -1; prev[a] = tail[c];
        if (tail[c] >= 0) next[tail[c]] = a; else head[c] = a;
        tail[c] = a;
    }
}

ORC_API int orc_boltzmann_step(int64_t rows, int64_t cols, int torus, int64_t n, int64_t* cell, int32_t* wealth,
                               int64_t* next, int64_t* prev, int64_t* head, int64_t* tail, uint64_t seed,
                               uint32_t epoch) {
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (!order || orc_activation_order(seed, epoch, n, order) != 0) { free(order); return -2; }
    for (int64_t q = 0; q < n; ++q) {
        const int64_t a = order[q];
        draw_stream ds;
        ds_init(&ds, seed, epoch, (uint64_t)a, STREAM_DRAW);
        /* move(): neighbourhood in connection order (grid.py:447-451), clamped or torus */
        const int64_t i = cell[a] / cols, j = cell[a] % cols;
        int64_t nbr[8]; int cnt = 0;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj) {
                if (di == 0 && dj == 0) continue;
                int64_t ni = i + di, nj = j + dj;
                if (torus) { ni = (ni + rows) % rows; nj = (nj + cols) % cols; }
                else if (ni < 0 || ni >= rows || nj < 0 || nj >= cols) continue;
                nbr[cnt++] = ni * cols + nj;
            }
        if (cnt == 0) { free(order); return -1; } /* LookupError in the reference */
        const int64_t dest = nbr[ds_randbelow(&ds, (uint64_t)cnt)];
        /* HasCell.cell setter (cell_agent.py:39-54): append to the new list, remove from the old */
        const int64_t old = cell[a];
        if (prev[a] >= 0) next[prev[a]] = next[a]; else head[old] = next[a];
        if (next[a] >= 0) prev[next[a]] = prev[a]; else tail[old] = prev[a];
        next[a] = -1; prev[a] = tail[dest];
        if (tail[dest] >= 0) next[tail[dest]] = a; else head[dest] = a;
        tail[dest] = a;
        cell[a] = dest;
        /* give_money() */
        if (wealth[a] > 0) {
            int64_t mates = 0;
            for (int64_t b = head[dest]; b >= 0; b = next[b]) mates += b != a;
            if (mates > 0) {
                int64_t m = (int64_t)ds_randbelow(&ds, (uint64_t)mates);
                int64_t other = -1;
                for (int64_t b = head[dest]; b >= 0; b = next[b]) {
                    if (b == a) continue;
                    if (m-- == 0) { other = b; break; }
                }
                wealth[other] += 1;
                wealth[a] -= 1;
            }
        }
    }
    free(order);
    return 0;
}
