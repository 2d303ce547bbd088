// Wolf-Sheep predation: typed activation with predation on dynamic populations.
//
// Replaces, for all sheep / all wolves at once,
//   mesa/examples/advanced/wolf_sheep/model.py:136-141    agents_by_type[Sheep].shuffle_do("step"); ...[Wolf].shuffle_do("step")
//   mesa/examples/advanced/wolf_sheep/agents.py:36-52     Animal.step: move; energy -= 1; feed; die (energy < 0) or
//                                                         reproduce (random() < p: energy /= 2, offspring in the same cell)
//   agents.py:58-99    Sheep.feed (eat the cell's patch if fully grown) / Sheep.move (a von Neumann neighbour without
//                      wolves, preferring fully grown grass; random.choice)
//   agents.py:105-122  Wolf.feed (random.choice of the sheep in the cell, in the cell's list order) / Wolf.move (a
//                      neighbour cell with sheep if any, select_random_cell)
//   agents.py:125-158  GrassPatch (eaten -> regrows grass_regrowth_time later)
//
// shuffle_do is sequential: animal #rho of the shuffled order sees what the animals before it did.  Both phases are
// computed in parallel as the unique fixed point of a Jacobi iteration, like the Schelling relocation (csrc/relocate.cu):
//   * sheep only interact through the grass: a patch is eaten by the EARLIEST sheep that ends its move on it, so the
//     only unknown is eaten_at[cell] = priority of that sheep (never if none).  Given a table E, every sheep decides
//     exactly like the reference with "patch c is grown at my time t" = grown at the start && E[c] >= t; the new table is
//     the minimum priority per destination.  The earliest sheep decides correctly for any E, by induction sheep k is
//     right after k passes, a pass that reproduces its input table ends the iteration -- in practice a handful of passes;
//   * wolves only interact through the sheep they eat: the unknown is eaten_by[sheep] = priority of the wolf that eats
//     it; "sheep s is alive at my time t" = D[s] >= t.  A wolf's victim is the m-th live sheep of its cell in the cell's
//     list order (arrival order: cell.py:143-170), m = _randbelow(#live).
// Every random number is keyed (csrc/philox.cuh): priority64(seed, epoch, unique_id - 1) orders the activations, the
// draws of a step are numbers 0, 1, 2, ... of the animal's own stream, so the reference can be driven with the same
// numbers (oracle/replay.py) and the trajectories compared bit for bit (float64 energies included).
#include "common.cuh"
#include "philox.cuh"

namespace {

constexpr unsigned long long kNever = ~0ull;

struct WsGeom {
    int h, w;
    // von Neumann neighbour k of flat cell c on the torus, in connection order (-1,0), (0,-1), (0,1), (1,0)  (grid.py:478-482)
    __device__ __forceinline__ int nbr(int c, int k) const {
        int i = c / w, j = c - i * w;
        if (k == 0) i = i == 0 ? h - 1 : i - 1;
        else if (k == 1) j = j == 0 ? w - 1 : j - 1;
        else if (k == 2) j = j == w - 1 ? 0 : j + 1;
        else i = i == h - 1 ? 0 : i + 1;
        return i * w + j;
    }
};

struct SheepDecision {
    int dest;        // cell after the move
    bool moved;
};

// agents.py:69-99 with the grass as it is at this sheep's time under the eaten-at table E
__device__ __forceinline__ SheepDecision sheep_move(const WsGeom& g, int c, unsigned long long t, const int32_t* __restrict__ wolves_in,
                                                    const uint8_t* __restrict__ grown, const unsigned long long* __restrict__ E,
                                                    mb::DrawStream& ds) {
    int safe[4], grassy[4], ns = 0, ng = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int n = g.nbr(c, k);
        if (wolves_in[n] > 0) continue;
        safe[ns++] = n;
        if (grown != nullptr && grown[n] && E[n] >= t) grassy[ng++] = n;
    }
    if (ns == 0) return SheepDecision{c, false};
    const int len = ng > 0 ? ng : ns;
    const int j = (int)ds.randbelow((uint64_t)len);
    return SheepDecision{ng > 0 ? grassy[j] : safe[j], true};
}

__global__ void ws_sheep_pass(WsGeom g, const int32_t* __restrict__ cell, const int64_t* __restrict__ uid, int64_t n, uint64_t seed,
                              uint32_t epoch, const int32_t* __restrict__ wolves_in, const uint8_t* __restrict__ grown,
                              const unsigned long long* __restrict__ E_prev, unsigned long long* __restrict__ E_next) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t a = (uint64_t)(uid[s] - 1);
    const unsigned long long t = mb::priority64(seed, epoch, a);
    mb::DrawStream ds(seed, epoch, a);
    const SheepDecision d = sheep_move(g, cell[s], t, wolves_in, grown, E_prev, ds);
    if (grown != nullptr && grown[d.dest]) atomicMin(&E_next[d.dest], t);
}

// changed |= (a != b); a is reset to "never" (it becomes the next pass's output table)
__global__ void ws_compare_reset(unsigned long long* __restrict__ a, const unsigned long long* __restrict__ b, int64_t n,
                                 unsigned* __restrict__ changed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (a[i] != b[i]) *changed = 1u;
    a[i] = kNever;
}

__global__ void ws_sheep_apply(WsGeom g, int32_t* __restrict__ cell, double* __restrict__ energy, const int64_t* __restrict__ uid,
                               int64_t n, uint64_t seed, uint32_t epoch, const int32_t* __restrict__ wolves_in,
                               uint8_t* __restrict__ grown_out, const uint8_t* __restrict__ grown, int32_t* __restrict__ regrow_at,
                               const unsigned long long* __restrict__ E, double gain, double p_reproduce, int32_t regrow_time,
                               uint8_t* __restrict__ moved, uint8_t* __restrict__ dead, uint8_t* __restrict__ repro,
                               unsigned long long* __restrict__ prio) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t a = (uint64_t)(uid[s] - 1);
    const unsigned long long t = mb::priority64(seed, epoch, a);
    mb::DrawStream ds(seed, epoch, a);
    const SheepDecision d = sheep_move(g, cell[s], t, wolves_in, grown, E, ds);
    double e = energy[s] - 1.0;                                   // agents.py:43
    if (grown != nullptr && grown[d.dest] && E[d.dest] == t) {    // the earliest sheep on a grown patch eats it (agents.py:61-67)
        e += gain;
        grown_out[d.dest] = 0;
        regrow_at[d.dest] = regrow_time;
    }
    bool is_dead = false, rep = false;
    if (e < 0.0) is_dead = true;                                  // agents.py:49-52
    else if (ds.uniform01() < p_reproduce) { rep = true; e = e / 2.0; }
    cell[s] = d.dest;
    energy[s] = e;
    moved[s] = d.moved;
    dead[s] = is_dead;
    repro[s] = rep;
    prio[s] = t;
}

struct WolfDecision {
    int dest;
    long long victim;   // sheep row, -1 if none
};

// agents.py:113-122 (move) + 106-111 (feed) with the sheep that are alive at this wolf's time under the eaten-by table D
__device__ __forceinline__ WolfDecision wolf_decide(const WsGeom& g, int c, unsigned long long t, const int32_t* __restrict__ slist,
                                                    const int32_t* __restrict__ cstart, const unsigned long long* __restrict__ D,
                                                    mb::DrawStream& ds) {
    int with[4], nbrs[4], nw = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int n = g.nbr(c, k);
        nbrs[k] = n;
        bool any = false;
        for (int q = cstart[n]; q < cstart[n + 1] && !any; ++q) any = D[slist[q]] >= t;
        if (any) with[nw++] = n;
    }
    const int len = nw > 0 ? nw : 4;
    const int j = (int)ds.randbelow((uint64_t)len);
    const int dest = nw > 0 ? with[j] : nbrs[j];
    int alive = 0;
    for (int q = cstart[dest]; q < cstart[dest + 1]; ++q) alive += D[slist[q]] >= t;
    long long victim = -1;
    if (alive > 0) {
        int m = (int)ds.randbelow((uint64_t)alive);
        for (int q = cstart[dest]; q < cstart[dest + 1]; ++q) {
            const int s = slist[q];
            if (D[s] >= t) {
                if (m == 0) { victim = s; break; }
                --m;
            }
        }
    }
    return WolfDecision{dest, victim};
}

__global__ void ws_wolf_pass(WsGeom g, const int32_t* __restrict__ cell, const int64_t* __restrict__ uid, int64_t n, uint64_t seed,
                             uint32_t epoch, const int32_t* __restrict__ slist, const int32_t* __restrict__ cstart,
                             const unsigned long long* __restrict__ D_prev, unsigned long long* __restrict__ D_next) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const uint64_t a = (uint64_t)(uid[w] - 1);
    const unsigned long long t = mb::priority64(seed, epoch, a);
    mb::DrawStream ds(seed, epoch, a);
    const WolfDecision d = wolf_decide(g, cell[w], t, slist, cstart, D_prev, ds);
    if (d.victim >= 0) atomicMin(&D_next[d.victim], t);
}

__global__ void ws_wolf_apply(WsGeom g, int32_t* __restrict__ cell, double* __restrict__ energy, const int64_t* __restrict__ uid,
                              int64_t n, uint64_t seed, uint32_t epoch, const int32_t* __restrict__ slist,
                              const int32_t* __restrict__ cstart, const unsigned long long* __restrict__ D, double gain,
                              double p_reproduce, uint8_t* __restrict__ dead, uint8_t* __restrict__ repro,
                              unsigned long long* __restrict__ prio) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const uint64_t a = (uint64_t)(uid[w] - 1);
    const unsigned long long t = mb::priority64(seed, epoch, a);
    mb::DrawStream ds(seed, epoch, a);
    const WolfDecision d = wolf_decide(g, cell[w], t, slist, cstart, D, ds);
    double e = energy[w] - 1.0;
    if (d.victim >= 0) e += gain;       // at the fixed point every wolf's victim is its own (D[victim] == t)
    bool is_dead = false, rep = false;
    if (e < 0.0) is_dead = true;
    else if (ds.uniform01() < p_reproduce) { rep = true; e = e / 2.0; }
    cell[w] = d.dest;
    energy[w] = e;
    dead[w] = is_dead;
    repro[w] = rep;
    prio[w] = t;
}

// counts[c] += 1 for every agent in cell c
__global__ void ws_count_cells(const int32_t* __restrict__ cell, int64_t n, int32_t* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[cell[i]], 1);
}

// cstart[c] = first position of cell c in the cell-sorted array (cstart[ncells] = n)
__global__ void ws_segment_starts(const int32_t* __restrict__ sorted_cell, int64_t n, int ncells, int32_t* __restrict__ cstart) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int lo = i == 0 ? 0 : sorted_cell[i - 1] + 1;
    const int hi = i == n ? ncells : sorted_cell[i];
    for (int c = lo; c <= hi; ++c) cstart[c] = (int32_t)i;
}

// grown |= regrow_at <= now (the regrowth events of agents.py:150-158 that are due), pending entries cleared
__global__ void ws_regrow(uint8_t* __restrict__ grown, int32_t* __restrict__ regrow_at, int64_t ncells, int32_t now) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const int32_t r = regrow_at[c];
    if (r >= 0 && r <= now) { grown[c] = 1; regrow_at[c] = -1; }
}

inline unsigned blocks_for(int64_t n) { return (unsigned)mb::ceil_div(n > 0 ? n : 1, 256); }

}  // namespace

#define WS_GEOM_CHECK(h, w) MB_REQUIRE((h) >= 3 && (w) >= 3 && (int64_t)(h) * (w) < (1ll << 30), "wolf-sheep: the torus needs 3 <= height, width and < 2^30 cells")

MB_API int mb_ws_count_cells(const int32_t* cell, int64_t n, int32_t* counts, int64_t ncells, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && ncells >= 0 && counts != nullptr && (n == 0 || cell != nullptr), "mb_ws_count_cells: bad argument");
    int rc = mb::check_cuda(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)ncells, stream), "memset");
    if (rc) return rc;
    if (n > 0) ws_count_cells<<<blocks_for(n), 256, 0, stream>>>(cell, n, counts);
    MB_LAUNCH_CHECK("mb_ws_count_cells");
    return MB_OK;
}

MB_API int mb_ws_segment_starts(const int32_t* sorted_cell, int64_t n, int64_t ncells, int32_t* cstart, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && ncells >= 0 && ncells < (1ll << 31) && cstart != nullptr && (n == 0 || sorted_cell != nullptr), "mb_ws_segment_starts: bad argument");
    ws_segment_starts<<<blocks_for(n + 1), 256, 0, stream>>>(sorted_cell, n, (int)ncells, cstart);
    MB_LAUNCH_CHECK("mb_ws_segment_starts");
    return MB_OK;
}

// One Jacobi pass of the sheep phase: E_next[c] = min priority of the sheep that end their move on the grown patch c,
// deciding with E_prev.  E_next must hold "never" (all bits set) on entry.  grown == NULL: the model runs without grass.
MB_API int mb_ws_sheep_pass(const int32_t* cell, const int64_t* uid, int64_t n, int height, int width, uint64_t seed, uint32_t epoch,
                            const int32_t* wolves_in, const uint8_t* grown, const uint64_t* E_prev, uint64_t* E_next,
                            cudaStream_t stream) {
    WS_GEOM_CHECK(height, width);
    MB_REQUIRE(n >= 0 && wolves_in && E_prev && E_next && (n == 0 || (cell && uid)), "mb_ws_sheep_pass: null buffer");
    if (n > 0)
        ws_sheep_pass<<<blocks_for(n), 256, 0, stream>>>(WsGeom{height, width}, cell, uid, n, seed, epoch, wolves_in, grown,
                                                         (const unsigned long long*)E_prev, (unsigned long long*)E_next);
    MB_LAUNCH_CHECK("mb_ws_sheep_pass");
    return MB_OK;
}

// *changed (device word, accumulates) = 1 if the two tables differ; `a` is reset to "never"
MB_API int mb_ws_compare_reset(uint64_t* a, const uint64_t* b, int64_t n, unsigned* changed, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && changed != nullptr && (n == 0 || (a && b)), "mb_ws_compare_reset: bad argument");
    if (n > 0) ws_compare_reset<<<blocks_for(n), 256, 0, stream>>>((unsigned long long*)a, (const unsigned long long*)b, n, changed);
    MB_LAUNCH_CHECK("mb_ws_compare_reset");
    return MB_OK;
}

// The sheep phase at its fixed point E: cell / energy updated in place, the eaten patches cleared (grown_out may alias
// nothing read here: pass a copy of the layer as `grown`), regrow_at set to regrow_time for them; per sheep: moved, dead
// (energy < 0), repro (offspring: same cell, the halved energy) and its priority.
MB_API int mb_ws_sheep_apply(int32_t* cell, double* energy, const int64_t* uid, int64_t n, int height, int width, uint64_t seed,
                             uint32_t epoch, const int32_t* wolves_in, const uint8_t* grown, uint8_t* grown_out, int32_t* regrow_at,
                             const uint64_t* E, double gain, double p_reproduce, int32_t regrow_time, uint8_t* moved, uint8_t* dead,
                             uint8_t* repro, uint64_t* prio, cudaStream_t stream) {
    WS_GEOM_CHECK(height, width);
    MB_REQUIRE(n >= 0 && wolves_in && E && (n == 0 || (cell && energy && uid && moved && dead && repro && prio)), "mb_ws_sheep_apply: null buffer");
    MB_REQUIRE(grown == nullptr || (grown_out != nullptr && regrow_at != nullptr && grown_out != grown), "mb_ws_sheep_apply: the grass layer needs a separate output layer");
    if (n > 0)
        ws_sheep_apply<<<blocks_for(n), 256, 0, stream>>>(WsGeom{height, width}, cell, energy, uid, n, seed, epoch, wolves_in, grown_out,
                                                          grown, regrow_at, (const unsigned long long*)E, gain, p_reproduce, regrow_time,
                                                          moved, dead, repro, (unsigned long long*)prio);
    MB_LAUNCH_CHECK("mb_ws_sheep_apply");
    return MB_OK;
}

// One Jacobi pass of the wolf phase: D_next[s] = min priority of the wolves that pick sheep s, deciding with D_prev.
// slist: sheep rows sorted by (cell, position in the cell's list); cstart [ncells + 1].
MB_API int mb_ws_wolf_pass(const int32_t* cell, const int64_t* uid, int64_t n, int height, int width, uint64_t seed, uint32_t epoch,
                           const int32_t* slist, const int32_t* cstart, const uint64_t* D_prev, uint64_t* D_next, cudaStream_t stream) {
    WS_GEOM_CHECK(height, width);
    MB_REQUIRE(n >= 0 && cstart && D_prev && D_next && (n == 0 || (cell && uid)), "mb_ws_wolf_pass: null buffer");
    if (n > 0)
        ws_wolf_pass<<<blocks_for(n), 256, 0, stream>>>(WsGeom{height, width}, cell, uid, n, seed, epoch, slist, cstart,
                                                        (const unsigned long long*)D_prev, (unsigned long long*)D_next);
    MB_LAUNCH_CHECK("mb_ws_wolf_pass");
    return MB_OK;
}

MB_API int mb_ws_wolf_apply(int32_t* cell, double* energy, const int64_t* uid, int64_t n, int height, int width, uint64_t seed,
                            uint32_t epoch, const int32_t* slist, const int32_t* cstart, const uint64_t* D, double gain,
                            double p_reproduce, uint8_t* dead, uint8_t* repro, uint64_t* prio, cudaStream_t stream) {
    WS_GEOM_CHECK(height, width);
    MB_REQUIRE(n >= 0 && cstart && D && (n == 0 || (cell && energy && uid && dead && repro && prio)), "mb_ws_wolf_apply: null buffer");
    if (n > 0)
        ws_wolf_apply<<<blocks_for(n), 256, 0, stream>>>(WsGeom{height, width}, cell, energy, uid, n, seed, epoch, slist, cstart,
                                                         (const unsigned long long*)D, gain, p_reproduce, dead, repro,
                                                         (unsigned long long*)prio);
    MB_LAUNCH_CHECK("mb_ws_wolf_apply");
    return MB_OK;
}

MB_API int mb_ws_regrow(uint8_t* grown, int32_t* regrow_at, int64_t ncells, int32_t now, cudaStream_t stream) {
    MB_REQUIRE(ncells >= 0 && (ncells == 0 || (grown && regrow_at)), "mb_ws_regrow: null buffer");
    if (ncells > 0) ws_regrow<<<blocks_for(ncells), 256, 0, stream>>>(grown, regrow_at, ncells, now);
    MB_LAUNCH_CHECK("mb_ws_regrow");
    return MB_OK;
}
