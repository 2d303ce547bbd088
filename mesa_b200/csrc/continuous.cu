// Continuous space: cell list build, radius-neighbour search and the Boids step (K7+K8 of
// SURVEY.md §2.4).
//
// Replaces, for all agents at once,
//   mesa/experimental/continuous_space/continuous_space.py:202-235  calculate_distances
//       torus: delta = abs(point - positions); delta = minimum(delta, size - delta);
//              dists = sqrt(delta[:,0]**2 + delta[:,1]**2)
//       non-torus: scipy cdist (Euclidean)
//   continuous_space.py:237-247   get_agents_in_radius: dists <= radius (INCLUSIVE)
//   continuous_space_agents.py:66-78   get_neighbors_in_radius: drops self
//   continuous_space.py:169-200   calculate_difference_vector (torus: pick delta or
//                                 delta - sign(delta)*size, whichever is STRICTLY smaller in magnitude)
//   mesa/examples/basic/boid_flockers/agents.py:63-92   Boid.step
//   continuous_space_agents.py:36-44 + continuous_space.py:285-298   position setter (wrap)
//
// The reference scans all N agents per query (O(N^2) per step).  Here agents are binned into
// a uniform cell grid with edge >= radius (radix sort by cell id + boundary detection), and a
// query only visits the 3x3 cells around it.
//
// State is fp32 (positions, directions); every pair predicate and the update arithmetic are
// evaluated in fp64 with the reference's operation order (no FMA contraction), so for inputs
// that are fp32-representable the neighbour SETS are bit-identical to the reference's and
// the new position/direction differ only by the final fp32 rounding.
//
// The Boids step is SYNCHRONOUS (every boid reads the pre-step state), whereas
// AgentSet.shuffle_do is sequential; DESIGN.md discusses the mode and how parity is checked
// (reference Boid.step driven on frozen snapshots).
#include "common.cuh"
#include "philox.cuh"
#include "sort.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

namespace {

struct Geometry {
    double x0, y0;      // lower corner (dimensions[:,0])
    double sx, sy;      // size = dimensions[:,1] - dimensions[:,0]
    int ncx, ncy;       // cells per axis (edge >= radius)
    int torus;
};

__device__ __forceinline__ int cell_coord(double p, double lo, double size, int nc) {
    int c = (int)floor((p - lo) * (double)nc / size);
    return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

__device__ __forceinline__ unsigned cell_of(const Geometry& g, float x, float y) {
    return (unsigned)(cell_coord((double)y, g.y0, g.sy, g.ncy) * g.ncx + cell_coord((double)x, g.x0, g.sx, g.ncx));
}

// ---- cell list: a counting sort by cell, deterministic ------------------------------------------------------------
// The agents are grouped by cell (not fully sorted: nothing needs an order between cells beyond the row-major cell id,
// which the prefix sum provides).  1. count_cells: every agent takes a ticket in its cell (atomicAdd on the cell's
// counter); 2. an exclusive prefix sum turns the counts into cell_start[0 .. ncells] (cell_end[c] = cell_start[c+1]);
// 3. scatter_cells writes the agent into slot cell_start[cell] + ticket; 4. the tickets depend on the scheduling of
// step 1, so every cell's few agents are put into ascending agent index (insertion sort by the thread that owns the
// cell; cells longer than kCellSortMax are handled by cooperative CTAs, rank by counting) -- the same order a stable
// sort of (cell, agent) gives, i.e. the storage-index order in which the reference lists a cell's agents;
// 5. gather_state packs (x, y, dx, dy) in that order (tried: writing the state in the scatter and permuting it with the
// ids -- 10^7 random 16-byte stores cost 0.57 ms, the gather of 8-byte reads 0.25 ms).  10^7 boids / 2.5 M cells:
// 0.52 ms instead of 0.79 ms with the three-pass radix sort (which other kernels of the library still use for 64-bit keys).
__global__ void count_cells(const float2* __restrict__ pos, int64_t n, Geometry g, uint32_t* __restrict__ counts,
                            uint32_t* __restrict__ cell_of_agent, uint32_t* __restrict__ ticket) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = pos[i];
    const uint32_t c = cell_of(g, p.x, p.y);
    cell_of_agent[i] = c;
    ticket[i] = atomicAdd(&counts[c], 1u);
}

__global__ void scatter_cells(const uint32_t* __restrict__ cell_of_agent, const uint32_t* __restrict__ ticket, int64_t n,
                              const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[cell_start[cell_of_agent[i]] + ticket[i]] = (uint32_t)i;
}

constexpr int kCellSortMax = 32;

// ascending agent index inside every cell; long cells are queued for sort_long_cells
__global__ void sort_cells(uint32_t* __restrict__ order, const uint32_t* __restrict__ cell_start, int64_t ncells,
                           uint32_t* __restrict__ long_cells, uint32_t* __restrict__ long_count) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t s = cell_start[c], e = cell_start[c + 1];
    const uint32_t len = e - s;
    if (len <= 1u) return;
    if (len > (uint32_t)kCellSortMax) {
        long_cells[atomicAdd(long_count, 1u)] = (uint32_t)c;
        return;
    }
    uint32_t v[kCellSortMax];
    for (uint32_t k = 0; k < len; ++k) v[k] = order[s + k];
    for (uint32_t a = 1; a < len; ++a) {
        const uint32_t x = v[a];
        uint32_t b = a;
        while (b > 0 && v[b - 1] > x) { v[b] = v[b - 1]; --b; }
        v[b] = x;
    }
    for (uint32_t k = 0; k < len; ++k) order[s + k] = v[k];
}

// cells with more than kCellSortMax agents (dense clusters): one CTA per cell, every agent's position = the number of
// smaller agent indices in the cell (rank by counting through a scratch copy; quadratic, but so is the neighbour work
// of such a cell)
__global__ void __launch_bounds__(256)
sort_long_cells(uint32_t* __restrict__ order, uint32_t* __restrict__ scratch, const uint32_t* __restrict__ cell_start,
                const uint32_t* __restrict__ long_cells, const uint32_t* __restrict__ long_count) {
    for (uint32_t q = blockIdx.x; q < *long_count; q += gridDim.x) {
        const uint32_t c = long_cells[q];
        const uint32_t s = cell_start[c], len = cell_start[c + 1] - s;
        for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) scratch[s + k] = order[s + k];
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) {
            const uint32_t x = scratch[s + k];
            uint32_t r = 0;
            for (uint32_t j = 0; j < len; ++j) r += scratch[s + j] < x;
            order[s + r] = x;
        }
        __syncthreads();
    }
}

// packed[i] = (x, y, dx, dy) of the agent at sorted slot i
__global__ void gather_state(const float2* __restrict__ pos, const float2* __restrict__ dir,
                             const uint32_t* __restrict__ order, int64_t n, float4* __restrict__ packed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = order[i];
    const float2 p = pos[a];
    const float2 d = dir != nullptr ? dir[a] : make_float2(0.f, 0.f);
    packed[i] = make_float4(p.x, p.y, d.x, d.y);
}

// squared distance exactly as calculate_distances computes it before the sqrt (continuous_space.py:223-233).
// The reference then tests sqrt(d2) <= radius; sqrt is monotone and correctly rounded, so that test equals
// d2 <= T with T = the largest double whose rounded sqrt is <= radius (computed on the host,
// sq_threshold_le) -- bit-identical neighbour sets without a double-precision sqrt per pair.
__device__ __forceinline__ double ref_distance2(const Geometry& g, double px, double py, double qx, double qy) {
    double dx = fabs(px - qx), dy = fabs(py - qy);
    if (g.torus) {
        dx = fmin(dx, g.sx - dx);
        dy = fmin(dy, g.sy - dy);
    }
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// one component of calculate_difference_vector (continuous_space.py:186-200)
__device__ __forceinline__ double ref_delta(double q, double p, double size, int torus) {
    const double d = q - p;
    if (!torus) return d;
    const double sgn = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);
    const double inv = d - sgn * size;
    return fabs(d) < fabs(inv) ? d : inv;
}

// visit the distinct cells of the 3x3 block around (cx, cy)
template <typename F>
__device__ __forceinline__ void for_neighbor_cells(const Geometry& g, int cx, int cy, F&& f) {
    const int nx = g.ncx < 3 ? g.ncx : 3, ny = g.ncy < 3 ? g.ncy : 3;
    for (int iy = 0; iy < ny; ++iy) {
        int yy = g.ncy < 3 ? iy : cy + iy - 1;
        if (yy < 0 || yy >= g.ncy) {
            if (!g.torus) continue;
            yy = (yy + g.ncy) % g.ncy;
        }
        for (int ix = 0; ix < nx; ++ix) {
            int xx = g.ncx < 3 ? ix : cx + ix - 1;
            if (xx < 0 || xx >= g.ncx) {
                if (!g.torus) continue;
                xx = (xx + g.ncx) % g.ncx;
            }
            f((unsigned)(yy * g.ncx + xx));
        }
    }
}

struct BoidParams {
    double vision2_le;      // d <= vision      <=>  d2 <= vision2_le
    double separation2_lt;  // d <  separation  <=>  d2 <  separation2_lt
    double cohere, separate, match, speed;
};

// accumulators of one boid over its neighbours (agents.py:75-85)
struct BoidAcc {
    double cohx = 0.0, cohy = 0.0, sepx = 0.0, sepy = 0.0, mx = 0.0, my = 0.0;
    unsigned k = 0;
};

// one candidate (position o, direction od; float64) against the boid at (px, py): agents.py:66-85
__device__ __forceinline__ void boid_pair_d(const Geometry& g, const BoidParams& bp, double px, double py,
                                            double ox, double oy, double odx, double ody, BoidAcc& acc) {
    const double d2 = ref_distance2(g, px, py, ox, oy);
    if (d2 <= bp.vision2_le) {
        const double ddx = ref_delta(ox, px, g.sx, g.torus);
        const double ddy = ref_delta(oy, py, g.sy, g.torus);
        acc.cohx += ddx; acc.cohy += ddy;
        if (d2 < bp.separation2_lt) { acc.sepx += ddx; acc.sepy += ddy; }
        acc.mx += odx; acc.my += ody;
        ++acc.k;
    }
}

__device__ __forceinline__ void boid_pair(const Geometry& g, const BoidParams& bp, double px, double py,
                                          const float4& o, BoidAcc& acc) {
    boid_pair_d(g, bp, px, py, (double)o.x, (double)o.y, (double)o.z, (double)o.w, acc);
}

// direction / position update of one boid from its accumulators (agents.py:75-92 + the position setter):
// returns (x, y, dx, dy) in float64
__device__ __forceinline__ double4 boid_finish_d(const Geometry& g, const BoidParams& bp, const float4& me,
                                                 const BoidAcc& acc) {
    double dirx = me.z, diry = me.w;
    double nx = me.x, ny = me.y;
    if (acc.k > 0) {
        // agents.py:75-89
        const double kk = (double)acc.k;
        dirx += (acc.cohx * bp.cohere + (-1.0 * acc.sepx * bp.separate) + acc.mx * bp.match) / kk;
        diry += (acc.cohy * bp.cohere + (-1.0 * acc.sepy * bp.separate) + acc.my * bp.match) / kk;
        const double norm = __dsqrt_rn(__dadd_rn(__dmul_rn(dirx, dirx), __dmul_rn(diry, diry)));
        dirx /= norm;
        diry /= norm;
    }
    nx += dirx * bp.speed;
    ny += diry * bp.speed;
    // position setter: in_bounds is inclusive on both ends (continuous_space.py:285-292)
    const bool inb = nx >= g.x0 && nx <= g.x0 + g.sx && ny >= g.y0 && ny <= g.y0 + g.sy;
    if (!inb && g.torus) {
        // torus_correct: lo + np.mod(p - lo, size)  (floored modulo)
        double rx = fmod(nx - g.x0, g.sx); if (rx != 0.0 && rx < 0.0) rx += g.sx;
        double ry = fmod(ny - g.y0, g.sy); if (ry != 0.0 && ry < 0.0) ry += g.sy;
        nx = g.x0 + rx;
        ny = g.y0 + ry;
    }
    return make_double4(nx, ny, dirx, diry);
}

// same, stored as float32 at agent index `a`
__device__ __forceinline__ void boid_finish(const Geometry& g, const BoidParams& bp, const float4& me,
                                            const BoidAcc& acc, uint32_t a, float2* __restrict__ pos_out,
                                            float2* __restrict__ dir_out, uint32_t* __restrict__ nbr_count) {
    const double4 s = boid_finish_d(g, bp, me, acc);
    pos_out[a] = make_float2((float)s.x, (float)s.y);
    dir_out[a] = make_float2((float)s.z, (float)s.w);
    if (nbr_count != nullptr) nbr_count[a] = acc.k;
}

// Generic path (any geometry): one thread per boid, neighbours through L1/L2.
__global__ void __launch_bounds__(128)
boids_update(const float4* __restrict__ packed, const uint32_t* __restrict__ order,
             const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end, int64_t n, Geometry g,
             BoidParams bp, float2* __restrict__ pos_out, float2* __restrict__ dir_out,
             uint32_t* __restrict__ nbr_count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 me = packed[i];
    const double px = me.x, py = me.y;
    const int cx = cell_coord(px, g.x0, g.sx, g.ncx), cy = cell_coord(py, g.y0, g.sy, g.ncy);
    BoidAcc acc;
    for_neighbor_cells(g, cx, cy, [&](unsigned c) {
        const uint32_t e = cell_end[c];
        for (uint32_t j = cell_start[c]; j < e; ++j) {
            if ((int64_t)j == i) continue;  // get_neighbors_in_radius drops self
            boid_pair(g, bp, px, py, packed[j], acc);
        }
    });
    boid_finish(g, bp, me, acc, order[i], pos_out, dir_out, nbr_count);
}

// Shared-memory tiled path: one CTA owns a tile of kTX x kTY cells.  The boids of the tile and of the
// one-cell ring around it ((kTX+2) x (kTY+2) cells) are staged once in shared memory; every owned boid then
// scans three contiguous shared-memory runs (the 3 cells of each of the 3 rows around it).
//
// Arithmetic: the reference works in float64 on positions that are float32-representable here, and its result is
// rounded to float32 at the end of the step.  What has to be EXACT is the membership of the neighbour set
// (`dist <= vision`, inclusive) and of the separation set (`dist < separation`); the sums only have to be good
// to float32.  So a pair is classified in float32 first -- d2 from the float32 differences carries a relative
// error below 3e-7 -- and
//   d2 <  vision^2 (1 - 1e-6)  surely a neighbour: its difference vector IS the float32 (dx, dy) (the unwrapped
//                              image is the nearest one: a tiled space is at least 10 cells wide), accumulated
//                              in float32 (5 adds);
//   d2 >  vision^2 (1 + 1e-6)  surely not (unless the pair may straddle the torus seam);
//   else                       the reference's float64 predicate and difference vector (boid_pair), ~1 pair in 1e5.
// Pairs that may be neighbours across the torus seam (|dx| > size - vision, only in tiles whose ring wraps) also
// take the float64 path: float32 cannot subtract the space size accurately.  The same three-way test is used for
// `dist < separation`.  The float64 accumulators of the rare path and the float32 ones are added in boid_finish,
// which is float64 throughout (normalisation, move, wrap) like the reference.
// History (10^7 boids, B200): thread per boid through L1 2.50 ms -> threshold on d2 instead of sqrt 2.07 ms ->
// shared-memory tiles 1.8 ms -> float32 reject filter + float64 survivor lists 1.73 ms -> this kernel.
// Tiles whose ring holds more than kTileCap boids (very dense regions) report themselves and are redone by the
// generic kernel.
constexpr int kTX = 8, kTY = 8;
constexpr int kHX = kTX + 2, kHY = kTY + 2;
constexpr int kTileCap = 1536;
constexpr int kTileThreads = 256;

struct TileSmem {
    float4 cand[kTileCap];
    uint32_t gidx[kTileCap];                  // cell-sorted index of each staged boid
    uint32_t cell_off[kHX * kHY + 1];         // start of each ring cell in cand[]
    uint32_t cell_src[kHX * kHY];             // start of each ring cell in the global sorted array
    uint32_t own_off[kTY + 1];
};

struct FastParams {
    float vis_lo2, vis_hi2;   // float32 band around vision2_le
    float sep_lo2, sep_hi2;   // float32 band around separation2_lt
    float seam_x, seam_y;     // |dx| above this may be a neighbour across the torus seam
};

__global__ void __launch_bounds__(kTileThreads, 4)
boids_update_tiled(const float4* __restrict__ packed, const uint32_t* __restrict__ order,
                   const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end, Geometry g,
                   BoidParams bp, FastParams fp, int tiles_x, float2* __restrict__ pos_out,
                   float2* __restrict__ dir_out, uint32_t* __restrict__ nbr_count,
                   uint32_t* __restrict__ overflow_tiles, uint32_t* __restrict__ overflow_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cx0 = (blockIdx.x % tiles_x) * kTX, cy0 = (blockIdx.x / tiles_x) * kTY;

    // 1. sizes of the ring cells (wrapped on a torus, absent outside a clamped space)
    if (tid < kHX * kHY) {
        int cx = cx0 - 1 + tid % kHX, cy = cy0 - 1 + tid / kHX;
        bool exists = true;
        if (g.torus) {
            cx = (cx + g.ncx) % g.ncx;
            cy = (cy + g.ncy) % g.ncy;
        } else {
            exists = cx >= 0 && cx < g.ncx && cy >= 0 && cy < g.ncy;
        }
        uint32_t s = 0, e = 0;
        if (exists) {
            const unsigned c = (unsigned)cy * g.ncx + cx;
            s = cell_start[c];
            e = cell_end[c];
        }
        sm.cell_src[tid] = s;
        sm.cell_off[tid + 1] = e - s;
    }
    if (tid == 0) sm.cell_off[0] = 0;
    __syncthreads();
    if (warp == 0) {  // inclusive scan of kHX*kHY = 100 counts, 4 per lane
        unsigned v[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane * 4 + k;
            v[k] = i < kHX * kHY ? sm.cell_off[i + 1] : 0u;
            sum += v[k];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        unsigned run = incl - sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane * 4 + k;
            run += v[k];
            if (i < kHX * kHY) sm.cell_off[i + 1] = run;
        }
    }
    __syncthreads();
    const unsigned total = sm.cell_off[kHX * kHY];
    if (total > kTileCap) {  // too dense for the staging buffer: hand the tile to the generic kernel
        if (tid == 0) overflow_tiles[atomicAdd(overflow_count, 1u)] = blockIdx.x;
        return;
    }
    // 2. stage, flat over all ring boids (a ring cell holds only a few: per-cell loops would idle most lanes)
    for (unsigned i = tid; i < total; i += kTileThreads) {
        int lo = 0, hi = kHX * kHY;  // last cell with cell_off <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sm.cell_off[mid] <= i) lo = mid; else hi = mid;
        }
        const unsigned src = sm.cell_src[lo] + (i - sm.cell_off[lo]);
        sm.cand[i] = packed[src];
        sm.gidx[i] = src;
    }
    // 3. owned boids: the interior cells of ring row iy form the contiguous run
    //    [cell_off[iy*kHX+1], cell_off[iy*kHX+kHX-1]); flatten the kTY runs so that all threads stay busy
    if (tid == 0) {
        unsigned run = 0;
        for (int iy = 1; iy <= kTY; ++iy) {
            sm.own_off[iy - 1] = run;
            if (cy0 + iy - 1 < g.ncy) run += sm.cell_off[iy * kHX + kHX - 1] - sm.cell_off[iy * kHX + 1];
        }
        sm.own_off[kTY] = run;
    }
    __syncthreads();
    const unsigned owned = sm.own_off[kTY];
    // only tiles whose ring wraps around the space can hold pairs that are neighbours across the seam
    const bool seam_tile = g.torus && (cx0 == 0 || cy0 == 0 || cx0 + kTX >= g.ncx || cy0 + kTY >= g.ncy);
    for (unsigned t = tid; t < owned; t += kTileThreads) {
        int iy = 1;
#pragma unroll
        for (int k = 1; k < kTY; ++k) iy += t >= sm.own_off[k];
        const unsigned p = sm.cell_off[iy * kHX + 1] + (t - sm.own_off[iy - 1]);
        const float4 me = sm.cand[p];
        const double px = me.x, py = me.y;
        const int cx = cell_coord(px, g.x0, g.sx, g.ncx);
        if (cx < cx0 || cx >= cx0 + kTX) continue;  // wrapped duplicate of a partial edge tile: owned elsewhere
        const int ix = cx - cx0 + 1;                // ring column of my cell
        BoidAcc acc;                                // float64: the rare exact path
        float cohx = 0.f, cohy = 0.f, sepx = 0.f, sepy = 0.f, mx = 0.f, my = 0.f;
        unsigned k = 0;
        // The boid itself is one of the candidates (d2 = 0): instead of testing `q == p` for every pair it is counted
        // like a neighbour -- it adds 0 to the difference sums and its own direction to the match sum -- and taken
        // out again after the loop (get_neighbors_in_radius drops self, continuous_space_agents.py:66-78).
        auto scan = [&](auto seam) {
#pragma unroll
            for (int r = iy - 1; r <= iy + 1; ++r) {
                const float4* q = sm.cand + sm.cell_off[r * kHX + ix - 1];
                const float4* const qe = sm.cand + sm.cell_off[r * kHX + ix + 2];
                for (; q < qe; ++q) {
                    const float4 o = *q;
                    const float dx = o.x - me.x, dy = o.y - me.y;
                    const float d2 = fmaf(dx, dx, dy * dy);
                    if (d2 > fp.vis_hi2) {  // surely not a neighbour -- unless the pair straddles the torus seam
                        if (decltype(seam)::value && (fabsf(dx) > fp.seam_x || fabsf(dy) > fp.seam_y))
                            boid_pair(g, bp, px, py, o, acc);
                        continue;
                    }
                    if (d2 < fp.vis_lo2) {
                        cohx += dx; cohy += dy;
                        mx += o.z; my += o.w;
                        ++k;
                        if (d2 < fp.sep_hi2) {
                            if (d2 < fp.sep_lo2 || ref_distance2(g, px, py, (double)o.x, (double)o.y) < bp.separation2_lt) {
                                sepx += dx; sepy += dy;
                            }
                        }
                    } else {
                        boid_pair(g, bp, px, py, o, acc);  // borderline: the reference's float64 test
                    }
                }
            }
        };
        if (seam_tile) scan(std::true_type{}); else scan(std::false_type{});
        mx -= me.z; my -= me.w;   // take the boid itself out again
        k -= 1u;
        acc.cohx += (double)cohx; acc.cohy += (double)cohy;
        acc.sepx += (double)sepx; acc.sepy += (double)sepy;
        acc.mx += (double)mx; acc.my += (double)my;
        acc.k += k;
        boid_finish(g, bp, me, acc, order[sm.gidx[p]], pos_out, dir_out, nbr_count);
    }
}

// generic kernel restricted to the boids of the overflowed tiles (rare)
__global__ void __launch_bounds__(128)
boids_update_overflow(const float4* __restrict__ packed, const uint32_t* __restrict__ order,
                      const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end, Geometry g,
                      BoidParams bp, int tiles_x, const uint32_t* __restrict__ overflow_tiles,
                      const uint32_t* __restrict__ overflow_count, float2* __restrict__ pos_out,
                      float2* __restrict__ dir_out, uint32_t* __restrict__ nbr_count) {
    for (unsigned t = blockIdx.x; t < *overflow_count; t += gridDim.x) {
        const unsigned tile = overflow_tiles[t];
        const int cx0 = (tile % tiles_x) * kTX, cy0 = (tile / tiles_x) * kTY;
        for (int c = 0; c < kTX * kTY; ++c) {
            const int cx = cx0 + c % kTX, cy = cy0 + c / kTX;
            if (cx >= g.ncx || cy >= g.ncy) continue;
            const unsigned cell = (unsigned)cy * g.ncx + cx;
            const uint32_t e = cell_end[cell];
            for (uint32_t i = cell_start[cell] + threadIdx.x; i < e; i += blockDim.x) {
                const float4 me = packed[i];
                BoidAcc acc;
                for_neighbor_cells(g, cx, cy, [&](unsigned nc) {
                    const uint32_t ne = cell_end[nc];
                    for (uint32_t j = cell_start[nc]; j < ne; ++j) {
                        if (j == i) continue;
                        boid_pair(g, bp, (double)me.x, (double)me.y, packed[j], acc);
                    }
                });
                boid_finish(g, bp, me, acc, order[i], pos_out, dir_out, nbr_count);
            }
        }
    }
}

// ---- sequential activation (the reference's shuffle_do semantics, exactly) -------------------------------------
// AgentSet.shuffle_do("step") (agentset.py:772-787) activates the boids one after the other: boid a sees the NEW
// position / direction of every boid that acted before it and the OLD state of the others.  With the activation
// order given by the Philox priorities (csrc/philox.cuh) this is a dependency graph, not a chain: a only depends
// on the earlier boids that can be within `vision` of it when it acts, i.e. whose OLD position is within
// vision + reach (reach >= the longest move of a step).  The step is therefore computed in passes over the boids
// that are still pending: a boid whose relevant earlier neighbours are all done computes its new state (float64,
// kept in `newst` until the end of the step so that later boids read exactly what the reference would), publishes
// it (st.release) and is done; the others go to the next pass.  The pending boid of lowest priority is always
// ready, so every pass makes progress; on uniform random flocks the number of passes is ~ the longest chain of
// decreasing priorities through overlapping neighbourhoods (a few tens), and the pending set shrinks geometrically.
// The cell list is built on the OLD positions with edge >= vision + reach, so both the old and the new position
// of every relevant boid are covered by the 3x3 cells around a.
struct SeqParams {
    float relevant2;   // float32 bound on d2(old, old) beyond which two boids cannot interact in this step
    float far2;        // float32 bound on d2 beyond which an OLD-state candidate is surely not within vision
    double reach2;     // (reach (1 + 1e-9))^2: a longer move is reported as an error
    float fsx, fsy;
};

// Publication of a new state without flags or fences: `newst` holds two 16-byte halves per boid, (x, y) and (dx, dy),
// preset to NaN at the start of the step.  A half is written with ONE 16-byte store and read with ONE 16-byte
// L2-coherent load (ld.global.cg bypasses the non-coherent L1), so a reader sees either NaN ("not finished") or the
// complete half; it needs both halves to be valid.  No ordering between the halves is required, hence no
// acquire/release pair -- on sm_100a every ld.acquire.gpu costs an L1 invalidation (CCTL.IVALL).
__device__ __forceinline__ bool load_new_state(const double2* newst, uint32_t j, double2& p, double2& d) {
    p = __ldcg(newst + 2 * (size_t)j);
    d = __ldcg(newst + 2 * (size_t)j + 1);
    return p.x == p.x && d.x == d.x;  // NaN != NaN
}
__device__ __forceinline__ void store_new_state(double2* newst, uint32_t i, const double4& s) {
    __stcg(newst + 2 * (size_t)i, make_double2(s.x, s.y));
    __stcg(newst + 2 * (size_t)i + 1, make_double2(s.z, s.w));
}

__global__ void boids_seq_priorities(const uint32_t* __restrict__ order, int64_t n, uint64_t seed, uint32_t epoch,
                                     uint64_t* __restrict__ prio) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) prio[i] = mb::priority64(seed, epoch, (uint64_t)order[i]);
}

// round of a slot = top bits of its priority; members = slots sorted by round (stable: cell order inside a round)
__global__ void boids_seq_round_keys(const uint64_t* __restrict__ prio, int64_t n, int shift, uint32_t* __restrict__ keys,
                                     uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint32_t)(prio[i] >> shift);
    vals[i] = (uint32_t)i;
}
// round_start[k] = first position of the sorted key array with key >= k (k = 0 .. rounds)
__global__ void boids_seq_round_starts(const uint32_t* __restrict__ keys, int64_t n, int rounds,
                                       uint32_t* __restrict__ round_start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    const uint32_t prev = i == 0 ? 0u : keys[i - 1] + 1u;
    if (i == 0 || keys[i - 1] != k)
        for (uint32_t kk = prev; kk <= k; ++kk) round_start[kk] = (uint32_t)i;
    if (i == n - 1)
        for (uint32_t kk = k + 1u; kk <= (uint32_t)rounds; ++kk) round_start[kk] = (uint32_t)n;
}

// One pass: the `fresh` boids of round `round` (members[round_start[round] .. round_start[round+1]), cell order;
// round < 0: none) plus the boids left pending by the previous pass.  Rounds partition the boids by priority (round k
// holds the k-th slice of the activation order), so when round k is fresh every earlier round has been offered at
// least one pass: almost all of a boid's earlier neighbours are finished, and because a round is processed in CELL
// order its neighbourhood reads are spatially coherent (the rank-ordered chained kernel below reads at random).
__global__ void __launch_bounds__(128)
boids_seq_pass(const float4* __restrict__ packed, const uint64_t* __restrict__ prio, double2* newst,
               const uint32_t* __restrict__ order, const uint32_t* __restrict__ cell_start,
               const uint32_t* __restrict__ cell_end, int64_t n, Geometry g, BoidParams bp, SeqParams sp,
               const uint32_t* __restrict__ members, const uint32_t* __restrict__ round_start, int round,
               const uint32_t* __restrict__ pending_in, const uint32_t* __restrict__ count_in,
               uint32_t* __restrict__ pending_out, uint32_t* __restrict__ count_out, float2* __restrict__ pos_out,
               float2* __restrict__ dir_out, uint32_t* __restrict__ nbr_count, uint32_t* __restrict__ error) {
    const int64_t fresh0 = round >= 0 ? (int64_t)round_start[round] : 0;
    const int64_t fresh = round >= 0 ? (int64_t)round_start[round + 1] - fresh0 : 0;
    const int64_t total = fresh + (count_in != nullptr ? (int64_t)*count_in : 0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // whole warps iterate together (the pending append below is warp-aggregated)
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < total; base += stride) {
        const int64_t idx = base + lane;
        const bool active = idx < total;
        uint32_t i = 0;
        bool ready = true;
        if (active) {
            if (idx < fresh) i = members != nullptr ? members[fresh0 + idx] : (uint32_t)(fresh0 + idx);
            else i = pending_in[idx - fresh];
            const float4 me = packed[i];
            const uint64_t pi = prio[i];
            const double px = me.x, py = me.y;
            const int cx = cell_coord(px, g.x0, g.sx, g.ncx), cy = cell_coord(py, g.y0, g.sy, g.ncy);
            BoidAcc acc;
            for_neighbor_cells(g, cx, cy, [&](unsigned c) {
                if (!ready) return;
                const uint32_t e = cell_end[c];
                for (uint32_t j = cell_start[c]; j < e; ++j) {
                    if (j == i) continue;
                    const float4 o = packed[j];
                    float dx = fabsf(me.x - o.x), dy = fabsf(me.y - o.y);
                    if (g.torus) {
                        dx = fminf(dx, sp.fsx - dx);
                        dy = fminf(dy, sp.fsy - dy);
                    }
                    const float d2 = fmaf(dx, dx, dy * dy);
                    if (d2 > sp.relevant2) continue;           // cannot be within vision before or after its move
                    if (prio[j] < pi) {                        // acts before me: I see its NEW state
                        double2 sp_, sd_;
                        if (!load_new_state(newst, j, sp_, sd_)) {   // not computed yet: try again in the next pass
                            ready = false;
                            return;
                        }
                        boid_pair_d(g, bp, px, py, sp_.x, sp_.y, sd_.x, sd_.y, acc);
                    } else if (d2 <= sp.far2) {                // acts after me: its OLD state
                        boid_pair(g, bp, px, py, o, acc);
                    }
                }
            });
            if (ready) {
                const double4 s = boid_finish_d(g, bp, me, acc);
                // the move of this step (before the wrap) must not exceed `reach`, or a dependency was missed
                const double mvx = (s.z * bp.speed), mvy = (s.w * bp.speed);
                if (mvx * mvx + mvy * mvy > sp.reach2) atomicExch(error, 1u);
                store_new_state(newst, i, s);
                const uint32_t a = order[i];
                pos_out[a] = make_float2((float)s.x, (float)s.y);
                dir_out[a] = make_float2((float)s.z, (float)s.w);
                if (nbr_count != nullptr) nbr_count[a] = acc.k;
            }
        }
        const unsigned waiting = __ballot_sync(0xffffffffu, active && !ready);
        if (waiting) {
            unsigned slot = 0;
            if (lane == __ffs(waiting) - 1) slot = atomicAdd(count_out, (unsigned)__popc(waiting));
            slot = __shfl_sync(0xffffffffu, slot, __ffs(waiting) - 1);
            if (active && !ready) pending_out[slot + __popc(waiting & ((1u << lane) - 1u))] = i;
        }
    }
}

// ---- sequential activation in ONE launch: rank-ordered chained kernel -------------------------------------------
// The pass scheme above re-launches over the pending boids until the dependency chains are exhausted (~50 passes on
// a uniform flock).  The chained kernel instead activates the boids in RANK order (rank = position in the Philox
// activation order): thread r of the grid owns the boid of rank r, CTAs take their rank range from a ticket counter
// (so a CTA with a lower range has always started earlier), and a boid whose relevant earlier neighbour is not
// finished simply waits for it -- the neighbour has a lower rank, hence sits in a CTA that is already running, so
// the wait always ends (the pattern of a decoupled look-back scan).  Within a warp nobody blocks: the lanes loop,
// every iteration each unfinished lane re-tries, lanes that depend on a lane of the same warp see its result one
// iteration later.  `rank` holds the activation rank of every cell-sorted slot (read-only during the launch); new
// states are published through `newst` (load_new_state / store_new_state).
// sort keys: the high word of priority64 of every cell-sorted slot (the low word, the agent index, breaks ties)
__global__ void boids_seq_keys(const uint32_t* __restrict__ order, int64_t n, uint64_t seed, uint32_t epoch,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint32_t)(mb::priority64(seed, epoch, order != nullptr ? (uint64_t)order[i] : (uint64_t)i) >> 32);
    vals[i] = (uint32_t)i;
}

// runs of equal high words (rare) are ordered by agent index, like priority64 orders them
__global__ void boids_seq_fix_ties(const uint32_t* __restrict__ keys, uint32_t* __restrict__ slots,
                                   const uint32_t* __restrict__ order, int64_t n) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint32_t k = keys[r];
    if ((r > 0 && keys[r - 1] == k) || r + 1 >= n || keys[r + 1] != k) return;  // not the start of a run
    int64_t e = r + 1;
    while (e < n && keys[e] == k) ++e;
    auto agent_of = [&](uint32_t slot) { return order != nullptr ? order[slot] : slot; };
    for (int64_t a = r + 1; a < e; ++a) {  // insertion sort by agent index
        const uint32_t s = slots[a];
        const uint32_t ag = agent_of(s);
        int64_t b = a;
        while (b > r && agent_of(slots[b - 1]) > ag) {
            slots[b] = slots[b - 1];
            --b;
        }
        slots[b] = s;
    }
}

__global__ void boids_seq_ranks(const uint32_t* __restrict__ slots, int64_t n, uint32_t* __restrict__ rank) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) rank[slots[r]] = (uint32_t)r;
}

constexpr int kChainThreads = 128;

__global__ void __launch_bounds__(kChainThreads, 8)
boids_seq_chain(const float4* __restrict__ packed, const uint32_t* __restrict__ byrank,
                const uint32_t* __restrict__ rank, double2* newst, const uint32_t* __restrict__ order,
                const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end, int64_t n, Geometry g,
                BoidParams bp, SeqParams sp, uint32_t* __restrict__ ticket, float2* __restrict__ pos_out,
                float2* __restrict__ dir_out, uint32_t* __restrict__ nbr_count, uint32_t* __restrict__ error) {
    __shared__ uint32_t s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t r = (int64_t)s_ticket * kChainThreads + threadIdx.x;
    bool finished = r >= n;
    uint32_t i = 0, blocker = 0xffffffffu;
    float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0;
    if (!finished) {
        i = byrank[r];
        me = packed[i];
        cx = cell_coord((double)me.x, g.x0, g.sx, g.ncx);
        cy = cell_coord((double)me.y, g.y0, g.sy, g.ncy);
    }
    const uint32_t myrank = (uint32_t)r;
    unsigned spins = 0;
    while (!__all_sync(0xffffffffu, finished)) {
        bool waiting = false;
        if (!finished) {
            double2 np_, nd_;
            if (blocker != 0xffffffffu && !load_new_state(newst, blocker, np_, nd_)) {
                waiting = true;  // the neighbour that stopped the last attempt is still not finished
            } else {
                blocker = 0xffffffffu;
                const double px = me.x, py = me.y;
                BoidAcc acc;
                for_neighbor_cells(g, cx, cy, [&](unsigned c) {
                    if (blocker != 0xffffffffu) return;
                    const uint32_t e = cell_end[c];
                    for (uint32_t j0 = cell_start[c]; j0 < e; j0 += 4) {
                        // four candidates per batch: their loads are issued together (memory-level parallelism)
                        float4 o[4];
                        uint32_t rj[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t j = min(j0 + u, e - 1u);
                            o[u] = __ldg(packed + j);
                            rj[u] = __ldg(rank + j);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            // ONE float64 pair evaluation per candidate slot for all lanes, switched on by a predicate: with
                            // the evaluation inside the branches (irrelevant / new state / old state) each copy ran at ~3.5 of
                            // 32 lanes and was 65 % of the kernel's warp instructions (profiles/r2_boids_seq_chain_ncu.csv).
                            // It only moved the step from 12.3 to 12.1 ms at 10^7 boids: the kernel is bound by the memory
                            // traffic of rank-ordered access (21 GB of DRAM reads per step, 29 % L2 hits), not by issue.
                            const uint32_t j = j0 + u;
                            float dx = fabsf(me.x - o[u].x), dy = fabsf(me.y - o[u].y);
                            if (g.torus) {
                                dx = fminf(dx, sp.fsx - dx);
                                dy = fminf(dy, sp.fsy - dy);
                            }
                            const float d2 = fmaf(dx, dx, dy * dy);
                            // relevant: can be within vision before or after its move
                            const bool valid = blocker == 0xffffffffu && j < e && j != i && d2 <= sp.relevant2;
                            bool use_new = valid && rj[u] < myrank;                     // acts before me: I see its NEW state
                            const bool use_old = valid && !use_new && d2 <= sp.far2;    // acts after me: its OLD state
                            if (use_new && !load_new_state(newst, j, np_, nd_)) {
                                blocker = j;
                                use_new = false;
                            }
                            const double ox = use_new ? np_.x : (double)o[u].x, oy = use_new ? np_.y : (double)o[u].y;
                            const double odx = use_new ? nd_.x : (double)o[u].z, ody = use_new ? nd_.y : (double)o[u].w;
                            const double dd2 = ref_distance2(g, px, py, ox, oy);
                            const double ddx = ref_delta(ox, px, g.sx, g.torus), ddy = ref_delta(oy, py, g.sy, g.torus);
                            if ((use_new || use_old) && dd2 <= bp.vision2_le) {         // boid_pair_d, agents.py:66-85
                                acc.cohx += ddx; acc.cohy += ddy;
                                if (dd2 < bp.separation2_lt) { acc.sepx += ddx; acc.sepy += ddy; }
                                acc.mx += odx; acc.my += ody;
                                ++acc.k;
                            }
                        }
                        if (blocker != 0xffffffffu) return;
                    }
                });
                if (blocker == 0xffffffffu) {
                    const double4 s = boid_finish_d(g, bp, me, acc);
                    const double mvx = s.z * bp.speed, mvy = s.w * bp.speed;
                    if (mvx * mvx + mvy * mvy > sp.reach2) atomicExch(error, 1u);
                    store_new_state(newst, i, s);
                    const uint32_t a = order[i];
                    pos_out[a] = make_float2((float)s.x, (float)s.y);
                    dir_out[a] = make_float2((float)s.z, (float)s.w);
                    if (nbr_count != nullptr) nbr_count[a] = acc.k;
                    finished = true;
                } else {
                    waiting = true;
                }
            }
        }
        if (__any_sync(0xffffffffu, waiting)) {
            if (++spins > (1u << 22)) {  // seconds: a NaN state or an inconsistent rank table; do not hang the GPU
                if (!finished) atomicExch(error, 2u);
                break;
            }
            __nanosleep(spins < 8 ? 20 : 200);
        }
    }
}

// ---- radius queries (get_agents_in_radius for a batch of points) ---------------------------
// pass 0 (out_idx == nullptr): counts[q] = #{agents with dist <= radius} (minus `exclude[q]` if given)
// pass 1: fills out_idx/out_dist at offsets[q] .. in ascending agent index (storage order)
__global__ void radius_query(const float4* __restrict__ packed, const uint32_t* __restrict__ order,
                             const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ cell_end,
                             Geometry g, const float2* __restrict__ queries, const int32_t* __restrict__ exclude,
                             int64_t nq, double radius2_le, uint32_t* __restrict__ counts,
                             const int64_t* __restrict__ offsets, int32_t* __restrict__ out_idx,
                             double* __restrict__ out_dist) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 p = queries[q];
    const double px = p.x, py = p.y;
    const int cx = cell_coord(px, g.x0, g.sx, g.ncx), cy = cell_coord(py, g.y0, g.sy, g.ncy);
    const int32_t skip = exclude != nullptr ? exclude[q] : -1;
    unsigned k = 0;
    const int64_t off = offsets != nullptr ? offsets[q] : 0;
    for_neighbor_cells(g, cx, cy, [&](unsigned c) {
        const uint32_t e = cell_end[c];
        for (uint32_t j = cell_start[c]; j < e; ++j) {
            const int32_t a = (int32_t)order[j];
            if (a == skip) continue;
            const float4 o = packed[j];
            const double d2 = ref_distance2(g, px, py, (double)o.x, (double)o.y);
            if (d2 <= radius2_le) {
                if (out_idx != nullptr) {
                    const double d = __dsqrt_rn(d2);
                    // insertion keeps the row sorted by agent index (rows are short)
                    int64_t w = off + k;
                    while (w > off && out_idx[w - 1] > a) {
                        out_idx[w] = out_idx[w - 1];
                        out_dist[w] = out_dist[w - 1];
                        --w;
                    }
                    out_idx[w] = a;
                    out_dist[w] = d;
                }
                ++k;
            }
        }
    });
    if (counts != nullptr) counts[q] = k;
}

// ---- k nearest agents (get_k_nearest_agents for a batch of points, continuous_space.py:249-283) -------------
// One thread per query walks the cell list in Chebyshev RINGS of cells around the query's cell.  After ring r is
// done every unvisited agent is at least r * min(cell edge) away, so the search stops as soon as the k-th best
// distance is strictly below that bound (strictly: an unvisited agent at exactly the k-th distance could still win
// the tie on the lower storage index).  The candidate list lives in the output row, kept sorted by (distance,
// agent index) -- "in case of ties, the earlier an agent is inserted the higher it will rank" (:256-257).  On a
// torus the ring offsets are minimal images (each cell visited once); the walk ends when the rings cover the grid.
// Rows are padded with index -1 / distance +inf when fewer than k agents exist.
__global__ void __launch_bounds__(128)
knn_query(const float4* __restrict__ packed, const uint32_t* __restrict__ order, const uint32_t* __restrict__ cell_start,
          const uint32_t* __restrict__ cell_end, Geometry g, const float2* __restrict__ queries,
          const int32_t* __restrict__ exclude, int64_t nq, int k, int32_t* __restrict__ out_idx,
          double* __restrict__ out_dist) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float2 p = queries[q];
    const double px = p.x, py = p.y;
    const int cx = cell_coord(px, g.x0, g.sx, g.ncx), cy = cell_coord(py, g.y0, g.sy, g.ncy);
    const int32_t skip = exclude != nullptr ? exclude[q] : -1;
    int32_t* idx = out_idx + q * k;
    double* d2s = out_dist + q * k;     // squared distances while searching, square-rooted at the end
    int have = 0;
    // offsets a cell may take on each axis
    const int lox = g.torus ? -((g.ncx - 1) / 2) : -cx, hix = g.torus ? g.ncx / 2 : g.ncx - 1 - cx;
    const int loy = g.torus ? -((g.ncy - 1) / 2) : -cy, hiy = g.torus ? g.ncy / 2 : g.ncy - 1 - cy;
    int rmax = -lox;
    if (hix > rmax) rmax = hix;
    if (-loy > rmax) rmax = -loy;
    if (hiy > rmax) rmax = hiy;
    const double ex = g.sx / (double)g.ncx, ey = g.sy / (double)g.ncy;
    const double emin = ex < ey ? ex : ey;

    auto visit = [&](int ox, int oy) {
        int xx = cx + ox, yy = cy + oy;
        if (g.torus) {
            xx = (xx % g.ncx + g.ncx) % g.ncx;
            yy = (yy % g.ncy + g.ncy) % g.ncy;
        }
        const unsigned c = (unsigned)(yy * g.ncx + xx);
        const uint32_t e = cell_end[c];
        for (uint32_t j = cell_start[c]; j < e; ++j) {
            const int32_t a = (int32_t)order[j];
            if (a == skip) continue;
            const float4 o = packed[j];
            const double d2 = ref_distance2(g, px, py, (double)o.x, (double)o.y);
            if (have == k && !(d2 < d2s[k - 1] || (d2 == d2s[k - 1] && a < idx[k - 1]))) continue;
            int w = have < k ? have : k - 1;
            while (w > 0 && (d2s[w - 1] > d2 || (d2s[w - 1] == d2 && idx[w - 1] > a))) {
                d2s[w] = d2s[w - 1];
                idx[w] = idx[w - 1];
                --w;
            }
            d2s[w] = d2;
            idx[w] = a;
            if (have < k) ++have;
        }
    };

    for (int r = 0; r <= rmax; ++r) {
        if (r == 0) {
            visit(0, 0);
        } else {
            for (int oy = -r; oy <= r; ++oy) {
                if (oy < loy || oy > hiy) continue;
                if (oy == -r || oy == r) {
                    const int a0 = -r > lox ? -r : lox, a1 = r < hix ? r : hix;
                    for (int ox = a0; ox <= a1; ++ox) visit(ox, oy);
                } else {
                    if (-r >= lox) visit(-r, oy);
                    if (r <= hix) visit(r, oy);
                }
            }
        }
        if (have == k) {
            const double bound = (double)r * emin * (1.0 - 1e-9);   // slack for the rounding of cell_coord
            if (d2s[k - 1] < bound * bound) break;
        }
    }
    for (int w = 0; w < have; ++w) d2s[w] = __dsqrt_rn(d2s[w]);
    for (int w = have; w < k; ++w) {
        idx[w] = -1;
        d2s[w] = __longlong_as_double(0x7ff0000000000000ll);
    }
}

// ---- calculate_distances / calculate_difference_vector of ONE point against all or selected agents
//      (continuous_space.py:169-235): float64, the reference's operation order ----------------------------------------
__global__ void point_distances(const float2* __restrict__ pos, const int32_t* __restrict__ sel, int64_t m, Geometry g,
                                double px, double py, double* __restrict__ out_dist, double* __restrict__ out_delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float2 a = pos[sel != nullptr ? (int64_t)sel[i] : i];
    if (out_dist != nullptr) out_dist[i] = __dsqrt_rn(ref_distance2(g, px, py, (double)a.x, (double)a.y));
    if (out_delta != nullptr) {
        out_delta[2 * i] = ref_delta((double)a.x, px, g.sx, g.torus);
        out_delta[2 * i + 1] = ref_delta((double)a.y, py, g.sy, g.torus);
    }
}

// The same two queries in a space of any dimension (continuous_space.py:48-101 keeps a (capacity, ndims) array; the
// reference's distance loop runs "for i in range(1, self.ndims)", :229-231): pos is [n, ndim] float32, squares are
// added axis by axis in axis order -- also what scipy's Euclidean cdist does on the bounded path (:234).
constexpr int kMaxSpaceDims = 8;
struct NdQuery {
    double p[kMaxSpaceDims], size[kMaxSpaceDims];
    int ndim, torus;
};

__global__ void point_distances_nd(const float* __restrict__ pos, const int32_t* __restrict__ sel, int64_t m, NdQuery g,
                                   double* __restrict__ out_dist, double* __restrict__ out_delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float* a = pos + (sel != nullptr ? (int64_t)sel[i] : i) * g.ndim;
    double acc = 0.0;
    for (int d = 0; d < g.ndim; ++d) {
        const double q = (double)a[d];
        double del = fabs(g.p[d] - q);
        if (g.torus) del = fmin(del, g.size[d] - del);
        const double sq = __dmul_rn(del, del);
        acc = d == 0 ? sq : __dadd_rn(acc, sq);
        if (out_delta != nullptr) out_delta[i * g.ndim + d] = ref_delta(q, g.p[d], g.size[d], g.torus);
    }
    if (out_dist != nullptr) out_dist[i] = __dsqrt_rn(acc);
}

struct CellListWs {
    uint32_t *keys0, *vals0, *keys1, *vals1, *cell_start, *cell_end, *spare;
    float4* packed;
    void* sort_ws;
    size_t sort_ws_bytes;
};

size_t carve(CellListWs* ws, char* base, int64_t n, int64_t ncells) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    CellListWs t;
    t.keys0 = (uint32_t*)take(4 * n);
    t.vals0 = (uint32_t*)take(4 * n);
    t.keys1 = (uint32_t*)take(4 * n);
    t.vals1 = (uint32_t*)take(4 * n);
    t.cell_start = (uint32_t*)take(4 * (ncells + 2));  // counts -> exclusive prefix sums; [ncells] = n
    t.cell_end = t.cell_start + 1;                     // cell_end[c] = cell_start[c + 1]
    t.spare = (uint32_t*)take(4 * 4);                  // [0] overflow-tile counter, [1] long-cell counter
    t.packed = (float4*)take(16 * n);
    // scratch of the radix sort (sequential Boids, activation order) and of the prefix sum over the cell counters
    t.sort_ws_bytes = mb::sort_workspace_bytes(n);
    const size_t scan_bytes = 4 * mb::scan_scratch_words(ncells + 1) + 256;
    if (scan_bytes > t.sort_ws_bytes) t.sort_ws_bytes = scan_bytes;
    t.sort_ws = take(t.sort_ws_bytes);
    if (ws) *ws = t;
    return off;
}

// largest double y with sqrt(y) <= r  (IEEE sqrt is correctly rounded and monotone)
double sq_threshold_le(double r) {
    if (!(r >= 0.0)) return -1.0;
    double y = r * r;
    while (y > 0.0 && sqrt(y) > r) y = nextafter(y, 0.0);
    while (sqrt(nextafter(y, INFINITY)) <= r) y = nextafter(y, INFINITY);
    return y;
}
// smallest double y with sqrt(y) >= s, so that  sqrt(d2) < s  <=>  d2 < y
double sq_threshold_ge(double s) {
    if (!(s > 0.0)) return 0.0;
    double y = s * s;
    while (sqrt(y) < s) y = nextafter(y, INFINITY);
    while (y > 0.0 && sqrt(nextafter(y, 0.0)) >= s) y = nextafter(y, 0.0);
    return y;
}

int make_geometry(Geometry* g, double x0, double y0, double sx, double sy, double radius, int torus) {
    MB_REQUIRE(sx > 0.0 && sy > 0.0, "continuous space: size must be positive");
    MB_REQUIRE(radius > 0.0, "continuous space: radius must be positive");
    g->x0 = x0; g->y0 = y0; g->sx = sx; g->sy = sy; g->torus = torus;
    // cell edge strictly larger than the radius so rounding can never push a neighbour two cells away
    const double edge = radius * (1.0 + 1e-6);
    int64_t ncx = (int64_t)floor(sx / edge), ncy = (int64_t)floor(sy / edge);
    // cap the cell count: at most 2^26 cells (a few agents per cell is the useful regime)
    while (ncx * ncy > (1ll << 26)) { ncx = (ncx + 1) / 2; ncy = (ncy + 1) / 2; }
    g->ncx = (int)(ncx < 1 ? 1 : ncx);
    g->ncy = (int)(ncy < 1 ? 1 : ncy);
    return MB_OK;
}

// bins the agents; leaves order = vals0 (sorted slot -> agent), cell_start / cell_end and the packed state in ws
int build_cell_list(const float2* pos, const float2* dir, int64_t n, const Geometry& g, CellListWs& ws,
                    uint32_t** order_out, cudaStream_t stream) {
    const int64_t ncells = (int64_t)g.ncx * g.ncy;
    int rc = mb::check_cuda(cudaMemsetAsync(ws.cell_start, 0, 4 * (ncells + 2), stream), "memset");
    if (rc) return rc;
    rc = mb::check_cuda(cudaMemsetAsync(ws.spare, 0, 4 * 4, stream), "memset");
    if (rc) return rc;
    *order_out = ws.vals0;
    if (n == 0) return MB_OK;
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    count_cells<<<blocks, 256, 0, stream>>>(pos, n, g, ws.cell_start, ws.keys0, ws.keys1);
    if (mb::exclusive_scan_u32(ws.cell_start, ncells + 1, reinterpret_cast<uint32_t*>(ws.sort_ws), stream) != MB_OK) return MB_ERR_CUDA;
    scatter_cells<<<blocks, 256, 0, stream>>>(ws.keys0, ws.keys1, n, ws.cell_start, ws.vals0);
    // keys0 / keys1 are free again: long-cell queue and scratch
    sort_cells<<<(unsigned)mb::ceil_div(ncells, 128), 128, 0, stream>>>(ws.vals0, ws.cell_start, ncells, ws.keys0, ws.spare + 1);
    sort_long_cells<<<mb::sm_count(), 256, 0, stream>>>(ws.vals0, ws.keys1, ws.cell_start, ws.keys0, ws.spare + 1);
    gather_state<<<blocks, 256, 0, stream>>>(pos, dir, ws.vals0, n, ws.packed);
    MB_LAUNCH_CHECK("build_cell_list");
    return MB_OK;
}

}  // namespace

MB_API int mb_celllist_workspace_bytes(int64_t n, double size_x, double size_y, double radius, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0, "mb_celllist_workspace_bytes: bad argument");
    Geometry g;
    int rc = make_geometry(&g, 0.0, 0.0, size_x, size_y, radius, 1);
    if (rc) return rc;
    *out = carve(nullptr, nullptr, n, (int64_t)g.ncx * g.ncy);
    return MB_OK;
}

MB_API int mb_boids_step_f32(const float* pos_in, const float* dir_in, float* pos_out, float* dir_out, int64_t n,
                             double x0, double y0, double size_x, double size_y, int torus, double vision,
                             double separation, double cohere, double separate, double match, double speed,
                             uint32_t* neighbor_count, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_boids_step_f32: n out of range");
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos_in && dir_in && pos_out && dir_out && workspace, "mb_boids_step_f32: null buffer");
    MB_REQUIRE(pos_in != pos_out && dir_in != dir_out, "mb_boids_step_f32: in/out must not alias (synchronous step)");
    Geometry g;
    int rc = make_geometry(&g, x0, y0, size_x, size_y, vision, torus);
    if (rc) return rc;
    CellListWs ws;
    if (carve(&ws, (char*)workspace, n, (int64_t)g.ncx * g.ncy) > workspace_bytes) {
        mb::set_error("mb_boids_step_f32: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    uint32_t* order = nullptr;
    rc = build_cell_list((const float2*)pos_in, (const float2*)dir_in, n, g, ws, &order, stream);
    if (rc) return rc;
    const BoidParams bp{sq_threshold_le(vision), sq_threshold_ge(separation), cohere, separate, match, speed};
    static const bool force_generic = getenv("MB_BOIDS_GENERIC") != nullptr;
    if (!force_generic && g.ncx >= kHX && g.ncy >= kHY) {
        // tiled path; keys0 is free after the sort when the order ended up in the other ping-pong buffer,
        // keys1 otherwise: use whichever key buffer does not alias `order`'s partner for the overflow list
        const int tiles_x = (int)mb::ceil_div(g.ncx, kTX), tiles_y = (int)mb::ceil_div(g.ncy, kTY);
        uint32_t* ovf_tiles = ws.vals1;    // n entries >= tiles (checked below); free: the order lives in vals0
        uint32_t* ovf_count = ws.spare;
        if ((int64_t)tiles_x * tiles_y <= n) {
            static bool attr_set = false;
            if (!attr_set) {
                rc = mb::check_cuda(cudaFuncSetAttribute(boids_update_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)sizeof(TileSmem)), "cudaFuncSetAttribute");
                if (rc) return rc;
                attr_set = true;
            }
            // float32 classification bands (see the kernel comment); d2 in float32 is within 3e-7 relative of the
            // reference's float64 value, the bands are +-1e-6.  Seam: a pair whose |dx| exceeds size - vision(1+)
            // may be a neighbour through the torus wrap.
            const double band = 1e-6;
            FastParams fp;
            fp.vis_lo2 = (float)(bp.vision2_le * (1.0 - band));
            fp.vis_hi2 = (float)(bp.vision2_le * (1.0 + band));
            fp.sep_lo2 = (float)(bp.separation2_lt * (1.0 - band));
            fp.sep_hi2 = (float)(bp.separation2_lt * (1.0 + band));
            fp.seam_x = (float)((g.sx - vision * 1.001) * (1.0 - band));
            fp.seam_y = (float)((g.sy - vision * 1.001) * (1.0 - band));
            boids_update_tiled<<<(unsigned)(tiles_x * tiles_y), kTileThreads, sizeof(TileSmem), stream>>>(
                ws.packed, order, ws.cell_start, ws.cell_end, g, bp, fp, tiles_x, (float2*)pos_out,
                (float2*)dir_out, neighbor_count, ovf_tiles, ovf_count);
            boids_update_overflow<<<mb::sm_count(), 128, 0, stream>>>(ws.packed, order, ws.cell_start, ws.cell_end, g, bp,
                                                                    tiles_x, ovf_tiles, ovf_count, (float2*)pos_out,
                                                                    (float2*)dir_out, neighbor_count);
            MB_LAUNCH_CHECK("mb_boids_step_f32");
            return MB_OK;
        }
    }
    boids_update<<<(unsigned)mb::ceil_div(n, 128), 128, 0, stream>>>(ws.packed, order, ws.cell_start, ws.cell_end, n, g, bp,
                                                                     (float2*)pos_out, (float2*)dir_out, neighbor_count);
    MB_LAUNCH_CHECK("mb_boids_step_f32");
    return MB_OK;
}

// ---- spatial re-sort of the storage order ------------------------------------------------------------------------
// The step kernels gather every boid's state by its cell-list position and scatter the results back to storage order;
// with a storage order that has nothing to do with space (agents created at random positions) those are 10^7 random
// 8-byte accesses each (gather_state 0.25 ms, 2.6x write amplification in the update kernel: profiles/r1_boids_*).
// mb_boids_resort_f32 rewrites the population in cell-list order -- row i of the outputs is the boid at sorted slot i,
// perm_out[i] its previous row -- so that a model which re-sorts every few steps (boids move at most `speed` per step,
// a cell edge is >= vision) keeps storage order and cell order nearly identical and all of those accesses coalesce.
// unique ids travel with the rows (ids_in NULL: the previous row numbers are the ids).
namespace {
__global__ void split_sorted(const float4* __restrict__ packed, const uint32_t* __restrict__ order,
                             const int64_t* __restrict__ ids_in, int64_t n, float2* __restrict__ pos_out,
                             float2* __restrict__ dir_out, int64_t* __restrict__ ids_out, uint32_t* __restrict__ perm_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = packed[i];
    const uint32_t a = order[i];
    pos_out[i] = make_float2(s.x, s.y);
    dir_out[i] = make_float2(s.z, s.w);
    if (ids_out != nullptr) ids_out[i] = ids_in != nullptr ? ids_in[a] : (int64_t)a;
    if (perm_out != nullptr) perm_out[i] = a;
}
}  // namespace

MB_API int mb_boids_resort_f32(const float* pos_in, const float* dir_in, const int64_t* ids_in, float* pos_out, float* dir_out,
                               int64_t* ids_out, uint32_t* perm_out, int64_t n, double x0, double y0, double size_x,
                               double size_y, int torus, double vision, void* workspace, size_t workspace_bytes,
                               cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_boids_resort_f32: n out of range");
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos_in && dir_in && pos_out && dir_out && workspace, "mb_boids_resort_f32: null buffer");
    MB_REQUIRE(pos_in != pos_out && dir_in != dir_out && (ids_in == nullptr || ids_in != ids_out),
               "mb_boids_resort_f32: in / out must not alias");
    Geometry g;
    int rc = make_geometry(&g, x0, y0, size_x, size_y, vision, torus);
    if (rc) return rc;
    CellListWs ws;
    if (carve(&ws, (char*)workspace, n, (int64_t)g.ncx * g.ncy) > workspace_bytes) {
        mb::set_error("mb_boids_resort_f32: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    uint32_t* order = nullptr;
    rc = build_cell_list((const float2*)pos_in, (const float2*)dir_in, n, g, ws, &order, stream);
    if (rc) return rc;
    split_sorted<<<(unsigned)mb::ceil_div(n, 256), 256, 0, stream>>>(ws.packed, order, ids_in, n, (float2*)pos_out, (float2*)dir_out,
                                                                    ids_out, perm_out);
    MB_LAUNCH_CHECK("mb_boids_resort_f32");
    return MB_OK;
}

// ---- sequential Boids step (shuffle_do semantics, see boids_seq_pass) -----------------------------------------
namespace {

constexpr int kSeqMaxPasses = 4096;
constexpr int kSeqMaxRounds = 256;

struct SeqWs {
    CellListWs cl;
    uint64_t* prio;
    double2* newst;
    uint32_t *pend0, *pend1, *counts, *error;
    uint32_t *pk0, *pk1, *pv0, *pv1, *rank, *ticket;   // priority / round sort ping-pong, rank table (chained kernel)
    uint32_t* round_start;                             // [kSeqMaxRounds + 1]
};

size_t carve_seq(SeqWs* ws, char* base, int64_t n, int64_t ncells) {
    size_t off = carve(ws ? &ws->cl : nullptr, base, n, ncells);
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += (bytes + 255) & ~(size_t)255;
        return p;
    };
    SeqWs t;
    if (ws) t.cl = ws->cl;
    t.prio = (uint64_t*)take(8 * n);
    t.newst = (double2*)take(32 * n);
    t.pend0 = (uint32_t*)take(4 * n);
    t.pend1 = (uint32_t*)take(4 * n);
    t.counts = (uint32_t*)take(4 * (kSeqMaxPasses + 1));
    t.error = (uint32_t*)take(4);
    t.pk0 = (uint32_t*)take(4 * n);
    t.pk1 = (uint32_t*)take(4 * n);
    t.pv0 = (uint32_t*)take(4 * n);
    t.pv1 = (uint32_t*)take(4 * n);
    t.rank = (uint32_t*)take(4 * n);
    t.ticket = (uint32_t*)take(4);
    t.round_start = (uint32_t*)take(4 * (kSeqMaxRounds + 1));
    if (ws) *ws = t;
    return off;
}

}  // namespace

MB_API int mb_boids_seq_workspace_bytes(int64_t n, double size_x, double size_y, double vision, double reach,
                                        size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0 && reach >= 0.0, "mb_boids_seq_workspace_bytes: bad argument");
    Geometry g;
    int rc = make_geometry(&g, 0.0, 0.0, size_x, size_y, vision + reach, 1);
    if (rc) return rc;
    *out = carve_seq(nullptr, nullptr, n, (int64_t)g.ncx * g.ncy);
    return MB_OK;
}

// One Boid step of all agents in the reference's SEQUENTIAL activation order (ascending Philox priority of
// (seed, epoch, agent)): identical to model.agents.shuffle_do("step") driven with that order.  `reach` must bound the
// distance any boid moves in one step (speed * max(1, max |direction|)); a longer move returns MB_ERR_INVALID.
// Synchronises `stream` (the host reads the number of pending boids between batches of passes).
MB_API int mb_boids_step_seq_f32(const float* pos_in, const float* dir_in, float* pos_out, float* dir_out, int64_t n,
                                 double x0, double y0, double size_x, double size_y, int torus, double vision,
                                 double separation, double cohere, double separate, double match, double speed,
                                 double reach, uint64_t seed, uint32_t epoch, uint32_t* neighbor_count,
                                 void* workspace, size_t workspace_bytes, int* out_passes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_boids_step_seq_f32: n out of range");
    if (out_passes) *out_passes = 0;
    if (n == 0) return MB_OK;
    MB_REQUIRE(pos_in && dir_in && pos_out && dir_out && workspace, "mb_boids_step_seq_f32: null buffer");
    MB_REQUIRE(pos_in != pos_out && dir_in != dir_out, "mb_boids_step_seq_f32: in/out must not alias");
    MB_REQUIRE(reach >= 0.0, "mb_boids_step_seq_f32: reach must be >= 0");
    Geometry g;
    int rc = make_geometry(&g, x0, y0, size_x, size_y, vision + reach, torus);
    if (rc) return rc;
    SeqWs ws;
    if (carve_seq(&ws, (char*)workspace, n, (int64_t)g.ncx * g.ncy) > workspace_bytes) {
        mb::set_error("mb_boids_step_seq_f32: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    uint32_t* order = nullptr;
    rc = build_cell_list((const float2*)pos_in, (const float2*)dir_in, n, g, ws.cl, &order, stream);
    if (rc) return rc;
    rc = mb::check_cuda(cudaMemsetAsync(ws.newst, 0xFF, 32 * n, stream), "memset");  // all NaN: nobody has finished
    if (rc) return rc;
    rc = mb::check_cuda(cudaMemsetAsync(ws.counts, 0, 4 * (kSeqMaxPasses + 1), stream), "memset");
    if (rc) return rc;
    rc = mb::check_cuda(cudaMemsetAsync(ws.error, 0, 4, stream), "memset");
    if (rc) return rc;
    const BoidParams bp{sq_threshold_le(vision), sq_threshold_ge(separation), cohere, separate, match, speed};
    // float32 pre-filters (conservative): d2 of the torus minimum image in float32 is off by at most ~ulp(size) per
    // component, i.e. a relative error of about 64 eps size / radius on d2; beyond relevant2 two boids cannot see each
    // other before or after a move of `reach`, beyond far2 an unmoved boid is surely outside `vision`
    const double smax = size_x > size_y ? size_x : size_y;
    const double margin = 64.0 * 1.1920929e-7 * (smax / vision) + 1e-5;
    SeqParams sp;
    const double rel = vision + reach;
    sp.relevant2 = margin < 0.5 ? (float)(rel * rel * (1.0 + margin)) : INFINITY;
    sp.far2 = margin < 0.5 ? (float)(vision * vision * (1.0 + margin)) : INFINITY;
    sp.reach2 = reach * reach * (1.0 + 1e-9) * (1.0 + 1e-9);
    sp.fsx = (float)size_x;
    sp.fsy = (float)size_y;
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    const char* mode = getenv("MB_BOIDS_SEQ");
    if (mode == nullptr || (strcmp(mode, "passes") != 0 && strcmp(mode, "rounds") != 0)) {
        // rank-ordered chained kernel: sort the slots by priority, then one launch
        boids_seq_keys<<<blocks, 256, 0, stream>>>(order, n, seed, epoch, ws.pk0, ws.pv0);
        const int where = mb::radix_sort_pairs<uint32_t>(ws.pk0, ws.pv0, ws.pk1, ws.pv1, n, 0, 32, ws.cl.sort_ws,
                                                         ws.cl.sort_ws_bytes, stream);
        if (where < 0) return where;
        uint32_t* keys = where ? ws.pk1 : ws.pk0;
        uint32_t* byrank = where ? ws.pv1 : ws.pv0;
        boids_seq_fix_ties<<<blocks, 256, 0, stream>>>(keys, byrank, order, n);
        boids_seq_ranks<<<blocks, 256, 0, stream>>>(byrank, n, ws.rank);
        rc = mb::check_cuda(cudaMemsetAsync(ws.ticket, 0, 4, stream), "memset");
        if (rc) return rc;
        boids_seq_chain<<<(unsigned)mb::ceil_div(n, kChainThreads), kChainThreads, 0, stream>>>(
            ws.cl.packed, byrank, ws.rank, ws.newst, order, ws.cl.cell_start, ws.cl.cell_end, n, g, bp, sp, ws.ticket,
            (float2*)pos_out, (float2*)dir_out, neighbor_count, ws.error);
        MB_LAUNCH_CHECK("mb_boids_step_seq_f32");
        uint32_t err = 0u;
        rc = mb::check_cuda(cudaMemcpyAsync(&err, ws.error, 4, cudaMemcpyDeviceToHost, stream), "read error");
        if (rc) return rc;
        rc = mb::check_cuda(cudaStreamSynchronize(stream), "mb_boids_step_seq_f32");
        if (rc) return rc;
        MB_REQUIRE(err != 1u, "mb_boids_step_seq_f32: a boid moved farther than `reach` in one step");
        if (err != 0u) {
            mb::set_error("mb_boids_step_seq_f32: the chained kernel timed out waiting for a dependency");
            return MB_ERR_CUDA;
        }
        if (out_passes) *out_passes = 1;
        return MB_OK;
    }
    // multi-pass schemes (MB_BOIDS_SEQ=rounds / passes; kept for cross-checking the chained kernel, 10^7 boids: chained
    // 19 ms, rounds 23 ms, passes 66 ms): the boids are split into `rounds` slices of the activation order; each round
    // is offered `subpasses` passes (the fresh members, then retries of whatever is still pending, which also rides
    // along with the following rounds); `passes` is the one-round special case (every boid fresh in pass 0)
    boids_seq_priorities<<<blocks, 256, 0, stream>>>(order, n, seed, epoch, ws.prio);
    int rounds = n < 65536 ? 1 : 16, subpasses = 2;  // small flocks: a pass is cheaper than a launch
    if (mode != nullptr && strcmp(mode, "passes") == 0) rounds = 1;
    else if (const char* e = getenv("MB_BOIDS_SEQ_ROUNDS")) rounds = atoi(e);
    if (const char* e = getenv("MB_BOIDS_SEQ_SUBPASSES")) subpasses = atoi(e);
    int log_rounds = 0;
    while ((1 << log_rounds) < rounds) ++log_rounds;
    rounds = 1 << log_rounds;
    MB_REQUIRE(rounds >= 1 && rounds <= kSeqMaxRounds && subpasses >= 1 && subpasses <= 8,
               "mb_boids_step_seq_f32: bad MB_BOIDS_SEQ_ROUNDS / MB_BOIDS_SEQ_SUBPASSES");
    const uint32_t* members = nullptr;
    if (rounds > 1) {
        boids_seq_round_keys<<<blocks, 256, 0, stream>>>(ws.prio, n, 64 - log_rounds, ws.pk0, ws.pv0);
        const int where = mb::radix_sort_pairs<uint32_t>(ws.pk0, ws.pv0, ws.pk1, ws.pv1, n, 0, log_rounds > 8 ? 16 : 8,
                                                         ws.cl.sort_ws, ws.cl.sort_ws_bytes, stream);
        if (where < 0) return where;
        members = where ? ws.pv1 : ws.pv0;
        boids_seq_round_starts<<<blocks, 256, 0, stream>>>(where ? ws.pk1 : ws.pk0, n, rounds, ws.round_start);
    } else {
        const uint32_t host_start[2] = {0u, (uint32_t)n};
        rc = mb::check_cuda(cudaMemcpyAsync(ws.round_start, host_start, 8, cudaMemcpyHostToDevice, stream), "round table");
        if (rc) return rc;
    }
    const int grid_cap = mb::sm_count() * 16;
    auto grid_for = [&](int64_t items) {
        const int64_t want = mb::ceil_div(items < 1 ? 1 : items, 128);
        return (unsigned)(want > grid_cap ? grid_cap : want);
    };
    int pass = 0;  // passes launched so far; pass p reads counts[p-1] / pend[(p-1)&1] and writes counts[p] / pend[p&1]
    auto launch_pass = [&](int round, int64_t items) {
        uint32_t* pin = ((pass - 1) & 1) ? ws.pend1 : ws.pend0;
        uint32_t* pout = (pass & 1) ? ws.pend1 : ws.pend0;
        boids_seq_pass<<<grid_for(items), 128, 0, stream>>>(
            ws.cl.packed, ws.prio, ws.newst, order, ws.cl.cell_start, ws.cl.cell_end, n, g, bp, sp, members, ws.round_start,
            round, pass > 0 ? pin : nullptr, pass > 0 ? ws.counts + (pass - 1) : nullptr, pout, ws.counts + pass,
            (float2*)pos_out, (float2*)dir_out, neighbor_count, ws.error);
        ++pass;
    };
    const int64_t per_round = mb::ceil_div(n, rounds);
    for (int k = 0; k < rounds; ++k)
        for (int sub = 0; sub < subpasses; ++sub) launch_pass(sub == 0 ? k : -1, sub == 0 ? 2 * per_round : per_round / 2);
    MB_LAUNCH_CHECK("mb_boids_step_seq_f32");
    uint32_t pending = (uint32_t)n;
    while (true) {  // drain: whatever is still pending after the scheduled passes
        uint32_t host[2] = {0u, 0u};
        rc = mb::check_cuda(cudaMemcpyAsync(&host[0], ws.counts + (pass - 1), 4, cudaMemcpyDeviceToHost, stream), "read pending");
        if (rc) return rc;
        rc = mb::check_cuda(cudaMemcpyAsync(&host[1], ws.error, 4, cudaMemcpyDeviceToHost, stream), "read error");
        if (rc) return rc;
        rc = mb::check_cuda(cudaStreamSynchronize(stream), "mb_boids_step_seq_f32");
        if (rc) return rc;
        MB_REQUIRE(host[1] == 0u, "mb_boids_step_seq_f32: a boid moved farther than `reach` in one step");
        pending = host[0];
        if (pending == 0u) break;
        if (pass + 8 >= kSeqMaxPasses) {
            mb::set_error("mb_boids_step_seq_f32: dependency chains longer than %d passes", kSeqMaxPasses);
            return MB_ERR_UNSUPPORTED;
        }
        for (int b = 0; b < 8; ++b) launch_pass(-1, pending);
        MB_LAUNCH_CHECK("mb_boids_step_seq_f32");
    }
    if (out_passes) *out_passes = pass;
    return MB_OK;
}

// The activation order of AgentSet.shuffle / shuffle_do number `epoch` (agentset.py:415-437, 772-787) as an explicit
// permutation: order[r] = index of the agent activated r-th = agents sorted by priority64(seed, epoch, agent)
// (csrc/philox.cuh).  Same kernels as the sequential Boids step: radix sort on the high word, ties by agent index.
// workspace: mb_activation_order_workspace_bytes(n).
MB_API int mb_activation_order_workspace_bytes(int64_t n, size_t* out) {
    MB_REQUIRE(out != nullptr && n >= 0, "mb_activation_order_workspace_bytes: bad argument");
    *out = 4 * ((size_t)(4 * n + 255) & ~(size_t)255) + mb::sort_workspace_bytes(n) + 1024;
    return MB_OK;
}

MB_API int mb_activation_order(uint64_t seed, uint32_t epoch, int64_t n, uint32_t* order_out, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31), "mb_activation_order: n out of range");
    if (n == 0) return MB_OK;
    MB_REQUIRE(order_out && workspace, "mb_activation_order: null buffer");
    size_t need = 0;
    mb_activation_order_workspace_bytes(n, &need);
    if (workspace_bytes < need) {
        mb::set_error("mb_activation_order: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    const size_t stride = ((size_t)(4 * n + 255) & ~(size_t)255);
    char* base = (char*)workspace;
    uint32_t *k0 = (uint32_t*)base, *k1 = (uint32_t*)(base + stride), *v0 = (uint32_t*)(base + 2 * stride),
             *v1 = (uint32_t*)(base + 3 * stride);
    void* sort_ws = base + 4 * stride;
    const unsigned blocks = (unsigned)mb::ceil_div(n, 256);
    // identity "order": slot i holds agent i.  v0 doubles as that identity table (vals = slot = agent).
    boids_seq_keys<<<blocks, 256, 0, stream>>>(nullptr, n, seed, epoch, k0, v0);
    const int where = mb::radix_sort_pairs<uint32_t>(k0, v0, k1, v1, n, 0, 32, sort_ws, mb::sort_workspace_bytes(n), stream);
    if (where < 0) return where;
    uint32_t* keys = where ? k1 : k0;
    uint32_t* vals = where ? v1 : v0;
    boids_seq_fix_ties<<<blocks, 256, 0, stream>>>(keys, vals, nullptr, n);
    MB_LAUNCH_CHECK("mb_activation_order");
    return mb::check_cuda(cudaMemcpyAsync(order_out, vals, 4 * n, cudaMemcpyDeviceToDevice, stream), "mb_activation_order");
}

// Batched get_agents_in_radius.  Call with out_idx == NULL to get counts[nq]; prefix-sum them into
// offsets[nq] (int64) and call again with out_idx/out_dist to fill the CSR rows (ascending agent index,
// the reference's storage order).  exclude[q] >= 0 drops that agent (get_neighbors_in_radius).
MB_API int mb_radius_query_f32(const float* pos, int64_t n, double x0, double y0, double size_x, double size_y,
                               int torus, const float* queries, const int32_t* exclude, int64_t nq, double radius,
                               uint32_t* counts, const int64_t* offsets, int32_t* out_idx, double* out_dist,
                               void* workspace, size_t workspace_bytes, int reuse_cell_list, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31) && nq >= 0, "mb_radius_query_f32: size out of range");
    if (nq == 0) return MB_OK;
    MB_REQUIRE(queries && workspace && (n == 0 || pos), "mb_radius_query_f32: null buffer");
    MB_REQUIRE((out_idx == nullptr) == (out_dist == nullptr), "mb_radius_query_f32: out_idx and out_dist go together");
    MB_REQUIRE(out_idx == nullptr || offsets != nullptr, "mb_radius_query_f32: fill pass needs offsets");
    Geometry g;
    int rc = make_geometry(&g, x0, y0, size_x, size_y, radius, torus);
    if (rc) return rc;
    CellListWs ws;
    if (carve(&ws, (char*)workspace, n, (int64_t)g.ncx * g.ncy) > workspace_bytes) {
        mb::set_error("mb_radius_query_f32: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    uint32_t* order = ws.vals0;
    if (!reuse_cell_list) {
        rc = build_cell_list((const float2*)pos, nullptr, n, g, ws, &order, stream);
        if (rc) return rc;
        // keep the order in vals0 so a later call with reuse_cell_list finds it
        if (order != ws.vals0) {
            rc = mb::check_cuda(cudaMemcpyAsync(ws.vals0, order, 4 * n, cudaMemcpyDeviceToDevice, stream), "copy order");
            if (rc) return rc;
            order = ws.vals0;
        }
    }
    radius_query<<<(unsigned)mb::ceil_div(nq, 128), 128, 0, stream>>>(ws.packed, order, ws.cell_start, ws.cell_end, g,
                                                                      (const float2*)queries, exclude, nq, sq_threshold_le(radius), counts,
                                                                      offsets, out_idx, out_dist);
    MB_LAUNCH_CHECK("mb_radius_query_f32");
    return MB_OK;
}

// Batched get_k_nearest_agents: out_idx / out_dist are [nq, k] (ascending (distance, agent index); -1 / +inf padding).
// `cell_edge` sizes the cell list (a few agents per cell is the useful regime); the workspace is
// mb_celllist_workspace_bytes(n, size_x, size_y, cell_edge).
MB_API int mb_knn_query_f32(const float* pos, int64_t n, double x0, double y0, double size_x, double size_y, int torus,
                            const float* queries, const int32_t* exclude, int64_t nq, int k, double cell_edge,
                            int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes,
                            int reuse_cell_list, cudaStream_t stream) {
    MB_REQUIRE(n >= 0 && n < (1ll << 31) && nq >= 0 && k >= 0, "mb_knn_query_f32: size out of range");
    if (nq == 0 || k == 0) return MB_OK;
    MB_REQUIRE(queries && workspace && out_idx && out_dist && (n == 0 || pos), "mb_knn_query_f32: null buffer");
    Geometry g;
    int rc = make_geometry(&g, x0, y0, size_x, size_y, cell_edge, torus);
    if (rc) return rc;
    CellListWs ws;
    if (carve(&ws, (char*)workspace, n, (int64_t)g.ncx * g.ncy) > workspace_bytes) {
        mb::set_error("mb_knn_query_f32: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    uint32_t* order = ws.vals0;
    if (!reuse_cell_list) {
        rc = build_cell_list((const float2*)pos, nullptr, n, g, ws, &order, stream);
        if (rc) return rc;
    }
    knn_query<<<(unsigned)mb::ceil_div(nq, 128), 128, 0, stream>>>(ws.packed, order, ws.cell_start, ws.cell_end, g,
                                                                   (const float2*)queries, exclude, nq, k, out_idx, out_dist);
    MB_LAUNCH_CHECK("mb_knn_query_f32");
    return MB_OK;
}

// calculate_distances (out_dist, [m]) and / or calculate_difference_vector (out_delta, [m, 2]) of the point
// (px, py) against agents sel[0..m) (sel == NULL: agents 0..m).  float64 results.
MB_API int mb_point_distances_f32(const float* pos, const int32_t* sel, int64_t m, double size_x, double size_y,
                                  int torus, double px, double py, double* out_dist, double* out_delta,
                                  cudaStream_t stream) {
    MB_REQUIRE(m >= 0, "mb_point_distances_f32: negative size");
    if (m == 0) return MB_OK;
    MB_REQUIRE(pos && (out_dist || out_delta), "mb_point_distances_f32: null buffer");
    Geometry g;
    g.x0 = 0.0; g.y0 = 0.0; g.sx = size_x; g.sy = size_y; g.ncx = 1; g.ncy = 1; g.torus = torus;
    point_distances<<<(unsigned)mb::ceil_div(m, 256), 256, 0, stream>>>((const float2*)pos, sel, m, g, px, py, out_dist, out_delta);
    MB_LAUNCH_CHECK("mb_point_distances_f32");
    return MB_OK;
}

// n-dimensional form of mb_point_distances_f32: pos [n, ndim] float32, size_host / point_host [ndim] doubles in HOST
// memory, out_dist [m], out_delta [m, ndim] float64 (either may be NULL).  1 <= ndim <= 8.
MB_API int mb_point_distances_nd_f32(const float* pos, int ndim, const int32_t* sel, int64_t m, const double* size_host,
                                     int torus, const double* point_host, double* out_dist, double* out_delta,
                                     cudaStream_t stream) {
    MB_REQUIRE(m >= 0, "mb_point_distances_nd_f32: negative size");
    MB_REQUIRE(ndim >= 1 && ndim <= kMaxSpaceDims, "mb_point_distances_nd_f32: 1 <= ndim <= %d", kMaxSpaceDims);
    MB_REQUIRE(size_host != nullptr && point_host != nullptr, "mb_point_distances_nd_f32: null host argument");
    if (m == 0) return MB_OK;
    MB_REQUIRE(pos && (out_dist || out_delta), "mb_point_distances_nd_f32: null buffer");
    NdQuery g;
    for (int d = 0; d < kMaxSpaceDims; ++d) {
        g.p[d] = d < ndim ? point_host[d] : 0.0;
        g.size[d] = d < ndim ? size_host[d] : 1.0;
    }
    g.ndim = ndim;
    g.torus = torus;
    point_distances_nd<<<(unsigned)mb::ceil_div(m, 256), 256, 0, stream>>>(pos, sel, m, g, out_dist, out_delta);
    MB_LAUNCH_CHECK("mb_point_distances_nd_f32");
    return MB_OK;
}
