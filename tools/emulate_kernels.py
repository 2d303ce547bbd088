"""CPU emulation of the index arithmetic of kernels that need no inter-thread communication.

The column-walker kernels of csrc/layers.cu (k_nsum_walk, k_diffuse_walk) and the ring search of
csrc/continuous.cu (knn_query) are extracted TEXTUALLY from the .cu sources, compiled with g++ against
tools/emu/cuda_shim.h (threads run one after the other) and checked against brute force on ragged shapes.
This is a development aid for a container without a GPU -- the parity tests proper run on the device
(tests/test_layers_gpu.py, tests/test_continuous_gpu.py).

    python tools/emulate_kernels.py          # exit code 0 = all cases agree
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mesa_b200", "csrc")


def between(text: str, start: str, end: str) -> str:
    a = text.index(start)
    b = text.index(end, a)
    return text[a:b]


HARNESS = r'''
#include "cuda_shim.h"
enum { DT_U8 = 0 };
%(layers)s
%(continuous)s
%(agents)s

// stable compaction of csrc/agents.cu: flags -> (host) exclusive scan -> grid-stride row scatter, rows of any width
template <typename U>
static int check_compact(int64_t n, int upr, int blocks, int pattern) {
    std::vector<uint8_t> keep(n);
    for (int64_t i = 0; i < n; ++i)
        keep[i] = pattern == 0 ? 0 : pattern == 1 ? 1 : pattern == 2 ? (uint8_t)(rand() %% 3 == 0) * 255 : (uint8_t)((i / 37) %% 2);
    std::vector<uint32_t> offs(n + 1, 77u);
    blockDim.x = 256;
    gridDim.x = (unsigned)((n + 1 + 255) / 256);
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
        for (threadIdx.x = 0; threadIdx.x < 256; ++threadIdx.x) flags_to_words(keep.data(), offs.data(), n);
    uint32_t run = 0;
    for (int64_t i = 0; i <= n; ++i) { const uint32_t f = offs[i]; if (f > 1) return 1; offs[i] = run; run += f; }
    const int64_t kept = offs[n];
    std::vector<U> in(n * upr), got(kept * upr), want;
    unsigned char* raw = (unsigned char*)in.data();
    for (size_t b = 0; b < in.size() * sizeof(U); ++b) raw[b] = (unsigned char)rand();
    for (int64_t i = 0; i < n; ++i)
        if (keep[i]) for (int k = 0; k < upr; ++k) want.push_back(in[i * upr + k]);
    gridDim.x = (unsigned)blocks;
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
        for (threadIdx.x = 0; threadIdx.x < 256; ++threadIdx.x)
            compact_rows<U>(in.data(), got.data(), n * upr, (unsigned)upr, keep.data(), offs.data());
    if ((int64_t)want.size() != kept * upr || memcmp(want.data(), got.data(), want.size() * sizeof(U)) != 0) {
        printf("compact mismatch unit=%%d n=%%ld upr=%%d blocks=%%d pattern=%%d\n", (int)sizeof(U), (long)n, upr, blocks, pattern);
        return 1;
    }
    return 0;
}

template <typename T, typename A>
static void naive_nsum(const std::vector<T>& in, std::vector<A>& out, int64_t rows, int64_t cols, int R, bool vn,
                       bool inc, bool torus) {
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            A acc = 0;
            for (int di = -R; di <= R; ++di)
                for (int dj = -R; dj <= R; ++dj) {
                    if (vn && abs(di) + abs(dj) > R) continue;
                    if (di == 0 && dj == 0 && !inc) continue;
                    int64_t r = i + di, c = j + dj;
                    if (torus) { r = (r %% rows + rows) %% rows; c = (c %% cols + cols) %% cols; }
                    else if (r < 0 || r >= rows || c < 0 || c >= cols) continue;
                    acc += (A)in[r * cols + c];
                }
            out[i * cols + j] = acc;
        }
}

template <typename T, typename A, int R, bool VN, int V>
static int check_nsum(int64_t rows, int64_t cols, bool inc, bool torus) {
    if (torus && (rows < 2 * R + 1 || cols < 2 * R + 1)) return 0;
    if (cols %% V) return 0;
    std::vector<T> in(rows * cols);
    for (auto& x : in) x = (T)(rand() %% 251);
    std::vector<A> want(rows * cols), got(rows * cols, (A)-7);
    naive_nsum<T, A>(in, want, rows, cols, R, VN, inc, torus);
    std::vector<T> top(R * cols), bot(R * cols);
    for (int k = 0; k < R; ++k)
        for (int64_t c = 0; c < cols; ++c) {
            top[k * cols + c] = torus ? in[(rows - R + k) * cols + c] : (T)0;
            bot[k * cols + c] = torus ? in[k * cols + c] : (T)0;
        }
    constexpr int SEG = 16;
    gridDim.x = (unsigned)((cols / V + 127) / 128);
    gridDim.y = (unsigned)((rows + SEG - 1) / SEG);
    blockDim.x = 128;
    for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)
        for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
            for (threadIdx.x = 0; threadIdx.x < 128; ++threadIdx.x)
                k_nsum_walk<T, A, R, VN, V, SEG, 1>(in.data(), torus ? top.data() : nullptr, torus ? bot.data() : nullptr,
                                                    got.data(), rows, cols, inc, torus);
    for (int64_t k = 0; k < rows * cols; ++k)
        if (got[k] != want[k]) {
            printf("nsum mismatch R=%%d VN=%%d V=%%d rows=%%ld cols=%%ld inc=%%d torus=%%d at %%ld: %%g vs %%g\n", R, (int)VN, V,
                   (long)rows, (long)cols, (int)inc, (int)torus, (long)k, (double)got[k], (double)want[k]);
            return 1;
        }
    return 0;
}

static int check_diffuse(int64_t rows, int64_t cols, bool torus) {
    if (torus && (rows < 3 || cols < 3)) return 0;
    std::vector<double> in(rows * cols), got(rows * cols, -1.0);
    for (auto& x : in) x = (double)(rand() %% 1000) / 8.0;
    const double rate = 0.25;
    std::vector<double> top(cols), bot(cols);
    for (int64_t c = 0; c < cols; ++c) { top[c] = in[(rows - 1) * cols + c]; bot[c] = in[c]; }
    gridDim.x = (unsigned)((cols + 127) / 128);
    gridDim.y = (unsigned)((rows + 15) / 16);
    for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)
        for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
            for (threadIdx.x = 0; threadIdx.x < 128; ++threadIdx.x)
                k_diffuse_walk<double, 16>(in.data(), torus ? top.data() : nullptr, torus ? bot.data() : nullptr, got.data(),
                                           rows, cols, rate, torus);
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            double acc = 0.0;
            int k = 0;
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    if (!di && !dj) continue;
                    int64_t r = i + di, c = j + dj;
                    if (torus) { r = (r + rows) %% rows; c = (c + cols) %% cols; }
                    else if (r < 0 || r >= rows || c < 0 || c >= cols) continue;
                    acc += in[r * cols + c];
                    ++k;
                }
            const double want = in[i * cols + j] * (1.0 - rate * k / 8.0) + rate / 8.0 * acc;
            if (fabs(want - got[i * cols + j]) > 1e-12) {
                printf("diffuse mismatch rows=%%ld cols=%%ld torus=%%d at (%%ld,%%ld): %%g vs %%g\n", (long)rows, (long)cols,
                       (int)torus, (long)i, (long)j, got[i * cols + j], want);
                return 1;
            }
        }
    return 0;
}

static int check_knn(int n, double sx, double sy, bool torus, int k, double edge, int nq) {
    Geometry g;
    g.x0 = 0; g.y0 = 0; g.sx = sx; g.sy = sy; g.torus = torus;
    g.ncx = std::max(1, (int)floor(sx / (edge * (1.0 + 1e-6))));
    g.ncy = std::max(1, (int)floor(sy / (edge * (1.0 + 1e-6))));
    std::vector<float2> pos(n);
    for (int i = 0; i < n; ++i) {
        // half of the agents clustered, some on a lattice (exact ties)
        if (i %% 3 == 0) pos[i] = float2{(float)(rand() %% 8) * (float)(sx / 8), (float)(rand() %% 8) * (float)(sy / 8)};
        else if (i %% 3 == 1) pos[i] = float2{(float)(0.6 * sx + 0.05 * sx * rand() / RAND_MAX), (float)(0.6 * sy + 0.05 * sy * rand() / RAND_MAX)};
        else pos[i] = float2{(float)(sx * (rand() / (RAND_MAX + 1.0))), (float)(sy * (rand() / (RAND_MAX + 1.0)))};
    }
    const int ncells = g.ncx * g.ncy;
    std::vector<uint32_t> cell(n), start(ncells + 1, 0), order(n);
    for (int i = 0; i < n; ++i) { cell[i] = cell_of(g, pos[i].x, pos[i].y); ++start[cell[i] + 1]; }
    for (int c = 0; c < ncells; ++c) start[c + 1] += start[c];
    std::vector<uint32_t> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) order[fill[cell[i]]++] = i;
    std::vector<float4> packed(n);
    for (int s = 0; s < n; ++s) packed[s] = make_float4(pos[order[s]].x, pos[order[s]].y, 0.f, 0.f);
    std::vector<float2> queries(nq);
    std::vector<int32_t> excl(nq);
    for (int q = 0; q < nq; ++q) {
        if (q %% 2 == 0 && n > 0) { excl[q] = rand() %% n; queries[q] = pos[excl[q]]; }
        else { excl[q] = -1; queries[q] = float2{(float)(sx * (rand() / (RAND_MAX + 1.0))), (float)(sy * (rand() / (RAND_MAX + 1.0)))}; }
    }
    std::vector<int32_t> idx((size_t)nq * k, -9);
    std::vector<double> dist((size_t)nq * k, -9.0);
    gridDim.x = (unsigned)((nq + 127) / 128);
    blockDim.x = 128;
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
        for (threadIdx.x = 0; threadIdx.x < 128; ++threadIdx.x)
            knn_query(packed.data(), order.data(), start.data(), start.data() + 1, g, queries.data(), excl.data(), nq, k,
                      idx.data(), dist.data());
    for (int q = 0; q < nq; ++q) {
        std::vector<std::pair<double, int>> all;
        for (int i = 0; i < n; ++i)
            if (i != excl[q]) all.push_back({ref_distance2(g, queries[q].x, queries[q].y, pos[i].x, pos[i].y), i});
        std::sort(all.begin(), all.end());
        for (int w = 0; w < k; ++w) {
            const int wi = w < (int)all.size() ? all[w].second : -1;
            const double wd = w < (int)all.size() ? sqrt(all[w].first) : INFINITY;
            if (idx[(size_t)q * k + w] != wi || dist[(size_t)q * k + w] != wd) {
                printf("knn mismatch n=%%d size=(%%g,%%g) torus=%%d k=%%d edge=%%g q=%%d w=%%d: (%%d,%%g) vs (%%d,%%g)\n", n, sx, sy,
                       (int)torus, k, edge, q, w, idx[(size_t)q * k + w], dist[(size_t)q * k + w], wi, wd);
                return 1;
            }
        }
    }
    return 0;
}

template <typename T, bool DIFFUSE>
static int check_box3(int64_t rows, int64_t cols, bool inc, bool torus) {
    constexpr int V = 16 / sizeof(T);
    if (cols %% V) return 0;
    if (torus && (rows < 3 || cols < 3)) return 0;
    std::vector<T> in(rows * cols);
    for (auto& x : in) x = (T)(rand() %% 4096) / (T)64;
    std::vector<T> top(cols), bot(cols);
    for (int64_t c = 0; c < cols; ++c) { top[c] = in[(rows - 1) * cols + c]; bot[c] = in[c]; }
    std::vector<double> gotd(rows * cols, -1.0);
    std::vector<T> gott(rows * cols, (T)-1);
    const double rate = 0.25;
    gridDim.x = (unsigned)((cols / V + 127) / 128);
    gridDim.y = (unsigned)((rows + 15) / 16);
    for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)
        for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
            for (threadIdx.x = 0; threadIdx.x < 128; ++threadIdx.x)
                k_box3<T, DIFFUSE, 16>(in.data(), torus ? top.data() : nullptr, torus ? bot.data() : nullptr,
                                       DIFFUSE ? (void*)gott.data() : (void*)gotd.data(), rows, cols, inc, torus, rate);
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) {
            double acc = 0.0;
            int k = 0;
            for (int di = -1; di <= 1; ++di)
                for (int dj = -1; dj <= 1; ++dj) {
                    if (!di && !dj) continue;
                    int64_t r = i + di, c = j + dj;
                    if (torus) { r = (r + rows) %% rows; c = (c + cols) %% cols; }
                    else if (r < 0 || r >= rows || c < 0 || c >= cols) continue;
                    acc += (double)in[r * cols + c];
                    ++k;
                }
            const double self = (double)in[i * cols + j];
            bool ok;
            if (DIFFUSE) ok = gott[i * cols + j] == (T)fma(rate / 8.0, acc, self * (1.0 - rate * k / 8.0));
            else ok = gotd[i * cols + j] == (inc ? acc + self : acc);
            if (!ok) {
                printf("box3 mismatch T=%%d diffuse=%%d rows=%%ld cols=%%ld inc=%%d torus=%%d at (%%ld,%%ld)\n", (int)sizeof(T), (int)DIFFUSE,
                       (long)rows, (long)cols, (int)inc, (int)torus, (long)i, (long)j);
                return 1;
            }
        }
    return 0;
}

static int check_nd(int ndim, const int64_t* dims, int radius, bool vn, bool inc, bool torus) {
    NdShape sh;
    sh.ndim = ndim;
    int64_t n = 1;
    for (int d = ndim - 1; d >= 0; --d) { sh.dim[d] = dims[d]; sh.stride[d] = n; n *= dims[d]; if (torus && dims[d] < 2 * radius + 1) return 0; }
    for (int d = ndim; d < kMaxDims; ++d) { sh.dim[d] = 1; sh.stride[d] = 0; }
    std::vector<int32_t> in(n), got(n, -1);
    for (auto& x : in) x = rand() %% 100;
    gridDim.x = (unsigned)((n + 255) / 256);
    blockDim.x = 256;
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
        for (threadIdx.x = 0; threadIdx.x < 256; ++threadIdx.x)
            k_neighborhood_sum_nd<int32_t, int32_t>(in.data(), got.data(), n, sh, radius, vn, inc, torus);
    for (int64_t idx = 0; idx < n; ++idx) {
        int64_t c[4] = {0, 0, 0, 0}, rem = idx;
        for (int d = 0; d < ndim; ++d) { c[d] = rem / sh.stride[d]; rem %%= sh.stride[d]; }
        int want = 0;
        int o[4];
        for (o[0] = -radius; o[0] <= radius; ++o[0])
        for (o[1] = (ndim > 1 ? -radius : 0); o[1] <= (ndim > 1 ? radius : 0); ++o[1])
        for (o[2] = (ndim > 2 ? -radius : 0); o[2] <= (ndim > 2 ? radius : 0); ++o[2])
        for (o[3] = (ndim > 3 ? -radius : 0); o[3] <= (ndim > 3 ? radius : 0); ++o[3]) {
            int man = abs(o[0]) + abs(o[1]) + abs(o[2]) + abs(o[3]);
            if ((vn && man > radius) || (man == 0 && !inc)) continue;
            int64_t at = 0; bool ok = true;
            for (int d = 0; d < ndim; ++d) {
                int64_t x = c[d] + o[d];
                if (torus) x = (x + 8 * dims[d]) %% dims[d]; else if (x < 0 || x >= dims[d]) ok = false;
                at += x * sh.stride[d];
            }
            if (ok) want += in[at];
        }
        if (want != got[idx]) { printf("nd mismatch ndim=%%d r=%%d vn=%%d inc=%%d torus=%%d at %%ld\n", ndim, radius, (int)vn, (int)inc, (int)torus, (long)idx); return 1; }
    }
    return 0;
}

// point_distances_nd of csrc/continuous.cu against the golden vectors of the live ContinuousSpace (1, 3 and 4
// dimensions), handed over as a flat binary file by main() below
static int check_point_distances_nd(const char* path) {
    FILE* fp = fopen(path, "rb");
    if (!fp) { printf("cannot open %%s\n", path); return 1; }
    int bad = 0, ncases = 0;
    int64_t hdr[4];
    while (fread(hdr, sizeof(int64_t), 4, fp) == 4) {
        const int ndim = (int)hdr[0], torus = (int)hdr[1];
        const int64_t n = hdr[2], nq = hdr[3];
        std::vector<double> size(ndim), points(nq * ndim), dist(nq * n), delta(nq * n * ndim);
        std::vector<float> pos(n * ndim);
        if (fread(size.data(), 8, ndim, fp) != (size_t)ndim || fread(pos.data(), 4, n * ndim, fp) != (size_t)(n * ndim) ||
            fread(points.data(), 8, nq * ndim, fp) != (size_t)(nq * ndim) || fread(dist.data(), 8, nq * n, fp) != (size_t)(nq * n) ||
            fread(delta.data(), 8, nq * n * ndim, fp) != (size_t)(nq * n * ndim)) { printf("short read\n"); return 1; }
        for (int64_t q = 0; q < nq; ++q) {
            NdQuery g;
            for (int d = 0; d < kMaxSpaceDims; ++d) { g.p[d] = d < ndim ? points[q * ndim + d] : 0.0; g.size[d] = d < ndim ? size[d] : 1.0; }
            g.ndim = ndim; g.torus = torus;
            std::vector<double> gd(n, -1.0), gv(n * ndim, -1.0);
            blockDim.x = 256;
            gridDim.x = (unsigned)((n + 255) / 256);
            for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
                for (threadIdx.x = 0; threadIdx.x < 256; ++threadIdx.x)
                    point_distances_nd(pos.data(), nullptr, n, g, gd.data(), gv.data());
            if (memcmp(gd.data(), &dist[q * n], 8 * n) != 0 || memcmp(gv.data(), &delta[q * n * ndim], 8 * n * ndim) != 0) {
                printf("point_distances_nd mismatch ndim=%%d torus=%%d query %%ld\n", ndim, torus, (long)q);
                ++bad;
            }
        }
        ++ncases;
    }
    fclose(fp);
    if (ncases < 7) { printf("only %%d n-d cases read\n", ncases); return 1; }
    return bad;
}

int main(int argc, char** argv) {
    if (argc > 1 && check_point_distances_nd(argv[1])) { printf("FAILED: n-d point distances\n"); return 1; }

    int bad = argc > 1 ? 0 : 1;  // the golden file is part of the check
    const int64_t shapes[][2] = {{1, 1}, {1, 40}, {40, 1}, {3, 3}, {1, 4}, {3, 4}, {5, 8}, {16, 128}, {17, 132}, {35, 260}, {5, 516}, {33, 1024}, {50, 2048}};
    for (auto& s : shapes)
        for (int inc = 0; inc < 2; ++inc)
            for (int torus = 0; torus < 2; ++torus) {
                bad += check_nsum<uint8_t, int32_t, 1, false, 4>(s[0], s[1], inc, torus);
                bad += check_nsum<uint8_t, int32_t, 1, true, 4>(s[0], s[1], inc, torus);
                bad += check_nsum<uint8_t, int32_t, 2, false, 4>(s[0], s[1], inc, torus);
                bad += check_nsum<uint8_t, int32_t, 2, true, 4>(s[0], s[1], inc, torus);
                bad += check_nsum<uint8_t, int32_t, 1, false, 1>(s[0], s[1], inc, torus);
                bad += check_nsum<uint8_t, int32_t, 2, true, 1>(s[0], s[1], inc, torus);
                bad += check_nsum<float, double, 1, false, 1>(s[0], s[1], inc, torus);
                bad += check_nsum<float, double, 2, false, 1>(s[0], s[1], inc, torus);
                bad += check_nsum<double, double, 1, true, 1>(s[0], s[1], inc, torus);
                bad += check_nsum<int32_t, int32_t, 2, true, 1>(s[0], s[1], inc, torus);
                if (!inc) bad += check_diffuse(s[0], s[1], torus);
                bad += check_box3<float, false>(s[0], s[1], inc, torus);
                if (!inc) bad += check_box3<float, true>(s[0], s[1], inc, torus);
                if (!inc) bad += check_box3<double, true>(s[0], s[1], inc, torus);
            }
    {
        const int64_t d1[] = {9}, d3[] = {4, 5, 6}, d3b[] = {7, 3, 5}, d4[] = {3, 4, 3, 5};
        for (int r = 1; r <= 2; ++r)
            for (int m = 0; m < 8; ++m) {
                bad += check_nd(1, d1, r, m & 1, m & 2, m & 4);
                bad += check_nd(3, d3, r, m & 1, m & 2, m & 4);
                bad += check_nd(3, d3b, r, m & 1, m & 2, m & 4);
                bad += check_nd(4, d4, r, m & 1, m & 2, m & 4);
            }
    }
    for (int torus = 0; torus < 2; ++torus) {
        bad += check_knn(2000, 300.0, 200.0, torus, 1, 5.0, 200);
        bad += check_knn(2000, 300.0, 200.0, torus, 4, 5.0, 200);
        bad += check_knn(2000, 300.0, 200.0, torus, 16, 2.0, 200);
        bad += check_knn(500, 40.0, 2.0, torus, 7, 0.5, 100);
        bad += check_knn(500, 40.0, 2.0, torus, 7, 30.0, 100);
        bad += check_knn(3, 10.0, 10.0, torus, 5, 1.0, 20);
        bad += check_knn(64, 16.0, 16.0, torus, 9, 4.0, 64);
        bad += check_knn(64, 16.0, 16.0, torus, 64, 7.9, 64);
        bad += check_knn(0, 16.0, 16.0, torus, 2, 4.0, 8);
    }
    for (int pattern = 0; pattern < 4; ++pattern)
        for (int64_t n : {0, 1, 2, 255, 256, 257, 5000}) {
            bad += check_compact<uint8_t>(n, 1, 3, pattern);
            bad += check_compact<uint8_t>(n, 3, 64, pattern);
            bad += check_compact<uint16_t>(n, 1, 1, pattern);
            bad += check_compact<uint32_t>(n, 3, 7, pattern);
            bad += check_compact<uint2>(n, 1, 2, pattern);
            bad += check_compact<uint4>(n, 2, 5, pattern);
        }
    printf(bad ? "FAILED: %%d case(s)\n" : "all cases agree\n", bad);
    return bad != 0;
}
'''


def _dump_nd_golden(path: str) -> None:
    """tests/golden/continuous_queries_nd.npz (live ContinuousSpace) as a flat binary file for the C++ harness."""
    import numpy as np

    g = np.load(os.path.join(ROOT, "tests", "golden", "continuous_queries_nd.npz"))
    with open(path, "wb") as f:
        for c, (n, torus, _radius, _k) in enumerate(g["cases"]):
            pos, points, dims = g[f"pos_{c}"], g[f"points_{c}"], g[f"dims_{c}"]
            np.array([pos.shape[1], int(torus), pos.shape[0], points.shape[0]], dtype=np.int64).tofile(f)
            (dims[:, 1] - dims[:, 0]).astype(np.float64).tofile(f)
            pos.astype(np.float32).tofile(f)
            points.astype(np.float64).tofile(f)
            g[f"dist_all_{c}"].astype(np.float64).tofile(f)
            g[f"delta_all_{c}"].astype(np.float64).tofile(f)


def main() -> int:
    layers = open(os.path.join(CSRC, "layers.cu")).read()
    cont = open(os.path.join(CSRC, "continuous.cu")).read()
    layers_part = between(layers, "template <typename T, int V> struct WalkVec;", "template <typename T, typename A, int V, int MINB1, int MINB2>")
    cont_part = (between(cont, "struct Geometry {", "// ---- cell list:")
                 + between(cont, "__device__ __forceinline__ double ref_distance2", "// one component of calculate_difference_vector")
                 + between(cont, "// one component of calculate_difference_vector", "// visit the distinct cells of the 3x3 block")
                 + between(cont, "__global__ void __launch_bounds__(128)\nknn_query", "// ---- calculate_distances / calculate_difference_vector of ONE point")
                 + between(cont, "constexpr int kMaxSpaceDims = 8;", "struct CellListWs {"))
    layers_part += between(layers, "constexpr int kMaxDims = 4;", "#define MB_DISPATCH_DTYPE")
    agents = open(os.path.join(CSRC, "agents.cu")).read()
    agents_part = between(agents, "__global__ void flags_to_words", "template <typename U>\nint launch_rows")
    src = HARNESS % {"layers": layers_part, "continuous": cont_part, "agents": agents_part}
    with tempfile.TemporaryDirectory() as tmp:
        cpp = os.path.join(tmp, "emu.cpp")
        exe = os.path.join(tmp, "emu")
        with open(cpp, "w") as f:
            f.write(src)
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I", os.path.join(ROOT, "tools", "emu"), cpp, "-o", exe],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            return 2
        nd_bin = os.path.join(tmp, "nd_cases.bin")
        _dump_nd_golden(nd_bin)
        r = subprocess.run([exe, nd_bin], capture_output=True, text=True)
        sys.stdout.write(r.stdout[-4000:])
        return r.returncode


if __name__ == "__main__":
    sys.exit(main())
