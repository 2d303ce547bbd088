// Host shims that let g++ compile the BODY of selected device functions / kernels of mesa_b200/csrc
// (extracted textually by tools/emulate_kernels.py) so that their index arithmetic can be exercised on a
// machine without a GPU.  Threads of a launch run one after the other -- only kernels whose threads do not
// communicate (no shuffles, no shared memory, no atomics ordering) can be emulated this way.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct dim3 { unsigned x = 1, y = 1, z = 1; };
static dim3 blockIdx, threadIdx, blockDim, gridDim;

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
