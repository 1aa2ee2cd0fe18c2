// tests/cpp/cuda_emul.h -- run __global__/__device__ code of mvs-texturing_b200/csrc serially on the host.
//
// Test infrastructure for the container without a GPU: the kernel SOURCE is compiled unchanged by g++
// (-ffp-contract=off mirrors nvcc -fmad=false; explicit __fmaf_rn stays an FMA) and every (block, thread)
// index is executed in order.  Good for checking index arithmetic, tree construction and traversal
// logic; it says nothing about races, memory ordering or performance.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __launch_bounds__(...)

struct EmulDim { unsigned x, y, z; };
static EmulDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

template <typename F>
static void emul_launch(unsigned grid, unsigned block, F body)
{
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; body(); }
}

static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline void __threadfence() {}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
#define __shared__ static
static inline void __syncthreads() {}   // kernels that need a block barrier must not be launched serially

// Warp shuffles are NOT emulated: the stub returns the caller's own value, so a shuffle reduction
// degenerates to "lane 0 only" -- a harness has to redo such reductions itself.
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }

// __ballot_sync via two passes over the warp (valid when the code before the ballot has no side effects
// that the second pass would see): pass 1 records every lane's predicate, pass 2 returns the ballot.
static bool g_emul_ballot_record = false;
static unsigned g_emul_ballot_acc = 0, g_emul_ballot_value = 0;
static inline unsigned __ballot_sync(unsigned, bool p)
{
    if (g_emul_ballot_record) { if (p) g_emul_ballot_acc |= 1u << (threadIdx.x & 31u); return 0u; }
    return g_emul_ballot_value;
}
template <typename F>
static void emul_launch_ballot(unsigned grid, unsigned block, F body)
{
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned w = 0; w < block; w += 32) {
            blockIdx.x = b;
            g_emul_ballot_record = true; g_emul_ballot_acc = 0;
            for (unsigned t = w; t < w + 32 && t < block; ++t) { threadIdx.x = t; body(); }
            g_emul_ballot_record = false; g_emul_ballot_value = g_emul_ballot_acc;
            for (unsigned t = w; t < w + 32 && t < block; ++t) { threadIdx.x = t; body(); }
        }
}
