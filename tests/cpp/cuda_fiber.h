// tests/cpp/cuda_fiber.h -- cooperative host emulation of CUDA kernels: one fiber (ucontext) per CUDA thread.
//
// Unlike cuda_emul.h (threads run to completion one after the other) this runs kernels that synchronise:
// __syncthreads, __syncwarp, warp shuffles / ballots, cooperative-groups grid.sync() and dataflow spin loops
// (they must poll through __nanosleep, which yields).  Scheduling is deterministic round robin, so this
// checks index arithmetic and protocol LOGIC (does every wait get released, are the results right); it says
// nothing about data races, memory ordering or performance.  Static __shared__ arrays become function statics,
// i.e. ONE instance: kernels that use shared memory must be launched with grid = 1 here.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __launch_bounds__(...)
#define __shared__ static

struct EmulDim { unsigned x, y, z; };
static EmulDim threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

namespace emul {

struct Barrier { unsigned expected = 0, arrived = 0, gen = 0; };
struct Warp { Barrier bar; unsigned long long slot[32]; };
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    unsigned tx = 0, bx = 0, rank = 0;
    bool done = false;
    Warp *warp = nullptr;
    Barrier *block = nullptr;
    Barrier *grid = nullptr;      // grid.sync() scope: one per emulated device
    std::vector<char> *shared = nullptr;   // per-block scratch for emul_block_shared()
};

static ucontext_t g_sched;
static Fiber *g_cur = nullptr;
static std::function<void()> g_body;
static Barrier g_grid;
static unsigned long long g_switches = 0, g_switch_limit = 0;
// 0 = round robin in thread order; otherwise every scheduler pass visits the live fibers in a fresh pseudo-random order
// (hardware gives no ordering guarantee either: protocols must not depend on who runs first)
static unsigned long long g_schedule_seed = 0;
static inline void set_schedule(unsigned long long seed) { g_schedule_seed = seed; }
static inline unsigned long long next_rand() { g_schedule_seed = g_schedule_seed * 6364136223846793005ull + 1442695040888963407ull; return g_schedule_seed >> 33; }

static inline void yield() { Fiber *f = g_cur; swapcontext(&f->ctx, &g_sched); }

static inline void wait(Barrier &b)
{
    const unsigned g = b.gen;
    if (++b.arrived >= b.expected) { b.arrived = 0; ++b.gen; return; }
    while (b.gen == g) yield();
}
// a thread that returns no longer takes part in barriers (sm_70+: exited threads count as arrived)
static inline void leave(Barrier &b)
{
    if (b.expected) --b.expected;
    if (b.expected && b.arrived >= b.expected) { b.arrived = 0; ++b.gen; }
}

static void entry()
{
    g_body();
    Fiber *f = g_cur;
    f->done = true;
    leave(f->warp->bar); leave(*f->block); leave(*f->grid);
    swapcontext(&f->ctx, &g_sched);
}

// returns false if the launch did not finish within `max_switches` context switches (a hang in the protocol)
template <typename F>
static bool launch(unsigned grid, unsigned block, F body, unsigned long long max_switches = 2000000000ull, size_t stack_bytes = 128 * 1024)
{
    gridDim.x = grid; blockDim.x = block;
    const unsigned wpb = (block + 31) / 32;
    std::vector<Fiber> fibers((size_t)grid * block);
    std::vector<Warp> warps((size_t)grid * wpb);
    std::vector<Barrier> blocks(grid);
    g_grid = Barrier(); g_grid.expected = grid * block;
    g_body = body;
    for (unsigned b = 0; b < grid; ++b) {
        blocks[b].expected = block;
        for (unsigned t = 0; t < block; ++t) {
            Fiber &f = fibers[(size_t)b * block + t];
            f.tx = t; f.bx = b; f.warp = &warps[(size_t)b * wpb + t / 32]; f.block = &blocks[b]; f.grid = &g_grid;
            f.warp->bar.expected++;
            f.stack.resize(stack_bytes);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, (void (*)())entry, 0);
        }
    }
    size_t remaining = fibers.size();
    g_switches = 0;
    std::vector<unsigned> order_(fibers.size());
    for (size_t i = 0; i < order_.size(); ++i) order_[i] = (unsigned)i;
    while (remaining) {
        if (g_schedule_seed) for (size_t i = order_.size(); i > 1; --i) std::swap(order_[i - 1], order_[next_rand() % i]);
        for (size_t oi = 0; oi < fibers.size(); ++oi) {
            Fiber &f = fibers[order_[oi]];
            if (f.done) continue;
            g_cur = &f; threadIdx.x = f.tx; blockIdx.x = f.bx;
            swapcontext(&g_sched, &f.ctx);
            if (f.done) --remaining;
            if (++g_switches > max_switches) { g_cur = nullptr; return false; }
        }
    }
    g_cur = nullptr;
    return true;
}

// Several emulated DEVICES at once (one process per GPU in the real run): `ranks` grids of `grid` x `block` threads, each
// with its own grid.sync() scope, all scheduled round robin, so that kernels which spin on flags in "peer memory" (plain
// host pointers here) make progress.  body(rank) runs as the kernel of that rank; blockIdx / gridDim are per rank.
template <typename F>
static bool launch_ranks(unsigned ranks, unsigned grid, unsigned block, F body, unsigned long long max_switches = 4000000000ull,
                         size_t stack_bytes = 128 * 1024)
{
    gridDim.x = grid; blockDim.x = block;
    const unsigned wpb = (block + 31) / 32;
    std::vector<Fiber> fibers((size_t)ranks * grid * block);
    std::vector<Warp> warps((size_t)ranks * grid * wpb);
    std::vector<Barrier> blocks((size_t)ranks * grid), grids(ranks);
    std::vector<std::vector<char> > shared((size_t)ranks * grid);
    g_body = [&] { body(g_cur->rank); };
    for (unsigned r = 0; r < ranks; ++r) {
        grids[r].expected = grid * block;
        for (unsigned b = 0; b < grid; ++b) {
            const size_t gb = (size_t)r * grid + b;
            blocks[gb].expected = block;
            for (unsigned t = 0; t < block; ++t) {
                Fiber &f = fibers[gb * block + t];
                f.tx = t; f.bx = b; f.rank = r; f.warp = &warps[gb * wpb + t / 32]; f.block = &blocks[gb]; f.grid = &grids[r];
                f.shared = &shared[gb];
                f.warp->bar.expected++;
                f.stack.resize(stack_bytes);
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = &g_sched;
                makecontext(&f.ctx, (void (*)())entry, 0);
            }
        }
    }
    size_t remaining = fibers.size();
    g_switches = 0;
    std::vector<unsigned> order_(fibers.size());
    for (size_t i = 0; i < order_.size(); ++i) order_[i] = (unsigned)i;
    while (remaining) {
        if (g_schedule_seed) for (size_t i = order_.size(); i > 1; --i) std::swap(order_[i - 1], order_[next_rand() % i]);
        for (size_t oi = 0; oi < fibers.size(); ++oi) {
            Fiber &f = fibers[order_[oi]];
            if (f.done) continue;
            g_cur = &f; threadIdx.x = f.tx; blockIdx.x = f.bx;
            swapcontext(&g_sched, &f.ctx);
            if (f.done) --remaining;
            if (++g_switches > max_switches) { g_cur = nullptr; return false; }
        }
    }
    g_cur = nullptr;
    return true;
}

// per-block scratch standing in for a static __shared__ array when several blocks are alive at once (launch_ranks):
// the harness replaces `__shared__ T name[N];` by `T *name = (T *)emul_block_shared(sizeof(T) * N);` in the kernel text
static inline void *block_shared(size_t bytes)
{
    std::vector<char> &s = *g_cur->shared;
    if (s.size() < bytes) s.resize(bytes);
    return s.data();
}

// kernels WITHOUT any synchronisation: every (block, thread) index runs to completion in order, no fibers (fast).
// A kernel that does reach a barrier / shuffle here dereferences g_cur == nullptr and crashes -- by design.
template <typename F>
static void launch_serial(unsigned grid, unsigned block, F body)
{
    gridDim.x = grid; blockDim.x = block;
    g_cur = nullptr;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) { blockIdx.x = b; threadIdx.x = t; body(); }
}

template <typename T> static inline unsigned long long to_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace emul

static inline void __syncthreads() { emul::wait(*emul::g_cur->block); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::wait(emul::g_cur->warp->bar); }
static inline void __nanosleep(unsigned) { emul::yield(); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
    emul::Warp *w = emul::g_cur->warp;
    const unsigned lane = threadIdx.x & 31u;
    w->slot[lane] = emul::to_bits(v);
    emul::wait(w->bar);
    const T r = emul::from_bits<T>(w->slot[(lane ^ (unsigned)lane_mask) & 31u]);
    emul::wait(w->bar);
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
    emul::Warp *w = emul::g_cur->warp;
    w->slot[threadIdx.x & 31u] = emul::to_bits(v);
    emul::wait(w->bar);
    const T r = emul::from_bits<T>(w->slot[(unsigned)src & 31u]);
    emul::wait(w->bar);
    return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta)
{
    emul::Warp *w = emul::g_cur->warp;
    const unsigned lane = threadIdx.x & 31u;
    w->slot[lane] = emul::to_bits(v);
    emul::wait(w->bar);
    const T r = lane + delta < 32u ? emul::from_bits<T>(w->slot[lane + delta]) : v;
    emul::wait(w->bar);
    return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta)
{
    emul::Warp *w = emul::g_cur->warp;
    const unsigned lane = threadIdx.x & 31u;
    w->slot[lane] = emul::to_bits(v);
    emul::wait(w->bar);
    const T r = lane >= delta ? emul::from_bits<T>(w->slot[lane - delta]) : v;
    emul::wait(w->bar);
    return r;
}
// redux.sync: every lane of the warp arrives (each with the mask of its own group); minimum over the lanes named in the mask
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
    emul::Warp *w = emul::g_cur->warp;
    w->slot[threadIdx.x & 31u] = v;
    emul::wait(w->bar);
    unsigned r = 0xFFFFFFFFu;
    for (unsigned l = 0; l < 32; ++l) if ((mask >> l) & 1u) r = std::min(r, (unsigned)w->slot[l]);
    emul::wait(w->bar);
    return r;
}
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    emul::Warp *w = emul::g_cur->warp;
    w->slot[threadIdx.x & 31u] = v;
    emul::wait(w->bar);
    unsigned r = 0u;
    for (unsigned l = 0; l < 32; ++l) if ((mask >> l) & 1u) r = std::max(r, (unsigned)w->slot[l]);
    emul::wait(w->bar);
    return r;
}
static inline unsigned __ballot_sync(unsigned, bool p)
{
    emul::Warp *w = emul::g_cur->warp;
    w->slot[threadIdx.x & 31u] = p ? 1ull : 0ull;
    emul::wait(w->bar);
    unsigned r = 0;
    const unsigned base = threadIdx.x & ~31u;
    for (unsigned l = 0; l < 32 && base + l < blockDim.x; ++l) if (w->slot[l]) r |= 1u << l;
    emul::wait(w->bar);
    return r;
}

static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
static inline uint2 __ldcg(const uint2 *p) { return *p; }
static inline uint4 __ldcg(const uint4 *p) { return *p; }
static inline float4 __ldcg(const float4 *p) { return *p; }
template <typename T> static inline void __stcg(T *p, T v) { *(volatile T *)p = v; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
