// tests/cpp/emul_include/cooperative_groups.h -- the slice of cooperative groups csrc/*.cu uses, on top of
// cuda_fiber.h.  Threads run one at a time, so coalesced_threads() is the group of the calling thread alone
// (any subset of the converged threads is a legal coalesced group).
#pragma once
#include "cuda_fiber.h"

namespace cooperative_groups {

struct grid_group {
    void sync() const { emul::wait(*emul::g_cur->grid); }
    unsigned long long thread_rank() const { return (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; }
    unsigned long long size() const { return (unsigned long long)gridDim.x * blockDim.x; }
};
static inline grid_group this_grid() { return grid_group(); }

struct thread_block {
    void sync() const { __syncthreads(); }
    unsigned thread_rank() const { return threadIdx.x; }
    unsigned size() const { return blockDim.x; }
};
static inline thread_block this_thread_block() { return thread_block(); }

struct coalesced_group {
    unsigned thread_rank() const { return 0; }
    unsigned size() const { return 1; }
    template <typename T> T shfl(T v, int) const { return v; }
    void sync() const {}
};
static inline coalesced_group coalesced_threads() { return coalesced_group(); }

}  // namespace cooperative_groups
