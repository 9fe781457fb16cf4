// tests/emul/cuda_emul.h — TEST INFRASTRUCTURE ONLY.
//
// A tiny CPU stand-in for the CUDA execution model so that the product's kernel
// sources (plonkish_b200/csrc/*.cuh) can be compiled with g++ and their *logic*
// (index arithmetic, histogram/scan/scatter, segmented reductions, edge cases)
// exercised against the oracle in the CPU test suite.  One OS thread per CUDA
// thread of a block, blocks run one after another, __syncthreads / warp shuffles
// are real barriers.  It never ships and nothing in plonkish_b200/ includes it
// unless PLONKISH_EMUL is defined by the test build.
#pragma once
#ifndef PLONKISH_EMUL
#error "cuda_emul.h is only for the PLONKISH_EMUL test build"
#endif

#include <stdint.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint3e { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

namespace emul {
inline thread_local uint3e t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char *g_dyn_smem = nullptr;
inline std::barrier<> *g_block_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
inline std::vector<uint64_t> g_xchg;  // [warp][32]

inline unsigned linear_tid() {
    return t_threadIdx.x + g_blockDim.x * (t_threadIdx.y + g_blockDim.y * t_threadIdx.z);
}
inline uint64_t warp_exchange(uint64_t v, unsigned src_lane) {
    unsigned tid = linear_tid(), w = tid >> 5, lane = tid & 31;
    g_xchg[w * 32 + lane] = v;
    g_warp_barrier[w]->arrive_and_wait();
    uint64_t r = g_xchg[w * 32 + (src_lane & 31)];
    g_warp_barrier[w]->arrive_and_wait();
    return r;
}

// Whole-struct shuffle: lane reads the blob of lane + delta (own blob if out of range).
inline std::vector<unsigned char> g_blob;  // [warp][32][256]
template <class T>
inline T warp_exchange_blob(const T &v, int delta) {
    static_assert(sizeof(T) <= 256, "blob too large");
    unsigned tid = linear_tid(), w = tid >> 5, lane = tid & 31;
    memcpy(&g_blob[(w * 32 + lane) * 256], &v, sizeof(T));
    g_warp_barrier[w]->arrive_and_wait();
    int src = (int)lane + delta;
    if (src < 0 || src > 31) src = (int)lane;
    T r;
    memcpy(&r, &g_blob[(w * 32 + (unsigned)src) * 256], sizeof(T));
    g_warp_barrier[w]->arrive_and_wait();
    return r;
}

// Runs fn() once per CUDA thread of the launch.
inline void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &fn) {
    unsigned nthreads = block.x * block.y * block.z;
    g_blockDim = block;
    g_gridDim = grid;
    std::vector<unsigned char> smem(dyn_smem + 16);
    g_dyn_smem = smem.data();
    std::barrier<> block_barrier(nthreads);
    g_block_barrier = &block_barrier;
    unsigned nwarps = (nthreads + 31) / 32;
    g_warp_barrier.clear();
    for (unsigned w = 0; w < nwarps; ++w) {
        unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        g_warp_barrier.emplace_back(new std::barrier<>(lanes));
    }
    g_xchg.assign(nwarps * 32, 0);
    g_blob.assign((size_t)nwarps * 32 * 256, 0);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, t] {
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        t_blockIdx.x = bx; t_blockIdx.y = by; t_blockIdx.z = bz;
                        fn();
                        block_barrier.arrive_and_wait();
                    }
        });
    }
    for (auto &th : pool) th.join();
    g_block_barrier = nullptr;
}
}  // namespace emul

#define threadIdx emul::t_threadIdx
#define blockIdx emul::t_blockIdx
#define blockDim emul::g_blockDim
#define gridDim emul::g_gridDim

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

static inline void __syncthreads() { emul::g_block_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_exchange(0, 0); }
static inline unsigned __shfl_sync(unsigned, unsigned v, int src) { return (unsigned)emul::warp_exchange(v, (unsigned)src); }
static inline unsigned __shfl_down_sync(unsigned, unsigned v, unsigned d) {
    unsigned lane = emul::linear_tid() & 31;
    unsigned src = lane + d < 32 ? lane + d : lane;
    return (unsigned)emul::warp_exchange(v, src);
}
static inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned d) {
    unsigned lane = emul::linear_tid() & 31;
    unsigned src = lane >= d ? lane - d : lane;
    return (unsigned)emul::warp_exchange(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned lane = emul::linear_tid() & 31;
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) {
        unsigned bit = (unsigned)emul::warp_exchange(pred ? 1u : 0u, l);
        r |= (bit & 1u) << l;
    }
    (void)lane;
    return r;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned atomicCAS(unsigned *p, unsigned expect, unsigned v) {
    __atomic_compare_exchange_n(p, &expect, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
    return expect;  // the old value, as CUDA returns it
}
static inline unsigned atomicMax(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
