// cuda_runtime.h (stand-in) -- TEST INFRASTRUCTURE: runs the repository's CUDA kernels on the CPU, thread for thread.
//
// g++ compiles the kernel headers against this file instead of the CUDA toolkit's.  A launch runs the blocks of the
// grid one after the other; the threads of a block are real OS threads, so everything the kernels do between
// barriers is genuinely concurrent:
//   __syncthreads()              a barrier of the block's threads
//   __shfl_*_sync / __ballot_sync / __any_sync (full mask)   a warp-wide exchange through a per-warp slot array and a
//                                warp barrier (all 32 lanes must arrive: a divergent call deadlocks here as it would
//                                hang or corrupt on the device)
//   atomicCAS / atomicAdd / atomicOr / atomicExch            GCC __atomic builtins (sequentially consistent)
//   __shared__                   `static`: one block runs at a time; dynamic shared memory is one global buffer
// Not modelled: the memory model beyond sequential consistency, warp-synchronous execution outside the *_sync calls,
// bank conflicts, occupancy.  What it does check: indexing, barrier placement, the arithmetic -- and, built with
// -fsanitize=thread or address, data races between barriers and out-of-bounds accesses in the kernels' own source.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

#define __CUSIM__ 1
#define __host__
#define __device__
#define __global__ static
#define __forceinline__ inline
#define __shared__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
typedef void* cudaStream_t;

namespace cusim {
struct Block {
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<unsigned long long> slots;      // 32 per warp
    unsigned n_threads = 0;
};
inline Block*& cur() { static Block* b = nullptr; return b; }
alignas(128) inline unsigned char dyn_smem[232448];
}  // namespace cusim

inline thread_local uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

static inline void __syncthreads() { pthread_barrier_wait(&cusim::cur()->bar); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { pthread_barrier_wait(&cusim::cur()->warp_bar[threadIdx.x >> 5]); }

namespace cusim {
// every lane of the warp deposits v, then reads lane `src`'s value
inline unsigned long long exchange(unsigned long long v, unsigned src) {
    Block* b = cur();
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    b->slots[w * 32 + lane] = v;
    pthread_barrier_wait(&b->warp_bar[w]);
    const unsigned long long r = b->slots[w * 32 + (src & 31u)];
    pthread_barrier_wait(&b->warp_bar[w]);      // nobody overwrites a slot before everybody has read
    return r;
}
}  // namespace cusim

template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return (T)cusim::exchange((unsigned long long)v, (unsigned)src); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const unsigned lane = threadIdx.x & 31u;
    const T r = (T)cusim::exchange((unsigned long long)v, lane >= d ? lane - d : lane);
    return lane >= d ? r : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return (T)cusim::exchange((unsigned long long)v, (threadIdx.x & 31u) ^ (unsigned)m); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    cusim::Block* b = cusim::cur();
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    b->slots[w * 32 + lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&b->warp_bar[w]);
    for (unsigned i = 0; i < 32; ++i) r |= (unsigned)b->slots[w * 32 + i] << i;
    pthread_barrier_wait(&b->warp_bar[w]);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
template <class T, class U> static inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicOr(T* p, U v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class U> static inline T atomicMax(T* p, U v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur < (T)v && !__atomic_compare_exchange_n(p, &cur, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
template <class T, class U> static inline T atomicExch(T* p, U v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_SEQ_CST); }

namespace cusim {
// kern<<<grid, block, smem>>>(args...): blocks one after the other, the threads of a block side by side
template <class F> void launch(dim3 grid, unsigned block, F body) {
    gridDim = grid; blockDim = dim3(block);
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            Block b;
            b.n_threads = block;
            pthread_barrier_init(&b.bar, nullptr, block);
            const unsigned n_warps = (block + 31) / 32;
            b.warp_bar.resize(n_warps);
            b.slots.assign((size_t)n_warps * 32, 0ull);
            for (unsigned w = 0; w < n_warps; ++w) pthread_barrier_init(&b.warp_bar[w], nullptr, std::min(32u, block - w * 32));
            cur() = &b;
            std::vector<std::thread> th;
            th.reserve(block);
            for (unsigned t = 0; t < block; ++t)
                th.emplace_back([&, t] { threadIdx = uint3{t, 0, 0}; blockIdx = uint3{bx, by, 0}; body(); });
            for (auto& x : th) x.join();
            pthread_barrier_destroy(&b.bar);
            for (auto& wb : b.warp_bar) pthread_barrier_destroy(&wb);
            cur() = nullptr;
        }
}
}  // namespace cusim
