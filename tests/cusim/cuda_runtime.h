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
// A thread that returns from the kernel stays behind as a ghost that keeps arriving at the block's barriers (exited threads
// count as arrived on the device).  CUSIM_TRACE=1 prints every launch; a watchdog (CUSIM_WATCHDOG_S, default 120 s) aborts
// a launch in which every thread has been waiting that long and says where each one waits (a split warp: some lanes at a
// full-mask vote, the others at __syncthreads -- how k_part_count's bare read of a shared counter was found).
// Not modelled: the memory model beyond sequential consistency, warp-synchronous execution outside the *_sync calls,
// bank conflicts, occupancy.  What it does check: indexing, barrier placement, the arithmetic -- and, built with
// -fsanitize=thread or address, data races between barriers and out-of-bounds accesses in the kernels' own source.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <math.h>

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
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
typedef void* cudaStream_t;
// CUDA's global min / max overloads (mixed signedness and widths resolve the way nvcc resolves them for the calls the
// kernels make: both arguments converted to the wider unsigned type)
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long min(unsigned a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned b) { return a < b ? a : b; }
static inline unsigned long long min(unsigned a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned b) { return a < b ? a : b; }

inline thread_local uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace cusim {
// __syncthreads() is a pthread barrier of ALL the block's threads.  On the device a thread that has returned from the
// kernel counts as arrived at every later barrier; here such a thread stays behind as a ghost that keeps arriving until
// the last thread of the block has returned (that one announces at which barrier generation everybody may leave).
struct Block {
    pthread_barrier_t bar;
    unsigned finished = 0;                      // threads that have returned from the kernel (atomic)
    long done_at = -1;                          // barrier generation after which the ghosts leave (atomic)
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<unsigned long long> slots;      // 32 per warp
    unsigned n_threads = 0;
    pthread_barrier_t named;                    // `bar.sync 1, n`: a barrier of the first n threads of the block
    unsigned named_n = 0;
    pthread_mutex_t named_mu = PTHREAD_MUTEX_INITIALIZER;
};
inline thread_local long bar_gen = 0;
// where every thread of the running block is waiting (for the watchdog's report of a deadlock)
struct alignas(64) WaitSlot { void* site; int kind; unsigned long passed; };     // one cache line per thread
inline WaitSlot wait_slot[1024];                // kind: 0 running, 1 __syncthreads, 2 warp-level sync, 3 returned (ghost)
struct Waiting {
    WaitSlot& s;
    Waiting(int kind, void* site) : s(wait_slot[threadIdx.x]) { s.site = site; __atomic_store_n(&s.kind, kind, __ATOMIC_RELAXED); }
    ~Waiting() { __atomic_store_n(&s.kind, 0, __ATOMIC_RELAXED); __atomic_store_n(&s.passed, s.passed + 1, __ATOMIC_RELAXED); }
};
inline Block*& cur() { static Block* b = nullptr; return b; }
alignas(128) inline unsigned char dyn_smem[232448];
}  // namespace cusim


static inline __attribute__((always_inline)) void __syncthreads() {
    cusim::Waiting w(1, __builtin_extract_return_addr(__builtin_return_address(0)));
    pthread_barrier_wait(&cusim::cur()->bar); ++cusim::bar_gen;
}
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { cusim::Waiting w(2, __builtin_return_address(0)); pthread_barrier_wait(&cusim::cur()->warp_bar[threadIdx.x >> 5]); }

namespace cusim {
inline void named_barrier(unsigned n) {
    Block* b = cur();
    pthread_mutex_lock(&b->named_mu);
    if (b->named_n == 0) { pthread_barrier_init(&b->named, nullptr, n); b->named_n = n; }
    pthread_mutex_unlock(&b->named_mu);
    pthread_barrier_wait(&b->named);
}
// every lane of the warp deposits v, then reads lane `src`'s value
inline unsigned long long exchange(unsigned long long v, unsigned src) {
    Block* b = cur();
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    Waiting wt(2, __builtin_return_address(0));
    b->slots[w * 32 + lane] = v;
    pthread_barrier_wait(&b->warp_bar[w]);
    const unsigned long long r = b->slots[w * 32 + (src & 31u)];
    pthread_barrier_wait(&b->warp_bar[w]);      // nobody overwrites a slot before everybody has read
    return r;
}
}  // namespace cusim

template <class T> static inline T __shfl_sync(unsigned m, T v, int src) {
    if (m == (1u << (threadIdx.x & 31u))) return v;           // __activemask() of a lone thread
    return (T)cusim::exchange((unsigned long long)v, (unsigned)src);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const unsigned lane = threadIdx.x & 31u;
    const T r = (T)cusim::exchange((unsigned long long)v, lane >= d ? lane - d : lane);
    return lane >= d ? r : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return (T)cusim::exchange((unsigned long long)v, (threadIdx.x & 31u) ^ (unsigned)m); }
static inline unsigned __ballot_sync(unsigned m, int pred) {
    if (m == (1u << (threadIdx.x & 31u))) return pred ? m : 0u;
    unsigned r = 0;
    cusim::Block* b = cusim::cur();
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    cusim::Waiting wt(2, __builtin_return_address(0));
    b->slots[w * 32 + lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&b->warp_bar[w]);
    for (unsigned i = 0; i < 32; ++i) r |= (unsigned)b->slots[w * 32 + i] << i;
    pthread_barrier_wait(&b->warp_bar[w]);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }

// partial masks: the kernels only ever pass the full mask or __activemask(); a thread of the simulator runs on its own,
// so its active mask is its own lane and a vote / shuffle over that mask involves nobody else
static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }
static inline int __syncthreads_or(int pred) {
    static int acc;
    if (threadIdx.x == 0) __atomic_store_n(&acc, 0, __ATOMIC_SEQ_CST);
    __syncthreads();
    if (pred) __atomic_store_n(&acc, 1, __ATOMIC_SEQ_CST);
    __syncthreads();
    const int r = __atomic_load_n(&acc, __ATOMIC_SEQ_CST);
    __syncthreads();
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
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
inline pthread_mutex_t launch_mu = PTHREAD_MUTEX_INITIALIZER;
struct LaunchLock { LaunchLock() { pthread_mutex_lock(&launch_mu); } ~LaunchLock() { pthread_mutex_unlock(&launch_mu); } };
template <class F> void launch(dim3 grid, unsigned block, F body, const char* name = "kernel") {
    LaunchLock one_at_a_time;
    static const bool trace = getenv("CUSIM_TRACE") != nullptr;        // CUSIM_TRACE=1: every launch on stderr (which kernel hangs?)
    if (trace) fprintf(stderr, "[cusim] %s <<<(%u,%u), %u>>>\n", name, grid.x, grid.y, block);                  // (host threads of the library may launch side by side: ok_sets_build_many)
    gridDim = grid; blockDim = dim3(block);
    // one set of `block` OS threads per launch; they run the blocks of the grid one after the other, together
    Block b;
    b.n_threads = block;
    pthread_barrier_init(&b.bar, nullptr, block);
    const unsigned n_warps = (block + 31) / 32;
    b.warp_bar.resize(n_warps);
    b.slots.assign((size_t)n_warps * 32, 0ull);
    for (unsigned w = 0; w < n_warps; ++w) pthread_barrier_init(&b.warp_bar[w], nullptr, std::min(32u, block - w * 32));
    pthread_barrier_t rearm;
    pthread_barrier_init(&rearm, nullptr, block);
    cur() = &b;
    int launch_over = 0;
    for (unsigned t = 0; t < block; ++t) { wait_slot[t].kind = 0; wait_slot[t].passed = 0; }
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
        th.emplace_back([&, t] {
            for (unsigned by = 0; by < grid.y; ++by)
                for (unsigned bx = 0; bx < grid.x; ++bx) {
                    threadIdx = uint3{t, 0, 0}; blockIdx = uint3{bx, by, 0};
                    bar_gen = 0;
                    body();
                    if (__atomic_add_fetch(&b.finished, 1u, __ATOMIC_SEQ_CST) == block) __atomic_store_n(&b.done_at, bar_gen + 1, __ATOMIC_SEQ_CST);
                    __atomic_store_n(&wait_slot[t].kind, 3, __ATOMIC_RELAXED);
                    for (;;) {                                  // ghost: arrive at the barriers the others still run into
                        pthread_barrier_wait(&b.bar); ++bar_gen;
                        if (__atomic_load_n(&b.done_at, __ATOMIC_SEQ_CST) == bar_gen) break;
                    }
                    pthread_barrier_wait(&rearm);               // everybody has seen the end of the block
                    if (t == 0) {                               // re-arm for the next one
                        b.finished = 0; b.done_at = -1;
                        if (b.named_n) { pthread_barrier_destroy(&b.named); b.named_n = 0; }
                    }
                    pthread_barrier_wait(&rearm);
                }
        });
    // watchdog: no barrier passed and no thread running for a while = a deadlock; say where the threads wait
    std::thread watchdog([&] {
        unsigned long last = ~0ul; int still = 0;
        const char* ev = getenv("CUSIM_WATCHDOG_S");
        const int limit = (ev ? atoi(ev) : 120) * 50;                    // in ticks of 20 ms
        while (!__atomic_load_n(&launch_over, __ATOMIC_SEQ_CST)) {
            timespec ts{0, 20000000}; nanosleep(&ts, nullptr);
            unsigned long now = 0; bool running = false;
            for (unsigned t = 0; t < block; ++t) {
                now += __atomic_load_n(&wait_slot[t].passed, __ATOMIC_RELAXED);
                running |= __atomic_load_n(&wait_slot[t].kind, __ATOMIC_RELAXED) == 0;
            }
            if (now != last || running) { last = now; still = 0; continue; }
            if (++still < limit) continue;
            fprintf(stderr, "[cusim] DEADLOCK in %s <<<(%u,%u), %u>>>: every thread has been waiting for %d s\n", name, grid.x, grid.y, block, limit / 50);
            for (unsigned t = 0; t < block; ++t)
                fprintf(stderr, "  thread %4u  %s\n", t,
                        wait_slot[t].kind == 1 ? "__syncthreads" : wait_slot[t].kind == 2 ? "warp-level sync (shuffle / vote / __syncwarp)" : "returned from the kernel (ghost)");
            fflush(stderr);
            abort();
        }
    });
    for (auto& x : th) x.join();
    __atomic_store_n(&launch_over, 1, __ATOMIC_SEQ_CST);
    watchdog.join();
    pthread_barrier_destroy(&b.bar);
    pthread_barrier_destroy(&rearm);
    for (auto& wb : b.warp_bar) pthread_barrier_destroy(&wb);
    cur() = nullptr;
}
}  // namespace cusim

// ================================================================================ runtime API ==
// Everything executes at once, in program order: streams and events are tokens (an event remembers the host clock),
// device memory is host memory.  A correctly synchronised program sees what it would see on the device; a MISSING
// stream dependency is not something this stand-in can show.
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <mutex>
#include <set>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaSimEvent { double ms; };
typedef cudaSimEvent* cudaEvent_t;
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
struct cudaPointerAttributes { cudaMemoryType type; void* devicePointer; void* hostPointer; int device; };
struct cudaIpcMemHandle_t { char reserved[64]; };

namespace cusim {
inline std::mutex& mem_mu() { static std::mutex m; return m; }
inline std::set<const void*>& pinned() { static std::set<const void*> s; return s; }
inline double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
inline void* alloc(size_t n) { void* p = nullptr; if (posix_memalign(&p, 256, (n + 255) & ~(size_t)255 ? (n + 255) & ~(size_t)255 : 256)) return nullptr; return p; }
}  // namespace cusim

template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)cusim::alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(const void* p) { free(const_cast<void*>(p)); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) {
    *p = (T*)cusim::alloc(n);
    if (!*p) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lk(cusim::mem_mu());
    cusim::pinned().insert(*p);
    return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) {
    { std::lock_guard<std::mutex> lk(cusim::mem_mu()); cusim::pinned().erase(p); }
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                            cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; ++r) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cudaSimEvent{cusim::now_ms()}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->ms = cusim::now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : e == cudaErrorMemoryAllocation ? "cudaErrorMemoryAllocation" : "cudaErrorInvalidValue"; }
static inline const char* cudaGetErrorString(cudaError_t e) { return cudaGetErrorName(e); }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    const char* ev = getenv("CUSIM_SMS");
    p->multiProcessorCount = ev ? atoi(ev) : 2;          // few "SMs": grids are sized in multiples of it
    strcpy(p->name, "cusim (CPU stand-in)");
    return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    std::lock_guard<std::mutex> lk(cusim::mem_mu());
    const bool host = cusim::pinned().count(p) != 0;
    a->type = host ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
    a->devicePointer = host ? const_cast<void*>(p) : nullptr;
    a->hostPointer = const_cast<void*>(p);
    a->device = 0;
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T** d, void* h, unsigned) { *d = (T*)h; return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
