// microbench.cu -- primitive rates that decide the shape of the partition + count kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

// 1. shared-memory atomicAdd with return, random bins
template <int BINS>
__global__ void k_smem_atom(unsigned* out, int iters) {
    __shared__ unsigned hist[BINS];
    for (int i = threadIdx.x; i < BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    uint64_t s = mix(blockIdx.x * 1024ull + threadIdx.x);
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; acc += atomicAdd(&hist[(s >> 40) % BINS], 1u); }
    if (acc == 0xFFFFFFFFu) out[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = hist[0];
}
// 2. match_any based ranking into warp-private histograms (no atomics)
template <int BINS>
__global__ void k_match(unsigned* out, int iters) {
    __shared__ unsigned hist[8][BINS];
    for (int i = threadIdx.x; i < 8 * BINS; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t s = mix(blockIdx.x * 1024ull + threadIdx.x);
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        unsigned b = (s >> 40) % BINS;
        unsigned m = __match_any_sync(0xFFFFFFFFu, b);
        unsigned rank = __popc(m & ((1u << lane) - 1));
        unsigned base = hist[w][b];
        __syncwarp();
        if (rank == 0) hist[w][b] = base + __popc(m);
        __syncwarp();
        acc += base + rank;
    }
    if (acc == 0xFFFFFFFFu) out[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = hist[0][0];
}
// 3. table updates on a region of `slots` 16-byte slots: load key + RED count (the hit path of the count table)
__global__ void k_table(unsigned long long* tab, uint64_t slots, uint64_t n, int mode) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h = mix(i) % slots;
        if (mode == 0) {            // load + RED
            unsigned long long k = __ldcg(&tab[2 * h]);
            if (k != 12345ull) atomicAdd(&tab[2 * h + 1], 1ull);
        } else if (mode == 1) {     // RED only
            atomicAdd(&tab[2 * h + 1], 1ull);
        } else if (mode == 2) {     // CAS + RED
            unsigned long long k = atomicCAS(&tab[2 * h], 0ull, i + 1);
            if (k != 12345ull) atomicAdd(&tab[2 * h + 1], 1ull);
        } else {                    // load only
            unsigned long long k = __ldcg(&tab[2 * h]);
            if (k == 12345ull) tab[2 * h + 1] = 1;
        }
    }
}
// 4. streaming 8-byte scatter into P partitions through per-partition cursors held in smem-staged runs is the real
//    kernel; here just the coalesced copy rate for reference
__global__ void k_copy(const uint4* a, uint4* b, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    unsigned* out; CK(cudaMalloc(&out, 1 << 20));
    const int iters = 4096, blocks = p.multiProcessorCount * 8;
    {
        float ms = timeit([&] { k_smem_atom<256><<<blocks, 256>>>(out, iters); });
        printf("smem atomicAdd(ret) 256 bins : %.3f ms  %.1f Gop/s\n", ms, (double)blocks * 256 * iters / ms / 1e6);
        ms = timeit([&] { k_smem_atom<1024><<<blocks, 256>>>(out, iters); });
        printf("smem atomicAdd(ret) 1024 bins: %.3f ms  %.1f Gop/s\n", ms, (double)blocks * 256 * iters / ms / 1e6);
        ms = timeit([&] { k_smem_atom<16><<<blocks, 256>>>(out, iters); });
        printf("smem atomicAdd(ret) 16 bins  : %.3f ms  %.1f Gop/s\n", ms, (double)blocks * 256 * iters / ms / 1e6);
        ms = timeit([&] { k_match<256><<<blocks, 256>>>(out, iters); });
        printf("match_any rank 256 bins      : %.3f ms  %.1f Gop/s\n", ms, (double)blocks * 256 * iters / ms / 1e6);
        ms = timeit([&] { k_match<1024><<<blocks, 256>>>(out, iters); });
        printf("match_any rank 1024 bins     : %.3f ms  %.1f Gop/s\n", ms, (double)blocks * 256 * iters / ms / 1e6);
    }
    CK(cudaGetLastError());
    const uint64_t max_slots = 1ull << 29;  // 8 GB
    unsigned long long* tab; CK(cudaMalloc(&tab, max_slots * 16)); CK(cudaMemset(tab, 0, max_slots * 16));
    const uint64_t n = 1ull << 28;
    const char* names[4] = {"ld+RED", "RED", "CAS+RED", "ld"};
    for (uint64_t slots : {1ull << 20, 1ull << 21, 1ull << 22, 1ull << 23, 1ull << 25, 1ull << 29}) {
        for (int mode = 0; mode < 4; ++mode) {
            float ms = timeit([&] { k_table<<<p.multiProcessorCount * 16, 256>>>(tab, slots, n, mode); }, 3);
            printf("table %7.0f MB %-8s: %8.3f ms  %6.1f Gop/s\n", slots * 16.0 / 1e6, names[mode], ms, n / ms / 1e6);
        }
    }
    uint64_t nc = (4ull << 30) / 16;
    uint4 *a, *b; CK(cudaMalloc(&a, nc * 16)); CK(cudaMalloc(&b, nc * 16));
    float ms = timeit([&] { k_copy<<<p.multiProcessorCount * 16, 256>>>(a, b, nc); });
    printf("copy 4 GB: %.3f ms  %.0f GB/s (r+w)\n", ms, 2.0 * nc * 16 / ms / 1e6);
    CK(cudaGetLastError());
    return 0;
}
