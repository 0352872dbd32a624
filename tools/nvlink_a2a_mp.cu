// nvlink_a2a_mp.cu -- the all-to-all of nvlink_a2a.cu with ONE PROCESS PER GPU and CUDA-IPC mapped buffers (what the
// multi-GPU count uses), to tell a mechanism limit from a multi-process / IPC limit.
//   ./nvlink_a2a_mp [mb per pair] [gpus] [alloc_gb]     alloc_gb: size of the receive allocation (the count's is ~20 GB)
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("rank %d CUDA error %s at line %d\n", rank, cudaGetErrorString(e_), __LINE__); fflush(stdout); _exit(1); } } while (0)

struct Shared {
    std::atomic<int> arrive[64];
    cudaIpcMemHandle_t h[8];
    double t[8];
};
static Shared* sh;
static int rank, G;
static int phase = 0;
static void barrier() {
    const int p = phase++;
    sh->arrive[p].fetch_add(1);
    while (sh->arrive[p].load() < G) usleep(50);
}

struct Peers { uint4* p[8]; };
__global__ void __launch_bounds__(512) k_push(Peers remote, const uint4* local, size_t n_per_peer, int me, int g) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int d = 1; d < g; ++d) {
        const int peer = (me + d) % g;
        const uint4* src = local + (size_t)peer * n_per_peer;
        uint4* dst = remote.p[peer] + (size_t)me * n_per_peer;
        size_t i = tid;
        for (; i + 3 * nth < n_per_peer; i += 4 * nth) {
            uint4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = src[i + q * nth];
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[i + q * nth] = v[q];
        }
        for (; i < n_per_peer; i += nth) dst[i] = src[i];
    }
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 512;
    G = argc > 2 ? atoi(argv[2]) : 8;
    const size_t alloc_gb = argc > 3 ? atoi(argv[3]) : 0;
    sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void*)sh, 0, sizeof(Shared));
    rank = 0;
    for (int r = 1; r < G; ++r) { pid_t p = fork(); if (p == 0) { rank = r; break; } }
    CK(cudaSetDevice(rank));
    const size_t n_per_peer = mb * (1 << 20) / 16;
    const size_t bytes = (size_t)G * n_per_peer * 16, recv_bytes = alloc_gb ? alloc_gb << 30 : bytes;
    uint4 *send, *recv;
    CK(cudaMalloc(&send, bytes));
    CK(cudaMalloc(&recv, recv_bytes));
    CK(cudaMemset(send, rank + 1, bytes));
    CK(cudaIpcGetMemHandle(&sh->h[rank], recv));
    barrier();
    Peers pr{};
    for (int h = 0; h < G; ++h) {
        if (h == rank) { pr.p[h] = recv; continue; }
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, sh->h[h], cudaIpcMemLazyEnablePeerAccess));
        pr.p[h] = (uint4*)p;
    }
    cudaStream_t st[8];
    for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    auto run = [&](const char* name, int mode) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaDeviceSynchronize());
            barrier();
            auto t0 = std::chrono::steady_clock::now();
            if (mode == 0) {
                for (int d = 1; d < G; ++d) {
                    const int h = (rank + d) % G;
                    CK(cudaMemcpyAsync(pr.p[h] + (size_t)rank * n_per_peer, send + (size_t)h * n_per_peer, n_per_peer * 16, cudaMemcpyDeviceToDevice, st[h]));
                }
            } else if (mode == 1) {      // the same in 8 pieces per peer, like the count's chunks
                for (int c = 0; c < 8; ++c)
                    for (int d = 1; d < G; ++d) {
                        const int h = (rank + d) % G;
                        const size_t a = n_per_peer * c / 8, b = n_per_peer * (c + 1) / 8;
                        CK(cudaMemcpyAsync(pr.p[h] + (size_t)rank * n_per_peer + a, send + (size_t)h * n_per_peer + a, (b - a) * 16, cudaMemcpyDeviceToDevice, st[h]));
                    }
            } else {
                k_push<<<296, 512, 0, st[0]>>>(pr, send, n_per_peer, rank, G);
            }
            CK(cudaDeviceSynchronize());
            sh->t[rank] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            barrier();
            double mx = 0;
            for (int r = 0; r < G; ++r) mx = sh->t[r] > mx ? sh->t[r] : mx;
            best = mx < best ? mx : best;
            barrier();
        }
        if (rank == 0) {
            printf("%-12s %d processes, IPC  %6.1f MB per pair  %8.3f ms  %7.1f GB/s per GPU and direction\n", name, G, (double)mb, best * 1e3,
                   (double)(G - 1) * n_per_peer * 16 / 1e9 / best);
            fflush(stdout);
        }
    };
    run("ce", 0);
    run("ce 8 pieces", 1);
    run("push kernel", 2);
    barrier();
    if (rank == 0) { for (int r = 1; r < G; ++r) wait(nullptr); }
    _exit(0);
}
