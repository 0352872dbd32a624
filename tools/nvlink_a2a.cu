// nvlink_a2a.cu -- how fast can 8 B200s exchange data all-to-all over NVLink 5 / NVSwitch, and by which mechanism?
//
// One process, every visible GPU, peer access enabled.  Every GPU sends `mb` MB to each of its peers, all at once:
//   ce<k>   cudaMemcpyPeerAsync, k streams per (source, destination) pair          (copy engines)
//   push    a kernel on every source writes into the peers' buffers, 16-byte stores  (SM-issued posted writes)
//   pull    a kernel on every destination reads the peers' buffers, 16-byte loads    (SM-issued remote loads)
// Prints GB/s per GPU and direction (bytes a GPU sends / time of the whole exchange).  The multi-GPU count's
// exchange is designed from these numbers.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_a2a tools/nvlink_a2a.cu && ./nvlink_a2a [mb] [gpus]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct Peers { uint4* p[8]; };

// grid-stride copy of n 16-byte words to (push) or from (pull) every peer; slot = which 1/G of the remote buffer is ours
template <bool PULL, int UNROLL>
__global__ void __launch_bounds__(512) k_a2a(Peers remote, uint4* local, size_t n_per_peer, int me, int G) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (int d = 1; d < G; ++d) {
        const int peer = (me + d) % G;
        // local layout: [G slots of n_per_peer]; slot `peer` = what I send to / receive from `peer`
        uint4* mine = local + (size_t)peer * n_per_peer;
        uint4* theirs = remote.p[peer] + (size_t)me * n_per_peer;
        const uint4* src = PULL ? theirs : mine;
        uint4* dst = PULL ? mine : theirs;
        size_t i = tid;
        for (; i + (UNROLL - 1) * nth < n_per_peer; i += UNROLL * nth) {
            uint4 v[UNROLL];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) v[q] = src[i + q * nth];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) dst[i + q * nth] = v[q];
        }
        for (; i < n_per_peer; i += nth) dst[i] = src[i];
    }
}

int main(int argc, char** argv) {
    const size_t mb = argc > 1 ? atoi(argv[1]) : 512;
    int G = 0;
    CK(cudaGetDeviceCount(&G));
    if (argc > 2) G = std::min(G, atoi(argv[2]));
    if (G < 2) { printf("needs >= 2 GPUs\n"); return 0; }
    const size_t n_per_peer = mb * (1 << 20) / 16;
    std::vector<uint4*> send(G), recv(G);
    std::vector<std::vector<cudaStream_t>> st(G);
    for (int g = 0; g < G; ++g) {
        CK(cudaSetDevice(g));
        for (int h = 0; h < G; ++h) if (h != g) { cudaError_t e = cudaDeviceEnablePeerAccess(h, 0); if (e != cudaSuccess) cudaGetLastError(); }
        CK(cudaMalloc(&send[g], (size_t)G * n_per_peer * 16));
        CK(cudaMalloc(&recv[g], (size_t)G * n_per_peer * 16));
        CK(cudaMemset(send[g], g + 1, (size_t)G * n_per_peer * 16));
        st[g].resize(G * 4 + 1);
        for (auto& s : st[g]) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
    auto sync_all = [&] { for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); } };
    auto report = [&](const char* name, double s) {
        const double gb = (double)(G - 1) * n_per_peer * 16 / 1e9;
        printf("%-10s %d GPUs  %6.1f MB per pair  %8.3f ms  %7.1f GB/s per GPU and direction\n", name, G, (double)mb, s * 1e3, gb / s);
        fflush(stdout);
    };
    for (int rep = 0; rep < 2; ++rep) {
        for (int k : {1, 2, 4}) {
            sync_all();
            auto t0 = std::chrono::steady_clock::now();
            for (int g = 0; g < G; ++g) {
                CK(cudaSetDevice(g));
                for (int d = 1; d < G; ++d) {
                    const int h = (g + d) % G;
                    const size_t piece = (n_per_peer + k - 1) / k;
                    for (int q = 0; q < k; ++q) {
                        const size_t a = std::min(n_per_peer, q * piece), b = std::min(n_per_peer, (q + 1) * piece);
                        if (b > a) CK(cudaMemcpyPeerAsync(recv[h] + (size_t)g * n_per_peer + a, h, send[g] + (size_t)h * n_per_peer + a, g, (b - a) * 16, st[g][h * 4 + q]));
                    }
                }
            }
            sync_all();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            char name[16]; snprintf(name, sizeof name, "ce%d", k);
            if (rep) report(name, s);
        }
        for (int mode = 0; mode < 4; ++mode) {
            for (int ctas : {148, 296, 592}) {
                sync_all();
                auto t0 = std::chrono::steady_clock::now();
                for (int g = 0; g < G; ++g) {
                    CK(cudaSetDevice(g));
                    Peers pr{};
                    for (int h = 0; h < G; ++h) pr.p[h] = mode & 1 ? send[h] : recv[h];      // pull reads the peers' SEND buffers
                    uint4* local = mode & 1 ? recv[g] : send[g];
                    cudaStream_t s0 = st[g][G * 4];
                    if (mode == 0) k_a2a<false, 4><<<ctas, 512, 0, s0>>>(pr, local, n_per_peer, g, G);
                    if (mode == 1) k_a2a<true, 4><<<ctas, 512, 0, s0>>>(pr, local, n_per_peer, g, G);
                    if (mode == 2) k_a2a<false, 8><<<ctas, 512, 0, s0>>>(pr, local, n_per_peer, g, G);
                    if (mode == 3) k_a2a<true, 8><<<ctas, 512, 0, s0>>>(pr, local, n_per_peer, g, G);
                }
                sync_all();
                const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                char name[24]; snprintf(name, sizeof name, "%s%d/%d", mode & 1 ? "pull" : "push", mode & 2 ? 8 : 4, ctas);
                if (rep) report(name, s);
            }
        }
    }
    // one pair only, for reference (what the profiling guide quotes: ~770 GB/s)
    for (int k : {1, 4}) {
        sync_all();
        auto t0 = std::chrono::steady_clock::now();
        CK(cudaSetDevice(0));
        const size_t tot = (size_t)(G - 1) * n_per_peer, piece = (tot + k - 1) / k;
        for (int q = 0; q < k; ++q) {
            const size_t a = std::min(tot, q * piece), b = std::min(tot, (q + 1) * piece);
            if (b > a) CK(cudaMemcpyPeerAsync(recv[1] + a, 1, send[0] + a, 0, (b - a) * 16, st[0][q]));
        }
        sync_all();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        char name[16]; snprintf(name, sizeof name, "pair ce%d", k);
        report(name, s);
    }
    return 0;
}
