// sim_setops.cpp -- TEST INFRASTRUCTURE: the set-operation kernels of setops.cuh / kernels.cuh executed on the CPU through
// the stand-in cuda_runtime.h of this directory (see its header).  Built by tests/test_kernels_cusim.py from a copy of
// the kernel headers in which only `extern __shared__` is rewritten (dynamic shared memory becomes one global buffer).
#include "cuda_runtime.h"
#include "setops.cuh"

alignas(128) unsigned char ava_smem_raw[232448];

#define SIM_EXPORT extern "C" __attribute__((visibility("default")))

// ok_sets_all_vs_all_part's keyed form (orion_gpu.cu ava_keyed): ends -> geometry (the caller's, from okx_ava_geometry)
// -> bounds -> tiles.  out: n x n, entries i < j.  Returns *failed.
SIM_EXPORT int sim_ava_keyed(const unsigned long long* const* keys, const unsigned long long* ns, unsigned n_sets, unsigned key_shift,
                             uint32_t phi_lo, uint64_t scale, unsigned n_tiles, unsigned grid, unsigned long long* out) {
    static_assert(sizeof(OkAvaSmem) <= sizeof(ava_smem_raw), "dynamic shared memory of k_ava_tiles");
    OkAvaGeo g{};
    g.phi_lo = phi_lo; g.key_shift = key_shift; g.scale = scale; g.n_tiles = n_tiles;
    std::vector<unsigned> bounds((size_t)n_sets * ((size_t)n_tiles + 1), 0xDEADBEEFu);
    uint64_t max_n = 1;
    for (unsigned s = 0; s < n_sets; ++s) max_n = std::max<uint64_t>(max_n, ns[s]);
    const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((max_n + 255) / 256, 3));
    cusim::launch(dim3(gx, n_sets), 256, [&] { k_ava_bounds(keys, ns, g, bounds.data()); });
    for (unsigned v : bounds) if (v == 0xDEADBEEFu) return -1;          // a bound nobody wrote
    unsigned failed = 0;
    cusim::launch(dim3(std::min(grid, n_tiles)), OK_AVA_THREADS, [&] { k_ava_tiles(keys, n_sets, g, bounds.data(), out, &failed); });
    return (int)failed;
}

// probe_reads_merge's steps 2: which keys of a (sorted) occur in b (sorted); returns the number of matches written to out
SIM_EXPORT uint64_t sim_member(const unsigned long long* a, uint64_t na, const unsigned long long* b, uint64_t nb, unsigned grid,
                               unsigned long long* out) {
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE;
    std::vector<unsigned long long> lo(n_tiles + 1);
    cusim::launch(dim3(1), 256, [&] { k_intersect_bounds(a, na, b, nb, lo.data()); });
    unsigned long long n_out = 0;
    cusim::launch(dim3(grid), 256, [&] { k_member_tiled(a, na, b, lo.data(), out, &n_out); });
    return n_out;
}

// compare.rs:58 for one pair through the tiled kernel
SIM_EXPORT uint64_t sim_intersect(const unsigned long long* a, uint64_t na, const unsigned long long* b, uint64_t nb, unsigned grid) {
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE;
    std::vector<unsigned long long> lo(n_tiles + 1);
    cusim::launch(dim3(1), 256, [&] { k_intersect_bounds(a, na, b, nb, lo.data()); });
    unsigned long long m = 0;
    cusim::launch(dim3(grid), 256, [&] { k_intersect_tiled(a, na, b, lo.data(), &m); });
    return m;
}
