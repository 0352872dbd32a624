// sim_main.cpp -- TEST INFRASTRUCTURE: a self-checking run of the simulated set-operation kernels, for the sanitizers:
//   g++ -std=c++17 -O1 -g -D__CUDACC__ -fsanitize=thread            ... -> data races between barriers
//   g++ -std=c++17 -O1 -g -D__CUDACC__ -fsanitize=address,undefined ... -> out-of-bounds accesses, UB
// (tests/test_kernels_cusim.py builds and runs both when the compiler supports them.)
#include "sim_setops.cpp"

#include <cstdio>
#include <iterator>
#include <random>

static OkAvaGeo geometry(unsigned k, const std::vector<std::vector<unsigned long long>>& sets, uint64_t total) {
    // as ava_geometry (orion_gpu.cu): the span of positions the sets cover, ~OK_AVA_TARGET keys per tile
    OkAvaGeo g{};
    g.key_shift = 64u - 2u * k;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (auto& s : sets) {
        if (s.empty()) continue;
        lo = std::min(lo, ok_phi32(s.front(), g.key_shift));
        hi = std::max(hi, ok_phi32(s.back(), g.key_shift));
    }
    if (lo > hi) { lo = 0; hi = 0; }
    const uint64_t span = (uint64_t)hi - lo + 1u;
    uint64_t tiles = std::max<uint64_t>(1, total / OK_AVA_TARGET);
    tiles = std::min<uint64_t>(tiles, std::min<uint64_t>(span, 1ull << 22));
    g.phi_lo = lo; g.n_tiles = (unsigned)tiles; g.scale = (tiles << 32) / span;
    return g;
}

int main(int argc, char** argv) {
    const unsigned n_sets = argc > 1 ? (unsigned)atoi(argv[1]) : 11u;
    const unsigned pool_n = argc > 2 ? (unsigned)atoi(argv[2]) : 12000u;
    const unsigned k = 21;
    std::mt19937_64 rng(7);
    std::vector<unsigned long long> pool;
    for (unsigned i = 0; i < pool_n; ++i) { const unsigned long long a = rng() >> 22, b = rng() >> 22; pool.push_back(std::min(a, b)); }
    std::sort(pool.begin(), pool.end());
    pool.erase(std::unique(pool.begin(), pool.end()), pool.end());
    std::vector<std::vector<unsigned long long>> sets(n_sets);
    uint64_t total = 0;
    for (unsigned s = 0; s < n_sets; ++s) {
        const double f = s == 3 ? 0.0 : (s % 5 + 1) / 6.0;
        for (auto v : pool) if ((rng() >> 11) * (1.0 / 9007199254740992.0) < f) sets[s].push_back(v);
        total += sets[s].size();
    }
    std::vector<const unsigned long long*> ptrs(n_sets);
    std::vector<unsigned long long> ns(n_sets);
    for (unsigned s = 0; s < n_sets; ++s) { ptrs[s] = sets[s].data(); ns[s] = sets[s].size(); }
    const OkAvaGeo g = geometry(k, sets, total);
    std::vector<unsigned long long> out((size_t)n_sets * n_sets, 0ull);
    const int failed = sim_ava_keyed(ptrs.data(), ns.data(), n_sets, g.key_shift, g.phi_lo, g.scale, g.n_tiles, 2, out.data());
    int bad = failed != 0;
    for (unsigned i = 0; i < n_sets; ++i)
        for (unsigned j = 0; j < n_sets; ++j) {
            std::vector<unsigned long long> both;
            std::set_intersection(sets[i].begin(), sets[i].end(), sets[j].begin(), sets[j].end(), std::back_inserter(both));
            const unsigned long long want = i < j ? both.size() : 0ull;
            if (out[(size_t)i * n_sets + j] != want) ++bad;
            if (i == 1 && j == 2) {
                if (sim_intersect(sets[i].data(), sets[i].size(), sets[j].data(), sets[j].size(), 2) != both.size()) ++bad;
                std::vector<unsigned long long> got(sets[i].size() + 1);
                const uint64_t m = sim_member(sets[i].data(), sets[i].size(), sets[j].data(), sets[j].size(), 2, got.data());
                got.resize(m);
                std::sort(got.begin(), got.end());
                if (got != both) ++bad;
            }
        }
    printf("sets %u keys %llu tiles %u failed %d mismatches %d\n", n_sets, (unsigned long long)total, g.n_tiles, failed, bad);
    return bad ? 1 : 0;
}
