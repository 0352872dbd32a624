// sim_lib_main.cpp -- TEST INFRASTRUCTURE: the whole simulated library (orion_gpu_sim.cpp = the rewritten orion_gpu.cu of
// build_sim.py) in ONE program with a self-checking driver, for the sanitizers:
//   -fsanitize=address,undefined   an index out of range in shared or global memory, in ANY kernel or in the host runtime
//   -fsanitize=thread              conflicting non-atomic accesses between barriers (the device's racecheck)
// The driver goes through the C ABI: a table-path count, the partitioned count of > 2^20 bases with a second batch merged,
// set builds + union + all-vs-all (keyed, forced) + a query by merge (forced); results against the oracle's counter.
#include "orion_gpu_sim.cpp"

#include <map>
#include <random>
#include <set>

extern "C" {
void* orc_counter_create(unsigned k);
void orc_counter_destroy(void* h);
void orc_counter_add_batch(void* h, const uint8_t* bases, const uint64_t* off, uint64_t n, int normalize);
void orc_counter_finish(void* h, uint64_t min_count, uint64_t** keys, uint64_t** counts, uint64_t* n);
void orc_free(void* p);
}

static std::vector<uint8_t> genome(std::mt19937_64& rng, size_t n) {
    std::vector<uint8_t> g(n);
    for (auto& b : g) b = "ACGT"[rng() & 3u];
    return g;
}
static void reads_of(std::mt19937_64& rng, const std::vector<uint8_t>& g, size_t n_reads, std::vector<uint8_t>& bases, std::vector<uint64_t>& off) {
    bases.clear(); off.assign(1, 0);
    for (size_t r = 0; r < n_reads; ++r) {
        const size_t p = rng() % (g.size() - 150);
        for (size_t i = 0; i < 150; ++i) {
            uint8_t b = g[p + i];
            const unsigned e = (unsigned)(rng() % 1000);
            if (e < 5) b = "ACGT"[rng() & 3u]; else if (e == 5) b = 'N'; else if (e == 6) b = (uint8_t)(b + 32);
            bases.push_back(b);
        }
        off.push_back(bases.size());
    }
}
static int check_count(unsigned k, const std::vector<uint8_t>& bases, const std::vector<uint64_t>& off, uint64_t min_count,
                       const uint64_t* keys, const uint64_t* counts, uint64_t n, const char* what) {
    void* o = orc_counter_create(k);
    orc_counter_add_batch(o, bases.data(), off.data(), off.size() - 1, 1);
    uint64_t *wk = nullptr, *wc = nullptr, wn = 0;
    orc_counter_finish(o, min_count, &wk, &wc, &wn);
    int bad = wn != n;
    for (uint64_t i = 0; i < n && !bad; ++i) bad = wk[i] != keys[i] || wc[i] != counts[i];
    printf("%-58s %8llu distinct  %s\n", what, (unsigned long long)n, bad ? "MISMATCH" : "ok");
    orc_free(wk); orc_free(wc); orc_counter_destroy(o);
    return bad;
}
#define MUST(call) do { const int rc_ = (call); if (rc_ != OK_SUCCESS) { printf("%s -> %d: %s\n", #call, rc_, ok_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    const size_t big_reads = argc > 1 ? (size_t)atoi(argv[1]) : 7200;       // x 150 bases: > 2^20 -> the partitioned path
    std::mt19937_64 rng(11);
    int bad = 0;
    MUST(ok_init(nullptr, 0));
    const auto g = genome(rng, 150000);
    std::vector<uint8_t> bases, bases2; std::vector<uint64_t> off, off2;
    // ---- table path
    reads_of(rng, g, 600, bases, off);
    ok_counter* c = nullptr;
    MUST(ok_counter_create(31, OK_NORM_NORMALIZED, 0, &c));
    MUST(ok_counter_add_batch(c, bases.data(), off.data(), off.size() - 1));
    uint64_t *keys = nullptr, *counts = nullptr, n = 0;
    MUST(ok_counter_finish(c, 1, &keys, &counts, &n));
    bad += check_count(31, bases, off, 1, keys, counts, n, "table path, 600 reads");
    ok_free(keys); ok_free(counts);
    MUST(ok_counter_destroy(c));
    // ---- partitioned path + a second batch merged
    reads_of(rng, g, big_reads, bases, off);
    reads_of(rng, g, big_reads, bases2, off2);
    MUST(ok_counter_create(31, OK_NORM_NORMALIZED, 0, &c));
    MUST(ok_counter_add_batch(c, bases.data(), off.data(), off.size() - 1));
    MUST(ok_counter_finish(c, 1, &keys, &counts, &n));
    ok_counter_stats st{};
    MUST(ok_counter_get_stats(c, &st));
    bad += check_count(31, bases, off, 1, keys, counts, n, st.partitioned ? "partitioned path" : "NOT partitioned (batch too small?)");
    bad += st.partitioned ? 0 : 1;
    ok_free(keys); ok_free(counts);
    MUST(ok_counter_add_batch(c, bases2.data(), off2.data(), off2.size() - 1));
    MUST(ok_counter_finish(c, 2, &keys, &counts, &n));
    {
        std::vector<uint8_t> both(bases); both.insert(both.end(), bases2.begin(), bases2.end());
        std::vector<uint64_t> offb(off);
        for (size_t i = 1; i < off2.size(); ++i) offb.push_back(bases.size() + off2[i]);
        bad += check_count(31, both, offb, 2, keys, counts, n, "two batches merged, min_count 2");
    }
    ok_free(keys); ok_free(counts);
    MUST(ok_counter_destroy(c));
    // ---- sets: build, union, all-vs-all (keyed form forced), query by merge (forced)
    setenv("ORION_AVA_KEYED", "1", 1);
    setenv("ORION_PROBE_MERGE", "1", 1);
    std::vector<ok_set*> sets;
    std::vector<std::set<uint64_t>> want_sets;
    for (int s = 0; s < 5; ++s) {
        std::vector<uint8_t> gs(g.begin() + s * 9000, g.begin() + s * 9000 + 40000);
        if (s == 4) gs = genome(rng, 30000);
        const uint64_t o2[2] = {0, gs.size()};
        ok_set* h = nullptr;
        MUST(ok_set_create(21, OK_NORM_NORMALIZED, 0, &h));
        MUST(ok_set_add_batch(h, gs.data(), o2, 1));
        sets.push_back(h);
        void* o = orc_counter_create(21);
        orc_counter_add_batch(o, gs.data(), o2, 1, 1);
        uint64_t *wk = nullptr, *wc = nullptr, wn = 0;
        orc_counter_finish(o, 1, &wk, &wc, &wn);
        want_sets.emplace_back(wk, wk + wn);
        orc_free(wk); orc_free(wc); orc_counter_destroy(o);
    }
    std::vector<uint64_t> sizes(5), inter(25);
    MUST(ok_sets_all_vs_all(sets.data(), 5, sizes.data(), inter.data()));
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            uint64_t w = 0;
            for (auto v : want_sets[i]) w += want_sets[j].count(v);
            if (inter[i * 5 + j] != w) { ++bad; printf("inter[%d][%d] = %llu, want %llu\n", i, j, (unsigned long long)inter[i * 5 + j], (unsigned long long)w); }
        }
    printf("%-58s %s\n", "all-vs-all of 5 sets (keyed form forced)", bad ? "MISMATCH" : "ok");
    ok_set* u = nullptr;
    MUST(ok_set_union(sets.data(), 5, &u));
    std::set<uint64_t> wu;
    for (auto& s : want_sets) wu.insert(s.begin(), s.end());
    uint64_t* uk = nullptr; uint64_t un = 0;
    MUST(ok_set_export(u, &uk, &un));
    bad += un != wu.size() || !std::equal(wu.begin(), wu.end(), uk);
    printf("%-58s %8llu keys      %s\n", "union of 5 sets", (unsigned long long)un, bad ? "MISMATCH" : "ok");
    ok_free(uk);
    reads_of(rng, g, 300, bases, off);
    std::vector<uint32_t> hits(300);
    MUST(ok_probe_reads(u, OK_NORM_RAW, bases.data(), off.data(), 300, hits.data()));
    for (size_t r = 0; r < 300; ++r) {
        uint32_t w = 0;
        for (size_t p = off[r]; p + 21 <= off[r + 1]; ++p) {
            uint64_t v = 0; bool okw = true;
            for (int i = 0; i < 21 && okw; ++i) {
                const uint8_t b = bases[p + i];
                const int code = (b == 'A' || b == 'a') ? 0 : (b == 'C' || b == 'c') ? 1 : (b == 'G' || b == 'g') ? 2 : (b == 'T' || b == 't') ? 3 : -1;
                if (code < 0) okw = false; else v = v * 4 + (uint64_t)code;
            }
            if (!okw) continue;
            uint64_t rc = 0;
            for (int i = 0; i < 21; ++i) rc |= (3ull - ((v >> (2 * i)) & 3ull)) << (2 * (20 - i));
            w += (uint32_t)wu.count(v < rc ? v : rc);
        }
        if (hits[r] != w) { ++bad; if (bad < 5) printf("read %zu: %u hits, want %u\n", r, hits[r], w); }
    }
    printf("%-58s %s\n", "query by merge, 300 reads", bad ? "MISMATCH" : "ok");
    for (auto h : sets) MUST(ok_set_destroy(h));
    MUST(ok_set_destroy(u));
    MUST(ok_shutdown());
    printf("mismatches %d\n", bad);
    return bad ? 1 : 0;
}
