// orion_oracle.cpp -- TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement of the reference's (motroy/orion-kmer) k-mer hot path, kept
// deliberately reference-shaped: every window is re-encoded from scratch (O(k)),
// reverse-complemented with a second O(k) loop, and pushed into a hash-map sink.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may load this library; the shipped CUDA path never does.
//
// PARITY PINNING: the Rust reference cannot be compiled in this image (no
// cargo/rustc, no vendored crates), so this restatement is pinned against the
// reference's own known-answer vectors instead (tests/golden/reference_vectors.json,
// copied by hand from orion-kmer/src/kmer.rs:108-341 and orion-kmer/tests/*.rs --
// see tests/test_oracle_golden.py).  The FASTA/FASTQ framing and normalize(false)
// live in the un-vendored crate needletail 0.5.1 (Cargo.lock:580-591); their
// behaviour is restated from its published semantics.  The parts of it that no
// valid reference test exercises (multi-line FASTA joining, U->T, CRLF, raw-mode
// newlines in `query`) are "parity unpinned".
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/orion-kmer/).

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- kmer.rs ---

// src/kmer.rs:12-20  dna_base_to_u64
inline bool base_code(uint8_t b, uint64_t* out) {
    switch (b) {
        case 'A': case 'a': *out = 0; return true;
        case 'C': case 'c': *out = 1; return true;
        case 'G': case 'g': *out = 2; return true;
        case 'T': case 't': *out = 3; return true;
        default: return false;
    }
}

// src/kmer.rs:37-57  seq_to_u64 (first base in the most significant used bits)
inline bool seq_to_u64(const uint8_t* seq, size_t len, unsigned k, uint64_t* out) {
    if (k == 0 || k > 32) return false;
    if (len != k) return false;
    uint64_t v = 0;
    for (unsigned i = 0; i < k; ++i) {
        uint64_t c;
        if (!base_code(seq[i], &c)) return false;
        v |= c << (2 * (k - 1 - i));
    }
    *out = v;
    return true;
}

// src/kmer.rs:79-94  reverse_complement_u64
inline uint64_t reverse_complement_u64(uint64_t v, unsigned k) {
    uint64_t rc = 0;
    for (unsigned i = 0; i < k; ++i) {
        uint64_t b = (v >> (2 * i)) & 3u;
        rc |= (b ^ 3u) << (2 * (k - 1 - i));
    }
    return rc;
}

// src/kmer.rs:99-106  canonical_u64
inline uint64_t canonical_u64(uint64_t v, unsigned k) {
    uint64_t rc = reverse_complement_u64(v, k);
    return v < rc ? v : rc;
}

// --------------------------------------------------------------- hash sink ---
// Stands in for DashMap<u64, AtomicUsize> (count.rs:48), DashSet<u64> (build.rs:95)
// and std HashMap/HashSet.  Iteration order is never observable in the reference
// (outputs are sorted or are integer sums), so any exact map is equivalent.
// SipHash-1-3 is used because it is what Rust's RandomState computes per lookup,
// which keeps the timed CPU baseline honest.

inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }

inline uint64_t siphash13_u64(uint64_t m) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;  // key = (0,0)
#define SIPROUND                                                     \
    do {                                                             \
        v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);    \
        v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;                       \
        v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;                       \
        v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);    \
    } while (0)
    v3 ^= m; SIPROUND; v0 ^= m;
    uint64_t b = 8ULL << 56;
    v3 ^= b; SIPROUND; v0 ^= b;
    v2 ^= 0xff; SIPROUND; SIPROUND; SIPROUND;
#undef SIPROUND
    return v0 ^ v1 ^ v2 ^ v3;
}

struct CountMap {
    std::vector<uint64_t> keys, vals;
    std::vector<uint8_t> used;
    size_t n = 0, mask = 0;
    explicit CountMap(size_t cap_pow2 = 1024) { init(cap_pow2); }
    void init(size_t cap) {
        keys.assign(cap, 0); vals.assign(cap, 0); used.assign(cap, 0);
        mask = cap - 1; n = 0;
    }
    void grow() {
        std::vector<uint64_t> ok, ov; std::vector<uint8_t> ou;
        ok.swap(keys); ov.swap(vals); ou.swap(used);
        init(ok.size() * 2);
        for (size_t i = 0; i < ok.size(); ++i) if (ou[i]) add(ok[i], ov[i]);
    }
    inline void add(uint64_t key, uint64_t by) {
        if ((n + 1) * 8 > (mask + 1) * 7) grow();
        size_t i = siphash13_u64(key) & mask;
        while (used[i] && keys[i] != key) i = (i + 1) & mask;
        if (!used[i]) { used[i] = 1; keys[i] = key; vals[i] = 0; ++n; }
        vals[i] += by;
    }
    inline bool find(uint64_t key, uint64_t* v) const {
        size_t i = siphash13_u64(key) & mask;
        while (used[i]) {
            if (keys[i] == key) { if (v) *v = vals[i]; return true; }
            i = (i + 1) & mask;
        }
        return false;
    }
};

// count.rs:23-38  process_sequence_chunk ; build.rs:50-58 ; classify.rs:167-174
inline void process_sequence_chunk(const uint8_t* seq, size_t len, unsigned k, CountMap& m) {
    if (len < k) return;
    for (size_t i = 0; i + k <= len; ++i) {
        uint64_t v;
        if (seq_to_u64(seq + i, k, k, &v)) m.add(canonical_u64(v, k), 1);
    }
}

// ------------------------------------------------- needletail (restated) ---

// needletail 0.5.1 sequence::normalize(seq, iupac=false):  ACGTN- kept; acgtn
// upper-cased; u/U -> T; '.' '~' -> '-'; space, tab, CR, LF removed; every other
// byte -> N.  Call sites: count.rs:71, build.rs:48, classify.rs:165.
size_t normalize_into(const uint8_t* in, size_t len, uint8_t* out) {
    size_t o = 0;
    for (size_t i = 0; i < len; ++i) {
        uint8_t c = in[i], r;
        switch (c) {
            case 'A': case 'C': case 'G': case 'T': case 'N': case '-': r = c; break;
            case 'a': r = 'A'; break;
            case 'c': r = 'C'; break;
            case 'g': r = 'G'; break;
            case 't': r = 'T'; break;
            case 'n': r = 'N'; break;
            case 'u': case 'U': r = 'T'; break;
            case '.': case '~': r = '-'; break;
            case ' ': case '\t': case '\r': case '\n': continue;
            default: r = 'N'; break;
        }
        out[o++] = r;
    }
    return o;
}

struct Fastx {
    // record i: id = buf[id_off[i] .. id_off[i]+id_len[i]), raw sequence (FASTA: with
    // embedded line breaks, as needletail's Sequence::sequence() hands it out)
    std::vector<uint64_t> id_off, id_len, seq_off, seq_len;
    int error = 0;  // 0 ok, 1 empty file, 2 invalid start byte, 3 truncated / malformed FASTQ
};

inline size_t line_end(const uint8_t* b, size_t len, size_t p) {
    const void* q = memchr(b + p, '\n', len - p);
    return q ? (size_t)((const uint8_t*)q - b) : len;
}

// needletail parse_fastx_reader: format by first byte ('>' FASTA, '@' FASTQ), FASTA
// records may span lines, FASTQ is strict 4-line.  Call sites count.rs:63, build.rs:42,
// query.rs:51, classify.rs:150.  (Compressed-stream sniffing is host I/O, out of scope.)
void parse_fastx(const uint8_t* b, size_t len, Fastx& fx) {
    if (len == 0) { fx.error = 1; return; }
    if (b[0] == '>') {
        size_t p = 0;
        while (p < len) {  // b[p] == '>' at a line start
            size_t he = line_end(b, len, p);
            size_t id0 = p + 1, id1 = he;
            if (id1 > id0 && b[id1 - 1] == '\r') --id1;
            size_t s0 = he < len ? he + 1 : len;
            // the sequence runs up to the next line that starts with '>' (or EOF)
            size_t nx = s0;
            while (nx < len && b[nx] != '>') { size_t le = line_end(b, len, nx); nx = le < len ? le + 1 : len; }
            size_t e = nx;
            while (e > s0 && (b[e - 1] == '\n' || b[e - 1] == '\r')) --e;  // trailing EOL is not sequence
            fx.id_off.push_back(id0); fx.id_len.push_back(id1 - id0);
            fx.seq_off.push_back(s0); fx.seq_len.push_back(e - s0);
            p = nx;
        }
    } else if (b[0] == '@') {
        size_t p = 0;
        while (p < len) {
            if (b[p] == '\n' || b[p] == '\r') { ++p; continue; }  // trailing blank lines
            if (b[p] != '@') { fx.error = 3; return; }
            size_t he = line_end(b, len, p);
            if (he >= len) { fx.error = 3; return; }
            size_t id0 = p + 1, id1 = he;
            if (id1 > id0 && b[id1 - 1] == '\r') --id1;
            size_t s0 = he + 1, se = line_end(b, len, s0);
            if (se >= len) { fx.error = 3; return; }
            size_t s1 = se; if (s1 > s0 && b[s1 - 1] == '\r') --s1;
            size_t pl = se + 1;
            if (pl >= len || b[pl] != '+') { fx.error = 3; return; }
            size_t pe = line_end(b, len, pl);
            if (pe >= len) { fx.error = 3; return; }
            size_t q0 = pe + 1, qe = line_end(b, len, q0);
            size_t q1 = qe; if (q1 > q0 && b[q1 - 1] == '\r') --q1;
            if (q1 - q0 != s1 - s0) { fx.error = 3; return; }
            fx.id_off.push_back(id0); fx.id_len.push_back(id1 - id0);
            fx.seq_off.push_back(s0); fx.seq_len.push_back(s1 - s0);
            p = qe < len ? qe + 1 : len;
        }
    } else {
        fx.error = 2;
    }
}

struct Counter { unsigned k; CountMap map; };

}  // namespace

// ===================================================================== C API ==
extern "C" {

// ---- src/kmer.rs public functions -------------------------------------------
int orc_seq_to_u64(const uint8_t* seq, uint64_t len, unsigned k, uint64_t* out) {
    return seq_to_u64(seq, (size_t)len, k, out) ? 1 : 0;
}
// src/kmer.rs:61-75 u64_to_seq ; returns 0 where the reference panics (k out of range)
int orc_u64_to_seq(uint64_t v, unsigned k, uint8_t* out) {
    if (k == 0 || k > 32) return 0;
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (unsigned i = 0; i < k; ++i) out[i] = (uint8_t)L[(v >> (2 * (k - 1 - i))) & 3u];
    return 1;
}
int orc_reverse_complement_u64(uint64_t v, unsigned k, uint64_t* out) {
    if (k == 0 || k > 32) return 0;  // reference panics (kmer.rs:80-82)
    *out = reverse_complement_u64(v, k); return 1;
}
int orc_canonical_u64(uint64_t v, unsigned k, uint64_t* out) {
    if (k == 0 || k > 32) return 0;
    *out = canonical_u64(v, k); return 1;
}

uint64_t orc_normalize(const uint8_t* in, uint64_t len, uint8_t* out) {
    return normalize_into(in, (size_t)len, out);
}

// ---- FASTA/FASTQ framing ------------------------------------------------------
void* orc_fastx_parse(const uint8_t* buf, uint64_t len) {
    Fastx* fx = new Fastx(); parse_fastx(buf, (size_t)len, *fx); return fx;
}
int orc_fastx_error(void* h) { return ((Fastx*)h)->error; }
uint64_t orc_fastx_n(void* h) { return ((Fastx*)h)->seq_off.size(); }
void orc_fastx_record(void* h, uint64_t i, uint64_t* id_off, uint64_t* id_len,
                      uint64_t* seq_off, uint64_t* seq_len) {
    Fastx* fx = (Fastx*)h;
    *id_off = fx->id_off[i]; *id_len = fx->id_len[i];
    *seq_off = fx->seq_off[i]; *seq_len = fx->seq_len[i];
}
void orc_fastx_free(void* h) { delete (Fastx*)h; }

// ---- count (count.rs:40-141) ---------------------------------------------------
// returns NULL for k outside 1..=32 (count.rs:43-45 -> InvalidKmerSize)
void* orc_counter_create(unsigned k) {
    if (k == 0 || k > 32) return nullptr;
    Counter* c = new Counter{k, CountMap(1 << 16)}; return c;
}
void orc_counter_destroy(void* h) { delete (Counter*)h; }
// one already-normalized record (count.rs:72)
void orc_counter_add_seq(void* h, const uint8_t* seq, uint64_t len) {
    Counter* c = (Counter*)h; process_sequence_chunk(seq, (size_t)len, c->k, c->map);
}
// a batch in the C-ABI layout: concatenated bases + n+1 offsets.  normalize!=0 applies
// needletail normalize(false) per record first (count/build/classify semantics);
// normalize==0 feeds the raw bytes (query semantics).
void orc_counter_add_batch(void* h, const uint8_t* bases, const uint64_t* off, uint64_t n,
                           int normalize) {
    Counter* c = (Counter*)h;
    std::vector<uint8_t> tmp;
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* s = bases + off[r]; size_t len = (size_t)(off[r + 1] - off[r]);
        if (normalize) {
            tmp.resize(len); size_t m = normalize_into(s, len, tmp.data());
            process_sequence_chunk(tmp.data(), m, c->k, c->map);
        } else {
            process_sequence_chunk(s, len, c->k, c->map);
        }
    }
}
// whole FASTA/FASTQ file content (count.rs:63-79).  Returns parser error code.
int orc_counter_add_fastx(void* h, const uint8_t* buf, uint64_t len) {
    Counter* c = (Counter*)h; Fastx fx; parse_fastx(buf, (size_t)len, fx);
    if (fx.error) return fx.error;
    std::vector<uint8_t> tmp;
    for (size_t r = 0; r < fx.seq_off.size(); ++r) {
        tmp.resize(fx.seq_len[r]);
        size_t m = normalize_into(buf + fx.seq_off[r], fx.seq_len[r], tmp.data());
        process_sequence_chunk(tmp.data(), m, c->k, c->map);
    }
    return 0;
}
uint64_t orc_counter_distinct(void* h) { return ((Counter*)h)->map.n; }
// count.rs:106-119: keep count >= min_count, sort ascending by key.  Caller frees
// with orc_free.
void orc_counter_finish(void* h, uint64_t min_count, uint64_t** keys, uint64_t** counts,
                        uint64_t* n) {
    Counter* c = (Counter*)h;
    std::vector<std::pair<uint64_t, uint64_t>> v; v.reserve(c->map.n);
    for (size_t i = 0; i <= c->map.mask; ++i)
        if (c->map.used[i] && c->map.vals[i] >= min_count) v.emplace_back(c->map.keys[i], c->map.vals[i]);
    std::sort(v.begin(), v.end());
    *n = v.size();
    *keys = (uint64_t*)malloc(sizeof(uint64_t) * (v.size() + 1));
    *counts = (uint64_t*)malloc(sizeof(uint64_t) * (v.size() + 1));
    for (size_t i = 0; i < v.size(); ++i) { (*keys)[i] = v[i].first; (*counts)[i] = v[i].second; }
}
void orc_free(void* p) { free(p); }

// count.rs:127-135: "KMER\tcount\n" per line.  Returns bytes written (buffer must hold
// n*(k+22)).
uint64_t orc_format_counts(const uint64_t* keys, const uint64_t* counts, uint64_t n, unsigned k,
                           char* out) {
    char* p = out;
    for (uint64_t i = 0; i < n; ++i) {
        orc_u64_to_seq(keys[i], k, (uint8_t*)p); p += k; *p++ = '\t';
        p += sprintf(p, "%llu", (unsigned long long)counts[i]); *p++ = '\n';
    }
    return (uint64_t)(p - out);
}

// ---- set algebra ----------------------------------------------------------------
// db_types.rs:43-48 get_all_kmers_unified over sorted-unique arrays; out must hold sum(n)
uint64_t orc_set_union(const uint64_t* const* sets, const uint64_t* ns, uint64_t n_sets,
                       uint64_t* out) {
    CountMap m(1 << 16);
    for (uint64_t s = 0; s < n_sets; ++s) for (uint64_t i = 0; i < ns[s]; ++i) m.add(sets[s][i], 1);
    uint64_t o = 0;
    for (size_t i = 0; i <= m.mask; ++i) if (m.used[i]) out[o++] = m.keys[i];
    std::sort(out, out + o);
    return o;
}
// compare.rs:51-66: out[0]=|A| out[1]=|B| out[2]=|A n B| out[3]=|A u B|; returns Jaccard
double orc_compare(const uint64_t* a, uint64_t na, const uint64_t* b, uint64_t nb, uint64_t* out) {
    CountMap mb(1 << 16);
    for (uint64_t i = 0; i < nb; ++i) mb.add(b[i], 1);
    uint64_t inter = 0;
    for (uint64_t i = 0; i < na; ++i) if (mb.find(a[i], nullptr)) ++inter;   // compare.rs:58
    uint64_t uni = na + nb - inter;                                           // compare.rs:60
    out[0] = na; out[1] = nb; out[2] = inter; out[3] = uni;
    return uni == 0 ? 0.0 : (double)inter / (double)uni;                      // compare.rs:62-66
}

// query.rs:79-108: per read, windows whose canonical k-mer is in the set (raw sequence,
// windows not de-duplicated).  `set` sorted-unique.  n_threads mirrors rayon par_iter.
void orc_query_hits(const uint64_t* set, uint64_t nset, unsigned k, const uint8_t* bases,
                    const uint64_t* off, uint64_t n_reads, uint64_t* hits, int n_threads) {
    CountMap m(1 << 16);
    for (uint64_t i = 0; i < nset; ++i) m.add(set[i], 1);
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t r = lo; r < hi; ++r) {
            const uint8_t* s = bases + off[r]; size_t len = (size_t)(off[r + 1] - off[r]);
            uint64_t h = 0;
            if (len >= k)
                for (size_t i = 0; i + k <= len; ++i) {
                    uint64_t v;
                    if (seq_to_u64(s + i, k, k, &v) && m.find(canonical_u64(v, k), nullptr)) ++h;
                }
            hits[r] = h;
        }
    };
    if (n_threads <= 1) { work(0, n_reads); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back(work, n_reads * t / n_threads, n_reads * (t + 1) / n_threads);
    for (auto& t : th) t.join();
}

// classify.rs:196-201,224-236: input (kmers,counts) already filtered by min frequency;
// matched = |In n R|, depth = sum of input counts over matched.
void orc_classify_ref(const uint64_t* in_keys, const uint64_t* in_counts, uint64_t n_in,
                      const uint64_t* ref, uint64_t n_ref, uint64_t* matched, uint64_t* depth) {
    CountMap m(1 << 16);
    for (uint64_t i = 0; i < n_ref; ++i) m.add(ref[i], 1);
    uint64_t mt = 0, d = 0;
    for (uint64_t i = 0; i < n_in; ++i) if (m.find(in_keys[i], nullptr)) { ++mt; d += in_counts[i]; }
    *matched = mt; *depth = d;
}

// ---- "not reference behaviour": all-cores count.  Reads are split over threads, each
// with a private table; the tables are then merged by key shard in parallel and the
// result sorted.  Reported beside the faithful single-thread number so the GPU ratio is
// not flattered.
void orc_count_batch_mt(unsigned k, const uint8_t* bases, const uint64_t* off, uint64_t n,
                        int n_threads, uint64_t min_count, uint64_t** keys, uint64_t** counts,
                        uint64_t* n_out) {
    if (n_threads < 1) n_threads = 1;
    const int T = n_threads;
    std::vector<CountMap> maps; maps.reserve(T);
    for (int t = 0; t < T; ++t) maps.emplace_back(1 << 16);
    auto shard_of = [T](uint64_t c) { return (int)(((c * 0x9E3779B97F4A7C15ULL) >> 40) % (uint64_t)T); };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                for (uint64_t r = n * t / T; r < n * (t + 1) / T; ++r)
                    process_sequence_chunk(bases + off[r], (size_t)(off[r + 1] - off[r]), k, maps[t]);
            });
        for (auto& t : th) t.join();
    }
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> parts(T);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                CountMap m(1 << 16);
                for (int s = 0; s < T; ++s) {
                    const CountMap& src = maps[s];
                    for (size_t i = 0; i <= src.mask; ++i)
                        if (src.used[i] && shard_of(src.keys[i]) == t) m.add(src.keys[i], src.vals[i]);
                }
                for (size_t i = 0; i <= m.mask; ++i)
                    if (m.used[i] && m.vals[i] >= min_count) parts[t].emplace_back(m.keys[i], m.vals[i]);
                std::sort(parts[t].begin(), parts[t].end());
            });
        for (auto& t : th) t.join();
    }
    std::vector<std::pair<uint64_t, uint64_t>> all;
    for (auto& p : parts) all.insert(all.end(), p.begin(), p.end());
    std::sort(all.begin(), all.end());
    *n_out = all.size();
    *keys = (uint64_t*)malloc(sizeof(uint64_t) * (all.size() + 1));
    *counts = (uint64_t*)malloc(sizeof(uint64_t) * (all.size() + 1));
    for (size_t i = 0; i < all.size(); ++i) { (*keys)[i] = all[i].first; (*counts)[i] = all[i].second; }
}


// ---- checkers for inputs too large for one hash map (also "not reference behaviour" in their ORGANISATION
// only: every window still goes through seq_to_u64 + canonical_u64 above and an exact map) -------------------
//
// orc_count_batch_slice_mt: the count table restricted to canonical k-mers in [lo, hi] (inclusive bounds), reads
// split over threads.  The multi-GPU bench verifies one narrow key slice of the global table with it: every rank
// runs it over its own reads, the partial tables are summed.
void orc_count_batch_slice_mt(unsigned k, const uint8_t* bases, const uint64_t* off, uint64_t n, uint64_t lo,
                              uint64_t hi, int n_threads, uint64_t** keys, uint64_t** counts, uint64_t* n_out) {
    if (n_threads < 1) n_threads = 1;
    const int T = n_threads;
    std::vector<CountMap> maps; maps.reserve(T);
    for (int t = 0; t < T; ++t) maps.emplace_back(1 << 12);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            CountMap& m = maps[t];
            for (uint64_t r = n * t / T; r < n * (t + 1) / T; ++r) {
                const uint8_t* seq = bases + off[r]; const size_t len = (size_t)(off[r + 1] - off[r]);
                if (len < k) continue;
                for (size_t i = 0; i + k <= len; ++i) {
                    uint64_t v;
                    if (!seq_to_u64(seq + i, k, k, &v)) continue;
                    const uint64_t c = canonical_u64(v, k);
                    if (c >= lo && c <= hi) m.add(c, 1);
                }
            }
        });
    for (auto& t : th) t.join();
    CountMap all(1 << 12);
    for (int t = 0; t < T; ++t)
        for (size_t i = 0; i <= maps[t].mask; ++i) if (maps[t].used[i]) all.add(maps[t].keys[i], maps[t].vals[i]);
    std::vector<std::pair<uint64_t, uint64_t>> v; v.reserve(all.n);
    for (size_t i = 0; i <= all.mask; ++i) if (all.used[i]) v.emplace_back(all.keys[i], all.vals[i]);
    std::sort(v.begin(), v.end());
    *n_out = v.size();
    *keys = (uint64_t*)malloc(sizeof(uint64_t) * (v.size() + 1));
    *counts = (uint64_t*)malloc(sizeof(uint64_t) * (v.size() + 1));
    for (size_t i = 0; i < v.size(); ++i) { (*keys)[i] = v[i].first; (*counts)[i] = v[i].second; }
}

// orc_count_batch_ranged_mt: the WHOLE count table of a batch whose distinct k-mers would not fit per-thread maps
// (config 2: 1.16e9 windows, 2.15e8 distinct).  Pass 1 (threads over reads): canonical k-mer of every window,
// filed by its top 8 key bits.  Pass 2 (threads over the 256 key ranges, ascending): exact map per range, sorted,
// appended -- ranges are ordered, so the concatenation is the sorted table of count.rs:106-119.
// Peak memory: 8 bytes per window + the maps of the ranges in flight.
void orc_count_batch_ranged_mt(unsigned k, const uint8_t* bases, const uint64_t* off, uint64_t n, int n_threads,
                               uint64_t min_count, uint64_t** keys, uint64_t** counts, uint64_t* n_out) {
    if (n_threads < 1) n_threads = 1;
    const int T = n_threads, R = 256;
    const unsigned shift = 2 * k >= 8 ? 2 * k - 8 : 0;
    std::vector<std::vector<std::vector<uint64_t>>> filed(T, std::vector<std::vector<uint64_t>>(R));
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                auto& mine = filed[t];
                for (uint64_t r = n * t / T; r < n * (t + 1) / T; ++r) {
                    const uint8_t* seq = bases + off[r]; const size_t len = (size_t)(off[r + 1] - off[r]);
                    if (len < k) continue;
                    for (size_t i = 0; i + k <= len; ++i) {
                        uint64_t v;
                        if (!seq_to_u64(seq + i, k, k, &v)) continue;
                        const uint64_t c = canonical_u64(v, k);
                        mine[(size_t)((c >> shift) & 255u)].push_back(c);
                    }
                }
            });
        for (auto& t : th) t.join();
    }
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> parts(R);
    {
        std::atomic<int> next{0};
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&] {
                for (;;) {
                    const int r = next.fetch_add(1);
                    if (r >= R) break;
                    size_t tot = 0;
                    for (int s = 0; s < T; ++s) tot += filed[s][r].size();
                    size_t cap = 1 << 12; while (cap < tot / 2 + 16) cap <<= 1;
                    CountMap m(cap);
                    for (int s = 0; s < T; ++s) {
                        for (uint64_t c : filed[s][r]) m.add(c, 1);
                        std::vector<uint64_t>().swap(filed[s][r]);
                    }
                    auto& out = parts[r]; out.reserve(m.n);
                    for (size_t i = 0; i <= m.mask; ++i)
                        if (m.used[i] && m.vals[i] >= min_count) out.emplace_back(m.keys[i], m.vals[i]);
                    std::sort(out.begin(), out.end());
                }
            });
        for (auto& t : th) t.join();
    }
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    *n_out = total;
    *keys = (uint64_t*)malloc(sizeof(uint64_t) * (total + 1));
    *counts = (uint64_t*)malloc(sizeof(uint64_t) * (total + 1));
    size_t o = 0;
    for (auto& p : parts) {
        for (auto& e : p) { (*keys)[o] = e.first; (*counts)[o] = e.second; ++o; }
        std::vector<std::pair<uint64_t, uint64_t>>().swap(p);
    }
}

}  // extern "C"
