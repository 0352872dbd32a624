// orion_host.cpp -- host-side pieces that sit ABOVE the C ABI (include/orion_gpu.h): what the
// reference's Rust drivers do on the CPU before and after the device path.  Built with g++
// into liborion_host.so; no CUDA here.
//
//   okh_fastx_to_batch    needletail parse_fastx_reader + the whitespace half of
//                         normalize(false)  (count.rs:63-72, build.rs:42-48, query.rs:51-66)
//                         -> the C-ABI batch layout (bases + offsets) and the record ids
//   okh_format_counts     count.rs:127-135  "KMER\tcount\n"
//   okh_json_f64          serde_json's (ryu) text of an f64: the ratios of the compare / classify reports
//   okh_synth_*           seeded synthetic workloads of SURVEY.md section 8(d) (SplitMix64)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define OKH_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

// bases / ids live in raw buffers sized once for the whole input (a parsed batch is never larger than its text;
// untouched pages of a large malloc cost nothing) and are filled through a cursor: no per-record vector growth and no
// zero-fill of bytes that are about to be overwritten
struct Batch {
    uint8_t* bases = nullptr;
    uint8_t* ids = nullptr;
    std::vector<uint64_t> offsets;
    std::vector<uint64_t> id_offsets;
    size_t nb = 0, ni = 0;               // bytes used in bases / ids
    Batch() : offsets(1, 0), id_offsets(1, 0) {}
    ~Batch() { free(bases); free(ids); }
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;
    bool begin(size_t text_len) {
        bases = (uint8_t*)malloc(text_len + 1); ids = (uint8_t*)malloc(text_len + 1);
        const size_t guess = text_len / 200 + 16;     // ~ one record per 300 bytes of FASTQ; grows if wrong
        offsets.reserve(guess); id_offsets.reserve(guess);
        return bases && ids;
    }
    void close_record() { offsets.push_back(nb); }
    void add_id(const uint8_t* p, size_t n) {
        if (n && p[n - 1] == '\r') --n;
        memcpy(ids + ni, p, n); ni += n;
        id_offsets.push_back(ni);
    }
};

enum { FX_OK = 0, FX_EMPTY = 1, FX_BAD_START = 2, FX_MALFORMED = 3 };

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; }

// one line = [p, e) without its '\n'; returns the start of the next line
inline size_t next_line(const uint8_t* b, size_t len, size_t p, size_t* e) {
    const uint8_t* nl = (const uint8_t*)memchr(b + p, '\n', len - p);
    *e = nl ? (size_t)(nl - b) : len;
    return nl ? *e + 1 : len;
}

// true if [p, p + n) holds a byte <= 0x20 (every whitespace byte normalize(false) removes is one); the loop has no
// early exit so that the compiler vectorises it
inline bool may_hold_ws(const uint8_t* p, size_t n) {
    unsigned any = 0;
    for (size_t i = 0; i < n; ++i) any |= (unsigned)(p[i] <= 0x20);
    return any != 0;
}

void append_sequence(Batch& out, const uint8_t* p, size_t n, bool strip_ws) {
    uint8_t* d = out.bases + out.nb;
    if (!strip_ws || !may_hold_ws(p, n)) { memcpy(d, p, n); out.nb += n; return; }    // the usual line: nothing to strip
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) { uint8_t c = p[i]; d[m] = c; m += is_ws(c) ? 0 : 1; }
    out.nb += m;
}

int parse_fasta(const uint8_t* b, size_t len, bool strip_ws, Batch& out) {
    size_t p = 0;
    bool open = false;
    size_t raw_begin = 0, raw_end = 0;  // raw mode: sequence region of the open record
    auto close = [&]() {
        if (!open) return;
        if (!strip_ws) {  // raw bytes keep interior line breaks, lose the final end-of-line run
            size_t e = raw_end;
            while (e > raw_begin && (b[e - 1] == '\n' || b[e - 1] == '\r')) --e;
            memcpy(out.bases + out.nb, b + raw_begin, e - raw_begin); out.nb += e - raw_begin;
        }
        out.close_record();
        open = false;
    };
    while (p < len) {
        size_t e, nx = next_line(b, len, p, &e);
        if (b[p] == '>') {
            close();
            out.add_id(b + p + 1, e - p - 1);
            open = true; raw_begin = raw_end = nx;
        } else {
            if (strip_ws) append_sequence(out, b + p, e - p, true);
            raw_end = nx;
        }
        p = nx;
    }
    close();
    return FX_OK;
}

int parse_fastq(const uint8_t* b, size_t len, bool strip_ws, Batch& out) {
    size_t p = 0;
    while (p < len) {
        if (b[p] == '\n' || b[p] == '\r') { ++p; continue; }
        if (b[p] != '@') return FX_MALFORMED;
        size_t he, s0 = next_line(b, len, p, &he);
        if (he >= len) return FX_MALFORMED;
        size_t se, pl = next_line(b, len, s0, &se);
        if (se >= len) return FX_MALFORMED;
        size_t s1 = se; if (s1 > s0 && b[s1 - 1] == '\r') --s1;
        if (pl >= len || b[pl] != '+') return FX_MALFORMED;
        size_t pe, q0 = next_line(b, len, pl, &pe);
        if (pe >= len) return FX_MALFORMED;
        size_t qe, nx = next_line(b, len, q0, &qe);
        size_t q1 = qe; if (q1 > q0 && b[q1 - 1] == '\r') --q1;
        if (q1 - q0 != s1 - s0) return FX_MALFORMED;
        out.add_id(b + p + 1, he - p - 1);
        append_sequence(out, b + s0, s1 - s0, strip_ws);
        out.close_record();
        p = nx;
    }
    return FX_OK;
}

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

inline uint8_t comp(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

}  // namespace

namespace {

// ---- parallel framing -----------------------------------------------------------------------------
// A large text is cut at record starts and the pieces are parsed by several threads into their own batches,
// which are then concatenated (offsets shifted).  Record starts can be recognised from any byte offset:
//   FASTA: a line that begins with '>';
//   FASTQ (strictly 4 lines per record, as parse_fastq requires): a line that begins with '@' whose second
//          successor begins with '+'.  A quality line may begin with '@' as well, but then the line two further
//          on is a sequence line, which never begins with '+'.
inline size_t line_start_after(const uint8_t* b, size_t len, size_t p) {
    if (p == 0) return 0;
    const uint8_t* nl = (const uint8_t*)memchr(b + p - 1, '\n', len - (p - 1));
    return nl ? (size_t)(nl - b) + 1 : len;
}

size_t next_record_start(const uint8_t* b, size_t len, size_t from, bool fastq) {
    size_t p = line_start_after(b, len, from);
    while (p < len) {
        size_t e, l1 = next_line(b, len, p, &e);
        if (!fastq) { if (b[p] == '>') return p; }
        else if (b[p] == '@' && l1 < len) {
            size_t e1, l2 = next_line(b, len, l1, &e1);
            if (l2 < len && b[l2] == '+') return p;
        }
        p = l1;
    }
    return len;
}

int parse_one(const uint8_t* b, size_t len, bool fastq, bool strip_ws, Batch& out) {
    if (!out.begin(len)) return FX_MALFORMED;            // out of host memory: reported as unparseable
    return fastq ? parse_fastq(b, len, strip_ws, out) : parse_fasta(b, len, strip_ws, out);
}

int parse_parallel(const uint8_t* b, size_t len, bool fastq, bool strip_ws, unsigned parts, Batch& out) {
    std::vector<size_t> cut(1, 0);
    for (unsigned i = 1; i < parts; ++i) {
        const size_t c = next_record_start(b, len, len / parts * i, fastq);
        if (c > cut.back() && c < len) cut.push_back(c);
    }
    cut.push_back(len);
    const size_t n = cut.size() - 1;
    if (n < 2) return parse_one(b, len, fastq, strip_ws, out);
    std::vector<Batch> piece(n);
    std::vector<int> st(n, FX_OK);
    {
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; ++i)
            th.emplace_back([&, i] { st[i] = parse_one(b + cut[i], cut[i + 1] - cut[i], fastq, strip_ws, piece[i]); });
        for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < n; ++i) if (st[i] != FX_OK) return st[i];
    std::vector<size_t> b0(n + 1, 0), i0(n + 1, 0), r0(n + 1, 0);
    for (size_t i = 0; i < n; ++i) {
        b0[i + 1] = b0[i] + piece[i].nb; i0[i + 1] = i0[i] + piece[i].ni; r0[i + 1] = r0[i] + piece[i].offsets.size() - 1;
    }
    out.bases = (uint8_t*)malloc(b0[n] + 1); out.ids = (uint8_t*)malloc(i0[n] + 1);
    if (!out.bases || !out.ids) return FX_MALFORMED;
    out.nb = b0[n]; out.ni = i0[n];
    out.offsets.resize(r0[n] + 1); out.id_offsets.resize(r0[n] + 1);
    out.offsets[0] = 0; out.id_offsets[0] = 0;
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            memcpy(out.bases + b0[i], piece[i].bases, piece[i].nb);
            memcpy(out.ids + i0[i], piece[i].ids, piece[i].ni);
            for (size_t r = 1; r < piece[i].offsets.size(); ++r) {
                out.offsets[r0[i] + r] = piece[i].offsets[r] + b0[i];
                out.id_offsets[r0[i] + r] = piece[i].id_offsets[r] + i0[i];
            }
        });
    for (auto& t : th) t.join();
    return FX_OK;
}

}  // namespace

// threads < 0: as many as the host offers (at most 16, at least 8 MB of text per thread); 1: the sequential parser
OKH_EXPORT void* okh_fastx_parse_mt(const uint8_t* buf, uint64_t len, int strip_ws, int threads, int* status) {
    Batch* out = new Batch();
    int st;
    if (len == 0) { out->begin(0); st = FX_EMPTY; }
    else if (buf[0] != '>' && buf[0] != '@') { out->begin(0); st = FX_BAD_START; }
    else {
        unsigned parts = threads < 0 ? std::thread::hardware_concurrency() : (unsigned)threads;
        if (threads < 0) parts = (unsigned)std::min<uint64_t>(std::min<unsigned>(parts, 16u), len / (8u << 20));
        if (threads < 0 && parts < 4) parts = 1;        // the merge copy makes 2-3 pieces slower than the sequential parser (measured)
        if (parts < 1) parts = 1;
        st = parts > 1 ? parse_parallel(buf, (size_t)len, buf[0] == '@', strip_ws != 0, parts, *out)
                       : parse_one(buf, (size_t)len, buf[0] == '@', strip_ws != 0, *out);
    }
    *status = st;
    return out;
}

// ---- FASTA/FASTQ framing ------------------------------------------------------------------
// strip_ws != 0: count/build/classify semantics (whitespace removed, as normalize(false) does;
// the device handles case, U and invalid bytes).  strip_ws == 0: query semantics (raw bytes).
OKH_EXPORT void* okh_fastx_parse(const uint8_t* buf, uint64_t len, int strip_ws, int* status) {
    return okh_fastx_parse_mt(buf, len, strip_ws, -1, status);
}
OKH_EXPORT uint64_t okh_batch_n_records(void* h) { return ((Batch*)h)->offsets.size() - 1; }
OKH_EXPORT uint64_t okh_batch_n_bases(void* h) { return ((Batch*)h)->nb; }
OKH_EXPORT const uint8_t* okh_batch_bases(void* h) { return ((Batch*)h)->bases; }
OKH_EXPORT const uint64_t* okh_batch_offsets(void* h) { return ((Batch*)h)->offsets.data(); }
OKH_EXPORT const uint8_t* okh_batch_ids(void* h) { return ((Batch*)h)->ids; }
OKH_EXPORT const uint64_t* okh_batch_id_offsets(void* h) { return ((Batch*)h)->id_offsets.data(); }
OKH_EXPORT void okh_batch_free(void* h) { delete (Batch*)h; }

// ---- count.rs:127-135 ------------------------------------------------------------------------
// "KMER\tcount\n" per entry.  The k-mer is decoded four bases at a time through a 256-entry table, the lines of a
// large table are formatted by several threads (a first pass sizes every chunk, so each thread writes at its final
// offset): 215 M lines of config 2 are 7.5 GB of text, a minute of single-threaded byte-at-a-time formatting.
namespace {
struct Quad { char c[256][4]; Quad() { for (int v = 0; v < 256; ++v) for (int j = 0; j < 4; ++j) c[v][j] = "ACGT"[(v >> (2 * (3 - j))) & 3]; } };
const Quad g_quad;

inline unsigned dec_digits(uint64_t c) { unsigned d = 1; while (c >= 10) { c /= 10; ++d; } return d; }

inline char* format_line(uint64_t kmer, uint64_t count, unsigned k, char* p) {
    const unsigned head = k & 3u;                       // the first k mod 4 bases one by one, then whole bytes of 2-bit codes
    for (unsigned j = 0; j < head; ++j) *p++ = "ACGT"[(kmer >> (2 * (k - 1 - j))) & 3u];
    for (int sh = (int)(2 * (k - head)) - 8; sh >= 0; sh -= 8) { memcpy(p, g_quad.c[(kmer >> sh) & 0xFFu], 4); p += 4; }
    *p++ = '\t';
    const unsigned d = dec_digits(count);
    for (unsigned i = d; i-- > 0;) { p[i] = (char)('0' + count % 10); count /= 10; }
    p += d;
    *p++ = '\n';
    return p;
}
}  // namespace

// exact size of the text okh_format_counts will write (so that the caller need not allocate k + 22 bytes per line)
OKH_EXPORT uint64_t okh_format_counts_size(const uint64_t* counts, uint64_t n, unsigned k) {
    uint64_t bytes = 0;
    for (uint64_t i = 0; i < n; ++i) bytes += k + 2 + dec_digits(counts[i]);
    return bytes;
}

OKH_EXPORT uint64_t okh_format_counts(const uint64_t* kmers, const uint64_t* counts, uint64_t n, unsigned k,
                                      char* out) {
    unsigned nt = std::thread::hardware_concurrency();
    nt = n < (1u << 18) ? 1u : (nt < 1 ? 1u : (nt > 32 ? 32u : nt));
    if (nt == 1) {
        char* p = out;
        for (uint64_t i = 0; i < n; ++i) p = format_line(kmers[i], counts[i], k, p);
        return (uint64_t)(p - out);
    }
    std::vector<uint64_t> start(nt + 1, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)                   // pass 1: bytes of every chunk
        th.emplace_back([&, t] {
            uint64_t bytes = 0;
            for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; ++i) bytes += k + 2 + dec_digits(counts[i]);
            start[t + 1] = bytes;
        });
    for (auto& x : th) x.join();
    th.clear();
    for (unsigned t = 0; t < nt; ++t) start[t + 1] += start[t];
    for (unsigned t = 0; t < nt; ++t)                   // pass 2: every chunk at its final offset
        th.emplace_back([&, t] {
            char* p = out + start[t];
            for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; ++i) p = format_line(kmers[i], counts[i], k, p);
        });
    for (auto& x : th) x.join();
    return start[nt];
}


// ---- serde_json's f64 text (compare.rs:16-25, classify.rs:22-52 reports) ----------------------------------------
// serde_json prints a finite f64 through the ryu crate: the SHORTEST decimal digits that round-trip, laid out by
// ryu's pretty printer -- digits D (length L), value = D x 10^k, kk = L + k:
//   k >= 0 and kk <= 16   ->  D followed by k zeros and ".0"           100.0   1000000000000000.0
//   0 < kk <= 16          ->  D with a point after kk digits            12.34   0.5 is the next case
//   -5 < kk <= 0          ->  "0." then -kk zeros then D                0.5     0.00001234
//   otherwise             ->  d[.ddd]e<kk-1>                            1e16    1.234e-7   1e-6
// Non-finite values print as null.  The digits come from the smallest precision of "%.*e" that reads back to the
// same double (the correctly rounded shortest form; ryu picks the same digits outside rare tie cases).
// out: at least 32 bytes.  Returns the length written (no terminator counted).
OKH_EXPORT int okh_json_f64(double v, char* out) {
    if (!std::isfinite(v)) { memcpy(out, "null", 5); return 4; }
    char b[48];
    int prec = 1;
    for (; prec <= 17; ++prec) { snprintf(b, sizeof b, "%.*e", prec - 1, v); if (strtod(b, nullptr) == v) break; }
    // b = [-]d[.ddd]e[+-]XX
    const char* p = b;
    std::string s;
    if (*p == '-') { s += '-'; ++p; }
    std::string digits;
    for (; *p && *p != 'e'; ++p) if (*p != '.') digits += *p;
    const int x = atoi(p + 1);                                    // decimal exponent of the first digit
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    if (digits == "0") { s += "0.0"; memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
    const int L = (int)digits.size(), kk = x + 1, k = kk - L;
    if (k >= 0 && kk <= 16) { s += digits; s.append((size_t)k, '0'); s += ".0"; }
    else if (kk > 0 && kk <= 16) { s += digits.substr(0, (size_t)kk); s += '.'; s += digits.substr((size_t)kk); }
    else if (kk > -5 && kk <= 0) { s += "0."; s.append((size_t)(-kk), '0'); s += digits; }
    else {
        s += digits[0];
        if (L > 1) { s += '.'; s += digits.substr(1); }
        s += 'e'; s += std::to_string(kk - 1);
    }
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// ---- synthetic workloads: SplitMix64, 2-bit fields low bits first -> ACGT ----
OKH_EXPORT void okh_synth_genome(uint64_t seed, uint64_t n, uint8_t* out) {
    SplitMix64 g(seed);
    uint64_t i = 0;
    while (i < n) {
        uint64_t r = g.next();
        for (int j = 0; j < 32 && i < n; ++j, ++i) out[i] = (uint8_t)"ACGT"[(r >> (2 * j)) & 3u];
    }
}

// reads: start uniform, strand 50/50, substitution sub_ppm per million bases (uniform over the
// other three), N n_ppm per million.  Each read has its own generator (seed, index), so the
// output does not depend on the thread count.  out holds n_reads*read_len bytes, no separators.
OKH_EXPORT void okh_synth_reads(const uint8_t* genome, uint64_t glen, uint64_t seed, uint64_t first_read,
                                uint64_t n_reads, uint32_t read_len, uint32_t sub_ppm, uint32_t n_ppm,
                                uint8_t* out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t r = lo; r < hi; ++r) {
            SplitMix64 g(seed ^ ((first_read + r + 1) * 0xD1B54A32D192ED03ull));
            const uint64_t start = g.next() % (glen - read_len + 1);
            const bool rev = g.next() & 1u;
            uint8_t* o = out + r * (uint64_t)read_len;
            for (uint32_t i = 0; i < read_len; ++i)
                o[i] = rev ? comp(genome[start + read_len - 1 - i]) : genome[start + i];
            for (uint32_t i = 0; i < read_len; ++i) {
                const uint64_t x = g.next();
                const uint32_t u = (uint32_t)(x % 1000000u);
                if (u < n_ppm) o[i] = 'N';
                else if (u < n_ppm + sub_ppm) {
                    const char* alt = "ACGT";
                    uint8_t c = o[i]; int pick = (int)((x >> 32) % 3u), seen = 0;
                    for (int a = 0; a < 4; ++a) if ((uint8_t)alt[a] != c) { if (seen == pick) { o[i] = (uint8_t)alt[a]; break; } ++seen; }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, n_reads * t / n_threads, n_reads * (t + 1) / n_threads);
    for (auto& t : th) t.join();
}

// descendant of a genome: every base substituted with probability sub_ppm per million
OKH_EXPORT void okh_synth_mutate(const uint8_t* genome, uint64_t n, uint64_t seed, uint32_t sub_ppm, uint8_t* out) {
    SplitMix64 g(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t x = g.next();
        uint8_t c = genome[i];
        if ((uint32_t)(x % 1000000u) < sub_ppm) {
            const char* alt = "ACGT"; int pick = (int)((x >> 32) % 3u), seen = 0;
            for (int a = 0; a < 4; ++a) if ((uint8_t)alt[a] != c) { if (seen == pick) { c = (uint8_t)alt[a]; break; } ++seen; }
        }
        out[i] = c;
    }
}
