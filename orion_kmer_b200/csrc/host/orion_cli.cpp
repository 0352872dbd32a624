// orion_cli.cpp -- `orion-kmer-b200`: the reference's command line (cli.rs:4-185) over liborion_gpu.so.
// Same subcommands, flags, defaults, output formats and error texts as the Rust CLI; the hot loop of
// every driver is a batch call into the C ABI (include/orion_gpu.h), everything around it is host code:
//
//   count     count.rs:40-141      -> ok_counter_*            TSV "KMER\tcount"
//   build     build.rs:80-160      -> ok_set_*                bincode KmerDbV2 (.db)
//   compare   compare.rs:29-97     -> ok_set_union / ok_set_intersection_size   pretty JSON
//   query     query.rs:24-134      -> ok_probe_reads          matching read ids, input order
//   classify  classify.rs:56-385   -> ok_counter_* + ok_probe_counts            pretty JSON (+ TSV)
//
// There is no CPU path: without a CUDA device every command fails with the library's message.
#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "orion_gpu.h"

extern "C" {
// liborion_host.so (orion_host.cpp, orion_io.cpp)
void* okh_fastx_parse(const uint8_t* buf, uint64_t len, int strip_ws, int* status);
uint64_t okh_batch_n_records(void* h);
uint64_t okh_batch_n_bases(void* h);
const uint8_t* okh_batch_bases(void* h);
const uint64_t* okh_batch_offsets(void* h);
const uint8_t* okh_batch_ids(void* h);
const uint64_t* okh_batch_id_offsets(void* h);
void okh_batch_free(void* h);
uint64_t okh_format_counts(const uint64_t* kmers, const uint64_t* counts, uint64_t n, unsigned k, char* out);
uint64_t okh_format_counts_size(const uint64_t* counts, uint64_t n, unsigned k);
int okh_json_f64(double v, char* out);
const char* okh_io_last_error();
void* okh_read_file(const char* path, int mode);
const uint8_t* okh_file_data(void* h);
uint64_t okh_file_size(void* h);
void okh_file_free(void* h);
int okh_write_file(const char* path, const uint8_t* data, uint64_t len, int by_extension);
void* okh_db_new(uint8_t k);
void okh_db_add_reference(void* h, const char* name, const uint64_t* kmers, uint64_t n);
uint8_t okh_db_k(void* h);
uint64_t okh_db_n_references(void* h);
const char* okh_db_name(void* h, uint64_t i);
uint64_t okh_db_n_kmers(void* h, uint64_t i);
const uint64_t* okh_db_kmers(void* h, uint64_t i);
void okh_db_free(void* h);
int okh_db_write(void* h, const char* path);
void* okh_db_read(const char* path);
}

namespace {

int g_verbose = 0;
struct Fail { std::string msg; };                       // anyhow::Error: main prints its outermost message
[[noreturn]] void fail(const std::string& m) { throw Fail{m}; }
void info(const std::string& m) { if (g_verbose >= 1) fprintf(stderr, "[INFO  orion_kmer] %s\n", m.c_str()); }
void gpu(int rc) { if (rc != OK_SUCCESS) fail(ok_last_error()); }
std::string quoted(const std::string& p) { return "\"" + p + "\""; }   // {:?} of a path

std::string invalid_k(unsigned k) { return "Invalid K-mer size: " + std::to_string(k) + ". Must be between 1 and 32."; }

// ---- clap-like argument handling ------------------------------------------------------------
struct Opt { char shrt; const char* lng; bool takes_value; bool multi; };
struct Parsed { std::vector<std::pair<std::string, std::vector<std::string>>> kv; };

const std::vector<std::string>* find(const Parsed& p, const char* lng) {
    for (auto& e : p.kv) if (e.first == lng) return &e.second;
    return nullptr;
}
[[noreturn]] void usage_error(const std::string& m) {
    fprintf(stderr, "error: %s\n\nFor more information, try '--help'.\n", m.c_str());
    exit(2);
}
uint64_t to_uint(const std::string& s, const char* what, uint64_t max) {
    if (s.empty() || s.find_first_not_of("0123456789") != std::string::npos) usage_error("invalid value '" + s + "' for '" + what + "'");
    errno = 0;
    const unsigned long long v = strtoull(s.c_str(), nullptr, 10);
    if (errno || v > max) usage_error("invalid value '" + s + "' for '" + what + "': number too large to fit in target type");
    return v;
}

Parsed parse_args(const std::vector<std::string>& args, const std::vector<Opt>& opts, int* threads) {
    Parsed out;
    auto add = [&](const Opt& o, const std::string& v) {
        for (auto& e : out.kv) if (e.first == o.lng) {
            if (!o.multi && o.takes_value) usage_error(std::string("the argument '--") + o.lng + "' cannot be used multiple times");
            e.second.push_back(v); return;
        }
        out.kv.push_back({o.lng, {v}});
    };
    const Opt* open_multi = nullptr;
    for (size_t i = 0; i < args.size(); ++i) {
        const std::string& a = args[i];
        const Opt* o = nullptr;
        std::string inline_val; bool has_inline = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (auto& c : opts) if (name == c.lng) o = &c;
            if (!o) usage_error("unexpected argument '" + a + "' found");
            if (eq != std::string::npos) { inline_val = a.substr(eq + 1); has_inline = true; }
        } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
            if (a.find_first_not_of('v', 1) == std::string::npos) { g_verbose += (int)a.size() - 1; open_multi = nullptr; continue; }
            for (auto& c : opts) if (c.shrt && a[1] == c.shrt) o = &c;
            if (!o) usage_error("unexpected argument '" + a + "' found");
            if (a.size() > 2) { inline_val = a.substr(a[2] == '=' ? 3 : 2); has_inline = true; }
        } else {
            if (open_multi) { add(*open_multi, a); continue; }
            usage_error("unexpected argument '" + a + "' found");
        }
        open_multi = nullptr;
        if (!o->takes_value) { add(*o, "1"); continue; }
        std::string v;
        if (has_inline) v = inline_val;
        else { if (i + 1 >= args.size()) usage_error(std::string("a value is required for '--") + o->lng + "' but none was supplied"); v = args[++i]; }
        if (!strcmp(o->lng, "threads")) { *threads = (int)to_uint(v, "--threads", 1 << 20); continue; }
        if (!strcmp(o->lng, "verbose")) { ++g_verbose; continue; }
        add(*o, v);
        if (o->multi) open_multi = o;
    }
    return out;
}
std::string need(const Parsed& p, const char* lng) {
    auto v = find(p, lng);
    if (!v) usage_error(std::string("the following required arguments were not provided:\n  --") + lng);
    return (*v)[0];
}
std::vector<std::string> need_all(const Parsed& p, const char* lng) {
    auto v = find(p, lng);
    if (!v) usage_error(std::string("the following required arguments were not provided:\n  --") + lng);
    return *v;
}

// ---- files -----------------------------------------------------------------------------------
struct Batch {
    void* h = nullptr;
    ~Batch() { if (h) okh_batch_free(h); }
    uint64_t n_records() const { return okh_batch_n_records(h); }
};

// count.rs:56-66 / query.rs:45-52 (codec by extension, then needletail's sniff of what comes out) and
// build.rs:38-43 / classify.rs:143-151 (raw bytes, needletail sniffs): read, decode, frame.  strip_ws: normalize(false) removes whitespace;
// query keeps record.sequence() as it is.
void load_fastx(const std::string& path, bool by_magic, bool strip_ws, const std::string& open_ctx,
                const std::string& parse_ctx, Batch& out) {
    void* f = okh_read_file(path.c_str(), by_magic ? 1 : 2);
    if (!f) {
        // File::open fails inside get_decompressed_input_reader / get_buffered_file_reader (the "open" context); a corrupt
        // compressed stream only shows when parse_fastx_reader reads its first bytes (the "parse" context)
        const std::string why = okh_io_last_error();
        fail(why.rfind("Failed to open", 0) == 0 ? open_ctx : parse_ctx);
    }
    int st = 0;
    out.h = okh_fastx_parse(okh_file_data(f), okh_file_size(f), strip_ws ? 1 : 0, &st);
    okh_file_free(f);
    if (st == 1 || st == 2) fail(parse_ctx);                                   // empty file / neither '>' nor '@'
    if (st != 0) fail("Error reading record from " + path);
}

struct Db {
    void* h = nullptr;
    ~Db() { if (h) okh_db_free(h); }
    uint8_t k() const { return okh_db_k(h); }
    uint64_t n_refs() const { return okh_db_n_references(h); }
};
void load_db(const std::string& path, Db& db) {     // utils.rs:37-55
    info("Loading k-mer database (KmerDbV2) from: " + quoted(path));
    db.h = okh_db_read(path.c_str());
    if (!db.h) {
        const std::string why = okh_io_last_error();
        fail(why.rfind("Failed to deserialize", 0) == 0 ? "Failed to deserialize KmerDbV2 from " + quoted(path)
                                                        : "Failed to get input reader for k-mer database: " + quoted(path));
    }
}

struct Set {
    ok_set* s = nullptr;
    Set() = default;
    Set(const Set&) = delete;
    Set(Set&& o) noexcept : s(o.s) { o.s = nullptr; }
    ~Set() { if (s) ok_set_destroy(s); }
    uint64_t size() const { uint64_t n = 0; gpu(ok_set_size(s, &n)); return n; }
};
// one reference of a loaded database as a device set (the file holds the keys in hash order)
Set set_of_reference(const Db& db, uint64_t i) {
    std::vector<uint64_t> v(okh_db_kmers(db.h, i), okh_db_kmers(db.h, i) + okh_db_n_kmers(db.h, i));
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    Set s;
    gpu(ok_set_from_sorted(db.k(), v.data(), v.size(), &s.s));
    return s;
}
// db_types.rs:43-48 get_all_kmers_unified
Set union_of(const Db& db, std::vector<Set>& refs) {
    Set u;
    if (refs.empty()) { gpu(ok_set_from_sorted(db.k(), nullptr, 0, &u.s)); return u; }
    std::vector<ok_set*> hs;
    for (auto& r : refs) hs.push_back(r.s);
    gpu(ok_set_union(hs.data(), hs.size(), &u.s));
    return u;
}

// ---- serde_json pretty printing ------------------------------------------------------------------
std::string json_str(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break; case '\t': o += "\\t"; break; case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
            default: if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; } else o += (char)c;
        }
    }
    return o + "\"";
}
std::string json_f64(double v) {      // serde_json's text of an f64 (ryu's shortest digits and layout): liborion_host.so
    char b[48];
    const int n = okh_json_f64(v, b);
    return std::string(b, (size_t)n);
}
struct Json {                           // objects and arrays only need what the two reports use
    std::string out; int depth = 0; std::vector<bool> first;
    void indent() { out += '\n'; out.append((size_t)depth * 2, ' '); }
    void sep() { if (!first.back()) out += ','; first.back() = false; indent(); }
    void begin(char c) { out += c; ++depth; first.push_back(true); }
    void end(char c) { const bool empty = first.back(); first.pop_back(); --depth; if (!empty) indent(); out += c; }
    void key(const char* k) { sep(); out += json_str(k) + ": "; }
    void kv(const char* k, const std::string& s) { key(k); out += json_str(s); }
    void kv(const char* k, uint64_t v) { key(k); out += std::to_string(v); }
    void kvf(const char* k, double v) { key(k); out += json_f64(v); }
};

void write_out(const std::string& path, const std::string& data, bool by_extension, const std::string& ctx) {
    if (okh_write_file(path.c_str(), (const uint8_t*)data.data(), data.size(), by_extension ? 1 : 0)) fail(ctx);
}

// ---- count (count.rs:40-141) ----------------------------------------------------------------------
void run_count(const Parsed& p) {
    const unsigned k = (unsigned)to_uint(need(p, "kmer-size"), "--kmer-size <KMER_SIZE>", 255);
    const auto inputs = need_all(p, "input-files");
    const std::string out_path = need(p, "output-file");
    uint64_t min_count = 1;
    if (auto v = find(p, "min-count")) min_count = to_uint((*v)[0], "--min-count <MIN_COUNT>", UINT64_MAX);
    if (k == 0 || k > 32) fail(invalid_k(k));
    ok_counter* c = nullptr;
    gpu(ok_counter_create((uint8_t)k, OK_NORM_NORMALIZED, 0, &c));
    for (auto& path : inputs) {
        info("Processing file: " + path);
        Batch b;
        load_fastx(path, false, true, "Failed to get input reader for file: " + path, "Failed to parse FASTA/Q content from: " + path, b);
        gpu(ok_counter_add_batch(c, okh_batch_bases(b.h), okh_batch_offsets(b.h), b.n_records()));
    }
    uint64_t *keys = nullptr, *counts = nullptr, n = 0;
    gpu(ok_counter_finish(c, min_count, &keys, &counts, &n));
    // exact-size, uninitialised text buffer: config 2's table is 215 M lines = 7.5 GB of text
    const uint64_t text_bytes = okh_format_counts_size(counts, n, k);
    std::unique_ptr<char[]> text(new char[text_bytes + 1]);
    okh_format_counts(keys, counts, n, k, text.get());
    ok_free(keys); ok_free(counts);
    ok_counter_destroy(c);
    if (okh_write_file(out_path.c_str(), (const uint8_t*)text.get(), text_bytes, 1))
        fail("Failed to get output writer for file: " + quoted(out_path));
    info("Successfully wrote k-mer counts to " + quoted(out_path));
}

// ---- build (build.rs:80-160) ----------------------------------------------------------------------
std::string basename_of(const std::string& p) {     // Path::file_name, whole path if there is none
    std::string t = p;
    while (t.size() > 1 && t.back() == '/') t.pop_back();
    const size_t s = t.find_last_of('/');
    const std::string b = s == std::string::npos ? t : t.substr(s + 1);
    return (b.empty() || b == "..") ? p : b;
}
void run_build(const Parsed& p) {
    const unsigned k = (unsigned)to_uint(need(p, "kmer-size"), "--kmer-size <KMER_SIZE>", 255);
    const auto genomes = need_all(p, "genomes");
    const std::string out_path = need(p, "output-file");
    if (k == 0 || k > 32) fail(invalid_k(k));
    Db db; db.h = okh_db_new((uint8_t)k);
    for (auto& path : genomes) {
        Batch b;
        // anyhow prints the OUTERMOST context (main.rs:10-13 logs "{}"): build.rs:39,43 with the lossy path, unquoted
        load_fastx(path, true, true, "Failed to get buffered file reader for file: " + path,
                   "Failed to parse FASTA/Q content from: " + path, b);
        Set s;
        gpu(ok_set_create((uint8_t)k, OK_NORM_NORMALIZED, 0, &s.s));
        gpu(ok_set_add_batch(s.s, okh_batch_bases(b.h), okh_batch_offsets(b.h), b.n_records()));
        uint64_t* keys = nullptr; uint64_t n = 0;
        gpu(ok_set_export(s.s, &keys, &n));
        const std::string name = basename_of(path);
        info("Adding " + std::to_string(n) + " unique k-mers from reference '" + name + "' to the database.");
        okh_db_add_reference(db.h, name.c_str(), keys, n);
        ok_free(keys);
    }
    if (okh_db_write(db.h, out_path.c_str())) fail("Failed to get output writer for database file: " + quoted(out_path));
    info("Successfully wrote k-mer database (KmerDbV2) to " + quoted(out_path));
}

// ---- compare (compare.rs:29-97) ---------------------------------------------------------------------
void run_compare(const Parsed& p) {
    const std::string p1 = need(p, "db1"), p2 = need(p, "db2"), out_path = need(p, "output-file");
    Db d1, d2;
    load_db(p1, d1); load_db(p2, d2);
    if (d1.k() != d2.k())      // errors.rs:24-25
        fail("K-mer databases have incompatible k-mer sizes (overall comparison): " + std::to_string(d1.k()) + " vs " + std::to_string(d2.k()));
    if (d1.k() == 0 || d1.k() > 32) fail(invalid_k(d1.k()));
    std::vector<Set> r1, r2;
    for (uint64_t i = 0; i < d1.n_refs(); ++i) r1.push_back(set_of_reference(d1, i));
    for (uint64_t i = 0; i < d2.n_refs(); ++i) r2.push_back(set_of_reference(d2, i));
    Set a = union_of(d1, r1), b = union_of(d2, r2);
    const uint64_t na = a.size(), nb = b.size();
    uint64_t inter = 0;
    gpu(ok_set_intersection_size(a.s, b.s, &inter));
    const uint64_t uni = na + nb - inter;
    Json j; j.begin('{');
    j.kv("db1_path", p1); j.kv("db2_path", p2); j.kv("kmer_size", (uint64_t)d1.k());
    j.kv("db1_total_unique_kmers_across_references", na); j.kv("db2_total_unique_kmers_across_references", nb);
    j.kv("intersection_size", inter); j.kv("union_size", uni);
    j.kvf("jaccard_index", uni == 0 ? 0.0 : (double)inter / (double)uni);
    j.end('}');
    write_out(out_path, j.out, false, "Failed to create output JSON file: " + quoted(out_path));   // plain File::create (compare.rs:85)
}

// ---- query (query.rs:24-134) ------------------------------------------------------------------------
void run_query(const Parsed& p) {
    const std::string db_path = need(p, "database"), reads_path = need(p, "reads"), out_path = need(p, "output-file");
    uint64_t min_hits = 1;
    if (auto v = find(p, "min-hits")) min_hits = to_uint((*v)[0], "--min-hits <MIN_HITS>", UINT64_MAX);
    Db db; load_db(db_path, db);
    if (db.k() == 0 || db.k() > 32) fail(invalid_k(db.k()));
    std::vector<Set> refs;
    for (uint64_t i = 0; i < db.n_refs(); ++i) refs.push_back(set_of_reference(db, i));
    Set all = union_of(db, refs);
    Batch b;
    load_fastx(reads_path, false, false, "Failed to get input reader for reads file: " + quoted(reads_path),
               "Failed to parse FASTQ content from: " + quoted(reads_path), b);
    const uint64_t n = b.n_records();
    std::vector<uint32_t> hits(n);
    gpu(ok_probe_reads(all.s, OK_NORM_RAW, okh_batch_bases(b.h), okh_batch_offsets(b.h), n, hits.data()));
    const uint8_t* ids = okh_batch_ids(b.h);
    const uint64_t* ido = okh_batch_id_offsets(b.h);
    const uint64_t* off = okh_batch_offsets(b.h);
    std::string text;
    for (uint64_t r = 0; r < n; ++r) {
        if (off[r + 1] - off[r] < db.k()) continue;               // query.rs:83-85: shorter than k -> never reported
        if ((uint64_t)hits[r] >= min_hits) { text.append((const char*)ids + ido[r], ido[r + 1] - ido[r]); text += '\n'; }
    }
    write_out(out_path, text, true, "Failed to get output writer for matching reads: " + quoted(out_path));
}

// ---- classify (classify.rs:56-385) --------------------------------------------------------------------
std::string fixed4(double v) { char b[64]; snprintf(b, sizeof b, "%.4f", v); return b; }
std::string tsv_field(const std::string& s) {    // csv crate: quote when the field holds the delimiter, a quote or a line break
    if (s.find_first_of("\t\"\r\n") == std::string::npos) return s;
    std::string o = "\"";
    for (char c : s) { if (c == '"') o += '"'; o += c; }
    return o + "\"";
}
void run_classify(const Parsed& p) {
    const std::string in_path = need(p, "input-file"), out_path = need(p, "output-file");
    const auto db_paths = need_all(p, "databases");
    uint64_t min_freq = 1; double min_cov = 0.0; int user_k = -1;
    if (auto v = find(p, "kmer-size")) user_k = (int)to_uint((*v)[0], "--kmer-size <KMER_SIZE>", 255);
    if (auto v = find(p, "min-kmer-frequency")) min_freq = to_uint((*v)[0], "--min-kmer-frequency <MIN_KMER_FREQUENCY>", UINT64_MAX);
    if (auto v = find(p, "min-coverage")) {
        char* e = nullptr; min_cov = strtod((*v)[0].c_str(), &e);
        if (!e || *e || (*v)[0].empty()) usage_error("invalid value '" + (*v)[0] + "' for '--min-coverage <MIN_COVERAGE>': invalid float literal");
    }
    fprintf(stderr, "DEBUG: Entered run_classify. Input file: %s, Num DBs: %zu, Output: %s\n", quoted(in_path).c_str(),
            db_paths.size(), quoted(out_path).c_str());                     // classify.rs:57-62 prints this unconditionally
    int k = -1;
    if (user_k >= 0) { if (user_k == 0 || user_k > 32) fail(invalid_k((unsigned)user_k)); k = user_k; }
    std::vector<Db> dbs(db_paths.size());
    for (size_t i = 0; i < db_paths.size(); ++i) {
        try { load_db(db_paths[i], dbs[i]); } catch (const Fail&) { fail("Failed to load database: " + quoted(db_paths[i])); }
        const int dk = dbs[i].k();
        if (k >= 0) {
            if (dk != k)
                fail(user_k >= 0 ? "User-provided k-mer size " + std::to_string(k) + " does not match k-mer size " + std::to_string(dk) +
                                       " from database: " + quoted(db_paths[i])
                                 : "Effective k-mer size " + std::to_string(k) + " (from first database) does not match k-mer size " +
                                       std::to_string(dk) + " from database: " + quoted(db_paths[i]));
        } else {
            if (dk == 0 || dk > 32) fail(invalid_k((unsigned)dk));
            k = dk;
        }
    }
    // input k-mer counts, filtered by min_kmer_frequency (classify.rs:135-201)
    Batch b;
    // classify.rs:143-155: outermost contexts, path formatted with {:?}
    load_fastx(in_path, true, true, "Failed to get buffered file reader for file: " + quoted(in_path),
               "Failed to parse FASTA/Q content from: " + quoted(in_path), b);
    ok_counter* c = nullptr;
    gpu(ok_counter_create((uint8_t)k, OK_NORM_NORMALIZED, 0, &c));
    gpu(ok_counter_add_batch(c, okh_batch_bases(b.h), okh_batch_offsets(b.h), b.n_records()));
    uint64_t *keys = nullptr, *counts = nullptr, n_in = 0;
    gpu(ok_counter_finish(c, min_freq, &keys, &counts, &n_in));
    auto ratio = [](uint64_t a, uint64_t d) { return d ? (double)a / (double)d : 0.0; };

    Json j; j.begin('{');
    j.kv("input_file_path", in_path); j.kv("total_unique_kmers_in_input", n_in); j.kv("min_kmer_frequency_filter", min_freq);
    j.key("databases_analyzed"); j.begin('[');
    std::string tsv = "InputFile\tDatabase\tReference\tTotalKmersInReference\tInputKmersHittingReference\tSumDepthMatchedKmers\t"
                      "AvgDepthMatchedKmers\tProportionInputKmersHittingReference\tReferenceBreadthOfCoverage\n";
    for (size_t d = 0; d < dbs.size(); ++d) {
        std::vector<Set> refs;
        for (uint64_t i = 0; i < dbs[d].n_refs(); ++i) refs.push_back(set_of_reference(dbs[d], i));
        Set all = union_of(dbs[d], refs);
        uint64_t om = 0, od = 0;
        gpu(ok_probe_counts(all.s, keys, counts, n_in, &om, &od));      // matched in ANY reference (classify.rs:272-277)
        const uint64_t db_total = all.size();
        j.sep(); j.begin('{');
        j.kv("database_path", db_paths[d]); j.kv("database_kmer_size", (uint64_t)dbs[d].k());
        j.kv("total_unique_kmers_in_db_across_references", db_total);
        j.kv("overall_input_kmers_matched_in_db", om); j.kv("overall_sum_depth_of_matched_kmers_in_input", od);
        j.kvf("overall_avg_depth_of_matched_kmers_in_input", ratio(od, om));
        j.kvf("proportion_input_kmers_in_db_overall", ratio(om, n_in));
        j.kvf("proportion_db_kmers_covered_overall", ratio(om, db_total));
        j.key("references"); j.begin('[');
        // one upload of the input count map, every reference probed on the device (ok_probe_counts_many)
        std::vector<ok_set*> ref_handles;
        for (uint64_t i = 0; i < dbs[d].n_refs(); ++i) ref_handles.push_back(refs[i].s);
        std::vector<uint64_t> ref_m(ref_handles.size(), 0), ref_dep(ref_handles.size(), 0);
        gpu(ok_probe_counts_many(ref_handles.data(), ref_handles.size(), keys, counts, n_in, ref_m.data(), ref_dep.data()));
        for (uint64_t i = 0; i < dbs[d].n_refs(); ++i) {
            const uint64_t m = ref_m[i], dep = ref_dep[i];
            const uint64_t rn = refs[i].size();
            const double breadth = ratio(m, rn);
            if (!(breadth >= min_cov)) continue;                        // classify.rs:247
            const std::string name = okh_db_name(dbs[d].h, i);
            j.sep(); j.begin('{');
            j.kv("reference_name", name); j.kv("total_kmers_in_reference", rn); j.kv("input_kmers_hitting_reference", m);
            j.kv("sum_depth_of_matched_kmers_in_input", dep); j.kvf("avg_depth_of_matched_kmers_in_input", ratio(dep, m));
            j.kvf("proportion_input_kmers_hitting_reference", ratio(m, n_in)); j.kvf("reference_breadth_of_coverage", breadth);
            j.end('}');
            tsv += tsv_field(in_path) + "\t" + tsv_field(db_paths[d]) + "\t" + tsv_field(name) + "\t" + std::to_string(rn) + "\t" +
                   std::to_string(m) + "\t" + std::to_string(dep) + "\t" + fixed4(ratio(dep, m)) + "\t" + fixed4(ratio(m, n_in)) + "\t" +
                   fixed4(breadth) + "\n";
        }
        j.end(']'); j.end('}');
    }
    j.end(']'); j.end('}');
    ok_free(keys); ok_free(counts); ok_counter_destroy(c);
    write_out(out_path, j.out, true, "Failed to get output writer for JSON file: " + quoted(out_path));
    if (auto v = find(p, "output-tsv")) write_out((*v)[0], tsv, true, "Failed to get output writer for TSV file: " + quoted((*v)[0]));
}

const char* HELP =
    "Usage: orion-kmer-b200 [OPTIONS] <COMMAND>\n\n"
    "Commands:\n"
    "  count     Count k-mers in FASTA/FASTQ files\n"
    "  build     Build a unique k-mer database from genome assemblies\n"
    "  compare   Compare two k-mer databases\n"
    "  query     Query short reads against a k-mer database\n"
    "  classify  Classify sequences against k-mer databases and report coverage statistics\n\n"
    "Options:\n"
    "  -t, --threads <THREADS>  Number of threads to use (0 for all logical cores) [default: 0]\n"
    "  -v, --verbose...         Verbosity level (e.g., -v, -vv)\n"
    "  -h, --help               Print help\n"
    "  -V, --version            Print version\n\n"
    "count:    -k <KMER_SIZE> -i <INPUT_FILES>... -o <OUTPUT_FILE> [-m <MIN_COUNT>]\n"
    "build:    -k <KMER_SIZE> -g <GENOME_FILES>... -o <OUTPUT_FILE>\n"
    "compare:  --db1 <DB1> --db2 <DB2> -o <OUTPUT_FILE>\n"
    "query:    -d <DATABASE_FILE> -r <READS_FILE> -o <OUTPUT_FILE> [-c <MIN_HITS>]\n"
    "classify: -i <INPUT_FILE> -d <DATABASE_FILES>... -o <OUTPUT_FILE> [-k <KMER_SIZE>] [--min-kmer-frequency <N>]\n"
    "          [--min-coverage <F>] [--output-tsv <FILE>]\n"
    "Inputs and outputs choose their codec by extension (.gz, .xz, .zst); every k-mer loop runs on the GPU.\n";

}  // namespace

int main(int argc, char** argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    const std::vector<Opt> global = {{'t', "threads", true, false}, {'v', "verbose", false, false}};
    const std::vector<Opt> count_o = {{'k', "kmer-size", true, false}, {'i', "input-files", true, true}, {'o', "output-file", true, false}, {'m', "min-count", true, false}};
    const std::vector<Opt> build_o = {{'k', "kmer-size", true, false}, {'g', "genomes", true, true}, {'o', "output-file", true, false}};
    const std::vector<Opt> compare_o = {{0, "db1", true, false}, {0, "db2", true, false}, {'o', "output-file", true, false}};
    const std::vector<Opt> query_o = {{'d', "database", true, false}, {'r', "reads", true, false}, {'o', "output-file", true, false}, {'c', "min-hits", true, false}};
    const std::vector<Opt> classify_o = {{'i', "input-file", true, false}, {'d', "databases", true, true}, {'o', "output-file", true, false},
                                         {'k', "kmer-size", true, false}, {0, "min-kmer-frequency", true, false}, {0, "min-coverage", true, false},
                                         {0, "output-tsv", true, false}};
    // global flags may come before or after the subcommand (clap `global = true`)
    size_t cmd_at = args.size();
    for (size_t i = 0; i < args.size(); ++i) {
        const std::string& a = args[i];
        if (a == "-h" || a == "--help") { fputs(HELP, stdout); return 0; }
        if (a == "-V" || a == "--version") { printf("orion-kmer-b200 %s\n", ok_version()); return 0; }
        if (a == "-t" || a == "--threads") { ++i; continue; }
        if (!a.empty() && a[0] != '-') { cmd_at = i; break; }
    }
    if (cmd_at == args.size()) { fputs(HELP, stderr); return 2; }
    const std::string cmd = args[cmd_at];
    std::vector<std::string> rest(args.begin(), args.begin() + cmd_at);
    rest.insert(rest.end(), args.begin() + cmd_at + 1, args.end());
    const std::vector<Opt>* spec = cmd == "count" ? &count_o : cmd == "build" ? &build_o : cmd == "compare" ? &compare_o
                                 : cmd == "query" ? &query_o : cmd == "classify" ? &classify_o : nullptr;
    if (!spec) usage_error("unrecognized subcommand '" + cmd + "'");
    std::vector<Opt> opts = *spec;
    opts.insert(opts.end(), global.begin(), global.end());
    int threads = 0;       // accepted for compatibility: the reference only ever uses it for `query` (query.rs:78)
    const Parsed p = parse_args(rest, opts, &threads);
    try {
        if (cmd == "count") run_count(p);
        else if (cmd == "build") run_build(p);
        else if (cmd == "compare") run_compare(p);
        else if (cmd == "query") run_query(p);
        else run_classify(p);
    } catch (const Fail& f) {          // main.rs:10-13: error!("Error: {}", e); exit(1)
        fprintf(stderr, "[ERROR orion_kmer] Error: %s\n", f.msg.c_str());
        return 1;
    }
    ok_shutdown();
    return 0;
}
