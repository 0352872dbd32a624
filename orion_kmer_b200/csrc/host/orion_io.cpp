// orion_io.cpp -- file codecs and the .db format: the host-side rows f1/f2 of SURVEY.md section 8.
// Part of liborion_host.so (g++, no CUDA).
//
//   okh_read_file     mode 0: utils.rs:125-152 get_decompressed_input_reader (codec chosen by EXTENSION:
//                     gz / xz / zst / zstd / plain) -- the .db files;
//                     mode 1: the way build.rs:38 and classify.rs:143 read: raw bytes handed to needletail's
//                     parse_fastx_reader, which sniffs gzip / bzip2 / xz magic itself (needletail 0.5.1 is
//                     locked with bzip2, flate2 and xz2, Cargo.lock:584-591, and without zstd);
//                     mode 2: count.rs:59-63 and query.rs:45-51: the extension codec FIRST, and what comes
//                     out goes through the same parse_fastx_reader sniff (so reads.fastq.bz2, or a gzip
//                     file without the .gz extension, are decoded there as well)
//   okh_write_file    utils.rs:167-199 get_output_writer (gz level 6, xz preset 6, zstd level 3)
//   okh_db_*          db_types.rs:8-14 KmerDbV2 { k: u8, references: HashMap<String, HashSet<u64>> } as
//                     bincode 1.3.3 default options write it (build.rs:141, utils.rs:44): little-endian,
//                     fixed-width integers, u64 lengths:
//                       k:u8 | n_refs:u64 | n_refs x { name_len:u64 | utf-8 | n_kmers:u64 | n_kmers x u64 }
//                     Entry and key order are arbitrary in the reference (hash iteration order); this
//                     writer emits references in insertion order and keys ascending.
//
// zlib is linked.  liblzma / libzstd / libbz2 ship in this image as runtime libraries only (no headers), so
// they are loaded with dlopen and the few entry points used are declared here; a missing library
// turns into the error "xz/zstd support unavailable".
#include <dlfcn.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define OKH_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

thread_local std::string g_io_err;
int io_fail(const std::string& m) { g_io_err = m; return 1; }

std::string ext_of(const char* path) {          // utils.rs:115-119: last extension, lower-cased
    std::string p(path);
    const size_t slash = p.find_last_of('/');
    const size_t dot = p.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot + 1 == p.size()) return "";
    if (dot == (slash == std::string::npos ? 0 : slash + 1)) return "";   // ".hidden" has no extension
    std::string e = p.substr(dot + 1);
    for (auto& c : e) c = (char)tolower((unsigned char)c);
    return e;
}

bool slurp(const char* path, std::vector<uint8_t>& out, const char* what) {
    FILE* f = fopen(path, "rb");
    if (!f) { io_fail(std::string("Failed to open ") + what + ": \"" + path + "\""); return false; }
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
    const bool ok = !ferror(f);
    fclose(f);
    if (!ok) io_fail(std::string("I/O error reading \"") + path + "\"");
    return ok;
}

// ---- gzip (multi-member, like flate2's MultiGzDecoder) -------------------------------------
bool gunzip(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    z_stream z{};
    if (inflateInit2(&z, 15 + 16) != Z_OK) { io_fail("zlib: inflateInit failed"); return false; }
    z.next_in = const_cast<Bytef*>(in.data());
    z.avail_in = (uInt)0;
    size_t in_pos = 0;
    std::vector<uint8_t> buf(1 << 20);
    for (;;) {
        if (z.avail_in == 0 && in_pos < in.size()) {
            const size_t take = std::min<size_t>(in.size() - in_pos, 1u << 30);
            z.next_in = const_cast<Bytef*>(in.data() + in_pos); z.avail_in = (uInt)take; in_pos += take;
        }
        z.next_out = buf.data(); z.avail_out = (uInt)buf.size();
        const int r = inflate(&z, Z_NO_FLUSH);
        out.insert(out.end(), buf.data(), buf.data() + (buf.size() - z.avail_out));
        if (r == Z_STREAM_END) {
            if (z.avail_in == 0 && in_pos >= in.size()) break;
            if (inflateReset(&z) != Z_OK) { inflateEnd(&z); io_fail("zlib: inflateReset failed"); return false; }   // next member
            continue;
        }
        if (r != Z_OK) { inflateEnd(&z); io_fail("corrupt gzip stream"); return false; }
        if (z.avail_in == 0 && in_pos >= in.size() && z.avail_out != 0) { inflateEnd(&z); io_fail("truncated gzip stream"); return false; }
    }
    inflateEnd(&z);
    return true;
}

bool gzip(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    z_stream z{};
    if (deflateInit2(&z, 6, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { io_fail("zlib: deflateInit failed"); return false; }
    std::vector<uint8_t> buf(1 << 20);
    size_t pos = 0;
    int r = Z_OK;
    do {
        const size_t take = std::min<size_t>(n - pos, 1u << 30);
        z.next_in = const_cast<Bytef*>(in + pos); z.avail_in = (uInt)take; pos += take;
        const int flush = pos >= n ? Z_FINISH : Z_NO_FLUSH;
        do {
            z.next_out = buf.data(); z.avail_out = (uInt)buf.size();
            r = deflate(&z, flush);
            out.insert(out.end(), buf.data(), buf.data() + (buf.size() - z.avail_out));
        } while (z.avail_out == 0);
    } while (pos < n || r != Z_STREAM_END);
    deflateEnd(&z);
    return true;
}

// ---- xz via liblzma.so.5 (declarations of the stable C ABI, lzma/base.h + container.h) ------
struct lzma_stream_abi {
    const uint8_t* next_in; size_t avail_in; uint64_t total_in;
    uint8_t* next_out; size_t avail_out; uint64_t total_out;
    const void* allocator; void* internal;
    void* reserved_ptr1; void* reserved_ptr2; void* reserved_ptr3; void* reserved_ptr4;
    uint64_t reserved_int1; uint64_t reserved_int2; size_t reserved_int3; size_t reserved_int4;
    int reserved_enum1; int reserved_enum2;
};
enum { LZMA_OK_ = 0, LZMA_STREAM_END_ = 1, LZMA_RUN_ = 0, LZMA_FINISH_ = 3, LZMA_CONCATENATED_ = 0x08, LZMA_CHECK_CRC64_ = 4 };
struct LzmaApi {
    int (*stream_decoder)(lzma_stream_abi*, uint64_t, uint32_t) = nullptr;
    int (*easy_encoder)(lzma_stream_abi*, uint32_t, int) = nullptr;
    int (*code)(lzma_stream_abi*, int) = nullptr;
    void (*end)(lzma_stream_abi*) = nullptr;
    bool ok = false;
};
LzmaApi& lzma() {
    static LzmaApi a = [] {
        LzmaApi x;
        void* h = dlopen("liblzma.so.5", RTLD_NOW);
        if (!h) h = dlopen("liblzma.so", RTLD_NOW);
        if (h) {
            x.stream_decoder = (int (*)(lzma_stream_abi*, uint64_t, uint32_t))dlsym(h, "lzma_stream_decoder");
            x.easy_encoder = (int (*)(lzma_stream_abi*, uint32_t, int))dlsym(h, "lzma_easy_encoder");
            x.code = (int (*)(lzma_stream_abi*, int))dlsym(h, "lzma_code");
            x.end = (void (*)(lzma_stream_abi*))dlsym(h, "lzma_end");
            x.ok = x.stream_decoder && x.easy_encoder && x.code && x.end;
        }
        return x;
    }();
    return a;
}
bool lzma_run(lzma_stream_abi& s, const uint8_t* in, size_t n, std::vector<uint8_t>& out, const char* what) {
    std::vector<uint8_t> buf(1 << 20);
    s.next_in = in; s.avail_in = n;
    for (;;) {
        s.next_out = buf.data(); s.avail_out = buf.size();
        const int r = lzma().code(&s, s.avail_in == 0 ? LZMA_FINISH_ : LZMA_RUN_);
        out.insert(out.end(), buf.data(), buf.data() + (buf.size() - s.avail_out));
        if (r == LZMA_STREAM_END_) break;
        if (r != LZMA_OK_) { lzma().end(&s); io_fail(std::string(what) + " (liblzma code " + std::to_string(r) + ")"); return false; }
    }
    lzma().end(&s);
    return true;
}
bool unxz(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    if (!lzma().ok) { io_fail("xz support unavailable: liblzma.so.5 not found"); return false; }
    lzma_stream_abi s{};
    if (lzma().stream_decoder(&s, UINT64_MAX, LZMA_CONCATENATED_) != LZMA_OK_) { io_fail("liblzma: decoder init failed"); return false; }
    return lzma_run(s, in.data(), in.size(), out, "corrupt xz stream");
}
bool xz(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    if (!lzma().ok) { io_fail("xz support unavailable: liblzma.so.5 not found"); return false; }
    lzma_stream_abi s{};
    if (lzma().easy_encoder(&s, 6, LZMA_CHECK_CRC64_) != LZMA_OK_) { io_fail("liblzma: encoder init failed"); return false; }
    return lzma_run(s, in, n, out, "xz encoding failed");
}

// ---- zstd via libzstd.so.1 (simple + streaming API of zstd.h) --------------------------------
struct ZBuf { void* p; size_t size; size_t pos; };
struct ZstdApi {
    void* (*createDStream)() = nullptr;
    size_t (*freeDStream)(void*) = nullptr;
    size_t (*decompressStream)(void*, ZBuf*, ZBuf*) = nullptr;
    size_t (*compressBound)(size_t) = nullptr;
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok = false;
};
ZstdApi& zstd() {
    static ZstdApi a = [] {
        ZstdApi x;
        void* h = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW);
        if (h) {
            x.createDStream = (void* (*)())dlsym(h, "ZSTD_createDStream");
            x.freeDStream = (size_t (*)(void*))dlsym(h, "ZSTD_freeDStream");
            x.decompressStream = (size_t (*)(void*, ZBuf*, ZBuf*))dlsym(h, "ZSTD_decompressStream");
            x.compressBound = (size_t (*)(size_t))dlsym(h, "ZSTD_compressBound");
            x.compress = (size_t (*)(void*, size_t, const void*, size_t, int))dlsym(h, "ZSTD_compress");
            x.isError = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
            x.ok = x.createDStream && x.freeDStream && x.decompressStream && x.compressBound && x.compress && x.isError;
        }
        return x;
    }();
    return a;
}
bool unzstd(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    if (!zstd().ok) { io_fail("zstd support unavailable: libzstd.so.1 not found"); return false; }
    void* ds = zstd().createDStream();
    std::vector<uint8_t> buf(1 << 20);
    ZBuf zi{const_cast<uint8_t*>(in.data()), in.size(), 0};
    size_t last = 0;
    while (zi.pos < zi.size) {
        ZBuf zo{buf.data(), buf.size(), 0};
        last = zstd().decompressStream(ds, &zo, &zi);
        if (zstd().isError(last)) { zstd().freeDStream(ds); io_fail("corrupt zstd stream"); return false; }
        out.insert(out.end(), buf.data(), buf.data() + zo.pos);
    }
    for (int guard = 0; last != 0 && guard < (1 << 20); ++guard) {   // flush what the decoder still holds
        ZBuf zo{buf.data(), buf.size(), 0};
        last = zstd().decompressStream(ds, &zo, &zi);
        if (zstd().isError(last)) { zstd().freeDStream(ds); io_fail("corrupt zstd stream"); return false; }
        out.insert(out.end(), buf.data(), buf.data() + zo.pos);
        if (zo.pos == 0) break;
    }
    zstd().freeDStream(ds);
    if (last != 0) { io_fail("truncated zstd stream"); return false; }
    return true;
}
bool zstd_compress(const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    if (!zstd().ok) { io_fail("zstd support unavailable: libzstd.so.1 not found"); return false; }
    out.resize(zstd().compressBound(n));
    const size_t r = zstd().compress(out.data(), out.size(), in, n, 3);   // the zstd crate's level 0 = the library default (3)
    if (zstd().isError(r)) { io_fail("zstd encoding failed"); return false; }
    out.resize(r);
    return true;
}

// ---- bzip2 via libbz2.so.1.0 (bzlib.h's streaming decoder), input only: needletail sniffs it ----
// Every stream of a multi-stream file (pbzip2 writes those) is decoded.  needletail 0.5.1 hands the reader to the
// bzip2 crate; whether that stops after the first stream is not pinned by any reference test (parity unpinned).
struct bz_stream_abi {
    char* next_in; unsigned avail_in; unsigned total_in_lo32; unsigned total_in_hi32;
    char* next_out; unsigned avail_out; unsigned total_out_lo32; unsigned total_out_hi32;
    void* state;
    void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};
enum { BZ_OK_ = 0, BZ_STREAM_END_ = 4 };
struct Bz2Api {
    int (*init)(bz_stream_abi*, int, int) = nullptr;
    int (*run)(bz_stream_abi*) = nullptr;
    int (*end)(bz_stream_abi*) = nullptr;
    bool ok = false;
};
Bz2Api& bz2() {
    static Bz2Api a = [] {
        Bz2Api x;
        void* h = dlopen("libbz2.so.1.0", RTLD_NOW);
        if (!h) h = dlopen("libbz2.so.1", RTLD_NOW);
        if (!h) h = dlopen("libbz2.so", RTLD_NOW);
        if (h) {
            x.init = (int (*)(bz_stream_abi*, int, int))dlsym(h, "BZ2_bzDecompressInit");
            x.run = (int (*)(bz_stream_abi*))dlsym(h, "BZ2_bzDecompress");
            x.end = (int (*)(bz_stream_abi*))dlsym(h, "BZ2_bzDecompressEnd");
            x.ok = x.init && x.run && x.end;
        }
        return x;
    }();
    return a;
}
bool unbz2(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
    if (!bz2().ok) { io_fail("bzip2 support unavailable: libbz2.so.1.0 not found"); return false; }
    std::vector<uint8_t> buf(1 << 20);
    size_t in_pos = 0;
    while (in_pos < in.size()) {                        // one decoder per stream
        if (in_pos && (in.size() - in_pos < 3 || memcmp(in.data() + in_pos, "BZh", 3))) break;     // trailing bytes that are no stream: ignored, as bzip2 does
        bz_stream_abi s{};
        if (bz2().init(&s, 0, 0) != BZ_OK_) { io_fail("libbz2: decoder init failed"); return false; }
        for (;;) {
            if (s.avail_in == 0 && in_pos < in.size()) {
                const size_t take = std::min<size_t>(in.size() - in_pos, 1u << 30);
                s.next_in = (char*)const_cast<uint8_t*>(in.data() + in_pos); s.avail_in = (unsigned)take; in_pos += take;
            }
            s.next_out = (char*)buf.data(); s.avail_out = (unsigned)buf.size();
            const int r = bz2().run(&s);
            out.insert(out.end(), buf.data(), buf.data() + (buf.size() - s.avail_out));
            if (r == BZ_STREAM_END_) { in_pos -= s.avail_in; break; }        // what follows is the next stream
            if (r != BZ_OK_) { bz2().end(&s); io_fail("corrupt bzip2 stream"); return false; }
            if (s.avail_in == 0 && in_pos >= in.size() && s.avail_out != 0) { bz2().end(&s); io_fail("truncated bzip2 stream"); return false; }
        }
        bz2().end(&s);
    }
    return true;
}

// needletail's parse_fastx_reader on a byte stream: gzip (1f 8b), bzip2 ("BZ") or xz (fd 37 7a 58 5a 00) magic
// selects a decoder, anything else is taken as it is
bool sniff_decode(std::vector<uint8_t>& raw, std::vector<uint8_t>& out) {
    if (raw.size() >= 2 && raw[0] == 0x1f && raw[1] == 0x8b) return gunzip(raw, out);
    if (raw.size() >= 2 && raw[0] == 'B' && raw[1] == 'Z') return unbz2(raw, out);
    if (raw.size() >= 6 && !memcmp(raw.data(), "\xfd" "7zXZ\0", 6)) return unxz(raw, out);
    out.swap(raw);
    return true;
}

struct FileData { std::vector<uint8_t> bytes; };

}  // namespace

OKH_EXPORT const char* okh_io_last_error() { return g_io_err.c_str(); }

// mode 0: codec from the extension (load_kmer_db_v2).  mode 1: plain read, then needletail's magic sniff (build /
// classify inputs).  mode 2: codec from the extension, then the sniff (count / query inputs).
OKH_EXPORT void* okh_read_file(const char* path, int mode) {
    std::vector<uint8_t> raw;
    if (!slurp(path, raw, mode == 1 ? "input file for buffered reading" : "input file")) return nullptr;
    FileData* fd = new FileData();
    bool ok = true;
    if (mode == 1) {
        ok = sniff_decode(raw, fd->bytes);
    } else {
        const std::string e = ext_of(path);
        std::vector<uint8_t> dec;
        if (e == "gz") ok = gunzip(raw, dec);
        else if (e == "xz") ok = unxz(raw, dec);
        else if (e == "zst" || e == "zstd") ok = unzstd(raw, dec);
        else dec.swap(raw);
        if (ok && mode == 2) ok = sniff_decode(dec, fd->bytes);
        else fd->bytes.swap(dec);
    }
    if (!ok) { delete fd; return nullptr; }
    return fd;
}
OKH_EXPORT const uint8_t* okh_file_data(void* h) { return ((FileData*)h)->bytes.data(); }
OKH_EXPORT uint64_t okh_file_size(void* h) { return ((FileData*)h)->bytes.size(); }
OKH_EXPORT void okh_file_free(void* h) { delete (FileData*)h; }

// by_extension == 0: plain File::create (compare.rs:85 writes its JSON that way whatever the extension)
OKH_EXPORT int okh_write_file(const char* path, const uint8_t* data, uint64_t len, int by_extension) {
    std::vector<uint8_t> enc;
    const uint8_t* p = data; size_t n = len;
    if (by_extension) {
        const std::string e = ext_of(path);
        bool ok = true, coded = true;
        if (e == "gz") ok = gzip(data, len, enc);
        else if (e == "xz") ok = xz(data, len, enc);
        else if (e == "zst" || e == "zstd") ok = zstd_compress(data, len, enc);
        else coded = false;
        if (!ok) return 1;
        if (coded) { p = enc.data(); n = enc.size(); }
    }
    FILE* f = fopen(path, "wb");
    if (!f) return io_fail(std::string("Failed to create output file: \"") + path + "\"");
    const bool ok = n == 0 || fwrite(p, 1, n, f) == n;
    if (fclose(f) != 0 || !ok) return io_fail(std::string("I/O error writing \"") + path + "\"");
    return 0;
}

// ---- KmerDbV2 <-> bincode -------------------------------------------------------------------
namespace {
struct Db {
    uint8_t k = 0;
    std::vector<std::string> names;
    std::vector<std::vector<uint64_t>> kmers;     // as stored in the file (the reference writes hash order)
};
inline void put_u64(std::vector<uint8_t>& o, uint64_t v) { for (int i = 0; i < 8; ++i) o.push_back((uint8_t)(v >> (8 * i))); }
}  // namespace

OKH_EXPORT void* okh_db_new(uint8_t k) { Db* d = new Db(); d->k = k; return d; }
// db_types.rs:38-40 add_reference: same name overwrites
OKH_EXPORT void okh_db_add_reference(void* h, const char* name, const uint64_t* kmers, uint64_t n) {
    Db* d = (Db*)h;
    for (size_t i = 0; i < d->names.size(); ++i)
        if (d->names[i] == name) { d->kmers[i].assign(kmers, kmers + n); return; }
    d->names.emplace_back(name);
    d->kmers.emplace_back(kmers, kmers + n);
}
OKH_EXPORT uint8_t okh_db_k(void* h) { return ((Db*)h)->k; }
OKH_EXPORT uint64_t okh_db_n_references(void* h) { return ((Db*)h)->names.size(); }
OKH_EXPORT const char* okh_db_name(void* h, uint64_t i) { return ((Db*)h)->names[i].c_str(); }
OKH_EXPORT uint64_t okh_db_n_kmers(void* h, uint64_t i) { return ((Db*)h)->kmers[i].size(); }
OKH_EXPORT const uint64_t* okh_db_kmers(void* h, uint64_t i) { return ((Db*)h)->kmers[i].data(); }
OKH_EXPORT void okh_db_free(void* h) { delete (Db*)h; }

OKH_EXPORT int okh_db_write(void* h, const char* path) {
    const Db* d = (const Db*)h;
    std::vector<uint8_t> o;
    size_t total = 1 + 8;
    for (size_t i = 0; i < d->names.size(); ++i) total += 16 + d->names[i].size() + 8 * d->kmers[i].size();
    o.reserve(total);
    o.push_back(d->k);
    put_u64(o, d->names.size());
    for (size_t i = 0; i < d->names.size(); ++i) {
        put_u64(o, d->names[i].size());
        o.insert(o.end(), d->names[i].begin(), d->names[i].end());
        put_u64(o, d->kmers[i].size());
        const size_t at = o.size();
        o.resize(at + 8 * d->kmers[i].size());
        memcpy(o.data() + at, d->kmers[i].data(), 8 * d->kmers[i].size());   // little-endian host
    }
    return okh_write_file(path, o.data(), o.size(), 1);
}

OKH_EXPORT void* okh_db_read(const char* path) {
    void* fh = okh_read_file(path, 0);
    if (!fh) return nullptr;
    const uint8_t* b = okh_file_data(fh);
    const uint64_t len = okh_file_size(fh);
    uint64_t p = 0;
    auto fail = [&](const char* why) -> void* {
        io_fail(std::string("Failed to deserialize KmerDbV2 from \"") + path + "\": " + why);
        okh_file_free(fh);
        return nullptr;
    };
    auto get_u64 = [&](uint64_t* v) { if (len - p < 8) return false; memcpy(v, b + p, 8); p += 8; return true; };
    if (len < 1) return fail("unexpected end of file");
    Db* d = new Db();
    d->k = b[p++];
    uint64_t n_refs = 0;
    if (!get_u64(&n_refs)) { delete d; return fail("unexpected end of file"); }
    for (uint64_t r = 0; r < n_refs; ++r) {
        uint64_t nl = 0, nk = 0;
        if (!get_u64(&nl) || len - p < nl) { delete d; return fail("unexpected end of file"); }
        std::string name((const char*)b + p, nl); p += nl;
        if (!get_u64(&nk) || (len - p) / 8 < nk) { delete d; return fail("unexpected end of file"); }
        std::vector<uint64_t> ks(nk);
        memcpy(ks.data(), b + p, 8 * nk); p += 8 * nk;
        bool replaced = false;   // a HashMap keeps the last of two equal keys
        for (size_t i = 0; i < d->names.size(); ++i) if (d->names[i] == name) { d->kmers[i].swap(ks); replaced = true; break; }
        if (!replaced) { d->names.push_back(std::move(name)); d->kmers.push_back(std::move(ks)); }
    }
    okh_file_free(fh);
    return d;
}
