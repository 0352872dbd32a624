// orion_gpu.cu -- C ABI (include/orion_gpu.h) over the sm_100a kernels in kernels.cuh.
// Host-side runtime of the hot path: device buffers, streams, the copy/compute pipeline,
// table sizing and growth.  No CPU fallback: every compute entry needs a CUDA device.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <time.h>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "orion_gpu.h"
#include "kernels.cuh"
#include "partition.cuh"
#include "merge.cuh"
#include "setops.cuh"

#define OK_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

// ------------------------------------------------------------------ errors and globals --
thread_local std::string g_err;
int set_err(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return set_err(e_ == cudaErrorMemoryAllocation ? OK_ERR_OUT_OF_MEMORY : OK_ERR_CUDA, \
                           "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, \
                           cudaGetErrorString(e_));                                           \
    } while (0)
#define TRY(call) do { int r_ = (call); if (r_ != OK_SUCCESS) return r_; } while (0)

int g_device = -1;
int g_sms = 0;
int g_push_tma = 0;         // ORION_PUSH_TMA=1: the sharded scatter's copy warp pushes with TMA bulk copies (measured slower at N=2: 11.4 vs 9.0 ms)
int g_scatter_p3 = 1;       // ORION_SCATTER_P3: 3 rounds per warp-tile in the level-1 scatter when it has <= 256 bins
int g_count_seed = 2;      // ORION_COUNT_SEED bits: 1 = seed round of one key per thread, 2 = first round unqueued (measured: 0 8.66, 1 7.42, 2 7.21 ms)
std::atomic<uint64_t> g_launches{0};
std::mutex g_mu;
std::vector<ok_counter*> g_spare_builders;   // cleared set builders waiting for the next ok_set_create (guarded by g_mu)

// Streams of the sets: a small shared pool, handed out round-robin and never destroyed.  A stream per set -- 1,000
// of them for the 1,000 genomes of BASELINE.json configs[3] -- made every device-wide synchronisation (cudaFree's
// implicit one, first of all) walk a thousand streams: closing the sets took 0.4 - 8 ms each, the union's large frees
// up to 270 ms.  Every set operation drains its stream before it returns, so sets can share them.
// Key arrays of the sets: carved out of 1 GiB slabs.  A cudaMalloc costs ~0.35 ms whatever its size (measured: 0.37 ms
// for 0.1 GB, 0.77 ms for 10 GB) -- 40 % of the time a 5 Mbp genome's set took to build -- and a cudaFree drains the
// device.  A slab goes back to the device when the last set inside it is destroyed; arrays above a quarter of a slab
// get an allocation of their own.
struct KeySlab { unsigned long long* base; uint64_t cap, used; uint64_t live; };
constexpr uint64_t KEY_SLAB_KEYS = 1ull << 27;          // 1 GiB
KeySlab* g_cur_slab = nullptr;                          // guarded by g_mu
cudaError_t keys_alloc(uint64_t n_keys, unsigned long long** out, KeySlab** owner) {
    *owner = nullptr;
    if (n_keys > KEY_SLAB_KEYS / 4) return cudaMalloc((void**)out, n_keys * 8);
    const uint64_t need = (std::max<uint64_t>(n_keys, 1) + 31u) & ~31ull;       // 256-byte granules: 16-byte loads, TMA sources
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cur_slab && g_cur_slab->live == 0) g_cur_slab->used = 0;              // nobody inside: start over
    if (!g_cur_slab || g_cur_slab->used + need > g_cur_slab->cap) {
        KeySlab* sl = new KeySlab{nullptr, KEY_SLAB_KEYS, 0, 0};
        const cudaError_t e = cudaMalloc((void**)&sl->base, KEY_SLAB_KEYS * 8);
        if (e != cudaSuccess) { delete sl; cudaGetLastError(); return cudaMalloc((void**)out, n_keys * 8); }     // (a device almost full: exact size)
        if (g_cur_slab && g_cur_slab->live == 0) { cudaFree(g_cur_slab->base); delete g_cur_slab; }
        g_cur_slab = sl;
    }
    *out = g_cur_slab->base + g_cur_slab->used;
    g_cur_slab->used += need;
    ++g_cur_slab->live;
    *owner = g_cur_slab;
    return cudaSuccess;
}
void keys_free(unsigned long long* p, KeySlab* owner) {
    if (!owner) { cudaFree(p); return; }
    std::lock_guard<std::mutex> lk(g_mu);
    if (--owner->live == 0 && owner != g_cur_slab) { cudaFree(owner->base); delete owner; }
}

constexpr int N_SET_STREAMS = 8;
cudaStream_t g_set_streams[N_SET_STREAMS] = {};
std::atomic<unsigned> g_set_stream_next{0};
cudaError_t set_stream(cudaStream_t* out) {
    const unsigned i = g_set_stream_next.fetch_add(1) % N_SET_STREAMS;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_set_streams[i]) {
        const cudaError_t e = cudaStreamCreateWithFlags(&g_set_streams[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
    }
    *out = g_set_streams[i];
    return cudaSuccess;
}


#define LAUNCH(kern, grid, block, smem, stream, ...)                 \
    do {                                                             \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);    \
        g_launches.fetch_add(1, std::memory_order_relaxed);          \
    } while (0)

// ORION_TRACE=1: wall-clock laps of the host-side steps of the set operations on stderr (every lap drains the device
// first, so a traced run is slower than a plain one; the laps say where the time goes)
const bool g_trace = getenv("ORION_TRACE") != nullptr;
struct TraceClock {
    double t0 = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    TraceClock() { if (g_trace) { cudaDeviceSynchronize(); t0 = now(); } }
    void lap(const char* fmt, ...) {
        if (!g_trace) return;
        cudaDeviceSynchronize();
        const double t1 = now();
        char buf[256];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        fprintf(stderr, "[orion trace] %-60s %9.3f ms\n", buf, t1 - t0);
        t0 = now();
    }
};

int ensure_init() {
    if (g_device >= 0) return OK_SUCCESS;
    return ok_init(nullptr, 0);
}

int invalid_k(unsigned k) {  // errors.rs:6-7
    return set_err(OK_ERR_INVALID_KMER_SIZE, "Invalid K-mer size: %u. Must be between 1 and 32.", k);
}

inline unsigned grid_for(uint64_t n, unsigned per_block = 256, unsigned waves = 8) {
    uint64_t need = (n + per_block - 1) / per_block;
    uint64_t cap = (uint64_t)(g_sms > 0 ? g_sms : 148) * waves;
    return (unsigned)std::max<uint64_t>(1, std::min(need, cap));
}

// --------------------------------------------------------------- pinned result buffers --
struct PinnedBlock { void* p; size_t bytes; bool used; };
std::vector<PinnedBlock> g_pool;

int pool_alloc(void** out, size_t bytes) {
    if (bytes == 0) bytes = 8;
    std::lock_guard<std::mutex> lk(g_mu);
    PinnedBlock* best = nullptr;
    for (auto& b : g_pool)
        if (!b.used && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && (!best || b.bytes < best->bytes)) best = &b;
    if (best) { best->used = true; *out = best->p; return OK_SUCCESS; }
    void* p = nullptr;
    CU(cudaMallocHost(&p, bytes));
    g_pool.push_back({p, bytes, true});
    *out = p;
    return OK_SUCCESS;
}
bool pool_release(void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& b : g_pool) if (b.p == p) { b.used = false; return true; }
    return false;
}

template <class T>
int dev_reserve(T** p, uint64_t* cap, uint64_t need) {  // grow-only device buffer, contents dropped
    if (*cap >= need && *p) return OK_SUCCESS;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    uint64_t n = std::max<uint64_t>(need, 1);
    CU(cudaMalloc((void**)p, n * sizeof(T)));
    *cap = n;
    return OK_SUCCESS;
}

constexpr double LF_TARGET = 0.5;    // load factor a fresh table is sized for
constexpr double LF_MAX = 0.7;       // occupancy at which the table is rebuilt larger
constexpr unsigned MAX_PROBE = 2048; // displacement bound L (also the tail padding)
constexpr uint64_t MIN_SLOTS = 1ull << 16;
constexpr uint64_t SPILL_CAP = 1ull << 22;
constexpr uint64_t COPY_CHUNK = 128ull << 20;  // bytes per H2D piece of the ingest pipeline (a multiple of the tile size)

}  // namespace

// =============================================================================== counter ==
// RUN_LEVEL1: a host batch has been scattered into level-1 bins and everything after that (level 2, count,
// compaction) is deferred, so that ok_counter_finish can run it slice by slice under the result's D2H copy
enum { RUN_NONE = 0, RUN_SPARSE = 1, RUN_DENSE = 2, RUN_LEVEL1 = 3 };

// plan of one partitioned batch: bin widths and the per-batch device arrays (inside d_meta)
struct PartPlan {
    OkPartCfg cfg{};
    unsigned n_sub = 1, n_bin1 = 1, stride = 1;
    bool sharded = false;                    // keys arrived through the fused multi-GPU scatter (level-1 regions per sender)
    bool hinted = false;                     // sub-partitions sized from the caller's capacity hint (distinct keys), not from the windows
    bool big_count = false;                  // sub-partitions may average > 5800 keys: the 16384-slot count kernel
    uint64_t cap_bound = 0, max_items = 0;
    unsigned *hist = nullptr, *beg = nullptr, *cursor = nullptr, *cap_end = nullptr, *deferred = nullptr;
    unsigned *beg1 = nullptr, *cursor1 = nullptr, *end1 = nullptr;
    unsigned *item_off = nullptr, *item_n = nullptr, *item_bin = nullptr;
    unsigned long long* scan = nullptr;      // exclusive scan of n_distinct, n_sub + 1 entries
    unsigned long long* chunk_sum = nullptr; // per-1024-chunk totals of the two-launch scans
    OkPartScalars* scal = nullptr;
    unsigned* bin_first = nullptr;           // first level-2 work item of every level-1 bin (+ the total)
    unsigned long long* lb = nullptr;        // look-back words of the dense count (one per sub-partition)
    unsigned* ticket = nullptr;              // sub-partition ticket of the dense count kernel
    uint64_t dense_cap = 0;                  // entries d_run_keys / d_run_counts hold for the dense attempt (0: sparse only)
    unsigned n_slices = 0, slice_step = 0;   // result slices: sub-partitions [i*step, (i+1)*step)
};

// fused multi-GPU exchange (ok_shard_*): geometry agreed by all ranks + the peer-mapped level-1 buffers
struct ShardState {
    bool ready = false, buffers = false;
    bool hinted = false;                     // sub-partitions sized from the capacity hint (expected distinct k-mers of THIS rank's shard)
    unsigned g = 0, sub_bits = 0, b1 = 0, b2 = 0, stride = 16;
    uint64_t n_bases_max = 0, cap_keys = 0;
    unsigned long long* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned* d_state = nullptr;             // reg_beg | reg_end | reg_fill | send_cur | send_end (1024 entries each) | OkShardBlocks
    OkShardBlocks* d_blk = nullptr;
    unsigned* d_snap = nullptr;              // cursor snapshots between the chunks of a sharded scatter, 9 x 1024
    OkShardBlocks* h_blk = nullptr;          // page-locked mirror
    unsigned *reg_beg = nullptr, *reg_end = nullptr, *reg_fill = nullptr, *send_cur = nullptr, *send_end = nullptr;
    unsigned long long* d_received = nullptr;
    // chunked exchange over the copy engines (ok_xchg_*)
    unsigned n_chunks = 0;                   // chunks per batch, agreed by all ranks
    // three-level form (>= 4 ranks): the sender splits by OWNER only (long contiguous runs: friendly to DRAM and to the
    // copy engines), the owner runs BOTH scatter levels of the one-GPU path on what arrives (level 1 per chunk, under
    // the exchange).  rb1 / rb2: the owner's level bits; b1 (sender-side level-1 bits) is 0 then.
    bool three = false;
    unsigned rb1 = 0, rb2 = 0;
    double margin = 1.25;                    // what a rank may receive, as a multiple of the largest batch (ok_shard_set_margin)
    uint64_t recv_cap = 0;                   // keys every peer-mapped buffer holds
    unsigned long long* d_send = nullptr; uint64_t cap_send = 0;   // sub-blocks for the other owners, built locally
    unsigned* d_xchg = nullptr;              // cur | end | beg (sender side), rbeg | rend | rfill (receiver side): n_chunks x 1024 each; then hdr_send | hdr_recv: n_chunks x 8
    cudaStream_t s_peer[8][4] = {};          // copy streams: up to 4 per peer (one stream's copies reach ~500 GB/s, NVLink takes more)
    cudaEvent_t ev_piece[8][4] = {};         // end of a (peer, stream) piece of the chunk in flight
    unsigned streams_per_peer = 1;
    cudaStream_t s_join = nullptr;           // joins the copy streams of a chunk: ev_sent[chunk] = "this sender's chunk has landed everywhere"
    cudaEvent_t ev_sent[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // the owner's receive pipeline (fills + level-2 scatter per chunk) on its own stream, under the exchange of later chunks
    cudaStream_t s_recv = nullptr;
    cudaEvent_t ev_recv[3] = {nullptr, nullptr, nullptr};     // start / end of the receive pipeline
    unsigned recv_chunks = 0;                // chunks whose receive work has been issued (ok_xchg_chunk_recv)
    bool begun = false;                      // ok_xchg_scatter_begin without its _end yet
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool xchg_pending = false;               // a batch has been scattered and exchanged, not yet counted
    unsigned xchg_chunks_used = 0;
    float ms_copy_tail = 0;                  // what the peer copies took beyond the last scatter kernel
};

struct PartHost {                            // page-locked mirror of the batch's scalars
    OkPartScalars scal;
    unsigned long long total;
    unsigned long long slice_base[64];       // output offset of every result slice, then the total
    unsigned long long received;             // sharded path: k-mers the peers wrote into this rank's buffer
    unsigned long long slice0_windows;       // sliced result pipeline: windows held by the first slice
    unsigned long long windows_now;          //   ... and by the whole batch
    unsigned dense_failed;                   // the dense count met a sub-partition it could not finish (zero-copy store)
};

struct ok_counter {
    unsigned k = 0;
    int norm_mode = 0;
    uint64_t hint = 0, user_hint = 0;   // user_hint: ok_counter_create's capacity_hint as given (hint follows the table)
    bool distrust_hint = false;         // the hint made shared-memory tables overflow once: sub-partitions are sized from the windows again
    int shard_rank = 0, n_shards = 1;   // multi-GPU: this table holds one key range of n_shards
    cudaStream_t s_main = nullptr, s_copy = nullptr, s_aux = nullptr;   // s_aux: compaction of the sliced result pipeline
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    std::vector<cudaEvent_t> ev_chunks;
    // table
    OkTableView tv{};                 // tv.slots == nullptr until the first batch
    OkDevStats* d_stats = nullptr;
    OkDevStats* h_stats = nullptr;    // pinned mirror
    OkSpill spill{};
    uint64_t occupied = 0, windows = 0, bases_seen = 0, max_disp = 0, spilled_total = 0, grows = 0;
    float ms_insert = 0, ms_readout = 0, ms_fill = 0, ms_route = 0, ms_push = 0;
    // staging for host batches
    uint8_t* d_bases = nullptr; uint64_t cap_bases = 0;
    uint64_t* d_off = nullptr; uint64_t cap_off = 0;
    // readout
    unsigned long long* d_tiles = nullptr; uint64_t cap_tiles = 0;   // [n_tiles] + total
    unsigned long long* d_out_keys = nullptr; uint64_t cap_out_keys = 0;
    unsigned long long* d_out_counts = nullptr; uint64_t cap_out_counts = 0;
    // partitioned (one-shot) path: its result is a sorted run instead of a table --
    // RUN_SPARSE: sorted sub-partition runs in d_buf2 (keys) / d_buf1 (counts); RUN_DENSE: d_run_*
    int path_mode = 0;                 // 0 auto, 1 table only, 2 partitioned whenever the counter is empty
    int run_state = RUN_NONE;
    PartPlan pl;
    PartHost* h_part = nullptr;
    PartHost* h_part_dev = nullptr;    // the same page-locked block as the kernels address it (zero-copy stores)
    uint64_t pend_bases = 0, pend_rec = 0, pend_windows_before = 0;   // RUN_LEVEL1: the batch still sitting in d_bases / d_off
    ShardState shard;
    bool buf1_external = false;        // d_buf1 is the caller's peer-mapped buffer (never reallocated or freed here)
    bool keys_sorted_runs = false;     // ok_set_union: the next key batch is a concatenation of sorted runs (strided level-1 gather)
    unsigned long long* d_run_keys = nullptr; uint64_t cap_run_keys = 0;
    unsigned long long* d_run_counts = nullptr; uint64_t cap_run_counts = 0;
    uint64_t n_run = 0, n_deferred = 0;
    uint64_t batch_windows = 0;         // windows the sharded batch in progress added (taken back by ok_counter_abort_batch)
    // earlier batches' result, set aside while the next large batch is counted on its own; merged afterwards (merge.cuh)
    unsigned long long* d_acc_keys = nullptr; uint64_t cap_acc_keys = 0;
    unsigned long long* d_acc_counts = nullptr; uint64_t cap_acc_counts = 0;
    uint64_t n_acc = 0;
    unsigned long long* d_mrg_keys = nullptr; uint64_t cap_mrg_keys = 0;      // merge output (swapped with the run afterwards)
    unsigned long long* d_mrg_counts = nullptr; uint64_t cap_mrg_counts = 0;
    ulonglong2* d_split = nullptr; uint64_t cap_split = 0;
    float ms_merge = 0; uint64_t n_merges = 0;
    unsigned long long* d_buf1 = nullptr; uint64_t cap_buf1 = 0;
    unsigned long long* d_buf2 = nullptr; uint64_t cap_buf2 = 0;
    unsigned long long* d_cnt = nullptr; uint64_t cap_cnt = 0;       // counts of the runs when d_buf1 is peer-mapped memory (measured: 25 % slower to write)
    unsigned* d_meta = nullptr; uint64_t cap_meta = 0;               // per-batch arrays of the plan (32-bit words)
    float ms_sample = 0, ms_scatter1 = 0, ms_scatter2 = 0, ms_count = 0, ms_compact = 0;
    cudaEvent_t ev_p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

uint64_t keyspace_bound(unsigned k) { return k >= 31 ? ~0ull : (1ull << (2 * k)); }

uint64_t slots_for(const ok_counter* c, uint64_t distinct) {
    uint64_t ks = keyspace_bound(c->k);
    if (distinct > ks) distinct = ks;
    uint64_t n = (uint64_t)((double)distinct / LF_TARGET) + 1;
    return std::max<uint64_t>(n, MIN_SLOTS);
}

int table_alloc(ok_counter* c, uint64_t n_home, OkTableView* out) {
    OkTableView t{};
    t.n_home = n_home;
    t.n_total = n_home + MAX_PROBE;
    t.n_home_all = n_home * (uint64_t)c->n_shards;
    t.home_base = n_home * (uint64_t)c->shard_rank;
    t.key_shift = 64 - 2 * c->k;
    t.map_mode = OK_MAP_CANON;
    t.max_probe = MAX_PROBE;
    cudaError_t e = cudaMalloc((void**)&t.slots, t.n_total * sizeof(OkSlot));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(OK_ERR_OUT_OF_MEMORY, "cannot allocate a %llu-slot k-mer table (%.1f GB): %s",
                       (unsigned long long)t.n_total, t.n_total * 16.0 / 1e9, cudaGetErrorString(e));
    }
    LAUNCH(k_fill_slots, grid_for(t.n_total, 256, 16), 256, 0, c->s_main, t.slots, t.n_total);
    *out = t;
    return OK_SUCCESS;
}

int read_stats(ok_counter* c) {
    CU(cudaMemcpyAsync(c->h_stats, c->d_stats, sizeof(OkDevStats), cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    c->occupied = c->h_stats->occupied;
    c->windows = c->h_stats->windows;
    c->max_disp = std::max<uint64_t>(c->max_disp, c->h_stats->max_disp);
    return OK_SUCCESS;
}

// Rebuild into a table of new_home slots, then re-add whatever the probe bound spilled.
int table_rebuild(ok_counter* c, uint64_t new_home) {
    for (int attempt = 0; attempt < 6; ++attempt) {
        cudaEvent_t e0 = c->ev_a, e1 = c->ev_b;
        CU(cudaEventRecord(e0, c->s_main));
        OkTableView nt{};
        TRY(table_alloc(c, new_home, &nt));
        const uint64_t n_spill = std::min<uint64_t>(c->h_stats->spill_n, c->spill.cap);
        // reset the per-table statistics (windows is cumulative and kept)
        CU(cudaMemsetAsync(&c->d_stats->occupied, 0, sizeof(unsigned long long), c->s_main));
        CU(cudaMemsetAsync(&c->d_stats->max_disp, 0, sizeof(unsigned long long), c->s_main));
        CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, sizeof(unsigned long long), c->s_main));
        // spilled entries move to a scratch copy first (the rebuild may spill again)
        uint64_t *sk = nullptr, *si = nullptr;
        if (n_spill) {
            CU(cudaMalloc((void**)&sk, n_spill * 8)); CU(cudaMalloc((void**)&si, n_spill * 8));
            CU(cudaMemcpyAsync(sk, c->spill.keys, n_spill * 8, cudaMemcpyDeviceToDevice, c->s_main));
            CU(cudaMemcpyAsync(si, c->spill.incs, n_spill * 8, cudaMemcpyDeviceToDevice, c->s_main));
        }
        if (c->tv.slots)
            LAUNCH(k_rehash, grid_for(c->tv.n_total), 256, 0, c->s_main, nt, c->d_stats, c->spill,
                   c->tv.slots, c->tv.n_total);
        if (n_spill)
            LAUNCH(k_add_kmers, grid_for(n_spill), 256, 0, c->s_main, nt, c->d_stats, c->spill,
                   (const unsigned long long*)sk, (const unsigned long long*)si, n_spill, 0);
        CU(cudaEventRecord(e1, c->s_main));
        TRY(read_stats(c));
        CU(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); c->ms_fill += ms;
        if (sk) cudaFree(sk);
        if (si) cudaFree(si);
        if (c->tv.slots) cudaFree(c->tv.slots);
        c->tv = nt;
        c->max_disp = c->h_stats->max_disp;
        ++c->grows;
        if (c->h_stats->spill_n == 0) return OK_SUCCESS;
        if (c->h_stats->spill_n > c->spill.cap)
            return set_err(OK_ERR_INTERNAL, "spill list overflow while rebuilding the k-mer table");
        c->spilled_total += c->h_stats->spill_n;
        new_home *= 2;  // still too clustered for the displacement bound: spread further
    }
    return set_err(OK_ERR_INTERNAL,
                   "k-mer keys are too clustered for the ordered table (displacement bound %u exceeded "
                   "after repeated growth)", MAX_PROBE);
}

// Make room for up to `incoming` new distinct keys; returns how many may be added before the
// next check.
int ensure_headroom(ok_counter* c, uint64_t incoming, uint64_t* allowed) {
    if (!c->tv.slots) {
        uint64_t want = c->hint ? c->hint : incoming;
        TRY(table_rebuild(c, slots_for(c, std::max<uint64_t>(want, 1))));
        c->grows = 0;
    }
    const uint64_t ks = keyspace_bound(c->k);
    for (;;) {
        const uint64_t limit = (uint64_t)(LF_MAX * (double)c->tv.n_home);
        // a table that already covers the whole key space at <= LF_TARGET can never overfill
        if (ks != ~0ull && (double)ks <= LF_TARGET * (double)c->tv.n_home) { *allowed = ~0ull; return OK_SUCCESS; }
        const uint64_t free_keys = limit > c->occupied ? limit - c->occupied : 0;
        const uint64_t min_step = std::min<uint64_t>(incoming, 16ull << 20);
        if (free_keys >= min_step) { *allowed = free_keys; return OK_SUCCESS; }
        uint64_t want = std::max<uint64_t>(2 * c->tv.n_home, slots_for(c, c->occupied + min_step));
        TRY(table_rebuild(c, want));
    }
}

template <class Sink>
void launch_extract(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off,
                    uint64_t n_rec, uint64_t t0, uint64_t t1, cudaStream_t st, const Sink& sink, int norm_mode,
                    unsigned k) {
    (void)c;
    const uint64_t n_tiles = t1 - t0;
    const uint64_t max_warps = (uint64_t)(g_sms > 0 ? g_sms : 148) * 8 /*blocks per SM*/ * 8 /*warps*/;
    const uint64_t tpw = std::max<uint64_t>(1, (n_tiles + max_warps - 1) / max_warps);
    const uint64_t warps = (n_tiles + tpw - 1) / tpw;
    const unsigned blocks = (unsigned)((warps + 7) / 8);
    auto kern = norm_mode == OK_NORM_NORMALIZED ? k_extract<true, Sink> : k_extract<false, Sink>;
    LAUNCH(kern, blocks, 256, 0, st, d_bases, n_bases, d_off, n_rec, t0, t1, tpw, k, sink);
}

// count the windows of tiles [t0, t1) of a device-resident batch
int counter_process_tiles(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off,
                          uint64_t n_rec, uint64_t t0, uint64_t t1) {
    while (t0 < t1) {
        const uint64_t remaining = std::min<uint64_t>((t1 - t0) * OK_TILE_BASES, n_bases - t0 * OK_TILE_BASES);
        uint64_t allowed = 0;
        TRY(ensure_headroom(c, remaining, &allowed));
        uint64_t nt = t1 - t0;
        if (allowed != ~0ull) nt = std::min<uint64_t>(nt, std::max<uint64_t>(1, allowed / OK_TILE_BASES));
        SinkCount sink{}; sink.t = c->tv; sink.st = c->d_stats; sink.sp = c->spill;
        CU(cudaEventRecord(c->ev_a, c->s_main));
        launch_extract(c, d_bases, n_bases, d_off, n_rec, t0, t0 + nt, c->s_main, sink, c->norm_mode, c->k);
        CU(cudaEventRecord(c->ev_b, c->s_main));
        TRY(read_stats(c));
        CU(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->ms_insert += ms;
        if (c->h_stats->spill_n) {
            if (c->h_stats->spill_n > c->spill.cap)
                return set_err(OK_ERR_INTERNAL, "spill list overflow (%llu entries): k-mer keys too clustered",
                               (unsigned long long)c->h_stats->spill_n);
            c->spilled_total += c->h_stats->spill_n;
            TRY(table_rebuild(c, 2 * c->tv.n_home));
        }
        t0 += nt;
    }
    return OK_SUCCESS;
}

int counter_readout(ok_counter* c, uint64_t min_count, uint64_t* n_out) {
    *n_out = 0;
    if (!c->tv.slots || c->occupied == 0) return OK_SUCCESS;
    if (min_count == 0) min_count = 1;
    const uint64_t n_tiles = (c->tv.n_total + OK_RT_SLOTS - 1) / OK_RT_SLOTS;
    TRY(dev_reserve(&c->d_tiles, &c->cap_tiles, n_tiles + 1));
    CU(cudaEventRecord(c->ev_a, c->s_main));
    LAUNCH(k_readout_count, (unsigned)n_tiles, 256, 0, c->s_main, c->tv.slots, c->tv.n_total, min_count, c->d_tiles);
    LAUNCH(k_scan_tiles, 1, 1024, 0, c->s_main, c->d_tiles, n_tiles, c->d_tiles + n_tiles);
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, c->d_tiles + n_tiles, 8, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (total) {
        TRY(dev_reserve(&c->d_out_keys, &c->cap_out_keys, total));
        TRY(dev_reserve(&c->d_out_counts, &c->cap_out_counts, total));
        if (min_count > 1)
            LAUNCH(k_readout_write<true>, (unsigned)n_tiles, 256, 0, c->s_main, c->tv, min_count, c->d_tiles,
                   c->d_out_keys, c->d_out_counts);
        else
            LAUNCH(k_readout_write<false>, (unsigned)n_tiles, 256, 0, c->s_main, c->tv, min_count, c->d_tiles,
                   c->d_out_keys, c->d_out_counts);
    }
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_readout, c->ev_a, c->ev_b);
    *n_out = total;
    return OK_SUCCESS;
}

}  // namespace

// ===================================================================== partitioned path ==
namespace {

constexpr uint64_t PART_MIN_BASES = 1ull << 20;   // below this the table path is as fast
constexpr unsigned PART_TARGET = 4096;            // keys per sub-partition if every key were distinct
constexpr unsigned PART_DISTINCT_TARGET = 4096;   // ... expected distinct keys per sub-partition when the caller gave a capacity hint
constexpr unsigned PART_WINDOWS_MAX = 24576;      // ... and windows per sub-partition at most (16-bit counts: < 61440 even when unbalanced)
constexpr unsigned PART_MAX_BITS = 18;            // up to 9 bits at level 1 + up to 10 at level 2
unsigned RESULT_SLICES = 16;            // the result pipeline compacts and ships the table in slices

// A counter that already holds a sorted RUN stays on the one-shot path: the run is set aside, the batch counted on
// its own and the two runs merged (run_stash / run_unstash).  Only a live device-wide TABLE rules the path out.
// Batches beyond PART_MAX_UNITS are cut into sub-batches by the callers (32-bit offsets inside a batch).
constexpr uint64_t PART_MAX_UNITS = 1792ull << 20;
bool part_eligible(const ok_counter* c, uint64_t n_units) {
    if (c->path_mode == 1) return false;
    if (c->run_state == RUN_NONE && c->occupied) return false;
    return c->path_mode == 2 || n_units >= PART_MIN_BASES;
}

void part_choose_bits(ok_counter* c, uint64_t n_units, PartPlan& pl, bool use_hint = false) {
    uint64_t want = (n_units + PART_TARGET - 1) / PART_TARGET;       // safe when every key is distinct
    uint64_t exp_distinct = n_units;
    if (use_hint && c->user_hint && !c->distrust_hint && !getenv("ORION_NO_HINT_BITS")) {
        // The count kernel's table holds a sub-partition's DISTINCT keys.  With the caller's capacity hint
        // (expected distinct k-mers) sub-partitions can hold several windows per slot: fewer, larger ones mean
        // fewer bins per scatter level and less per-sub-partition overhead.  A hint that is too low costs
        // speed only: what outgrows a table is deferred to the generic kernel.
        exp_distinct = std::min<uint64_t>(c->user_hint, n_units);
        const uint64_t by_distinct = (exp_distinct + PART_DISTINCT_TARGET - 1) / PART_DISTINCT_TARGET;
        const uint64_t by_windows = (n_units + PART_WINDOWS_MAX - 1) / PART_WINDOWS_MAX;
        const uint64_t hinted_want = std::max(by_distinct, by_windows);
        if (hinted_want < want) { want = hinted_want; pl.hinted = true; }
    }
    unsigned bits = 0;
    while ((1ull << bits) < want && bits < PART_MAX_BITS) ++bits;
    if (const char* ev = getenv("ORION_SUB_BITS")) {   // tuning knob: total bits of the two scatter levels
        const unsigned v = (unsigned)atoi(ev);
        if (v >= 1 && v <= PART_MAX_BITS) bits = v;
    }
    pl.cfg.key_shift = 64 - 2 * c->k;
    pl.cfg.shard_log2 = 0;
    for (int g = c->n_shards; g > 1; g >>= 1) ++pl.cfg.shard_log2;
    // one level up to 256 bins; two balanced levels beyond (staging runs stay >= 8 keys per bin)
    pl.cfg.b1 = bits <= 8 ? bits : (bits + 1) / 2;
    if (const char* ev = getenv("ORION_B1")) {   // tuning knob: bits of the first scatter level
        unsigned v = (unsigned)atoi(ev);
        if (v >= 1 && v <= 10 && v <= bits && bits - v <= 10) pl.cfg.b1 = v;
    }
    pl.cfg.b2 = bits - pl.cfg.b1;
    pl.n_sub = 1u << bits;
    pl.n_bin1 = 1u << pl.cfg.b1;
    pl.big_count = exp_distinct / pl.n_sub > (pl.hinted ? 4600u : 5800u);   // units bound the keys: a 6144-key table would defer too often
    if (const char* ev = getenv("ORION_BIG_COUNT")) pl.big_count = atoi(ev) != 0;   // test hook: force the 16384-slot count kernel
}

// kernels specialised on k: 31 and 21 (BASELINE.json's configurations) get compile-time shifts, any other k the generic code
#define OK_BY_K(k, KERN, ...) ((k) == 31 ? KERN<__VA_ARGS__, 31> : (k) == 21 ? KERN<__VA_ARGS__, 21> : KERN<__VA_ARGS__, 0>)
#define OK_BY_K3(k, KERN, ...) ((k) == 31 ? KERN<__VA_ARGS__, 31, true> : (k) == 21 ? KERN<__VA_ARGS__, 21, true> : KERN<__VA_ARGS__, 0, true>)

template <class K> int set_smem(K kern, size_t bytes) {
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return OK_SUCCESS;
}

// Upper bound of the sum of the sub-partition capacities k_part_plan will hand out, so that the
// buffers can be allocated before the sample has been looked at (no host round trip):
//   sum(est_p) <= S (every sampled unit stands for `stride` units),
//   sum(6 sqrt(est_p stride)) <= 6 sqrt(stride) sqrt(n_sub S)   (Cauchy-Schwarz).
uint64_t part_cap_bound(uint64_t n_units, unsigned n_sub, unsigned stride, uint64_t unit_chunk) {
    const double S = (double)n_units + (double)unit_chunk * (stride + 1.0);
    double b = S;
    if (stride > 1) b += 6.0 * std::sqrt((double)stride) * std::sqrt((double)n_sub * (S + (double)n_sub * stride)) + (129.0 + 10.0 * stride) * n_sub;
    return (uint64_t)b + 2ull * n_sub + 64;
}

// carve the per-batch device arrays out of c->d_meta and make sure the key buffers are large enough
int part_layout(ok_counter* c, uint64_t n_units, uint64_t unit_chunk, uint64_t flat_keys, PartPlan& pl) {
    pl.cap_bound = part_cap_bound(n_units, pl.n_sub, pl.stride, unit_chunk);
    if (pl.cap_bound >= (1ull << 32)) return set_err(OK_ERR_INTERNAL, "partitioned path: batch too large for 32-bit offsets");
    pl.max_items = std::max<uint64_t>(pl.cap_bound / OK_PART_TILE + OK_PART_MAXBINS + 1, flat_keys / OK_PART_TILE + 2);   // one partial item per bin / region
    const uint64_t n_sub = pl.n_sub;
    uint64_t words = 0;                                   // 32-bit words
    auto take = [&](uint64_t n) { uint64_t o = words; words += (n + 3) & ~3ull; return o; };
    const uint64_t o_scan = take(2 * (n_sub + 2)), o_chunk = take(2 * (n_sub / 1024 + 2)), o_scal = take(sizeof(OkPartScalars) / 4), o_hist = take(n_sub), o_beg = take(n_sub),
                   o_cur = take(n_sub), o_end = take(n_sub), o_def = take(n_sub), o_b1 = take(OK_PART_MAXBINS),
                   o_c1 = take(OK_PART_MAXBINS), o_e1 = take(OK_PART_MAXBINS), o_io = take(pl.max_items),
                   o_in = take(pl.max_items), o_ib = take(pl.max_items), o_bf = take(OK_PART_MAXBINS + 1),
                   o_lb = take(2 * (n_sub + 2)), o_tk = take(4);
    TRY(dev_reserve(&c->d_meta, &c->cap_meta, words));
    unsigned* m = c->d_meta;
    pl.scan = (unsigned long long*)(m + o_scan); pl.chunk_sum = (unsigned long long*)(m + o_chunk); pl.scal = (OkPartScalars*)(m + o_scal);
    pl.hist = m + o_hist; pl.beg = m + o_beg; pl.cursor = m + o_cur; pl.cap_end = m + o_end; pl.deferred = m + o_def;
    pl.beg1 = m + o_b1; pl.cursor1 = m + o_c1; pl.end1 = m + o_e1;
    pl.item_off = m + o_io; pl.item_n = m + o_in; pl.item_bin = m + o_ib; pl.bin_first = m + o_bf;
    pl.lb = (unsigned long long*)(m + o_lb); pl.ticket = m + o_tk;
    TRY(dev_reserve(&c->d_buf2, &c->cap_buf2, pl.cap_bound + 16));
    if (c->buf1_external) {
        if (c->cap_buf1 < pl.cap_bound + 16) return set_err(OK_ERR_INVALID_ARGUMENT, "the peer buffer is too small for this batch (%llu < %llu keys)",
                                                            (unsigned long long)c->cap_buf1, (unsigned long long)pl.cap_bound + 16);
        TRY(dev_reserve(&c->d_cnt, &c->cap_cnt, pl.cap_bound + 16));
    } else {
        TRY(dev_reserve(&c->d_buf1, &c->cap_buf1, pl.cap_bound + 16));   // level-1 output, later the counts of the runs
    }
    return OK_SUCCESS;
}

int run_to_table(ok_counter* c);
int run_make_dense(ok_counter* c);

constexpr int PART_RETRY = 100;   // internal: the one-shot path gave up, count the batch through the table instead

// work items of the level-2 scatter, from the level-1 fills (one small launch per batch)
void part_launch_items(ok_counter* c, PartPlan& pl) {
    if (pl.cfg.b2 == 0) return;
    if (pl.sharded)   // (bin, sender) regions filled by the peers
        LAUNCH(k_part_items, 32, 1024, 0, c->s_main, c->shard.reg_beg, c->shard.reg_fill, c->shard.reg_end, pl.n_bin1 << c->shard.g,
               pl.n_bin1 - 1u, pl.item_off, pl.item_n, pl.item_bin, pl.scal, (unsigned*)nullptr);
    else
        LAUNCH(k_part_items, 32, 1024, 0, c->s_main, pl.beg1, pl.cursor1, pl.end1, pl.n_bin1, 0xFFFFFFFFu, pl.item_off, pl.item_n, pl.item_bin,
               pl.scal, pl.bin_first);
}

// Level 2 + count + scan of sub-partitions [p0, p1) on the compute stream.  The whole batch is one
// such range; the sliced result pipeline issues one per slice (p0, p1: multiples of 1024 and of
// the sub-partitions per level-1 bin).  Sorted runs land in place (keys in d_buf2, counts in
// d_buf1); pl.scan[p] = output offset of sub-partition p; host_total (page-locked, optional)
// receives the running total of distinct k-mers after this range.
int part_launch_range(ok_counter* c, PartPlan& pl, unsigned p0, unsigned p1, bool whole, unsigned long long* host_total, bool level2 = true,
                      bool dense = false) {
    const OkPartSpill ps{c->spill, c->d_stats};
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    if (pl.cfg.b2 > 0 && level2) {
        auto k_l2 = OK_BY_K(c->k, k_part_scatter_keys, 2, true);
        TRY(set_smem(k_l2, sizeof(OkScatterKeysSmem)));
        if (whole)
            LAUNCH(k_l2, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), c->s_main, c->d_buf1, pl.item_off, pl.item_n,
                   pl.item_bin, pl.scal, pl.cfg, pl.cursor, pl.cap_end, c->d_buf2, ps, (const unsigned*)nullptr, 0u, 0u);
        else
            LAUNCH(k_l2, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), c->s_main, c->d_buf1, pl.item_off, pl.item_n,
                   pl.item_bin, pl.scal, pl.cfg, pl.cursor, pl.cap_end, c->d_buf2, ps, (const unsigned*)pl.bin_first,
                   p0 >> pl.cfg.b2, p1 >> pl.cfg.b2);
    }
    if (whole) CU(cudaEventRecord(c->ev_p[3], c->s_main));
    // count every sub-partition in shared memory; sorted runs land in place, counts in d_buf1
    unsigned* d_nd = pl.hist;   // the sample histogram is no longer needed (k_part_plan zeroed it)
    unsigned* cnt_out = reinterpret_cast<unsigned*>(c->buf1_external ? c->d_cnt : c->d_buf1);   // 32-bit counts, same indices as the keys
    OkDenseOut dn{};
    if (dense) {
        dn.lb = pl.lb; dn.keys = c->d_run_keys; dn.counts = c->d_run_counts; dn.cap = pl.dense_cap; dn.ticket = pl.ticket;
        dn.failed = &c->h_part_dev->dense_failed;
        CU(cudaMemsetAsync(pl.ticket, 0, sizeof(unsigned), c->s_main));
    }
    if (pl.big_count) {
        auto k_cnt = OK_BY_K(c->k, k_part_count, 14);
        TRY(set_smem(k_cnt, sizeof(OkCount2Smem<14>)));
        LAUNCH(k_cnt, std::min<unsigned>(p1 - p0, grid_sm), OkCount2Cfg<14>::THREADS, sizeof(OkCount2Smem<14>), c->s_main,
               c->d_buf2, pl.beg, pl.cursor, pl.cap_end, p0, p1, pl.cfg, cnt_out, d_nd, pl.deferred, pl.scal,
               (g_count_seed & 1) ? OkCount2Cfg<14>::THREADS : 0u, (g_count_seed & 2) != 0, dn);
    } else {
        auto k_cnt = OK_BY_K(c->k, k_part_count, 13);
        TRY(set_smem(k_cnt, sizeof(OkCount2Smem<13>)));
        LAUNCH(k_cnt, std::min<unsigned>(p1 - p0, grid_sm * 2), OkCount2Cfg<13>::THREADS, sizeof(OkCount2Smem<13>), c->s_main,
               c->d_buf2, pl.beg, pl.cursor, pl.cap_end, p0, p1, pl.cfg, cnt_out, d_nd, pl.deferred, pl.scal,
               (g_count_seed & 1) ? OkCount2Cfg<13>::THREADS : 0u, (g_count_seed & 2) != 0, dn);
    }
    if (!dense) {        // (a dense attempt with a deferred sub-partition has failed as a whole: the range is recounted sparse)
        const size_t ct_smem = (size_t)OK_CT_SLOTS * 12;
        TRY(set_smem(k_part_count_generic, ct_smem));
        LAUNCH(k_part_count_generic, grid_sm, OK_CT_THREADS, ct_smem, c->s_main, c->d_buf2, pl.beg, pl.cursor, pl.cap_end,
               pl.deferred, pl.scal, pl.cfg, cnt_out, d_nd, ps);
    }
    if (whole) CU(cudaEventRecord(c->ev_p[4], c->s_main));
    const unsigned chunk0 = p0 / 1024u, n_chunks = (p1 - p0 + 1023u) / 1024u;
    LAUNCH(k_part_scan_sums, n_chunks, 1024, 0, c->s_main, d_nd, pl.n_sub, pl.chunk_sum, chunk0);
    LAUNCH(k_part_scan, n_chunks, 1024, 0, c->s_main, d_nd, pl.n_sub, pl.chunk_sum, pl.scan, chunk0, host_total, pl.scal);
    return OK_SUCCESS;
}

// level 2 + count + scan, shared by the two entry points.  The keys are already scattered into
// level-1 bins in d_buf1 (b2 > 0) or straight into sub-partitions in d_buf2 (b2 == 0).  Leaves
// the result as sorted sub-partition runs (RUN_SPARSE); compaction happens when it is asked for.
// The count kernel writes the dense result itself (OkDenseOut): reserve the outputs and clear the look-back words.
// Capacity: every window could be a distinct k-mer, so without a capacity hint the outputs are sized for the
// windows; with one, for the hint + 25 % -- a result that outgrows them fails the attempt (recounted sparse).
int part_dense_prepare(ok_counter* c, PartPlan& pl) {
    static const bool off = getenv("ORION_NO_DENSE") != nullptr;      // A/B knob: sorted runs in place + compaction pass
    pl.dense_cap = 0;
    if (off) return OK_SUCCESS;
    uint64_t cap = pl.cap_bound;
    if (c->user_hint && !c->distrust_hint) cap = std::min<uint64_t>(cap, c->user_hint + c->user_hint / 4 + (1ull << 20));
    TRY(dev_reserve(&c->d_run_keys, &c->cap_run_keys, cap));
    TRY(dev_reserve(&c->d_run_counts, &c->cap_run_counts, cap));
    CU(cudaMemsetAsync(pl.lb, 0, (size_t)pl.n_sub * 8, c->s_main));
    c->h_part->dense_failed = 0;
    pl.dense_cap = cap;
    return OK_SUCCESS;
}

// a dense attempt failed (a sub-partition was deferred, or the result outgrew the outputs): count [0, n_sub) again in
// the sparse form.  The dense pass left the sub-partitions (d_buf2) untouched.
int part_recount_sparse(ok_counter* c, PartPlan& pl) {
    CU(cudaMemsetAsync(&pl.scal->n_deferred, 0, sizeof(unsigned), c->s_main));
    CU(cudaMemsetAsync(&pl.scal->def_done, 0, sizeof(unsigned), c->s_main));
    pl.dense_cap = 0;
    return part_launch_range(c, pl, 0, pl.n_sub, true, nullptr, /*level2=*/false, /*dense=*/false);
}

int part_finish(ok_counter* c, PartPlan& pl, bool level2 = true) {
    // level2 == false: the caller has already run the level-2 scatter (chunk by chunk: ok_xchg_count_device)
    if (level2) part_launch_items(c, pl);
    TRY(part_dense_prepare(c, pl));
    bool dense = pl.dense_cap != 0;
    TRY(part_launch_range(c, pl, 0, pl.n_sub, true, nullptr, level2, dense));
    if (dense) {
        CU(cudaStreamSynchronize(c->s_main));
        if (c->h_part->dense_failed) { dense = false; TRY(part_recount_sparse(c, pl)); }
    }
    CU(cudaEventRecord(c->ev_p[5], c->s_main));
    // the one host round trip of the batch: totals, slice boundaries of the result, statistics
    const unsigned step = std::max<unsigned>(1, pl.n_sub / RESULT_SLICES);
    pl.n_slices = (pl.n_sub + step - 1) / step;
    CU(cudaMemcpy2DAsync(c->h_part->slice_base, 8, pl.scan, (size_t)step * 8, 8, pl.n_slices, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaMemcpyAsync(&c->h_part->total, pl.scan + pl.n_sub, 8, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaMemcpyAsync(&c->h_part->scal, pl.scal, sizeof(OkPartScalars), cudaMemcpyDeviceToHost, c->s_main));
    TRY(read_stats(c));   // synchronises the stream
    CU(cudaGetLastError());
    c->h_part->slice_base[pl.n_slices] = c->h_part->total;
    pl.slice_step = step;
    cudaEventElapsedTime(&c->ms_sample, c->ev_p[0], c->ev_p[1]);
    cudaEventElapsedTime(&c->ms_scatter1, c->ev_p[1], c->ev_p[2]);
    cudaEventElapsedTime(&c->ms_scatter2, c->ev_p[2], c->ev_p[3]);
    cudaEventElapsedTime(&c->ms_count, c->ev_p[3], c->ev_p[4]);
    float ms_scan = 0; cudaEventElapsedTime(&ms_scan, c->ev_p[4], c->ev_p[5]);
    c->ms_count += ms_scan;
    c->ms_compact = 0;
    c->ms_insert = c->ms_sample + c->ms_scatter1 + c->ms_scatter2 + c->ms_count;
    c->ms_readout = 0;
    c->n_run = c->h_part->total; c->run_state = dense ? RUN_DENSE : RUN_SPARSE;
    c->n_deferred = c->h_part->scal.n_deferred;
    c->occupied = c->n_run;
    return OK_SUCCESS;
}

// whatever the displacement / capacity bounds spilled is exact but unsorted: fold the run and
// the spill list into the general table.  If even the spill list overflowed, nothing of this
// batch is kept and the caller re-counts it through the table path.
// forget everything the one-shot path did with the current batch
int part_discard(ok_counter* c, uint64_t windows_before) {
    c->run_state = RUN_NONE; c->n_run = 0; c->occupied = 0; c->windows = windows_before; c->n_deferred = 0;
    CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, 8, c->s_main));
    CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    c->h_stats->spill_n = 0;
    return OK_SUCCESS;
}

// A capacity hint far below the truth makes the shared-memory tables of the count kernel overflow; the
// deferred sub-partitions would crawl through the generic kernel and its spill list.  Cheaper and
// cleaner: drop the attempt and count the batch again with sub-partitions sized from the windows.
bool part_hint_misled(const ok_counter* c, const PartPlan& pl) {
    return pl.hinted && c->n_deferred > pl.n_sub / 256u + 4u;
}

int part_count_keys_absorb(ok_counter* c, const uint64_t* d_keys, uint64_t n, int depth);
int run_stash(ok_counter* c);
int run_unstash(ok_counter* c);

// What the one-shot path spilled (keys past a sampled capacity or past a sub-partition's table) is exact but
// unsorted.  The spilled keys are counted as a small batch of their own by the same path and that run is MERGED
// into the batch's run (a few rounds at most: every round takes another table-full of distinct keys out of a hot
// sub-partition).  Folding everything into the device-wide ordered table -- round 1's way -- fails exactly when
// spills happen: thousands of distinct k-mers on one 16-base prefix (microsatellites) share one home slot there.
// If even the spill list overflowed, nothing of this batch is kept and the caller re-counts it through the table path.
int part_absorb_spills(ok_counter* c, uint64_t windows_before, int depth = 0) {
    if (c->h_stats->spill_n == 0) return OK_SUCCESS;
    if (c->h_stats->spill_n > c->spill.cap) {
        TRY(part_discard(c, windows_before));
        return PART_RETRY;
    }
    const uint64_t n = c->h_stats->spill_n;
    c->spilled_total += n;
    if (depth >= 6 || c->n_acc) return run_to_table(c);          // (a set-aside run of earlier batches: keep the old way)
    unsigned long long* d_tmp = nullptr;
    CU(cudaMalloc((void**)&d_tmp, n * 8));
    CU(cudaMemcpyAsync(d_tmp, c->spill.keys, n * 8, cudaMemcpyDeviceToDevice, c->s_main));
    CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, 8, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    c->h_stats->spill_n = 0;
    const uint64_t windows = c->windows;                        // the spilled windows are already in it
    const int saved_mode = c->path_mode;
    int r = run_stash(c);
    if (r == OK_SUCCESS) {
        c->path_mode = 2;
        r = part_count_keys_absorb(c, (const uint64_t*)d_tmp, n, depth + 1);     // (plans afresh: c->pl now describes the spill batch)
    }
    c->path_mode = saved_mode;
    cudaFree(d_tmp);
    if (r != OK_SUCCESS) return r == PART_RETRY ? set_err(OK_ERR_INTERNAL, "the spill list overflowed while its own keys were being counted") : r;
    c->windows = windows;
    CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    return run_unstash(c);
}

// pieces of a batch that is still landing in device memory: piece i covers bases
// [i*piece_bytes, (i+1)*piece_bytes) and is complete once ev[i] has fired on the copy stream
struct PieceSchedule { uint64_t n_pieces, piece_bytes; const cudaEvent_t* ev; };

// d_bases: the batch in device memory (possibly still landing, see `pieces`); sample_src: where the
// sampling kernel reads the bases from -- d_bases itself, or the caller's page-locked host buffer
// (zero-copy), so that the plan exists before the first piece has arrived.
// defer: stop after the level-1 scatter when the plan can be finished slice by slice (RUN_LEVEL1);
// part_settle or the sliced ok_counter_finish does the rest.
bool part_sliceable(const PartPlan& pl);
// tile_begin / tile_end: count the windows ending in tiles [tile_begin, tile_end) only -- a sub-batch of a batch too
// large for one pass (the walk takes its k-1 base halo from the tile before, so sub-batches lose no window)
int part_count_bases(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec,
                     const uint8_t* sample_src, const PieceSchedule* pieces, bool defer = false,
                     uint64_t tile_begin = 0, uint64_t tile_end = ~0ull) {
    PartPlan& pl = c->pl; pl = PartPlan{};
    const uint64_t windows_before = c->windows;
    const uint64_t n_tiles = std::min<uint64_t>(tile_end, (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES);      // first tile NOT counted
    const uint64_t n_units = std::min<uint64_t>(n_tiles * OK_TILE_BASES, n_bases) - tile_begin * OK_TILE_BASES;   // window ends in range
    const bool whole = tile_begin == 0 && n_tiles * OK_TILE_BASES >= n_bases;
    part_choose_bits(c, n_units, pl, /*use_hint=*/true);
    // every `stride`-th tile is sampled; large sub-partitions (sized from a capacity hint) still see > 500 samples each at 1/32
    pl.stride = n_tiles - tile_begin > 4096 ? (n_units / pl.n_sub >= 12288 ? 32 : 16) : 1;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    TRY(part_layout(c, n_units, OK_TILE_BASES, 0, pl));
    CU(cudaEventRecord(c->ev_p[0], c->s_main));
    CU(cudaMemsetAsync(pl.hist, 0, pl.n_sub * sizeof(unsigned), c->s_main));
    {
        const uint64_t sampled = (n_tiles - tile_begin + pl.stride - 1) / pl.stride;
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((sampled + 7) / 8, (uint64_t)grid_sm * 8));
        auto kern = c->norm_mode == OK_NORM_NORMALIZED ? k_part_sample<true> : k_part_sample<false>;
        LAUNCH(kern, blocks, 256, 0, c->s_main, sample_src, n_bases, d_off, n_rec, n_tiles, (uint64_t)pl.stride, c->k, pl.cfg, pl.hist,
               /*halo=*/sample_src == d_bases, tile_begin);
        CU(cudaEventRecord(c->ev_b, c->s_main));   // the last reader of sample_src (possibly the caller's own buffer, zero-copy)
    }
    LAUNCH(k_part_plan_sums, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)n_units, pl.chunk_sum);
    LAUNCH(k_part_plan, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)n_units, pl.cfg.b2,
           pl.chunk_sum, (unsigned)pl.cap_bound, pl.beg, pl.cursor, pl.cap_end, pl.beg1, pl.cursor1, pl.end1, pl.scal);
    CU(cudaEventRecord(c->ev_p[1], c->s_main));
    {
        const bool two = pl.cfg.b2 > 0;
        auto kern = c->norm_mode == OK_NORM_NORMALIZED ? OK_BY_K(c->k, k_part_scatter_bases, true, false) : OK_BY_K(c->k, k_part_scatter_bases, false, false);
        const unsigned bins1 = 1u << (two ? pl.cfg.b1 : pl.cfg.b1 + pl.cfg.b2);
        if (bins1 <= 256 && g_scatter_p3)      // 32 staging slots per bin: three fuller rounds per warp-tile instead of four
            kern = c->norm_mode == OK_NORM_NORMALIZED ? OK_BY_K3(c->k, k_part_scatter_bases, true, false) : OK_BY_K3(c->k, k_part_scatter_bases, false, false);
        TRY(set_smem(kern, sizeof(OkScatterSmem)));
        const uint64_t max_warps = (uint64_t)grid_sm * (OK_SB_KPT == 16 ? 3 : 2) * OK_SB_WARPS;   // resident warps (72 KB smem per CTA)
        const uint64_t n_launch = pieces ? pieces->n_pieces : 1;
        for (uint64_t p = 0; p < n_launch; ++p) {
            uint64_t t0 = tile_begin, t1 = n_tiles, visible = n_bases;
            if (pieces) {
                CU(cudaStreamWaitEvent(c->s_main, pieces->ev[p], 0));
                t0 = p * (pieces->piece_bytes / OK_TILE_BASES);
                if (p + 1 < n_launch) { visible = (p + 1) * pieces->piece_bytes; t1 = visible / OK_TILE_BASES; }
            }
            if (t1 <= t0) continue;
            const uint64_t tpw = std::max<uint64_t>(1, (t1 - t0 + max_warps - 1) / max_warps);
            const unsigned blocks = (unsigned)((t1 - t0 + OK_SB_WARPS * tpw - 1) / (OK_SB_WARPS * tpw));
            LAUNCH(kern, blocks, OK_SB_THREADS, sizeof(OkScatterSmem), c->s_main, d_bases, visible, d_off, n_rec, t0, t1, tpw, c->k,
                   pl.cfg, two ? pl.cursor1 : pl.cursor, (const unsigned*)(two ? pl.end1 : pl.cap_end), two ? c->d_buf1 : c->d_buf2,
                   (OkPartSpill{c->spill, c->d_stats}), &c->d_stats->windows, OkPeerOut{}, OkPushDesc{});
        }
    }
    CU(cudaEventRecord(c->ev_p[2], c->s_main));
    if (defer && whole && part_sliceable(pl)) {
        c->run_state = RUN_LEVEL1;
        c->pend_bases = n_bases; c->pend_rec = n_rec; c->pend_windows_before = windows_before;
        return OK_SUCCESS;
    }
    TRY(part_finish(c, pl));
    if (part_hint_misled(c, pl)) {
        TRY(part_discard(c, windows_before));
        c->distrust_hint = true;
        return part_count_bases(c, d_bases, n_bases, d_off, n_rec, d_bases, pieces, false, tile_begin, tile_end);
    }
    return part_absorb_spills(c, windows_before);
}

int part_count_keys_absorb(ok_counter* c, const uint64_t* d_keys, uint64_t n, int depth);
int part_count_keys(ok_counter* c, const uint64_t* d_keys, uint64_t n) { return part_count_keys_absorb(c, d_keys, n, 0); }
int part_count_keys_absorb(ok_counter* c, const uint64_t* d_keys, uint64_t n, int depth) {
    PartPlan& pl = c->pl; pl = PartPlan{};
    const uint64_t windows_before = c->windows;
    part_choose_bits(c, n, pl);
    const uint64_t n_chunks = (n + 255) / 256;
    pl.stride = n_chunks > 16384 ? 16 : 1;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    // a concatenation of sorted runs (ok_set_union): single-key sample + strided level-1 gather (partition.cuh)
    const bool no_stride = getenv("ORION_UNION_NO_STRIDE") != nullptr;       // A/B knob: contiguous items as for any key array
    const bool strided = c->keys_sorted_runs && depth == 0 && !((uintptr_t)d_keys & 15u) && !no_stride;
    TRY(part_layout(c, n, 256, n, pl));
    CU(cudaEventRecord(c->ev_p[0], c->s_main));
    CU(cudaMemsetAsync(pl.hist, 0, pl.n_sub * sizeof(unsigned), c->s_main));
    if (strided)
        LAUNCH(k_part_sample_keys_single, (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n / pl.stride + 255) / 256, (uint64_t)grid_sm * 16)),
               256, 0, c->s_main, (const unsigned long long*)d_keys, n, (uint64_t)pl.stride, pl.cfg, pl.hist);
    else
        LAUNCH(k_part_sample_keys, (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_chunks + pl.stride - 1) / pl.stride, (uint64_t)grid_sm * 16)),
               256, 0, c->s_main, (const unsigned long long*)d_keys, n, (uint64_t)pl.stride, pl.cfg, pl.hist);
    LAUNCH(k_part_plan_sums, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)n, pl.chunk_sum);
    LAUNCH(k_part_plan, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)n, pl.cfg.b2,
           pl.chunk_sum, (unsigned)pl.cap_bound, pl.beg, pl.cursor, pl.cap_end, pl.beg1, pl.cursor1, pl.end1, pl.scal);
    CU(cudaEventRecord(c->ev_p[1], c->s_main));
    {
        // one level-1 scatter over the whole key array, in items of 4096 keys
        const bool two = pl.cfg.b2 > 0;
        if (strided) {
            LAUNCH(k_part_set_items, 1, 1, 0, c->s_main, pl.scal, (unsigned)ok_strided_items(n));
            auto kern = k_part_scatter_keys<1, false, 0, true>;
            TRY(set_smem(kern, sizeof(OkScatterKeysSmem)));
            LAUNCH(kern, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), c->s_main, (const unsigned long long*)d_keys,
                   pl.item_off, pl.item_n, pl.item_bin, pl.scal, pl.cfg, two ? pl.cursor1 : pl.cursor,
                   (const unsigned*)(two ? pl.end1 : pl.cap_end), two ? c->d_buf1 : c->d_buf2, (OkPartSpill{c->spill, c->d_stats}),
                   (const unsigned*)nullptr, (unsigned)(n & 0xFFFFFFFFu), (unsigned)(n >> 32));
        } else {
            LAUNCH(k_part_flat_items, 64, 1024, 0, c->s_main, (unsigned)n, pl.item_off, pl.item_n, pl.item_bin, pl.scal);
            auto kern = ((uintptr_t)d_keys & 15u) ? k_part_scatter_keys<1, false> : k_part_scatter_keys<1, true>;
            TRY(set_smem(kern, sizeof(OkScatterKeysSmem)));
            LAUNCH(kern, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), c->s_main, (const unsigned long long*)d_keys,
                   pl.item_off, pl.item_n, pl.item_bin, pl.scal, pl.cfg, two ? pl.cursor1 : pl.cursor,
                   (const unsigned*)(two ? pl.end1 : pl.cap_end), two ? c->d_buf1 : c->d_buf2, (OkPartSpill{c->spill, c->d_stats}),
                   (const unsigned*)nullptr, 0u, 0u);
        }
        c->windows += n;
        CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
    }
    CU(cudaEventRecord(c->ev_p[2], c->s_main));
    TRY(part_finish(c, pl));
    return part_absorb_spills(c, windows_before, depth);
}

// compact sub-partitions [p0, p1) of the sparse run into the dense arrays
void launch_compact(ok_counter* c, unsigned p0, unsigned p1, unsigned long long* out_keys, unsigned long long* out_counts,
                    cudaStream_t stream = nullptr) {
    const PartPlan& pl = c->pl;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    LAUNCH(k_part_compact, std::min<unsigned>(p1 - p0, grid_sm * 8), 256, 0, stream ? stream : c->s_main, c->d_buf2, reinterpret_cast<const unsigned*>(c->buf1_external ? c->d_cnt : c->d_buf1), pl.beg, pl.hist,
           pl.scan, p0, p1, out_keys, out_counts);
}

// sorted sub-partition runs -> one dense sorted run in d_run_keys / d_run_counts
int run_make_dense(ok_counter* c) {
    if (c->run_state != RUN_SPARSE) return OK_SUCCESS;
    const uint64_t total = c->n_run;
    TRY(dev_reserve(&c->d_run_keys, &c->cap_run_keys, total));
    TRY(dev_reserve(&c->d_run_counts, &c->cap_run_counts, total));
    CU(cudaEventRecord(c->ev_a, c->s_main));
    if (total) launch_compact(c, 0, c->pl.n_sub, c->d_run_keys, c->d_run_counts);
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_compact, c->ev_a, c->ev_b);
    c->ms_readout = c->ms_compact;
    c->run_state = RUN_DENSE;
    return OK_SUCCESS;
}

// fold the sorted run into the general table (a later batch arrived, or the run spilled)
int run_to_table(ok_counter* c) {
    if (c->run_state == RUN_NONE) return OK_SUCCESS;
    TRY(run_make_dense(c));
    c->run_state = RUN_NONE;
    const uint64_t n = c->n_run;
    c->occupied = 0;
    c->hint = std::max<uint64_t>(c->hint, n + n / 4);
    uint64_t done = 0;
    if (n == 0 && c->h_stats->spill_n) TRY(table_rebuild(c, slots_for(c, std::max<uint64_t>(c->hint, 1))));
    while (done < n) {
        uint64_t allowed = 0;
        TRY(ensure_headroom(c, n - done, &allowed));
        const uint64_t m = std::min<uint64_t>(n - done, allowed);
        LAUNCH(k_add_kmers, grid_for(m), 256, 0, c->s_main, c->tv, c->d_stats, c->spill, c->d_run_keys + done,
               c->d_run_counts + done, m, 0);
        TRY(read_stats(c));
        CU(cudaGetLastError());
        if (c->h_stats->spill_n) {
            if (c->h_stats->spill_n > c->spill.cap) return set_err(OK_ERR_INTERNAL, "spill list overflow while folding a run");
            TRY(table_rebuild(c, 2 * c->tv.n_home));
        }
        done += m;
    }
    c->n_run = 0;
    if (c->h_stats->spill_n) TRY(table_rebuild(c, 2 * c->tv.n_home));   // also re-adds what the one-shot path spilled
    return OK_SUCCESS;
}
int acc_to_table(ok_counter* c);

// min_count filter of the run into the d_out arrays
int run_filter(ok_counter* c, uint64_t min_count, uint64_t* n_out) {
    TRY(run_make_dense(c));
    const uint64_t n = c->n_run;
    *n_out = 0;
    if (n == 0) return OK_SUCCESS;
    const uint64_t n_tiles = (n + 2047) / 2048;
    TRY(dev_reserve(&c->d_tiles, &c->cap_tiles, n_tiles + 1));
    LAUNCH(k_filter_count, (unsigned)n_tiles, 256, 0, c->s_main, c->d_run_counts, n, min_count, c->d_tiles);
    LAUNCH(k_scan_tiles, 1, 1024, 0, c->s_main, c->d_tiles, n_tiles, c->d_tiles + n_tiles);
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, c->d_tiles + n_tiles, 8, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (total) {
        TRY(dev_reserve(&c->d_out_keys, &c->cap_out_keys, total));
        TRY(dev_reserve(&c->d_out_counts, &c->cap_out_counts, total));
        LAUNCH(k_filter_write, (unsigned)n_tiles, 256, 0, c->s_main, c->d_run_keys, c->d_run_counts, n, min_count, c->d_tiles,
               c->d_out_keys, c->d_out_counts);
        CU(cudaStreamSynchronize(c->s_main));
        CU(cudaGetLastError());
    }
    *n_out = total;
    return OK_SUCCESS;
}

// ---- more than one large batch: runs are merged, not folded into a table (count.rs:48: ONE table across all files) ----
int run_unstash(ok_counter* c);

// set the counter's sorted run aside so that the next batch can take the one-shot path on its own
int run_stash(ok_counter* c) {
    if (c->n_acc) TRY(run_unstash(c));              // (never in the normal flow: every batch ends with run_unstash)
    if (c->run_state == RUN_NONE) return OK_SUCCESS;
    TRY(run_make_dense(c));
    if (c->n_run) {
        std::swap(c->d_run_keys, c->d_acc_keys); std::swap(c->cap_run_keys, c->cap_acc_keys);
        std::swap(c->d_run_counts, c->d_acc_counts); std::swap(c->cap_run_counts, c->cap_acc_counts);
        c->n_acc = c->n_run;
    }
    c->n_run = 0; c->run_state = RUN_NONE; c->occupied = 0;
    return OK_SUCCESS;
}

// the set-aside run goes into the device-wide table (the batch that followed it ended there)
int acc_to_table(ok_counter* c) {
    const uint64_t n = c->n_acc;
    c->n_acc = 0;
    uint64_t done = 0;
    c->hint = std::max<uint64_t>(c->hint, c->occupied + n + n / 4);
    while (done < n) {
        uint64_t allowed = 0;
        TRY(ensure_headroom(c, n - done, &allowed));
        const uint64_t m = std::min<uint64_t>(n - done, allowed);
        LAUNCH(k_add_kmers, grid_for(m), 256, 0, c->s_main, c->tv, c->d_stats, c->spill, c->d_acc_keys + done, c->d_acc_counts + done, m, 0);
        TRY(read_stats(c));
        CU(cudaGetLastError());
        if (c->h_stats->spill_n) {
            if (c->h_stats->spill_n > c->spill.cap) return set_err(OK_ERR_INTERNAL, "spill list overflow while folding a run");
            TRY(table_rebuild(c, 2 * c->tv.n_home));
        }
        done += m;
    }
    return OK_SUCCESS;
}

// after a batch: merge the set-aside run with the batch's run (merge.cuh); the result is the counter's run again
int run_unstash(ok_counter* c) {
    if (c->n_acc == 0) return OK_SUCCESS;
    if (c->run_state == RUN_NONE) {
        if (c->occupied || (c->tv.slots && c->h_stats->spill_n)) return acc_to_table(c);     // the batch went to the table
        // the batch held no countable window: the set-aside run is the result
        std::swap(c->d_run_keys, c->d_acc_keys); std::swap(c->cap_run_keys, c->cap_acc_keys);
        std::swap(c->d_run_counts, c->d_acc_counts); std::swap(c->cap_run_counts, c->cap_acc_counts);
        c->n_run = c->n_acc; c->n_acc = 0; c->occupied = c->n_run; c->run_state = RUN_DENSE;
        return OK_SUCCESS;
    }
    TRY(run_make_dense(c));
    const uint64_t na = c->n_acc, nb = c->n_run;
    const uint64_t n_tiles = (na + nb + OK_MG_TILE - 1) / OK_MG_TILE;
    TRY(dev_reserve(&c->d_split, &c->cap_split, n_tiles + 1));
    TRY(dev_reserve(&c->d_tiles, &c->cap_tiles, n_tiles + 1));
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    TRY(set_smem(k_merge_count, sizeof(OkMergeSmem)));
    TRY(set_smem(k_merge_write<true>, sizeof(OkMergeSmem)));
    CU(cudaEventRecord(c->ev_a, c->s_main));
    LAUNCH(k_merge_partition, grid_for(n_tiles + 1), 256, 0, c->s_main, c->d_acc_keys, na, c->d_run_keys, nb, n_tiles, c->d_split);
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)grid_sm * 4));
    LAUNCH(k_merge_count, blocks, OK_MG_THREADS, sizeof(OkMergeSmem), c->s_main, c->d_acc_keys, c->d_run_keys, c->d_split, n_tiles, c->d_tiles);
    LAUNCH(k_scan_tiles, 1, 1024, 0, c->s_main, c->d_tiles, n_tiles, c->d_tiles + n_tiles);
    unsigned long long total = 0;
    CU(cudaMemcpyAsync(&total, c->d_tiles + n_tiles, 8, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (c->cap_mrg_keys < total || c->cap_mrg_counts < total) {       // later batches add ever fewer new k-mers: some headroom saves reallocations
        TRY(dev_reserve(&c->d_mrg_keys, &c->cap_mrg_keys, total + total / 8));
        TRY(dev_reserve(&c->d_mrg_counts, &c->cap_mrg_counts, total + total / 8));
    }
    LAUNCH(k_merge_write<true>, blocks, OK_MG_THREADS, sizeof(OkMergeSmem), c->s_main, c->d_acc_keys, c->d_acc_counts, c->d_run_keys, c->d_run_counts,
           c->d_split, n_tiles, c->d_tiles, c->d_mrg_keys, c->d_mrg_counts);
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->ms_merge += ms; ++c->n_merges;
    std::swap(c->d_run_keys, c->d_mrg_keys); std::swap(c->cap_run_keys, c->cap_mrg_keys);
    std::swap(c->d_run_counts, c->d_mrg_counts); std::swap(c->cap_run_counts, c->cap_mrg_counts);
    c->n_run = total; c->occupied = total; c->n_acc = 0; c->run_state = RUN_DENSE;
    return OK_SUCCESS;
}

// a device-resident batch of any size through the one-shot path, whatever sorted run the counter already holds:
// cut into sub-batches of <= PART_MAX_UNITS bases, each counted on its own and merged into the run.
// PART_RETRY: nothing of the batch was counted (the caller goes through the table).
int part_add_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec) {
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    uint64_t max_units = PART_MAX_UNITS;
    if (const char* ev = getenv("ORION_MAX_BATCH_BASES")) { const long long v = atoll(ev); if (v >= (1 << 16)) max_units = std::min<uint64_t>((uint64_t)v, PART_MAX_UNITS); }   // test hook: sub-batches on small inputs
    const uint64_t n_sub = (n_bases + max_units - 1) / max_units;
    const uint64_t per = (n_tiles + n_sub - 1) / n_sub;
    for (uint64_t t0 = 0; t0 < n_tiles; t0 += per) {
        const uint64_t t1 = std::min(n_tiles, t0 + per);
        TRY(run_stash(c));
        int r = part_count_bases(c, d_bases, n_bases, d_off, n_rec, d_bases, nullptr, false, t0, t1);
        if (r == PART_RETRY) {
            // the one-shot path gave up on this sub-batch (nothing of it is kept): everything so far goes into the
            // table and the rest of the batch is counted there
            if (t0 == 0 && c->n_acc == 0) return PART_RETRY;
            TRY(run_to_table(c));
            TRY(acc_to_table(c));
            return counter_process_tiles(c, d_bases, n_bases, d_off, n_rec, t0, n_tiles);
        }
        if (r != OK_SUCCESS) return r;
        TRY(run_unstash(c));
        if (c->run_state == RUN_NONE && c->occupied && t1 < n_tiles)      // ended up in the table (spills): stay there
            return counter_process_tiles(c, d_bases, n_bases, d_off, n_rec, t1, n_tiles);
    }
    return OK_SUCCESS;
}

// ---- deferred batches (RUN_LEVEL1) ------------------------------------------------------------
// result slices of the sliced pipeline: whole level-1 bins and whole 1024-sub-partition scan chunks
unsigned part_slice_step(const PartPlan& pl) {
    const unsigned align = std::max(1024u, 1u << pl.cfg.b2);
    const unsigned step = std::max(align, pl.n_sub / RESULT_SLICES);
    return (step + align - 1) / align * align;
}
bool part_sliceable(const PartPlan& pl) {
    if (getenv("ORION_NO_DEFER")) return false;
    if (pl.cfg.b2 == 0 || pl.sharded) return false;
    const unsigned step = part_slice_step(pl);
    return pl.n_sub % step == 0 && pl.n_sub / step >= 4 && pl.n_sub / step <= 60;
}

// the batch of a RUN_LEVEL1 counter, re-counted through the general table (the one-shot path gave up)
int part_recount_pending(ok_counter* c) {
    TRY(run_to_table(c));
    const uint64_t n_tiles = (c->pend_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    return counter_process_tiles(c, c->d_bases, c->pend_bases, c->d_off, c->pend_rec, 0, n_tiles);
}

// finish a deferred batch in one go: whatever needs the counted result calls this first
int part_settle(ok_counter* c) {
    if (c->run_state != RUN_LEVEL1) return OK_SUCCESS;
    c->run_state = RUN_NONE;
    TRY(part_finish(c, c->pl));
    int r;
    if (part_hint_misled(c, c->pl)) {
        TRY(part_discard(c, c->pend_windows_before));
        c->distrust_hint = true;
        r = part_count_bases(c, c->d_bases, c->pend_bases, c->d_off, c->pend_rec, c->d_bases, nullptr, false);
    } else {
        r = part_absorb_spills(c, c->pend_windows_before);
    }
    return r == PART_RETRY ? part_recount_pending(c) : r;
}

// Sliced result pipeline of a deferred batch (min_count <= 1): for every slice of the key space
//   compute stream: level-2 scatter -> count -> scan        (~1 ms per slice)
//   aux stream    : compaction into the dense arrays          (as soon as the slice's offsets exist)
//   copy stream   : D2H of the slice                          (~4 ms per slice: the bottleneck)
// so all compute after the level-1 scatter hides under the copy of the result.  The result buffers are
// sized after the first slice (distinct per window and per key range, + 4 %); if a later slice does
// not fit, the remaining slices are only counted and the caller falls back to the exact-size path.
// *shipped = false: counted, but the result has not (completely) reached the host: the state is RUN_SPARSE (or the
// table, after a spill) and the caller goes on with the exact-size path
int part_finish_sliced(ok_counter* c, uint64_t** kmers, uint64_t** counts, uint64_t* n, bool* shipped) {
    *shipped = false;
    PartPlan& pl = c->pl;
    const unsigned step = part_slice_step(pl), n_slices = pl.n_sub / step;
    PartHost* h = c->h_part;
    PartHost* hd = c->h_part_dev;
    c->run_state = RUN_NONE;
    if (!c->s_aux) CU(cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
    while (c->ev_chunks.size() < 2 * (size_t)n_slices) {
        cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_chunks.push_back(e);
    }
    cudaEvent_t* ev_scan = c->ev_chunks.data();
    cudaEvent_t* ev_comp = c->ev_chunks.data() + n_slices;
    part_launch_items(c, pl);
    TRY(part_dense_prepare(c, pl));
    const bool dense = pl.dense_cap != 0;      // the count kernel writes the dense result: a slice ships as soon as it is counted
    auto issue = [&](unsigned i) -> int {
        const unsigned p0 = i * step, p1 = p0 + step;
        TRY(part_launch_range(c, pl, p0, p1, false, &hd->slice_base[i + 1], true, dense));
        if (i == 0) LAUNCH(k_part_slice_fill, 1, 1024, 0, c->s_main, pl.beg, pl.cursor, pl.cap_end, 0u, p1, &hd->slice0_windows);
        CU(cudaEventRecord(ev_scan[i], c->s_main));
        return OK_SUCCESS;
    };
    h->slice_base[0] = 0;
    CU(cudaMemcpyAsync(&h->windows_now, &c->d_stats->windows, 8, cudaMemcpyDeviceToHost, c->s_main));   // the copy engines are idle here
    TRY(issue(0));
    if (n_slices > 1) TRY(issue(1));
    void *hk = nullptr, *hc = nullptr;
    struct Blocks {                      // page-locked result blocks go back to the pool unless they are handed to the caller
        void *&k, *&c; bool keep = false;
        ~Blocks() { if (!keep) { if (k) pool_release(k); if (c) pool_release(c); } }
    } blocks{hk, hc};
    uint64_t cap = 0;
    bool shipping = true;
    for (unsigned i = 0; i < n_slices; ++i) {
        CU(cudaEventSynchronize(ev_scan[i]));
        const volatile unsigned long long* sb = h->slice_base;
        const uint64_t o0 = sb[i], o1 = sb[i + 1];
        if (i == 0) {
            // distinct k-mers of the batch, estimated from the first slice two ways: per key range (the
            // position map balances the slices) and per window (slice_fill = windows the slice holds)
            const double d0 = (double)o1;
            const double w0 = (double)*(volatile unsigned long long*)&h->slice0_windows;
            const double w_all = (double)*(volatile unsigned long long*)&h->windows_now - (double)c->pend_windows_before;
            double est = d0 * (double)n_slices;
            if (w0 > 0) est = std::max(est, d0 * (w_all / w0));
            est = std::min(est * 1.04 + 262144.0, std::max(w_all, d0) + 16.0);   // distinct <= windows
            cap = (uint64_t)est;
            if (dense) cap = std::min<uint64_t>(cap, pl.dense_cap);
            TRY(pool_alloc(&hk, cap * 8));
            TRY(pool_alloc(&hc, cap * 8));
            if (!dense) {
                TRY(dev_reserve(&c->d_run_keys, &c->cap_run_keys, cap));
                TRY(dev_reserve(&c->d_run_counts, &c->cap_run_counts, cap));
            }
        }
        if (shipping && (o1 > cap || (dense && *(volatile unsigned*)&h->dense_failed))) shipping = false;
        if (shipping && o1 > o0) {
            if (dense) {
                CU(cudaStreamWaitEvent(c->s_copy, ev_scan[i], 0));
            } else {
                CU(cudaStreamWaitEvent(c->s_aux, ev_scan[i], 0));
                launch_compact(c, i * step, (i + 1) * step, c->d_run_keys, c->d_run_counts, c->s_aux);
                CU(cudaEventRecord(ev_comp[i], c->s_aux));
                CU(cudaStreamWaitEvent(c->s_copy, ev_comp[i], 0));
            }
            CU(cudaMemcpyAsync((uint64_t*)hk + o0, c->d_run_keys + o0, (o1 - o0) * 8, cudaMemcpyDeviceToHost, c->s_copy));
            CU(cudaMemcpyAsync((uint64_t*)hc + o0, c->d_run_counts + o0, (o1 - o0) * 8, cudaMemcpyDeviceToHost, c->s_copy));
        }
        if (i + 2 < n_slices) TRY(issue(i + 2));
    }
    CU(cudaStreamSynchronize(c->s_main));
    const bool dense_ok = dense && !h->dense_failed;
    if (dense && !dense_ok) { shipping = false; TRY(part_recount_sparse(c, pl)); }      // a deferred sub-partition: everything again, sparse
    CU(cudaMemcpyAsync(&h->scal, pl.scal, sizeof(OkPartScalars), cudaMemcpyDeviceToHost, c->s_main));
    if (dense && !dense_ok) CU(cudaMemcpyAsync(&h->slice_base[n_slices], pl.scan + pl.n_sub, 8, cudaMemcpyDeviceToHost, c->s_main));
    TRY(read_stats(c));   // synchronises the compute stream
    CU(cudaStreamSynchronize(c->s_aux));
    CU(cudaStreamSynchronize(c->s_copy));
    CU(cudaGetLastError());
    const uint64_t total = h->slice_base[n_slices];
    h->total = total;
    c->ms_sample = c->ms_scatter1 = c->ms_scatter2 = c->ms_count = c->ms_compact = 0;
    c->ms_insert = c->ms_readout = 0;
    c->n_run = total; c->occupied = total;
    c->n_deferred = h->scal.n_deferred;
    c->run_state = dense_ok ? RUN_DENSE : RUN_SPARSE;
    const bool spilled = c->h_stats->spill_n != 0;
    if (part_hint_misled(c, pl)) {       // see part_hint_misled: count the batch again, sized from its windows
        TRY(part_discard(c, c->pend_windows_before));
        c->distrust_hint = true;
        const int r = part_count_bases(c, c->d_bases, c->pend_bases, c->d_off, c->pend_rec, c->d_bases, nullptr, false);
        return r == PART_RETRY ? part_recount_pending(c) : r;
    }
    if (shipping && !spilled) {
        c->run_state = RUN_DENSE;
        *kmers = (uint64_t*)hk; *counts = (uint64_t*)hc; *n = total;
        *shipped = true; blocks.keep = true;
        return OK_SUCCESS;
    }
    {   // slice boundaries for the exact-size result path of ok_counter_finish
        const unsigned fstep = std::max<unsigned>(1, pl.n_sub / RESULT_SLICES);
        pl.n_slices = (pl.n_sub + fstep - 1) / fstep; pl.slice_step = fstep;
        CU(cudaMemcpy2DAsync(h->slice_base, 8, pl.scan, (size_t)fstep * 8, 8, pl.n_slices, cudaMemcpyDeviceToHost, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        h->slice_base[pl.n_slices] = total;
    }
    if (spilled) {
        const int r = part_absorb_spills(c, c->pend_windows_before);
        if (r == PART_RETRY) TRY(part_recount_pending(c)); else if (r != OK_SUCCESS) return r;
    }
    return OK_SUCCESS;
}

// result of the counter, wherever it lives: -> device pointers
int counter_result(ok_counter* c, uint64_t min_count, const unsigned long long** dk, const unsigned long long** dc, uint64_t* n) {
    if (c->run_state != RUN_NONE) {
        if (min_count <= 1) {
            TRY(run_make_dense(c));
            *dk = c->d_run_keys; *dc = c->d_run_counts; *n = c->n_run;
            return OK_SUCCESS;
        }
        TRY(run_filter(c, min_count, n));
        *dk = c->d_out_keys; *dc = c->d_out_counts;
        return OK_SUCCESS;
    }
    TRY(counter_readout(c, min_count, n));
    *dk = c->d_out_keys; *dc = c->d_out_counts;
    return OK_SUCCESS;
}

}  // namespace

// ============================================================================ lifecycle ==
OK_EXPORT int ok_init(const int* device_ids, int n_devices) {
    if (n_devices > 1)
        return set_err(OK_ERR_INVALID_ARGUMENT, "one process drives one GPU: n_devices must be 1 (got %d)", n_devices);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return set_err(OK_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    int dev = (device_ids && n_devices == 1) ? device_ids[0] : 0;
    if (dev < 0 || dev >= count) return set_err(OK_ERR_INVALID_ARGUMENT, "device id %d out of range (0..%d)", dev, count - 1);
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, dev));
    g_sms = prop.multiProcessorCount;
    g_device = dev;
    if (const char* ev = getenv("ORION_COUNT_SEED")) g_count_seed = atoi(ev);
    if (const char* ev = getenv("ORION_SCATTER_P3")) g_scatter_p3 = atoi(ev);
    if (const char* ev = getenv("ORION_PUSH_TMA")) g_push_tma = atoi(ev);
    if (const char* ev = getenv("ORION_SLICES")) { const int v = atoi(ev); if (v >= 4 && v <= 60) RESULT_SLICES = (unsigned)v; }
    return OK_SUCCESS;
}

OK_EXPORT int ok_shutdown(void) {
    std::vector<ok_counter*> spare;
    { std::lock_guard<std::mutex> lk(g_mu); spare.swap(g_spare_builders); }
    for (ok_counter* c : spare) ok_counter_destroy(c);
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& b : g_pool) cudaFreeHost(b.p);
    g_pool.clear();
    for (auto& st : g_set_streams) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; }
    if (g_cur_slab && g_cur_slab->live == 0) { cudaFree(g_cur_slab->base); delete g_cur_slab; g_cur_slab = nullptr; }
    g_device = -1;
    return OK_SUCCESS;
}

OK_EXPORT const char* ok_last_error(void) { return g_err.c_str(); }
OK_EXPORT const char* ok_version(void) { return "orion-kmer-b200 0.1 (sm_100a)"; }
OK_EXPORT uint64_t ok_launch_count(void) { return g_launches.load(); }
OK_EXPORT int ok_synchronize(void) { TRY(ensure_init()); CU(cudaDeviceSynchronize()); return OK_SUCCESS; }

OK_EXPORT int ok_host_alloc(void** out, uint64_t bytes) {
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_host_alloc: out is NULL");
    TRY(ensure_init());
    CU(cudaMallocHost(out, bytes ? bytes : 8));
    return OK_SUCCESS;
}
OK_EXPORT int ok_host_free(void* p) { if (p) CU(cudaFreeHost(p)); return OK_SUCCESS; }
OK_EXPORT int ok_free(void* p) {
    if (!p) return OK_SUCCESS;
    if (!pool_release(p)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_free: pointer was not handed out by this library");
    return OK_SUCCESS;
}

// ==================================================================== host k-mer arithmetic ==
OK_EXPORT int ok_seq_to_u64(const uint8_t* seq, uint64_t len, uint8_t k, uint64_t* out) {
    if (k == 0 || k > 32 || len != k || !seq || !out) return 0;  // kmer.rs:38-43
    uint64_t v = 0;
    for (unsigned i = 0; i < k; ++i) {
        uint32_t c8, v4;
        ok_pack4<false>((uint32_t)seq[i], c8, v4);   // same table the kernels use
        if (!(v4 & 8u)) return 0;
        v = (v << 2) | (c8 >> 6);
    }
    *out = v;
    return 1;
}
OK_EXPORT int ok_u64_to_seq(uint64_t kmer, uint8_t k, uint8_t* out) {
    if (k == 0 || k > 32) return set_err(OK_ERR_INVALID_KMER_SIZE, "Invalid k-mer length for decoding: %u", (unsigned)k);
    for (unsigned i = 0; i < k; ++i) out[i] = (uint8_t)"ACGT"[(kmer >> (2 * (k - 1 - i))) & 3u];
    return OK_SUCCESS;
}
OK_EXPORT int ok_reverse_complement_u64(uint64_t kmer, uint8_t k, uint64_t* out) {
    if (k == 0 || k > 32) return set_err(OK_ERR_INVALID_KMER_SIZE, "Invalid k-mer length for reverse complement: %u", (unsigned)k);
    *out = ok_revcomp(kmer, k);
    return OK_SUCCESS;
}
OK_EXPORT int ok_canonical_u64(uint64_t kmer, uint8_t k, uint64_t* out) {
    if (k == 0 || k > 32) return set_err(OK_ERR_INVALID_KMER_SIZE, "Invalid k-mer length for reverse complement: %u", (unsigned)k);
    *out = ok_canonical(kmer, k);
    return OK_SUCCESS;
}

// ================================================================================ counter ==
OK_EXPORT int ok_counter_create(uint8_t k, int norm_mode, uint64_t capacity_hint, ok_counter** out) {
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_create: out is NULL");
    *out = nullptr;
    if (k == 0 || k > 32) return invalid_k(k);  // count.rs:43-45
    if (norm_mode != OK_NORM_NORMALIZED && norm_mode != OK_NORM_RAW)
        return set_err(OK_ERR_INVALID_ARGUMENT, "unknown norm_mode %d", norm_mode);
    TRY(ensure_init());
    ok_counter* c = new ok_counter();
    c->k = k; c->norm_mode = norm_mode; c->hint = capacity_hint; c->user_hint = capacity_hint;
    auto fail = [&](int code) { ok_counter_destroy(c); return code; };
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(set_err(OK_ERR_CUDA, "CUDA error %s in ok_counter_create", cudaGetErrorName(e_))); } while (0)
    CUF(cudaStreamCreateWithFlags(&c->s_main, cudaStreamNonBlocking));
    CUF(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CUF(cudaEventCreate(&c->ev_a));
    CUF(cudaEventCreate(&c->ev_b));
    for (auto& e : c->ev_p) CUF(cudaEventCreate(&e));
    CUF(cudaMalloc((void**)&c->d_stats, sizeof(OkDevStats)));
    CUF(cudaMemset(c->d_stats, 0, sizeof(OkDevStats)));
    CUF(cudaMallocHost((void**)&c->h_stats, sizeof(OkDevStats)));
    memset(c->h_stats, 0, sizeof(OkDevStats));
    CUF(cudaMallocHost((void**)&c->h_part, sizeof(PartHost)));
    memset(c->h_part, 0, sizeof(PartHost));
    CUF(cudaHostGetDevicePointer((void**)&c->h_part_dev, c->h_part, 0));
    c->spill.cap = SPILL_CAP;
    CUF(cudaMalloc((void**)&c->spill.keys, SPILL_CAP * 8));
    CUF(cudaMalloc((void**)&c->spill.incs, SPILL_CAP * 8));
#undef CUF
    *out = c;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_destroy(ok_counter* c) {
    if (!c) return OK_SUCCESS;
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->s_copy) cudaStreamSynchronize(c->s_copy);
    if (c->s_aux) { cudaStreamSynchronize(c->s_aux); cudaStreamDestroy(c->s_aux); }
    cudaFree(c->tv.slots); cudaFree(c->d_stats); cudaFreeHost(c->h_stats);
    cudaFree(c->spill.keys); cudaFree(c->spill.incs);
    cudaFree(c->d_bases); cudaFree(c->d_off); cudaFree(c->d_tiles);
    cudaFree(c->d_out_keys); cudaFree(c->d_out_counts);
    cudaFree(c->d_run_keys); cudaFree(c->d_run_counts); if (!c->buf1_external) cudaFree(c->d_buf1); cudaFree(c->d_buf2); cudaFree(c->d_cnt);
    cudaFree(c->shard.d_state); cudaFree(c->shard.d_received); cudaFree(c->shard.d_snap); cudaFreeHost(c->shard.h_blk);
    cudaFree(c->shard.d_send); cudaFree(c->shard.d_xchg);
    for (auto& row : c->shard.s_peer) for (auto st : row) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto& row : c->shard.ev_piece) for (auto e : row) if (e) cudaEventDestroy(e);
    for (auto st : {c->shard.s_join, c->shard.s_recv}) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto e : c->shard.ev_sent) if (e) cudaEventDestroy(e);
    for (auto e : c->shard.ev_recv) if (e) cudaEventDestroy(e);
    for (auto e : c->shard.ev_chunk) if (e) cudaEventDestroy(e);
    cudaFree(c->d_meta); cudaFreeHost(c->h_part);
    cudaFree(c->d_acc_keys); cudaFree(c->d_acc_counts); cudaFree(c->d_mrg_keys); cudaFree(c->d_mrg_counts); cudaFree(c->d_split);
    for (auto e : c->ev_p) if (e) cudaEventDestroy(e);
    for (auto e : c->ev_chunks) cudaEventDestroy(e);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->s_main) cudaStreamDestroy(c->s_main);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    delete c;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_clear(ok_counter* c) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_clear: NULL handle");
    if (c->run_state == RUN_LEVEL1) { CU(cudaStreamSynchronize(c->s_main)); c->run_state = RUN_NONE; }   // a deferred batch is simply dropped
    if (c->tv.slots && c->occupied && c->run_state == RUN_NONE)
        LAUNCH(k_fill_slots, grid_for(c->tv.n_total, 256, 16), 256, 0, c->s_main, c->tv.slots, c->tv.n_total);
    c->run_state = RUN_NONE; c->n_run = 0; c->n_acc = 0;
    CU(cudaMemsetAsync(c->d_stats, 0, sizeof(OkDevStats), c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    c->occupied = c->windows = c->bases_seen = c->max_disp = c->spilled_total = 0;
    c->ms_insert = c->ms_readout = c->ms_fill = 0;
    c->ms_merge = 0; c->n_merges = 0;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_add_batch_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                          const uint64_t* d_rec_offsets, uint64_t n_records) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_add_batch_device: NULL handle");
    if (n_bases == 0 || n_records == 0) return OK_SUCCESS;
    if (!d_bases || !d_rec_offsets) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL batch pointer");
    if ((uintptr_t)d_bases & 15u) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    TRY(part_settle(c));
    c->ms_insert = 0; c->ms_fill = 0;
    if (part_eligible(c, n_bases)) {
        const int r = part_add_device(c, d_bases, n_bases, d_rec_offsets, n_records);
        if (r != PART_RETRY) { if (r == OK_SUCCESS) c->bases_seen += n_bases; return r; }
    }
    TRY(run_to_table(c));
    TRY(acc_to_table(c));
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    TRY(counter_process_tiles(c, d_bases, n_bases, d_rec_offsets, n_records, 0, n_tiles));
    c->bases_seen += n_bases;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_add_batch(ok_counter* c, const uint8_t* bases, const uint64_t* rec_offsets,
                                   uint64_t n_records) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_add_batch: NULL handle");
    if (n_records == 0) return OK_SUCCESS;
    if (!rec_offsets) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL rec_offsets");
    if (rec_offsets[0] != 0) return set_err(OK_ERR_INVALID_ARGUMENT, "rec_offsets[0] must be 0");
    const uint64_t n_bases = rec_offsets[n_records];
    if (n_bases == 0) return OK_SUCCESS;
    if (!bases) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL bases");
    TRY(part_settle(c));               // a deferred batch still lives in d_bases
    c->ms_insert = 0; c->ms_fill = 0;
    TRY(dev_reserve(&c->d_bases, &c->cap_bases, n_bases + 64));
    TRY(dev_reserve(&c->d_off, &c->cap_off, n_records + 1));
    // copy stream runs ahead piece by piece; the compute stream follows the events
    const uint64_t n_pieces = (n_bases + COPY_CHUNK - 1) / COPY_CHUNK;
    while (c->ev_chunks.size() < n_pieces + 1) {
        cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_chunks.push_back(e);
    }
    CU(cudaMemcpyAsync(c->d_off, rec_offsets, (n_records + 1) * 8, cudaMemcpyHostToDevice, c->s_copy));
    CU(cudaEventRecord(c->ev_chunks[n_pieces], c->s_copy));
    for (uint64_t p = 0; p < n_pieces; ++p) {
        const uint64_t b0 = p * COPY_CHUNK, b1 = std::min(n_bases, b0 + COPY_CHUNK);
        CU(cudaMemcpyAsync(c->d_bases + b0, bases + b0, b1 - b0, cudaMemcpyHostToDevice, c->s_copy));
        CU(cudaEventRecord(c->ev_chunks[p], c->s_copy));
    }
    CU(cudaStreamWaitEvent(c->s_main, c->ev_chunks[n_pieces], 0));
    if (part_eligible(c, n_bases) && (c->run_state != RUN_NONE || n_bases > PART_MAX_UNITS || getenv("ORION_MAX_BATCH_BASES"))) {
        // a later batch of a multi-batch job, or one too large for a single pass: land it, then sub-batches + merge
        CU(cudaStreamWaitEvent(c->s_main, c->ev_chunks[n_pieces - 1], 0));
        const int r = part_add_device(c, c->d_bases, n_bases, c->d_off, n_records);
        CU(cudaStreamSynchronize(c->s_copy));
        if (r != PART_RETRY) { if (r == OK_SUCCESS) c->bases_seen += n_bases; return r; }
    } else if (part_eligible(c, n_bases)) {
        // The level-1 scatter follows the pieces as they land.  The plan needs a sample of the WHOLE
        // batch first: a page-locked caller buffer is sampled in place (zero-copy reads over PCIe,
        // 1/16 of the tiles); a pageable one only after the last piece has landed.
        const uint8_t* sample_src = c->d_bases;
        cudaPointerAttributes attr{};
        const bool mapped = cudaPointerGetAttributes(&attr, bases) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
                            attr.devicePointer && ((uintptr_t)attr.devicePointer & 15u) == 0 && !getenv("ORION_NO_ZEROCOPY");
        cudaGetLastError();
        if (mapped) sample_src = (const uint8_t*)attr.devicePointer;
        else CU(cudaStreamWaitEvent(c->s_main, c->ev_chunks[n_pieces - 1], 0));
        const PieceSchedule ps{n_pieces, COPY_CHUNK, c->ev_chunks.data()};
        // Deferred: level 2, count and compaction run in ok_counter_finish, slice by slice under the D2H copy
        // of the result (or in part_settle, if anything else is asked of the counter first).
        const int r = part_count_bases(c, c->d_bases, n_bases, c->d_off, n_records, sample_src, &ps, /*defer=*/true);
        if (r == OK_SUCCESS && c->run_state == RUN_LEVEL1) {     // deferred: return as soon as the caller's buffers are free again
            CU(cudaStreamSynchronize(c->s_copy));               //   every piece has landed
            CU(cudaEventSynchronize(c->ev_b));                  //   and the sampling kernel (zero-copy reads of `bases`) is done
        }
        if (r != PART_RETRY) { if (r == OK_SUCCESS) c->bases_seen += n_bases; return r; }
    }
    TRY(run_to_table(c));
    TRY(acc_to_table(c));
    const uint64_t tiles_per_piece = COPY_CHUNK / OK_TILE_BASES;
    for (uint64_t p = 0; p < n_pieces; ++p) {
        CU(cudaStreamWaitEvent(c->s_main, c->ev_chunks[p], 0));
        const uint64_t t0 = p * tiles_per_piece;
        const uint64_t b1 = std::min(n_bases, (p + 1) * COPY_CHUNK);
        // a tile is only complete once the piece holding its last base has landed
        const uint64_t t1 = (p + 1 == n_pieces) ? (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES : b1 / OK_TILE_BASES;
        // the kernel only reads bases below n_visible: pass the landed prefix as the batch length
        TRY(counter_process_tiles(c, c->d_bases, (p + 1 == n_pieces) ? n_bases : b1, c->d_off, n_records, t0, t1));
    }
    c->bases_seen += n_bases;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_add_kmers_device(ok_counter* c, const uint64_t* d_kmers, uint64_t n) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_add_kmers_device: NULL handle");
    if (n == 0) return OK_SUCCESS;
    if (!d_kmers) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL d_kmers");
    TRY(part_settle(c));
    c->ms_insert = 0; c->ms_fill = 0;
    if (part_eligible(c, n) && n < (1ull << 31)) {
        TraceClock tc;
        TRY(run_stash(c));
        const int r = part_count_keys(c, d_kmers, n);
        tc.lap("add_kmers %llu keys: r=%d run=%llu deferred=%llu spilled=%llu sample %.2f l1 %.2f l2 %.2f count %.2f", (unsigned long long)n, r,
               (unsigned long long)c->n_run, (unsigned long long)c->n_deferred, (unsigned long long)c->spilled_total, c->ms_sample, c->ms_scatter1,
               c->ms_scatter2, c->ms_count);
        if (r == OK_SUCCESS) return run_unstash(c);
        if (r != PART_RETRY) return r;
    }
    TRY(run_to_table(c));
    TRY(acc_to_table(c));
    uint64_t done = 0;
    while (done < n) {
        uint64_t allowed = 0;
        TRY(ensure_headroom(c, n - done, &allowed));
        const uint64_t m = std::min<uint64_t>(n - done, allowed);
        CU(cudaEventRecord(c->ev_a, c->s_main));
        LAUNCH(k_add_kmers, grid_for(m), 256, 0, c->s_main, c->tv, c->d_stats, c->spill,
               (const unsigned long long*)d_kmers + done, (const unsigned long long*)nullptr, m, 1);
        CU(cudaEventRecord(c->ev_b, c->s_main));
        TRY(read_stats(c));
        CU(cudaGetLastError());
        float ms = 0; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->ms_insert += ms;
        if (c->h_stats->spill_n) {
            if (c->h_stats->spill_n > c->spill.cap) return set_err(OK_ERR_INTERNAL, "spill list overflow: k-mer keys too clustered");
            c->spilled_total += c->h_stats->spill_n;
            TRY(table_rebuild(c, 2 * c->tv.n_home));
        }
        done += m;
    }
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_set_shard(ok_counter* c, int rank, int n_ranks) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_set_shard: NULL handle");
    if (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8)
        return set_err(OK_ERR_INVALID_ARGUMENT, "n_ranks must be 1, 2, 4 or 8 (got %d)", n_ranks);
    if (rank < 0 || rank >= n_ranks) return set_err(OK_ERR_INVALID_ARGUMENT, "rank %d out of range", rank);
    if (c->tv.slots) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_set_shard must precede the first batch");
    c->shard_rank = rank; c->n_shards = n_ranks;
    return OK_SUCCESS;
}

namespace {
template <int G, int PASS>
void launch_route(const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec, unsigned k,
                  int norm_mode, unsigned long long* cursors, unsigned long long* out, cudaStream_t st) {
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    const uint64_t max_warps = (uint64_t)(g_sms > 0 ? g_sms : 148) * 8 * 8;
    const uint64_t tpw = std::max<uint64_t>(1, (n_tiles + max_warps - 1) / max_warps);
    const unsigned blocks = (unsigned)(((n_tiles + tpw - 1) / tpw + 7) / 8);
    auto kern = norm_mode == OK_NORM_NORMALIZED ? k_route<true, G, PASS> : k_route<false, G, PASS>;
    LAUNCH(kern, blocks, 256, 0, st, d_bases, n_bases, d_off, n_rec, (uint64_t)0, n_tiles, tpw, k, 64 - 2 * k,
           (int)OK_MAP_CANON, cursors, out);
}
template <int PASS>
void launch_route_g(int g, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec,
                    unsigned k, int norm_mode, unsigned long long* cursors, unsigned long long* out, cudaStream_t st) {
    switch (g) {
        case 1: launch_route<1, PASS>(d_bases, n_bases, d_off, n_rec, k, norm_mode, cursors, out, st); break;
        case 2: launch_route<2, PASS>(d_bases, n_bases, d_off, n_rec, k, norm_mode, cursors, out, st); break;
        case 4: launch_route<4, PASS>(d_bases, n_bases, d_off, n_rec, k, norm_mode, cursors, out, st); break;
        default: launch_route<8, PASS>(d_bases, n_bases, d_off, n_rec, k, norm_mode, cursors, out, st); break;
    }
}
}  // namespace

OK_EXPORT int ok_counter_route_batch_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                            const uint64_t* d_rec_offsets, uint64_t n_records, int n_ranks,
                                            uint64_t* d_out, uint64_t* out_counts) {
    if (!c || !out_counts) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_route_batch_device: NULL argument");
    if (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8)
        return set_err(OK_ERR_INVALID_ARGUMENT, "n_ranks must be 1, 2, 4 or 8 (got %d)", n_ranks);
    for (int r = 0; r < n_ranks; ++r) out_counts[r] = 0;
    if (n_bases == 0 || n_records == 0) return OK_SUCCESS;
    if (!d_bases || !d_rec_offsets || !d_out) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL batch pointer");
    if ((uintptr_t)d_bases & 15u) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    unsigned long long* cur = c->d_stats->route_counts;
    CU(cudaEventRecord(c->ev_a, c->s_main));
    CU(cudaMemsetAsync(cur, 0, 8 * sizeof(unsigned long long), c->s_main));
    launch_route_g<0>(n_ranks, d_bases, n_bases, d_rec_offsets, n_records, c->k, c->norm_mode, cur, nullptr, c->s_main);
    unsigned long long h[8] = {0};
    CU(cudaMemcpyAsync(h, cur, sizeof h, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    unsigned long long base[8] = {0}, run = 0;
    for (int r = 0; r < n_ranks; ++r) { out_counts[r] = h[r]; base[r] = run; run += h[r]; }
    CU(cudaMemcpyAsync(cur, base, sizeof base, cudaMemcpyHostToDevice, c->s_main));
    launch_route_g<1>(n_ranks, d_bases, n_bases, d_rec_offsets, n_records, c->k, c->norm_mode, cur,
                      (unsigned long long*)d_out, c->s_main);
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_route, c->ev_a, c->ev_b);
    return OK_SUCCESS;
}

// ---- fused multi-GPU routing: the owner multisplit writes straight into the owners' receive buffers ----
OK_EXPORT int ok_peer_buffer_create(uint64_t bytes, void** d_ptr, uint8_t handle[64]) {
    if (!d_ptr || !handle) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_peer_buffer_create: NULL argument");
    TRY(ensure_init());
    CU(cudaMalloc(d_ptr, bytes ? bytes : 8));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle, &h, 64);
    return OK_SUCCESS;
}
OK_EXPORT int ok_peer_buffer_open(const uint8_t handle[64], void** d_ptr) {
    if (!d_ptr || !handle) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_peer_buffer_open: NULL argument");
    TRY(ensure_init());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OK_SUCCESS;
}
OK_EXPORT int ok_peer_buffer_close(void* d_ptr) { if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr)); return OK_SUCCESS; }
OK_EXPORT int ok_peer_buffer_destroy(void* d_ptr) { if (d_ptr) CU(cudaFree(d_ptr)); return OK_SUCCESS; }

// pass 0 of the routing: how many k-mers of this batch each rank owns (out_counts: n_ranks host entries)
OK_EXPORT int ok_counter_route_count_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                            const uint64_t* d_rec_offsets, uint64_t n_records, int n_ranks,
                                            uint64_t* out_counts) {
    if (!c || !out_counts) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_route_count_device: NULL argument");
    if (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8)
        return set_err(OK_ERR_INVALID_ARGUMENT, "n_ranks must be 1, 2, 4 or 8 (got %d)", n_ranks);
    for (int r = 0; r < n_ranks; ++r) out_counts[r] = 0;
    if (n_bases == 0 || n_records == 0) return OK_SUCCESS;
    if ((uintptr_t)d_bases & 15u) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    unsigned long long* cur = c->d_stats->route_counts;
    CU(cudaEventRecord(c->ev_a, c->s_main));
    CU(cudaMemsetAsync(cur, 0, 8 * sizeof(unsigned long long), c->s_main));
    launch_route_g<0>(n_ranks, d_bases, n_bases, d_rec_offsets, n_records, c->k, c->norm_mode, cur, nullptr, c->s_main);
    unsigned long long h[8] = {0};
    CU(cudaMemcpyAsync(h, cur, sizeof h, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_route, c->ev_a, c->ev_b);
    for (int r = 0; r < n_ranks; ++r) out_counts[r] = h[r];
    return OK_SUCCESS;
}

// pass 1: extract + multisplit by owner; rank r's k-mers are written, in runs, to d_dst[r][0 .. counts[r])
// (d_dst[r] = the slice of rank r's receive buffer reserved for this sender: peer memory over NVLink)
OK_EXPORT int ok_counter_route_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                              const uint64_t* d_rec_offsets, uint64_t n_records, int n_ranks,
                                              uint64_t* const* d_dst, const uint64_t* counts) {
    if (!c || !d_dst || !counts) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_route_scatter_device: NULL argument");
    if (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8)
        return set_err(OK_ERR_INVALID_ARGUMENT, "n_ranks must be 1, 2, 4 or 8 (got %d)", n_ranks);
    if (n_bases == 0 || n_records == 0) return OK_SUCCESS;
    if ((uintptr_t)d_bases & 15u) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    OkPartCfg cfg{};
    cfg.key_shift = 64 - 2 * c->k; cfg.shard_log2 = 0; cfg.b2 = 0; cfg.b1 = 0;
    for (int g = n_ranks; g > 1; g >>= 1) ++cfg.b1;
    OkPeerOut po{}; po.shift = 0;
    unsigned zero[8] = {0}, ends[8] = {0};
    for (int r = 0; r < n_ranks; ++r) {
        if (counts[r] >= (1ull << 32)) return set_err(OK_ERR_INVALID_ARGUMENT, "more than 2^32 k-mers for one rank in one batch");
        po.p[r] = (unsigned long long*)d_dst[r]; ends[r] = (unsigned)counts[r];
    }
    TRY(part_settle(c)); TRY(run_stash(c));
    TRY(dev_reserve(&c->d_meta, &c->cap_meta, 64));
    unsigned *d_cur = c->d_meta, *d_end = c->d_meta + 8;
    CU(cudaEventRecord(c->ev_a, c->s_main));
    CU(cudaMemcpyAsync(d_cur, zero, sizeof zero, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_end, ends, sizeof ends, cudaMemcpyHostToDevice, c->s_main));
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    const uint64_t max_warps = (uint64_t)grid_sm * (OK_SB_KPT == 16 ? 3 : 2) * OK_SB_WARPS;
    const uint64_t tpw = std::max<uint64_t>(1, (n_tiles + max_warps - 1) / max_warps);
    const unsigned blocks = (unsigned)((n_tiles + OK_SB_WARPS * tpw - 1) / (OK_SB_WARPS * tpw));
    auto kern = c->norm_mode == OK_NORM_NORMALIZED ? k_part_scatter_bases<true, true> : k_part_scatter_bases<false, true>;
    TRY(set_smem(kern, sizeof(OkScatterSmem)));
    LAUNCH(kern, blocks, OK_SB_THREADS, sizeof(OkScatterSmem), c->s_main, d_bases, n_bases, d_rec_offsets, n_records, (uint64_t)0, n_tiles, tpw, c->k,
           cfg, d_cur, (const unsigned*)d_end, (unsigned long long*)nullptr, (OkPartSpill{c->spill, c->d_stats}),
           c->d_stats->route_counts, po, OkPushDesc{});
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->ms_route += ms;
    return OK_SUCCESS;
}

// ---- fused multi-GPU exchange, second form: no counting pass, and the level-1 scatter of the owner happens on the sender ----
// Per batch and rank:  ok_shard_sample_device -> [reduce-scatter of the fine histogram, all-gather of the
// level-1 histograms] -> ok_shard_scatter_device (extract + multisplit by (owner, level-1 bin) straight
// into the owners' level-1 regions over NVLink) -> [all-gather of the cursors = the barrier] ->
// ok_shard_count_device (level-2 scatter + count of what arrived).
OK_EXPORT int ok_shard_geometry(ok_counter* c, uint64_t n_bases_max, uint32_t* sub_bits, uint32_t* l1_bits, uint64_t* buffer_keys) {
    if (!c || !sub_bits || !l1_bits || !buffer_keys) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_geometry: NULL argument");
    if (c->n_shards < 2) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_geometry: call ok_counter_set_shard first (n_ranks >= 2)");
    if (n_bases_max == 0 || n_bases_max >= (1ull << 31)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_geometry: batch size out of range");
    ShardState& sh = c->shard;
    sh.g = 0; for (int g = c->n_shards; g > 1; g >>= 1) ++sh.g;
    // what a rank receives is balanced by the canonical prior: plan for 1.25 x the largest batch
    const uint64_t n_units = (uint64_t)((double)n_bases_max * c->shard.margin);
    PartPlan pl;
    part_choose_bits(c, n_bases_max, pl, /*use_hint=*/true);
    sh.hinted = pl.hinted;
    unsigned bits = pl.cfg.b1 + pl.cfg.b2;
    sh.b1 = std::min(pl.cfg.b1, 10u - sh.g);                 // sender bins = owners x level-1 bins <= 1024
    if (sh.b1 >= bits && bits >= 2) sh.b1 = bits / 2;          // the sharded count always runs a second level
    if (const char* ev = getenv("ORION_SHARD_B1")) { unsigned v = (unsigned)atoi(ev); if (v >= 1 && v + sh.g <= 10 && v <= bits) sh.b1 = v; }
    if (bits - sh.b1 > 10) bits = sh.b1 + 10;                // level 2 has at most 1024 bins: larger sub-partitions instead
    sh.b2 = bits - sh.b1; sh.sub_bits = bits;
    sh.stride = 16; sh.n_bases_max = n_bases_max;
    sh.cap_keys = part_cap_bound(n_units, 1u << bits, sh.stride, OK_TILE_BASES) + 16;
    if (sh.cap_keys >= (1ull << 32)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_geometry: batch too large for 32-bit offsets");
    if (!sh.d_state) {
        CU(cudaMalloc((void**)&sh.d_state, 5 * 1024 * sizeof(unsigned) + sizeof(OkShardBlocks)));
        CU(cudaMallocHost((void**)&sh.h_blk, sizeof(OkShardBlocks)));
        CU(cudaMalloc((void**)&sh.d_snap, 9 * 1024 * sizeof(unsigned)));
        CU(cudaMalloc((void**)&sh.d_received, 8));
        sh.reg_beg = sh.d_state; sh.reg_end = sh.d_state + 1024; sh.reg_fill = sh.d_state + 2048;
        sh.send_cur = sh.d_state + 3072; sh.send_end = sh.d_state + 4096; sh.d_blk = (OkShardBlocks*)(sh.d_state + 5120);
    }
    sh.ready = true; sh.buffers = false;
    *sub_bits = sh.sub_bits; *l1_bits = sh.b1; *buffer_keys = sh.cap_keys;
    return OK_SUCCESS;
}

// What a rank receives is balanced by the canonical-k-mer prior: exactly for uniform base composition, within a few
// per cent for repeats and microsatellites, but a 35 %-GC genome puts 1.6 x the mean on one of 8 owners
// (tools/skew.py).  The caller learns the real shares from the summed sample (before anything is sent) and can
// raise the margin; the geometry must be taken again afterwards.
OK_EXPORT int ok_shard_set_margin(ok_counter* c, double margin) {
    if (!c || !(margin >= 1.0 && margin <= 8.0)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_set_margin: margin must be in [1, 8]");
    c->shard.margin = margin; c->shard.ready = false;
    return OK_SUCCESS;
}

// d_peer_buffers[r] = rank r's level-1 buffer as mapped in THIS process (own buffer included), cap_keys each
OK_EXPORT int ok_shard_set_buffers(ok_counter* c, void* const* d_peer_buffers, uint64_t cap_keys) {
    if (!c || !d_peer_buffers) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_set_buffers: NULL argument");
    ShardState& sh = c->shard;
    if (!sh.ready) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_set_buffers: call ok_shard_geometry first");
    if (cap_keys < sh.cap_keys) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_set_buffers: buffers hold %llu keys, %llu needed",
                                               (unsigned long long)cap_keys, (unsigned long long)sh.cap_keys);
    for (int r = 0; r < c->n_shards; ++r) {
        if (!d_peer_buffers[r] || ((uintptr_t)d_peer_buffers[r] & 15u)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_set_buffers: bad buffer of rank %d", r);
        sh.peer[r] = (unsigned long long*)d_peer_buffers[r];
    }
    sh.recv_cap = cap_keys;
    if (sh.three) {
        // the peer-mapped buffer only RECEIVES (raw keys per sender and chunk); the level-1 buffer is the counter's own
        if (c->buf1_external) { c->d_buf1 = nullptr; c->cap_buf1 = 0; c->buf1_external = false; }
    } else {
        if (!c->buf1_external && c->d_buf1) { cudaFree(c->d_buf1); }
        c->d_buf1 = sh.peer[c->shard_rank]; c->cap_buf1 = cap_keys; c->buf1_external = true;
    }
    sh.buffers = true;
    return OK_SUCCESS;
}

namespace {
int shard_check(ok_counter* c, const char* who, uint64_t n_bases) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "%s: NULL handle", who);
    if (!c->shard.ready || !c->shard.buffers) return set_err(OK_ERR_INVALID_ARGUMENT, "%s: ok_shard_geometry / ok_shard_set_buffers first", who);
    if (n_bases > c->shard.n_bases_max) return set_err(OK_ERR_INVALID_ARGUMENT, "%s: batch larger than the agreed maximum", who);
    if (c->run_state != RUN_NONE || c->occupied) return set_err(OK_ERR_INVALID_ARGUMENT, "%s: the counter holds a result; clear it first", who);
    return OK_SUCCESS;
}
OkPartCfg shard_global_cfg(const ok_counter* c, unsigned bits) {   // bins over the WHOLE key space: (owner, ...) ids
    OkPartCfg cfg{}; cfg.key_shift = 64 - 2 * c->k; cfg.shard_log2 = 0; cfg.b1 = c->shard.g + bits; cfg.b2 = 0;
    return cfg;
}
}  // namespace

// d_hist_fine[n_ranks << sub_bits] += sampled k-mers per (owner, sub-partition); d_hist_l1[n_ranks << l1_bits] = per (owner, level-1 bin)
OK_EXPORT int ok_shard_sample_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                     uint64_t n_records, uint32_t* d_hist_fine, uint32_t* d_hist_l1) {
    if (c) { TRY(part_settle(c)); TRY(run_stash(c)); }      // an earlier batch's run is set aside and merged after the count
    TRY(shard_check(c, "ok_shard_sample_device", n_bases));
    if (!d_hist_fine || !d_hist_l1) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_sample_device: NULL histogram");
    if (n_bases && ((uintptr_t)d_bases & 15u)) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    const ShardState& sh = c->shard;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    CU(cudaEventRecord(c->ev_a, c->s_main));
    CU(cudaMemsetAsync(d_hist_fine, 0, ((size_t)c->n_shards << sh.sub_bits) * sizeof(unsigned), c->s_main));
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    if (n_tiles && n_records) {
        const uint64_t sampled = (n_tiles + sh.stride - 1) / sh.stride;
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((sampled + 7) / 8, (uint64_t)grid_sm * 8));
        auto kern = c->norm_mode == OK_NORM_NORMALIZED ? k_part_sample<true> : k_part_sample<false>;
        LAUNCH(kern, blocks, 256, 0, c->s_main, d_bases, n_bases, d_rec_offsets, n_records, n_tiles, (uint64_t)sh.stride, c->k,
               shard_global_cfg(c, sh.sub_bits), d_hist_fine, /*halo=*/true, (uint64_t)0);
    }
    LAUNCH(k_shard_l1_hist, (unsigned)(c->n_shards << sh.b1), 128, 0, c->s_main, d_hist_fine, sh.b2, d_hist_l1);
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_route, c->ev_a, c->ev_b);
    return OK_SUCCESS;
}

// d_hist_mine[1 << sub_bits]: the fine histogram summed over the ranks, this rank's slice (reduce-scatter);
// d_hist_l1_all[n_ranks][n_ranks << l1_bits]: every rank's level-1 histogram (all-gather);
// d_cursors_out[n_ranks << l1_bits]: where this sender stopped in each of its regions (to be all-gathered).
OK_EXPORT int ok_shard_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                      uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* d_hist_l1_all,
                                      uint32_t* d_cursors_out) {
    TRY(shard_check(c, "ok_shard_scatter_device", n_bases));
    if (!d_hist_mine || !d_hist_l1_all || !d_cursors_out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_scatter_device: NULL argument");
    ShardState& sh = c->shard;
    PartPlan& pl = c->pl; pl = PartPlan{};
    pl.cfg.key_shift = 64 - 2 * c->k; pl.cfg.shard_log2 = sh.g; pl.cfg.b1 = sh.b1; pl.cfg.b2 = sh.b2;
    pl.n_sub = 1u << sh.sub_bits; pl.n_bin1 = 1u << sh.b1; pl.stride = sh.stride; pl.sharded = true;
    const uint64_t n_units = (uint64_t)((double)sh.n_bases_max * sh.margin);
    pl.hinted = sh.hinted;
    pl.big_count = sh.hinted ? std::min<uint64_t>(c->user_hint, sh.n_bases_max) / pl.n_sub > 4600     // expected distinct keys per sub-partition
                             : sh.n_bases_max / pl.n_sub > 5800;     // what arrives is balanced: about one batch worth of k-mers
    if (const char* ev = getenv("ORION_BIG_COUNT")) pl.big_count = atoi(ev) != 0;
    TRY(part_layout(c, n_units, OK_TILE_BASES, 0, pl));
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    CU(cudaEventRecord(c->ev_p[0], c->s_main));
    // my sub-partitions (as the receiver) from the summed sample; my regions and my cursors in every owner's buffer
    CU(cudaMemcpyAsync(pl.hist, d_hist_mine, pl.n_sub * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->s_main));
    LAUNCH(k_part_plan_sums, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)sh.cap_keys, pl.chunk_sum);
    LAUNCH(k_part_plan, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)sh.cap_keys, pl.cfg.b2,
           pl.chunk_sum, (unsigned)pl.cap_bound, pl.beg, pl.cursor, pl.cap_end, pl.beg1, pl.cursor1, pl.end1, pl.scal);
    LAUNCH(k_shard_plan, 1, 1024, 0, c->s_main, d_hist_l1_all, sh.g, (unsigned)c->shard_rank, sh.b1, sh.stride,
           (unsigned)(std::min<uint64_t>(c->cap_buf1, 0xFFFFFFF0ull) & ~1ull), sh.reg_beg, sh.reg_end, sh.send_cur, sh.send_end, sh.d_blk);
    CU(cudaEventRecord(c->ev_p[1], c->s_main));
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    const unsigned n_regs = (unsigned)c->n_shards << sh.b1;
    // bins = (owner, level-1 bin).  Own keys land in the own level-1 buffer; the other owners' blocks are built
    // in the level-2 buffer (idle until the count) and pushed over NVLink by the pusher CTAs of the NEXT chunk's
    // launch, so the transfer of chunk i overlaps the extraction of chunk i+1.
    OkPushDesc pd{};
    pd.end = sh.send_end; pd.blk = sh.d_blk; pd.local = c->d_buf2; pd.n_regs = n_regs; pd.b1 = sh.b1; pd.me = (unsigned)c->shard_rank;
    for (int r = 0; r < c->n_shards; ++r) pd.peer[r] = sh.peer[r];
    unsigned n_chunks = 8;
    if (const char* ev = getenv("ORION_SHARD_CHUNKS")) n_chunks = (unsigned)std::max(1, atoi(ev));
    if (n_tiles < 64 * (uint64_t)n_chunks) n_chunks = 1;
    n_chunks = std::min(n_chunks, 8u);
    unsigned* snap = sh.d_snap;                                   // cursor snapshots: [0] = before chunk 0, [i+1] = after chunk i
    CU(cudaMemcpyAsync(snap, sh.send_cur, n_regs * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->s_main));
    if (n_tiles && n_records) {
        OkPeerOut po{}; po.shift = sh.b1;
        for (int r = 0; r < c->n_shards; ++r) po.p[r] = r == c->shard_rank ? c->d_buf1 : c->d_buf2;
        auto kern = c->norm_mode == OK_NORM_NORMALIZED ? OK_BY_K(c->k, k_part_scatter_bases, true, true) : OK_BY_K(c->k, k_part_scatter_bases, false, true);
        const size_t smem_push = sizeof(OkScatterSmem) + sizeof(OkPushSmem);      // staging buffers of the copy warp (TMA push)
        TRY(set_smem(kern, smem_push));
        { const char* ev = getenv("ORION_PUSH_TMA"); pd.tma = (ev ? atoi(ev) : g_push_tma) ? 1u : 0u; }   // read per batch: A/B runs
        const OkPartCfg cfg = shard_global_cfg(c, sh.b1);      // level-1 bin id = (owner, bin)
        const uint64_t per_chunk = (n_tiles + n_chunks - 1) / n_chunks;
        for (unsigned ch = 0; ch < n_chunks; ++ch) {
            const uint64_t t0 = ch * per_chunk, t1 = std::min<uint64_t>(n_tiles, t0 + per_chunk);
            if (t1 <= t0) { CU(cudaMemcpyAsync(snap + (size_t)(ch + 1) * 1024, sh.send_cur, n_regs * 4, cudaMemcpyDeviceToDevice, c->s_main)); continue; }
            pd.enabled = ch ? 1u : 0u;
            pd.prev = snap + (size_t)(ch ? ch - 1 : 0) * 1024; pd.cur = snap + (size_t)ch * 1024;
            const uint64_t max_warps = (uint64_t)grid_sm * (OK_SB_KPT == 16 ? 3 : 2) * OK_SB_WARPS;
            const uint64_t tpw = std::max<uint64_t>(1, (t1 - t0 + max_warps - 1) / max_warps);
            const unsigned blocks = (unsigned)((t1 - t0 + OK_SB_WARPS * tpw - 1) / (OK_SB_WARPS * tpw));
            LAUNCH(kern, blocks, OK_SB_THREADS + 32, smem_push, c->s_main, d_bases, n_bases, d_rec_offsets, n_records, t0, t1, tpw, c->k,
                   cfg, sh.send_cur, (const unsigned*)sh.send_end, (unsigned long long*)nullptr, (OkPartSpill{c->spill, c->d_stats}),
                   c->d_stats->route_counts, po, pd);
            CU(cudaMemcpyAsync(snap + (size_t)(ch + 1) * 1024, sh.send_cur, n_regs * 4, cudaMemcpyDeviceToDevice, c->s_main));
        }
    } else {
        for (unsigned ch = 0; ch < n_chunks; ++ch)
            CU(cudaMemcpyAsync(snap + (size_t)(ch + 1) * 1024, sh.send_cur, n_regs * 4, cudaMemcpyDeviceToDevice, c->s_main));
    }
    CU(cudaEventRecord(c->ev_a, c->s_main));
    // the last chunk's push is the only exposed one
    pd.enabled = 1u; pd.prev = snap + (size_t)(n_chunks - 1) * 1024; pd.cur = snap + (size_t)n_chunks * 1024;
    LAUNCH(k_shard_push, grid_sm * 4, 256, 0, c->s_main, pd);
    LAUNCH(k_shard_export, 1, 1024, 0, c->s_main, sh.send_cur, sh.g, (unsigned)c->shard_rank, sh.b1, sh.d_blk, d_cursors_out);
    CU(cudaEventRecord(c->ev_p[2], c->s_main));
    TRY(read_stats(c));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_scatter1, c->ev_p[1], c->ev_p[2]);
    cudaEventElapsedTime(&c->ms_push, c->ev_a, c->ev_p[2]);
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_p[0], c->ev_p[2]); c->ms_route += ms;
    if (c->h_stats->spill_n) {
        // a region overflowed: the spilled k-mers belong to OTHER ranks, this rank cannot count them
        CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, 8, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        c->h_stats->spill_n = 0;
        return set_err(OK_ERR_INTERNAL, "sharded scatter overflowed a sampled region; count this batch through the two-pass route instead");
    }
    return OK_SUCCESS;
}

// d_cursors_all[n_ranks][n_ranks << l1_bits]: the all-gathered d_cursors_out of every rank
OK_EXPORT int ok_shard_count_device(ok_counter* c, const uint32_t* d_cursors_all) {
    if (!c || !d_cursors_all) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_count_device: NULL argument");
    if (!c->pl.sharded || c->run_state != RUN_NONE) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_shard_count_device: no scattered batch pending");
    ShardState& sh = c->shard;
    PartPlan& pl = c->pl;
    const uint64_t windows_before = c->windows;
    LAUNCH(k_shard_fills, 1, 1024, 0, c->s_main, d_cursors_all, sh.g, (unsigned)c->shard_rank, sh.b1, sh.reg_beg, sh.reg_end,
           sh.reg_fill, sh.d_received);
    CU(cudaMemcpyAsync(&c->h_part->received, sh.d_received, 8, cudaMemcpyDeviceToHost, c->s_main));
    if (pl.cfg.b2 == 0) return set_err(OK_ERR_INTERNAL, "sharded path needs two scatter levels");
    TRY(part_finish(c, pl));
    c->windows = windows_before + c->h_part->received;
    c->batch_windows = c->h_part->received;
    CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (part_hint_misled(c, pl)) {
        // the keys came from the peers and are gone: this rank cannot recount alone.  The caller clears the
        // counter and routes the batch again (multi.py: every rank agrees through one all-reduce).
        TRY(part_discard(c, windows_before));
        c->distrust_hint = true; sh.ready = false;
        return set_err(OK_ERR_INTERNAL, "capacity hint too low for the sharded count (%llu sub-partitions outgrew their tables); "
                                        "recount the batch: the hint is ignored from now on", (unsigned long long)c->h_part->scal.n_deferred);
    }
    const int r = part_absorb_spills(c, windows_before);
    if (r == PART_RETRY) return set_err(OK_ERR_INTERNAL, "sharded count spilled beyond the spill list");
    return r;       // an earlier batch's run stays set aside until ok_counter_commit_batch (the ranks agree first)
}

// ---- multi-GPU exchange, third form: chunked scatter + one copy-engine peer copy per (peer, chunk) ----
// Per batch and rank:  ok_xchg_sample_device -> [reduce-scatter of the fine histogram, all-gather of the per-chunk
// level-1 histograms] -> ok_xchg_scatter_device (host plan; per chunk: extraction + multisplit by (owner, level-1
// bin) into the chunk's sub-blocks, then one plain async peer copy per owner on that owner's copy stream, under the
// next chunk's extraction; returns once every copy of this sender has landed) -> [any collective = the barrier]
// -> ok_xchg_count_device (per chunk: fills from the sub-block headers, level-2 scatter; then the count).
namespace {
constexpr unsigned XCHG_STRIDE = 1024;   // entries per chunk row of the d_xchg arrays
}  // namespace

OK_EXPORT int ok_xchg_geometry(ok_counter* c, uint64_t n_bases_max, uint32_t* sub_bits, uint32_t* l1_bits, uint32_t* n_chunks,
                               uint64_t* buffer_keys) {
    if (!n_chunks) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_geometry: NULL argument");
    TRY(ok_shard_geometry(c, n_bases_max, sub_bits, l1_bits, buffer_keys));
    ShardState& sh = c->shard;
    // measured at 8 GPUs (profiles/r2_*): two levels 33.0 ms per step, three levels 35.8 -- the extra pass over the
    // received keys costs more than the friendlier runs save; the three-level form stays selectable (ORION_XCHG_LEVELS=3)
    sh.three = false;
    if (const char* ev = getenv("ORION_XCHG_LEVELS")) sh.three = atoi(ev) == 3;
    sh.rb1 = sh.b1; sh.rb2 = sh.b2;
    if (sh.three) {
        // 2^g owners x 2^sub_bits sub-partitions are g + sub_bits bits of position: in two passes that is 512-1024 bins
        // per pass at 8 GPUs -- 64-byte runs, one DRAM page activation each, which starves the copy engines (measured:
        // 300 GB/s per GPU against 640 with idle SMs).  Three passes of <= 256 bins keep the runs at 256 bytes and more.
        PartPlan pl;
        part_choose_bits(c, n_bases_max, pl, /*use_hint=*/true);
        sh.hinted = pl.hinted;
        sh.rb1 = pl.cfg.b1; sh.rb2 = pl.cfg.b2; sh.sub_bits = pl.cfg.b1 + pl.cfg.b2; sh.b1 = 0; sh.b2 = sh.sub_bits;
        sh.cap_keys = (uint64_t)((double)n_bases_max * sh.margin) + 4096;        // raw keys; the per-region slack is added below
        *sub_bits = sh.sub_bits; *l1_bits = 0;
    }
    unsigned nc = 8;
    if (const char* ev = getenv("ORION_XCHG_CHUNKS")) nc = (unsigned)std::min(8, std::max(1, atoi(ev)));
    const uint64_t n_tiles = (n_bases_max + OK_TILE_BASES - 1) / OK_TILE_BASES;
    if (n_tiles < 64ull * nc) nc = 1;             // small batches: the per-chunk regions would be mostly slack
    sh.n_chunks = nc;
    // a sub-block carries a header and every region its own 6-sigma slack: n_chunks x senders x bins regions per owner
    const uint64_t regions = (uint64_t)nc << (sh.g + sh.b1);
    sh.cap_keys += regions * (200ull + 10ull * sh.stride) +
                   (uint64_t)(6.0 * std::sqrt((double)sh.stride) * std::sqrt((double)regions * (((double)n_bases_max * sh.margin) + (double)regions * sh.stride)));
    if (sh.cap_keys >= (1ull << 32)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_geometry: batch too large for 32-bit offsets");
    if (!sh.d_xchg) CU(cudaMalloc((void**)&sh.d_xchg, (6 * 8 * XCHG_STRIDE + 2 * 8 * 8) * sizeof(unsigned)));
    // with few peers one copy per (peer, chunk) leaves copy engines idle (measured at 2 GPUs: 500 GB/s on one stream):
    // the sub-block is cut into pieces on several streams
    sh.streams_per_peer = std::max(1u, std::min(4u, 8u / (unsigned)c->n_shards));
    if (const char* ev = getenv("ORION_XCHG_STREAMS")) sh.streams_per_peer = (unsigned)std::min(4, std::max(1, atoi(ev)));
    for (int r = 0; r < c->n_shards; ++r)
        for (unsigned q = 0; q < sh.streams_per_peer; ++q)
            if (!sh.s_peer[r][q] && r != c->shard_rank) {
                CU(cudaStreamCreateWithFlags(&sh.s_peer[r][q], cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&sh.ev_piece[r][q], cudaEventDisableTiming));
            }
    for (auto& e : sh.ev_chunk) if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : sh.ev_sent) if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!sh.s_recv) {
        CU(cudaStreamCreateWithFlags(&sh.s_recv, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&sh.s_join, cudaStreamNonBlocking));
        for (auto& e : sh.ev_recv) CU(cudaEventCreate(&e));
    }
    *n_chunks = nc; *buffer_keys = sh.cap_keys;
    return OK_SUCCESS;
}

// d_hist_fine[n_ranks << sub_bits] (owner, sub-partition); d_hist_l1c[n_chunks][n_ranks << l1_bits] (chunk, owner, level-1 bin)
OK_EXPORT int ok_xchg_sample_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                    uint64_t n_records, uint32_t* d_hist_fine, uint32_t* d_hist_l1c) {
    if (c) { TRY(part_settle(c)); TRY(run_stash(c)); }      // an earlier batch's run is set aside and merged after the count
    TRY(shard_check(c, "ok_xchg_sample_device", n_bases));
    if (!d_hist_fine || !d_hist_l1c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_sample_device: NULL histogram");
    if (n_bases && ((uintptr_t)d_bases & 15u)) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    ShardState& sh = c->shard;
    if (!sh.n_chunks) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_sample_device: ok_xchg_geometry first");
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    const unsigned n_regs = (unsigned)c->n_shards << sh.b1;
    CU(cudaEventRecord(c->ev_a, c->s_main));
    CU(cudaMemsetAsync(d_hist_fine, 0, ((size_t)c->n_shards << sh.sub_bits) * sizeof(unsigned), c->s_main));
    CU(cudaMemsetAsync(d_hist_l1c, 0, (size_t)sh.n_chunks * n_regs * sizeof(unsigned), c->s_main));
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    if (n_tiles && n_records) {
        const uint64_t sampled = (n_tiles + sh.stride - 1) / sh.stride;
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((sampled + 7) / 8, (uint64_t)grid_sm * 4));
        auto kern = c->norm_mode == OK_NORM_NORMALIZED ? k_xchg_sample<true> : k_xchg_sample<false>;
        const size_t smem = (size_t)sh.n_chunks * n_regs * sizeof(unsigned);
        TRY(set_smem(kern, smem));
        const uint64_t per_chunk = (n_tiles + sh.n_chunks - 1) / sh.n_chunks;
        LAUNCH(kern, blocks, 256, smem, c->s_main, d_bases, n_bases, d_rec_offsets, n_records, n_tiles, (uint64_t)sh.stride, c->k,
               shard_global_cfg(c, sh.sub_bits), d_hist_fine, 32u - (sh.g + sh.b1), n_regs, per_chunk, sh.n_chunks, d_hist_l1c);
    }
    CU(cudaEventRecord(c->ev_b, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_route, c->ev_a, c->ev_b);
    return OK_SUCCESS;
}

// d_hist_mine[1 << sub_bits]: the fine histogram summed over the ranks, this rank's slice (device);
// h_l1c_all[n_ranks][n_chunks][n_ranks << l1_bits]: every rank's per-chunk level-1 histogram (HOST memory).
namespace {
int xchg_begin(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
               uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* h_l1c_all) {
    TRY(shard_check(c, "ok_xchg_scatter_device", n_bases));
    if (!d_hist_mine || !h_l1c_all) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_scatter_device: NULL argument");
    ShardState& sh = c->shard;
    if (!sh.n_chunks) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_scatter_device: ok_xchg_geometry first");
    const unsigned W = (unsigned)c->n_shards, me = (unsigned)c->shard_rank, NC = sh.n_chunks, n_bin1 = 1u << sh.b1, n_regs = W << sh.b1;
    unsigned long long* const own = sh.peer[me];          // my peer-mapped buffer (== the level-1 buffer in the two-level form)
    PartPlan& pl = c->pl; pl = PartPlan{};
    pl.cfg.key_shift = 64 - 2 * c->k; pl.cfg.shard_log2 = sh.g; pl.cfg.b1 = sh.rb1; pl.cfg.b2 = sh.rb2;
    pl.n_sub = 1u << sh.sub_bits; pl.n_bin1 = 1u << sh.rb1; pl.stride = sh.stride; pl.sharded = !sh.three;
    const uint64_t n_units = (uint64_t)((double)sh.n_bases_max * sh.margin);
    pl.hinted = sh.hinted;
    pl.big_count = sh.hinted ? std::min<uint64_t>(c->user_hint, sh.n_bases_max) / pl.n_sub > 4600 : sh.n_bases_max / pl.n_sub > 5800;
    if (const char* ev = getenv("ORION_BIG_COUNT")) pl.big_count = atoi(ev) != 0;
    TRY(part_layout(c, n_units, OK_TILE_BASES, 0, pl));
    TRY(dev_reserve(&sh.d_send, &sh.cap_send, sh.cap_keys));
    // ---- the layout, identical on every rank: owner o's buffer = for every sender, for every chunk, one sub-block
    const unsigned hdr_keys = std::max(2u, (n_bin1 / 2u + 1u) & ~1u);
    const uint64_t limit = std::min<uint64_t>(sh.recv_cap, 0xFFFFFFF0ull) & ~1ull;
    std::vector<unsigned> cur((size_t)NC * XCHG_STRIDE, 0), end((size_t)NC * XCHG_STRIDE, 0), rbeg((size_t)NC * XCHG_STRIDE, 0),
                          rend((size_t)NC * XCHG_STRIDE, 0), hdr_send(NC * 8, 0), hdr_recv(NC * 8, 0);
    struct Copy { uint64_t src, dst, len; unsigned peer, chunk; };
    std::vector<Copy> copies;
    uint64_t local_run = 0;
    for (unsigned o = 0; o < W; ++o) {
        uint64_t run = 0;
        for (unsigned s2 = 0; s2 < W; ++s2)
            for (unsigned ch = 0; ch < NC; ++ch) {
                const uint32_t* h = h_l1c_all + ((size_t)s2 * NC + ch) * n_regs + (size_t)o * n_bin1;
                const uint64_t sub_start = std::min(run, limit);
                run += hdr_keys;
                const uint64_t local_base = local_run;                  // only meaningful for s2 == me, o != me
                for (unsigned b = 0; b < n_bin1; ++b) {
                    const unsigned cap = ok_part_capacity(h[b], sh.stride, limit);
                    const uint64_t lo = std::min(run, limit), hi = std::min(run + cap, limit);
                    if (s2 == me) {
                        const uint64_t shift = o == me ? 0 : local_base - sub_start;      // remote -> local coordinates (mod 2^64)
                        cur[(size_t)ch * XCHG_STRIDE + o * n_bin1 + b] = (unsigned)(lo + shift);
                        end[(size_t)ch * XCHG_STRIDE + o * n_bin1 + b] = (unsigned)(hi + shift);
                    }
                    if (o == me) { rbeg[(size_t)ch * XCHG_STRIDE + s2 * n_bin1 + b] = (unsigned)lo; rend[(size_t)ch * XCHG_STRIDE + s2 * n_bin1 + b] = (unsigned)hi; }
                    run += cap;
                }
                const uint64_t sub_len = std::min(run, limit) - sub_start;
                if (o == me) hdr_recv[ch * 8 + s2] = (unsigned)sub_start;
                if (s2 == me) {
                    hdr_send[ch * 8 + o] = (unsigned)(o == me ? sub_start : local_base);
                    if (o != me) { copies.push_back({local_base, sub_start, sub_len, o, ch}); local_run += sub_len; }
                }
            }
    }
    if (local_run > sh.cap_send) return set_err(OK_ERR_INTERNAL, "chunked exchange: the send buffer is too small (%llu > %llu keys); count this batch through the two-pass route instead",
                                                (unsigned long long)local_run, (unsigned long long)sh.cap_send);
    // (a region clipped by `limit` simply spills below and the batch falls back to the exact route)
    unsigned* dx = sh.d_xchg;
    unsigned *d_cur = dx, *d_end = dx + 8 * XCHG_STRIDE, *d_beg = dx + 16 * XCHG_STRIDE, *d_rbeg = dx + 24 * XCHG_STRIDE, *d_rend = dx + 32 * XCHG_STRIDE;
    unsigned *d_hs = dx + 48 * XCHG_STRIDE, *d_hr = d_hs + 64;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    CU(cudaEventRecord(c->ev_p[0], c->s_main));
    CU(cudaMemcpyAsync(d_cur, cur.data(), cur.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_beg, cur.data(), cur.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_end, end.data(), end.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_rbeg, rbeg.data(), rbeg.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_rend, rend.data(), rend.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_hs, hdr_send.data(), hdr_send.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaMemcpyAsync(d_hr, hdr_recv.data(), hdr_recv.size() * 4, cudaMemcpyHostToDevice, c->s_main));
    // my sub-partitions (as the receiver) from the summed sample
    CU(cudaMemcpyAsync(pl.hist, d_hist_mine, pl.n_sub * sizeof(unsigned), cudaMemcpyDeviceToDevice, c->s_main));
    LAUNCH(k_part_plan_sums, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)sh.cap_keys, pl.chunk_sum);
    LAUNCH(k_part_plan, (pl.n_sub + 1023) / 1024, 1024, 0, c->s_main, pl.hist, pl.n_sub, pl.stride, (unsigned)sh.cap_keys, pl.cfg.b2,
           pl.chunk_sum, (unsigned)pl.cap_bound, pl.beg, pl.cursor, pl.cap_end, pl.beg1, pl.cursor1, pl.end1, pl.scal);
    CU(cudaMemsetAsync(sh.d_received, 0, 8, c->s_main));
    CU(cudaEventRecord(c->ev_p[1], c->s_main));
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    const uint64_t per_chunk = (n_tiles + NC - 1) / NC;
    OkPeerOut po{}; po.shift = sh.b1;
    for (unsigned r = 0; r < W; ++r) po.p[r] = r == me ? own : sh.d_send;
    auto kern = c->norm_mode == OK_NORM_NORMALIZED ? OK_BY_K(c->k, k_part_scatter_bases, true, true) : OK_BY_K(c->k, k_part_scatter_bases, false, true);
    if (n_regs <= 256 && g_scatter_p3)      // >= 32 staging slots per bin: three fuller rounds per warp-tile
        kern = c->norm_mode == OK_NORM_NORMALIZED ? OK_BY_K3(c->k, k_part_scatter_bases, true, true) : OK_BY_K3(c->k, k_part_scatter_bases, false, true);
    TRY(set_smem(kern, sizeof(OkScatterSmem)));
    const OkPartCfg cfg = shard_global_cfg(c, sh.b1);      // scatter bin id = (owner, sender-side level-1 bin)
    size_t next_copy = 0;
    // Issue order matters: the copy engines take the copies roughly in the order they were enqueued, whatever their
    // streams.  With every sender starting at owner 0 all ranks hit the same destination at once and its ingress caps
    // the whole exchange (measured: 320 GB/s per GPU; staggered, tools/nvlink_a2a_mp reaches 700).  Sender s starts at
    // owner s+1: at any moment the copies in flight form a permutation.
    std::sort(copies.begin(), copies.end(), [&](const Copy& a, const Copy& b) {
        return a.chunk != b.chunk ? a.chunk < b.chunk : (a.peer + W - me) % W < (b.peer + W - me) % W; });
    for (unsigned ch = 0; ch < NC; ++ch) {
        const uint64_t t0 = ch * per_chunk, t1 = std::min<uint64_t>(n_tiles, t0 + per_chunk);
        if (t1 > t0 && n_records) {
            const uint64_t max_warps = (uint64_t)grid_sm * (OK_SB_KPT == 16 ? 3 : 2) * OK_SB_WARPS;
            const uint64_t tpw = std::max<uint64_t>(1, (t1 - t0 + max_warps - 1) / max_warps);
            const unsigned blocks = (unsigned)((t1 - t0 + OK_SB_WARPS * tpw - 1) / (OK_SB_WARPS * tpw));
            LAUNCH(kern, blocks, OK_SB_THREADS, sizeof(OkScatterSmem), c->s_main, d_bases, n_bases, d_rec_offsets, n_records, t0, t1, tpw, c->k,
                   cfg, d_cur + (size_t)ch * XCHG_STRIDE, (const unsigned*)(d_end + (size_t)ch * XCHG_STRIDE), (unsigned long long*)nullptr,
                   (OkPartSpill{c->spill, c->d_stats}), c->d_stats->route_counts, po, OkPushDesc{});
        }
        LAUNCH(k_xchg_headers, 1, 1024, 0, c->s_main, d_cur + (size_t)ch * XCHG_STRIDE, d_end + (size_t)ch * XCHG_STRIDE, d_beg + (size_t)ch * XCHG_STRIDE,
               n_regs, sh.b1, me, d_hs + ch * 8, own, sh.d_send);
        CU(cudaEventRecord(sh.ev_chunk[ch], c->s_main));
        // one plain asynchronous peer copy per (owner, chunk), each owner on its own stream: they run on the copy
        // engines under the extraction of the next chunk
        for (; next_copy < copies.size() && copies[next_copy].chunk == ch; ++next_copy) {
            const Copy& cp = copies[next_copy];
            const uint64_t piece = ((cp.len + sh.streams_per_peer - 1) / sh.streams_per_peer + 1) & ~1ull;      // even: 16-byte aligned cuts
            for (unsigned q = 0; q < sh.streams_per_peer; ++q) {
                const uint64_t a = std::min<uint64_t>(cp.len, q * piece), b2 = std::min<uint64_t>(cp.len, (q + 1) * piece);
                CU(cudaStreamWaitEvent(sh.s_peer[cp.peer][q], sh.ev_chunk[ch], 0));
                if (b2 > a) CU(cudaMemcpyAsync(sh.peer[cp.peer] + cp.dst + a, sh.d_send + cp.src + a, (b2 - a) * 8, cudaMemcpyDeviceToDevice, sh.s_peer[cp.peer][q]));
                CU(cudaEventRecord(sh.ev_piece[cp.peer][q], sh.s_peer[cp.peer][q]));
                CU(cudaStreamWaitEvent(sh.s_join, sh.ev_piece[cp.peer][q], 0));
            }
        }
        CU(cudaStreamWaitEvent(sh.s_join, sh.ev_chunk[ch], 0));
        CU(cudaEventRecord(sh.ev_sent[ch], sh.s_join));       // chunk ch of this sender has landed in every owner's buffer
    }
    CU(cudaEventRecord(c->ev_p[2], c->s_main));
    sh.xchg_chunks_used = NC; sh.recv_chunks = 0; sh.begun = true;
    return OK_SUCCESS;
}

int xchg_end(ok_counter* c) {
    ShardState& sh = c->shard;
    if (!sh.begun) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_scatter_end: no chunked scatter in progress");
    sh.begun = false;
    CU(cudaStreamWaitEvent(c->s_main, sh.ev_sent[sh.xchg_chunks_used - 1], 0));     // s_join is in order: the last chunk's event covers them all
    CU(cudaEventRecord(c->ev_b, c->s_main));
    TRY(read_stats(c));          // drains the compute stream, which has waited for every peer copy
    CU(cudaGetLastError());
    cudaEventElapsedTime(&c->ms_scatter1, c->ev_p[1], c->ev_p[2]);
    cudaEventElapsedTime(&c->ms_push, c->ev_p[2], c->ev_b);
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_p[0], c->ev_b); c->ms_route += ms;
    if (c->h_stats->spill_n) {
        // a region overflowed: the spilled k-mers belong to OTHER ranks, this rank cannot count them
        CU(cudaStreamSynchronize(sh.s_recv));
        CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, 8, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        c->h_stats->spill_n = 0;
        return set_err(OK_ERR_INTERNAL, "chunked exchange overflowed a sampled region; count this batch through the two-pass route instead");
    }
    sh.xchg_pending = true;
    return OK_SUCCESS;
}

// receive work of one chunk on the receive stream: fills from the sub-block headers, work items, level-2 scatter
int xchg_recv_chunk(ok_counter* c, unsigned ch, cudaStream_t st) {
    ShardState& sh = c->shard;
    PartPlan& pl = c->pl;
    const unsigned W = (unsigned)c->n_shards, n_regs = W << sh.b1;
    unsigned* dx = sh.d_xchg;
    unsigned *d_rbeg = dx + 24 * XCHG_STRIDE, *d_rend = dx + 32 * XCHG_STRIDE, *d_rfill = dx + 40 * XCHG_STRIDE, *d_hr = dx + 48 * XCHG_STRIDE + 64;
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    auto k_l2 = OK_BY_K(c->k, k_part_scatter_keys, 2, true);
    TRY(set_smem(k_l2, sizeof(OkScatterKeysSmem)));
    unsigned long long* const own = sh.peer[c->shard_rank];
    LAUNCH(k_xchg_fills, 1, 1024, 0, st, own, d_hr + ch * 8, n_regs, sh.b1, d_rbeg + (size_t)ch * XCHG_STRIDE,
           d_rend + (size_t)ch * XCHG_STRIDE, d_rfill + (size_t)ch * XCHG_STRIDE, sh.d_received);
    if (sh.three) {
        // what arrived is raw keys per sender: LEVEL 1 of the owner's own two levels, straight from the receive buffer
        const bool two = pl.cfg.b2 > 0;
        auto k_l1 = OK_BY_K(c->k, k_part_scatter_keys, 1, true);
        TRY(set_smem(k_l1, sizeof(OkScatterKeysSmem)));
        LAUNCH(k_part_items, 32, 1024, 0, st, d_rbeg + (size_t)ch * XCHG_STRIDE, d_rfill + (size_t)ch * XCHG_STRIDE, d_rend + (size_t)ch * XCHG_STRIDE,
               n_regs, 0xFFFFFFFFu, pl.item_off, pl.item_n, pl.item_bin, pl.scal, (unsigned*)nullptr);
        LAUNCH(k_l1, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), st, own, pl.item_off, pl.item_n,
               pl.item_bin, pl.scal, pl.cfg, two ? pl.cursor1 : pl.cursor, (const unsigned*)(two ? pl.end1 : pl.cap_end), two ? c->d_buf1 : c->d_buf2,
               (OkPartSpill{c->spill, c->d_stats}), (const unsigned*)nullptr, 0u, 0u);
        return OK_SUCCESS;
    }
    LAUNCH(k_part_items, 32, 1024, 0, st, d_rbeg + (size_t)ch * XCHG_STRIDE, d_rfill + (size_t)ch * XCHG_STRIDE, d_rend + (size_t)ch * XCHG_STRIDE,
           n_regs, pl.n_bin1 - 1u, pl.item_off, pl.item_n, pl.item_bin, pl.scal, (unsigned*)nullptr);
    LAUNCH(k_l2, grid_sm * 2, OK_SK_THREADS, sizeof(OkScatterKeysSmem), st, own, pl.item_off, pl.item_n,
           pl.item_bin, pl.scal, pl.cfg, pl.cursor, pl.cap_end, c->d_buf2, (OkPartSpill{c->spill, c->d_stats}), (const unsigned*)nullptr, 0u, 0u);
    return OK_SUCCESS;
}
}  // namespace

OK_EXPORT int ok_xchg_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                     uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* h_l1c_all) {
    TRY(xchg_begin(c, d_bases, n_bases, d_rec_offsets, n_records, d_hist_mine, h_l1c_all));
    return xchg_end(c);
}

// The same in steps, so that the owner's level-2 work overlaps the exchange: _begin enqueues the whole chunked scatter
// and its peer copies and returns at once; for every chunk the caller (1) waits with ok_xchg_chunk_sent until THIS
// sender's copies of the chunk have landed, (2) synchronises with the other ranks on the host (any barrier), then
// (3) calls ok_xchg_chunk_recv, which enqueues the receive work of that chunk -- every sender's sub-block of it is
// complete -- on a stream of its own; _end drains the sender side and reports overflowed regions.
// No device-side waiting on remote state: all cross-rank synchronisation stays with the caller's host code.
OK_EXPORT int ok_xchg_scatter_begin(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                    uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* h_l1c_all) {
    return xchg_begin(c, d_bases, n_bases, d_rec_offsets, n_records, d_hist_mine, h_l1c_all);
}
OK_EXPORT int ok_xchg_chunk_sent(ok_counter* c, uint32_t chunk) {
    if (!c || !c->shard.begun || chunk >= c->shard.xchg_chunks_used) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_chunk_sent: no such chunk in flight");
    CU(cudaEventSynchronize(c->shard.ev_sent[chunk]));
    return OK_SUCCESS;
}
OK_EXPORT int ok_xchg_chunk_recv(ok_counter* c, uint32_t chunk) {
    if (!c || !c->shard.begun || chunk != c->shard.recv_chunks || chunk >= c->shard.xchg_chunks_used)
        return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_chunk_recv: chunks are received in order, after ok_xchg_scatter_begin");
    ShardState& sh = c->shard;
    if (chunk == 0) {
        CU(cudaStreamWaitEvent(sh.s_recv, c->ev_p[1], 0));      // the plan (cursors of my sub-partitions) exists
        CU(cudaEventRecord(sh.ev_recv[0], sh.s_recv));
    }
    CU(cudaStreamWaitEvent(sh.s_recv, sh.ev_chunk[chunk], 0));  // my own sub-block of the chunk (written directly) and its header
    TRY(xchg_recv_chunk(c, chunk, sh.s_recv));
    ++sh.recv_chunks;
    if (sh.recv_chunks == sh.xchg_chunks_used) { CU(cudaEventRecord(sh.ev_recv[1], sh.s_recv)); CU(cudaEventRecord(sh.ev_recv[2], sh.s_recv)); }
    return OK_SUCCESS;
}
OK_EXPORT int ok_xchg_scatter_end(ok_counter* c) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_scatter_end: NULL handle");
    return xchg_end(c);
}

// every rank has returned from ok_xchg_scatter_device (the caller's collective in between is the barrier)
OK_EXPORT int ok_xchg_count_device(ok_counter* c) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_count_device: NULL handle");
    ShardState& sh = c->shard;
    if (!sh.xchg_pending || c->run_state != RUN_NONE) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_xchg_count_device: no exchanged batch pending");
    sh.xchg_pending = false;
    PartPlan& pl = c->pl;
    if (pl.cfg.b2 == 0 && !sh.three) return set_err(OK_ERR_INTERNAL, "sharded path needs two scatter levels");
    const uint64_t windows_before = c->windows;
    const unsigned NC = sh.xchg_chunks_used;
    const float ms_scatter1 = c->ms_scatter1;
    const bool overlapped = sh.recv_chunks == NC;
    if (sh.recv_chunks) CU(cudaStreamWaitEvent(c->s_main, overlapped ? sh.ev_recv[2] : sh.ev_chunk[0], 0));
    if (sh.recv_chunks && !overlapped) CU(cudaStreamSynchronize(sh.s_recv));      // (a caller that stopped half-way)
    CU(cudaEventRecord(c->ev_p[2], c->s_main));      // part_finish times the level-2 scatter from here
    for (unsigned ch = sh.recv_chunks; ch < NC; ++ch) TRY(xchg_recv_chunk(c, ch, c->s_main));     // not overlapped: everything after the barrier
    CU(cudaMemcpyAsync(&c->h_part->received, sh.d_received, 8, cudaMemcpyDeviceToHost, c->s_main));
    TRY(part_finish(c, pl, /*level2=*/sh.three));   // three-level form: the receive work was level 1, level 2 follows now
    c->ms_scatter1 = ms_scatter1;                     // measured by ok_xchg_scatter_device (part_finish re-read moved events)
    if (overlapped && !sh.three) cudaEventElapsedTime(&c->ms_scatter2, sh.ev_recv[0], sh.ev_recv[1]);     // ran under the exchange (includes its waits)
    sh.recv_chunks = 0;
    c->ms_insert = c->ms_sample + c->ms_scatter1 + c->ms_push + c->ms_scatter2 + c->ms_count;
    c->windows = windows_before + c->h_part->received;
    c->batch_windows = c->h_part->received;
    CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (part_hint_misled(c, pl)) {
        TRY(part_discard(c, windows_before));
        c->distrust_hint = true; sh.ready = false;
        return set_err(OK_ERR_INTERNAL, "capacity hint too low for the sharded count (%llu sub-partitions outgrew their tables); "
                                        "recount the batch: the hint is ignored from now on", (unsigned long long)c->h_part->scal.n_deferred);
    }
    const int r = part_absorb_spills(c, windows_before);
    if (r == PART_RETRY) return set_err(OK_ERR_INTERNAL, "sharded count spilled beyond the spill list");
    return r;       // an earlier batch's run stays set aside until ok_counter_commit_batch (the ranks agree first)
}

OK_EXPORT int ok_counter_finish_device(ok_counter* c, uint64_t min_count, const uint64_t** d_kmers,
                                       const uint64_t** d_counts, uint64_t* n) {
    if (!c || !n) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_finish_device: NULL argument");
    TRY(part_settle(c));
    uint64_t total = 0;
    const unsigned long long *dk = nullptr, *dc = nullptr;
    TRY(counter_result(c, min_count, &dk, &dc, &total));
    if (d_kmers) *d_kmers = (const uint64_t*)dk;
    if (d_counts) *d_counts = (const uint64_t*)dc;
    *n = total;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_finish(ok_counter* c, uint64_t min_count, uint64_t** kmers, uint64_t** counts,
                                uint64_t* n) {
    if (!c || !kmers || !counts || !n) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_finish: NULL argument");
    void *hk = nullptr, *hc = nullptr;
    if (c->run_state == RUN_LEVEL1) {
        if (min_count <= 1) {
            bool shipped = false;
            TRY(part_finish_sliced(c, kmers, counts, n, &shipped));
            if (shipped) return OK_SUCCESS;      // otherwise: counted, result not shipped -> the paths below
        } else {
            TRY(part_settle(c));
        }
    }
    if (c->run_state == RUN_SPARSE && min_count <= 1 && c->n_run) {
        // result pipeline: compact the sorted sub-partition runs slice by slice on the compute stream
        // while the copy stream ships the slices already compacted to the host
        const uint64_t total = c->n_run;
        const PartPlan& pl = c->pl;
        TRY(pool_alloc(&hk, total * 8));
        TRY(pool_alloc(&hc, total * 8));
        TRY(dev_reserve(&c->d_run_keys, &c->cap_run_keys, total));
        TRY(dev_reserve(&c->d_run_counts, &c->cap_run_counts, total));
        while (c->ev_chunks.size() < pl.n_slices) {
            cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_chunks.push_back(e);
        }
        CU(cudaEventRecord(c->ev_a, c->s_main));
        for (unsigned i = 0; i < pl.n_slices; ++i) {
            const unsigned p0 = i * pl.slice_step, p1 = std::min(pl.n_sub, p0 + pl.slice_step);
            const uint64_t o0 = c->h_part->slice_base[i], o1 = c->h_part->slice_base[i + 1];
            if (o1 <= o0) continue;
            launch_compact(c, p0, p1, c->d_run_keys, c->d_run_counts);
            CU(cudaEventRecord(c->ev_chunks[i], c->s_main));
            CU(cudaStreamWaitEvent(c->s_copy, c->ev_chunks[i], 0));
            CU(cudaMemcpyAsync((uint64_t*)hk + o0, c->d_run_keys + o0, (o1 - o0) * 8, cudaMemcpyDeviceToHost, c->s_copy));
            CU(cudaMemcpyAsync((uint64_t*)hc + o0, c->d_run_counts + o0, (o1 - o0) * 8, cudaMemcpyDeviceToHost, c->s_copy));
        }
        CU(cudaEventRecord(c->ev_b, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        CU(cudaStreamSynchronize(c->s_copy));
        CU(cudaGetLastError());
        cudaEventElapsedTime(&c->ms_compact, c->ev_a, c->ev_b);
        c->ms_readout = c->ms_compact;
        c->run_state = RUN_DENSE;
        *kmers = (uint64_t*)hk; *counts = (uint64_t*)hc; *n = total;
        return OK_SUCCESS;
    }
    uint64_t total = 0;
    const unsigned long long *dk = nullptr, *dc = nullptr;
    TRY(counter_result(c, min_count, &dk, &dc, &total));
    TRY(pool_alloc(&hk, total * 8));
    TRY(pool_alloc(&hc, total * 8));
    if (total) {
        CU(cudaMemcpyAsync(hk, dk, total * 8, cudaMemcpyDeviceToHost, c->s_main));
        CU(cudaMemcpyAsync(hc, dc, total * 8, cudaMemcpyDeviceToHost, c->s_copy));
        CU(cudaStreamSynchronize(c->s_main));
        CU(cudaStreamSynchronize(c->s_copy));
    }
    *kmers = (uint64_t*)hk; *counts = (uint64_t*)hc; *n = total;
    return OK_SUCCESS;
}

// multi-GPU: the exchange of the batch in progress failed on some rank (a sampled region overflowed, a hint proved
// too low) and every rank recounts it through another route.  Drops what this rank holds of THAT batch only;
// the result of earlier batches (set aside while the batch was being counted) stays.
OK_EXPORT int ok_counter_commit_batch(ok_counter* c) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_commit_batch: NULL handle");
    TRY(part_settle(c));
    return run_unstash(c);
}

OK_EXPORT int ok_counter_abort_batch(ok_counter* c) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_abort_batch: NULL handle");
    c->shard.xchg_pending = false;
    c->pl.sharded = false;
    if (c->shard.begun) { cudaStreamSynchronize(c->s_main); cudaStreamSynchronize(c->shard.s_join); c->shard.begun = false; }
    if (c->shard.s_recv) CU(cudaStreamSynchronize(c->shard.s_recv));
    c->shard.recv_chunks = 0;
    // Whatever run the counter holds now belongs to the aborted batch (the result of earlier batches was set aside
    // when the batch began and is only merged by ok_counter_commit_batch); its windows are taken back as well.
    CU(cudaStreamSynchronize(c->s_main));
    if (c->run_state != RUN_NONE) {
        uint64_t w = 0;
        if (c->run_state == RUN_SPARSE || c->run_state == RUN_DENSE) w = c->batch_windows;
        c->windows = c->windows >= w ? c->windows - w : 0;
        CU(cudaMemcpyAsync(&c->d_stats->windows, &c->windows, 8, cudaMemcpyHostToDevice, c->s_main));
        c->run_state = RUN_NONE; c->n_run = 0; c->occupied = 0;
    }
    CU(cudaMemsetAsync(&c->d_stats->spill_n, 0, 8, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    c->h_stats->spill_n = 0;
    return run_unstash(c);       // the set-aside run (if any) is the counter's result again
}

OK_EXPORT int ok_counter_set_capacity_hint(ok_counter* c, uint64_t capacity_hint) {
    if (!c) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_set_capacity_hint: NULL handle");
    c->user_hint = capacity_hint; c->distrust_hint = false;
    if (capacity_hint > c->hint) c->hint = capacity_hint;
    c->shard.ready = false;            // a sharded geometry derived from the old hint is void
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_set_path(ok_counter* c, int mode) {
    if (!c || mode < 0 || mode > 2) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_set_path: bad argument");
    c->path_mode = mode;
    return OK_SUCCESS;
}

OK_EXPORT int ok_counter_get_stats(ok_counter* c, ok_counter_stats* out) {
    if (c && out) TRY(part_settle(c));
    if (!c || !out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_counter_get_stats: NULL argument");
    out->n_slots = c->run_state != RUN_NONE ? 0 : c->tv.n_total; out->n_distinct = c->occupied; out->n_windows = c->windows;
    out->n_bases = c->bases_seen; out->max_displacement = c->max_disp; out->n_spilled = c->spilled_total;
    out->n_grows = c->grows; out->ms_insert = c->ms_insert; out->ms_readout = c->ms_readout; out->ms_fill = c->ms_fill; out->ms_route = c->ms_route;
    out->ms_sample = c->ms_sample; out->ms_scatter1 = c->ms_scatter1; out->ms_scatter2 = c->ms_scatter2;
    out->ms_count = c->ms_count; out->ms_compact = c->ms_compact; out->partitioned = c->run_state != RUN_NONE ? 1 : 0;
    out->ms_push = c->ms_push; out->n_deferred = c->n_deferred;
    out->ms_merge = c->ms_merge; out->n_merges = c->n_merges;
    return OK_SUCCESS;
}

// =================================================================================== sets ==
// Set builders are counters, and a counter owns streams, page-locked mirrors, a spill list and grow-only device
// buffers: creating and destroying one per reference genome cost 130 ms per 5 Mbp genome, a hundred times the
// count itself (tools/bench_build_query.py).  Sealed sets hand their (cleared) builder back to a small pool.
namespace {
constexpr size_t MAX_SPARE_BUILDERS = 4;

int take_builder(uint8_t k, int norm_mode, uint64_t capacity_hint, ok_counter** out) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_spare_builders.empty() && k >= 1 && k <= 32 && (norm_mode == OK_NORM_NORMALIZED || norm_mode == OK_NORM_RAW)) {
            ok_counter* c = g_spare_builders.back();
            g_spare_builders.pop_back();
            if (c->tv.slots && c->tv.key_shift != 64u - 2u * k) {
                // the pooled table was laid out for another k: its home slots (key << key_shift) would no longer be
                // monotone in the new keys and every sorted consumer of the set would be handed an unsorted array
                cudaFree(c->tv.slots); c->tv = OkTableView{}; c->grows = 0; c->occupied = 0;
            }
            c->k = k; c->norm_mode = norm_mode; c->hint = capacity_hint; c->user_hint = capacity_hint;
            c->distrust_hint = false; c->path_mode = 0;
            *out = c;
            return OK_SUCCESS;
        }
    }
    return ok_counter_create(k, norm_mode, capacity_hint, out);   // build.rs:83-85 validates k the same way
}

uint64_t builder_footprint(const ok_counter* c) {
    return (c->cap_buf1 + c->cap_buf2 + c->cap_run_keys + c->cap_run_counts + c->cap_acc_keys + c->cap_mrg_keys) * 8 + c->cap_bases;
}
constexpr uint64_t BIG_BUILDER = 4ull << 30;
std::atomic<int> g_keep_big_builders{0};       // > 0 while a union in groups is running: its multi-GB builder serves every group

void give_builder(ok_counter* c) {
    if (!c) return;
    const uint64_t footprint = builder_footprint(c);
    if (c->n_shards == 1 && !c->buf1_external && (footprint < BIG_BUILDER || g_keep_big_builders.load() > 0) && ok_counter_clear(c) == OK_SUCCESS) {      // multi-GB scratch goes back to the device
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_spare_builders.size() < MAX_SPARE_BUILDERS || footprint >= BIG_BUILDER) { g_spare_builders.push_back(c); return; }
    }
    TraceClock tc;
    ok_counter_destroy(c);
    tc.lap("give_builder: destroy (%.1f GB)", footprint * 1e-9);
}

// a union in groups (ok_set_union): the builder that counted one group, tens of GB of scratch, is kept for the next
// one (destroying and regrowing it cost 30 - 270 ms per group); whatever multi-GB builder is pooled when the union
// ends goes back to the device
struct KeepBigBuilders {
    KeepBigBuilders() { g_keep_big_builders.fetch_add(1); }
    ~KeepBigBuilders() {
        if (g_keep_big_builders.fetch_sub(1) != 1) return;
        std::vector<ok_counter*> big;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            for (size_t i = 0; i < g_spare_builders.size();)
                if (builder_footprint(g_spare_builders[i]) >= BIG_BUILDER) { big.push_back(g_spare_builders[i]); g_spare_builders.erase(g_spare_builders.begin() + i); }
                else ++i;
        }
        for (ok_counter* c : big) ok_counter_destroy(c);
    }
};
}  // namespace

struct ok_set {
    unsigned k = 0;
    int norm_mode = 0;
    ok_counter* builder = nullptr;            // while batches are still being added
    unsigned long long* d_keys = nullptr;     // sorted, duplicate-free, once sealed
    KeySlab* slab = nullptr;                  // the slab d_keys lives in (nullptr: an allocation of its own)
    uint64_t n = 0;
    bool sealed = false;
    int has_max = 0;                          // contains 0xFFFF...F (only possible for foreign k=32 sets)
    unsigned long long* d_table = nullptr;    // hashed membership table, built on first probe
    uint64_t n_table = 0;
    cudaStream_t st = nullptr;
    // staging of ok_probe_reads, kept across calls (three cudaMalloc / cudaFree per call cost more than a small probe)
    uint8_t* d_pb = nullptr; uint64_t cap_pb = 0;
    uint64_t* d_po = nullptr; uint64_t cap_po = 0;
    unsigned* d_ph = nullptr; uint64_t cap_ph = 0;
    // scratch of the probe by merge (probe_reads_merge), kept across calls as well
    unsigned long long* d_mlo = nullptr; uint64_t cap_mlo = 0;        // tile bounds in the set + the match counter
    unsigned long long* d_mkeys = nullptr; uint64_t cap_mkeys = 0;    // the batch's distinct k-mers found in the set
    unsigned long long* d_mtab = nullptr; uint64_t cap_mtab = 0;      // hashed table of those
};

namespace {

int set_seal(ok_set* s) {
    if (s->sealed) return OK_SUCCESS;
    if (!s->st) CU(set_stream(&s->st));
    uint64_t n = 0;
    if (s->builder) {
        TraceClock tc;
        const uint64_t *dk = nullptr, *dc = nullptr;
        TRY(ok_counter_finish_device(s->builder, 1, &dk, &dc, &n));
        tc.lap("seal: finish_device %llu keys", (unsigned long long)n);
        const uint64_t total = n + (s->has_max ? 1 : 0);
        if (total) {
            CU(keys_alloc(total, &s->d_keys, &s->slab));
            // on the set's own stream and drained here: every consumer runs on non-blocking streams, which the legacy
            // default stream does not order against, and the builder's buffer (dk) goes back to the pool right below
            if (n) CU(cudaMemcpyAsync(s->d_keys, dk, n * 8, cudaMemcpyDeviceToDevice, s->st));
            const unsigned long long m = OK_EMPTY_KEY;
            if (s->has_max) CU(cudaMemcpyAsync(s->d_keys + n, &m, 8, cudaMemcpyHostToDevice, s->st));
            CU(cudaStreamSynchronize(s->st));
        }
        n = total;
        tc.lap("seal: malloc + copy");
        give_builder(s->builder);
        s->builder = nullptr;
        tc.lap("seal: give_builder");
    }
    s->n = n;
    s->sealed = true;
    return OK_SUCCESS;
}

int set_table(ok_set* s) {
    TRY(set_seal(s));
    if (s->d_table) return OK_SUCCESS;
    s->n_table = std::max<uint64_t>(1024, 2 * s->n);
    CU(cudaMalloc((void**)&s->d_table, s->n_table * 8));
    LAUNCH(k_fill_u64, grid_for(s->n_table), 256, 0, s->st, s->d_table, s->n_table, OK_EMPTY_KEY);
    if (s->n) LAUNCH(k_keytable_build, grid_for(s->n), 256, 0, s->st, s->d_table, s->n_table, s->d_keys, s->n);
    CU(cudaStreamSynchronize(s->st));
    CU(cudaGetLastError());
    return OK_SUCCESS;
}

int kmer_size_mismatch(unsigned a, unsigned b) {  // errors.rs:24-25
    return set_err(OK_ERR_KMER_SIZE_MISMATCH,
                   "K-mer databases have incompatible k-mer sizes (overall comparison): %u vs %u", a, b);
}

}  // namespace

OK_EXPORT int ok_set_create(uint8_t k, int norm_mode, uint64_t capacity_hint, ok_set** out) {
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_create: out is NULL");
    *out = nullptr;
    ok_counter* b = nullptr;
    TRY(take_builder(k, norm_mode, capacity_hint, &b));
    ok_set* s = new ok_set();
    s->k = k; s->norm_mode = norm_mode; s->builder = b;
    *out = s;
    return OK_SUCCESS;
}

OK_EXPORT int ok_set_add_batch(ok_set* s, const uint8_t* bases, const uint64_t* rec_offsets, uint64_t n_records) {
    if (!s) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_add_batch: NULL handle");
    if (s->sealed || !s->builder) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_add_batch: the set is already sealed");
    return ok_counter_add_batch(s->builder, bases, rec_offsets, n_records);
}

OK_EXPORT int ok_set_add_batch_device(ok_set* s, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets, uint64_t n_records) {
    if (!s) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_add_batch_device: NULL handle");
    if (s->sealed || !s->builder) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_add_batch_device: the set is already sealed");
    TraceClock tc;
    const int r = ok_counter_add_batch_device(s->builder, d_bases, n_bases, d_rec_offsets, n_records);
    tc.lap("set add_batch_device %llu bases: sample %.3f l1 %.3f l2 %.3f count %.3f (device ms)", (unsigned long long)n_bases, s->builder->ms_sample,
           s->builder->ms_scatter1, s->builder->ms_scatter2, s->builder->ms_count);
    return r;
}

// build.rs:93-116 over many files: independent units, a few host threads each with its own builder (see the header).
// n_bases == nullptr: host batches (every worker copies its own file in, under the kernels of the others).
namespace {
int sets_build_many(uint8_t k, int norm_mode, uint64_t n_files, const uint8_t* const* d_bases, const uint64_t* n_bases,
                    const uint64_t* const* d_rec_offsets, const uint64_t* n_records, ok_set** out) {
    const bool on_device = n_bases != nullptr;
    if (n_files && (!d_bases || !d_rec_offsets || !n_records || !out))
        return set_err(OK_ERR_INVALID_ARGUMENT, "ok_sets_build_many: NULL argument");
    for (uint64_t i = 0; i < n_files; ++i) out[i] = nullptr;
    if (k == 0 || k > 32) return invalid_k(k);          // build.rs:83-85
    if (norm_mode != OK_NORM_NORMALIZED && norm_mode != OK_NORM_RAW) return set_err(OK_ERR_INVALID_ARGUMENT, "unknown norm_mode %d", norm_mode);
    if (n_files == 0) return OK_SUCCESS;
    TRY(ensure_init());
    unsigned n_threads = 2;          // measured, 200 genomes of 5 Mbp: 0.43 / 0.32 / 0.61 / 2.04 ms per genome with 1 / 2 / 4 / 8 threads
    if (const char* ev = getenv("ORION_BUILD_THREADS")) n_threads = (unsigned)std::min(16, std::max(1, atoi(ev)));
    n_threads = (unsigned)std::min<uint64_t>(n_threads, n_files);
    std::atomic<uint64_t> next{0};
    std::atomic<int> first_error{OK_SUCCESS};
    std::mutex err_mu;
    std::string err_text;
    auto worker = [&] {
        cudaSetDevice(g_device);                         // the current device is per-thread state
        for (;;) {
            const uint64_t i = next.fetch_add(1);
            if (i >= n_files || first_error.load() != OK_SUCCESS) return;
            ok_set* s = nullptr;
            int r = ok_set_create(k, norm_mode, 0, &s);
            if (r == OK_SUCCESS && n_records[i]) {
                if (on_device) { if (n_bases[i]) r = ok_set_add_batch_device(s, d_bases[i], n_bases[i], d_rec_offsets[i], n_records[i]); }
                else r = ok_set_add_batch(s, d_bases[i], d_rec_offsets[i], n_records[i]);
            }
            if (r == OK_SUCCESS) r = set_seal(s);
            if (r != OK_SUCCESS) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (first_error.load() == OK_SUCCESS) { first_error.store(r); err_text = g_err; }      // g_err is this thread's own
                ok_set_destroy(s);
                return;
            }
            out[i] = s;
        }
    };
    TraceClock tc;
    if (n_threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_threads; ++t) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    tc.lap("build_many: %llu files, %u threads", (unsigned long long)n_files, n_threads);
    if (first_error.load() != OK_SUCCESS) {
        for (uint64_t i = 0; i < n_files; ++i) { ok_set_destroy(out[i]); out[i] = nullptr; }
        g_err = err_text;
        return first_error.load();
    }
    return OK_SUCCESS;
}
}  // namespace

OK_EXPORT int ok_sets_build_many_device(uint8_t k, int norm_mode, uint64_t n_files, const uint8_t* const* d_bases,
                                        const uint64_t* n_bases, const uint64_t* const* d_rec_offsets, const uint64_t* n_records,
                                        ok_set** out) {
    if (n_files && !n_bases) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_sets_build_many_device: NULL n_bases");
    static const uint64_t none = 0;
    return sets_build_many(k, norm_mode, n_files, d_bases, n_files ? n_bases : &none, d_rec_offsets, n_records, out);
}
OK_EXPORT int ok_sets_build_many(uint8_t k, int norm_mode, uint64_t n_files, const uint8_t* const* bases,
                                 const uint64_t* const* rec_offsets, const uint64_t* n_records, ok_set** out) {
    return sets_build_many(k, norm_mode, n_files, bases, nullptr, rec_offsets, n_records, out);
}

OK_EXPORT int ok_set_from_sorted(uint8_t k, const uint64_t* kmers, uint64_t n, ok_set** out) {
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_from_sorted: out is NULL");
    *out = nullptr;
    if (k == 0 || k > 32) return invalid_k(k);
    if (n && !kmers) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL kmers");
    TRY(ensure_init());
    // (the order is checked on the device, after the upload: a host loop over the 1.28 G keys of 256 genome sets took
    // longer than their transfer)
    ok_set* s = new ok_set();
    s->k = k; s->n = n; s->sealed = true;
    s->has_max = (n && kmers[n - 1] == OK_EMPTY_KEY) ? 1 : 0;
    unsigned* d_bad = nullptr; unsigned bad = 0;
    cudaError_t e = set_stream(&s->st);
    if (e == cudaSuccess && n > 1) e = cudaMalloc((void**)&d_bad, 4);
    if (e == cudaSuccess && n > 1) e = cudaMemsetAsync(d_bad, 0, 4, s->st);
    if (e == cudaSuccess && n) e = keys_alloc(n, &s->d_keys, &s->slab);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(s->d_keys, kmers, n * 8, cudaMemcpyHostToDevice, s->st);
    if (e == cudaSuccess && n > 1) {
        LAUNCH(k_check_ascending, grid_for(n), 256, 0, s->st, s->d_keys, n, d_bad);
        e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s->st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);          // pageable source: the DMA has landed before anyone reads the set
    cudaFree(d_bad);
    if (e != cudaSuccess) { ok_set_destroy(s); return set_err(OK_ERR_CUDA, "CUDA error %s in ok_set_from_sorted", cudaGetErrorName(e)); }
    if (bad) {
        ok_set_destroy(s);
        uint64_t i = 1;
        while (i < n && kmers[i] > kmers[i - 1]) ++i;
        return set_err(OK_ERR_INVALID_ARGUMENT, "kmers must be strictly ascending (index %llu)", (unsigned long long)i);
    }
    *out = s;
    return OK_SUCCESS;
}

// a set from a sorted, duplicate-free DEVICE array (a key-range slice received from a peer, multi-GPU set algebra)
OK_EXPORT int ok_set_from_sorted_device(uint8_t k, const uint64_t* d_kmers, uint64_t n, ok_set** out) {
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_from_sorted_device: out is NULL");
    *out = nullptr;
    if (k == 0 || k > 32) return invalid_k(k);
    if (n && !d_kmers) return set_err(OK_ERR_INVALID_ARGUMENT, "NULL d_kmers");
    TRY(ensure_init());
    ok_set* s = new ok_set();
    s->k = k; s->n = n; s->sealed = true;
    unsigned* d_bad = nullptr; unsigned bad = 0; unsigned long long last = 0;
    cudaError_t e = set_stream(&s->st);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_bad, 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, s->st);
    if (e == cudaSuccess && n) e = keys_alloc(n, &s->d_keys, &s->slab);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(s->d_keys, d_kmers, n * 8, cudaMemcpyDeviceToDevice, s->st);
    if (e == cudaSuccess && n > 1) LAUNCH(k_check_ascending, grid_for(n), 256, 0, s->st, s->d_keys, n, d_bad);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(&last, s->d_keys + n - 1, 8, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    cudaFree(d_bad);
    if (e != cudaSuccess) { ok_set_destroy(s); return set_err(OK_ERR_CUDA, "CUDA error %s in ok_set_from_sorted_device", cudaGetErrorName(e)); }
    if (bad) { ok_set_destroy(s); return set_err(OK_ERR_INVALID_ARGUMENT, "kmers must be strictly ascending"); }
    s->has_max = (n && last == OK_EMPTY_KEY) ? 1 : 0;
    *out = s;
    return OK_SUCCESS;
}

// the sealed set's sorted keys in device memory (valid until the set is destroyed)
OK_EXPORT int ok_set_keys_device(ok_set* s, const uint64_t** d_kmers, uint64_t* n) {
    if (!s || !d_kmers || !n) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_keys_device: NULL argument");
    TRY(set_seal(s));
    *d_kmers = (const uint64_t*)s->d_keys; *n = s->n;
    return OK_SUCCESS;
}

OK_EXPORT int ok_set_copy_keys_device(ok_set* s, uint64_t first, uint64_t n, uint64_t* d_out) {
    if (!s || (n && !d_out)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_copy_keys_device: NULL argument");
    TRY(set_seal(s));
    if (first > s->n || n > s->n - first) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_copy_keys_device: range outside the set");
    if (n) { CU(cudaMemcpyAsync(d_out, s->d_keys + first, n * 8, cudaMemcpyDeviceToDevice, s->st)); CU(cudaStreamSynchronize(s->st)); }
    return OK_SUCCESS;
}

// multi-GPU set algebra: where the key ranges of n_ranks owners begin inside this set (the same ownership rule as
// the sharded count: equal shares of the canonical k-mer position).  bounds[r] .. bounds[r+1] = the keys rank r owns.
OK_EXPORT int ok_set_shard_bounds(ok_set* s, int n_ranks, uint64_t* bounds) {
    if (!s || !bounds) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_shard_bounds: NULL argument");
    if (n_ranks < 1 || n_ranks > 1024) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_shard_bounds: n_ranks out of range");
    TRY(set_seal(s));
    unsigned long long* d = nullptr;
    CU(cudaMalloc((void**)&d, (n_ranks + 1) * 8));
    LAUNCH(k_set_shard_bounds, 1, 1024, 0, s->st, s->d_keys, s->n - (s->has_max ? 1 : 0), 64u - 2u * s->k, (unsigned)n_ranks, d);
    cudaError_t e = cudaMemcpyAsync(bounds, d, (n_ranks + 1) * 8, cudaMemcpyDeviceToHost, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    cudaFree(d);
    if (e != cudaSuccess) return set_err(OK_ERR_CUDA, "CUDA error %s in ok_set_shard_bounds", cudaGetErrorName(e));
    bounds[n_ranks] = s->n;          // a foreign u64::MAX key (k = 32) belongs to the last owner
    return OK_SUCCESS;
}

OK_EXPORT int ok_set_size(ok_set* s, uint64_t* n) {
    if (!s || !n) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_size: NULL argument");
    TRY(set_seal(s));
    *n = s->n;
    return OK_SUCCESS;
}
OK_EXPORT int ok_set_k(ok_set* s, uint8_t* k) {
    if (!s || !k) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_k: NULL argument");
    *k = (uint8_t)s->k;
    return OK_SUCCESS;
}

OK_EXPORT int ok_set_export(ok_set* s, uint64_t** kmers, uint64_t* n) {
    if (!s || !kmers || !n) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_export: NULL argument");
    TRY(set_seal(s));
    void* h = nullptr;
    TRY(pool_alloc(&h, s->n * 8));
    if (s->n) { CU(cudaMemcpyAsync(h, s->d_keys, s->n * 8, cudaMemcpyDeviceToHost, s->st)); CU(cudaStreamSynchronize(s->st)); }
    *kmers = (uint64_t*)h; *n = s->n;
    return OK_SUCCESS;
}

namespace {
// union of two sealed sets by the keys-only merge (merge.cuh); -> a new sealed set
int set_merge_keys(const ok_set* a, const ok_set* b, ok_set** out) {
    *out = nullptr;
    ok_set* u = new ok_set();
    u->k = a->k; u->norm_mode = a->norm_mode; u->sealed = true; u->has_max = a->has_max | b->has_max;
    ulonglong2* d_split = nullptr; unsigned long long* d_tiles = nullptr;
    auto fail = [&](int code) { cudaFree(d_split); cudaFree(d_tiles); ok_set_destroy(u); return code; };
#define CUM(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(set_err(e_ == cudaErrorMemoryAllocation ? OK_ERR_OUT_OF_MEMORY : OK_ERR_CUDA, "CUDA error %s while merging two sets", cudaGetErrorName(e_))); } while (0)
    CUM(set_stream(&u->st));
    const uint64_t na = a->n, nb = b->n, n_tiles = (na + nb + OK_MG_TILE - 1) / OK_MG_TILE;
    CUM(cudaMalloc((void**)&d_split, (n_tiles + 1) * sizeof(ulonglong2)));
    CUM(cudaMalloc((void**)&d_tiles, (n_tiles + 1) * 8));
    if (set_smem(k_merge_count, sizeof(OkMergeSmem)) != OK_SUCCESS || set_smem(k_merge_write<false>, sizeof(OkMergeSmem)) != OK_SUCCESS) return fail(OK_ERR_CUDA);
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)grid_sm * 4));
    LAUNCH(k_merge_partition, grid_for(n_tiles + 1), 256, 0, u->st, a->d_keys, na, b->d_keys, nb, n_tiles, d_split);
    LAUNCH(k_merge_count, blocks, OK_MG_THREADS, sizeof(OkMergeSmem), u->st, a->d_keys, b->d_keys, d_split, n_tiles, d_tiles);
    LAUNCH(k_scan_tiles, 1, 1024, 0, u->st, d_tiles, n_tiles, d_tiles + n_tiles);
    unsigned long long total = 0;
    CUM(cudaMemcpyAsync(&total, d_tiles + n_tiles, 8, cudaMemcpyDeviceToHost, u->st));
    CUM(cudaStreamSynchronize(u->st));
    if (total) CUM(cudaMalloc((void**)&u->d_keys, total * 8));
    LAUNCH(k_merge_write<false>, blocks, OK_MG_THREADS, sizeof(OkMergeSmem), u->st, a->d_keys, (const unsigned long long*)nullptr, b->d_keys,
           (const unsigned long long*)nullptr, d_split, n_tiles, d_tiles, u->d_keys, (unsigned long long*)nullptr);
    CUM(cudaStreamSynchronize(u->st));
    CUM(cudaGetLastError());
#undef CUM
    cudaFree(d_split); cudaFree(d_tiles);
    u->n = total;
    *out = u;
    return OK_SUCCESS;
}
constexpr uint64_t UNION_GROUP_KEYS = 1200ull << 20;      // keys one pass of the partitioned path takes (32-bit offsets)
}  // namespace

OK_EXPORT int ok_set_union(ok_set* const* sets, uint64_t n_sets, ok_set** out) {
    if (!out || (n_sets && !sets)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_union: NULL argument");
    *out = nullptr;
    if (n_sets == 0) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_union: no sets");
    uint64_t sum = 0; int has_max = 0;
    for (uint64_t i = 0; i < n_sets; ++i) {
        if (!sets[i]) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_union: NULL set");
        if (sets[i]->k != sets[0]->k) return kmer_size_mismatch(sets[0]->k, sets[i]->k);
        TRY(set_seal(sets[i]));
        sum += sets[i]->n; has_max |= sets[i]->has_max;
    }
    uint64_t group_keys = UNION_GROUP_KEYS;
    if (const char* ev = getenv("ORION_UNION_GROUP_KEYS")) { const long long v = atoll(ev); if (v > 0) group_keys = (uint64_t)v; }     // test hook
    if (sum > group_keys && n_sets > 1) {
        // more keys than one pass takes (1,000 genomes are 5e9): groups of sets, each group one pass, and the group
        // results folded together with the keys-only merge -- every key is read and written once per fold step
        KeepBigBuilders keep;
        ok_set* acc = nullptr;
        uint64_t i = 0;
        while (i < n_sets) {
            uint64_t j = i, keys = 0;
            while (j < n_sets && (j == i || keys + sets[j]->n <= group_keys)) keys += sets[j++]->n;
            ok_set* g = nullptr;
            int r;
            TraceClock tc;
            if (j - i == 1) {      // one (large) set: it is its own union
                r = ok_set_from_sorted_device((uint8_t)sets[i]->k, (const uint64_t*)sets[i]->d_keys, sets[i]->n, &g);
                if (r == OK_SUCCESS) g->norm_mode = sets[i]->norm_mode;
            } else {
                r = ok_set_union(sets + i, j - i, &g);
                if (r == OK_SUCCESS) r = set_seal(g);
            }
            if (r != OK_SUCCESS) { ok_set_destroy(g); ok_set_destroy(acc); return r; }
            tc.lap("union: group of %llu sets, %llu keys -> %llu", (unsigned long long)(j - i), (unsigned long long)keys, (unsigned long long)g->n);
            if (!acc) acc = g;
            else {
                ok_set* m = nullptr;
                r = set_merge_keys(acc, g, &m);
                tc.lap("union: merge %llu + %llu", (unsigned long long)acc->n, (unsigned long long)g->n);
                ok_set_destroy(acc); ok_set_destroy(g);
                tc.lap("union: destroy merged inputs");
                if (r != OK_SUCCESS) return r;
                acc = m;
            }
            i = j;
        }
        *out = acc;
        return OK_SUCCESS;
    }
    ok_set* u = nullptr;
    TRY(ok_set_create((uint8_t)sets[0]->k, sets[0]->norm_mode, std::max<uint64_t>(sum, 1), &u));
    u->has_max = has_max;
    // All keys in one array -> ONE pass of the partitioned path (scatter by key range, dedupe in shared memory): the
    // union comes out sorted.  Set by set, the first set would take that path and every later one random-access
    // inserts into a device-wide table sized for the sum (db_types.rs:43-48 re-hashes every key the same way).
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_sets; ++i) total += sets[i]->n - (sets[i]->has_max ? 1 : 0);
    if (n_sets > 1 && total >= PART_MIN_BASES && total < (1ull << 31) && !getenv("ORION_UNION_SETWISE")) {
        TraceClock tc;
        unsigned long long* d_all = nullptr;
        cudaStream_t st = u->builder->s_main;
        cudaError_t e = cudaMalloc((void**)&d_all, total * 8);
        tc.lap("union: malloc d_all %.1f GB", total * 8e-9);
        uint64_t at = 0;
        for (uint64_t i = 0; i < n_sets && e == cudaSuccess; ++i) {
            const uint64_t m = sets[i]->n - (sets[i]->has_max ? 1 : 0);
            if (m) e = cudaMemcpyAsync(d_all + at, sets[i]->d_keys, m * 8, cudaMemcpyDeviceToDevice, st);
            at += m;
        }
        if (e != cudaSuccess) { cudaFree(d_all); ok_set_destroy(u); return set_err(OK_ERR_CUDA, "CUDA error %s in ok_set_union", cudaGetErrorName(e)); }
        tc.lap("union: concatenate");
        u->builder->keys_sorted_runs = true;
        const int r = ok_counter_add_kmers_device(u->builder, (const uint64_t*)d_all, total);   // returns with the stream drained
        u->builder->keys_sorted_runs = false;
        cudaStreamSynchronize(st);
        tc.lap("union: add_kmers_device (incl. its buffers)");
        cudaFree(d_all);
        tc.lap("union: free d_all");
        if (r != OK_SUCCESS) { ok_set_destroy(u); return r; }
        *out = u;
        return OK_SUCCESS;
    }
    for (uint64_t i = 0; i < n_sets; ++i) {
        const uint64_t m = sets[i]->n - (sets[i]->has_max ? 1 : 0);
        int r = ok_counter_add_kmers_device(u->builder, (const uint64_t*)sets[i]->d_keys, m);
        if (r != OK_SUCCESS) { ok_set_destroy(u); return r; }
    }
    *out = u;
    return OK_SUCCESS;
}

OK_EXPORT int ok_set_destroy(ok_set* s) {
    if (!s) return OK_SUCCESS;
    if (s->builder) give_builder(s->builder);
    if (s->d_keys) keys_free(s->d_keys, s->slab);
    cudaFree(s->d_table); cudaFree(s->d_pb); cudaFree(s->d_po); cudaFree(s->d_ph);
    cudaFree(s->d_mlo); cudaFree(s->d_mkeys); cudaFree(s->d_mtab);
    delete s;
    return OK_SUCCESS;
}

namespace {
// |A n B| into *d_out (device, zeroed by the caller) on stream st; A is the smaller set.  Small sets keep the plain
// per-key search (one launch); larger ones the tiled two-launch form.  d_lo: scratch of >= |A| / OK_IS_TILE + 2 words.
void launch_intersection(const ok_set* a, const ok_set* b, unsigned long long* d_lo, unsigned long long* d_out, cudaStream_t st) {
    static const bool plain = getenv("ORION_INTERSECT_PLAIN") != nullptr;      // A/B knob: the per-key search for every size
    if (a->n < 4 * OK_IS_TILE || !d_lo || plain) {
        LAUNCH(k_intersect_sorted, grid_for(a->n), 256, 0, st, a->d_keys, a->n, b->d_keys, b->n, d_out);
        return;
    }
    const uint64_t n_tiles = (a->n + OK_IS_TILE - 1) / OK_IS_TILE;
    LAUNCH(k_intersect_bounds, grid_for(n_tiles + 1), 256, 0, st, a->d_keys, a->n, b->d_keys, b->n, d_lo);
    LAUNCH(k_intersect_tiled, (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)(g_sms > 0 ? g_sms : 148) * 8), 256, 0, st,
           a->d_keys, a->n, b->d_keys, d_lo, d_out);
}
}  // namespace

OK_EXPORT int ok_set_intersection_size(ok_set* a, ok_set* b, uint64_t* out) {
    if (!a || !b || !out) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_set_intersection_size: NULL argument");
    if (a->k != b->k) return kmer_size_mismatch(a->k, b->k);  // compare.rs:37-39
    TRY(set_seal(a)); TRY(set_seal(b));
    if (a->n > b->n) std::swap(a, b);
    *out = 0;
    if (a->n == 0) return OK_SUCCESS;
    unsigned long long* d = nullptr;
    const uint64_t n_lo = a->n / OK_IS_TILE + 2;
    CU(cudaMalloc((void**)&d, (1 + n_lo) * 8));
    CU(cudaMemsetAsync(d, 0, 8, a->st));
    launch_intersection(a, b, d + 1, d, a->st);
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, a->st));
    CU(cudaStreamSynchronize(a->st));
    cudaFree(d);
    *out = h;
    return OK_SUCCESS;
}

namespace {
// Tile geometry of the keyed all-vs-all (setops.cuh) from the first and last key of every non-empty set and the total
// number of keys.  Host logic only (okx_ava_geometry replays it for the CPU tests).
OkAvaGeo ava_geometry(unsigned k, const unsigned long long* ends /* 2 per set */, const uint64_t* ns, uint64_t n_sets, uint64_t total) {
    OkAvaGeo g{};
    g.key_shift = 64u - 2u * k;
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for (uint64_t s = 0; s < n_sets; ++s) {
        if (!ns[s]) continue;
        lo = std::min(lo, ok_phi32(ends[2 * s], g.key_shift));
        hi = std::max(hi, ok_phi32(ends[2 * s + 1], g.key_shift));
    }
    if (lo > hi) { lo = 0; hi = 0; }                     // no keys at all
    const uint64_t span = (uint64_t)hi - lo + 1u;
    uint64_t tiles = std::max<uint64_t>(1, total / OK_AVA_TARGET);
    tiles = std::min<uint64_t>(tiles, std::min<uint64_t>(span, 1ull << 22));
    g.phi_lo = lo;
    g.n_tiles = (unsigned)tiles;
    g.scale = (tiles << 32) / span;                      // <= 2^32: (span - 1) * scale < tiles * 2^32
    return g;
}

// 0: row by row (kernels.cuh); 1: key by key (setops.cuh).  ORION_AVA_KEYED=1 forces the keyed form onto any input
// it can take (test hook), =0 switches it off (A/B knob).
bool ava_keyed_wanted(ok_set* const* sets, uint64_t n, uint64_t n_parts) {
    if (n_parts != 1 || n < 2 || n > OK_AVA_MAX_SETS) return false;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (sets[i]->has_max || sets[i]->n >= 0xFFFFFFFFull) return false;
        total += sets[i]->n;
    }
    if (const char* ev = getenv("ORION_AVA_KEYED")) return atoi(ev) != 0;
    return n >= 8 && total >= (1ull << 20);
}

// inter[i * n + j] (i < j) = |set i n set j| for all pairs at once; *done = false: a tile did not fit (clustered keys),
// nothing was written and the caller takes the row form
int ava_keyed(ok_set* const* sets, uint64_t n, uint64_t* inter, bool* done) {
    *done = false;
    cudaStream_t st = sets[0]->st;
    const unsigned long long** d_ptrs = nullptr; unsigned long long *d_ns = nullptr, *d_ends = nullptr, *d_out = nullptr;
    unsigned *d_bounds = nullptr, *d_failed = nullptr;
    auto release = [&] { cudaFree((void*)d_ptrs); cudaFree(d_ns); cudaFree(d_ends); cudaFree(d_out); cudaFree(d_bounds); cudaFree(d_failed); };
#define CUR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return set_err(e_ == cudaErrorMemoryAllocation ? OK_ERR_OUT_OF_MEMORY : OK_ERR_CUDA, "CUDA error %s in the keyed all-vs-all", cudaGetErrorName(e_)); } } while (0)
    std::vector<const unsigned long long*> h_ptrs(n);
    std::vector<unsigned long long> h_ns(n), h_ends(2 * n, 0ull);
    uint64_t total = 0, max_n = 0;
    for (uint64_t i = 0; i < n; ++i) { h_ptrs[i] = sets[i]->d_keys; h_ns[i] = sets[i]->n; total += sets[i]->n; max_n = std::max<uint64_t>(max_n, sets[i]->n); }
    CUR(cudaMalloc((void**)&d_ptrs, n * sizeof(void*)));
    CUR(cudaMalloc((void**)&d_ns, n * 8));
    CUR(cudaMalloc((void**)&d_ends, 2 * n * 8));
    CUR(cudaMalloc((void**)&d_out, n * n * 8));
    CUR(cudaMalloc((void**)&d_failed, 4));
    CUR(cudaMemcpyAsync((void*)d_ptrs, h_ptrs.data(), n * sizeof(void*), cudaMemcpyHostToDevice, st));
    CUR(cudaMemcpyAsync(d_ns, h_ns.data(), n * 8, cudaMemcpyHostToDevice, st));
    CUR(cudaMemsetAsync(d_ends, 0, 2 * n * 8, st));
    CUR(cudaMemsetAsync(d_out, 0, n * n * 8, st));
    CUR(cudaMemsetAsync(d_failed, 0, 4, st));
    LAUNCH(k_ava_ends, (unsigned)((n + 255) / 256), 256, 0, st, d_ptrs, (const unsigned long long*)d_ns, (unsigned)n, d_ends);
    CUR(cudaMemcpyAsync(h_ends.data(), d_ends, 2 * n * 8, cudaMemcpyDeviceToHost, st));
    CUR(cudaStreamSynchronize(st));
    if (total == 0) { release(); memset(inter, 0, n * n * 8); *done = true; return OK_SUCCESS; }
    const OkAvaGeo g = ava_geometry(sets[0]->k, h_ends.data(), (const uint64_t*)h_ns.data(), n, total);
    CUR(cudaMalloc((void**)&d_bounds, n * ((size_t)g.n_tiles + 1) * sizeof(unsigned)));
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    const dim3 grid_b((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((max_n + 255) / 256, 64)), (unsigned)n);
    LAUNCH(k_ava_bounds, grid_b, 256, 0, st, d_ptrs, (const unsigned long long*)d_ns, g, d_bounds);
    if (set_smem(k_ava_tiles, sizeof(OkAvaSmem)) != OK_SUCCESS) { release(); return OK_ERR_CUDA; }
    LAUNCH(k_ava_tiles, std::min<unsigned>(g.n_tiles, grid_sm), OK_AVA_THREADS, sizeof(OkAvaSmem), st, d_ptrs, (unsigned)n, g,
           (const unsigned*)d_bounds, d_out, d_failed);
    unsigned failed = 0;
    CUR(cudaMemcpyAsync(&failed, d_failed, 4, cudaMemcpyDeviceToHost, st));
    CUR(cudaStreamSynchronize(st));
    CUR(cudaGetLastError());
    if (!failed) {
        CUR(cudaMemcpyAsync(inter, d_out, n * n * 8, cudaMemcpyDeviceToHost, st));
        CUR(cudaStreamSynchronize(st));
        *done = true;
    }
#undef CUR
    release();
    return OK_SUCCESS;
}
}  // namespace

// pairs (i < j) in row-major order, pair p belongs to part p % n_parts: sizes[n] and the upper-triangle entries of
// this part's pairs (everything else in inter[n*n] is zero).  Multi-GPU: every rank holds all sets, takes one
// part, and one all-reduce(sum) of the matrix completes it.
OK_EXPORT int ok_sets_all_vs_all_part(ok_set* const* sets, uint64_t n, uint64_t part, uint64_t n_parts, uint64_t* sizes,
                                      uint64_t* inter) {
    if ((n && !sets) || !sizes || !inter) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_sets_all_vs_all: NULL argument");
    if (n_parts == 0 || part >= n_parts) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_sets_all_vs_all_part: part %llu of %llu",
                                                         (unsigned long long)part, (unsigned long long)n_parts);
    if (n == 0) return OK_SUCCESS;
    for (uint64_t i = 0; i < n; ++i) {
        if (!sets[i]) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_sets_all_vs_all: NULL set");
        if (sets[i]->k != sets[0]->k) return kmer_size_mismatch(sets[0]->k, sets[i]->k);
        TRY(set_seal(sets[i]));
        sizes[i] = sets[i]->n;
    }
    // Key by key (setops.cuh): every key of every set is read once and the whole matrix accumulated in one pass.
    if (ava_keyed_wanted(sets, n, n_parts)) {
        bool done = false;
        TRY(ava_keyed(sets, n, inter, &done));
        if (done) return OK_SUCCESS;
    }
    // Row by row: set i against every later set of this part in TWO launches (k_intersect_row_*), set i resident in
    // L2 for the whole row.  (Per pair -- two launches each, 65,280 for 256 sets -- it took 2.0 s at 256 x 5 M keys.)
    static const bool pairwise = getenv("ORION_AVA_PAIRWISE") != nullptr;      // A/B knob: the per-pair launches
    cudaStream_t st = sets[0]->st;
    unsigned long long *d = nullptr, *d_lo = nullptr;
    const unsigned long long** d_ptrs = nullptr; unsigned long long* d_ns = nullptr; unsigned* d_cols = nullptr;
    auto release = [&] { cudaFree(d); cudaFree(d_lo); cudaFree((void*)d_ptrs); cudaFree(d_ns); cudaFree(d_cols); };
#define CUR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return set_err(e_ == cudaErrorMemoryAllocation ? OK_ERR_OUT_OF_MEMORY : OK_ERR_CUDA, "CUDA error %s in ok_sets_all_vs_all", cudaGetErrorName(e_)); } } while (0)
    uint64_t max_tiles = 1;
    for (uint64_t i = 0; i < n; ++i) max_tiles = std::max<uint64_t>(max_tiles, (sets[i]->n + OK_IS_TILE - 1) / OK_IS_TILE);
    CUR(cudaMalloc((void**)&d, n * n * 8));
    CUR(cudaMalloc((void**)&d_lo, (max_tiles + 1) * n * 8));
    CUR(cudaMalloc((void**)&d_ptrs, n * sizeof(void*)));
    CUR(cudaMalloc((void**)&d_ns, n * 8));
    std::vector<const unsigned long long*> h_ptrs(n);
    std::vector<unsigned long long> h_ns(n);
    for (uint64_t i = 0; i < n; ++i) { h_ptrs[i] = sets[i]->d_keys; h_ns[i] = sets[i]->n; }
    CUR(cudaMemcpyAsync((void*)d_ptrs, h_ptrs.data(), n * sizeof(void*), cudaMemcpyHostToDevice, st));
    CUR(cudaMemcpyAsync(d_ns, h_ns.data(), n * 8, cudaMemcpyHostToDevice, st));
    CUR(cudaMemsetAsync(d, 0, n * n * 8, st));
    const unsigned grid_sm = (unsigned)(g_sms > 0 ? g_sms : 148);
    // the columns of every row of this part, uploaded once
    std::vector<unsigned> cols; std::vector<uint64_t> row_first(n + 1, 0);
    uint64_t p = 0;
    for (uint64_t i = 0; i < n; ++i) {
        row_first[i] = cols.size();
        for (uint64_t j = i + 1; j < n; ++j, ++p) {
            if (p % n_parts != part) continue;
            if (sets[i]->n == 0 || sets[j]->n == 0) continue;
            if (pairwise) {
                ok_set *a = sets[i], *b = sets[j];
                if (a->n > b->n) std::swap(a, b);
                launch_intersection(a, b, d_lo, d + i * n + j, st);
            } else {
                cols.push_back((unsigned)j);
            }
        }
    }
    row_first[n] = cols.size();
    CUR(cudaMalloc((void**)&d_cols, std::max<size_t>(cols.size(), 1) * sizeof(unsigned)));
    if (!cols.empty()) CUR(cudaMemcpyAsync(d_cols, cols.data(), cols.size() * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    for (uint64_t i = 0; i < n; ++i) {
        const unsigned n_cols = (unsigned)(row_first[i + 1] - row_first[i]);
        if (!n_cols) continue;
        const OkRowSets rs{d_ptrs, d_ns, d_cols + row_first[i]};
        const uint64_t n_tiles = (sets[i]->n + OK_IS_TILE - 1) / OK_IS_TILE;
        LAUNCH(k_intersect_row_bounds, grid_for((n_tiles + 1) * n_cols), 256, 0, st, sets[i]->d_keys, sets[i]->n, rs, n_cols, d_lo);
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles * n_cols, (uint64_t)grid_sm * 8));
        LAUNCH(k_intersect_row_tiled, blocks, 256, 0, st, sets[i]->d_keys, sets[i]->n, rs, n_cols, d_lo, d + i * n, (uint64_t)1);
    }
    CUR(cudaMemcpyAsync(inter, d, n * n * 8, cudaMemcpyDeviceToHost, st));
    CUR(cudaStreamSynchronize(st));
    CUR(cudaGetLastError());
#undef CUR
    release();
    return OK_SUCCESS;
}

OK_EXPORT int ok_sets_all_vs_all(ok_set* const* sets, uint64_t n, uint64_t* sizes, uint64_t* inter) {
    TRY(ok_sets_all_vs_all_part(sets, n, 0, 1, sizes, inter));
    for (uint64_t i = 0; i < n; ++i) {
        inter[i * n + i] = sizes[i];
        for (uint64_t j = 0; j < i; ++j) inter[i * n + j] = inter[j * n + i];
    }
    return OK_SUCCESS;
}

// ================================================================================= probes ==
namespace {
// query.rs:83-107 on a device-resident batch: d_hits[r] = windows of read r whose canonical k-mer is in the set.
//
// Small sets: a hashed table of the whole set (built once, on the first probe), one random probe per window.
// Large sets (PROBE_MERGE_MIN_KEYS and up) are never hashed -- 2.5 G keys made a 40 GB table that answered
// 0.5 G probes/s (TLB and DRAM-page misses on every probe) after a 40 GB build.  Instead, by merge:
//   1. the batch's DISTINCT canonical k-mers, sorted           (the count path, a pooled builder)
//   2. which of them are in the set                             (k_member_tiled: the set is streamed once, in order)
//   3. a hashed table of just those (<= the batch's distinct k-mers, typically L2- or TLB-friendly)
//   4. the per-read probe of the windows against that table     (the same k_extract<SinkProbeReads> as for small sets)
// A window's k-mer is in the set iff it is in (batch k-mers n set), so the hits are the same integers.
constexpr uint64_t PROBE_MERGE_MIN_KEYS = 1ull << 26;

int probe_use_merge(const ok_set* s, uint64_t n_bases) {
    if (const char* ev = getenv("ORION_PROBE_MERGE")) return atoi(ev) != 0;       // A/B knob and test hook: 1 always, 0 never
    return s->n >= PROBE_MERGE_MIN_KEYS && n_bases >= PART_MIN_BASES && !s->d_table;
}

int probe_reads_merge(ok_set* s, int norm_mode, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec,
                      unsigned* d_hits) {
    ok_counter* c = nullptr;
    TRY(take_builder((uint8_t)s->k, norm_mode, 0, &c));
    struct Giveback { ok_counter* c; ~Giveback() { give_builder(c); } } gb{c};
    TraceClock tc;
    TRY(ok_counter_add_batch_device(c, d_bases, n_bases, d_off, n_rec));
    const uint64_t *dq = nullptr, *dqc = nullptr; uint64_t nq = 0;
    TRY(ok_counter_finish_device(c, 1, &dq, &dqc, &nq));           // returns with the builder's stream drained
    tc.lap("probe by merge: %llu bases -> %llu distinct k-mers", (unsigned long long)n_bases, (unsigned long long)nq);
    const uint64_t nb = s->n - (s->has_max ? 1 : 0);               // (u64::MAX is never a canonical k-mer of a read)
    if (nq == 0 || nb == 0) return OK_SUCCESS;                     // d_hits is already zero
    const uint64_t n_tiles = (nq + OK_IS_TILE - 1) / OK_IS_TILE;
    TRY(dev_reserve(&s->d_mlo, &s->cap_mlo, n_tiles + 2));
    TRY(dev_reserve(&s->d_mkeys, &s->cap_mkeys, nq));
    unsigned long long* d_nm = s->d_mlo + n_tiles + 1;
    CU(cudaMemsetAsync(d_nm, 0, 8, s->st));
    LAUNCH(k_intersect_bounds, grid_for(n_tiles + 1), 256, 0, s->st, (const unsigned long long*)dq, nq, s->d_keys, nb, s->d_mlo);
    LAUNCH(k_member_tiled, (unsigned)std::min<uint64_t>(n_tiles, (uint64_t)(g_sms > 0 ? g_sms : 148) * 8), 256, 0, s->st,
           (const unsigned long long*)dq, nq, s->d_keys, (const unsigned long long*)s->d_mlo, s->d_mkeys, d_nm);
    unsigned long long nm = 0;
    CU(cudaMemcpyAsync(&nm, d_nm, 8, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaGetLastError());
    tc.lap("probe by merge: %llu of them in the set of %llu keys", nm, (unsigned long long)nb);
    if (nm == 0) return OK_SUCCESS;
    const uint64_t n_tab = std::max<uint64_t>(1024, 2 * (uint64_t)nm);
    TRY(dev_reserve(&s->d_mtab, &s->cap_mtab, n_tab));
    LAUNCH(k_fill_u64, grid_for(n_tab), 256, 0, s->st, s->d_mtab, n_tab, OK_EMPTY_KEY);
    LAUNCH(k_keytable_build, grid_for(nm), 256, 0, s->st, s->d_mtab, n_tab, (const unsigned long long*)s->d_mkeys, (uint64_t)nm);
    SinkProbeReads sink{};
    sink.t = OkKeyTableView{s->d_mtab, n_tab, s->has_max};
    sink.rec_off = d_off; sink.n_rec = n_rec; sink.hits = d_hits;
    const uint64_t n_ex = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    launch_extract(nullptr, d_bases, n_bases, d_off, n_rec, 0, n_ex, s->st, sink, norm_mode, s->k);
    CU(cudaStreamSynchronize(s->st));
    CU(cudaGetLastError());
    tc.lap("probe by merge: table of the matches + per-read probe");
    return OK_SUCCESS;
}

// d_hits zeroed on the set's stream, then one of the two forms; returns with the stream drained
int probe_reads_device(ok_set* s, int norm_mode, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_off, uint64_t n_rec,
                       unsigned* d_hits) {
    TRY(set_seal(s));
    CU(cudaMemsetAsync(d_hits, 0, n_rec * 4, s->st));
    if (n_bases == 0) { CU(cudaStreamSynchronize(s->st)); return OK_SUCCESS; }
    if (probe_use_merge(s, n_bases)) {
        CU(cudaStreamSynchronize(s->st));            // the builder works on its own stream: the reads and the zeroed hits are in place
        return probe_reads_merge(s, norm_mode, d_bases, n_bases, d_off, n_rec, d_hits);
    }
    TRY(set_table(s));
    SinkProbeReads sink{};
    sink.t = OkKeyTableView{s->d_table, s->n_table, s->has_max};
    sink.rec_off = d_off; sink.n_rec = n_rec; sink.hits = d_hits;
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    launch_extract(nullptr, d_bases, n_bases, d_off, n_rec, 0, n_tiles, s->st, sink, norm_mode, s->k);
    CU(cudaStreamSynchronize(s->st));
    CU(cudaGetLastError());
    return OK_SUCCESS;
}
}  // namespace

OK_EXPORT int ok_probe_reads(ok_set* s, int norm_mode, const uint8_t* bases, const uint64_t* rec_offsets,
                             uint64_t n_records, uint32_t* hits_per_read) {
    if (!s || (n_records && (!rec_offsets || !hits_per_read)))
        return set_err(OK_ERR_INVALID_ARGUMENT, "ok_probe_reads: NULL argument");
    if (n_records == 0) return OK_SUCCESS;
    if (rec_offsets[0] != 0) return set_err(OK_ERR_INVALID_ARGUMENT, "rec_offsets[0] must be 0");
    if (norm_mode != OK_NORM_NORMALIZED && norm_mode != OK_NORM_RAW) return set_err(OK_ERR_INVALID_ARGUMENT, "unknown norm_mode %d", norm_mode);
    TRY(set_seal(s));
    const uint64_t n_bases = rec_offsets[n_records];
    memset(hits_per_read, 0, n_records * sizeof(uint32_t));
    if (n_bases == 0) return OK_SUCCESS;
    TRY(dev_reserve(&s->d_pb, &s->cap_pb, n_bases + 64));
    TRY(dev_reserve(&s->d_po, &s->cap_po, n_records + 1));
    TRY(dev_reserve(&s->d_ph, &s->cap_ph, n_records));
    uint8_t* d_b = s->d_pb; uint64_t* d_o = s->d_po; unsigned* d_h = s->d_ph;
    CU(cudaMemcpyAsync(d_b, bases, n_bases, cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(d_o, rec_offsets, (n_records + 1) * 8, cudaMemcpyHostToDevice, s->st));
    TRY(probe_reads_device(s, norm_mode, d_b, n_bases, d_o, n_records, d_h));
    CU(cudaMemcpyAsync(hits_per_read, d_h, n_records * 4, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaGetLastError());
    return OK_SUCCESS;
}

// the same probe with the reads and the result resident in device memory (bench `value` leg, multi-GPU probes)
OK_EXPORT int ok_probe_reads_device(ok_set* s, int norm_mode, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                                    uint64_t n_records, uint32_t* d_hits_per_read) {
    if (!s || (n_records && (!d_rec_offsets || !d_hits_per_read))) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_probe_reads_device: NULL argument");
    if (n_records == 0) return OK_SUCCESS;
    if (n_bases && ((uintptr_t)d_bases & 15u)) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    if (norm_mode != OK_NORM_NORMALIZED && norm_mode != OK_NORM_RAW) return set_err(OK_ERR_INVALID_ARGUMENT, "unknown norm_mode %d", norm_mode);
    return probe_reads_device(s, norm_mode, d_bases, n_bases, d_rec_offsets, n_records, d_hits_per_read);
}

// classify.rs:224-277 probes the SAME input count map against every reference: the input is uploaded once and
// every reference's hashed table is probed on the device (per reference a 5 M-key table is L2-resident; uploading
// 1 GB of input per reference, as one ok_probe_counts call per reference does, took 150 ms each).
OK_EXPORT int ok_probe_counts_many(ok_set* const* refs, uint64_t n_refs, const uint64_t* kmers, const uint64_t* counts,
                                   uint64_t n, uint64_t* matched, uint64_t* depth_sum) {
    if ((n_refs && (!refs || !matched || !depth_sum)) || (n && !kmers))
        return set_err(OK_ERR_INVALID_ARGUMENT, "ok_probe_counts_many: NULL argument");
    for (uint64_t r = 0; r < n_refs; ++r) {
        if (!refs[r]) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_probe_counts_many: NULL set");
        matched[r] = 0; depth_sum[r] = 0;
    }
    if (n == 0 || n_refs == 0) return OK_SUCCESS;
    for (uint64_t r = 0; r < n_refs; ++r) TRY(set_table(refs[r]));
    cudaStream_t st = refs[0]->st;
    unsigned long long *d_k = nullptr, *d_c = nullptr, *d_o = nullptr;
    auto release = [&] { cudaFree(d_k); cudaFree(d_c); cudaFree(d_o); };
#define CUR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { release(); return set_err(OK_ERR_CUDA, "CUDA error %s in ok_probe_counts_many", cudaGetErrorName(e_)); } } while (0)
    CUR(cudaMalloc((void**)&d_k, n * 8));
    if (counts) CUR(cudaMalloc((void**)&d_c, n * 8));
    CUR(cudaMalloc((void**)&d_o, n_refs * 16));
    CUR(cudaMemcpyAsync(d_k, kmers, n * 8, cudaMemcpyHostToDevice, st));
    if (counts) CUR(cudaMemcpyAsync(d_c, counts, n * 8, cudaMemcpyHostToDevice, st));
    CUR(cudaMemsetAsync(d_o, 0, n_refs * 16, st));
    for (uint64_t r = 0; r < n_refs; ++r)
        LAUNCH(k_probe_counts, grid_for(n), 256, 0, st, (OkKeyTableView{refs[r]->d_table, refs[r]->n_table, refs[r]->has_max}),
               (const unsigned long long*)d_k, (const unsigned long long*)d_c, n, d_o + 2 * r);
    std::vector<unsigned long long> h(2 * n_refs, 0);
    CUR(cudaMemcpyAsync(h.data(), d_o, n_refs * 16, cudaMemcpyDeviceToHost, st));
    CUR(cudaStreamSynchronize(st));
    CUR(cudaGetLastError());
#undef CUR
    release();
    for (uint64_t r = 0; r < n_refs; ++r) { matched[r] = h[2 * r]; depth_sum[r] = h[2 * r + 1]; }
    return OK_SUCCESS;
}

OK_EXPORT int ok_probe_counts(ok_set* ref, const uint64_t* kmers, const uint64_t* counts, uint64_t n,
                              uint64_t* matched, uint64_t* depth_sum) {
    if (!ref || !matched || !depth_sum || (n && !kmers)) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_probe_counts: NULL argument");
    return ok_probe_counts_many(&ref, 1, kmers, counts, n, matched, depth_sum);
}

// =================================================================================== pack ==
OK_EXPORT int ok_pack_2bit_device(const uint8_t* d_bases, uint64_t n_bases, int norm_mode, uint64_t* d_codes,
                                  uint32_t* d_valid) {
    TRY(ensure_init());
    if (n_bases == 0) return OK_SUCCESS;
    if (!d_bases || !d_codes || !d_valid) return set_err(OK_ERR_INVALID_ARGUMENT, "ok_pack_2bit_device: NULL argument");
    if ((uintptr_t)d_bases & 15u) return set_err(OK_ERR_INVALID_ARGUMENT, "d_bases must be 16-byte aligned");
    const uint64_t n_groups = (n_bases + 31) / 32;
    if (norm_mode == OK_NORM_NORMALIZED)
        LAUNCH(k_pack_2bit<true>, grid_for(n_groups), 256, 0, 0, d_bases, n_bases, n_groups, (unsigned long long*)d_codes, d_valid);
    else
        LAUNCH(k_pack_2bit<false>, grid_for(n_groups), 256, 0, 0, d_bases, n_bases, n_groups, (unsigned long long*)d_codes, d_valid);
    CU(cudaStreamSynchronize(0));
    CU(cudaGetLastError());
    return OK_SUCCESS;
}

// ============================================================================== test hooks ==
// Not part of the public ABI (not declared in include/): host emulations of the exact
// per-lane code the kernels run, so the arithmetic can be checked against the oracle on a
// machine without a GPU, and a device hook that materialises the extracted k-mers.

// Host walk of k_extract: same tiles, same lanes, same ok_pack32 / ok_window_mask /
// ok_lane_windows.  Appends every canonical k-mer (in stream order) to out.
OK_EXPORT int okx_emulate_extract(const uint8_t* bases, uint64_t n_bases, const uint64_t* rec_off, uint64_t n_rec,
                                  unsigned k, int norm_mode, uint64_t* out, uint64_t cap, uint64_t* n_out) {
    if (k == 0 || k > 32) return invalid_k(k);
    uint64_t n = 0, r_next = 0;
    uint64_t pc = 0; uint32_t pv = 0, ps = 0;
    const uint64_t n_groups = (n_bases + 31) / 32;
    for (uint64_t g = 0; g < n_groups; ++g) {
        uint32_t w[8];
        for (int i = 0; i < 8; ++i) {
            uint32_t x = 0;
            for (int b = 0; b < 4; ++b) { uint64_t p = g * 32 + 4 * i + b; if (p < n_bases) x |= (uint32_t)bases[p] << (8 * b); }
            w[i] = x;
        }
        uint64_t cc; uint32_t cv, cs = 0;
        if (norm_mode == OK_NORM_NORMALIZED) ok_pack32<true>(w, cc, cv); else ok_pack32<false>(w, cc, cv);
        while (r_next < n_rec && rec_off[r_next] < (g + 1) * 32) { cs |= 0x80000000u >> (unsigned)(rec_off[r_next] - g * 32); ++r_next; }
        const uint32_t okm = ok_window_mask(pv, cv, ps, cs, k);
        ok_lane_windows(pc, cc, okm, k, [&](int, uint64_t key) { if (n < cap) out[n] = key; ++n; });
        pc = cc; pv = cv; ps = cs;
    }
    *n_out = n;
    return OK_SUCCESS;
}

// Host model of the ordered table: sequential inserts with the same home/probe rule, then the
// same rank rule as k_readout_write.  out_* must hold n entries.
OK_EXPORT int okx_emulate_table(const uint64_t* keys, uint64_t n, unsigned k, int map_mode, uint64_t n_home,
                                unsigned max_probe, uint64_t min_count, uint64_t* out_keys, uint64_t* out_counts,
                                uint64_t* n_out, uint64_t* n_spilled) {
    const uint64_t n_total = n_home + max_probe;
    std::vector<uint64_t> sk(n_total, OK_EMPTY_KEY), sc(n_total, 0);
    uint64_t spilled = 0;
    const unsigned shift = 64 - 2 * k;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t h = ok_home_slot(keys[i], shift, map_mode, n_home);
        const uint64_t lim = std::min<uint64_t>(h + max_probe, n_total);
        bool placed = false;
        for (uint64_t q = h; q < lim; ++q) {
            if (sk[q] == OK_EMPTY_KEY) sk[q] = keys[i];
            if (sk[q] == keys[i]) { ++sc[q]; placed = true; break; }
        }
        if (!placed) ++spilled;
    }
    auto ld = [&](uint64_t q, uint64_t& kq, uint64_t& cq) { kq = sk[q]; cq = sc[q]; };
    uint64_t before = 0, total = 0;
    for (uint64_t s = 0; s < n_total; ++s) if (sk[s] != OK_EMPTY_KEY && sc[s] >= min_count) ++total;
    for (uint64_t s = 0; s < n_total; ++s) {
        if (sk[s] == OK_EMPTY_KEY || sc[s] < min_count) continue;
        const uint64_t h = ok_home_slot(sk[s], shift, map_mode, n_home);
        const long long adj = min_count > 1 ? ok_rank_adjust<true>(ld, s, sk[s], h, max_probe, n_total, min_count)
                                            : ok_rank_adjust<false>(ld, s, sk[s], h, max_probe, n_total, min_count);
        const uint64_t idx = before + adj;
        if (idx >= total) return set_err(OK_ERR_INTERNAL, "rank out of range");
        out_keys[idx] = sk[s]; out_counts[idx] = sc[s];
        ++before;
    }
    *n_out = total; *n_spilled = spilled;
    return OK_SUCCESS;
}

// Device: run k_extract with the materialising sink on host buffers (unordered output).
OK_EXPORT int okx_device_extract(const uint8_t* bases, const uint64_t* rec_off, uint64_t n_rec, unsigned k,
                                 int norm_mode, uint64_t* out, uint64_t cap, uint64_t* n_out) {
    if (k == 0 || k > 32) return invalid_k(k);
    TRY(ensure_init());
    const uint64_t n_bases = rec_off[n_rec];
    *n_out = 0;
    if (n_bases == 0) return OK_SUCCESS;
    uint8_t* d_b = nullptr; uint64_t* d_o = nullptr; unsigned long long *d_out = nullptr, *d_n = nullptr;
    CU(cudaMalloc((void**)&d_b, n_bases + 64));
    CU(cudaMalloc((void**)&d_o, (n_rec + 1) * 8));
    CU(cudaMalloc((void**)&d_out, std::max<uint64_t>(cap, 1) * 8));
    CU(cudaMalloc((void**)&d_n, 8));
    CU(cudaMemcpy(d_b, bases, n_bases, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_o, rec_off, (n_rec + 1) * 8, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_n, 0, 8));
    SinkEmit sink{d_out, d_n, cap};
    const uint64_t n_tiles = (n_bases + OK_TILE_BASES - 1) / OK_TILE_BASES;
    launch_extract(nullptr, d_b, n_bases, d_o, n_rec, 0, n_tiles, 0, sink, norm_mode, k);
    CU(cudaDeviceSynchronize());
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpy(&n, d_n, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(out, d_out, std::min<uint64_t>(n, cap) * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_b); cudaFree(d_o); cudaFree(d_out); cudaFree(d_n);
    *n_out = n;
    return OK_SUCCESS;
}

// owner rank of each key under the routing rule of k_route (host evaluation of the same code)
OK_EXPORT int okx_owner_of(const uint64_t* keys, uint64_t n, unsigned k, int n_ranks, int32_t* out) {
    if (k == 0 || k > 32) return invalid_k(k);
    for (uint64_t i = 0; i < n; ++i) out[i] = (int32_t)ok_home_slot(keys[i], 64 - 2 * k, OK_MAP_CANON, (uint64_t)n_ranks);
    return OK_SUCCESS;
}

// the plan of a one-shot batch as part_choose_bits / part_slice_step derive it (host logic only, no device needed):
// out[0] = total bits, [1] = level-1 bits, [2] = level-2 bits, [3] = sized from the capacity hint (0/1),
// [4] = 16384-slot count kernel (0/1), [5] = result slices of the deferred pipeline (0: not sliceable)
OK_EXPORT int okx_plan_bits(uint64_t n_units, uint64_t capacity_hint, unsigned k, uint32_t* out) {
    if (k == 0 || k > 32) return invalid_k(k);
    if (!out) return set_err(OK_ERR_INVALID_ARGUMENT, "okx_plan_bits: NULL out");
    ok_counter c;
    c.k = k; c.hint = c.user_hint = capacity_hint;
    PartPlan pl;
    part_choose_bits(&c, n_units, pl, /*use_hint=*/true);
    out[0] = pl.cfg.b1 + pl.cfg.b2; out[1] = pl.cfg.b1; out[2] = pl.cfg.b2; out[3] = pl.hinted ? 1u : 0u; out[4] = pl.big_count ? 1u : 0u;
    out[5] = part_sliceable(pl) ? pl.n_sub / part_slice_step(pl) : 0u;
    return OK_SUCCESS;
}

// the strided level-1 gather of a union (partition.cuh, k_part_scatter_keys<1, false, 0, true>) replayed on the host:
// out[w * 4096 + t * 8 + q] = index of the key thread t of item w holds in register q (u64::MAX: none).  Host logic
// only; the CPU tests check that every key of the array is read exactly once.
OK_EXPORT int okx_strided_order(uint64_t n_keys, uint64_t* out, uint64_t cap, uint64_t* n_items) {
    if (!out || !n_items) return set_err(OK_ERR_INVALID_ARGUMENT, "okx_strided_order: NULL argument");
    const uint64_t rl = ok_strided_rows_len(n_keys), items = ok_strided_items(n_keys);
    *n_items = items;
    if (cap < items * OK_PART_TILE) return set_err(OK_ERR_INVALID_ARGUMENT, "okx_strided_order: out holds %llu entries, %llu needed",
                                                   (unsigned long long)cap, (unsigned long long)(items * OK_PART_TILE));
    for (uint64_t w = 0; w < items; ++w)
        for (unsigned t = 0; t < OK_SK_THREADS; ++t)
            for (unsigned q = 0; q < OK_SK_KPT; q += 2) {
                const uint64_t i = ok_strided_index((uint64_t)(q / 2) * OK_SK_THREADS + t, w, rl);
                uint64_t* o = out + w * OK_PART_TILE + (uint64_t)t * OK_SK_KPT + q;
                o[0] = i < n_keys ? i : ~0ull;
                o[1] = i + 1 < n_keys ? i + 1 : ~0ull;
            }
    return OK_SUCCESS;
}


// tile geometry of the keyed all-vs-all as ava_geometry derives it, and the tile of every key (host logic only):
// geo[0] = n_tiles, [1] = phi_lo, [2] = scale; tiles[i] = ok_ava_tile(keys[i])
OK_EXPORT int okx_ava_geometry(unsigned k, const uint64_t* ends, const uint64_t* ns, uint64_t n_sets, uint64_t total,
                               const uint64_t* keys, uint64_t n_keys, uint64_t* geo, uint32_t* tiles) {
    if (k == 0 || k > 32) return invalid_k(k);
    if (!ends || !ns || !geo) return set_err(OK_ERR_INVALID_ARGUMENT, "okx_ava_geometry: NULL argument");
    const OkAvaGeo g = ava_geometry(k, (const unsigned long long*)ends, ns, n_sets, total);
    geo[0] = g.n_tiles; geo[1] = g.phi_lo; geo[2] = g.scale;
    for (uint64_t i = 0; i < n_keys; ++i) tiles[i] = ok_ava_tile(keys[i], g);
    return OK_SUCCESS;
}
// block (bi, bj) of thread b in k_ava_tiles: out[0] = owns a block (0/1), out[1] = bi, out[2] = bj
OK_EXPORT int okx_ava_block(unsigned b, unsigned n_sets, uint32_t* out) {
    unsigned bi, bj;
    out[0] = ok_ava_block(b, (n_sets + 7u) / 8u, bi, bj) ? 1u : 0u; out[1] = bi; out[2] = bj;
    return OK_SUCCESS;
}
