/* orion_gpu.h -- C ABI of liborion_gpu.so, the B200 (sm_100a) implementation of
 * motroy/orion-kmer's k-mer hot path.
 *
 * The reference (a pure-Rust crate) has no FFI seam of its own; its hot path is the inner
 * loop  normalize -> windows(k) -> seq_to_u64 -> canonical_u64 -> sink  repeated in five
 * drivers.  This header is the batch-level boundary a Rust host would bind with
 * `extern "C"` (see INTEGRATION.md for the binding) -- one group of calls per sink.  Every
 * entry cites the reference code it replaces (paths relative to orion-kmer/ in the
 * reference repository).
 *
 * Conventions
 *   - every function returns an ok_status (0 = success); ok_last_error() gives the
 *     thread-local message, worded like the reference's errors.rs where one exists.
 *   - no C++ types, exceptions or panics cross the boundary.
 *   - batch layout: one contiguous byte buffer of sequences + n_records+1 uint64 offsets
 *     (offsets[0] == 0, record r = bases[offsets[r] .. offsets[r+1])).
 *   - norm_mode OK_NORM_NORMALIZED = count/build/classify semantics: the host has already
 *     removed whitespace (needletail normalize(false)); the device maps acgt/ACGT and u/U->T,
 *     every other byte invalidates the windows that contain it.  OK_NORM_RAW = query
 *     semantics (query.rs:66,81): bytes as they are, only acgtACGT valid.
 *   - result arrays returned through `uint64_t**` are page-locked host memory owned by the
 *     library; release them with ok_free().  Handles are opaque, owned by the caller, and
 *     must not be used from two threads at once.
 *   - one process drives one GPU (ok_init picks it); multi-GPU runs are one process per GPU
 *     with the k-mer exchange between the *_route / *_add_kmers calls.
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with
 *     OK_ERR_NO_DEVICE.
 */
#ifndef ORION_GPU_H
#define ORION_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ok_status {
    OK_SUCCESS = 0,
    OK_ERR_INVALID_KMER_SIZE = 1,   /* errors.rs:6-7   InvalidKmerSize(k)            */
    OK_ERR_KMER_SIZE_MISMATCH = 2,  /* errors.rs:24-25 KmerSizeMismatch(k1,k2)       */
    OK_ERR_INVALID_ARGUMENT = 3,
    OK_ERR_OUT_OF_MEMORY = 4,
    OK_ERR_CUDA = 5,
    OK_ERR_NO_DEVICE = 6,
    OK_ERR_INTERNAL = 7
} ok_status;

typedef enum ok_norm_mode { OK_NORM_NORMALIZED = 0, OK_NORM_RAW = 1 } ok_norm_mode;

typedef struct ok_counter ok_counter; /* replaces DashMap<u64,AtomicUsize>  count.rs:48      */
typedef struct ok_set ok_set;         /* replaces HashSet<u64> of one reference, db_types.rs:13 */

/* ---- lifecycle (commands/mod.rs:10-33 dispatch_command) ------------------------------- */
int ok_init(const int* device_ids, int n_devices); /* n_devices must be 1; NULL -> device 0 */
/* releases what the library keeps between calls (pooled set builders, page-locked result blocks, the sets' stream
 * pool, an empty key slab).  Destroy every counter and set first: handles do not survive a shutdown. */
int ok_shutdown(void);
const char* ok_last_error(void);
const char* ok_version(void);
/* number of kernels this library has launched since ok_init (bench.py's gpu_launches) */
uint64_t ok_launch_count(void);
int ok_synchronize(void);

/* page-locked host buffers for batches and results */
int ok_host_alloc(void** out, uint64_t bytes);
int ok_host_free(void* p);
int ok_free(void* p); /* result arrays handed out by the library */

/* ---- k-mer arithmetic on the host (src/kmer.rs, public API of the crate) --------------- */
/* kmer.rs:37-57  returns 1 and writes *out for Some(v), 0 for None */
int ok_seq_to_u64(const uint8_t* seq, uint64_t len, uint8_t k, uint64_t* out);
/* kmer.rs:61-75  writes k bytes; OK_ERR_INVALID_KMER_SIZE where the reference panics */
int ok_u64_to_seq(uint64_t kmer, uint8_t k, uint8_t* out);
/* kmer.rs:79-94 */
int ok_reverse_complement_u64(uint64_t kmer, uint8_t k, uint64_t* out);
/* kmer.rs:99-106 */
int ok_canonical_u64(uint64_t kmer, uint8_t k, uint64_t* out);

/* ---- counter (count.rs:23-38,48,106-119 ; classify.rs:135-199) ------------------------- */
/* k outside 1..=32 -> OK_ERR_INVALID_KMER_SIZE (count.rs:43-45).  capacity_hint = expected
 * number of distinct canonical k-mers (0 = unknown; the table grows as needed).  The hint is a
 * speed knob only (the reference's DashMap::new() at count.rs:48 takes none): large one-shot
 * batches are cut into sub-partitions sized for their expected DISTINCT k-mers when a hint is
 * given (several windows per shared-memory table slot), for their windows otherwise.  A hint
 * that proves too low costs one recount of the batch and is ignored from then on. */
int ok_counter_create(uint8_t k, int norm_mode, uint64_t capacity_hint, ok_counter** out);
/* replace the capacity hint (0 = none); for a sharded counter: expected distinct k-mers of its shard */
int ok_counter_set_capacity_hint(ok_counter* c, uint64_t capacity_hint);
/* ONE table across all batches (count.rs:48,52-79): batches add up whatever their sizes.  A large batch into a
 * counter that already holds the sorted result of earlier large batches is counted on its own by the same
 * one-shot path and the two sorted runs are merged on the device (counts of equal k-mers summed); batches of more
 * than ~1.8 G bases are cut into sub-batches internally.
 * process_sequence_chunk over every record of the batch; host buffers (pinned or pageable).
 * The buffers are free again when the call returns.  For a large batch into an empty counter
 * the call returns once the batch has landed and been scattered by key range; the rest of the
 * count runs inside ok_counter_finish, slice by slice under the device-to-host copy of the
 * result (or at the next call on the handle that needs the counts). */
int ok_counter_add_batch(ok_counter* c, const uint8_t* bases, const uint64_t* rec_offsets,
                         uint64_t n_records);
/* same, buffers already resident in device memory (n_bases == rec_offsets[n_records]) */
int ok_counter_add_batch_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                const uint64_t* d_rec_offsets, uint64_t n_records);
/* multi-GPU: declare this counter to be shard `rank` of `n_ranks` (1, 2, 4 or 8) key ranges.
 * Must precede the first batch.  A sharded counter only accepts k-mers it owns. */
int ok_counter_set_shard(ok_counter* c, int rank, int n_ranks);
/* multi-GPU: add canonical k-mers received from peers (device pointer), +1 each */
int ok_counter_add_kmers_device(ok_counter* c, const uint64_t* d_kmers, uint64_t n);
/* multi-GPU: extract the batch's canonical k-mers and bucket them by owner rank
 * (owner = key range, balanced by the canonical-k-mer prior).  d_out must hold n_bases
 * entries; on return out_counts[r] (host, n_ranks entries) k-mers for rank r are stored
 * contiguously, rank after rank.  The caller exchanges them (NCCL all-to-all) and feeds
 * what it receives to ok_counter_add_kmers_device. */
int ok_counter_route_batch_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                  const uint64_t* d_rec_offsets, uint64_t n_records,
                                  int n_ranks, uint64_t* d_out, uint64_t* out_counts);
/* multi-GPU, fused form of the exchange: the routing kernel itself writes every owner's k-mers
 * into that owner's receive buffer (peer memory over NVLink), so the transfer overlaps the
 * extraction and no separate all-to-all runs.
 *   ok_peer_buffer_*            a device buffer other ranks can map (CUDA IPC, 64-byte handle)
 *   ok_counter_route_count_device    pass 0: k-mers of this batch per owner rank
 *   ok_counter_route_scatter_device  pass 1: d_dst[r] = this sender's slice of rank r's buffer,
 *                                    exactly counts[r] k-mers are written to it
 * The receiver then calls ok_counter_add_kmers_device on its own buffer (after a barrier). */
int ok_peer_buffer_create(uint64_t bytes, void** d_ptr, uint8_t handle[64]);
int ok_peer_buffer_open(const uint8_t handle[64], void** d_ptr);
int ok_peer_buffer_close(void* d_ptr);
int ok_peer_buffer_destroy(void* d_ptr);
int ok_counter_route_count_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                  const uint64_t* d_rec_offsets, uint64_t n_records, int n_ranks,
                                  uint64_t* out_counts);
int ok_counter_route_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                                    const uint64_t* d_rec_offsets, uint64_t n_records, int n_ranks,
                                    uint64_t* const* d_dst, const uint64_t* counts);
/* multi-GPU, fused exchange without a counting pass ("sharded scatter").  Every rank samples its
 * batch; the ranks exchange the (small) histograms; then each sender's extraction kernel multisplits
 * by (owner rank, level-1 bin) and writes straight into the owners' level-1 regions over NVLink, each
 * region sized from that sender's sample -- so what arrives is already level-1 partitioned and the
 * owner continues with its level-2 scatter.  Collectives stay with the caller (NCCL):
 *   ok_shard_geometry        agree the geometry for batches of up to n_bases_max bases per rank;
 *                            returns the histogram sizes and the size every peer buffer must have
 *   ok_shard_set_buffers     the ok_peer_buffer_* level-1 buffers of all ranks as mapped here
 *   ok_shard_sample_device   -> d_hist_fine[n_ranks << sub_bits], d_hist_l1[n_ranks << l1_bits]
 *        caller: reduce-scatter(sum) d_hist_fine -> d_hist_mine[1 << sub_bits];
 *                all-gather d_hist_l1 -> d_hist_l1_all[n_ranks][n_ranks << l1_bits]
 *   ok_shard_scatter_device  plan + scatter; -> d_cursors_out[n_ranks << l1_bits]
 *        caller: all-gather d_cursors_out -> d_cursors_all (this collective is also the barrier)
 *   ok_shard_count_device    level-2 scatter + count of what the peers wrote
 * A region that overflows its sampled capacity makes ok_shard_scatter_device fail on that rank; all
 * ranks then clear and count the batch through the two-pass ok_counter_route_* calls. */
int ok_shard_geometry(ok_counter* c, uint64_t n_bases_max, uint32_t* sub_bits, uint32_t* l1_bits,
                      uint64_t* buffer_keys);
int ok_shard_set_buffers(ok_counter* c, void* const* d_peer_buffers, uint64_t cap_keys);
/* how many k-mers a rank may receive, as a multiple of the largest batch (default 1.25).  Owners are equal shares of
 * the canonical-k-mer position, balanced for uniform base composition; the summed sample shows the real shares
 * before anything is sent, and a skewed input (35 % GC: 1.6 x on one of 8 owners) needs a larger margin.  Take the
 * geometry again afterwards. */
int ok_shard_set_margin(ok_counter* c, double margin);
int ok_shard_sample_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                           const uint64_t* d_rec_offsets, uint64_t n_records, uint32_t* d_hist_fine,
                           uint32_t* d_hist_l1);
int ok_shard_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases,
                            const uint64_t* d_rec_offsets, uint64_t n_records,
                            const uint32_t* d_hist_mine, const uint32_t* d_hist_l1_all,
                            uint32_t* d_cursors_out);
int ok_shard_count_device(ok_counter* c, const uint32_t* d_cursors_all);
/* multi-GPU, chunked exchange over the copy engines (the default form).  Same sampling and the same key-range
 * ownership as the sharded scatter above, but the SMs never touch NVLink: the batch is scattered in n_chunks
 * chunks; chunk c of a sender writes what it extracts for owner o into its own contiguous sub-block (a header with
 * the fills of its level-1 regions, then the regions), laid out identically in the sender's send buffer and in
 * the owner's level-1 buffer, and ONE plain asynchronous peer copy per (owner, chunk) moves it while the next chunk
 * is being extracted.  The layout is computed on the host from the all-gathered per-chunk histograms.
 *   ok_xchg_geometry        like ok_shard_geometry, also fixes n_chunks; then ok_shard_set_buffers
 *   ok_xchg_sample_device   -> d_hist_fine[n_ranks << sub_bits], d_hist_l1c[n_chunks][n_ranks << l1_bits]
 *        caller: reduce-scatter(sum) d_hist_fine -> d_hist_mine; all-gather d_hist_l1c -> HOST array
 *                h_l1c_all[n_ranks][n_chunks][n_ranks << l1_bits]
 *   ok_xchg_scatter_device  plan + chunked scatter + peer copies; returns when this sender's copies have landed
 *        caller: any collective (e.g. the all-reduce of the ranks' success flags) is the barrier
 *   ok_xchg_count_device    fills from the headers, level-2 scatter chunk by chunk, count
 * An overflowing region fails ok_xchg_scatter_device on that rank; all ranks then fall back as above. */
int ok_xchg_geometry(ok_counter* c, uint64_t n_bases_max, uint32_t* sub_bits, uint32_t* l1_bits, uint32_t* n_chunks,
                     uint64_t* buffer_keys);
int ok_xchg_sample_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                          uint64_t n_records, uint32_t* d_hist_fine, uint32_t* d_hist_l1c);
int ok_xchg_scatter_device(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                           uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* h_l1c_all);
int ok_xchg_count_device(ok_counter* c);
/* ok_xchg_scatter_device in steps, so that the owner's level-2 scatter overlaps the exchange of the later chunks:
 *   ok_xchg_scatter_begin   plan, enqueue every chunk's scatter and its peer copies; returns at once
 *   for chunk = 0 .. n_chunks-1:
 *       ok_xchg_chunk_sent  (host) wait until THIS sender's copies of the chunk have landed in every owner's buffer
 *       caller: a host-side barrier among the ranks (after it, every sender's sub-block of the chunk is complete)
 *       ok_xchg_chunk_recv  enqueue the receive work of the chunk (fills from the headers, level-2 scatter) on the
 *                           library's receive stream
 *   ok_xchg_scatter_end     drain the sender side; reports an overflowed region like ok_xchg_scatter_device
 * then the agreement collective and ok_xchg_count_device as above.  The device never waits on another rank: all
 * cross-rank synchronisation is the caller's host code. */
int ok_xchg_scatter_begin(ok_counter* c, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                          uint64_t n_records, const uint32_t* d_hist_mine, const uint32_t* h_l1c_all);
int ok_xchg_chunk_sent(ok_counter* c, uint32_t chunk);
int ok_xchg_chunk_recv(ok_counter* c, uint32_t chunk);
int ok_xchg_scatter_end(ok_counter* c);
/* count.rs:106-119: entries with count >= min_count, ascending by k-mer value. */
int ok_counter_finish(ok_counter* c, uint64_t min_count, uint64_t** kmers, uint64_t** counts,
                      uint64_t* n);
/* same, results left in device memory owned by the handle (valid until the next call on it) */
int ok_counter_finish_device(ok_counter* c, uint64_t min_count, const uint64_t** d_kmers,
                             const uint64_t** d_counts, uint64_t* n);
/* which count path to use: 0 = automatic (large one-shot batches go through the partitioned
 * shared-memory path, everything else through the device-wide table), 1 = table only,
 * 2 = partitioned whenever the counter is still empty.  Results are identical. */
int ok_counter_set_path(ok_counter* c, int mode);
/* multi-GPU: drop what this rank holds of the batch whose exchange just failed on some rank (every rank then
 * recounts it through another route); the result of earlier batches is kept */
int ok_counter_abort_batch(ok_counter* c);
/* multi-GPU: every rank has counted its share of the batch (ok_shard_count_device / ok_xchg_count_device succeeded
 * everywhere): merge it into the result of the earlier batches.  A no-op for the first batch. */
int ok_counter_commit_batch(ok_counter* c);
/* forget all counts, keep the allocation (bench loops) */
int ok_counter_clear(ok_counter* c);
int ok_counter_destroy(ok_counter* c);

typedef struct ok_counter_stats {
    uint64_t n_slots;        /* table slots                                   */
    uint64_t n_distinct;     /* occupied slots                                */
    uint64_t n_windows;      /* valid windows added so far (sum of counts)    */
    uint64_t n_bases;        /* bases seen so far                             */
    uint64_t max_displacement;
    uint64_t n_spilled;      /* inserts that overflowed the probe limit       */
    uint64_t n_grows;        /* table rebuilds                                */
    float ms_insert;         /* device time of the last add_batch kernels     */
    float ms_readout;        /* device time of the last finish kernels        */
    float ms_fill;           /* device time of the last table fill / rebuild  */
    float ms_route;          /* device time of the last route_batch kernels   */
    /* partitioned (one-shot) path: device time per phase of the last batch          */
    float ms_sample, ms_scatter1, ms_scatter2, ms_count, ms_compact;
    int partitioned;         /* 1 when the current result is a sorted run     */
    float ms_push;           /* sharded scatter: bulk copies into the peers' buffers (NVLink) */
    uint64_t n_deferred;     /* sub-partitions the fast count kernel left to the generic one  */
    float ms_merge;          /* device time spent merging batch runs since the last clear     */
    uint64_t n_merges;       /* merges of a batch's run into the accumulated run              */
} ok_counter_stats;
int ok_counter_get_stats(ok_counter* c, ok_counter_stats* out);

/* ---- k-mer sets (build.rs:23-78,95-116 ; db_types.rs:38-58) ---------------------------- */
int ok_set_create(uint8_t k, int norm_mode, uint64_t capacity_hint, ok_set** out);
int ok_set_add_batch(ok_set* s, const uint8_t* bases, const uint64_t* rec_offsets,
                     uint64_t n_records);                  /* DashSet::insert build.rs:55 */
/* same, batch already resident in device memory */
int ok_set_add_batch_device(ok_set* s, const uint8_t* d_bases, uint64_t n_bases, const uint64_t* d_rec_offsets,
                            uint64_t n_records);
/* build.rs:93-116 for MANY files at once: file i = the device-resident batch (d_bases[i], n_bases[i],
 * d_rec_offsets[i], n_records[i]); out[i] = its sealed set.  The files are independent units: a few host threads
 * (ORION_BUILD_THREADS, default 2) each drive their own pooled builder and streams, so one file's host round trips,
 * allocations and launch gaps are covered by the kernels of the others.  Same sets as n calls of ok_set_create +
 * ok_set_add_batch_device.  On error nothing is returned (all out[i] NULL). */
int ok_sets_build_many_device(uint8_t k, int norm_mode, uint64_t n_files, const uint8_t* const* d_bases,
                              const uint64_t* n_bases, const uint64_t* const* d_rec_offsets, const uint64_t* n_records,
                              ok_set** out);
/* the same with host batches (bases[i], rec_offsets[i], n_records[i]; rec_offsets[i][0] == 0): every worker copies
 * its own file in, under the kernels of the others */
int ok_sets_build_many(uint8_t k, int norm_mode, uint64_t n_files, const uint8_t* const* bases,
                       const uint64_t* const* rec_offsets, const uint64_t* n_records, ok_set** out);
/* a set from an existing sorted, duplicate-free host array (a reference loaded from a .db) */
int ok_set_from_sorted(uint8_t k, const uint64_t* kmers, uint64_t n, ok_set** out);
/* ... or device array (checked on the device; a key-range slice received from a peer in multi-GPU set algebra) */
int ok_set_from_sorted_device(uint8_t k, const uint64_t* d_kmers, uint64_t n, ok_set** out);
/* the sealed set's sorted keys in device memory, valid until the set is destroyed */
int ok_set_keys_device(ok_set* s, const uint64_t** d_kmers, uint64_t* n);
/* keys[first .. first + n) of the sealed set into a caller's device buffer (the send buffer of the key exchange) */
int ok_set_copy_keys_device(ok_set* s, uint64_t first, uint64_t n, uint64_t* d_out);
/* multi-GPU set algebra (compare.rs:51-66, query.rs:77-109, classify.rs:224-277 over key-range shards): where
 * the key ranges of n_ranks owners begin inside this sorted set -- the ownership rule of the sharded count (equal
 * shares of the canonical k-mer position, monotone in the key).  bounds[n_ranks + 1]; rank r owns
 * keys[bounds[r] .. bounds[r+1]).  Intersection sizes, hit counts and (matched, depth) sums over the shards add up
 * to the whole: one all-reduce(sum) of integers completes them. */
int ok_set_shard_bounds(ok_set* s, int n_ranks, uint64_t* bounds);
int ok_set_size(ok_set* s, uint64_t* n);                   /* HashSet::len                */
int ok_set_k(ok_set* s, uint8_t* k);
/* sorted ascending copy on the host (build_tests.rs compares decoded set contents) */
int ok_set_export(ok_set* s, uint64_t** kmers, uint64_t* n);
/* db_types.rs:43-48 get_all_kmers_unified ; all sets must share k */
int ok_set_union(ok_set* const* sets, uint64_t n_sets, ok_set** out);
int ok_set_destroy(ok_set* s);

/* ---- set algebra (compare.rs:51-66) ------------------------------------------------------ */
/* |A n B| ; k mismatch -> OK_ERR_KMER_SIZE_MISMATCH (compare.rs:37-39) */
int ok_set_intersection_size(ok_set* a, ok_set* b, uint64_t* out);
/* sizes[n] and the full symmetric n*n matrix of intersection sizes (row-major) */
int ok_sets_all_vs_all(ok_set* const* sets, uint64_t n, uint64_t* sizes, uint64_t* inter);
/* multi-GPU all-vs-all (BASELINE.json configs[4]): the unordered pairs i < j, numbered row by row, are dealt
 * round robin to n_parts ranks; this call computes the pairs of `part` only: sizes[n] and inter[i*n+j] for its
 * pairs, every other entry 0.  Each rank holds all sets; one all-reduce(sum) of the matrix completes it (the
 * diagonal is sizes[], the lower triangle the mirror image) -- the compare.rs:51-60 loop has no other exchange. */
int ok_sets_all_vs_all_part(ok_set* const* sets, uint64_t n, uint64_t part, uint64_t n_parts, uint64_t* sizes,
                            uint64_t* inter);

/* ---- probes (query.rs:79-108 ; classify.rs:224-236) -------------------------------------- */
/* hits_per_read[r] = number of windows of read r whose canonical k-mer is in the set */
int ok_probe_reads(ok_set* s, int norm_mode, const uint8_t* bases, const uint64_t* rec_offsets,
                   uint64_t n_records, uint32_t* hits_per_read);
/* same, reads and result resident in device memory */
int ok_probe_reads_device(ok_set* s, int norm_mode, const uint8_t* d_bases, uint64_t n_bases,
                          const uint64_t* d_rec_offsets, uint64_t n_records, uint32_t* d_hits_per_read);
/* matched = |{i : kmers[i] in ref}| , depth_sum = sum of counts[i] over those */
int ok_probe_counts(ok_set* ref, const uint64_t* kmers, const uint64_t* counts, uint64_t n,
                    uint64_t* matched, uint64_t* depth_sum);
/* the same input count map against n_refs references (the loop of classify.rs:224-277): the input is uploaded
 * once; matched[r], depth_sum[r] for reference r */
int ok_probe_counts_many(ok_set* const* refs, uint64_t n_refs, const uint64_t* kmers, const uint64_t* counts,
                         uint64_t n, uint64_t* matched, uint64_t* depth_sum);

/* ---- standalone 2-bit packing kernel (north-star subsystem 1) ---------------------------- */
/* d_codes: one uint64 per 32 bases, first base in the top two bits; d_valid: one uint32 per
 * 32 bases, first base in the top bit.  Both arrays hold ceil(n_bases/32) words. */
int ok_pack_2bit_device(const uint8_t* d_bases, uint64_t n_bases, int norm_mode,
                        uint64_t* d_codes, uint32_t* d_valid);

#ifdef __cplusplus
}
#endif
#endif /* ORION_GPU_H */
