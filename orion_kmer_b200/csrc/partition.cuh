// partition.cuh -- the partitioned count path (one-shot batches).
//
// Measured on B200 (tools/microbench.cu, profiles/): a DRAM-resident count table sustains only
// ~17 G load+RED/s (random 32 B sector traffic, DRAM-activation bound) while shared-memory
// atomics run at ~2.4 T/s.  So large batches are counted without a global table at all:
//
//   k_part_sample         1/16 of the tiles -> histogram of sub-partition ids
//   k_part_plan           histogram -> capacity and start of every sub-partition (on the device:
//                         no host round trip between the phases)
//   k_part_scatter_bases  ASCII -> pack -> rolling canonical k-mers -> level-1 partitions
//   k_part_items          level-1 fills -> 4096-key work items
//   k_part_scatter_keys   level-1 partition -> level-2 sub-partitions (TMA bulk loads of the items)
//   k_part_count          one CTA per sub-partition: hashed shared-memory table dedupes and counts
//                         (CAS claim + add), then the distinct keys are bucketed by position and
//                         emitted in key order
//   k_part_count_generic  the few sub-partitions k_part_count defers (oversized / clustered)
//   k_part_compact        sub-partition runs -> the final sorted (k-mer, count) arrays
//
// Partitions are ranges of the monotone position x(key) (kmer_math.cuh ok_canon_pos), so
// sub-partition order == key order and the concatenation of the sorted sub-partitions is the
// sorted count table of count.rs:106-119 -- no sort pass.  Both scatters are shared-memory
// multisplits: every thread holds 16 k-mers in registers, ranks them with one smem atomicAdd
// each, the tile is staged in shared memory in bin order and copied out in runs.
#pragma once
#include <type_traits>
#include "kernels.cuh"

#define OK_PART_TILE 4096u        // keys per CTA round of a scatter (256 threads x 16)
#define OK_PART_MAXBINS 1024u     // bins per scatter level
#define OK_STAGE_SLOTS 8192u      // staging slots of a scatter round
// generic (fallback) count kernel: hashed shared-memory table, sorted afterwards
#define OK_CT_SLOTS 8192u
#define OK_CT_THREADS 512u
// fast count kernel: hashed shared-memory table + position buckets (sizes: OkCount2Cfg)
#define OK_C2_BUCKET_MAX 32u      // a fuller bucket defers the sub-partition to the generic kernel

struct OkPartCfg {
    unsigned key_shift;    // 64 - 2k
    unsigned shard_log2;   // multi-GPU: this rank holds 1/2^shard_log2 of the position space
    unsigned b1, b2;       // bits of the level-1 / level-2 bin id
};

// 64-bit position of a key inside this rank's slice of the key space (monotone)
__device__ __forceinline__ uint64_t ok_part_pos(uint64_t key, const OkPartCfg& c) {
    return ok_canon_pos(key << c.key_shift) << c.shard_log2;
}
// its top 32 bits, which is all the bin arithmetic needs (shard + level-1 + level-2 + bucket bits
// <= 3 + 18 + 10 <= 32 - shard_log2): three 32-bit instructions instead of 64-bit shifts
__device__ __forceinline__ uint32_t ok_part_phi(uint64_t key, const OkPartCfg& c) {
    const uint32_t w = ~(uint32_t)((key << c.key_shift) >> 32);
    return (~__umulhi(w, w)) << c.shard_log2;
}
__device__ __forceinline__ unsigned ok_phi_bin1(uint32_t phi, const OkPartCfg& c) { return c.b1 ? phi >> (32 - c.b1) : 0u; }
__device__ __forceinline__ unsigned ok_phi_bin2(uint32_t phi, const OkPartCfg& c) { return c.b2 ? (phi << c.b1) >> (32 - c.b2) : 0u; }
__device__ __forceinline__ unsigned ok_phi_sub(uint32_t phi, const OkPartCfg& c) {
    const unsigned b = c.b1 + c.b2;
    return b ? phi >> (32 - b) : 0u;
}
template <int LEVEL>
__device__ __forceinline__ unsigned ok_part_bin(uint64_t key, const OkPartCfg& c) {
    const uint32_t phi = ok_part_phi(key, c);
    return LEVEL == 1 ? ok_phi_bin1(phi, c) : ok_phi_bin2(phi, c);
}

struct OkPartSpill { OkSpill sp; OkDevStats* st; };

// multi-GPU routing: bin b of the scatter is owner rank b and its keys go to that rank's
// receive buffer -- peer memory mapped over NVLink (CUDA IPC), or local memory for b == self
struct OkPeerOut { unsigned long long* p[8]; unsigned shift; };   // owner of scatter bin b = b >> shift

// scalars of one batch, device resident (host reads them once, at the end)
struct OkPartScalars {
    unsigned n_items;        // level-2 work items
    unsigned n_deferred;     // sub-partitions left to the generic count kernel
    unsigned total_cap;      // sum of the sub-partition capacities
    unsigned def_done;       // deferred sub-partitions the generic kernel has already counted (sliced runs)
};

// ----------------------------------------------------------------------------- sampling --
// every `stride`-th warp-tile; hist[sub] += 1 per k-mer (global RED; the sample is small)
template <bool MAP_U>
__global__ void __launch_bounds__(256)
k_part_sample(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
              uint64_t n_rec, uint64_t n_tiles, uint64_t stride, unsigned k, OkPartCfg cfg,
              unsigned* __restrict__ hist, bool halo, uint64_t tile_begin = 0) {
    // halo = false: a sampled tile stands alone (no read of the tile before it: half the PCIe traffic when
    // the source is the caller's host buffer); the k-1 windows reaching back are not seen, a ~3 % low
    // bias that only the generously padded single-GPU plan tolerates.  Exact counts need the halo.
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = tile_begin + warp * stride; t < n_tiles; t += warps * stride)      // tiles [tile_begin, n_tiles): a sub-batch
        ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t, t + 1, t + 1, k, lane,
            [&](uint64_t, uint64_t pc, uint64_t cc, uint32_t okmask) {
                ok_lane_windows(pc, cc, okmask, k, [&](int, uint64_t key) {
                    atomicAdd(&hist[ok_phi_sub(ok_part_phi(key, cfg), cfg)], 1u);
                });
            }, halo || stride == 1);
}
__global__ void __launch_bounds__(256)
k_part_sample_keys(const unsigned long long* __restrict__ keys, uint64_t n, uint64_t stride, OkPartCfg cfg,
                   unsigned* __restrict__ hist) {
    // sample whole 256-key chunks so the loads stay coalesced
    const uint64_t n_chunks = (n + 255) / 256;
    for (uint64_t c = blockIdx.x * stride; c < n_chunks; c += (uint64_t)gridDim.x * stride) {
        const uint64_t i = c * 256 + threadIdx.x;
        if (i < n) atomicAdd(&hist[ok_phi_sub(ok_part_phi(keys[i], cfg), cfg)], 1u);
    }
}
// keys = a concatenation of SORTED runs (the union of sealed sets, db_types.rs:43-48): a 256-key chunk of a run covers a
// handful of sub-partitions only and a sample of whole chunks says nothing about the others.  Here every `stride`-th
// KEY is sampled (one 32-byte sector per sample, n / stride of them), which is the Poisson sample the plan assumes.
__global__ void __launch_bounds__(256)
k_part_sample_keys_single(const unsigned long long* __restrict__ keys, uint64_t n, uint64_t stride, OkPartCfg cfg,
                          unsigned* __restrict__ hist) {
    for (uint64_t i = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * stride; i < n; i += (uint64_t)gridDim.x * blockDim.x * stride)
        atomicAdd(&hist[ok_phi_sub(ok_part_phi(__ldg(keys + i), cfg), cfg)], 1u);
}

// ------------------------------------------------------------------------------ planning --
// block-wide exclusive scan helper (1024 threads): returns the exclusive prefix of v, total in *tot
__device__ __forceinline__ unsigned ok_block_excl_scan_1024(unsigned v, unsigned* wsum /*[33]*/, unsigned* tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
    __syncthreads();                       // wsum may still be read from a previous call
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const unsigned w = wsum[lane];
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, winc, o); if (lane >= o) winc += y; }
        wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    *tot = wsum[32];
    return wsum[wid] + inc - v;
}

// capacity of a sub-partition from its sampled count: estimate + 6 sigma of the sampling error
// (+ slack), even so that every region starts 16-byte aligned (TMA bulk loads).  The sigma is taken from
// sampled + 1 and ten more samples' worth of room is added: a SMALL region whose sample happens to be empty
// (probability e^-mean: the chunked multi-GPU exchange lays out ~10^5 regions of a few hundred keys in the
// tests) would otherwise get a capacity below its true size.  Host twin: host_part_capacity (orion_gpu.cu).
__host__ __device__ __forceinline__ unsigned ok_part_capacity(unsigned sampled, unsigned stride, unsigned long long n_units) {
    const unsigned long long est = (unsigned long long)sampled * stride;
    unsigned long long cap = est;
    if (stride > 1) cap += (unsigned long long)(6.0f * sqrtf((float)(est + stride) * (float)stride)) + 10ull * stride + 128ull;
    if (cap > n_units) cap = n_units;
    return (unsigned)((cap + 1ull) & ~1ull);
}

// block-wide sum (1024 threads), result valid in every thread
__device__ __forceinline__ unsigned long long ok_block_sum_1024(unsigned long long v, unsigned long long* wsum /*[33]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = ok_warp_sum(v);
    __syncthreads();
    if (lane == 0) wsum[wid] = v;
    __syncthreads();
    if (wid == 0) { unsigned long long w = ok_warp_sum(wsum[lane]); if (lane == 0) wsum[32] = w; }
    __syncthreads();
    return wsum[32];
}

// Two launches over chunks of 1024 sub-partitions (coalesced; no serial pass over all of them):
// k_part_plan_sums: capacity total of every chunk.   k_part_plan: hist[n_sub] (sample counts) ->
// beg / cursor / cap_end per sub-partition, beg1 / cursor1 / end1 per level-1 bin.  hist is zeroed
// again (it becomes n_distinct later).
__global__ void __launch_bounds__(1024)
k_part_plan_sums(const unsigned* __restrict__ hist, unsigned n_sub, unsigned stride, unsigned n_units,
                 unsigned long long* __restrict__ chunk_sum) {
    __shared__ unsigned long long wsum[33];
    const unsigned p = blockIdx.x * 1024u + threadIdx.x;
    const unsigned long long t = ok_block_sum_1024(p < n_sub ? ok_part_capacity(hist[p], stride, n_units) : 0u, wsum);
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = t;
}
__global__ void __launch_bounds__(1024)
k_part_plan(unsigned* __restrict__ hist, unsigned n_sub, unsigned stride, unsigned n_units, unsigned b2,
            const unsigned long long* __restrict__ chunk_sum, unsigned cap_limit,
            unsigned* __restrict__ beg, unsigned* __restrict__ cursor, unsigned* __restrict__ cap_end,
            unsigned* __restrict__ beg1, unsigned* __restrict__ cursor1, unsigned* __restrict__ end1,
            OkPartScalars* __restrict__ sc) {
    __shared__ unsigned wsum[33];
    __shared__ unsigned long long wsum64[33];
    const unsigned long long before = ok_block_sum_1024(threadIdx.x < blockIdx.x ? chunk_sum[threadIdx.x] : 0ull, wsum64);
    const unsigned p = blockIdx.x * 1024u + threadIdx.x;
    const unsigned cap = p < n_sub ? ok_part_capacity(hist[p], stride, n_units) : 0u;
    unsigned total;
    const unsigned long long run64 = before + ok_block_excl_scan_1024(cap, wsum, &total);
    if (p < n_sub) {
        // the host sized the buffers from a bound on the sum; regions past it (never, unless the
        // caller's estimate of a sharded batch was off) get no room and spill instead
        const unsigned run = (unsigned)min(run64, (unsigned long long)cap_limit);
        const unsigned end = (unsigned)min(run64 + cap, (unsigned long long)cap_limit);
        beg[p] = run; cursor[p] = run; cap_end[p] = end;
        hist[p] = 0;
        if (b2 && (p & ((1u << b2) - 1u)) == 0) { beg1[p >> b2] = run; cursor1[p >> b2] = run; }
        if (b2 && (p & ((1u << b2) - 1u)) == (1u << b2) - 1u) end1[p >> b2] = end;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { sc->total_cap = (unsigned)before + total; sc->n_items = 0; sc->n_deferred = 0; sc->def_done = 0; }
}

// exclusive scan of n_distinct[n_sub] -> base[n_sub + 1] (u64), same two-launch shape
// A sliced run scans chunks [chunk0, chunk0 + gridDim.x) only: the chunk sums of the earlier slices are
// already in chunk_sum, so `before` carries the running total across slices.
__global__ void __launch_bounds__(1024)
k_part_scan_sums(const unsigned* __restrict__ v, unsigned n, unsigned long long* __restrict__ chunk_sum, unsigned chunk0) {
    __shared__ unsigned long long wsum[33];
    const unsigned cb = blockIdx.x + chunk0;
    const unsigned i = cb * 1024u + threadIdx.x;
    const unsigned long long t = ok_block_sum_1024(i < n ? v[i] : 0u, wsum);
    if (threadIdx.x == 0) chunk_sum[cb] = t;
}
__global__ void __launch_bounds__(1024)
k_part_scan(const unsigned* __restrict__ v, unsigned n, const unsigned long long* __restrict__ chunk_sum,
            unsigned long long* __restrict__ base, unsigned chunk0, unsigned long long* __restrict__ host_total,
            OkPartScalars* __restrict__ sc) {
    __shared__ unsigned wsum[33];
    __shared__ unsigned long long wsum64[33];
    const unsigned cb = blockIdx.x + chunk0;
    const unsigned long long before = ok_block_sum_1024(threadIdx.x < cb ? chunk_sum[threadIdx.x] : 0ull, wsum64);
    const unsigned i = cb * 1024u + threadIdx.x;
    unsigned total;
    const unsigned ex = ok_block_excl_scan_1024(i < n ? v[i] : 0u, wsum, &total);
    if (i < n) base[i] = before + ex;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        base[min(n, (cb + 1u) * 1024u)] = before + total;   // running total: the next slice rewrites it with the same value
        if (host_total) *host_total = before + total;        // page-locked host word (zero-copy store): no D2H copy to queue
        if (sc) sc->def_done = sc->n_deferred;               // what the generic kernel has counted so far
    }
}
// windows held by sub-partitions [p0, p1) after the scatters (one CTA): the sliced result pipeline
// sizes the result buffers from the first slice's distinct / windows ratio
__global__ void __launch_bounds__(1024)
k_part_slice_fill(const unsigned* __restrict__ beg, const unsigned* __restrict__ fill_end, const unsigned* __restrict__ cap_end,
                  unsigned p0, unsigned p1, unsigned long long* __restrict__ host_out) {
    __shared__ unsigned long long wsum[33];
    unsigned long long t = 0;
    for (unsigned p = p0 + threadIdx.x; p < p1; p += 1024u) {
        const unsigned b0 = beg[p], e0 = min(fill_end[p], cap_end[p]);
        t += e0 > b0 ? e0 - b0 : 0u;
    }
    t = ok_block_sum_1024(t, wsum);
    if (threadIdx.x == 0) *host_out = t;
}

// CTAs of 1024 threads (each repeats the cheap scan over the bins, then builds its share of the items):
// level-1 fills -> work items of <= OK_PART_TILE keys for the level-2 scatter
__global__ void __launch_bounds__(1024)
k_part_items(const unsigned* __restrict__ beg1, const unsigned* __restrict__ cursor1, const unsigned* __restrict__ end1,
             unsigned n_bin1, unsigned bin_mask, unsigned* __restrict__ item_off, unsigned* __restrict__ item_n,
             unsigned* __restrict__ item_bin, OkPartScalars* __restrict__ sc, unsigned* __restrict__ bin_first /* [MAXBINS + 1] or null */) {
    // the sharded path hands in (sender, bin) regions: region r belongs to level-1 bin r & bin_mask
    __shared__ unsigned wsum[33];
    __shared__ unsigned s_first[OK_PART_MAXBINS + 1], s_b0[OK_PART_MAXBINS], s_fill[OK_PART_MAXBINS];
    const unsigned b = threadIdx.x;
    unsigned fill = 0, b0 = 0;
    if (b < n_bin1) { b0 = beg1[b]; const unsigned e = min(cursor1[b], end1[b]); fill = e > b0 ? e - b0 : 0u; }
    const unsigned ni = (fill + OK_PART_TILE - 1u) / OK_PART_TILE;
    unsigned total;
    const unsigned first = ok_block_excl_scan_1024(ni, wsum, &total);
    s_first[b] = first; s_b0[b] = b0; s_fill[b] = fill;
    if (b == 0) s_first[OK_PART_MAXBINS] = total;
    if (bin_first && blockIdx.x == 0) { bin_first[b] = first; if (b == 0) bin_first[OK_PART_MAXBINS] = total; }   // bins past n_bin1 hold no items: first == total
    __syncthreads();
    for (unsigned o = blockIdx.x * 1024u + threadIdx.x; o < total; o += gridDim.x * 1024u) {     // item o belongs to the last bin whose first item is <= o
        unsigned lo = 0, hi = OK_PART_MAXBINS - 1u;
        while (lo < hi) { const unsigned mid = (lo + hi + 1u) >> 1; if (s_first[mid] <= o) lo = mid; else hi = mid - 1u; }
        const unsigned i = o - s_first[lo];
        item_off[o] = s_b0[lo] + i * OK_PART_TILE;
        item_n[o] = min(OK_PART_TILE, s_fill[lo] - i * OK_PART_TILE);
        item_bin[o] = lo & bin_mask;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) sc->n_items = total;
}

// ------------------------------------------------------- sharded (multi-GPU) planning kernels --
// Fused exchange without a counting pass.  An owner's level-1 buffer is cut into one BLOCK per sender
// and each block into one region per level-1 bin, sized from THAT sender's sample -- so a sender
// reserves space with local atomics only, and what it produces for one owner is one contiguous
// block.  The sender's extraction kernel multisplits by (owner, bin): its own keys go straight into
// its own buffer, the other owners' blocks are built in local memory (the level-2 buffer, idle at
// that point) and pushed over NVLink as one bulk copy per peer -- measured 4x faster than scattering
// 24-byte runs into peer memory directly.  All ranks derive the same layout from the all-gathered
// level-1 histograms l1_all[sender][owner][bin].
__global__ void __launch_bounds__(128)
k_shard_l1_hist(const unsigned* __restrict__ hist_fine, unsigned b2, unsigned* __restrict__ l1) {
    __shared__ unsigned wsum[4];
    unsigned v = 0;
    for (unsigned j = threadIdx.x; j < (1u << b2); j += 128u) v += hist_fine[((size_t)blockIdx.x << b2) + j];
    v = (unsigned)ok_warp_sum(v);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) l1[blockIdx.x] = wsum[0] + wsum[1] + wsum[2] + wsum[3];
}

struct OkShardBlocks {              // per owner o: the block this sender fills
    unsigned remote_start[8];       // where it starts in owner o's buffer
    unsigned local_base[8];         // where it is built locally (index into the send buffer; unused for o == me)
    unsigned cap[8];                // its capacity in keys
};

// one CTA, 1024 threads.  Region r = sender * n_bin1 + bin of owner o starts where the capacities before it end.
// send_cur / send_end are in the coordinates the scatter kernel writes with: the own buffer for o == me,
// the local send buffer otherwise.
__global__ void __launch_bounds__(1024)
k_shard_plan(const unsigned* __restrict__ l1_all, unsigned g_log2, unsigned me, unsigned b1, unsigned stride,
             unsigned buf_cap, unsigned* __restrict__ reg_beg, unsigned* __restrict__ reg_end,
             unsigned* __restrict__ send_cur, unsigned* __restrict__ send_end, OkShardBlocks* __restrict__ blk) {
    __shared__ unsigned wsum[33];
    __shared__ OkShardBlocks sb;
    const unsigned G = 1u << g_log2, n_bin1 = 1u << b1, n_reg = n_bin1 << g_log2;
    const unsigned r = threadIdx.x, s = r >> b1, b = r & (n_bin1 - 1u);
    for (unsigned o = 0; o < G; ++o) {
        const unsigned cap = r < n_reg ? ok_part_capacity(l1_all[((size_t)s * G + o) * n_bin1 + b], stride, buf_cap) : 0u;
        unsigned total;
        const unsigned start = ok_block_excl_scan_1024(cap, wsum, &total);
        if (r < n_reg) {
            const unsigned lo = min(start, buf_cap), hi = (unsigned)min((unsigned long long)start + cap, (unsigned long long)buf_cap);
            if (s == me) {
                send_cur[o * n_bin1 + b] = lo; send_end[o * n_bin1 + b] = hi;
                if (b == 0) sb.remote_start[o] = lo;
                if (b == n_bin1 - 1u) sb.cap[o] = hi;      // block end for now
            }
            if (o == me) { reg_beg[r] = lo; reg_end[r] = hi; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (unsigned o = 0; o < G; ++o) {
            sb.cap[o] -= sb.remote_start[o];
            sb.local_base[o] = run;
            if (o != me) run += sb.cap[o];
        }
        *blk = sb;
    }
    __syncthreads();
    const unsigned i = threadIdx.x, o = i >> b1;
    if (i < n_reg && o != me) {      // other owners' regions: remote coordinates -> the local send buffer
        const unsigned shift = sb.local_base[o] - sb.remote_start[o];
        send_cur[i] += shift; send_end[i] += shift;
    }
}

// this sender's cursors after the scatter, back in the owners' coordinates (to be all-gathered)
__global__ void __launch_bounds__(1024)
k_shard_export(const unsigned* __restrict__ send_cur, unsigned g_log2, unsigned me, unsigned b1,
               const OkShardBlocks* __restrict__ blk, unsigned* __restrict__ out) {
    const unsigned i = threadIdx.x, o = i >> b1;
    if (i < ((1u << b1) << g_log2)) out[i] = o == me ? send_cur[i] : send_cur[i] - blk->local_base[o] + blk->remote_start[o];
}

// cursors of every sender after the scatter (all-gathered) -> fill of each of my regions; their sum
// is the number of k-mers this rank received
__global__ void __launch_bounds__(1024)
k_shard_fills(const unsigned* __restrict__ cur_all, unsigned g_log2, unsigned me, unsigned b1,
              const unsigned* __restrict__ reg_beg, const unsigned* __restrict__ reg_end,
              unsigned* __restrict__ reg_fill, unsigned long long* __restrict__ n_received) {
    __shared__ unsigned long long wsum[33];
    const unsigned G = 1u << g_log2, n_bin1 = 1u << b1, n_reg = n_bin1 << g_log2;
    const unsigned r = threadIdx.x, s = r >> b1, b = r & (n_bin1 - 1u);
    unsigned long long got = 0;
    if (r < n_reg) {
        const unsigned e = min(cur_all[((size_t)s * G + me) * n_bin1 + b], reg_end[r]);
        reg_fill[r] = e;
        got = e > reg_beg[r] ? e - reg_beg[r] : 0u;
    }
    got = ok_block_sum_1024(got, wsum);
    if (threadIdx.x == 0) *n_received = got;
}

// ------------------------------------------ chunked exchange over the copy engines (ok_xchg_*) --
// Third form of the multi-GPU exchange.  The batch is scattered in a few CHUNKS; chunk c of sender s writes
// what it extracts for owner o into its own SUB-BLOCK (s -> o, c): a header (the fills of its level-1
// regions) followed by one region per level-1 bin, sized from the sender's per-chunk sample.  A sub-block
// is contiguous and laid out identically in the sender's send buffer and in the owner's level-1 buffer, so
// ONE plain asynchronous peer copy per (peer, chunk) moves it -- issued on a per-peer copy stream as soon
// as the chunk's scatter kernel has finished, i.e. under the extraction of the next chunk.  The SMs never
// touch NVLink (measured: SM-issued remote stores saturate near 400 GB/s per GPU in an 8-way all-to-all,
// a copy-engine peer copy reaches 720-770).  The layout is computed on the host from the all-gathered
// per-chunk level-1 histograms, identically on every rank.
template <bool MAP_U>
__global__ void __launch_bounds__(256)
k_xchg_sample(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
              uint64_t n_rec, uint64_t n_tiles, uint64_t stride, unsigned k, OkPartCfg cfg /* global: b1 = g + sub_bits */,
              unsigned* __restrict__ hist_fine, unsigned l1_shift /* 32 - (g + level-1 bits) */, unsigned n_regs,
              uint64_t tiles_per_chunk, unsigned n_chunks, unsigned* __restrict__ hist_l1c /* [n_chunks][n_regs] */) {
    extern __shared__ unsigned sh_l1c[];      // the (chunk, owner, level-1 bin) histogram of this CTA
    for (unsigned i = threadIdx.x; i < n_chunks * n_regs; i += blockDim.x) sh_l1c[i] = 0;
    __syncthreads();
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = warp * stride; t < n_tiles; t += warps * stride) {
        const unsigned chunk = (unsigned)min((uint64_t)(n_chunks - 1u), t / tiles_per_chunk);
        unsigned* l1 = sh_l1c + chunk * n_regs;
        ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t, t + 1, t + 1, k, lane,
            [&](uint64_t, uint64_t pc, uint64_t cc, uint32_t okmask) {
                ok_lane_windows(pc, cc, okmask, k, [&](int, uint64_t key) {
                    const uint32_t phi = ok_part_phi(key, cfg);
                    atomicAdd(&hist_fine[ok_phi_sub(phi, cfg)], 1u);
                    atomicAdd(&l1[phi >> l1_shift], 1u);
                });
            }, true);
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n_chunks * n_regs; i += blockDim.x) { const unsigned v = sh_l1c[i]; if (v) atomicAdd(&hist_l1c[i], v); }
}

// after a chunk's scatter: the fills of this sender's regions go into the header of every sub-block, so they
// travel with the data (no second collective).  hdr[o] = key offset of the header of sub-block (me -> o, chunk)
// in the buffer the scatter wrote it to (own level-1 buffer for o == me, the send buffer otherwise).
__global__ void __launch_bounds__(1024)
k_xchg_headers(const unsigned* __restrict__ cur, const unsigned* __restrict__ end, const unsigned* __restrict__ beg,
               unsigned n_regs, unsigned b1, unsigned me, const unsigned* __restrict__ hdr,
               unsigned long long* __restrict__ own, unsigned long long* __restrict__ send) {
    const unsigned i = threadIdx.x;
    if (i >= n_regs) return;
    const unsigned o = i >> b1, b = i & ((1u << b1) - 1u);
    const unsigned e = min(cur[i], end[i]);
    reinterpret_cast<unsigned*>((o == me ? own : send) + hdr[o])[b] = e > beg[i] ? e - beg[i] : 0u;
}

// receiver: headers of the sub-blocks (s -> me, chunk) -> fill cursor of each of my regions of that chunk
__global__ void __launch_bounds__(1024)
k_xchg_fills(const unsigned long long* __restrict__ own, const unsigned* __restrict__ hdr /* [senders] */, unsigned n_regs, unsigned b1,
             const unsigned* __restrict__ reg_beg, const unsigned* __restrict__ reg_end, unsigned* __restrict__ reg_fill,
             unsigned long long* __restrict__ n_received /* += */) {
    __shared__ unsigned long long wsum[33];
    const unsigned r = threadIdx.x, s = r >> b1, b = r & ((1u << b1) - 1u);
    unsigned long long got = 0;
    if (r < n_regs) {
        const unsigned fill = reinterpret_cast<const unsigned*>(own + hdr[s])[b];
        const unsigned cap = reg_end[r] - reg_beg[r];
        const unsigned f = min(fill, cap);
        reg_fill[r] = reg_beg[r] + f;
        got = f;
    }
    got = ok_block_sum_1024(got, wsum);
    if (threadIdx.x == 0 && got) atomicAdd(n_received, got);
}

// ------------------------------------------------------------- shared multisplit machinery --
// Slotted staging: the 8192 staging slots are split evenly among the bins of the level, a key's
// rank inside its bin (one shared-memory atomicAdd) is its slot, so one pass stages the round.
// A bin that outgrows its slots in a round sends the excess straight to global memory.
// CTA barrier of the scatter threads.  The sharded (PEER) scatter carries one extra warp that only
// copies to the peers and never joins these barriers, hence a named barrier with an explicit count.
#ifndef OK_SB_KPT
#define OK_SB_KPT 8                                   // keys per thread and round in the level-1 scatter
#endif
#define OK_SB_THREADS (OK_PART_TILE / OK_SB_KPT)      // scatter threads per CTA (512 at 8 keys per thread)
#define OK_SB_WARPS (OK_SB_THREADS / 32)
template <bool PEER> __device__ __forceinline__ void ok_scatter_sync() {
    if (PEER) asm volatile("bar.sync 1, %0;" :: "n"((int)OK_SB_THREADS) : "memory");
    else __syncthreads();
}

struct OkScatterSmem {
    unsigned long long stage[OK_STAGE_SLOTS];
    uint2 hg[OK_PART_MAXBINS];   // x: keys of the bin this round (zero between rounds), y: global index of its first staged key
};

__device__ __forceinline__ void ok_part_put(unsigned long long key, unsigned dst, unsigned end,
                                            unsigned long long* __restrict__ out, const OkPartSpill& ps) {
    if (dst < end) out[dst] = key;
    else ok_spill(ps.sp, ps.st, key, 1);     // past the sampled capacity of the bin: exact, slow path
}

// One multisplit round of the CTA (all 4096/KPT threads call it together): thread-held keys
// key[0..KPT) (bit q of vmask says key[q] exists) -> out, grouped by bin.  sm.hg[].x must be zero
// on entry and is zero again on exit.  after_stage() runs once the round's keys have left the
// registers of every thread (the level-2 kernel issues its next TMA load there).
template <int LEVEL, bool PEER, int KPT, class AfterStage, unsigned NW = OK_PART_TILE / KPT / 32u /* warps of the CTA */>
__device__ __forceinline__ void ok_multisplit(OkScatterSmem& sm, const uint64_t (&key)[KPT], unsigned vmask,
                                                const OkPartCfg& cfg, unsigned bins_log2,
                                                unsigned* __restrict__ cursors, const unsigned* __restrict__ bin_end,
                                                unsigned long long* __restrict__ out, const OkPartSpill& ps,
                                                const OkPeerOut* peer, AfterStage&& after_stage) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned cap_log2 = 13u - bins_log2, cap = 1u << cap_log2, n_bins = 1u << bins_log2;
    // (Ranking a group of keys first and storing them afterwards, so that the shared-memory atomics overlap, was
    // measured SLOWER: level 1 5.8 -> 6.4 ms, level 2 3.8 -> 4.1 ms.  Key by key it stays.)
#pragma unroll
    for (int q = 0; q < KPT; ++q)
        if (vmask >> q & 1u) {
            const unsigned b = ok_part_bin<LEVEL>(key[q], cfg);
            const unsigned r = atomicAdd(&sm.hg[b].x, 1u);
            // rank r of bin b sits in slot (r + b) mod cap of the bin: without the rotation the bank of a staged
            // key depends on r alone and most ranks are 0..3, an 8-way conflict on every store (548 M per pass, ncu)
            if (r < cap) sm.stage[(b << cap_log2) + ((r + b) & (cap - 1u))] = key[q];
            else ok_part_put(key[q], atomicAdd(&cursors[b], 1u), bin_end[b], PEER ? peer->p[b >> peer->shift] : out, ps);
        }
    ok_scatter_sync<PEER>();
    after_stage();
    // copy out.  Warp w owns an equal share of the staging slots = a contiguous range of bins.
    // (1) one global cursor bump per non-empty bin; a bin whose region is full spills its tail here
    const unsigned bins_per_warp = n_bins >= NW ? n_bins / NW : 1u;
    const unsigned wb0 = wid * bins_per_warp;
    for (unsigned i = lane; i < bins_per_warp && wb0 + i < n_bins; i += 32) {
        const unsigned b = wb0 + i;
        unsigned c = sm.hg[b].x, g = 0;
        if (c > cap) c = cap;
        if (c) {
            g = atomicAdd(&cursors[b], c);
            const unsigned e = bin_end[b];
            if (g + c > e) {
                const unsigned keep = g < e ? e - g : 0u;
                for (unsigned r = keep; r < c; ++r) ok_spill(ps.sp, ps.st, sm.stage[(b << cap_log2) + ((r + b) & (cap - 1u))], 1);
                c = keep;
            }
        }
        sm.hg[b] = make_uint2(c, g);
    }
    __syncwarp();
    // (2) dense walk over the warp's staging slots: consecutive lanes = consecutive slots of a bin
    if (NW == 16u && cap_log2 >= 3u && cap_log2 <= 5u) {
        // 256 / 512 / 1024 bins with 16 warps: every warp owns 512 staging slots = 16 passes of 32 lanes, a pass covers
        // 1 / 2 / 4 whole bins.  Unrolled with the slots-per-bin as a constant, the staging and header addresses become
        // immediates and the slot -> rank rotation one subtraction per pass: ~9 instructions per pass instead of 17
        // in the generic walk below (the copy-out was 28 % of the level-1 scatter's instructions).
        auto walk = [&](auto cl2) {
            constexpr unsigned C = decltype(cl2)::value, BPP = 32u >> C /* bins per pass */, M = (1u << C) - 1u;
            const unsigned sub = lane >> C;
            const unsigned long long* __restrict__ st = sm.stage + (wb0 << C) + lane;
            const uint2* __restrict__ hg = sm.hg + wb0 + sub;
            unsigned r = ((lane & M) - wb0 - sub) & M;
#pragma unroll 8
            for (int it = 0; it < 16; ++it) {
                const uint2 h = hg[it * BPP];
                if (r < h.x) (PEER ? peer->p[(wb0 + sub + it * BPP) >> peer->shift] : out)[h.y + r] = st[it * 32];
                r = (r - BPP) & M;
            }
        };
        if (cap_log2 == 5u) walk(std::integral_constant<unsigned, 5>{});
        else if (cap_log2 == 4u) walk(std::integral_constant<unsigned, 4>{});
        else walk(std::integral_constant<unsigned, 3>{});
    } else if (wb0 < n_bins) {
        const unsigned t_end = (wb0 + bins_per_warp) << cap_log2;
#pragma unroll 4
        for (unsigned t = (wb0 << cap_log2) + lane; t < t_end; t += 32) {
            const unsigned b = t >> cap_log2, r = (t - b) & (cap - 1u);      // undo the rotation: slot -> rank
            const uint2 h = sm.hg[b];
            if (r < h.x) (PEER ? peer->p[b >> peer->shift] : out)[h.y + r] = sm.stage[t];
        }
    }
    __syncwarp();
    for (unsigned i = lane; i < bins_per_warp && wb0 + i < n_bins; i += 32) sm.hg[wb0 + i].x = 0;
    ok_scatter_sync<PEER>();
}

// TMA bulk copy global -> shared, completion on an mbarrier (one elected thread issues it)
__device__ __forceinline__ uint32_t ok_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ok_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ok_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ok_tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ok_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(ok_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(ok_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ok_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ok_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ok_mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OK_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OK_MBAR_DONE_%=;\n"
        "bra OK_MBAR_WAIT_%=;\n"
        "OK_MBAR_DONE_%=:\n"
        "}\n" :: "r"(ok_smem_u32(bar)), "r"(phase) : "memory");
}

// TMA bulk store shared -> global (local or peer-mapped), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void ok_tma_store_1d(void* dst_gmem, const void* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst_gmem), "r"(ok_smem_u32(src_smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void ok_tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ok_tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------- sharded scatter: push fused into the scatter --
// The batch is scattered in a few chunks (launches).  Every CTA of a launch carries a ninth warp that
// does no scattering: it copies what the PREVIOUS chunk appended to the other owners' blocks (cursor
// deltas prev -> cur per region, contiguous runs) from the local send buffer into the owners' buffers
// over NVLink while the other eight warps extract and scatter the next chunk.  One kernel computes and
// communicates; only the last chunk's push is exposed.  (Measured: an SM sustains only ~4 GB/s of
// remote stores -- outstanding-request bound -- so the copy has to be spread over ALL SMs; dedicated
// pusher CTAs on a few SMs reached 300 GB/s, direct 24-byte scatter runs into peer memory 175 GB/s.)
struct OkPushDesc {
    const unsigned* prev;                 // cursors before the chunk being pushed (local coordinates)
    const unsigned* cur;                  // ... and after it
    const unsigned* end;                  // region ends (a cursor past it means the rest was spilled)
    const OkShardBlocks* blk;
    const unsigned long long* local;      // the local send buffer
    unsigned long long* peer[8];          // the owners' level-1 buffers
    unsigned n_regs, b1, me, enabled;     // enabled == 0: nothing to push
    unsigned tma;                         // 1: the copy warp moves the data with TMA bulk copies (ok_shard_push_tma)
};

// the calling group of `nthreads` threads (rank `tid` in it) is worker `worker` of `n_workers`: the regions are
// cut into 8192-key pieces and dealt round robin, so the workers stay balanced whatever the region sizes are
__device__ __forceinline__ void ok_shard_push(const OkPushDesc& d, unsigned worker, unsigned n_workers, unsigned tid, unsigned nthreads) {
    constexpr unsigned PIECE = 8192u;
    unsigned turn = worker;               // pieces until my next one
    for (unsigned reg = 0; reg < d.n_regs; ++reg) {
        const unsigned o = reg >> d.b1;
        if (o == d.me) continue;
        const unsigned lo0 = d.prev[reg], hi0 = min(d.cur[reg], d.end[reg]);
        if (hi0 <= lo0) continue;
        const unsigned n_pieces = (hi0 - lo0 + PIECE - 1u) / PIECE;
        if (turn >= n_pieces) { turn -= n_pieces; continue; }
        const unsigned long long* __restrict__ src = d.local;
        // block starts are even in both coordinate systems, so index parity == 16-byte alignment on both sides
        unsigned long long* __restrict__ dst = d.peer[o] + (long long)d.blk->remote_start[o] - (long long)d.blk->local_base[o];
        for (; turn < n_pieces; turn += n_workers) {
            unsigned lo = lo0 + turn * PIECE;
            const unsigned hi = min(lo + PIECE, hi0);
            if (lo & 1u) { if (tid == 0) dst[lo] = src[lo]; ++lo; }
            const unsigned n2 = (hi - lo) >> 1;                         // 16-byte units
            const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src + lo);
            uint4* __restrict__ d4 = reinterpret_cast<uint4*>(dst + lo);
            unsigned i = tid;
            for (; i + 7u * nthreads < n2; i += 8u * nthreads) {       // 128 bytes in flight per thread: NVLink latency is ~2 us
                uint4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ok_ld_stream16(s4 + i + q * nthreads);
#pragma unroll
                for (int q = 0; q < 8; ++q) d4[i + q * nthreads] = v[q];
            }
            for (; i < n2; i += nthreads) d4[i] = ok_ld_stream16(s4 + i);
            if (((hi - lo) & 1u) && tid == 0) dst[hi - 1] = src[hi - 1];
        }
        turn -= n_pieces;
    }
}

// The same copy through the TMA unit: lanes 0..OK_PUSH_STAGES-1 of the copy warp each own one 4 KB staging buffer
// and stream their share of a piece through it -- bulk load local -> shared (mbarrier), bulk store shared -> peer
// (bulk async-group).  No registers hold data and a lane only waits until the engine has READ its buffer, not until
// the remote write has landed, so the bytes in flight per SM are no longer bounded by what 32 threads can keep in
// registers.  MEASURED AND NOT ADOPTED (ORION_PUSH_TMA=1 selects it): at 2 GPUs the fused scatter+push takes 11.4 ms
// with it against 9.0 ms with the register copies, i.e. it tops out near 400 GB/s as well.  At 8 GPUs the register
// copies (386 GB/s per GPU inside the scatter) and the stand-alone k_shard_push with 16x the bytes in flight
// (435 GB/s) land at the same figure, so the all-to-all of SM stores, not the copy loop, is the bound there.
#define OK_PUSH_STAGES 4u
#define OK_PUSH_CHUNK 512u                 // keys per staging buffer (4 KB)
struct OkPushSmem {
    unsigned long long stage[OK_PUSH_STAGES][OK_PUSH_CHUNK];
    unsigned long long bar[OK_PUSH_STAGES];
};
__device__ __forceinline__ void ok_shard_push_tma(const OkPushDesc& d, unsigned worker, unsigned n_workers, unsigned lane, OkPushSmem& ps) {
    constexpr unsigned PIECE = 8192u;
    if (lane < OK_PUSH_STAGES) ok_mbar_init(&ps.bar[lane], 1);
    __syncwarp();
    unsigned phase = 0;
    unsigned turn = worker;               // pieces until my next one
    for (unsigned reg = 0; reg < d.n_regs; ++reg) {
        const unsigned o = reg >> d.b1;
        if (o == d.me) continue;
        const unsigned lo0 = d.prev[reg], hi0 = min(d.cur[reg], d.end[reg]);
        if (hi0 <= lo0) continue;
        const unsigned n_pieces = (hi0 - lo0 + PIECE - 1u) / PIECE;
        if (turn >= n_pieces) { turn -= n_pieces; continue; }
        const unsigned long long* __restrict__ src = d.local;
        // block starts are even in both coordinate systems, so index parity == 16-byte alignment on both sides
        unsigned long long* __restrict__ dst = d.peer[o] + (long long)d.blk->remote_start[o] - (long long)d.blk->local_base[o];
        for (; turn < n_pieces; turn += n_workers) {
            unsigned lo = lo0 + turn * PIECE;
            unsigned hi = min(lo + PIECE, hi0);
            if (lo & 1u) { if (lane == 0) dst[lo] = src[lo]; ++lo; }
            if ((hi - lo) & 1u) { --hi; if (lane == 0) dst[hi] = src[hi]; }
            if (lane < OK_PUSH_STAGES) {
                for (unsigned c0 = lo + lane * OK_PUSH_CHUNK; c0 < hi; c0 += OK_PUSH_STAGES * OK_PUSH_CHUNK) {
                    const unsigned bytes = (min(c0 + OK_PUSH_CHUNK, hi) - c0) * 8u;      // a multiple of 16
                    ok_tma_store_wait_read();                                             // my buffer's previous store has read it
                    ok_tma_load_1d(ps.stage[lane], src + c0, bytes, &ps.bar[lane]);
                    ok_mbar_wait(&ps.bar[lane], phase); phase ^= 1u;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    ok_tma_store_1d(dst + c0, ps.stage[lane], bytes);
                }
            }
        }
        turn -= n_pieces;
    }
    if (lane < OK_PUSH_STAGES) ok_tma_store_wait_all();       // the remote writes are done before the kernel ends
    __syncwarp();
}

__global__ void __launch_bounds__(256) k_shard_push(const __grid_constant__ OkPushDesc d) {
    ok_shard_push(d, blockIdx.x, gridDim.x, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------ level 1: from the bases --
// The warps of a CTA walk their own runs of tiles in lock step; each warp-tile (32 window ends per
// lane) is split in rounds of OK_SB_KPT window ends per lane, so a round holds <= 4096 k-mers per CTA.
// The launch covers tiles [tile_begin, tile_end) -- the ingest pipeline launches it once per landed piece.
// KC: k as a compile-time constant (31 and 21, the configurations of BASELINE.json; 0 = any k at run time).
// With k fixed the 64-bit shifts of the rolling update and of the position become immediate-operand funnel shifts.
// P3: a warp-tile (32 window ends per lane) is split in 3 rounds of 11/11/10 instead of 4 rounds of 8: the 8192 staging
// slots are filled to ~2/3 instead of 1/2 per round (fewer empty slots walked by the copy-out, fewer barriers and cursor
// bumps per key).  Only for <= 256 bins (32 slots per bin: a bin overflows its slots in < 2 % of the rounds).
template <bool MAP_U, bool PEER = false, int KC = 0, bool P3 = false>
__global__ void __launch_bounds__(PEER ? OK_SB_THREADS + 32 : OK_SB_THREADS, OK_SB_KPT == 16 ? 3 : 2)
k_part_scatter_bases(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
                     uint64_t n_rec, uint64_t tile_begin, uint64_t tile_end, uint64_t tiles_per_warp, unsigned k_in,
                     OkPartCfg cfg_in, unsigned* __restrict__ cursors, const unsigned* __restrict__ bin_end,
                     unsigned long long* __restrict__ out, OkPartSpill ps, unsigned long long* __restrict__ n_keys,
                     const __grid_constant__ OkPeerOut peer_out, const __grid_constant__ OkPushDesc push) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OkScatterSmem& sm = *reinterpret_cast<OkScatterSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const unsigned k = KC ? (unsigned)KC : k_in;
    OkPartCfg cfg = cfg_in;
    if (KC) cfg.key_shift = 64u - 2u * KC;
    if (PEER && threadIdx.x >= OK_SB_THREADS) {      // the copy warp of a sharded scatter (launched with 32 more threads)
        if (push.enabled) {
            if (push.tma) ok_shard_push_tma(push, blockIdx.x, gridDim.x, threadIdx.x - OK_SB_THREADS, *reinterpret_cast<OkPushSmem*>(smem_raw + sizeof(OkScatterSmem)));
            else ok_shard_push(push, blockIdx.x, gridDim.x, threadIdx.x - OK_SB_THREADS, 32u);
        }
        return;
    }
    const uint64_t warp = blockIdx.x * (uint64_t)OK_SB_WARPS + (threadIdx.x >> 5);
    const uint64_t t0 = tile_begin + warp * tiles_per_warp;
    if (tile_begin + (uint64_t)blockIdx.x * OK_SB_WARPS * tiles_per_warp >= tile_end) return;   // whole CTA idle
    for (unsigned i = threadIdx.x; i < OK_PART_MAXBINS; i += OK_SB_THREADS) sm.hg[i] = make_uint2(0u, 0u);
    ok_scatter_sync<PEER>();
    unsigned long long my_keys = 0;
    ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t0, t0 + tiles_per_warp, tile_end, k, lane,
        [&](uint64_t, uint64_t pc, uint64_t cc, uint32_t okmask) {
            OkRoll roll; roll.init(pc, cc, k);
            my_keys += __popc(okmask);
            const unsigned rev = __brev(okmask);      // window end j lives in okmask bit 31-j: bit j of rev
            constexpr int KP = P3 ? 11 : OK_SB_KPT, NP = P3 ? 3 : 32 / OK_SB_KPT;
#pragma unroll
            for (int part = 0; part < NP; ++part) {
                uint64_t key[KP];
#pragma unroll
                for (int q = 0; q < KP; ++q) key[q] = KP * part + q < 32 ? roll.step(KP * part + q) : 0ull;
                const unsigned vm = rev >> (KP * part) & ((1u << KP) - 1u);   // bit q <=> key[q]; window ends past 31 do not exist
                auto nop = [] {};
                ok_multisplit<1, PEER, KP, decltype(nop)&, OK_SB_WARPS>(sm, key, vm, cfg, cfg.b1, cursors, bin_end, out, ps, &peer_out, nop);
            }
        });
    my_keys = ok_warp_sum(my_keys);
    if (lane == 0 && my_keys) atomicAdd(n_keys, my_keys);
}

// --------------------------------------------------------------------- level 2: from keys --
struct OkScatterKeysSmem {
    OkScatterSmem sc;
    unsigned long long in[OK_PART_TILE];     // landing buffer of the next work item (TMA)
    unsigned long long bar;                  // mbarrier of the landing buffer
};

// work item w: keys src[item_off[w] .. +item_n[w]) (<= 4096), all of level-1 bin item_bin[w].
// item_off is even (16-byte aligned regions) and the copy is rounded up to an even key count.
// TMA = false: plain loads (a caller's key array that is not 16-byte aligned).
// 512 threads x 8 keys: with two CTAs per SM that is 32 resident warps to cover the shared-memory atomics
// and the global cursor bumps of a round (the kernel is latency-, not issue-bound).
#define OK_SK_KPT 8
#define OK_SK_THREADS (OK_PART_TILE / OK_SK_KPT)
// STRIDED (LEVEL 1, plain loads): src is a concatenation of SORTED runs (the union of sealed sets).  4096 consecutive
// keys of a run fall into one or two bins: every rank comes from the same shared-memory counter and all but `cap` keys
// of the round leave through the per-key overflow path (measured: 2.5 G keys/s against 70 G for shuffled keys).  So an
// item is gathered from OK_SK_PLACES places spread evenly over the whole array, one 16-byte pair of keys from each
// (ok_strided_index): the places of an item belong to different runs and key ranges, the bins fill evenly again, and
// the eight items that share a 128-byte line run side by side (consecutive CTAs), so DRAM still sees whole lines.
#define OK_SK_PLACES (OK_PART_TILE / 2u)
// rows of the strided view: the array is cut into OK_SK_PLACES rows of `rl` 128-byte lines (16 keys)
__host__ __device__ __forceinline__ uint64_t ok_strided_rows_len(uint64_t n_keys) {
    const uint64_t n_lines = (n_keys + 15u) / 16u;
    return (n_lines + OK_SK_PLACES - 1u) / OK_SK_PLACES;
}
__host__ __device__ __forceinline__ uint64_t ok_strided_items(uint64_t n_keys) { return 8u * ok_strided_rows_len(n_keys); }
// first key index of pair `place` (0 .. OK_SK_PLACES) of item w: a bijection (place, w, 0/1) <-> [0, 16 rl OK_SK_PLACES)
__host__ __device__ __forceinline__ uint64_t ok_strided_index(uint64_t place, uint64_t w, uint64_t rl) {
    return (place * rl + (w >> 3)) * 16u + (w & 7u) * 2u;
}
__global__ void k_part_set_items(OkPartScalars* __restrict__ sc, unsigned n_items) { sc->n_items = n_items; }

template <int LEVEL, bool TMA = true, int KC = 0, bool STRIDED = false>  // LEVEL 1: bin by bin1 (keys arriving from peers); LEVEL 2: bin by bin2 inside a bin1
__global__ void __launch_bounds__(OK_SK_THREADS, 2)
k_part_scatter_keys(const unsigned long long* __restrict__ src, const unsigned* __restrict__ item_off,
                    const unsigned* __restrict__ item_n, const unsigned* __restrict__ item_bin,
                    const OkPartScalars* __restrict__ scal, OkPartCfg cfg_in, unsigned* __restrict__ cursors,
                    const unsigned* __restrict__ bin_end, unsigned long long* __restrict__ out, OkPartSpill ps,
                    const unsigned* __restrict__ bin_first = nullptr, unsigned bin_lo = 0, unsigned bin_hi = 0) {
    static_assert(!STRIDED || (LEVEL == 1 && !TMA), "the strided gather is a level-1 scatter with plain loads");
    // STRIDED: there is no sliced run of a flat key array; (bin_lo, bin_hi) carry the number of keys in src instead
    const uint64_t flat_n = STRIDED ? ((uint64_t)bin_hi << 32 | bin_lo) : 0;
    if (STRIDED) { bin_first = nullptr; }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    OkScatterKeysSmem& sm = *reinterpret_cast<OkScatterKeysSmem*>(smem_raw);
    OkPartCfg cfg = cfg_in;
    if (KC) cfg.key_shift = 64u - 2u * KC;
    const unsigned bins_log2 = LEVEL == 1 ? cfg.b1 : cfg.b2;
    // a sliced run handles the items of level-1 bins [bin_lo, bin_hi) only
    const unsigned w_begin = bin_first ? bin_first[bin_lo] : 0u;
    const unsigned n_items = bin_first ? bin_first[bin_hi] : scal->n_items;
    for (unsigned i = threadIdx.x; i < OK_PART_MAXBINS; i += OK_SK_THREADS) sm.sc.hg[i] = make_uint2(0u, 0u);
    // LEVEL 2 reads our own level-1 buffer (even capacities, slack at the end): an odd item is
    // rounded UP to whole 16-byte units.  LEVEL 1 reads a caller's array: rounded DOWN, and the
    // odd last key is fetched with a plain load.
    auto load_item = [&](unsigned w, unsigned n) {
        const unsigned bytes = (LEVEL == 2 ? (n + 1u) & ~1u : n & ~1u) * 8u;
        if (bytes) ok_tma_load_1d(sm.in, src + item_off[w], bytes, &sm.bar);
        else ok_mbar_arrive(&sm.bar);
    };
    // the descriptors of a round are fetched one round ahead (two dependent L2 round trips per round otherwise)
    unsigned n = 0, bin = 0;
    if (!STRIDED && w_begin + blockIdx.x < n_items) { n = item_n[w_begin + blockIdx.x]; if (LEVEL == 2) bin = item_bin[w_begin + blockIdx.x]; }
    const uint64_t strided_rl = STRIDED ? ok_strided_rows_len(flat_n) : 0;
    if (TMA && threadIdx.x == 0) {
        ok_mbar_init(&sm.bar, 1);
        if (w_begin + blockIdx.x < n_items) load_item(w_begin + blockIdx.x, n);
    }
    __syncthreads();
    unsigned phase = 0;
    for (unsigned w = w_begin + blockIdx.x; w < n_items; w += gridDim.x) {
        const unsigned wn = w + gridDim.x;
        unsigned n_next = 0, bin_next = 0;
        if (!STRIDED && wn < n_items) { n_next = item_n[wn]; if (LEVEL == 2) bin_next = item_bin[wn]; }
        const unsigned bin_base = LEVEL == 1 ? 0u : bin << cfg.b2;
        uint64_t key[OK_SK_KPT]; unsigned vm = 0;
        if (TMA) {
            ok_mbar_wait(&sm.bar, phase); phase ^= 1u;
#pragma unroll
            for (int q = 0; q < OK_SK_KPT; ++q) {
                const unsigned i = q * OK_SK_THREADS + threadIdx.x;
                key[q] = sm.in[i];
                if (LEVEL == 1 && i + 1 == n && (n & 1u)) key[q] = src[item_off[w] + i];
                if (i < n) vm |= 1u << q;
            }
        } else if (STRIDED) {
#pragma unroll
            for (int q = 0; q < OK_SK_KPT; q += 2) {
                const uint64_t i = ok_strided_index((uint64_t)(q / 2) * OK_SK_THREADS + threadIdx.x, w, strided_rl);
                key[q] = 0; key[q + 1] = 0;
                if (i + 1 < flat_n) {        // (normal L2 policy: the neighbouring items read the rest of the line)
                    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(src + i));
                    key[q] = v.x; key[q + 1] = v.y; vm |= 3u << q;
                } else if (i < flat_n) {
                    key[q] = __ldg(src + i); vm |= 1u << q;
                }
            }
        } else {
            const unsigned long long* __restrict__ keys = src + item_off[w];
#pragma unroll
            for (int q = 0; q < OK_SK_KPT; ++q) {
                const unsigned i = q * OK_SK_THREADS + threadIdx.x;
                key[q] = 0;
                if (i < n) { key[q] = __ldcs(keys + i); vm |= 1u << q; }
            }
        }
        auto next_load = [&] {   // every thread holds its keys in registers: the landing buffer is free again
            if (TMA && threadIdx.x == 0 && wn < n_items) load_item(wn, n_next);
        };
        ok_multisplit<LEVEL, false, OK_SK_KPT, decltype(next_load)&, OK_SK_THREADS / 32u>(
            sm.sc, key, vm, cfg, bins_log2, cursors + bin_base, bin_end + bin_base, out, ps, nullptr, next_load);
        n = n_next; bin = bin_next;
    }
}

// plain work items over a flat key array (keys received from peers): item w = keys [w*4096, ...)
__global__ void __launch_bounds__(1024)
k_part_flat_items(unsigned n_keys, unsigned* __restrict__ item_off, unsigned* __restrict__ item_n,
                  unsigned* __restrict__ item_bin, OkPartScalars* __restrict__ sc) {
    const unsigned ni = (n_keys + OK_PART_TILE - 1u) / OK_PART_TILE;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < ni; i += gridDim.x * blockDim.x) {
        item_off[i] = i * OK_PART_TILE;
        item_n[i] = min(OK_PART_TILE, n_keys - i * OK_PART_TILE);
        item_bin[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->n_items = ni;
}

// ------------------------------------------------ count one sub-partition in shared memory --
// sub-partition p holds keys src[beg[p] .. fill_end[p]).  Its distinct keys come out sorted in
// place (keys -> src[beg[p] ..], counts -> cnt_out[beg[p] ..]); n_distinct[p] says how many.
//
// Fast kernel.  (1) dedupe + count in a HASHED table (CAS claim + add): duplicates -- ~80 % of the
// windows at 30x coverage -- hit their key on the first probe, and a k-mer and its tail-error
// variants no longer share a slot the way they must under an order-preserving placement.
// (2) order the distinct keys only: bucket them by the next 10 position bits (two sweeps over the
// table around one block scan), insertion-sort each bucket (~1 key per bucket), emit coalesced.
// Two sizes: <13> = 8192 slots, sub-partitions up to 6144 keys, 512 threads, 2 CTAs per SM (the
// normal case); <14> = 16384 slots, up to 12288 keys, 1024 threads, 1 CTA per SM (batches whose
// sub-partitions cannot be made smaller: 8-GPU routing, > 1.2 G keys per batch).
template <int LOG2> struct OkCount2Cfg {
    static constexpr unsigned SLOTS = 1u << LOG2, MAXKEYS = 3u << (LOG2 - 2), THREADS = 1u << (LOG2 - 4),
                              BUCKET_BITS = LOG2 - 2, BUCKETS = 1u << BUCKET_BITS, BPT = BUCKETS / THREADS /* buckets per thread in the scan: 4 */,
                              WARPS = THREADS / 32u,
                              WQ = MAXKEYS * 2u / 8u / WARPS,      // pending-key queue entries per warp (carved from sidx): 96
                              MAXN = 61440u,                       // windows per sub-partition: counts are 16-bit
                              LIMIT = MAXKEYS - 64u;               // distinct keys after which no new round is started: the rounds
                                                                   // in flight then claim < 4 * THREADS more, so an EMPTY slot always remains
};
template <int LOG2> struct OkCount2Smem {
    using C = OkCount2Cfg<LOG2>;
    unsigned long long tkey[C::SLOTS];             // 64 KB / 128 KB
    unsigned tcnt[C::SLOTS / 2];                   // two 16-bit counts per word (a count <= MAXKEYS < 65536)
    unsigned short newl[C::MAXKEYS];               // table slots of the distinct keys, in claim order
    unsigned short sidx[C::MAXKEYS];               // the same, grouped by bucket
    unsigned boff[C::BUCKETS];                     // bucket histogram -> bucket end offsets
    unsigned wsum[36];                             // [0..31] warp sums, [32..33] the CTA's dense base (u64), [34] its ticket
    unsigned n_new;
};

template <int LOG2> __device__ __forceinline__ unsigned ok_c2_hash(uint64_t key) {
    uint32_t x = (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u);
    x *= 0x85EBCA6Bu;
    return x >> (32 - LOG2);
}
template <int LOG2> __device__ __forceinline__ unsigned ok_c2_bucket(uint64_t key, const OkPartCfg& cfg, unsigned sub_bits) {
    return (ok_part_phi(key, cfg) << sub_bits) >> (32 - OkCount2Cfg<LOG2>::BUCKET_BITS);     // the next position bits
}
template <int LOG2> __device__ __forceinline__ void ok_c2_add(OkCount2Smem<LOG2>& sm, unsigned s) {
    atomicAdd(&sm.tcnt[s >> 1], 1u << ((s & 1u) << 4));
}
// key not found at its first probe (cur = what was read there): walk on, claim an empty slot.
// The thread that claims a slot also files the new distinct key: claim list + bucket histogram.
// The walk only decides WHERE the key lives; the count and the filing happen after the lanes have
// reconverged, once per call instead of once per divergent exit of the loop, and the claim-list
// cursor is bumped once per warp (ballot) instead of once per claiming lane on one shared word.
template <int LOG2>
__device__ __forceinline__ void ok_c2_insert_slow(OkCount2Smem<LOG2>& sm, unsigned long long key, unsigned s,
                                                  unsigned long long cur, const OkPartCfg& cfg, unsigned sub_bits) {
    bool claimed = false;
    for (;;) {
        if (cur == OK_EMPTY_KEY) {
            cur = atomicCAS(&sm.tkey[s], OK_EMPTY_KEY, key);
            claimed = cur == OK_EMPTY_KEY;
        }
        if (claimed || cur == key) break;
        s = (s + 1u) & (OkCount2Cfg<LOG2>::SLOTS - 1u);
        cur = sm.tkey[s];
    }
    ok_c2_add(sm, s);
    const unsigned act = __activemask();
    const unsigned bal = __ballot_sync(act, claimed);
    if (bal) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(bal) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&sm.n_new, __popc(bal));
        base = __shfl_sync(act, base, leader);
        if (claimed) {
            const unsigned at = base + __popc(bal & ((1u << lane) - 1u));
            if (at < OkCount2Cfg<LOG2>::MAXKEYS) sm.newl[at] = (unsigned short)s;     // past it the sub-partition is given up anyway
            atomicAdd(&sm.boff[ok_c2_bucket<LOG2>(key, cfg, sub_bits)], 1u);
        }
    }
}

// Dense output straight from the count kernel (no compaction pass): sub-partition p's distinct keys go to
// out[base(p) ..], base(p) = distinct keys of all sub-partitions before it, found by a decoupled look-back over
// lb[]: after its insert phase a CTA publishes its own total (AGG), then sums its predecessors' totals backwards
// until it meets one that already knows its inclusive prefix (PFX), and publishes its own.  Sub-partitions are
// handed out by a TICKET, so every predecessor a CTA waits for belongs to a CTA that is already running: no
// deadlock whatever else shares the GPU.  A sub-partition the fast kernel cannot finish (too many windows or
// distinct keys, crowded buckets) FAILS the dense attempt (*failed = 1, it publishes a total of 0 so that nobody
// hangs) and the host recounts the range in the sparse form below; the dense pass never modifies its input.
struct OkDenseOut {
    unsigned long long* lb;            // [n_sub] look-back words; nullptr: sparse output (sorted runs in place)
    unsigned long long* keys;          // dense outputs
    unsigned long long* counts;
    unsigned long long cap;            // entries the dense outputs hold
    unsigned* ticket;                  // zeroed before every launch
    unsigned* failed;
};
#define OK_LB_AGG 0x4000000000000000ull
#define OK_LB_PFX 0x8000000000000000ull
#define OK_LB_VAL 0x3FFFFFFFFFFFFFFFull

// warp 0 of the CTA (all 32 lanes): publish `tot` for sub-partition p, return the exclusive prefix
__device__ __forceinline__ unsigned long long ok_dense_lookback(unsigned long long* lb, unsigned p, unsigned long long tot) {
    const int lane = threadIdx.x & 31;
    if (p == 0) {
        if (lane == 0) { __threadfence(); atomicExch(&lb[0], OK_LB_PFX | tot); }
        return 0ull;
    }
    if (lane == 0) atomicExch(&lb[p], OK_LB_AGG | tot);
    unsigned long long excl = 0;
    long long q = (long long)p - 1;               // lane i looks at q - i
    for (;;) {
        const long long at = q - lane;
        unsigned long long v;
        do {
            v = at >= 0 ? *((volatile unsigned long long*)&lb[at]) : OK_LB_PFX;      // before sub-partition 0: prefix 0
        } while (__any_sync(OK_FULL, (v & (OK_LB_AGG | OK_LB_PFX)) == 0ull));
        const unsigned pfx = __ballot_sync(OK_FULL, (v & OK_LB_PFX) != 0ull);
        const int first = pfx ? __ffs(pfx) - 1 : 32;                                   // nearest predecessor with a prefix
        unsigned long long part = lane <= first ? (v & OK_LB_VAL) : 0ull;
        part = ok_warp_sum(part);
        excl += part;
        if (pfx) break;
        q -= 32;
    }
    if (lane == 0) atomicExch(&lb[p], OK_LB_PFX | (excl + tot));
    return excl;
}

template <int LOG2, int KC = 0>
__global__ void __launch_bounds__(OkCount2Cfg<LOG2>::THREADS, LOG2 == 13 ? 2 : 1)
k_part_count(unsigned long long* __restrict__ src, const unsigned* __restrict__ beg,
             const unsigned* __restrict__ fill_end /* cursor after the scatter */,
             const unsigned* __restrict__ cap_end, unsigned p_begin, unsigned p_end, OkPartCfg cfg_in,
             unsigned* __restrict__ cnt_out /* 32-bit counts of the runs, same indices as src */, unsigned* __restrict__ n_distinct,
             unsigned* __restrict__ deferred, OkPartScalars* __restrict__ scal, unsigned seed_keys /* 0 or THREADS */,
             bool direct_first, const __grid_constant__ OkDenseOut dense) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using C = OkCount2Cfg<LOG2>;
    constexpr unsigned OK_C2_SLOTS = C::SLOTS, OK_C2_MAXKEYS = C::MAXKEYS, OK_C2_THREADS = C::THREADS, OK_C2_BUCKETS = C::BUCKETS;
    OkCount2Smem<LOG2>& sm = *reinterpret_cast<OkCount2Smem<LOG2>*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    OkPartCfg cfg = cfg_in;
    if (KC) cfg.key_shift = 64u - 2u * KC;
    const unsigned sub_bits = cfg.b1 + cfg.b2;
    {
        ulonglong2* k2 = reinterpret_cast<ulonglong2*>(sm.tkey);
        uint4* c4 = reinterpret_cast<uint4*>(sm.tcnt);
        for (unsigned i = threadIdx.x; i < OK_C2_SLOTS / 2; i += OK_C2_THREADS) k2[i] = make_ulonglong2(OK_EMPTY_KEY, OK_EMPTY_KEY);
        for (unsigned i = threadIdx.x; i < OK_C2_SLOTS / 8; i += OK_C2_THREADS) c4[i] = make_uint4(0u, 0u, 0u, 0u);
        for (unsigned i = threadIdx.x; i < OK_C2_BUCKETS; i += OK_C2_THREADS) sm.boff[i] = 0;
        if (threadIdx.x == 0) sm.n_new = 0;
    }
    __syncthreads();
    const bool DENSE = dense.lb != nullptr;
    unsigned long long* const dbase = reinterpret_cast<unsigned long long*>(sm.wsum + 32);     // (wsum[32..33]: the CTA's dense base)
    for (unsigned p = p_begin + blockIdx.x;; p += gridDim.x) {
        if (DENSE) {         // sub-partitions in ticket order (see OkDenseOut)
            __syncthreads();
            if (threadIdx.x == 0) sm.wsum[34] = atomicAdd(dense.ticket, 1u);
            __syncthreads();
            p = p_begin + sm.wsum[34];
        }
        if (p >= p_end) break;
        // invariant here: table empty, counts zero, boff zero, n_new zero
        const unsigned b0 = beg[p];
        const unsigned e0 = min(fill_end[p], cap_end[p]);          // the rest was spilled by the scatter
        const unsigned n = e0 > b0 ? e0 - b0 : 0u;
        if (n == 0) {
            if (threadIdx.x == 0) n_distinct[p] = 0;
            if (DENSE && wid == 0) ok_dense_lookback(dense.lb, p, 0ull);
            continue;
        }
        // The table has to hold the DISTINCT keys only: a sub-partition may bring many more windows than slots
        // (the host sizes sub-partitions from the caller's capacity hint).  If the distinct keys do outgrow the
        // table the insert phase stops early and the sub-partition goes to the generic kernel.
        if (n > C::MAXN) {
            if (threadIdx.x == 0) { deferred[atomicAdd(&scal->n_deferred, 1u)] = p; if (DENSE) *dense.failed = 1u; }
            if (DENSE && wid == 0) ok_dense_lookback(dense.lb, p, 0ull);
            continue;
        }
        const bool may_overflow = n > OK_C2_MAXKEYS;
        const unsigned long long* __restrict__ keys = src + b0;
        // ---- (1) insert: 4 keys in flight per thread, the next 4 already loading.  A key found at its first
        // probe (a duplicate: ~80 % of the windows at 30x coverage) is counted on the spot by all lanes alike.
        // The rest -- new keys and collisions -- are compacted into a per-warp queue (ballot ranks, shared
        // memory borrowed from sidx, unused until pass 3) and drained with all lanes busy, instead of running
        // the divergent claim loop once per key slot with a quarter of the lanes.
        unsigned long long* wq = reinterpret_cast<unsigned long long*>(sm.sidx) + wid * C::WQ;
        unsigned qn = 0;
        auto drain = [&] {
            __syncwarp();
            for (unsigned i = lane; i < qn; i += 32) {
                const unsigned long long key = wq[i];
                const unsigned h = ok_c2_hash<LOG2>(key);
                ok_c2_insert_slow(sm, key, h, sm.tkey[h], cfg, sub_bits);
            }
            __syncwarp();
            qn = 0;
        };
        // Round 0: one key per thread goes straight through the claim loop.  The table is empty, so a first
        // probe cannot hit; with 4 keys per thread the whole first round (half of a typical sub-partition)
        // used to miss, queue up and take the slow path even though most of those keys are duplicates of
        // one another.  Seeding the table with 1/8 of the keys first (at 30x coverage ~95 % of the repeated
        // k-mers are among them) lets the rounds that follow hit on their first probe.
        // (16-byte loads of key pairs were measured SLOWER here: 8.0 vs 7.4 ms; the four 8-byte loads stay)
        unsigned long long nx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned i = seed_keys + q * OK_C2_THREADS + threadIdx.x;
            nx[q] = i < n ? __ldcs(keys + i) : OK_EMPTY_KEY;
        }
        if (seed_keys) {
            if (threadIdx.x < n) {
                const unsigned long long key = __ldcs(keys + threadIdx.x);
                const unsigned h = ok_c2_hash<LOG2>(key);
                ok_c2_insert_slow(sm, key, h, sm.tkey[h], cfg, sub_bits);
            }
            __syncthreads();
        }
        for (unsigned base = seed_keys; base < n; base += 4 * OK_C2_THREADS) {
            // (a vote, not a bare read of the shared word: other warps bump n_new meanwhile, and the full-mask ballots below
            // need every lane of the warp to take the same way out whether or not the lanes read it in the same instant)
            if (may_overflow && __any_sync(OK_FULL, *(volatile unsigned*)&sm.n_new > C::LIMIT)) break;
            unsigned long long kk[4], cur[4]; unsigned hs[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                kk[q] = nx[q];
                const unsigned i = base + (4 + q) * OK_C2_THREADS + threadIdx.x;
                nx[q] = i < n ? __ldcs(keys + i) : OK_EMPTY_KEY;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { hs[q] = ok_c2_hash<LOG2>(kk[q]); cur[q] = sm.tkey[hs[q]]; }
            if (direct_first && base == seed_keys) {     // empty table: no first probe can hit, skip the queue
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (kk[q] != OK_EMPTY_KEY) ok_c2_insert_slow(sm, kk[q], hs[q], sm.tkey[hs[q]], cfg, sub_bits);
                continue;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool live = kk[q] != OK_EMPTY_KEY;             // canonical k-mers never equal the sentinel
                if (live && cur[q] == kk[q]) ok_c2_add(sm, hs[q]);   // duplicate of a key already placed: the common case
                const bool pending = live && cur[q] != kk[q];
                const unsigned bal = __ballot_sync(OK_FULL, pending);
                if (bal) {
                    const unsigned cnt = __popc(bal);
                    if (qn + cnt > C::WQ) drain();                   // warp-uniform
                    if (pending) wq[qn + __popc(bal & ((1u << lane) - 1u))] = kk[q];
                    qn += cnt;
                }
            }
            drain();
        }
        __syncthreads();
        if (may_overflow && sm.n_new > C::LIMIT) {       // block-uniform: too many distinct keys, start over clean
            __syncthreads();
            ulonglong2* k2 = reinterpret_cast<ulonglong2*>(sm.tkey);
            uint4* c4 = reinterpret_cast<uint4*>(sm.tcnt);
            for (unsigned i = threadIdx.x; i < OK_C2_SLOTS / 2; i += OK_C2_THREADS) k2[i] = make_ulonglong2(OK_EMPTY_KEY, OK_EMPTY_KEY);
            for (unsigned i = threadIdx.x; i < OK_C2_SLOTS / 8; i += OK_C2_THREADS) c4[i] = make_uint4(0u, 0u, 0u, 0u);
            for (unsigned i = threadIdx.x; i < OK_C2_BUCKETS; i += OK_C2_THREADS) sm.boff[i] = 0;
            if (threadIdx.x == 0) { sm.n_new = 0; deferred[atomicAdd(&scal->n_deferred, 1u)] = p; if (DENSE) *dense.failed = 1u; }
            if (DENSE && wid == 0) ok_dense_lookback(dense.lb, p, 0ull);
            __syncthreads();
            continue;
        }
        // ---- (2) exclusive scan of the bucket counts, BPT per thread
        const unsigned tot = sm.n_new;
        if (DENSE && threadIdx.x == 0 && p) atomicExch(&dense.lb[p], OK_LB_AGG | tot);      // early: successors can already add it up
        unsigned hh[C::BPT], hsum = 0, hmax = 0;
#pragma unroll
        for (unsigned q = 0; q < C::BPT; ++q) { hh[q] = sm.boff[C::BPT * threadIdx.x + q]; hsum += hh[q]; hmax = max(hmax, hh[q]); }
        unsigned inc = hsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) sm.wsum[wid] = inc;
        const int crowded = __syncthreads_or(hmax > OK_C2_BUCKET_MAX);
        if (wid == 0) {
            const unsigned w = lane < (int)C::WARPS ? sm.wsum[lane] : 0u;
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, winc, o); if (lane >= o) winc += y; }
            if (lane < (int)C::WARPS) sm.wsum[lane] = winc - w;
        }
        __syncthreads();
        if (!crowded) {
            unsigned excl = sm.wsum[wid] + inc - hsum;
#pragma unroll
            for (unsigned q = 0; q < C::BPT; ++q) { sm.boff[C::BPT * threadIdx.x + q] = excl; excl += hh[q]; }
            __syncthreads();
            // ---- (3) the distinct keys into bucket order (boff[b] ends up as the END of bucket b)
            for (unsigned i = threadIdx.x; i < tot; i += OK_C2_THREADS) {
                const unsigned s = sm.newl[i];
                sm.sidx[atomicAdd(&sm.boff[ok_c2_bucket<LOG2>(sm.tkey[s], cfg, sub_bits)], 1u)] = (unsigned short)s;
            }
            if (DENSE && wid == 0) {          // by now the predecessors have usually published their prefix: one look
                const unsigned long long b = ok_dense_lookback(dense.lb, p, tot);
                if (lane == 0) { *dbase = b; if (b + tot > dense.cap) *dense.failed = 1u; }
            }
            __syncthreads();
            // ---- (4) rank inside the bucket (~1 key per bucket) = final position; emit
            for (unsigned i = threadIdx.x; i < tot; i += OK_C2_THREADS) {
                const unsigned s = sm.sidx[i];
                const unsigned long long key = sm.tkey[s];
                const unsigned b = ok_c2_bucket<LOG2>(key, cfg, sub_bits);
                const unsigned lo = b ? sm.boff[b - 1] : 0u, hi = sm.boff[b];
                unsigned pos = lo;
                for (unsigned j = lo; j < hi; ++j) pos += sm.tkey[sm.sidx[j]] < key ? 1u : 0u;
                if (DENSE) {
                    const unsigned long long at = *dbase + pos;
                    if (at < dense.cap) { dense.keys[at] = key; dense.counts[at] = reinterpret_cast<unsigned short*>(sm.tcnt)[s]; }
                } else {
                    src[b0 + pos] = key;
                    cnt_out[b0 + pos] = reinterpret_cast<unsigned short*>(sm.tcnt)[s];
                }
            }
        } else {                          // keys too clustered for per-bucket ranking: leave it to the generic kernel
            if (threadIdx.x == 0) { deferred[atomicAdd(&scal->n_deferred, 1u)] = p; if (DENSE) *dense.failed = 1u; }
            if (DENSE && wid == 0) ok_dense_lookback(dense.lb, p, 0ull);
        }
        __syncthreads();
        // ---- (5) clean what this sub-partition used
        for (unsigned i = threadIdx.x; i < tot; i += OK_C2_THREADS) {
            const unsigned s = sm.newl[i];
            sm.tkey[s] = OK_EMPTY_KEY;
            reinterpret_cast<unsigned short*>(sm.tcnt)[s] = 0;
        }
#pragma unroll
        for (unsigned q = 0; q < C::BPT; ++q) sm.boff[C::BPT * threadIdx.x + q] = 0;
        if (threadIdx.x == 0) { sm.n_new = 0; if (!crowded) n_distinct[p] = tot; }
        __syncthreads();
    }
}

// Generic kernel for the sub-partitions the fast kernel deferred (any number of windows; keys as clustered as they
// come): HASHED shared-memory table with 32-bit counts, then a bitonic sort of the table itself -- empty slots hold
// the sentinel u64::MAX and sort to the end, so the first n_distinct slots are the sorted run.  Round 1 used a
// MONOTONE table here (home slot = position inside the sub-partition): microsatellite reads put thousands of
// distinct k-mers on one 16-base prefix, i.e. on ONE home slot, the displacement bound spilled every window of
// them and the spill list overflowed (tools/skew.py, 2 % microsatellites).  Hashing does not care where keys sit.
// Only a sub-partition with more than OK_CT_MAXKEYS distinct keys spills (exact, slow path).
#define OK_CT_MAXKEYS 6144u
__global__ void __launch_bounds__(OK_CT_THREADS, 2)
k_part_count_generic(unsigned long long* __restrict__ src, const unsigned* __restrict__ beg,
                     const unsigned* __restrict__ fill_end, const unsigned* __restrict__ cap_end,
                     const unsigned* __restrict__ deferred, const OkPartScalars* __restrict__ scal, OkPartCfg cfg,
                     unsigned* __restrict__ cnt_out, unsigned* __restrict__ n_distinct, OkPartSpill ps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* tkey = reinterpret_cast<unsigned long long*>(smem_raw);   // [OK_CT_SLOTS]
    unsigned* tcnt = reinterpret_cast<unsigned*>(tkey + OK_CT_SLOTS);             // [OK_CT_SLOTS]
    __shared__ unsigned n_keys;
    (void)cfg;
    const unsigned n_def = scal->n_deferred;
    for (unsigned d = scal->def_done + blockIdx.x; d < n_def; d += gridDim.x) {
        const unsigned p = deferred[d];
        const unsigned b0 = beg[p];
        const unsigned e0 = min(fill_end[p], cap_end[p]);
        const unsigned n = e0 > b0 ? e0 - b0 : 0u;
        if (n == 0) { if (threadIdx.x == 0) n_distinct[p] = 0; continue; }
        __syncthreads();
        for (unsigned i = threadIdx.x; i < OK_CT_SLOTS; i += OK_CT_THREADS) { tkey[i] = OK_EMPTY_KEY; tcnt[i] = 0u; }
        if (threadIdx.x == 0) n_keys = 0;
        __syncthreads();
        // ---- insert: CAS claim + add; a key that finds the table full of other keys is spilled
        for (unsigned i = threadIdx.x; i < n; i += OK_CT_THREADS) {
            const unsigned long long key = __ldcs(src + b0 + i);
            if (key == OK_EMPTY_KEY) continue;          // canonical k-mers never equal the sentinel
            unsigned s = ok_c2_hash<13>(key);
            bool placed = false;
            for (unsigned probes = 0; probes < OK_CT_SLOTS; ++probes) {
                unsigned long long cur = tkey[s];
                if (cur == OK_EMPTY_KEY) {
                    if (*(volatile unsigned*)&n_keys >= OK_CT_MAXKEYS) break;      // full enough: the rest of the new keys spill
                    cur = atomicCAS(&tkey[s], OK_EMPTY_KEY, key);
                    if (cur == OK_EMPTY_KEY) { atomicAdd(&n_keys, 1u); cur = key; }
                }
                if (cur == key) { atomicAdd(&tcnt[s], 1u); placed = true; break; }
                s = (s + 1u) & (OK_CT_SLOTS - 1u);
            }
            if (!placed) ok_spill(ps.sp, ps.st, key, 1);
        }
        __syncthreads();
        // ---- bitonic sort of the whole table by key (8192 slots, 16 per thread and stage)
        for (unsigned k2 = 2; k2 <= OK_CT_SLOTS; k2 <<= 1)
            for (unsigned j = k2 >> 1; j > 0; j >>= 1) {
                for (unsigned t = threadIdx.x; t < OK_CT_SLOTS / 2; t += OK_CT_THREADS) {
                    const unsigned i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));       // the lower index of pair t
                    const unsigned l = i | j;
                    const bool up = (i & k2) == 0u;
                    const unsigned long long a = tkey[i], bkey = tkey[l];
                    if ((a > bkey) == up) {
                        tkey[i] = bkey; tkey[l] = a;
                        const unsigned ca = tcnt[i]; tcnt[i] = tcnt[l]; tcnt[l] = ca;
                    }
                }
                __syncthreads();
            }
        const unsigned tot = n_keys;                   // (a spilled key may also sit in the table: the spill list is exact either way)
        for (unsigned i = threadIdx.x; i < tot; i += OK_CT_THREADS) { src[b0 + i] = tkey[i]; cnt_out[b0 + i] = tcnt[i]; }
        if (threadIdx.x == 0) n_distinct[p] = tot;
    }
}

// sub-partition runs -> final arrays; base[p] = exclusive scan of n_distinct (as u64).
// Sub-partitions [p_begin, p_end): the result pipeline compacts and ships the table in slices.
__global__ void __launch_bounds__(256)
k_part_compact(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ counts /* 32-bit, widened here */,
               const unsigned* __restrict__ beg, const unsigned* __restrict__ n_distinct,
               const unsigned long long* __restrict__ base, unsigned p_begin, unsigned p_end,
               unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_counts) {
    for (unsigned p = p_begin + blockIdx.x; p < p_end; p += gridDim.x) {
        const unsigned n = n_distinct[p];
        const unsigned b0 = beg[p];
        const unsigned long long o0 = base[p];
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            out_keys[o0 + i] = keys[b0 + i];
            out_counts[o0 + i] = counts[b0 + i];
        }
    }
}

// n_distinct (u32) -> u64 copy for the scan kernel
__global__ void __launch_bounds__(256) k_widen_u32(const unsigned* __restrict__ a, unsigned long long* __restrict__ b, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// ---------------------------------------------------------------- min_count filter of a run --
__global__ void __launch_bounds__(256)
k_filter_count(const unsigned long long* __restrict__ counts, uint64_t n, uint64_t min_count,
               unsigned long long* __restrict__ tile_counts) {
    __shared__ unsigned wsum[8];
    const uint64_t a = (uint64_t)blockIdx.x * 2048;
    unsigned c = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) { uint64_t i = a + it * 256u + threadIdx.x; c += (i < n && counts[i] >= min_count) ? 1u : 0u; }
    c = (unsigned)ok_warp_sum(c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int i = 0; i < 8; ++i) t += wsum[i]; tile_counts[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(256)
k_filter_write(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ counts, uint64_t n,
               uint64_t min_count, const unsigned long long* __restrict__ tile_base,
               unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_counts) {
    __shared__ unsigned wsum[8];
    __shared__ unsigned long long running;
    const uint64_t a = (uint64_t)blockIdx.x * 2048;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = tile_base[blockIdx.x];
    __syncthreads();
    for (int it = 0; it < 8; ++it) {
        const uint64_t i = a + it * 256u + threadIdx.x;
        unsigned long long key = 0, cnt = 0;
        if (i < n) { key = keys[i]; cnt = counts[i]; }
        const bool keep = i < n && cnt >= min_count;
        const unsigned bal = __ballot_sync(OK_FULL, keep);
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        unsigned woff = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { unsigned w = wsum[q]; woff += q < wid ? w : 0u; tot += w; }
        if (keep) {
            const unsigned long long idx = running + woff + __popc(bal & ((1u << lane) - 1u));
            out_keys[idx] = key; out_counts[idx] = cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
}
