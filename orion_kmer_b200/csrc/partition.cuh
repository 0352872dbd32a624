// partition.cuh -- the partitioned count path (one-shot batches).
//
// Measured on B200 (tools/microbench.cu, profiles/): a DRAM-resident count table sustains only
// ~17 G load+RED/s (random 32 B sector traffic, DRAM-activation bound) while shared-memory
// atomics run at ~2.4 T/s.  So large batches are counted without a global table at all:
//
//   k_part_sample         1/16 of the tiles -> histogram of sub-partition ids (sizes the buffers)
//   k_part_scatter_bases  ASCII -> pack -> rolling canonical k-mers -> level-1 partitions
//   k_part_scatter_keys   level-1 partition -> level-2 sub-partitions
//   k_part_count          one CTA per sub-partition: shared-memory table (CAS claim + add),
//                         then an ordered sweep emits the sub-partition already sorted
//   k_part_compact        sub-partition runs -> the final sorted (k-mer, count) arrays
//
// Partitions are ranges of the monotone position x(key) (kmer_math.cuh ok_canon_pos), so
// sub-partition order == key order and the concatenation of the sorted sub-partitions is the
// sorted count table of count.rs:106-119 -- no sort pass.  Both scatters are shared-memory
// multisplits: every thread holds 16 k-mers in registers, ranks them with one smem atomicAdd
// each, the tile is staged in shared memory in bin order and copied out in runs.
#pragma once
#include "kernels.cuh"

#define OK_PART_TILE 4096u        // keys per CTA round of a scatter (256 threads x 16)
#define OK_PART_MAXBINS 1024u     // bins per scatter level
#define OK_CT_SLOTS 8192u         // slots of the shared-memory count table
#define OK_CT_PAD 512u            // tail padding = displacement bound of the smem table
#define OK_CT_THREADS 512u

struct OkPartCfg {
    unsigned key_shift;    // 64 - 2k
    unsigned shard_log2;   // multi-GPU: this rank holds 1/2^shard_log2 of the position space
    unsigned b1, b2;       // bits of the level-1 / level-2 bin id
};

// position of a key inside this rank's slice of the key space (64-bit fraction, monotone)
__device__ __forceinline__ uint64_t ok_part_pos(uint64_t key, const OkPartCfg& c) {
    return ok_canon_pos(key << c.key_shift) << c.shard_log2;
}
__device__ __forceinline__ unsigned ok_part_bin1(uint64_t x, const OkPartCfg& c) { return c.b1 ? (unsigned)(x >> (64 - c.b1)) : 0u; }
__device__ __forceinline__ unsigned ok_part_bin2(uint64_t x, const OkPartCfg& c) { return c.b2 ? (unsigned)((x << c.b1) >> (64 - c.b2)) : 0u; }
__device__ __forceinline__ unsigned ok_part_sub(uint64_t x, const OkPartCfg& c) {
    const unsigned b = c.b1 + c.b2;
    return b ? (unsigned)(x >> (64 - b)) : 0u;
}
template <int LEVEL>
__device__ __forceinline__ unsigned ok_part_bin(uint64_t key, const OkPartCfg& c) {
    const uint64_t x = ok_part_pos(key, c);
    return LEVEL == 1 ? ok_part_bin1(x, c) : ok_part_bin2(x, c);
}

struct OkPartSpill { OkSpill sp; OkDevStats* st; };

// multi-GPU routing: bin b of the scatter is owner rank b and its keys go to that rank's
// receive buffer -- peer memory mapped over NVLink (CUDA IPC), or local memory for b == self
struct OkPeerOut { unsigned long long* p[8]; };

// ----------------------------------------------------------------------------- sampling --
// every `stride`-th warp-tile; hist[sub] += 1 per k-mer (global RED; the sample is small)
template <bool MAP_U>
__global__ void __launch_bounds__(256)
k_part_sample(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
              uint64_t n_rec, uint64_t n_tiles, uint64_t stride, unsigned k, OkPartCfg cfg,
              unsigned* __restrict__ hist) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = warp * stride; t < n_tiles; t += warps * stride)
        ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t, t + 1, t + 1, k, lane,
            [&](uint64_t, uint64_t pc, uint64_t cc, uint32_t okmask) {
                ok_lane_windows(pc, cc, okmask, k, [&](int, uint64_t key) {
                    atomicAdd(&hist[ok_part_sub(ok_part_pos(key, cfg), cfg)], 1u);
                });
            });
}
__global__ void __launch_bounds__(256)
k_part_sample_keys(const unsigned long long* __restrict__ keys, uint64_t n, uint64_t stride, OkPartCfg cfg,
                   unsigned* __restrict__ hist) {
    // sample whole 256-key chunks so the loads stay coalesced
    const uint64_t n_chunks = (n + 255) / 256;
    for (uint64_t c = blockIdx.x * stride; c < n_chunks; c += (uint64_t)gridDim.x * stride) {
        const uint64_t i = c * 256 + threadIdx.x;
        if (i < n) atomicAdd(&hist[ok_part_sub(ok_part_pos(keys[i], cfg), cfg)], 1u);
    }
}

// ------------------------------------------------------------- shared multisplit machinery --
// Slotted staging: the 8192 staging slots are split evenly among the bins of the level, a key's
// rank inside its bin (one shared-memory atomicAdd) is its slot, so one pass stages the round.
// A bin that outgrows its slots in a round sends the excess straight to global memory.
#define OK_STAGE_SLOTS 8192u
struct OkScatterSmem {
    unsigned long long stage[OK_STAGE_SLOTS];
    unsigned hist[OK_PART_MAXBINS];              // keys per bin this round; zero between rounds
    unsigned gbase[OK_PART_MAXBINS];             // global index of the bin's first staged key (buffers < 2^32 keys)
};

__device__ __forceinline__ void ok_part_put(unsigned long long key, unsigned long long dst, unsigned long long end,
                                            unsigned long long* __restrict__ out, const OkPartSpill& ps) {
    if (dst < end) out[dst] = key;
    else ok_spill(ps.sp, ps.st, key, 1);     // past the sampled capacity of the bin: exact, slow path
}

// One multisplit round of the CTA (all 256 threads call it together): thread-held keys
// key[0..15] (bit q of vmask says key[q] exists) -> out, grouped by bin.  sm.hist must be zero on
// entry and is zero again on exit.
template <int LEVEL, bool PEER = false>
__device__ __forceinline__ void ok_multisplit16(OkScatterSmem& sm, const uint64_t (&key)[16], unsigned vmask,
                                                const OkPartCfg& cfg, unsigned bins_log2,
                                                unsigned long long* __restrict__ cursors,
                                                const unsigned long long* __restrict__ bin_end,
                                                unsigned long long* __restrict__ out, const OkPartSpill& ps,
                                                const OkPeerOut* peer = nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned cap_log2 = 13u - bins_log2, cap = 1u << cap_log2, n_bins = 1u << bins_log2;
#pragma unroll
    for (int q = 0; q < 16; ++q)
        if (vmask >> q & 1u) {
            const unsigned b = ok_part_bin<LEVEL>(key[q], cfg);
            const unsigned r = atomicAdd(&sm.hist[b], 1u);
            if (r < cap) sm.stage[(b << cap_log2) + r] = key[q];
            else ok_part_put(key[q], atomicAdd(&cursors[b], 1ull), bin_end[b], PEER ? peer->p[b] : out, ps);
        }
    __syncthreads();
    // copy out.  Warp w owns staging slots [w*1024, (w+1)*1024) = a contiguous range of bins.
    // (1) one global cursor bump per non-empty bin; a bin whose region is full spills its tail here
    const unsigned bins_per_warp = n_bins >= 8 ? n_bins >> 3 : 1u;
    const unsigned wb0 = wid * bins_per_warp;
    for (unsigned i = lane; i < bins_per_warp && wb0 + i < n_bins; i += 32) {
        const unsigned b = wb0 + i;
        unsigned c = sm.hist[b];
        if (c > cap) c = cap;
        if (c) {
            const unsigned long long g = atomicAdd(&cursors[b], (unsigned long long)c), e = bin_end[b];
            if (g + c > e) {
                const unsigned keep = g < e ? (unsigned)(e - g) : 0u;
                for (unsigned r = keep; r < c; ++r) ok_spill(ps.sp, ps.st, sm.stage[(b << cap_log2) + r], 1);
                c = keep;
            }
            sm.gbase[b] = (unsigned)g;
        }
        sm.hist[b] = c;
    }
    __syncwarp();
    // (2) dense walk over the warp's staging slots: consecutive lanes = consecutive slots of a bin
    if (wb0 < n_bins) {
        const unsigned t_end = (wb0 + bins_per_warp) << cap_log2;
        for (unsigned t = (wb0 << cap_log2) + lane; t < t_end; t += 32) {
            const unsigned b = t >> cap_log2, r = t & (cap - 1u);
            if (r < sm.hist[b]) (PEER ? peer->p[b] : out)[(unsigned long long)sm.gbase[b] + r] = sm.stage[t];
        }
    }
    __syncwarp();
    for (unsigned i = lane; i < bins_per_warp && wb0 + i < n_bins; i += 32) sm.hist[wb0 + i] = 0;
    __syncthreads();
}

// ------------------------------------------------------------------ level 1: from the bases --
// The 8 warps of a CTA walk their own runs of tiles in lock step; each warp-tile is split in
// two rounds of 16 window ends per lane, so a round holds <= 4096 k-mers per CTA.
template <bool MAP_U, bool PEER = false>
__global__ void __launch_bounds__(256, 3)
k_part_scatter_bases(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
                     uint64_t n_rec, uint64_t n_tiles, uint64_t tiles_per_warp, unsigned k, OkPartCfg cfg,
                     unsigned long long* __restrict__ cursors, const unsigned long long* __restrict__ bin_end,
                     unsigned long long* __restrict__ out, OkPartSpill ps, unsigned long long* __restrict__ n_keys,
                     const __grid_constant__ OkPeerOut peer_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OkScatterSmem& sm = *reinterpret_cast<OkScatterSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t t0 = warp * tiles_per_warp;
    if ((uint64_t)blockIdx.x * 8 * tiles_per_warp >= n_tiles) return;   // whole CTA idle
    for (unsigned i = threadIdx.x; i < OK_PART_MAXBINS; i += 256) sm.hist[i] = 0;
    __syncthreads();
    unsigned long long my_keys = 0;
    ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t0, t0 + tiles_per_warp, n_tiles, k, lane,
        [&](uint64_t, uint64_t pc, uint64_t cc, uint32_t okmask) {
            OkRoll roll; roll.init(pc, cc, k);
            my_keys += __popc(okmask);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint64_t key[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) key[q] = roll.step(16 * half + q);
                // window end j = 16*half + q lives in okmask bit 31-j; make bit q mean key[q]
                const unsigned vm = __brev(okmask) >> (16 * half) & 0xFFFFu;
                ok_multisplit16<1, PEER>(sm, key, vm, cfg, cfg.b1, cursors, bin_end, out, ps, &peer_out);
            }
        });
    my_keys = ok_warp_sum(my_keys);
    if (lane == 0 && my_keys) atomicAdd(n_keys, my_keys);
}

// --------------------------------------------------------------------- level 2: from keys --
// work item w: keys src[item_off[w] .. +item_n[w]) (<= 4096), all of level-1 bin item_bin[w]
template <int LEVEL>  // LEVEL 1: bin by bin1 (keys arriving from peers); LEVEL 2: bin by bin2 inside a bin1
__global__ void __launch_bounds__(256, 3)
k_part_scatter_keys(const unsigned long long* __restrict__ src, const unsigned long long* __restrict__ item_off,
                    const unsigned* __restrict__ item_n, const unsigned* __restrict__ item_bin, unsigned n_items,
                    OkPartCfg cfg, unsigned long long* __restrict__ cursors,
                    const unsigned long long* __restrict__ bin_end, unsigned long long* __restrict__ out,
                    OkPartSpill ps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OkScatterSmem& sm = *reinterpret_cast<OkScatterSmem*>(smem_raw);
    const unsigned bins_log2 = LEVEL == 1 ? cfg.b1 : cfg.b2;
    for (unsigned i = threadIdx.x; i < OK_PART_MAXBINS; i += 256) sm.hist[i] = 0;
    __syncthreads();
    for (unsigned w = blockIdx.x; w < n_items; w += gridDim.x) {
        const unsigned long long* __restrict__ keys = src + item_off[w];
        const unsigned n = item_n[w];
        const unsigned bin_base = LEVEL == 1 ? 0u : item_bin[w] << cfg.b2;
        uint64_t key[16]; unsigned vm = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const unsigned i = q * 256 + threadIdx.x;
            key[q] = 0;
            if (i < n) { key[q] = __ldcs(keys + i); vm |= 1u << q; }
        }
        ok_multisplit16<LEVEL>(sm, key, vm, cfg, bins_log2, cursors + bin_base, bin_end + bin_base, out, ps);
    }
}

// ------------------------------------------------------- count one sub-partition in smem --
// sub-partition p holds keys src[beg[p] .. fill_end[p]).  Its distinct keys come out sorted in
// place (keys -> src[beg[p] ..], counts -> cnt_out[beg[p] ..]); n_distinct[p] says how many.
__device__ __forceinline__ unsigned ok_ct_home(uint64_t key, const OkPartCfg& cfg, unsigned sub_bits) {
    const uint64_t f = ok_part_pos(key, cfg) << sub_bits;     // position inside the sub-partition
    return (unsigned)(((f >> 32) * (uint64_t)OK_CT_SLOTS) >> 32);
}

__global__ void __launch_bounds__(OK_CT_THREADS, 2)
k_part_count(unsigned long long* __restrict__ src, const unsigned long long* __restrict__ beg,
             const unsigned long long* __restrict__ fill_end /* cursor after the scatter */,
             const unsigned long long* __restrict__ cap_end, unsigned n_sub, OkPartCfg cfg,
             unsigned long long* __restrict__ cnt_out, unsigned* __restrict__ n_distinct, OkPartSpill ps) {
    constexpr unsigned NT = OK_CT_SLOTS + OK_CT_PAD;            // 8704
    constexpr unsigned ROUNDS = NT / OK_CT_THREADS;             // 17 strided rounds in the sweep
    constexpr unsigned NW = OK_CT_THREADS / 32;                 // 16 warps
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* tkey = reinterpret_cast<unsigned long long*>(smem_raw);   // [NT]
    unsigned* tcnt = reinterpret_cast<unsigned*>(tkey + NT);                      // [NT]
    __shared__ unsigned seg[ROUNDS * NW + 1];       // occupied slots per (round, warp), then exclusive scan
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned sub_bits = cfg.b1 + cfg.b2;
    for (unsigned p = blockIdx.x; p < n_sub; p += gridDim.x) {
        const unsigned long long b0 = beg[p];
        unsigned long long e0 = fill_end[p];
        if (e0 > cap_end[p]) e0 = cap_end[p];          // the rest was spilled by the scatter
        const unsigned n = (unsigned)(e0 - b0);
        if (n == 0) { if (threadIdx.x == 0) n_distinct[p] = 0; continue; }
        {   // 128-bit stores: two keys / four counts at a time
            ulonglong2* k2 = reinterpret_cast<ulonglong2*>(tkey);
            uint4* c4 = reinterpret_cast<uint4*>(tcnt);
            for (unsigned i = threadIdx.x; i < NT / 2; i += OK_CT_THREADS) k2[i] = make_ulonglong2(OK_EMPTY_KEY, OK_EMPTY_KEY);
            for (unsigned i = threadIdx.x; i < NT / 4; i += OK_CT_THREADS) c4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        // ---- insert: CAS claim + add, 4 keys in flight per thread
        for (unsigned base = 0; base < n; base += 4 * OK_CT_THREADS) {
            unsigned long long kk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned i = base + q * OK_CT_THREADS + threadIdx.x;
                kk[q] = i < n ? __ldcs(src + b0 + i) : OK_EMPTY_KEY;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned long long key = kk[q];
                if (key == OK_EMPTY_KEY) continue;      // canonical k-mers never equal the sentinel
                const unsigned h = ok_ct_home(key, cfg, sub_bits);
                bool placed = false;
                for (unsigned s = h; s < h + OK_CT_PAD; ++s) {
                    unsigned long long cur = tkey[s];
                    if (cur == OK_EMPTY_KEY) {
                        cur = atomicCAS(&tkey[s], OK_EMPTY_KEY, key);
                        if (cur == OK_EMPTY_KEY) cur = key;
                    }
                    if (cur == key) { atomicAdd(&tcnt[s], 1u); placed = true; break; }
                }
                if (!placed) ok_spill(ps.sp, ps.st, key, 1);
            }
        }
        __syncthreads();
        // ---- ordered sweep, round r covers slots [r*512, (r+1)*512), one per thread
        unsigned occ = 0;                               // bit r: my slot of round r is occupied
        unsigned long long pre_lo = 0, pre_hi = 0;      // 5 bits per round: occupied slots of my warp segment before mine
#pragma unroll
        for (unsigned r = 0; r < ROUNDS; ++r) {
            const bool o = tkey[r * OK_CT_THREADS + threadIdx.x] != OK_EMPTY_KEY;
            const unsigned bal = __ballot_sync(OK_FULL, o);
            occ |= (o ? 1u : 0u) << r;
            const unsigned long long before = __popc(bal & ((1u << lane) - 1u));
            if (r < 12) pre_lo |= before << (5 * r); else pre_hi |= before << (5 * (r - 12));
            if (lane == 0) seg[r * NW + wid] = __popc(bal);
        }
        __syncthreads();
        if (wid == 0) {                                 // exclusive scan of the 272 segment counts
            unsigned carry = 0;
            for (unsigned i0 = 0; i0 < ROUNDS * NW; i0 += 32) {
                const unsigned i = i0 + lane;
                const unsigned v = i < ROUNDS * NW ? seg[i] : 0u;
                unsigned inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
                if (i < ROUNDS * NW) seg[i] = carry + inc - v;
                carry += __shfl_sync(OK_FULL, inc, 31);
            }
            if (lane == 0) seg[ROUNDS * NW] = carry;
        }
        __syncthreads();
        const unsigned tot = seg[ROUNDS * NW];
        // second pass over MY occupied slots only (lanes pack their work, sparse tables cost little)
        while (occ) {
            const unsigned r = __ffs(occ) - 1; occ &= occ - 1;
            const unsigned s = r * OK_CT_THREADS + threadIdx.x;
            const unsigned long long key = tkey[s];
            const unsigned before = (unsigned)((r < 12 ? pre_lo >> (5 * r) : pre_hi >> (5 * (r - 12))) & 31u);
            int adj = 0;
            const bool lone = (s == 0 || tkey[s - 1] == OK_EMPTY_KEY) && (s + 1 >= NT || tkey[s + 1] == OK_EMPTY_KEY);
            if (!lone) {
                const unsigned h = ok_ct_home(key, cfg, sub_bits);
                for (unsigned t = h; t < s; ++t) adj -= tkey[t] > key ? 1 : 0;          // parked before us, larger
                for (unsigned t = s + 1; t < h + OK_CT_PAD; ++t) {                       // pushed past us, smaller
                    const unsigned long long kt = tkey[t];
                    if (kt == OK_EMPTY_KEY) break;
                    adj += kt < key ? 1 : 0;
                }
            }
            const unsigned idx = seg[r * NW + wid] + before + adj;
            src[b0 + idx] = key;                       // idx < tot <= n: stays inside the region
            cnt_out[b0 + idx] = tcnt[s];
        }
        if (threadIdx.x == 0) n_distinct[p] = tot;
        __syncthreads();
    }
}

// sub-partition runs -> final arrays; base[p] = exclusive scan of n_distinct (as u64)
__global__ void __launch_bounds__(256)
k_part_compact(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ counts,
               const unsigned long long* __restrict__ beg, const unsigned* __restrict__ n_distinct,
               const unsigned long long* __restrict__ base, unsigned n_sub,
               unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_counts) {
    for (unsigned p = blockIdx.x; p < n_sub; p += gridDim.x) {
        const unsigned n = n_distinct[p];
        const unsigned long long b0 = beg[p], o0 = base[p];
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
