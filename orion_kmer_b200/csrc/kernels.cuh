// kernels.cuh -- sm_100a kernels of the k-mer hot path (no tensor-core work: nothing here is
// a dense contraction; the path is HBM / L2-atomic bound integer work).
//
//   k_extract<MAP_U,Sink>   fused ingest: 128-bit loads of ASCII bases -> 2-bit pack (SWAR) ->
//                           warp-cooperative rolling canonical k-mers -> Sink
//                           (replaces count.rs:23-38 process_sequence_chunk and its four copies)
//   SinkCount               open-addressing table: CAS claim + 64-bit RED increment (count.rs:31-34)
//   k_readout_*             ordered sweep of the monotone table = filter + sort (count.rs:106-119)
//   k_keytable_* / k_probe* set membership (query.rs:87-94, classify.rs:230-236, compare.rs:58)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmer_math.cuh"

#define OK_TILE_BASES 1024u          // one warp-tile: 32 lanes x 32 bases
#define OK_FULL 0xFFFFFFFFu

struct OkSlot { unsigned long long key; unsigned long long count; };  // 16 B, one DRAM sector half

struct OkTableView {
    OkSlot* slots;
    uint64_t n_home;      // home slots held here
    uint64_t n_total;     // n_home + tail padding (probing never wraps)
    uint64_t n_home_all;  // home slots of the whole (possibly sharded) table: the map's range
    uint64_t home_base;   // first home slot of this shard within that range
    unsigned key_shift;   // 64 - 2k
    int map_mode;         // OK_MAP_*
    unsigned max_probe;   // displacement bound L; inserts beyond it are spilled, never lost
};

struct OkDevStats {                       // one per counter, device resident
    unsigned long long occupied;          // distinct keys claimed
    unsigned long long windows;           // sum of increments
    unsigned long long spill_n;           // entries appended to the spill list
    unsigned long long max_disp;          // largest displacement from the home slot
    unsigned long long route_counts[8];   // scratch for k_route
};

struct OkSpill { uint64_t* keys; uint64_t* incs; uint64_t cap; };

// ------------------------------------------------------------------------- small helpers --
__device__ __forceinline__ uint4 ok_ld_stream16(const void* p) {  // streaming 128-bit load
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned long long ok_ld_key(const unsigned long long* p) {
    return __ldcg(p);  // L2-coherent: a stale EMPTY is harmless (the CAS arbitrates)
}
__device__ __forceinline__ unsigned long long ok_warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(OK_FULL, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long ok_warp_max(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { unsigned long long w = __shfl_xor_sync(OK_FULL, v, o); v = w > v ? w : v; }
    return v;
}

// ------------------------------------------------------------------------------- fill --
__global__ void __launch_bounds__(256) k_fill_slots(OkSlot* slots, uint64_t n) {
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    uint4* p = reinterpret_cast<uint4*>(slots);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        __stcs(p + i, e);
}
__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// ------------------------------------------------------------------------- table insert --
// DashMap::entry(k).or_insert(0).fetch_add(inc) (count.rs:31-34).  Returns false when the
// key could not be placed within max_probe slots of its home (caller spills it).
__device__ __forceinline__ bool ok_table_add(const OkTableView& t, uint64_t key, uint64_t inc,
                                             unsigned& newkeys, unsigned& maxd) {
    const uint64_t h = ok_home_slot(key, t.key_shift, t.map_mode, t.n_home_all) - t.home_base;
    uint64_t lim = h + t.max_probe;
    if (lim > t.n_total) lim = t.n_total;
    for (uint64_t i = h; i < lim; ++i) {
        unsigned long long cur = ok_ld_key(&t.slots[i].key);
        if (cur == OK_EMPTY_KEY) {
            cur = atomicCAS(&t.slots[i].key, OK_EMPTY_KEY, (unsigned long long)key);
            if (cur == OK_EMPTY_KEY) { ++newkeys; cur = key; }
        }
        if (cur == key) {
            atomicAdd(&t.slots[i].count, (unsigned long long)inc);  // RED.64
            unsigned d = (unsigned)(i - h);
            maxd = d > maxd ? d : maxd;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ void ok_spill(const OkSpill& sp, OkDevStats* st, uint64_t key, uint64_t inc) {
    unsigned long long p = atomicAdd(&st->spill_n, 1ull);
    if (p < sp.cap) { sp.keys[p] = key; sp.incs[p] = inc; }
}

// ------------------------------------------------------------------------------ sinks --
struct SinkCount {            // count.rs:31-34
    OkTableView t; OkDevStats* st; OkSpill sp;
    unsigned newkeys, maxd; unsigned long long windows;
    __device__ __forceinline__ void begin() { newkeys = 0; maxd = 0; windows = 0; }
    __device__ __forceinline__ void group_begin(uint64_t) {}
    __device__ __forceinline__ void operator()(uint64_t /*pos*/, uint64_t key) {
        if (!ok_table_add(t, key, 1, newkeys, maxd)) ok_spill(sp, st, key, 1);
        ++windows;
    }
    __device__ __forceinline__ void group_end() {}
    __device__ __forceinline__ void end() {
        unsigned long long nk = ok_warp_sum(newkeys), w = ok_warp_sum(windows), md = ok_warp_max(maxd);
        if ((threadIdx.x & 31) == 0) {
            if (nk) atomicAdd(&st->occupied, nk);
            if (w) atomicAdd(&st->windows, w);
            if (md) atomicMax(&st->max_disp, md);
        }
    }
};

struct SinkEmit {             // test hook: materialise every canonical k-mer (unordered)
    unsigned long long* out; unsigned long long* n_out; uint64_t cap;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void group_begin(uint64_t) {}
    __device__ __forceinline__ void operator()(uint64_t, uint64_t key) {
        unsigned long long p = atomicAdd(n_out, 1ull);
        if (p < cap) out[p] = key;
    }
    __device__ __forceinline__ void group_end() {}
    __device__ __forceinline__ void end() {}
};

// 8-byte-slot membership table of a sealed set (hashed; probing wraps)
struct OkKeyTableView { const unsigned long long* keys; uint64_t n_slots; int has_max; };

__device__ __forceinline__ bool ok_keytable_contains(const OkKeyTableView& t, uint64_t key) {
    if (key == OK_EMPTY_KEY) return t.has_max != 0;
    uint64_t i = ok_mulhi64(ok_mix64(key), t.n_slots);
    for (;;) {
        unsigned long long cur = __ldg(&t.keys[i]);
        if (cur == key) return true;
        if (cur == OK_EMPTY_KEY) return false;
        if (++i == t.n_slots) i = 0;
    }
}

struct SinkProbeReads {       // query.rs:87-94: windows (not distinct k-mers) found in the set
    OkKeyTableView t; const uint64_t* rec_off; uint64_t n_rec; unsigned* hits;
    uint64_t rec; unsigned acc;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void flush() { if (acc) { atomicAdd(&hits[rec], acc); acc = 0; } }
    // first base of this lane's group: find the record that holds it
    __device__ __forceinline__ void group_begin(uint64_t pos) {
        uint64_t lo = 0, hi = n_rec - 1;  // last r with rec_off[r] <= pos
        while (lo < hi) { uint64_t mid = (lo + hi + 1) >> 1; if (__ldg(&rec_off[mid]) <= pos) lo = mid; else hi = mid - 1; }
        rec = lo; acc = 0;
    }
    __device__ __forceinline__ void operator()(uint64_t pos, uint64_t key) {
        if (!ok_keytable_contains(t, key)) return;
        // the window's record is the one holding its last base
        if (pos >= __ldg(&rec_off[rec + 1])) {
            flush();
            do { ++rec; } while (pos >= __ldg(&rec_off[rec + 1]));
        }
        ++acc;
    }
    __device__ __forceinline__ void group_end() { flush(); }
    __device__ __forceinline__ void end() {}
};

// -------------------------------------------------------------------- fused extraction --
// One warp owns a contiguous run of 1024-base tiles.  Per tile each lane loads its 32 bases
// with two 128-bit loads, packs them to 2 bits + validity in registers, receives the
// previous group from its neighbour by shuffle (the k-1 base halo), builds the 32-bit mask of
// countable windows and rolls forward / reverse-complement k-mers over its 32 positions.
// Record starts come from the offsets array: the warp walks it in step with the tiles.
template <bool MAP_U>
__device__ __forceinline__ void ok_load_group(const uint8_t* __restrict__ bases, uint64_t n_bases,
                                              uint64_t pos, uint64_t& codes, uint32_t& valid) {
    uint32_t w[8];
    if (pos + 32 <= n_bases) {
        uint4 a = ok_ld_stream16(bases + pos), b = ok_ld_stream16(bases + pos + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t x = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint64_t p = pos + 4 * i + b;
                if (p < n_bases) x |= (uint32_t)bases[p] << (8 * b);
            }
            w[i] = x;
        }
    }
    ok_pack32<MAP_U>(w, codes, valid);
}

// start flags of the tile [ws, ws+1024): bit 31-i of lane L's word <=> a record starts at
// ws + 32 L + i.  r_next = first record whose offset is >= ws on entry, >= ws+1024 on exit.
__device__ __forceinline__ uint32_t ok_tile_starts(const uint64_t* __restrict__ rec_off, uint64_t n_rec,
                                                   uint64_t ws, uint64_t& r_next, int lane) {
    uint32_t sw = 0;
    for (;;) {
        uint64_t r = r_next + lane;
        uint64_t o = r < n_rec ? __ldg(&rec_off[r]) : ~0ull;
        bool in = o < ws + OK_TILE_BASES;
        unsigned m = __ballot_sync(OK_FULL, in);
        unsigned rel = (unsigned)(o - ws);
        for (unsigned mm = m; mm; mm &= mm - 1) {
            int src = __ffs(mm) - 1;
            unsigned rr = __shfl_sync(OK_FULL, rel, src);
            if ((int)(rr >> 5) == lane) sw |= 0x80000000u >> (rr & 31u);
        }
        int cnt = __popc(m);
        r_next += cnt;
        if (cnt < 32) break;
    }
    return sw;
}

__device__ __forceinline__ uint64_t ok_lower_bound(const uint64_t* __restrict__ a, uint64_t n, uint64_t v) {
    uint64_t lo = 0, hi = n;  // first index with a[i] >= v
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (__ldg(&a[mid]) < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// The walk shared by every extraction kernel: per_group(pos, prev_codes, cur_codes, okmask) is
// called by ALL 32 lanes once per tile (so it may use warp collectives); pos is the stream
// position of the lane's first base, okmask the countable windows ending in its group.
template <bool MAP_U, class PerGroup>
__device__ __forceinline__ void ok_walk_tiles(const uint8_t* __restrict__ bases, uint64_t n_bases,
                                              const uint64_t* __restrict__ rec_off, uint64_t n_rec,
                                              uint64_t t0, uint64_t t1, uint64_t t_live, unsigned k, int lane,
                                              PerGroup&& per_group, bool halo = true) {
    // tiles >= t_live do not exist: per_group is still called (okmask 0) so that kernels whose
    // warps iterate in lock step keep their barriers aligned
    uint64_t carry_codes = 0; uint32_t carry_valid = 0, carry_start = 0;
    uint64_t r_next = 0;
    if (!halo) {                  // sampling: windows reaching back into the previous tile are simply not seen
        if (t0 > 0 && t0 < t_live) r_next = ok_lower_bound(rec_off, n_rec, t0 * OK_TILE_BASES);
    } else if (t0 > 0 && t0 < t_live) {  // warm-up on the tile before ours: its last group is our first halo
        const uint64_t ws = (t0 - 1) * OK_TILE_BASES;
        r_next = ok_lower_bound(rec_off, n_rec, ws);
        uint64_t c; uint32_t v;
        ok_load_group<MAP_U>(bases, n_bases, ws + 32u * lane, c, v);
        uint32_t s = ok_tile_starts(rec_off, n_rec, ws, r_next, lane);
        carry_codes = __shfl_sync(OK_FULL, c, 31);
        carry_valid = __shfl_sync(OK_FULL, v, 31);
        carry_start = __shfl_sync(OK_FULL, s, 31);
    }
    for (uint64_t t = t0; t < t1; ++t) {
        const uint64_t ws = t * OK_TILE_BASES;
        const uint64_t pos = ws + 32u * lane;
        if (t >= t_live) { per_group(pos, 0ull, 0ull, 0u); continue; }
        uint64_t cur_codes; uint32_t cur_valid;
        ok_load_group<MAP_U>(bases, n_bases, pos, cur_codes, cur_valid);
        // (an L2 prefetch of the warp's next tile here was measured: no effect, 5.78 vs 5.79 ms in the level-1 scatter)
        uint32_t cur_start = ok_tile_starts(rec_off, n_rec, ws, r_next, lane);
        uint64_t prev_codes = __shfl_up_sync(OK_FULL, cur_codes, 1);
        uint32_t prev_valid = __shfl_up_sync(OK_FULL, cur_valid, 1);
        uint32_t prev_start = __shfl_up_sync(OK_FULL, cur_start, 1);
        if (lane == 0) { prev_codes = carry_codes; prev_valid = carry_valid; prev_start = carry_start; }
        carry_codes = __shfl_sync(OK_FULL, cur_codes, 31);
        carry_valid = __shfl_sync(OK_FULL, cur_valid, 31);
        carry_start = __shfl_sync(OK_FULL, cur_start, 31);
        per_group(pos, prev_codes, cur_codes, ok_window_mask(prev_valid, cur_valid, prev_start, cur_start, k));
    }
}

// warp -> its run of tiles; false when the warp has nothing to do
__device__ __forceinline__ bool ok_warp_tiles(uint64_t tile_begin, uint64_t tile_end, uint64_t tiles_per_warp,
                                              uint64_t& t0, uint64_t& t1) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    t0 = tile_begin + warp * tiles_per_warp;
    if (t0 >= tile_end) return false;
    t1 = t0 + tiles_per_warp;
    if (t1 > tile_end) t1 = tile_end;
    return true;
}

template <bool MAP_U, class Sink>
__global__ void __launch_bounds__(256)
k_extract(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
          uint64_t n_rec, uint64_t tile_begin, uint64_t tile_end, uint64_t tiles_per_warp, unsigned k,
          Sink sink_in) {
    Sink sink = sink_in;
    uint64_t t0, t1;
    if (!ok_warp_tiles(tile_begin, tile_end, tiles_per_warp, t0, t1)) return;
    sink.begin();
    ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t0, t1, t1, k, threadIdx.x & 31,
        [&](uint64_t pos, uint64_t prev_codes, uint64_t cur_codes, uint32_t okmask) {
            if (okmask) {
                sink.group_begin(pos);
                ok_lane_windows(prev_codes, cur_codes, okmask, k, [&](int j, uint64_t key) { sink(pos + j, key); });
                sink.group_end();
            }
        });
    sink.end();
}

// ---------------------------------------------------------------- multi-GPU routing kernels --
// Owner of a canonical k-mer = its slice of the (prior-straightened) key space, so every rank
// ends up with a contiguous key range: per-GPU tables are disjoint and the global sorted table
// is the concatenation of the ranks' outputs.  PASS 0 counts k-mers per owner; PASS 1 writes
// them, owner after owner, at the cursors the host derived from pass 0.
template <bool MAP_U, int G, int PASS>
__global__ void __launch_bounds__(256)
k_route(const uint8_t* __restrict__ bases, uint64_t n_bases, const uint64_t* __restrict__ rec_off,
        uint64_t n_rec, uint64_t tile_begin, uint64_t tile_end, uint64_t tiles_per_warp, unsigned k,
        unsigned key_shift, int map_mode, unsigned long long* __restrict__ cursors /*[G]*/,
        unsigned long long* __restrict__ out) {
    uint64_t t0, t1;
    if (!ok_warp_tiles(tile_begin, tile_end, tiles_per_warp, t0, t1)) return;
    const int lane = threadIdx.x & 31;
    unsigned long long tot[G];
#pragma unroll
    for (int r = 0; r < G; ++r) tot[r] = 0;
    ok_walk_tiles<MAP_U>(bases, n_bases, rec_off, n_rec, t0, t1, t1, k, lane,
        [&](uint64_t, uint64_t prev_codes, uint64_t cur_codes, uint32_t okmask) {
            unsigned cnt[G];
#pragma unroll
            for (int r = 0; r < G; ++r) cnt[r] = 0;
            ok_lane_windows(prev_codes, cur_codes, okmask, k, [&](int, uint64_t key) {
                const int o = (int)ok_home_slot(key, key_shift, map_mode, (uint64_t)G);
#pragma unroll
                for (int r = 0; r < G; ++r) cnt[r] += (o == r) ? 1u : 0u;
            });
            if (PASS == 0) {
#pragma unroll
                for (int r = 0; r < G; ++r) tot[r] += cnt[r];
            } else {
                unsigned long long wr[G];   // this lane's next write position per owner
#pragma unroll
                for (int r = 0; r < G; ++r) {
                    unsigned inc = cnt[r];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
                    const unsigned total = __shfl_sync(OK_FULL, inc, 31);
                    unsigned long long base = 0;
                    if (lane == 0 && total) base = atomicAdd(&cursors[r], (unsigned long long)total);
                    base = __shfl_sync(OK_FULL, base, 0);
                    wr[r] = base + inc - cnt[r];
                }
                ok_lane_windows(prev_codes, cur_codes, okmask, k, [&](int, uint64_t key) {
                    const int o = (int)ok_home_slot(key, key_shift, map_mode, (uint64_t)G);
                    unsigned long long p = 0;
#pragma unroll
                    for (int r = 0; r < G; ++r) if (o == r) { p = wr[r]; wr[r] = p + 1; }
                    out[p] = key;
                });
            }
        });
    if (PASS == 0) {
#pragma unroll
        for (int r = 0; r < G; ++r) {
            unsigned long long v = ok_warp_sum(tot[r]);
            if (lane == 0 && v) atomicAdd(&cursors[r], v);
        }
    }
}

// standalone 2-bit packing kernel (subsystem 1 of the north star); one group per thread
template <bool MAP_U>
__global__ void __launch_bounds__(256)
k_pack_2bit(const uint8_t* __restrict__ bases, uint64_t n_bases, uint64_t n_groups,
            unsigned long long* __restrict__ codes, uint32_t* __restrict__ valid) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < n_groups;
         g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t c; uint32_t v;
        ok_load_group<MAP_U>(bases, n_bases, g * 32, c, v);
        codes[g] = c; valid[g] = v;
    }
}

// ---------------------------------------------------- k-mers that arrive already extracted --
// (multi-GPU receive side, table rebuilds, set unions)
__global__ void __launch_bounds__(256)
k_add_kmers(OkTableView t, OkDevStats* st, OkSpill sp, const unsigned long long* __restrict__ keys,
            const unsigned long long* __restrict__ incs, uint64_t n, int count_windows) {
    unsigned newkeys = 0, maxd = 0; unsigned long long windows = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = keys[i], inc = incs ? incs[i] : 1ull;
        if (!ok_table_add(t, key, inc, newkeys, maxd)) ok_spill(sp, st, key, inc);
        windows += count_windows ? inc : 0ull;
    }
    unsigned long long nk = ok_warp_sum(newkeys), w = ok_warp_sum(windows), md = ok_warp_max(maxd);
    if ((threadIdx.x & 31) == 0) {
        if (nk) atomicAdd(&st->occupied, nk);
        if (w) atomicAdd(&st->windows, w);
        if (md) atomicMax(&st->max_disp, md);
    }
}

// move every entry of an old table into a new one (growth).  `windows` is not re-counted.
__global__ void __launch_bounds__(256)
k_rehash(OkTableView dst, OkDevStats* st, OkSpill sp, const OkSlot* __restrict__ src, uint64_t n_src) {
    unsigned newkeys = 0, maxd = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_src;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = ok_ld_stream16(src + i);
        uint64_t key = ((uint64_t)v.y << 32) | v.x, cnt = ((uint64_t)v.w << 32) | v.z;
        if (key == OK_EMPTY_KEY) continue;
        if (!ok_table_add(dst, key, cnt, newkeys, maxd)) ok_spill(sp, st, key, cnt);
    }
    unsigned long long nk = ok_warp_sum(newkeys), md = ok_warp_max(maxd);
    if ((threadIdx.x & 31) == 0) {
        if (nk) atomicAdd(&st->occupied, nk);
        if (md) atomicMax(&st->max_disp, md);
    }
}

// ------------------------------------------------------------------------------ readout --
// count.rs:106-119 (filter count >= min_count, sort ascending by key) as an ordered sweep.
// Because the home slot is monotone in the key and probing only moves forward, an entry e in
// slot s with home h can only be out of order with (a) entries in [h, s) holding a larger key
// and (b) entries in the occupied run right of s, within the displacement bound, holding a
// smaller key.  rank(e) = #survivors before s - #(a) + #(b).
#define OK_RT_SLOTS 2048u   // slots per readout tile (256 threads x 8)

__device__ __forceinline__ void ok_ld_slot(const OkSlot* p, uint64_t& key, uint64_t& cnt) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    key = ((uint64_t)v.y << 32) | v.x; cnt = ((uint64_t)v.w << 32) | v.z;
}

__global__ void __launch_bounds__(256)
k_readout_count(const OkSlot* __restrict__ slots, uint64_t n_total, uint64_t min_count,
                unsigned long long* __restrict__ tile_counts) {
    __shared__ unsigned wsum[8];
    const uint64_t a = (uint64_t)blockIdx.x * OK_RT_SLOTS;
    unsigned c = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        uint64_t s = a + it * 256u + threadIdx.x;
        if (s < n_total) {
            uint64_t key, cnt; ok_ld_slot(slots + s, key, cnt);
            c += (key != OK_EMPTY_KEY && cnt >= min_count) ? 1u : 0u;
        }
    }
    c = (unsigned)ok_warp_sum(c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int i = 0; i < 8; ++i) tot += wsum[i];
        tile_counts[blockIdx.x] = tot;
    }
}

// exclusive scan of tile_counts in place (single CTA, 1024 threads); total -> *total
__global__ void __launch_bounds__(1024)
k_scan_tiles(unsigned long long* __restrict__ v, uint64_t n, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry, chunk_total;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long x = i < n ? v[i] : 0ull;
        unsigned long long inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = wsum[lane];
            unsigned long long winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(OK_FULL, winc, o); if (lane >= o) winc += y; }
            wsum[lane] = winc - w;               // exclusive offset of each warp
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        if (i < n) v[i] = carry + wsum[wid] + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

template <bool FILTER>
__global__ void __launch_bounds__(256)
k_readout_write(OkTableView t, uint64_t min_count, const unsigned long long* __restrict__ tile_base,
                unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_counts) {
    __shared__ unsigned wsum[8];
    __shared__ unsigned long long running;
    const OkSlot* __restrict__ slots = t.slots;
    const uint64_t a = (uint64_t)blockIdx.x * OK_RT_SLOTS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = tile_base[blockIdx.x];
    __syncthreads();
    for (int it = 0; it < 8; ++it) {
        const uint64_t s = a + it * 256u + threadIdx.x;
        uint64_t key = OK_EMPTY_KEY, cnt = 0;
        if (s < t.n_total) ok_ld_slot(slots + s, key, cnt);
        const bool surv = key != OK_EMPTY_KEY && cnt >= min_count;
        const unsigned bal = __ballot_sync(OK_FULL, surv);
        if (lane == 0) wsum[wid] = __popc(bal);
        __syncthreads();
        unsigned woff = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { unsigned w = wsum[i]; woff += i < wid ? w : 0u; tot += w; }
        const unsigned long long before = running + woff + __popc(bal & ((1u << lane) - 1u));
        if (surv) {
            const uint64_t h = ok_home_slot(key, t.key_shift, t.map_mode, t.n_home_all) - t.home_base;
            const long long adj = ok_rank_adjust<FILTER>(
                [&](uint64_t q, uint64_t& kq, uint64_t& cq) { ok_ld_slot(slots + q, kq, cq); },
                s, key, h, t.max_probe, t.n_total, min_count);
            const unsigned long long idx = before + adj;
            out_keys[idx] = key;
            out_counts[idx] = cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) running += tot;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- key tables and probes --
__global__ void __launch_bounds__(256)
k_keytable_build(unsigned long long* __restrict__ slots, uint64_t n_slots,
                 const unsigned long long* __restrict__ keys, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key == OK_EMPTY_KEY) continue;  // remembered in has_max
        uint64_t p = ok_mulhi64(ok_mix64(key), n_slots);
        for (;;) {
            unsigned long long cur = atomicCAS(&slots[p], OK_EMPTY_KEY, key);
            if (cur == OK_EMPTY_KEY || cur == key) break;
            if (++p == n_slots) p = 0;
        }
    }
}

// classify.rs:230-236: matched = |In n R| , depth = sum of input counts over the matches
__global__ void __launch_bounds__(256)
k_probe_counts(OkKeyTableView t, const unsigned long long* __restrict__ keys,
               const unsigned long long* __restrict__ counts, uint64_t n,
               unsigned long long* __restrict__ out /* [0]=matched [1]=depth */) {
    unsigned long long m = 0, d = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        if (ok_keytable_contains(t, keys[i])) { ++m; d += counts ? counts[i] : 0ull; }
    m = ok_warp_sum(m); d = ok_warp_sum(d);
    if ((threadIdx.x & 31) == 0) { if (m) atomicAdd(&out[0], m); if (d) atomicAdd(&out[1], d); }
}

// compare.rs:58 |A n B| for two sorted duplicate-free arrays, tile by tile.  A is cut into tiles of OK_IS_TILE
// keys; k_intersect_bounds finds, for every tile, where its first key would go in B (one binary search per
// tile, all tiles at once), so tile t can only match B[lo[t], lo[t+1]).  k_intersect_tiled streams that range
// through shared memory in chunks and lets every A key binary-search the chunk it falls into: each key of A and
// of B is read from global memory once (the per-key search of k_intersect_sorted makes ~23 dependent global
// loads per key of A: measured 13 s for the 32,640 pairs of config 5).
#define OK_IS_TILE 2048u
// A thread holds OK_IS_TILE / 256 CONSECUTIVE keys of A (ascending): one binary search places the first in the chunk of
// B, the others walk on from there -- on average one step per key for sets of similar size -- and fall back to a
// binary search of the rest when a gap is long.  (One binary search per key: 11 dependent shared-memory loads each;
// measured 61 us per pair of 5 M-key sets against 49 us with the rows batched and ... with the walk.)
__device__ __forceinline__ unsigned ok_is_match_run(const unsigned long long* __restrict__ sb, unsigned cn,
                                                    const unsigned long long (&ka)[OK_IS_TILE / 256], unsigned n_valid) {
    unsigned m = 0;
    if (n_valid == 0 || ka[n_valid - 1] < sb[0] || ka[0] > sb[cn - 1]) return 0u;
    unsigned l = 0, h = cn;
    { const unsigned long long key = ka[0]; while (l < h) { const unsigned mid = (l + h) >> 1; if (sb[mid] < key) l = mid + 1; else h = mid; } }
#pragma unroll
    for (unsigned q = 0; q < OK_IS_TILE / 256; ++q) {
        if (q >= n_valid) break;
        const unsigned long long key = ka[q];
        unsigned steps = 0;
        while (l < cn && sb[l] < key) {
            ++l;
            if (++steps == 8u) {        // a long gap: binary search of what is left
                unsigned lo = l, hi = cn;
                while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (sb[mid] < key) lo = mid + 1; else hi = mid; }
                l = lo;
                break;
            }
        }
        if (l >= cn) break;
        m += sb[l] == key ? 1u : 0u;
    }
    return m;
}

__global__ void __launch_bounds__(256)
k_intersect_bounds(const unsigned long long* __restrict__ a, uint64_t na, const unsigned long long* __restrict__ b,
                   uint64_t nb, unsigned long long* __restrict__ lo_out /* n_tiles + 1 */) {
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t <= n_tiles; t += (uint64_t)gridDim.x * blockDim.x) {
        if (t == n_tiles) { lo_out[t] = nb; continue; }
        const unsigned long long key = a[t * OK_IS_TILE];
        uint64_t lo = 0, hi = nb;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(&b[mid]) < key) lo = mid + 1; else hi = mid; }
        lo_out[t] = lo;
    }
}
__global__ void __launch_bounds__(256)
k_intersect_tiled(const unsigned long long* __restrict__ a, uint64_t na, const unsigned long long* __restrict__ b,
                  const unsigned long long* __restrict__ tile_lo, unsigned long long* __restrict__ out) {
    __shared__ unsigned long long sb[OK_IS_TILE];
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE;
    unsigned long long m = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t i0 = t * OK_IS_TILE;
        constexpr unsigned KPT = OK_IS_TILE / 256;
        unsigned long long ka[KPT];
        const uint64_t i1 = i0 + (uint64_t)threadIdx.x * KPT;          // this thread's consecutive keys
        const unsigned n_valid = i1 >= na ? 0u : (unsigned)(na - i1 < KPT ? na - i1 : KPT);
#pragma unroll
        for (unsigned q = 0; q < KPT; ++q) ka[q] = q < n_valid ? a[i1 + q] : OK_EMPTY_KEY;
        const uint64_t lo = tile_lo[t], hi = tile_lo[t + 1];
        for (uint64_t c = lo; c < hi; c += OK_IS_TILE) {
            const unsigned cn = (unsigned)(hi - c < OK_IS_TILE ? hi - c : OK_IS_TILE);
            __syncthreads();                               // the previous chunk is no longer being searched
            for (unsigned j = threadIdx.x; j < cn; j += 256u) sb[j] = b[c + j];
            __syncthreads();
            m += ok_is_match_run(sb, cn, ka, n_valid);
        }
    }
    m = ok_warp_sum(m);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, m);
}

// ---- membership by merge (query.rs:87-94 for a large database) ----
// Which keys of the sorted duplicate-free array A occur in the sorted duplicate-free array B?  Same tiling as
// k_intersect_tiled (A tile in registers, its range of B streamed through shared memory, every key of A and B read
// once); the matching keys are appended to `out` in no particular order, *n_out counts them.  ok_probe_reads uses it
// with A = the distinct k-mers of a read batch and B = a set of billions of keys: B is read front to back once
// instead of being probed at random (a 40 GB hashed table answered 0.5 G probes/s: TLB and DRAM-page misses).
__device__ __forceinline__ unsigned ok_is_match_mask(const unsigned long long* __restrict__ sb, unsigned cn,
                                                     const unsigned long long (&ka)[OK_IS_TILE / 256], unsigned n_valid) {
    unsigned hit = 0;
    if (n_valid == 0 || ka[n_valid - 1] < sb[0] || ka[0] > sb[cn - 1]) return 0u;
    unsigned l = 0, h = cn;
    { const unsigned long long key = ka[0]; while (l < h) { const unsigned mid = (l + h) >> 1; if (sb[mid] < key) l = mid + 1; else h = mid; } }
#pragma unroll
    for (unsigned q = 0; q < OK_IS_TILE / 256; ++q) {
        if (q >= n_valid) break;
        const unsigned long long key = ka[q];
        unsigned steps = 0;
        while (l < cn && sb[l] < key) {
            ++l;
            if (++steps == 8u) {        // a long gap: binary search of what is left
                unsigned lo = l, hi = cn;
                while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (sb[mid] < key) lo = mid + 1; else hi = mid; }
                l = lo;
                break;
            }
        }
        if (l >= cn) break;
        if (sb[l] == key) hit |= 1u << q;
    }
    return hit;
}
__global__ void __launch_bounds__(256)
k_member_tiled(const unsigned long long* __restrict__ a, uint64_t na, const unsigned long long* __restrict__ b,
               const unsigned long long* __restrict__ tile_lo, unsigned long long* __restrict__ out,
               unsigned long long* __restrict__ n_out) {
    __shared__ unsigned long long sb[OK_IS_TILE];
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t i0 = t * OK_IS_TILE;
        constexpr unsigned KPT = OK_IS_TILE / 256;
        unsigned long long ka[KPT];
        const uint64_t i1 = i0 + (uint64_t)threadIdx.x * KPT;          // this thread's consecutive keys
        const unsigned n_valid = i1 >= na ? 0u : (unsigned)(na - i1 < KPT ? na - i1 : KPT);
#pragma unroll
        for (unsigned q = 0; q < KPT; ++q) ka[q] = q < n_valid ? a[i1 + q] : OK_EMPTY_KEY;
        const uint64_t lo = tile_lo[t], hi = tile_lo[t + 1];
        unsigned hit = 0;                                              // B is duplicate-free: a key matches in one chunk at most
        for (uint64_t c = lo; c < hi; c += OK_IS_TILE) {
            const unsigned cn = (unsigned)(hi - c < OK_IS_TILE ? hi - c : OK_IS_TILE);
            __syncthreads();                               // the previous chunk is no longer being searched
            for (unsigned j = threadIdx.x; j < cn; j += 256u) sb[j] = b[c + j];
            __syncthreads();
            hit |= ok_is_match_mask(sb, cn, ka, n_valid);
        }
        // one reservation per warp: inclusive scan of the lanes' match counts
        const unsigned cnt = __popc(hit);
        unsigned inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(OK_FULL, inc, o); if (lane >= o) inc += y; }
        const unsigned total = __shfl_sync(OK_FULL, inc, 31);
        unsigned long long base = 0;
        if (lane == 31 && total) base = atomicAdd(n_out, (unsigned long long)total);
        base = __shfl_sync(OK_FULL, base, 31);
        unsigned long long pos = base + inc - cnt;
#pragma unroll
        for (unsigned q = 0; q < KPT; ++q)
            if (hit >> q & 1u) out[pos++] = ka[q];
    }
}

// ---- one ROW of an all-vs-all (compare.rs:51-60 for every pair): set A against many sets B_j in two launches ----
// Per pair the tiled form above costs two launches; the 32,640 pairs of BASELINE.json configs[4] are 65,280 launches
// of ~40 us each on one stream, and every pair re-reads A from DRAM.  Here the work items of a row are
// (j, tile of A): item w = jj * n_tiles + t.  Consecutive items share B_j (its matching range is streamed once,
// contiguously) and A (40 MB at most) stays resident in the 126 MB L2 for the whole row.
struct OkRowSets { const unsigned long long* const* keys; const unsigned long long* n; const unsigned* col; };   // B_j = keys[col[jj]], n[col[jj]]

__global__ void __launch_bounds__(256)
k_intersect_row_bounds(const unsigned long long* __restrict__ a, uint64_t na, OkRowSets sets, unsigned n_cols,
                       unsigned long long* __restrict__ lo_out /* n_cols x (n_tiles + 1) */) {
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE, per = n_tiles + 1;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < per * n_cols; w += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned jj = (unsigned)(w / per);
        const uint64_t t = w - (uint64_t)jj * per;
        const unsigned c = sets.col[jj];
        const unsigned long long* __restrict__ b = sets.keys[c];
        const uint64_t nb = sets.n[c];
        if (t == n_tiles) { lo_out[w] = nb; continue; }
        const unsigned long long key = a[t * OK_IS_TILE];
        uint64_t lo = 0, hi = nb;
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(&b[mid]) < key) lo = mid + 1; else hi = mid; }
        lo_out[w] = lo;
    }
}

// every CTA takes a contiguous share of the row's items (so it stays on one B_j for a while) and flushes its match
// count whenever the column changes: out_row[col] += |A n B_col|
__global__ void __launch_bounds__(256)
k_intersect_row_tiled(const unsigned long long* __restrict__ a, uint64_t na, OkRowSets sets, unsigned n_cols,
                      const unsigned long long* __restrict__ lo_all, unsigned long long* __restrict__ out_row, uint64_t out_stride) {
    __shared__ unsigned long long sb[OK_IS_TILE];
    __shared__ unsigned long long wsum[8];
    const uint64_t n_tiles = (na + OK_IS_TILE - 1) / OK_IS_TILE, per = n_tiles + 1;
    const uint64_t n_items = n_tiles * n_cols;
    const uint64_t share = (n_items + gridDim.x - 1) / gridDim.x;
    const uint64_t w0 = blockIdx.x * share, w1 = w0 + share < n_items ? w0 + share : n_items;
    unsigned long long m = 0;
    unsigned cur_jj = 0xFFFFFFFFu;
    auto flush = [&] {                                   // block-wide: every thread calls it at the same items
        m = ok_warp_sum(m);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long tot = 0;
            for (int i = 0; i < 8; ++i) tot += wsum[i];
            if (tot) atomicAdd(&out_row[(uint64_t)sets.col[cur_jj] * out_stride], tot);
        }
        m = 0;
    };
    for (uint64_t w = w0; w < w1; ++w) {
        const unsigned jj = (unsigned)(w / n_tiles);
        const uint64_t t = w - (uint64_t)jj * n_tiles;
        if (jj != cur_jj) { if (cur_jj != 0xFFFFFFFFu) flush(); cur_jj = jj; }
        const unsigned long long* __restrict__ b = sets.keys[sets.col[jj]];
        const uint64_t i0 = t * OK_IS_TILE;
        constexpr unsigned KPT = OK_IS_TILE / 256;
        unsigned long long ka[KPT];
        const uint64_t i1 = i0 + (uint64_t)threadIdx.x * KPT;          // this thread's consecutive keys
        const unsigned n_valid = i1 >= na ? 0u : (unsigned)(na - i1 < KPT ? na - i1 : KPT);
#pragma unroll
        for (unsigned q = 0; q < KPT; ++q) ka[q] = q < n_valid ? a[i1 + q] : OK_EMPTY_KEY;
        const uint64_t lo = lo_all[(uint64_t)jj * per + t], hi = lo_all[(uint64_t)jj * per + t + 1];
        for (uint64_t c = lo; c < hi; c += OK_IS_TILE) {
            const unsigned cn = (unsigned)(hi - c < OK_IS_TILE ? hi - c : OK_IS_TILE);
            __syncthreads();
            for (unsigned j = threadIdx.x; j < cn; j += 256u) sb[j] = b[c + j];
            __syncthreads();
            m += ok_is_match_run(sb, cn, ka, n_valid);
        }
    }
    if (cur_jj != 0xFFFFFFFFu) flush();
}

// strictly ascending? (a set handed over as a device array)  *bad = 1 otherwise
__global__ void __launch_bounds__(256)
k_check_ascending(const unsigned long long* __restrict__ a, uint64_t n, unsigned* __restrict__ bad) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (a[i] <= a[i - 1]) *bad = 1u;
}

// first index of every owner's key range inside a sorted set: bounds[r] = #keys owned by ranks < r
__global__ void k_set_shard_bounds(const unsigned long long* __restrict__ a, uint64_t n, unsigned key_shift, unsigned n_ranks,
                                   unsigned long long* __restrict__ bounds /* n_ranks + 1 */) {
    const unsigned r = threadIdx.x;
    if (r > n_ranks) return;
    uint64_t lo = 0, hi = n;     // first key whose owner >= r (the owner is monotone in the key)
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (ok_home_slot(a[mid], key_shift, OK_MAP_CANON, (uint64_t)n_ranks) < r) lo = mid + 1; else hi = mid;
    }
    bounds[r] = lo;
}

// compare.rs:58 |A n B| for two sorted duplicate-free arrays: every element of A binary-
// searches B (A is the smaller one).
__global__ void __launch_bounds__(256)
k_intersect_sorted(const unsigned long long* __restrict__ a, uint64_t na,
                   const unsigned long long* __restrict__ b, uint64_t nb,
                   unsigned long long* __restrict__ out) {
    unsigned long long m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < na;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = a[i];
        uint64_t lo = 0, hi = nb;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (__ldg(&b[mid]) < key) lo = mid + 1; else hi = mid; }
        if (lo < nb && __ldg(&b[lo]) == key) ++m;
    }
    m = ok_warp_sum(m);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, m);
}
