// setops.cuh -- all-vs-all intersection sizes of many sorted sets in ONE pass over their keys.
//
// compare.rs:51-60 computes |A n B| pair by pair: for n sets that is n (n-1) / 2 merges, every one of them reading
// both sets (BASELINE.json configs[4]: 32,640 pairs of 5 M-key sets = 326 G key steps, 1.6 s on one B200 in the
// row-by-row form of kernels.cuh).  The matrix is a co-occurrence count, though: inter[i][j] = number of keys held by
// both i and j, so it can be accumulated KEY by key -- and a key held by one set only (most keys of diverged genomes)
// contributes nothing.  Every key of every set is read once:
//
//   k_ava_ends     first and last key of every set (the host derives the tile geometry from them)
//   k_ava_bounds   the key space is cut into tiles of ~3072 keys (all sets together) by the monotone position
//                  x(key) of kmer_math.cuh; bounds[s][t] = first index of set s inside tile t (one streaming pass:
//                  a thread writes a bound wherever the tile id changes between neighbouring keys)
//   k_ava_tiles    one CTA per tile, persistent: (A) the tile's range of every set goes into a hashed shared-memory
//                  table, key -> number of sets holding it (the ranges of all sets are one flat list of entries, four
//                  loads per thread in flight); (B) keys held by >= 2 sets get dense ids (block scan); (C) bit `id` of
//                  row s is set for every shared key of set s (from the entries' table slots, kept in shared memory);
//                  (D) thread (bi, bj) owns an 8 x 8 block of the pair matrix and adds popc(row_i & row_j) word by
//                  word into 64 registers, kept across all tiles of the CTA and flushed once at the end.
//
// Integer work throughout (AND + POPC on bit rows, shared-memory atomics): nothing for the tensor cores.  A tile
// whose keys do not fit the table (clustered keys: one 16-base prefix shared by thousands of keys) fails the pass as
// a whole and the caller computes the matrix row by row instead (kernels.cuh) -- same integers, slower.
#pragma once
#include "kernels.cuh"

#define OK_AVA_THREADS 576u        // 18 warps; 528 threads own an 8 x 8 block of the upper triangle of a 256 x 256 matrix
#define OK_AVA_SLOTS 8192u         // hashed table of a tile
#define OK_AVA_MAX_ENTRIES 6144u   // keys of all sets in one tile, at most (table load <= 0.75 even if all are distinct)
#define OK_AVA_ROUND_IDS 1024u     // shared keys per round of (C) + (D)
#define OK_AVA_MAX_SETS 256u
#define OK_AVA_TARGET 3072u        // keys per tile the host aims for

struct OkAvaGeo {
    uint32_t phi_lo;       // smallest 32-bit position of any key
    unsigned key_shift;    // 64 - 2k
    uint64_t scale;        // tile = ((phi - phi_lo) * scale) >> 32, scale = floor(n_tiles 2^32 / span)
    unsigned n_tiles;
};

// top 32 bits of the monotone position of a key (the same number every partition of the count path is a prefix of)
OK_HD uint32_t ok_phi32(uint64_t key, unsigned key_shift) { return (uint32_t)(ok_canon_pos(key << key_shift) >> 32); }
OK_HD unsigned ok_ava_tile(uint64_t key, const OkAvaGeo& g) {
    return (unsigned)(((uint64_t)(uint32_t)(ok_phi32(key, g.key_shift) - g.phi_lo) * g.scale) >> 32);
}
OK_HD unsigned ok_ava_hash(uint64_t key) { return (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 51); }   // 13 bits
// thread b of k_ava_tiles owns block (bi, bj), bi <= bj, of the nb x nb grid of 8 x 8 blocks (upper triangle, row-major)
OK_HD bool ok_ava_block(unsigned b, unsigned nb, unsigned& bi, unsigned& bj) {
    bi = 0; bj = 0;
    if (b >= nb * (nb + 1u) / 2u) return false;
    unsigned rem = b;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    bj = bi + rem;
    return true;
}

#if defined(__CUDACC__)
// ends[2 s] = first key, ends[2 s + 1] = last key of set s (untouched for an empty set)
__global__ void __launch_bounds__(256)
k_ava_ends(const unsigned long long* const* __restrict__ keys, const unsigned long long* __restrict__ ns, unsigned n_sets,
           unsigned long long* __restrict__ ends) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_sets && ns[s]) { ends[2 * s] = keys[s][0]; ends[2 * s + 1] = keys[s][ns[s] - 1]; }
}

// bounds[s * (n_tiles + 1) + t] = number of keys of set s in tiles < t (so tile t holds [bounds[t], bounds[t + 1]))
__global__ void __launch_bounds__(256)
k_ava_bounds(const unsigned long long* const* __restrict__ keys, const unsigned long long* __restrict__ ns, OkAvaGeo g,
             unsigned* __restrict__ bounds) {
    const unsigned s = blockIdx.y;
    const unsigned long long* __restrict__ a = keys[s];
    const uint64_t n = ns[s];
    unsigned* __restrict__ b = bounds + (size_t)s * ((size_t)g.n_tiles + 1u);
    if (n == 0) {
        for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t <= g.n_tiles; t += (uint64_t)gridDim.x * blockDim.x) b[t] = 0u;
        return;
    }
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned t1 = ok_ava_tile(a[i], g);
        const unsigned t0 = i ? ok_ava_tile(a[i - 1], g) + 1u : 0u;      // tiles (tile of the key before, t1] begin at i
        for (unsigned t = t0; t <= t1; ++t) b[t] = (unsigned)i;
        if (i + 1 == n) for (unsigned t = t1 + 1u; t <= g.n_tiles; ++t) b[t] = (unsigned)n;
    }
}

struct OkAvaSmem {
    unsigned long long key[OK_AVA_SLOTS];            // 64 KB
    unsigned cnt[OK_AVA_SLOTS];                      // 32 KB: sets holding the key; after (B): id + 1 of a shared key, 0 otherwise
    unsigned col[(OK_AVA_ROUND_IDS / 32u) * OK_AVA_MAX_SETS];   // 32 KB: word w of row s at [w * 256 + s]
    unsigned ent[OK_AVA_MAX_ENTRIES];                // 24 KB: entry e of the tile = table slot | set << 13
    const unsigned long long* ptr[OK_AVA_MAX_SETS];  // first key of this tile's range of every set
    unsigned off[OK_AVA_MAX_SETS + 1u];              // entries of the sets before s (exclusive prefix of the range lengths)
    unsigned wsum[32];
    unsigned n_shared;
};

// (registers are handed out to 4 warps at a time: 18 warps count as 20, which leaves 96 registers per thread)
#ifndef OK_AVA_MLP
#define OK_AVA_MLP 4               // loads a thread has in flight in (A)
#endif
__global__ void __launch_bounds__(OK_AVA_THREADS, 1)
k_ava_tiles(const unsigned long long* const* __restrict__ keys, unsigned n_sets, OkAvaGeo g, const unsigned* __restrict__ bounds,
            unsigned long long* __restrict__ out /* n_sets x n_sets, entries i < j */, unsigned* __restrict__ failed) {
    extern __shared__ __align__(128) unsigned char ava_smem_raw[];
    OkAvaSmem& sm = *reinterpret_cast<OkAvaSmem*>(ava_smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    constexpr unsigned NW = OK_AVA_THREADS / 32u;
    constexpr unsigned SPT = (OK_AVA_SLOTS + OK_AVA_THREADS - 1u) / OK_AVA_THREADS;      // table slots per thread in (B)
    unsigned bi, bj;
    const bool has_block = ok_ava_block(tid, (n_sets + 7u) / 8u, bi, bj);
    unsigned acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0u;
    const size_t brow = (size_t)g.n_tiles + 1u;

    for (unsigned t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        // ---- the tile's range of every set (thread s: set s), their exclusive prefix; an empty table
        unsigned len = 0u, inc = 0u;
        if (tid < OK_AVA_MAX_SETS) {
            const unsigned long long* p = nullptr;
            if (tid < n_sets) { const unsigned lo = bounds[tid * brow + t]; len = bounds[tid * brow + t + 1u] - lo; p = keys[tid] + lo; }
            sm.ptr[tid] = p;
            inc = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(OK_FULL, inc, o); if ((int)lane >= o) inc += y; }
            if (lane == 31u) sm.wsum[wid] = inc;
        }
        for (unsigned i = tid; i < OK_AVA_SLOTS; i += OK_AVA_THREADS) { sm.key[i] = OK_EMPTY_KEY; sm.cnt[i] = 0u; }
        __syncthreads();
        if (tid < OK_AVA_MAX_SETS) {
            unsigned before = 0u;
            for (unsigned w = 0; w < wid; ++w) before += sm.wsum[w];
            sm.off[tid] = before + inc - len;
            if (tid == OK_AVA_MAX_SETS - 1u) sm.off[OK_AVA_MAX_SETS] = before + inc;
        }
        __syncthreads();
        const unsigned n_entries = sm.off[OK_AVA_MAX_SETS];     // (rewritten after the next tile's first barrier only)
        if (n_entries == 0u) continue;
        if (n_entries > OK_AVA_MAX_ENTRIES) { if (tid == 0) *failed = 1u; continue; }
        // ---- (A) key -> number of sets holding it (a set holds a key once).  Entry e of the tile belongs to the last
        // set whose prefix is <= e; four loads per thread are in flight before the first insert.
        for (unsigned e0 = tid; e0 < n_entries; e0 += OK_AVA_THREADS * OK_AVA_MLP) {
            unsigned long long kk[OK_AVA_MLP];
            unsigned ss[OK_AVA_MLP];
#pragma unroll
            for (int u = 0; u < OK_AVA_MLP; ++u) {
                const unsigned e = e0 + (unsigned)u * OK_AVA_THREADS;
                kk[u] = OK_EMPTY_KEY; ss[u] = 0u;
                if (e < n_entries) {
                    unsigned lo = 0u, hi = OK_AVA_MAX_SETS - 1u;
                    while (lo < hi) { const unsigned mid = (lo + hi + 1u) >> 1; if (sm.off[mid] <= e) lo = mid; else hi = mid - 1u; }
                    ss[u] = lo;
                    kk[u] = __ldg(sm.ptr[lo] + (e - sm.off[lo]));
                }
            }
#pragma unroll
            for (int u = 0; u < OK_AVA_MLP; ++u) {
                const unsigned e = e0 + (unsigned)u * OK_AVA_THREADS;
                if (e < n_entries) {
                    const unsigned long long key = kk[u];
                    unsigned h = ok_ava_hash(key);
                    for (;;) {
                        const unsigned long long cur = atomicCAS(&sm.key[h], OK_EMPTY_KEY, key);
                        if (cur == OK_EMPTY_KEY || cur == key) { atomicAdd(&sm.cnt[h], 1u); break; }
                        h = (h + 1u) & (OK_AVA_SLOTS - 1u);
                    }
                    sm.ent[e] = h | (ss[u] << 13);
                }
            }
        }
        __syncthreads();
        // ---- (B) dense ids for the keys held by two sets or more
        const unsigned s0 = tid * SPT;
        unsigned mine = 0u;
        for (unsigned i = 0; i < SPT; ++i) { const unsigned q = s0 + i; if (q < OK_AVA_SLOTS && sm.cnt[q] >= 2u) ++mine; }
        inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(OK_FULL, inc, o); if ((int)lane >= o) inc += y; }
        if (lane == 31u) sm.wsum[wid] = inc;
        __syncthreads();
        if (wid == 0u) {
            const unsigned w = lane < NW ? sm.wsum[lane] : 0u;
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(OK_FULL, winc, o); if ((int)lane >= o) winc += y; }
            sm.wsum[lane] = winc - w;
            if (lane == 31u) sm.n_shared = winc;
        }
        __syncthreads();
        const unsigned n_shared = sm.n_shared;
        {
            unsigned id = sm.wsum[wid] + inc - mine;
            for (unsigned i = 0; i < SPT; ++i) {
                const unsigned q = s0 + i;
                if (q < OK_AVA_SLOTS) { const unsigned c = sm.cnt[q]; sm.cnt[q] = c >= 2u ? ++id : 0u; }
            }
        }
        __syncthreads();
        // ---- (C) + (D), OK_AVA_ROUND_IDS shared keys at a time (one round, unless most keys of the tile are shared by few sets)
        for (unsigned base = 0u; base < n_shared; base += OK_AVA_ROUND_IDS) {
            const unsigned ids = n_shared - base < OK_AVA_ROUND_IDS ? n_shared - base : OK_AVA_ROUND_IDS;
            const unsigned n_words = (ids + 31u) / 32u;
            for (unsigned i = tid; i < n_words * OK_AVA_MAX_SETS; i += OK_AVA_THREADS) sm.col[i] = 0u;
            __syncthreads();
            for (unsigned e = tid; e < n_entries; e += OK_AVA_THREADS) {
                const unsigned v = sm.ent[e];
                const unsigned id1 = sm.cnt[v & (OK_AVA_SLOTS - 1u)];
                if (id1) {
                    const unsigned r = id1 - 1u - base;                                 // (wraps for ids of other rounds)
                    if (r < ids) atomicOr(&sm.col[(r >> 5) * OK_AVA_MAX_SETS + (v >> 13)], 1u << (r & 31u));
                }
            }
            __syncthreads();
            if (has_block) {
                for (unsigned w = 0u; w < n_words; ++w) {
                    const uint4* __restrict__ rp = reinterpret_cast<const uint4*>(&sm.col[w * OK_AVA_MAX_SETS + 8u * bi]);
                    const uint4* __restrict__ cp = reinterpret_cast<const uint4*>(&sm.col[w * OK_AVA_MAX_SETS + 8u * bj]);
                    const uint4 r0 = rp[0], r1 = rp[1], c0 = cp[0], c1 = cp[1];
                    const unsigned r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    const unsigned c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int x = 0; x < 8; ++x)
#pragma unroll
                        for (int y = 0; y < 8; ++y) acc[x * 8 + y] += __popc(r[x] & c[y]);
                }
            }
            __syncthreads();                               // the rows are cleared again by the next round / the next tile's ranges follow
        }
    }
    // one flush per CTA: entries above the diagonal only (a diagonal block also counted x >= y; dropped here)
    if (has_block) {
#pragma unroll
        for (int x = 0; x < 8; ++x)
#pragma unroll
            for (int y = 0; y < 8; ++y) {
                const unsigned i = 8u * bi + x, j = 8u * bj + y;
                if (i < j && j < n_sets && acc[x * 8 + y]) atomicAdd(&out[(size_t)i * n_sets + j], (unsigned long long)acc[x * 8 + y]);
            }
    }
}
#endif  // __CUDACC__
