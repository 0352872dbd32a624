// merge.cuh -- merge of two sorted (k-mer, count) runs, counts of equal k-mers summed.
//
// The reference keeps ONE table across all input files (count.rs:48, loop :52-79).  Here every large batch is
// counted on its own by the partitioned path (partition.cuh) and comes out as a sorted, duplicate-free run; a
// later batch's run is merged into the accumulated one on the device: C = A u B, count_C(x) = count_A(x) + count_B(x).
// Both inputs are strictly ascending, so a key occurs at most once per side.
//
//   k_merge_partition   merge-path split of the two runs into tiles of OK_MG_TILE merged elements (ties: A first;
//                       a pair of equal keys is never cut by a tile boundary)
//   k_merge_count       distinct keys per tile (elements - pairs)           -> k_scan_tiles (kernels.cuh)
//   k_merge_write       per tile: rank every element against the other side in shared memory, sum the pairs,
//                       write the survivors at the tile's offset
// HBM stream: keys are read twice, counts once, C is written once: 24 (|A| + |B|) + 16 |C| bytes.
#pragma once
#include "kernels.cuh"

#define OK_MG_TILE 2048u
#define OK_MG_THREADS 256u

// first d merged elements = split[t].x elements of A and .y of B, d = min(t * TILE, nA + nB)
__global__ void __launch_bounds__(256)
k_merge_partition(const unsigned long long* __restrict__ a, uint64_t na, const unsigned long long* __restrict__ b, uint64_t nb,
                  uint64_t n_tiles, ulonglong2* __restrict__ split /* n_tiles + 1 */) {
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t <= n_tiles; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t d = t * OK_MG_TILE < na + nb ? t * OK_MG_TILE : na + nb;
        uint64_t lo = d > nb ? d - nb : 0, hi = d < na ? d : na;      // i = elements taken from A
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (a[mid] <= b[d - 1 - mid]) lo = mid + 1; else hi = mid;   // ties: A first
        }
        uint64_t i = lo, j = d - lo;
        if (i > 0 && j < nb && a[i - 1] == b[j]) ++j;                 // keep the pair (A's copy, B's copy) in one tile
        split[t] = make_ulonglong2(i, j);
    }
}

struct OkMergeSmem {
    unsigned long long sa[OK_MG_TILE + 2], sb[OK_MG_TILE + 2];       // the tile's slices of A and B
    unsigned long long mk[OK_MG_TILE + 2], mc[OK_MG_TILE + 2];       // merged keys / counts (k_merge_write)
    unsigned wsum[8];
    unsigned long long running;
};

__device__ __forceinline__ unsigned ok_mg_lower(const unsigned long long* s, unsigned n, unsigned long long v) {   // #elements < v
    unsigned lo = 0, hi = n;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (s[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
// the same for a thread that asks for ascending v's one after the other: walk on from the previous answer (both sides are
// sorted, so a step or two on average), binary search of the rest after 8 steps.  *at: previous answer, updated.
__device__ __forceinline__ unsigned ok_mg_lower_from(const unsigned long long* s, unsigned n, unsigned long long v, unsigned at) {
    unsigned steps = 0;
    while (at < n && s[at] < v) {
        ++at;
        if (++steps == 8u) return at + ok_mg_lower(s + at, n - at, v);
    }
    return at;
}

__global__ void __launch_bounds__(OK_MG_THREADS)
k_merge_count(const unsigned long long* __restrict__ a, const unsigned long long* __restrict__ b,
              const ulonglong2* __restrict__ split, uint64_t n_tiles, unsigned long long* __restrict__ tile_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OkMergeSmem& sm = *reinterpret_cast<OkMergeSmem*>(smem_raw);
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const ulonglong2 s0 = split[t], s1 = split[t + 1];
        const unsigned ca = (unsigned)(s1.x - s0.x), cb = (unsigned)(s1.y - s0.y);
        __syncthreads();
        for (unsigned i = threadIdx.x; i < ca; i += OK_MG_THREADS) sm.sa[i] = a[s0.x + i];
        __syncthreads();
        unsigned pairs = 0;
        {   // every thread takes consecutive elements of B: one binary search, then a walk
            const unsigned per = (cb + OK_MG_THREADS - 1) / OK_MG_THREADS, i0 = threadIdx.x * per, i1 = min(cb, i0 + per);
            unsigned p = i0 < i1 ? ok_mg_lower(sm.sa, ca, b[s0.y + i0]) : 0u;
            for (unsigned i = i0; i < i1; ++i) {
                const unsigned long long v = b[s0.y + i];
                p = ok_mg_lower_from(sm.sa, ca, v, p);
                pairs += (p < ca && sm.sa[p] == v) ? 1u : 0u;
            }
        }
        pairs = (unsigned)ok_warp_sum(pairs);
        if ((threadIdx.x & 31) == 0) sm.wsum[threadIdx.x >> 5] = pairs;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int w = 0; w < 8; ++w) tot += sm.wsum[w];
            tile_counts[t] = (unsigned long long)(ca + cb - tot);
        }
    }
}

// COUNTS = false: keys only (union of two k-mer sets, db_types.rs:43-48): ac, bc and out_counts are not touched
template <bool COUNTS>
__global__ void __launch_bounds__(OK_MG_THREADS)
k_merge_write(const unsigned long long* __restrict__ a, const unsigned long long* __restrict__ ac,
              const unsigned long long* __restrict__ b, const unsigned long long* __restrict__ bc,
              const ulonglong2* __restrict__ split, uint64_t n_tiles, const unsigned long long* __restrict__ tile_base,
              unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OkMergeSmem& sm = *reinterpret_cast<OkMergeSmem*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const ulonglong2 s0 = split[t], s1 = split[t + 1];
        const unsigned ca = (unsigned)(s1.x - s0.x), cb = (unsigned)(s1.y - s0.y), n = ca + cb;
        __syncthreads();                                       // the previous tile has been written out
        for (unsigned i = threadIdx.x; i < ca; i += OK_MG_THREADS) sm.sa[i] = a[s0.x + i];
        for (unsigned i = threadIdx.x; i < cb; i += OK_MG_THREADS) sm.sb[i] = b[s0.y + i];
        if (threadIdx.x == 0) sm.running = tile_base[t];
        __syncthreads();
        // merged rank of every element: its own index + the elements of the other side before it (ties: A first)
        // (every thread takes consecutive elements of a side: one binary search in the other side, then a walk)
        {
            const unsigned per = (ca + OK_MG_THREADS - 1) / OK_MG_THREADS, i0 = threadIdx.x * per, i1 = min(ca, i0 + per);
            unsigned p = i0 < i1 ? ok_mg_lower(sm.sb, cb, sm.sa[i0]) : 0u;
            for (unsigned i = i0; i < i1; ++i) {
                const unsigned long long v = sm.sa[i];
                p = ok_mg_lower_from(sm.sb, cb, v, p);              // #B elements < v
                sm.mk[i + p] = v; if (COUNTS) sm.mc[i + p] = ac[s0.x + i];
            }
        }
        {
            const unsigned per = (cb + OK_MG_THREADS - 1) / OK_MG_THREADS, i0 = threadIdx.x * per, i1 = min(cb, i0 + per);
            unsigned p = i0 < i1 ? ok_mg_lower(sm.sa, ca, sm.sb[i0]) : 0u;
            for (unsigned i = i0; i < i1; ++i) {
                const unsigned long long v = sm.sb[i];
                p = ok_mg_lower_from(sm.sa, ca, v, p);              // #A elements < v
                const unsigned le = p + ((p < ca && sm.sa[p] == v) ? 1u : 0u);    // #A elements <= v
                sm.mk[i + le] = v; if (COUNTS) sm.mc[i + le] = bc[s0.y + i];
            }
        }
        __syncthreads();
        // an element equal to its predecessor is the B copy of a pair: its count goes to the predecessor
        for (unsigned base = 0; base < n; base += OK_MG_THREADS) {
            const unsigned r = base + threadIdx.x;
            unsigned long long key = 0, cnt = 0;
            bool keep = false;
            if (r < n) {
                key = sm.mk[r];
                keep = r == 0 || sm.mk[r - 1] != key;
                if (COUNTS) cnt = sm.mc[r] + ((r + 1 < n && sm.mk[r + 1] == key) ? sm.mc[r + 1] : 0ull);
            }
            const unsigned bal = __ballot_sync(OK_FULL, keep);
            if (lane == 0) sm.wsum[wid] = __popc(bal);
            __syncthreads();
            unsigned woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const unsigned x = sm.wsum[w]; woff += w < wid ? x : 0u; tot += x; }
            if (keep) {
                const unsigned long long idx = sm.running + woff + __popc(bal & ((1u << lane) - 1u));
                out_keys[idx] = key; if (COUNTS) out_counts[idx] = cnt;
            }
            __syncthreads();
            if (threadIdx.x == 0) sm.running += tot;
            __syncthreads();
        }
    }
}
