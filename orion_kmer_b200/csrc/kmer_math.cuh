// kmer_math.cuh -- k-mer arithmetic shared by the sm_100a kernels and the host-side checks.
//
// What it replaces in the reference (paths relative to orion-kmer/):
//   dna_base_to_u64            src/kmer.rs:12-20   -> ok_pack4 / ok_pack32 (SWAR, 4 bases per step)
//   seq_to_u64                 src/kmer.rs:37-57   -> rolling update inside ok_lane_windows
//   reverse_complement_u64     src/kmer.rs:79-94   -> ok_revcomp (bit-reverse form) + rolling update
//   canonical_u64              src/kmer.rs:99-106  -> min(fwd, rc)
//   seq.windows(k) + "skip windows holding a non-ACGT byte"  count.rs:28-36 -> validity masks
//
// Layout: a group is 32 consecutive bases.  Its codes are one uint64 with base i in bits
// [63-2i, 62-2i] (first base most significant, the same orientation seq_to_u64 uses, so a
// k-mer is a plain bit-field of the stream); its validity and record-start flags are uint32
// words with base i in bit 31-i.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OK_HD __host__ __device__ __forceinline__
#else
#define OK_HD inline
#endif

#define OK_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

OK_HD uint64_t ok_mask_k(unsigned k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

OK_HD uint64_t ok_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// Reverse complement of a full 32-mer: complement every base (~), reverse the 2-bit groups.
OK_HD uint64_t ok_revcomp32(uint64_t v) {
    uint64_t x = ~v;
#if defined(__CUDA_ARCH__)
    x = __brevll(x);  // reverses single bits; put the two bits of each base back in order
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
#else
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = __builtin_bswap64(x);
#endif
    return x;
}
// src/kmer.rs:79-94 in closed form (k in 1..=32, v < 4^k): v sits in the low 2k bits, so its
// reversed, complemented bases come out in the TOP 2k bits of revcomp32(v).
OK_HD uint64_t ok_revcomp(uint64_t v, unsigned k) { return ok_revcomp32(v) >> (64 - 2 * k); }

OK_HD uint64_t ok_canonical(uint64_t v, unsigned k) {
    uint64_t rc = ok_revcomp(v, k);
    return v < rc ? v : rc;
}

// ---------------------------------------------------------------------------------------
// ASCII -> 2-bit, four bases per step.  x holds 4 ASCII bytes, lowest address in the low byte.
// Returns the 8 code bits (first base in bits 7:6) and 4 validity bits (first base in bit 3).
// Valid bytes: A C G T a c g t, plus U u (read as T) when map_u (OK_NORM_NORMALIZED).
template <bool MAP_U>
OK_HD void ok_pack4(uint32_t x, uint32_t& code8, uint32_t& valid4) {
    // (c>>1)&3 : A->0 C->1 T->2 G->3 (U->2); xor with its own high bit swaps 2<->3.
    uint32_t r = (x >> 1) & 0x03030303u;
    uint32_t c = r ^ ((r >> 1) & 0x01010101u);
    code8 = (c * 0x40100401u) >> 24;  // gather the four 2-bit fields, first base highest
    // expected upper-case letter for each code, fetched with one byte permute
    uint32_t t = c | (c >> 4);
    uint32_t sel = (t & 0x0033u) | ((t >> 8) & 0x3300u);
#if defined(__CUDA_ARCH__)
    uint32_t want = __byte_perm(0x54474341u, 0u, sel);  // "ACGT"
#else
    uint32_t want = 0;
    for (int i = 0; i < 4; ++i) want |= ((0x54474341u >> (8 * ((sel >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
#endif
    uint32_t u = x & 0xDFDFDFDFu;                                   // fold lower case
    uint32_t z = u ^ want;                                          // zero byte <=> valid
    uint32_t nz = (((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z);          // bit 7 of a byte set <=> byte != 0
    uint32_t ok = ~nz;
    if (MAP_U) {
        uint32_t zu = u ^ 0x55555555u;                              // 'U'
        uint32_t nzu = (((zu & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | zu);
        ok |= ~nzu;
    }
    ok &= 0x80808080u;
    valid4 = ((ok >> 7) * 0x80402010u) >> 28;
}

// 32 ASCII bytes (eight little-endian words, w[0] = first four bases) -> one group
template <bool MAP_U>
OK_HD void ok_pack32(const uint32_t w[8], uint64_t& codes, uint32_t& valid) {
    uint64_t c = 0; uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t c8, v4;
        ok_pack4<MAP_U>(w[i], c8, v4);
        c |= (uint64_t)c8 << (56 - 8 * i);
        v |= v4 << (28 - 4 * i);
    }
    codes = c; valid = v;
}

// AND (or OR) of m >> i for i in [0, n), n <= 32, by run doubling.
OK_HD uint64_t ok_run_and(uint64_t m, unsigned n) {
    uint64_t pw = m, res = ~0ull; unsigned p = 1, off = 0;
    while (n) { if (n & 1u) { res &= pw >> off; off += p; } n >>= 1; pw &= pw >> p; p <<= 1; }
    return res;
}
OK_HD uint64_t ok_run_or(uint64_t m, unsigned n) {
    uint64_t pw = m, res = 0ull; unsigned p = 1, off = 0;
    while (n) { if (n & 1u) { res |= pw >> off; off += p; } n >>= 1; pw |= pw >> p; p <<= 1; }
    return res;
}

// Bit 31-j of the result is set when the window ENDING at base j of the current group is
// countable: all k bases valid (count.rs:29 seq_to_u64 == Some) and the window does not
// cross a record start (count.rs:28 windows() never spans records).
OK_HD uint32_t ok_window_mask(uint32_t prev_valid, uint32_t cur_valid, uint32_t prev_start,
                              uint32_t cur_start, unsigned k) {
    uint64_t v64 = ((uint64_t)prev_valid << 32) | cur_valid;
    uint64_t s64 = ((uint64_t)prev_start << 32) | cur_start;
    return (uint32_t)ok_run_and(v64, k) & ~(uint32_t)ok_run_or(s64, k - 1);
}

// Rolling canonical k-mers for the 32 windows that end in the current group.
// emit(j, canonical) is called for every countable window (j = index of its last base).
template <class Emit>
OK_HD void ok_lane_windows(uint64_t prev_codes, uint64_t cur_codes, uint32_t okmask, unsigned k,
                           Emit&& emit) {
    if (okmask == 0) return;
    const uint64_t mask = ok_mask_k(k);
    const unsigned hs = 2 * (k - 1);
    uint64_t fwd = prev_codes & mask;   // the k bases that end at the previous group's last base
    uint64_t rc = ok_revcomp(fwd, k);
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
        uint64_t c = (cur_codes >> (62 - 2 * j)) & 3ull;
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((3ull - c) << hs);
        if (okmask & (0x80000000u >> j)) emit(j, fwd < rc ? fwd : rc);
    }
}

// same walk, fully unrolled: j is a compile-time constant inside emit (register-array indexing)
template <class Emit>
OK_HD void ok_lane_windows_full(uint64_t prev_codes, uint64_t cur_codes, uint32_t okmask, unsigned k,
                                Emit&& emit) {
    if (okmask == 0) return;
    const uint64_t mask = ok_mask_k(k);
    const unsigned hs = 2 * (k - 1);
    uint64_t fwd = prev_codes & mask;
    uint64_t rc = ok_revcomp(fwd, k);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint64_t c = (cur_codes >> (62 - 2 * j)) & 3ull;
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((3ull - c) << hs);
        if (okmask & (0x80000000u >> j)) emit(j, fwd < rc ? fwd : rc);
    }
}

// ---------------------------------------------------------------------------------------
// Slot placement.  MONOTONE maps send a key to a home slot that never decreases with the key,
// so linear probing leaves the table sorted up to short local displacements and the sorted
// count table (count.rs:119 sort_by_key) falls out of one ordered sweep -- no sort pass.
// OK_MAP_CANON straightens the density of canonical k-mers (min of a k-mer and its reverse
// complement has density 2(1-u) over the key space) with its CDF 1-(1-u)^2.
enum { OK_MAP_LINEAR = 0, OK_MAP_CANON = 1, OK_MAP_HASH = 2 };

OK_HD uint64_t ok_mix64(uint64_t x) {  // murmur3 finaliser
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// position of a key as a 64-bit fraction of the key space, straightened by the canonical
// prior: x = 1 - (1-u)^2 evaluated on the top 32 bits of u (one 32x32->64 multiply).  Monotone
// non-decreasing in the key; keys sharing their first 16 bases tie, which the probing and the
// rank rule of the readout absorb.  Every placement decision (table home, partition bins,
// multi-GPU owner) is a prefix of this one number, so they nest consistently.
OK_HD uint64_t ok_canon_pos(uint64_t u) {
    const uint64_t w = (~u) >> 32;
    return ~(w * w);
}

OK_HD uint64_t ok_scale_pos(uint64_t x, uint64_t n) {   // floor(x * n / 2^64), cheap when n < 2^32
    return n < (1ull << 32) ? ((x >> 32) * n) >> 32 : ok_mulhi64(x, n);
}

OK_HD uint64_t ok_home_slot(uint64_t key, unsigned key_shift, int map_mode, uint64_t n_home) {
    uint64_t x;
    if (map_mode == OK_MAP_HASH) x = ok_mix64(key);
    else {
        x = key << key_shift;                 // key as a fraction of the key space
        if (map_mode == OK_MAP_CANON) x = ok_canon_pos(x);
    }
    return ok_scale_pos(x, n_home);
}

// resumable rolling canonical k-mer over one lane's 32 window ends (same arithmetic as
// ok_lane_windows, but the caller drives the loop and may pause between steps)
struct OkRoll {
    uint64_t fwd, rc, cur, mask; unsigned hs;
    OK_HD void init(uint64_t prev_codes, uint64_t cur_codes, unsigned k) {
        mask = ok_mask_k(k); hs = 2 * (k - 1); cur = cur_codes;
        fwd = prev_codes & mask; rc = ok_revcomp(fwd, k);
    }
    OK_HD uint64_t step(int j) {
        const uint64_t c = (cur >> (62 - 2 * j)) & 3ull;
        fwd = ((fwd << 2) | c) & mask;
        rc = (rc >> 2) | ((3ull - c) << hs);
        return fwd < rc ? fwd : rc;
    }
};

// Ordered readout of a monotone table (see kernels.cuh k_readout_write): how far the entry
// in slot s (key, home h) sits from its sorted position among the surviving entries.
//   (a) entries parked in [h, s) with a larger key        -> each moves us one place earlier
//   (b) entries in the occupied run right of s, within the displacement bound, with a
//       smaller key                                        -> each moves us one place later
template <bool FILTER, class LoadSlot>
OK_HD long long ok_rank_adjust(LoadSlot&& ld, uint64_t s, uint64_t key, uint64_t h, unsigned max_probe,
                               uint64_t n_total, uint64_t min_count) {
    long long adj = 0;
    for (uint64_t q = h; q < s; ++q) {
        uint64_t kq, cq; ld(q, kq, cq);
        if (kq > key && (!FILTER || cq >= min_count)) --adj;
    }
    uint64_t lim = h + max_probe;
    if (lim > n_total) lim = n_total;
    for (uint64_t q = s + 1; q < lim; ++q) {
        uint64_t kq, cq; ld(q, kq, cq);
        if (kq == OK_EMPTY_KEY) break;
        if (kq < key && (!FILTER || cq >= min_count)) ++adj;
    }
    return adj;
}
