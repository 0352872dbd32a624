"""Host-side checks of the arithmetic the kernels run (shared __host__ __device__ code in
orion_kmer_b200/csrc/kmer_math.cuh), against the reference's golden vectors and the oracle.
CPU only -- the device runs of the same code are in test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

import orion_kmer_b200 as ok


def emulate_extract(bases, offsets, k, norm_mode=ok.NORMALIZED):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cap = max(len(bases), 1)
    out = np.zeros(cap, dtype=np.uint64)
    n = C.c_uint64()
    rc = ok.lib().okx_emulate_extract(ok._ptr(bases), len(bases), ok._ptr(offsets), len(offsets) - 1, k, norm_mode,
                                      ok._ptr(out), cap, C.byref(n))
    assert rc == 0
    return out[:n.value]


def emulate_table(keys, k, n_home, max_probe=2048, map_mode=1, min_count=1):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    ok_ = np.zeros(max(len(keys), 1), dtype=np.uint64)
    oc = np.zeros(max(len(keys), 1), dtype=np.uint64)
    n, sp = C.c_uint64(), C.c_uint64()
    rc = ok.lib().okx_emulate_table(ok._ptr(keys), len(keys), k, map_mode, n_home, max_probe, min_count,
                                    ok._ptr(ok_), ok._ptr(oc), C.byref(n), C.byref(sp))
    assert rc == 0, ok.lib().ok_last_error()
    return ok_[:n.value], oc[:n.value], sp.value


# ---- src/kmer.rs vectors through the C ABI's host functions -----------------------------------
def test_kmer_golden_vectors(golden):
    g = golden["kmer"]
    for s, k, v in g["seq_to_u64_valid"]:
        assert ok.seq_to_u64(s.encode(), k) == v
    for s, k in g["seq_to_u64_none"]:
        assert ok.seq_to_u64(s.encode(), k) is None
    for v, k, s in g["u64_to_seq"]:
        assert ok.u64_to_seq(v, k) == s.encode()
    for k in g["panic_k"]:
        with pytest.raises(ok.InvalidKmerSize):
            ok.u64_to_seq(0, k)
        with pytest.raises(ok.InvalidKmerSize):
            ok.reverse_complement_u64(0, k)
    for a, b in g["reverse_complement"]:
        assert ok.reverse_complement_u64(ok.seq_to_u64(a.encode(), len(a)), len(a)) == ok.seq_to_u64(b.encode(), len(b))
    for a, b in g["canonical"]:
        assert ok.canonical_u64(ok.seq_to_u64(a.encode(), len(a)), len(a)) == ok.seq_to_u64(b.encode(), len(b))


def test_every_byte_value_classified_like_the_reference(oracle):
    """dna_base_to_u64 (kmer.rs:12-20) over all 256 byte values, k=1."""
    for b in range(256):
        s = bytes([b])
        assert ok.seq_to_u64(s, 1) == oracle.seq_to_u64(s, 1), b


def test_revcomp_canonical_random(oracle):
    rng = np.random.default_rng(11)
    for k in (1, 2, 3, 15, 16, 21, 31, 32):
        for v in rng.integers(0, 2 ** 63, size=200, dtype=np.uint64):
            v = int(v) * 2 + int(rng.integers(0, 2))
            v &= (1 << (2 * k)) - 1
            assert ok.reverse_complement_u64(v, k) == oracle.reverse_complement_u64(v, k)
            assert ok.canonical_u64(v, k) == oracle.canonical_u64(v, k)


# ---- the extraction walk (pack + window masks + rolling k-mers) ---------------------------------
ALPHABET = np.frombuffer(b"ACGTacgtNnUuRYKM-.*X \n", dtype=np.uint8)


def random_batch(rng, n_bases, mean_len, p_junk=0.02):
    p = np.full(len(ALPHABET), p_junk / (len(ALPHABET) - 4))
    p[:4] = (1 - p_junk) / 4
    bases = rng.choice(ALPHABET, size=n_bases, p=p)
    cuts = np.unique(rng.integers(0, n_bases + 1, size=max(1, n_bases // mean_len)))
    off = np.concatenate([[0], cuts, [n_bases]]).astype(np.uint64)
    # a few empty records
    off = np.sort(np.concatenate([off, off[1:4]]))
    return bases, off


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 15, 16, 17, 21, 31, 32])
@pytest.mark.parametrize("norm", [ok.NORMALIZED, ok.RAW])
def test_emulated_extract_matches_oracle(oracle, k, norm):
    rng = np.random.default_rng(100 + k)
    for n_bases, mean_len in ((0, 1), (1, 1), (31, 7), (32, 40), (33, 5), (1500, 150), (5000, 37), (4096, 4096), (3000, 2)):
        if n_bases == 0:
            bases, off = np.zeros(0, np.uint8), np.zeros(1, np.uint64)
        else:
            bases, off = random_batch(rng, n_bases, mean_len)
        got = emulate_extract(bases, off, k, norm)
        # oracle: NORMALIZED batches are whitespace-free by contract, so strip per record first
        if norm == ok.NORMALIZED:
            recs = [bytes(bases[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
            recs = [r.replace(b" ", b"").replace(b"\n", b"") for r in recs]
            b2, o2 = oracle.batch_from_records(recs)
            got = emulate_extract(b2, o2, k, norm)
            want_k, want_c = oracle.count_batch(k, b2, o2, 1, True)
        else:
            want_k, want_c = oracle.count_batch(k, bases, off, 1, False)
        gk, gc = np.unique(got, return_counts=True)
        assert np.array_equal(gk, want_k), (k, norm, n_bases)
        assert np.array_equal(gc.astype(np.uint64), want_c)


def test_emulated_extract_long_record_with_n_runs(oracle):
    rng = np.random.default_rng(5)
    bases = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000)
    for i in range(5):
        bases[3000 * i + 500:3000 * i + 600] = ord("N")
    bases[::97] |= 0x20  # lower case
    off = np.array([0, 20000], dtype=np.uint64)
    got = emulate_extract(bases, off, 21)
    want_k, want_c = oracle.count_batch(21, bases, off)
    gk, gc = np.unique(got, return_counts=True)
    assert np.array_equal(gk, want_k) and np.array_equal(gc.astype(np.uint64), want_c)


# ---- the ordered table (monotone home slot + rank rule of the readout) -------------------------------
@pytest.mark.parametrize("k,n_home", [(3, 65536), (5, 64), (11, 4096), (21, 3000), (31, 2500), (32, 2500)])
def test_emulated_table_is_sorted_and_exact(oracle, k, n_home):
    rng = np.random.default_rng(k)
    raw = rng.integers(0, 2 ** 63, size=1500, dtype=np.uint64)
    keys = np.array([oracle.canonical_u64(int(v) & ((1 << (2 * k)) - 1), k) for v in raw], dtype=np.uint64)
    keys = np.concatenate([keys, keys[:400], keys[:50]])  # duplicates
    rng.shuffle(keys)
    for min_count in (1, 2, 3):
        gk, gc, spilled = emulate_table(keys, k, n_home, min_count=min_count)
        assert spilled == 0
        wk, wc = np.unique(keys, return_counts=True)
        sel = wc >= min_count
        assert np.array_equal(gk, wk[sel]) and np.array_equal(gc, wc[sel].astype(np.uint64))


def test_emulated_table_clustered_keys_stay_sorted(oracle):
    """keys sharing a long prefix land in one neighbourhood: long displacement runs"""
    k = 31
    rng = np.random.default_rng(9)
    base = int(rng.integers(0, 2 ** 40)) << 20
    keys = (base + rng.integers(0, 2 ** 20, size=600, dtype=np.uint64)).astype(np.uint64)
    keys = np.concatenate([keys, rng.integers(0, 2 ** 60, size=600, dtype=np.uint64)])
    rng.shuffle(keys)
    gk, gc, spilled = emulate_table(keys, k, 4096, map_mode=0)
    assert spilled == 0
    wk, wc = np.unique(keys, return_counts=True)
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc.astype(np.uint64))
    # with a tight displacement bound the overflow is reported, never silently dropped
    gk2, gc2, spilled2 = emulate_table(keys, k, 4096, max_probe=64, map_mode=0)
    assert spilled2 > 0 and int(gc2.sum()) + spilled2 == len(keys)
    assert np.all(np.diff(gk2.astype(object)) > 0)


# ------------------------------------------------ the plan of a one-shot batch (host logic) --
def _plan(n_units, hint=0, k=31):
    out = np.zeros(6, dtype=np.uint32)
    assert ok.lib().okx_plan_bits(n_units, hint, k, ok._ptr(out)) == 0
    return dict(bits=int(out[0]), b1=int(out[1]), b2=int(out[2]), hinted=bool(out[3]), big=bool(out[4]), slices=int(out[5]))


def test_plan_without_hint_sizes_sub_partitions_for_their_windows():
    """no capacity hint: <= 4096 windows per sub-partition (every key could be distinct), at most 18 bits in two
    balanced levels, one level up to 256 bins"""
    for n in (1 << 20, 3_000_000, 24_000_000, 300_000_000, 1_500_000_000):
        p = _plan(n)
        assert not p["hinted"]
        assert p["bits"] == min(18, int(np.ceil(np.log2(n / 4096))))
        assert p["b1"] + p["b2"] == p["bits"] and (p["b2"] == 0) == (p["bits"] <= 8)
        assert abs(p["b1"] - p["b2"]) <= 1 or p["b2"] == 0
    assert _plan(1_500_000_000)["big"] is False           # 5722 windows per sub-partition: the 8192-slot tables
    assert _plan(2_000_000_000)["big"] is True            # 7629: the 16384-slot tables


def test_plan_with_hint_sizes_sub_partitions_for_their_distinct_keys():
    # BASELINE.json configs[1]: 1.5 G bases, 255 M expected distinct k-mers -> 2^16 sub-partitions of ~3.9 K distinct keys
    p = _plan(1_500_000_000, 255_000_000)
    assert p == dict(bits=16, b1=8, b2=8, hinted=True, big=False, slices=16)
    # a hint can only make sub-partitions larger, never smaller than the window-sized plan, and never beyond 24576
    # windows per sub-partition (16-bit counts in the shared-memory tables)
    for n in (3_000_000, 24_000_000, 300_000_000, 1_500_000_000):
        base = _plan(n)["bits"]
        for hint in (1, 1000, n // 100, n // 6, n // 2, n, 10 * n):
            p = _plan(n, hint)
            assert p["bits"] <= base
            assert n / (1 << p["bits"]) <= 24576 or p["bits"] == base
            assert p["hinted"] == (p["bits"] < base) or p["bits"] == 18
            if hint >= n:
                assert p["bits"] == base and not p["hinted"]      # every key distinct: the safe plan
    # the deferred (sliced) result pipeline needs >= 4 slices of whole level-1 bins and whole 1024-entry scan chunks
    assert _plan(3_000_000)["slices"] == 0
    assert _plan(24_000_000)["slices"] == 8
    assert _plan(1_500_000_000)["slices"] == 16


def test_format_counts_threaded_chunks_match_the_oracle_byte_for_byte(oracle):
    """count.rs:127-135 text: above 2^18 lines the formatter sizes chunks first and lets several threads write them at
    their final offsets; k mod 4 != 0 exercises the per-base head before the four-bases-per-lookup body"""
    rng = np.random.default_rng(17)
    for k, n in ((31, 400_000), (21, 300_000), (32, 270_000), (5, 1000), (1, 4)):
        hi = 2 ** (2 * k) if k < 32 else 2 ** 64
        keys = np.unique(rng.integers(0, hi, n, dtype=np.uint64))
        counts = rng.integers(1, 10 ** int(rng.integers(1, 12)), len(keys)).astype(np.uint64)
        counts[:2] = [1, 2 ** 64 - 1]
        text = ok.format_counts(keys, counts, k)
        assert text == oracle.format_counts(keys, counts, k), k
        assert text.count(b"\n") == len(keys)
    assert ok.format_counts(np.zeros(0, np.uint64), np.zeros(0, np.uint64), 31) == b""


# ------------------------------- the strided level-1 gather of a union (host replay of the kernel's index map) --
@pytest.mark.parametrize("n", [1, 2, 3, 15, 16, 17, 4095, 4096, 4097, 32767, 32768, 32769, 100_001, 1_234_567])
def test_strided_gather_reads_every_key_once(n):
    """k_part_scatter_keys<1, false, 0, true> gathers an item from 2048 places of the key array (ok_strided_index).
    Every key index below n must be read by exactly one (item, thread, register), pairs must be 16-byte aligned, and
    the places of one item must be spread over the whole array (that is the point: a sorted run no longer lands in
    one bin)."""
    import ctypes as C
    n_items = C.c_uint64()
    cap = (n // 32768 + 2) * 8 * 4096
    out = np.full(cap, np.iinfo(np.uint64).max, dtype=np.uint64)
    assert ok.lib().okx_strided_order(n, ok._ptr(out), cap, C.byref(n_items)) == 0
    items = n_items.value
    assert items % 8 == 0 and items * 4096 >= n and items * 4096 <= cap
    got = out[:items * 4096]
    valid = got[got != np.iinfo(np.uint64).max]
    assert len(valid) == n and np.array_equal(np.sort(valid), np.arange(n, dtype=np.uint64))
    pairs = got.reshape(-1, 2)
    both = (pairs[:, 0] != np.iinfo(np.uint64).max) & (pairs[:, 1] != np.iinfo(np.uint64).max)
    assert np.all(pairs[both, 0] % 2 == 0) and np.all(pairs[both, 1] == pairs[both, 0] + 1)
    assert not np.any((pairs[:, 0] == np.iinfo(np.uint64).max) & (pairs[:, 1] != np.iinfo(np.uint64).max))
    if n >= 1_000_000:
        first = got[:4096].reshape(-1, 2)[:, 0]
        first = np.sort(first[first != np.iinfo(np.uint64).max])
        assert len(first) > 2000 and np.min(np.diff(first)) >= 16 * (n // 16 // 2048)      # one pair per row of the view


# ------------------------------------------- keyed all-vs-all (setops.cuh): tile geometry and block ownership --
def _phi32(key, k):
    u = (int(key) << (64 - 2 * k)) & (2 ** 64 - 1)
    w = ((~u) & (2 ** 64 - 1)) >> 32
    return ((~(w * w)) & (2 ** 64 - 1)) >> 32


@pytest.mark.parametrize("k,n_keys,target_sets", [(21, 200_000, 5), (31, 50_000, 3), (5, 600, 2), (32, 100_000, 4), (21, 4000, 1)])
def test_ava_tiles_are_monotone_and_in_range(k, n_keys, target_sets):
    """tile(key) = ((phi32(key) - phi_lo) * scale) >> 32 must be monotone in the key (the bounds kernel writes a bound
    where the tile id changes between neighbours), below n_tiles for every key of every set, and the geometry must
    follow from the sets' end keys alone.  A narrow key range (a multi-GPU shard) still spreads over all tiles."""
    rng = np.random.default_rng(k * 1000 + n_keys)
    hi = 2 ** (2 * k) if k < 32 else 2 ** 64
    for lo_frac, hi_frac in ((0.0, 1.0), (0.25, 0.375)):
        sets = []
        for _ in range(target_sets):
            a, b = (rng.integers(int(hi * lo_frac), max(int(hi * lo_frac) + 1, int(hi * hi_frac) - 1), size=n_keys, dtype=np.uint64, endpoint=True)
                    for _ in range(2))
            ks = np.unique(np.minimum(a, b) if lo_frac == 0.0 else a)      # canonical k-mers: the smaller of x and rc(x), density 2 (1 - u)
            sets.append(ks)
        sets.append(np.zeros(0, np.uint64))                  # an empty set takes no part in the geometry
        ns = np.array([len(s) for s in sets], np.uint64)
        ends = np.zeros(2 * len(sets), np.uint64)
        for i, s in enumerate(sets):
            if len(s):
                ends[2 * i], ends[2 * i + 1] = s[0], s[-1]
        allk = np.unique(np.concatenate(sets))
        total = int(ns.sum())
        geo = np.zeros(3, np.uint64)
        tiles = np.zeros(len(allk), np.uint32)
        assert ok.lib().okx_ava_geometry(k, ok._ptr(ends), ok._ptr(ns), len(sets), total, ok._ptr(allk), len(allk), ok._ptr(geo), ok._ptr(tiles)) == 0
        n_tiles, phi_lo, scale = int(geo[0]), int(geo[1]), int(geo[2])
        assert 1 <= n_tiles <= max(1, total // 3072) and n_tiles <= 1 << 22
        assert np.all(np.diff(tiles.astype(np.int64)) >= 0) and int(tiles.max()) < n_tiles and int(tiles.min()) == 0
        for i in rng.integers(0, len(allk), size=200):
            assert int(tiles[i]) == ((_phi32(allk[i], k) - phi_lo) * scale) >> 32
        if n_tiles >= 8:
            assert int(tiles.max()) >= n_tiles - 2         # the last tiles are used: the range was rescaled, not clipped
            fill = np.bincount(tiles, minlength=n_tiles)
            assert fill.max() <= 6144                       # uniform keys: no tile near the table's limit


@pytest.mark.parametrize("n_sets", [2, 7, 8, 9, 64, 200, 255, 256])
def test_ava_blocks_cover_the_upper_triangle_once(n_sets):
    nb = (n_sets + 7) // 8
    seen = set()
    out = np.zeros(3, np.uint32)
    for b in range(576):
        assert ok.lib().okx_ava_block(b, n_sets, ok._ptr(out)) == 0
        if b < nb * (nb + 1) // 2:
            assert out[0] == 1 and out[1] <= out[2] < nb
            seen.add((int(out[1]), int(out[2])))
        else:
            assert out[0] == 0
    assert seen == {(i, j) for i in range(nb) for j in range(i, nb)}


def test_ava_keyed_algorithm_replayed_on_the_host_matches_pairwise_intersections():
    """The keyed all-vs-all (setops.cuh) replayed in numpy over the library's own tile geometry: cut every set at the
    tile bounds, per tile count the holders of each key, give the keys with >= 2 holders a bit each, and add
    popcount(row_i & row_j) to the pair matrix.  The sum over tiles must be |A n B| for every pair (compare.rs:58) --
    the claim the CUDA kernel rests on -- and no tile may exceed the shared-memory table (6144 entries)."""
    rng = np.random.default_rng(5)
    k = 21
    pool = np.unique(np.minimum(rng.integers(0, 1 << 42, 60_000, dtype=np.uint64), rng.integers(0, 1 << 42, 60_000, dtype=np.uint64)))
    sets = [pool[rng.random(len(pool)) < f] for f in (0.6, 0.5, 0.9, 0.0, 0.2, 1.0, 0.05, 0.5, 0.7)]
    sets[7] = np.unique(np.concatenate([sets[7], rng.integers(0, 1 << 42, 5000, dtype=np.uint64)]))     # keys nobody else holds
    n = len(sets)
    ns = np.array([len(s) for s in sets], np.uint64)
    ends = np.zeros(2 * n, np.uint64)
    for i, s in enumerate(sets):
        if len(s):
            ends[2 * i], ends[2 * i + 1] = s[0], s[-1]
    geo = np.zeros(3, np.uint64)
    inter = np.zeros((n, n), np.int64)
    per_set_tiles = []
    for s in sets:
        tiles = np.zeros(len(s), np.uint32)
        assert ok.lib().okx_ava_geometry(k, ok._ptr(ends), ok._ptr(ns), n, int(ns.sum()), ok._ptr(s), len(s), ok._ptr(geo), ok._ptr(tiles)) == 0
        assert np.all(np.diff(tiles.astype(np.int64)) >= 0)         # monotone: a tile is one contiguous range of the sorted set
        per_set_tiles.append(tiles)
    n_tiles = int(geo[0])
    assert n_tiles == int(ns.sum()) // 3072
    for t in range(n_tiles):
        ranges = [s[tl == t] for s, tl in zip(sets, per_set_tiles)]
        entries = np.concatenate(ranges)
        assert len(entries) <= 6144
        keys, holders = np.unique(entries, return_counts=True)
        shared = keys[holders >= 2]                                 # dense ids = positions in `shared`
        rows = np.zeros((n, len(shared)), bool)
        for i, r in enumerate(ranges):
            rows[i, np.searchsorted(shared, r[np.isin(r, shared, assume_unique=True)])] = True
        inter += rows.astype(np.int64) @ rows.astype(np.int64).T    # popcount(row_i & row_j) for every pair
    for i in range(n):
        for j in range(i + 1, n):
            assert inter[i, j] == len(np.intersect1d(sets[i], sets[j], assume_unique=True)), (i, j)


# ------------------------------------------------------------- serde_json's f64 text in the compare / classify reports --
def test_json_f64_follows_ryu_layout_and_round_trips():
    """compare.rs:16-25 / classify.rs:22-52 serialise their ratios with serde_json, which prints an f64 through ryu: the
    shortest digits that round-trip, in ryu's layout (plain decimals for 1e-5 <= |v| < 1e16, d.ddde<exp> outside, always
    a fraction or an exponent).  Known layouts, then 20,000 random doubles: the text reads back to the same double and
    carries the digits of the shortest representation (Python's repr)."""
    import random
    import struct
    want = [(1.0, "1.0"), (0.0, "0.0"), (-0.0, "-0.0"), (0.5, "0.5"), (10.0, "10.0"), (20.0, "20.0"), (100.0, "100.0"), (12.5, "12.5"),
            (1 / 3, "0.3333333333333333"), (2 / 3, "0.6666666666666666"), (0.1, "0.1"), (0.30000000000000004, "0.30000000000000004"),
            (1e15, "1000000000000000.0"), (1e16, "1e16"), (1.2345678901234568e17, "1.2345678901234568e17"),
            (123456789.125, "123456789.125"), (0.0001, "0.0001"), (1e-5, "0.00001"), (1.234e-5, "0.00001234"), (1e-6, "1e-6"),
            (1.234e-7, "1.234e-7"), (5e-324, "5e-324"), (1.7976931348623157e308, "1.7976931348623157e308"), (-2.5, "-2.5"),
            (-1e-7, "-1e-7"), (float("nan"), "null"), (float("inf"), "null"), (float("-inf"), "null")]
    for v, text in want:
        assert ok.json_f64(v) == text, (v, ok.json_f64(v), text)

    def digits(s):
        return s.lstrip("-").split("e")[0].replace(".", "").lstrip("0").rstrip("0") or "0"
    rnd = random.Random(11)
    for i in range(20_000):
        v = (rnd.random(), rnd.randint(0, 10 ** 7) / rnd.randint(1, 10 ** 7), rnd.random() * 10.0 ** rnd.randint(-30, 30),
             struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0])[i % 4]
        if v != v or abs(v) == float("inf"):
            continue
        t = ok.json_f64(v)
        assert float(t) == v and digits(t) == digits(repr(v)), (v, t)
        assert ("." in t or "e" in t) and not t.endswith(".")
        if 1e-5 <= abs(v) < 1e16:
            assert "e" not in t, (v, t)
