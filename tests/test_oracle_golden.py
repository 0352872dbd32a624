"""Pins the CPU oracle (oracle/orion_oracle.cpp) against the reference's own known-answer
vectors (tests/golden/reference_vectors.json, transcribed from orion-kmer/src/kmer.rs:108-341
and orion-kmer/tests/*.rs).  CPU only."""
import numpy as np
import pytest

U64 = 2 ** 64 - 1


def enc(o, s, k=None):
    k = len(s) if k is None else k
    return o.seq_to_u64(s.encode(), k)


def canon(o, s):
    return o.canonical_u64(enc(o, s), len(s))


def as_dict(o, keys, counts, k):
    return {o.u64_to_seq(int(a), k).decode(): int(c) for a, c in zip(keys, counts)}


# ---- src/kmer.rs unit vectors -------------------------------------------------------------
def test_seq_to_u64_valid(oracle, golden):
    for s, k, v in golden["kmer"]["seq_to_u64_valid"]:
        assert oracle.seq_to_u64(s.encode(), k) == v, (s, k)


def test_seq_to_u64_none(oracle, golden):
    for s, k in golden["kmer"]["seq_to_u64_none"]:
        assert oracle.seq_to_u64(s.encode(), k) is None, (s, k)


def test_u64_to_seq(oracle, golden):
    for v, k, s in golden["kmer"]["u64_to_seq"]:
        assert oracle.u64_to_seq(v, k) == s.encode()
    for k in golden["kmer"]["panic_k"]:
        with pytest.raises(ValueError):
            oracle.u64_to_seq(0, k)
        with pytest.raises(ValueError):
            oracle.reverse_complement_u64(0, k)


def test_reverse_complement(oracle, golden):
    for a, b in golden["kmer"]["reverse_complement"]:
        assert oracle.reverse_complement_u64(enc(oracle, a), len(a)) == enc(oracle, b)


def test_canonical(oracle, golden):
    for a, b in golden["kmer"]["canonical"]:
        assert canon(oracle, a) == enc(oracle, b), (a, b)
    assert canon(oracle, "TGGG") != canon(oracle, "GGGA")  # kmer.rs:153-156


def test_revcomp_closed_form(oracle):
    """reverse_complement == 2-bit-group reversal of ~v shifted right by 64-2k"""
    rng = np.random.default_rng(7)
    for k in (1, 2, 5, 21, 31, 32):
        for v in rng.integers(0, 2 ** 63, size=50, dtype=np.uint64):
            v = int(v) & ((1 << (2 * k)) - 1) if k < 32 else int(v)
            x = (~v) & U64
            r = 0
            for i in range(32):
                r |= ((x >> (2 * i)) & 3) << (2 * (31 - i))
            assert oracle.reverse_complement_u64(v, k) == r >> (64 - 2 * k)


# ---- count (tests/count_tests.rs) -----------------------------------------------------------
def test_count_cases(oracle, golden):
    g = golden["count"]
    for case in g["cases"]:
        contents = [g["files"][n].encode() for n in case["inputs"]]
        keys, counts = oracle.count_fastx(case["k"], contents, case["min_count"])
        assert list(keys) == sorted(keys)
        assert as_dict(oracle, keys, counts, case["k"]) == case["expected"], case["name"]
        # exact TSV text, ascending (count.rs:119,133)
        text = oracle.format_counts(keys, counts, case["k"]).decode()
        assert text == "".join(f"{s}\t{c}\n" for s, c in sorted(case["expected"].items()))


def test_count_invalid_k(oracle, golden):
    for k in golden["count"]["invalid_k"]:
        with pytest.raises(oracle.InvalidKmerSize) as e:
            oracle.Counter(k)
        assert str(e.value) == golden["count"]["invalid_k_message"].format(k=k)


# ---- build (tests/build_tests.rs) ------------------------------------------------------------
def test_build_cases(oracle, golden):
    for case in golden["build"]["cases"]:
        k = case["k"]
        sets = {}
        for name, content in case["files"].items():
            sets[name] = oracle.kmer_set_fastx(k, content.encode())
            want = sorted({canon(oracle, s) for s in case["expected"][name]})
            assert list(sets[name]) == want, (case["name"], name)
        assert len(oracle.set_union(list(sets.values()))) == case["total_unique"], case["name"]


# ---- compare (tests/compare_tests.rs) ---------------------------------------------------------
def test_compare_cases(oracle, golden):
    g = golden["compare"]
    for case in g["cases"]:
        a = oracle.kmer_set_fastx(case["k"], case["db1"].encode())
        b = oracle.kmer_set_fastx(case["k"], case["db2"].encode())
        r = oracle.compare(a, b)
        assert r["db1"] == case["db1_size"] and r["db2"] == case["db2_size"], case["name"]
        assert r["intersection_size"] == case["intersection_size"]
        assert r["union_size"] == case["union_size"]
        assert abs(r["jaccard_index"] - case["jaccard"]) < g["jaccard_tolerance"]
    assert oracle.compare(np.zeros(0, np.uint64), np.zeros(0, np.uint64))["jaccard_index"] == 0.0


# ---- query (tests/query_tests.rs) --------------------------------------------------------------
def test_query_hits(oracle, golden):
    g = golden["query"]
    kset = oracle.kmer_set_fastx(g["k"], g["db"].encode())
    recs = oracle.parse_fastx(g["reads"].encode())
    bases, off = oracle.batch_from_records([s for _, s in recs])
    for nt in (1, 3):
        hits = oracle.query_hits(kset, g["k"], bases, off, nt)
        assert list(hits) == g["hits"]
    for mh, ids in g["ids_by_min_hits"].items():
        got = [i.decode() for (i, _), h in zip(recs, hits) if h >= int(mh)]
        assert got == ids


# ---- classify (tests/classify_tests.rs) ---------------------------------------------------------
def test_classify_cases(oracle, golden):
    for case in golden["classify"]["cases"]:
        k = case["k"]
        keys, counts = oracle.count_fastx(k, [case["input"].encode()], case["min_kmer_frequency"])
        assert len(keys) == case["total_unique_kmers_in_input"], case["name"]
        for db in case["databases"]:
            refs = {n: oracle.kmer_set_fastx(k, c.encode()) for n, c in db["refs"].items()}
            union = oracle.set_union(list(refs.values()))
            assert len(union) == db["total_unique_kmers_in_db"]
            m, d = oracle.classify_ref(keys, counts, union)  # classify.rs:272-277
            assert (m, d) == (db["overall_matched"], db["overall_sum_depth"])
            for name, want in db["per_ref"].items():
                m, d = oracle.classify_ref(keys, counts, refs[name])
                assert len(refs[name]) == want["total"]
                assert (m, d) == (want["matched"], want["sum_depth"]), (case["name"], name)


# ---- values re-derived from src/ for the reference's fixtures (not reference-pinned) -----------
def test_fixture_derived(oracle, golden):
    g = golden["derived_from_src"]
    keys, counts = oracle.count_fastx(7, [g["test_input1.fasta"].encode()])
    assert as_dict(oracle, keys, counts, 7) == g["input1_k7"]
    keys, counts = oracle.count_fastx(6, [g["test_input2.fastq"].encode()])
    assert as_dict(oracle, keys, counts, 6) == g["input2_k6"]
    ids = [i.decode() for i, _ in oracle.parse_fastx(g["test_input2.fastq"].encode())]
    assert ids == g["input2_ids"]
    # multi-line record: raw sequence keeps the line breaks, normalize() removes them
    recs = oracle.parse_fastx(g["test_input1.fasta"].encode())
    assert recs[2][1] == b"GATTACA\nNNNNN\nGATTACA"
    assert oracle.normalize(recs[2][1]) == b"GATTACANNNNNGATTACA"


def test_framing_edges(oracle):
    assert oracle.parse_fastx(b">h1\n>h2\n") == [(b"h1", b""), (b"h2", b"")]
    assert oracle.parse_fastx(b">h desc\r\nAC\r\nGT\r\n") == [(b"h desc", b"AC\r\nGT")]
    assert oracle.normalize(b"AC\r\nGu.x") == b"ACGT-N"
    for bad in (b"", b"This is not fasta\nACGT"):
        with pytest.raises(oracle.FastxError):
            oracle.parse_fastx(bad)
    with pytest.raises(oracle.FastxError):
        oracle.parse_fastx(b"@r\nACGT\n+\n!!!\n")  # quality length mismatch


def test_mt_variant_equals_faithful(oracle):
    rng = np.random.default_rng(3)
    bases = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=20000, p=[.245, .245, .245, .245, .02])
    off = np.arange(0, 20001, 100, dtype=np.uint64)
    a = oracle.count_batch(11, bases, off)
    b = oracle.count_batch_mt(11, bases, off, 4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_large_input_checkers_equal_the_plain_restatement(oracle):
    """orc_count_batch_ranged_mt (whole table of a batch too large for per-thread maps) and orc_count_batch_slice_mt (one
    key slice) only reorganise the work: every window still goes through seq_to_u64 + canonical_u64 and an exact map.
    Both must reproduce the single-threaded restatement of count.rs:23-38,106-119."""
    import numpy as np
    rng = np.random.default_rng(12)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    for k in (1, 3, 4, 5, 11, 21, 31, 32):
        bases = alphabet[rng.integers(0, len(alphabet), 60_000)]
        cuts = np.unique(rng.integers(0, len(bases), 400))
        off = np.concatenate([[0], cuts, [len(bases)]]).astype(np.uint64)
        wk, wc = oracle.count_batch(k, bases, off, 1, False)
        for nt in (1, 3, 8):
            gk, gc = oracle.count_batch_ranged_mt(k, bases, off, nt)
            assert np.array_equal(gk, wk) and np.array_equal(gc, wc), (k, nt)
            for mc in (2, 3):
                fk, fc = oracle.count_batch_ranged_mt(k, bases, off, nt, mc)
                assert np.array_equal(fk, wk[wc >= mc]) and np.array_equal(fc, wc[wc >= mc])
        if len(wk) > 10:
            lo, hi = int(wk[len(wk) // 4]), int(wk[3 * len(wk) // 4])
            sk, sc = oracle.count_batch_slice_mt(k, bases, off, lo, hi, 4)
            sel = (wk >= lo) & (wk <= hi)
            assert np.array_equal(sk, wk[sel]) and np.array_equal(sc, wc[sel]), k
        ek, ec = oracle.count_batch_slice_mt(k, bases, off, 0, 2 ** 64 - 1, 2)
        assert np.array_equal(ek, wk) and np.array_equal(ec, wc)
