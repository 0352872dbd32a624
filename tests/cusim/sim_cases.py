"""TEST INFRASTRUCTURE: cases run against liborion_gpu_sim.so (the whole library on the CPU stand-in, ORION_GPU_LIB set by
tests/test_library_cusim.py before this process imports the package).  They go through the same ctypes mirror and the
same C ABI as the GPU tests; sizes are what a CPU can simulate in seconds."""
import os

import numpy as np
import pytest

import orion_kmer_b200 as ok
from orion_kmer_b200 import synth

assert ok.gpu_library_path().endswith("liborion_gpu_sim.so"), "run me through tests/test_library_cusim.py"


@pytest.fixture(scope="module")
def oracle():
    import oracle as orc
    orc.build()
    return orc


FULL = bool(os.environ.get("CUSIM_FULL"))


def test_partitioned_count_and_a_second_batch_merged(oracle):
    """the hot path end to end -- sample, plan, level-1 scatter, TMA-fed level-2 scatter, shared-memory count with the
    dense look-back output, sliced result pipeline -- then a second large batch counted on its own and merged
    (count.rs:48: ONE table across batches)"""
    g = synth.genome(42, 200_000)
    n = 7200
    b1, off = synth.reads(g, 43, n), synth.read_offsets(n)
    b2 = synth.reads(g, 44, n)
    c = ok.KmerCounter(31)
    c.add_batch(b1, off)
    keys, counts = c.finish(1)
    st = c.stats()
    assert st["partitioned"] == 1 and st["n_spilled"] == 0
    wk, wc = oracle.count_batch(31, b1, off)
    assert np.array_equal(keys, wk) and np.array_equal(counts, wc)
    if not FULL:                        # (the merge of a second batch: another 25 s of simulation, CUSIM_FULL=1)
        c.close()
        return
    c.add_batch(b2, off)
    keys, counts = c.finish(2)
    assert c.stats()["n_merges"] >= 1
    both = np.concatenate([b1, b2])
    off2 = np.arange(2 * n + 1, dtype=np.uint64) * np.uint64(150)
    wk, wc = oracle.count_batch(31, both, off2, 2)
    assert np.array_equal(keys, wk) and np.array_equal(counts, wc)
    c.close()


@pytest.mark.parametrize("mode", ["1", "0"])
def test_query_by_merge_and_by_table(oracle, monkeypatch, mode):
    """query.rs:83-107: the probe by merge (k_member_tiled, forced) and the hashed table of the whole set"""
    monkeypatch.setenv("ORION_PROBE_MERGE", mode)
    k = 31
    g, other = synth.genome(50, 60_000), synth.genome(51, 60_000)
    kset = ok.KmerSet.from_fastx(k, synth.fasta_text(b"g", g))
    oset = oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64))
    n = 600
    bases = np.concatenate([synth.reads(g, 52, n), synth.reads(other, 53, n)])
    off = synth.read_offsets(2 * n)
    want = oracle.query_hits(oset, k, bases, off, 2)
    assert np.array_equal(kset.probe_reads(bases, off, ok.RAW).astype(np.uint64), want)
    assert want[:n].min() > 10 and want[n:].max() < 3
    empty = ok.KmerSet.from_sorted(k, np.zeros(0, np.uint64))
    assert not empty.probe_reads(bases[:1500], off[:11], ok.RAW).any()
    w = ok.KmerSet.from_sorted(32, np.array([0, 5, 2 ** 64 - 1], np.uint64))
    polya = np.frombuffer(b"A" * 40 + b"T" * 40 + b"ACGT" * 10, np.uint8)
    assert list(w.probe_reads(polya, np.array([0, 40, 80, 120], np.uint64), ok.RAW)) == [9, 9, 0]
    for x in (kset, empty, w):
        x.close()


@pytest.mark.parametrize("threads", ["4", "1"] if FULL else ["4"])
def test_build_many_side_by_side(oracle, monkeypatch, threads):
    """build.rs:93-116 over many files: several host threads, each with its own pooled builder, launching side by side"""
    monkeypatch.setenv("ORION_BUILD_THREADS", threads)
    k = 21
    files = []
    for i, n in enumerate([30_000, 0, 20, 12_000, 5000, 25_000, 40, 9000]):
        g = synth.genome(300 + i, max(n, 1))[:n]
        off = np.array([0, n // 3, n], np.uint64) if n else np.array([0, 0], np.uint64)
        files.append((np.ascontiguousarray(g, dtype=np.uint8), off))
    want = [oracle.kmer_set_batch(k, b, o) for b, o in files]
    for s, w in zip(ok.KmerSet.build_many(k, files), want):
        assert len(s) == len(w) and np.array_equal(s.to_array(), w)
        s.close()
    with pytest.raises(ok.OrionError, match="Invalid K-mer size"):
        ok.KmerSet.build_many(33, files[:2])


def test_sets_union_compare_classify(oracle):
    """db_types.rs:43-48, compare.rs:51-66, classify.rs:224-236 on small sets (table path, setwise union, row form)"""
    k = 21
    a, b = synth.genome(7, 20_000), synth.genome(8, 15_000)
    gens = [a, synth.mutate(a, 1, 300), b, a[:5000]]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    u = ok.KmerSet.union(sets)
    assert np.array_equal(u.to_array(), oracle.set_union(osets))
    sizes, inter = ok.all_vs_all(sets)
    for i in range(4):
        for j in range(4):
            assert inter[i, j] == (len(osets[i]) if i == j else len(np.intersect1d(osets[i], osets[j], assume_unique=True)))
    r = ok.compare(sets[0], sets[1])
    w = oracle.compare(osets[0], osets[1])
    assert (r["intersection_size"], r["union_size"]) == (w["intersection_size"], w["union_size"])
    reads, off = synth.reads(a, 9, 300), synth.read_offsets(300)
    ik, ic = oracle.count_batch(k, reads, off)
    m, d = ok.probe_counts_many(sets, ik, ic)
    for i, s in enumerate(osets):
        assert (int(m[i]), int(d[i])) == oracle.classify_ref(ik, ic, s)
    for x in sets + [u]:
        x.close()


@pytest.mark.skipif(not os.environ.get("CUSIM_FULL"), reason="minutes of simulation: CUSIM_FULL=1")
def test_union_of_sorted_sets_through_the_strided_gather(oracle, monkeypatch):
    """db_types.rs:43-48 through ONE partitioned pass (>= 2^20 keys): the strided level-1 gather and the contiguous form"""
    k = 31
    base = synth.genome(61, 400_000)
    gens = [base, synth.mutate(base, 10, 3000), synth.genome(62, 333_333), base[:77_777]]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    assert sum(len(o) for o in osets) > (1 << 20)
    want = oracle.set_union(osets)
    for no_stride in (False, True):
        if no_stride:
            monkeypatch.setenv("ORION_UNION_NO_STRIDE", "1")
        u = ok.KmerSet.union(sets)
        assert np.array_equal(u.to_array(), want), no_stride
        u.close()
    for x in sets:
        x.close()
