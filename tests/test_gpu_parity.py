"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden
vectors.  Needs a B200: run with `pytest -m gpu` under gpurun."""
import ctypes as C

import numpy as np
import pytest

import orion_kmer_b200 as ok
from orion_kmer_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    ok.init(0)


def enc(s):
    return ok.seq_to_u64(s.encode(), len(s))


def canon(s):
    return ok.canonical_u64(enc(s), len(s))


def as_dict(keys, counts, k):
    return {ok.u64_to_seq(int(a), k).decode(): int(c) for a, c in zip(keys, counts)}


ALPHABET = np.frombuffer(b"ACGTacgtNnUuRYKM-.*X", dtype=np.uint8)


def random_batch(rng, n_bases, mean_len, p_junk=0.02):
    p = np.full(len(ALPHABET), p_junk / (len(ALPHABET) - 4))
    p[:4] = (1 - p_junk) / 4
    bases = rng.choice(ALPHABET, size=n_bases, p=p)
    cuts = np.unique(rng.integers(0, n_bases + 1, size=max(1, n_bases // mean_len)))
    off = np.concatenate([[0], cuts, [n_bases]]).astype(np.uint64)
    off = np.sort(np.concatenate([off, off[1:4]]))  # a few empty records
    return bases, off


# ---------------------------------------------------------------- extraction kernel alone --
@pytest.mark.parametrize("k", [1, 2, 3, 5, 16, 21, 31, 32])
@pytest.mark.parametrize("norm", [ok.NORMALIZED, ok.RAW])
def test_device_extract_matches_oracle(oracle, k, norm):
    rng = np.random.default_rng(1000 + k)
    for n_bases, mean_len in ((1, 1), (33, 5), (1500, 150), (70000, 37), (100000, 100000), (50000, 2)):
        bases, off = random_batch(rng, n_bases, mean_len)
        out = np.zeros(n_bases, dtype=np.uint64)
        n = C.c_uint64()
        rc = ok.lib().okx_device_extract(ok._ptr(bases), ok._ptr(off), len(off) - 1, k, norm, ok._ptr(out),
                                         len(out), C.byref(n))
        assert rc == 0, ok.lib().ok_last_error()
        gk, gc = np.unique(out[:n.value], return_counts=True)
        wk, wc = oracle.count_batch(k, bases, off, 1, norm == ok.NORMALIZED)
        assert np.array_equal(gk, wk), (k, norm, n_bases)
        assert np.array_equal(gc.astype(np.uint64), wc)


def test_pack_kernel_matches_host_code(oracle):
    import torch
    rng = np.random.default_rng(3)
    bases, _ = random_batch(rng, 100_001, 1000, p_junk=0.1)
    d = torch.from_numpy(bases).cuda()
    ng = (len(bases) + 31) // 32
    codes = torch.zeros(ng, dtype=torch.int64, device="cuda")
    valid = torch.zeros(ng, dtype=torch.int32, device="cuda")
    for norm in (ok.NORMALIZED, ok.RAW):
        ok._check(ok.lib().ok_pack_2bit_device(d.data_ptr(), len(bases), norm, codes.data_ptr(), valid.data_ptr()))
        c = codes.cpu().numpy().view(np.uint64)
        v = valid.cpu().numpy().view(np.uint32)
        for g in rng.integers(0, ng, size=300):
            chunk = bytes(bases[g * 32:(g + 1) * 32])
            for i, b in enumerate(chunk):
                code = oracle.seq_to_u64(bytes([b]), 1)
                is_u = b in b"Uu" and norm == ok.NORMALIZED
                want_valid = code is not None or is_u
                assert bool((int(v[g]) >> (31 - i)) & 1) == want_valid
                if want_valid:
                    assert (int(c[g]) >> (62 - 2 * i)) & 3 == (3 if is_u else code)


# ------------------------------------------------------- count: the reference's own vectors --
def test_count_golden_cases(golden):
    g = golden["count"]
    for case in g["cases"]:
        contents = [g["files"][n].encode() for n in case["inputs"]]
        keys, counts = ok.count_fastx(case["k"], contents, case["min_count"])
        assert as_dict(keys, counts, case["k"]) == case["expected"], case["name"]
        text = ok.run_count(case["k"], contents, case["min_count"]).decode()
        assert text == "".join(f"{s}\t{c}\n" for s, c in sorted(case["expected"].items()))


def test_count_fixture_files(golden):
    g = golden["derived_from_src"]
    keys, counts = ok.count_fastx(7, [g["test_input1.fasta"].encode()])
    assert as_dict(keys, counts, 7) == g["input1_k7"]
    keys, counts = ok.count_fastx(6, [g["test_input2.fastq"].encode()])
    assert as_dict(keys, counts, 6) == g["input2_k6"]


def test_count_edge_cases(oracle):
    c = ok.KmerCounter(5)
    assert c.finish()[0].size == 0                               # nothing added
    c.add_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))  # empty batch
    c.add_batch(np.frombuffer(b"ACG", np.uint8), np.array([0, 3], np.uint64))  # shorter than k
    c.add_batch(np.frombuffer(b"NNNNNNNN", np.uint8), np.array([0, 8], np.uint64))
    assert c.finish()[0].size == 0
    c.add_batch(np.frombuffer(b"ACGTA", np.uint8), np.array([0, 5], np.uint64))
    k, n = c.finish()
    assert list(k) == [canon("ACGTA")] and list(n) == [1]
    c.close()
    # k = 32: all-T and all-A reads collapse on the key 0; u64::MAX is never canonical
    c = ok.KmerCounter(32)
    b = np.frombuffer(b"T" * 40 + b"A" * 40, np.uint8)
    c.add_batch(b, np.array([0, 40, 80], np.uint64))
    k, n = c.finish()
    assert list(k) == [0] and list(n) == [18]
    c.close()


@pytest.mark.parametrize("k", [1, 2, 4, 11, 21, 31, 32])
def test_count_random_batches_match_oracle(oracle, k):
    rng = np.random.default_rng(k)
    c = ok.KmerCounter(k)
    o = oracle.Counter(k)
    for n_bases, mean_len in ((300_000, 150), (200_000, 5000), (65_536, 64), (1, 1)):
        bases, off = random_batch(rng, n_bases, mean_len)
        c.add_batch(bases, off)
        o.add_batch(bases, off, True)
    for min_count in (1, 2, 5):
        gk, gc = c.finish(min_count)
        wk, wc = o.finish(min_count)
        assert np.array_equal(gk, wk), (k, min_count)
        assert np.array_equal(gc, wc), (k, min_count)
    st = c.stats()
    assert st["n_windows"] == int(o.finish(1)[1].sum())
    c.close()


def test_count_grows_without_hint_and_with_small_hint(oracle):
    g = synth.genome(5, 3_000_000)
    off = np.array([0, len(g)], np.uint64)
    wk, wc = oracle.count_batch(25, g, off)
    for hint in (0, 1000):
        c = ok.KmerCounter(25, capacity_hint=hint)
        third = len(g) // 3
        # three batches that overlap by k-1 bases would double count: feed disjoint records instead
        c.add_batch(g, off)
        gk, gc = c.finish()
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
        c.add_batch(g[:third], np.array([0, third], np.uint64))   # a second batch on top
        gk2, gc2 = c.finish()
        o = oracle.Counter(25)
        o.add_batch(g, off)
        o.add_batch(g[:third], np.array([0, third], np.uint64))
        wk2, wc2 = o.finish()
        assert np.array_equal(gk2, wk2) and np.array_equal(gc2, wc2)
        assert c.stats()["n_spilled"] == 0
        c.close()


def test_count_device_resident_input(oracle):
    import torch
    g = synth.genome(8, 400_000)
    n_reads = 30_000
    bases = synth.reads(g, 9, n_reads)
    off = synth.read_offsets(n_reads)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    c = ok.KmerCounter(31, capacity_hint=4_000_000)
    c.add_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), n_reads)
    dk, dc, n = c.finish_device()
    wk, wc = oracle.count_batch(31, bases, off)
    assert n == len(wk)
    assert dk and dc
    h = c.finish()
    assert np.array_equal(h[0], wk) and np.array_equal(h[1], wc)
    c.clear()
    assert c.finish()[0].size == 0
    c.add_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), n_reads)
    h = c.finish()
    assert np.array_equal(h[0], wk) and np.array_equal(h[1], wc)
    c.close()


def test_count_config1_genome_exact(oracle):
    """BASELINE.json configs[0]: canonical 21-mers of one 5 Mbp genome record (80-column FASTA)."""
    g = synth.config1_genome()
    text = synth.fasta_text(b"chr1", g)
    keys, counts = ok.count_fastx(21, [text])
    wk, wc = oracle.count_fastx(21, [text])
    assert np.array_equal(keys, wk) and np.array_equal(counts, wc)
    assert ok.format_counts(keys[:1000], counts[:1000], 21) == oracle.format_counts(wk[:1000], wc[:1000], 21)


def test_count_config2_sample_exact(oracle):
    """BASELINE.json configs[1] recipe on a 300k-read sample (the oracle finishes it in seconds)."""
    g = synth.genome(3, 1_500_000)
    n_reads = 300_000
    bases = synth.reads(g, 3, n_reads)
    off = synth.read_offsets(n_reads)
    c = ok.KmerCounter(31)
    c.add_batch(bases, off)
    keys, counts = c.finish()
    wk, wc = oracle.count_batch(31, bases, off)
    assert np.array_equal(keys, wk) and np.array_equal(counts, wc)
    c.close()


def _available_host_memory():
    avail = None
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                avail = int(ln.split()[1]) * 1024
    except OSError:
        pass
    for f in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(f).read().strip()
            if v.isdigit():
                lim = int(v)
                try:
                    lim -= int(open(f.replace("memory.max", "memory.current").replace("limit_in_bytes", "usage_in_bytes")).read())
                except (OSError, ValueError):
                    pass
                avail = lim if avail is None else min(avail, lim)
        except OSError:
            pass
    return avail


def test_count_config2_full_batch_with_bench_hint_exact(oracle):
    """BASELINE.json configs[1] AT FULL SIZE, counted exactly the way bench.py times it -- device-resident batch of
    10 M x 150 bp reads, capacity hint 0.17 x bases (hint-sized sub-partitions, k_part_count with packed 16-bit counts
    and its early-out) -- and compared, whole table, with the oracle (count.rs:23-38,106-119).  Also through the
    host-buffer entry (sliced result pipeline) and without the hint.  The oracle pass needs ~20 GB of host memory;
    on a smaller host four key slices (1/64 of the key space each) are compared instead."""
    import json
    import os
    import time
    import torch
    n_reads, k = 10_000_000, 31
    g = synth.genome(3, 5 * n_reads)
    bases = synth.reads(g, 3, n_reads)
    off = synth.read_offsets(n_reads)
    n_bases = len(bases)
    hint = int(n_bases * 0.17)
    nt = min(os.cpu_count() or 1, 64)
    t0 = time.time()
    mem = _available_host_memory()
    full = mem is None or mem > 28 * 2 ** 30
    if full:
        wk, wc = oracle.count_batch_ranged_mt(k, bases, off, nt)
        slices = [(0, 2 ** 64 - 1)]
    else:
        span = (1 << 62) >> 6
        slices = [(a, a + span - 1) for a in (0, (1 << 62) // 3, (1 << 62) // 2 + 12345, (1 << 62) - span)]
        parts = [oracle.count_batch_slice_mt(k, bases, off, lo, hi, nt) for lo, hi in slices]
        wk, wc = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    t_oracle = time.time() - t0

    def check(keys, counts, what):
        if not full:
            sel = np.zeros(len(keys), bool)
            for lo, hi in slices:
                sel[np.searchsorted(keys, np.uint64(lo), "left"):np.searchsorted(keys, np.uint64(hi), "right")] = True
            assert np.all(keys[1:] > keys[:-1]), what
            keys, counts = keys[sel], counts[sel]
        assert len(keys) == len(wk), (what, len(keys), len(wk))
        assert np.array_equal(keys, wk), what
        assert np.array_equal(counts, wc), what

    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    record = {"reads": n_reads, "k": k, "hint": hint, "oracle": "count_batch_ranged_mt" if full else "4 key slices of 1/64",
              "oracle_threads": nt, "oracle_seconds": round(t_oracle, 1), "oracle_distinct": int(len(wk)),
              "oracle_windows": int(wc.sum()), "checked": []}
    for what, h, host in (("device batch, bench hint", hint, False), ("host batch (sliced result pipeline), bench hint", hint, True),
                          ("device batch, no hint", 0, False)):
        c = ok.KmerCounter(k, ok.NORMALIZED, h)
        if host:
            c.add_batch(bases, off)
        else:
            c.add_batch_device(d_b.data_ptr(), n_bases, d_o.data_ptr(), n_reads)
        keys, counts = c.finish(1)
        st = c.stats()
        c.close()
        assert st["partitioned"] == 1, what
        check(keys, counts, what)
        record["checked"].append({"what": what, "distinct": int(len(keys)), "windows": int(st["n_windows"]),
                                  "n_deferred": int(st["n_deferred"]), "n_spilled": int(st["n_spilled"]), "equal": True})
        del keys, counts
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_config2_full.json"), "w") as f:
            json.dump(record, f, indent=1)


def test_count_large_properties():
    """size-independent properties at a size the oracle is not run on: total = number of
    countable windows, strictly ascending keys, and both-strand invariance."""
    g = synth.genome(21, 20_000_000)
    n_reads = 2_000_000
    bases = synth.reads(g, 22, n_reads)
    off = synth.read_offsets(n_reads)
    k = 31
    c = ok.KmerCounter(k, capacity_hint=60_000_000)
    c.add_batch(bases, off)
    keys, counts = c.finish()
    c.close()
    assert np.all(keys[1:] > keys[:-1])
    # countable windows = windows with no N: count them with a prefix sum of N flags
    isn = (bases.reshape(n_reads, 150) == ord("N")).astype(np.int32)
    cs = np.concatenate([np.zeros((n_reads, 1), np.int32), np.cumsum(isn, axis=1)], axis=1)
    windows = int(((cs[:, k:] - cs[:, :-k]) == 0).sum())
    assert int(counts.sum()) == windows
    # reverse-complementing every read leaves the canonical count table unchanged
    comp = np.zeros(256, np.uint8)
    comp[:] = np.arange(256)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    rc = comp[bases.reshape(n_reads, 150)[:, ::-1]].reshape(-1).copy()
    c = ok.KmerCounter(k, capacity_hint=60_000_000)
    c.add_batch(rc, off)
    k2, n2 = c.finish()
    c.close()
    assert np.array_equal(keys, k2) and np.array_equal(counts, n2)


def test_invalid_k_messages(golden):
    for k in golden["count"]["invalid_k"]:
        with pytest.raises(ok.InvalidKmerSize) as e:
            ok.KmerCounter(k)
        assert str(e.value) == golden["count"]["invalid_k_message"].format(k=k)
        with pytest.raises(ok.InvalidKmerSize):
            ok.run_build(k, {"a.fa": b">a\nACGT"})


# ----------------------------------------------------------------- build / sets / compare --
def test_build_golden_cases(golden):
    for case in golden["build"]["cases"]:
        db = ok.run_build(case["k"], {n: c.encode() for n, c in case["files"].items()})
        assert db.k == case["k"] and db.num_references() == len(case["files"])
        for name, want in case["expected"].items():
            assert list(db.references[name].to_array()) == sorted({canon(s) for s in want}), (case["name"], name)
        assert db.total_unique_kmers() == case["total_unique"], case["name"]


def test_compare_golden_cases(golden):
    g = golden["compare"]
    for case in g["cases"]:
        a = ok.run_build(case["k"], {"db1.fa": case["db1"].encode()}).get_all_kmers_unified()
        b = ok.run_build(case["k"], {"db2.fa": case["db2"].encode()}).get_all_kmers_unified()
        r = ok.compare(a, b)
        assert (r["db1"], r["db2"]) == (case["db1_size"], case["db2_size"]), case["name"]
        assert r["intersection_size"] == case["intersection_size"]
        assert r["union_size"] == case["union_size"]
        assert abs(r["jaccard_index"] - case["jaccard"]) < g["jaccard_tolerance"]
    a = ok.KmerSet.from_fastx(3, b">a\nACGTACGT")
    b = ok.KmerSet.from_fastx(4, b">b\nACGTACGT")
    with pytest.raises(ok.KmerSizeMismatch) as e:
        a.intersection_size(b)
    assert str(e.value) == g["mismatch_message"]
    empty = ok.KmerSet.from_sorted(4, np.zeros(0, np.uint64))
    assert ok.compare(empty, empty)["jaccard_index"] == 0.0   # compare.rs:62-63


def test_sets_random_against_oracle(oracle):
    rng = np.random.default_rng(77)
    k = 21
    base = synth.genome(31, 300_000)
    gens = [base, synth.mutate(base, 1, 20_000), synth.mutate(base, 2, 50_000), synth.genome(32, 200_000),
            base[:1000]]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    for s, o in zip(sets, osets):
        assert len(s) == len(o) and np.array_equal(s.to_array(), o)
    u = ok.KmerSet.union(sets)
    assert np.array_equal(u.to_array(), oracle.set_union(osets))
    sizes, inter = ok.all_vs_all(sets)
    for i in range(len(sets)):
        assert sizes[i] == len(osets[i])
        for j in range(len(sets)):
            want = oracle.compare(osets[i], osets[j])["intersection_size"]
            assert inter[i, j] == want, (i, j)
    # classify's loop over references in one call: same numbers as one probe per reference
    pk = np.concatenate([osets[0][::3], osets[3][::5], np.array([1, 2, 3], np.uint64)])
    pc = (np.arange(len(pk), dtype=np.uint64) % np.uint64(7)) + np.uint64(1)
    m_many, d_many = ok.probe_counts_many(sets, pk, pc)
    for i, st in enumerate(sets):
        assert st.probe_counts(pk, pc) == (int(m_many[i]), int(d_many[i]))
        sel = np.isin(pk, osets[i])
        assert int(m_many[i]) == int(sel.sum()) and int(d_many[i]) == int(pc[sel].sum())
    # the multi-GPU split of the same job: every part's pairs, summed (the all-reduce), give the same matrix
    for n_parts in (2, 3, 8):
        parts = [ok.all_vs_all_part(sets, r, n_parts) for r in range(n_parts)]
        assert all(np.array_equal(p[0], sizes) for p in parts)
        assert sum(int(np.count_nonzero(p[1])) for p in parts) <= len(sets) * (len(sets) - 1) // 2
        assert np.array_equal(ok.finish_all_vs_all(sizes, sum(p[1] for p in parts)), inter)
    # foreign (non-canonical) k=32 set holding u64::MAX
    weird = np.array([0, 5, 2 ** 64 - 1], dtype=np.uint64)
    w = ok.KmerSet.from_sorted(32, weird)
    assert w.probe_counts(np.array([2 ** 64 - 1, 5, 6], np.uint64), np.array([7, 1, 1], np.uint64)) == (2, 8)
    assert list(ok.KmerSet.union([w, ok.KmerSet.from_sorted(32, np.array([5, 9], np.uint64))]).to_array()) == [0, 5, 9, 2 ** 64 - 1]


def test_pooled_set_builder_reused_for_another_k(oracle):
    """A sealed set hands its builder (table included) back to the pool; the next set may have another k.  The pooled
    table's home slots were laid out for the old k: reused as is, the readout of a small k=21 / k=31 set after a k=5
    one is no longer sorted.  Small genomes: they stay on the table path, where the bug lived."""
    rng = np.random.default_rng(5)
    for ks in ([5, 21, 31], [31, 21, 5, 32, 3]):
        for i, k in enumerate(ks):
            g = synth.genome(900 + 10 * k + i, 10_000 + 1000 * i)
            s = ok.KmerSet.from_fastx(k, synth.fasta_text(b"g", g))
            want = oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64))
            got = s.to_array()
            assert np.array_equal(got, want), (ks, k)
            other = oracle.kmer_set_batch(k, g[: len(g) // 2], np.array([0, len(g) // 2], np.uint64))
            t = ok.KmerSet.from_sorted(k, other)
            assert s.intersection_size(t) == len(np.intersect1d(want, other)), (ks, k)
            s.close(); t.close()


def test_key_range_shards_add_up_to_the_whole(oracle):
    """multi-GPU set algebra on one GPU: every set cut at the same owner boundaries (ok_set_shard_bounds); intersection
    sizes, per-read hits and (matched, depth) summed over the shards equal the unsharded results and the oracle's."""
    k = 21
    base = synth.genome(41, 400_000)
    gens = [base, synth.mutate(base, 1, 30_000), synth.genome(42, 250_000), base[:5000], synth.mutate(base, 2, 200_000)]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    sizes, inter = ok.all_vs_all(sets)
    n_reads = 5000
    reads = np.concatenate([synth.reads(base, 43, n_reads), synth.reads(gens[2], 44, n_reads)])
    off = synth.read_offsets(2 * n_reads)
    union = ok.KmerSet.union(sets)
    hits = union.probe_reads(reads, off, ok.RAW)
    assert np.array_equal(hits.astype(np.uint64), oracle.query_hits(oracle.set_union(osets), k, reads, off, 8))
    pk, pc = oracle.count_batch(k, reads, off)
    m_all, d_all = ok.probe_counts_many(sets, pk, pc)
    for world in (2, 8):
        owners = np.zeros(len(osets[0]), np.int32)
        ok._check(ok.lib().okx_owner_of(ok._ptr(osets[0]), len(osets[0]), k, world, ok._ptr(owners)))
        acc_inter, acc_hits = np.zeros_like(inter), np.zeros(len(hits), np.uint64)
        acc_m, acc_d = np.zeros(len(sets), np.uint64), np.zeros(len(sets), np.uint64)
        for r in range(world):
            shards = []
            for s, o in zip(sets, osets):
                b = s.shard_bounds(world)
                assert b[0] == 0 and b[-1] == len(o) and np.all(b[1:] >= b[:-1])
                ptr, n = s.keys_device()
                shards.append(ok.KmerSet.from_sorted_device(k, ptr + 8 * int(b[r]), int(b[r + 1] - b[r])))
            assert np.array_equal(shards[0].to_array(), osets[0][owners == r])      # the ownership rule of the sharded count
            sz, up = ok.all_vs_all_part(shards, 0, 1)
            acc_inter += ok.finish_all_vs_all(sz, up)
            su = ok.KmerSet.union(shards)
            acc_hits += su.probe_reads(reads, off, ok.RAW)
            m, d = ok.probe_counts_many(shards, pk, pc)
            acc_m += m; acc_d += d
            for sh in shards + [su]:
                sh.close()
        assert np.array_equal(acc_inter, inter) and np.array_equal(acc_hits, hits.astype(np.uint64))
        assert np.array_equal(acc_m, m_all) and np.array_equal(acc_d, d_all)
    with pytest.raises(ok.OrionError, match="strictly ascending"):
        import torch
        bad = torch.tensor([5, 9, 9, 12], dtype=torch.int64, device="cuda")
        ok.KmerSet.from_sorted_device(k, bad.data_ptr(), 4)
    # the host entry checks the order on the device as well (after the upload) and still names the offending index
    with pytest.raises(ok.OrionError, match=r"strictly ascending \(index 3\)"):
        ok.KmerSet.from_sorted(k, np.array([1, 5, 9, 9, 12], np.uint64))
    with pytest.raises(ok.OrionError, match=r"strictly ascending \(index 1\)"):
        ok.KmerSet.from_sorted(k, np.array([7, 3], np.uint64))
    assert list(ok.KmerSet.from_sorted(k, np.array([7], np.uint64)).to_array()) == [7]


def test_union_in_groups_matches_oracle(oracle, monkeypatch):
    """db_types.rs:43-48 for more keys than one pass of the partitioned path takes (1,000 genomes are 5e9 keys): groups
    of sets are unified one pass each and the group results folded with the keys-only merge.  The group size is
    shrunk so that the oracle can check it."""
    k = 31
    base = synth.genome(45, 300_000)
    gens = [base, synth.mutate(base, 3, 40_000), synth.genome(46, 50_000), base[:50_000], synth.genome(47, 5_000), base[100_000:250_000]]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    want = oracle.set_union(osets)
    for group in (400_000, 120_000, 10_000_000):
        monkeypatch.setenv("ORION_UNION_GROUP_KEYS", str(group))
        u = ok.KmerSet.union(sets)
        assert np.array_equal(u.to_array(), want), group
        assert u.probe_counts(want[::7], np.ones(len(want[::7]), np.uint64)) == (len(want[::7]), len(want[::7]))
        u.close()
    # foreign k = 32 sets holding u64::MAX, folded
    monkeypatch.setenv("ORION_UNION_GROUP_KEYS", "3")
    a = ok.KmerSet.from_sorted(32, np.array([1, 5, 2 ** 64 - 1], np.uint64))
    b = ok.KmerSet.from_sorted(32, np.array([5, 9], np.uint64))
    c = ok.KmerSet.from_sorted(32, np.array([0, 9, 2 ** 64 - 1], np.uint64))
    assert list(ok.KmerSet.union([a, b, c]).to_array()) == [0, 1, 5, 9, 2 ** 64 - 1]


@pytest.mark.parametrize("base_len", [500_000, 760_000])
def test_union_of_sorted_sets_strided_gather(oracle, monkeypatch, base_len):
    """db_types.rs:43-48 through ONE partitioned pass over the concatenated (sorted) sets.  The level-1 scatter gathers
    its items from 2048 places of the array (k_part_scatter_keys<1, false, 0, true>) and the plan samples single keys;
    the contiguous form (ORION_UNION_NO_STRIDE) and the oracle must give the same set.  The two sizes sit either side
    of the sampling threshold (every key / every 16th key); the totals are odd (a last pair with one key)."""
    k = 31
    base = synth.genome(61, base_len)
    gens = [base] + [synth.mutate(base, 10 + i, 3_000 * (i + 1)) for i in range(5)] + [synth.genome(62, 333_333), base[:77_777]]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    osets = [oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64)) for g in gens]
    total = sum(len(o) for o in osets)
    assert total > (1 << 20) and (total > 16384 * 256) == (base_len > 600_000)
    want = oracle.set_union(osets)
    for no_stride in (False, True):
        if no_stride:
            monkeypatch.setenv("ORION_UNION_NO_STRIDE", "1")
        else:
            monkeypatch.delenv("ORION_UNION_NO_STRIDE", raising=False)
        u = ok.KmerSet.union(sets)
        assert len(u) == len(want) and np.array_equal(u.to_array(), want), no_stride
        u.close()
    monkeypatch.delenv("ORION_UNION_NO_STRIDE", raising=False)
    # an odd number of keys, and the sets in another order
    three = len(osets[3]) + len(osets[7]) + len(osets[0])
    last = osets[6][:len(osets[6]) - (three + len(osets[6]) + 1) % 2]          # makes the total odd
    odd = [sets[3], sets[7], sets[0], ok.KmerSet.from_sorted(k, last)]
    assert sum(len(x) for x in odd) % 2 == 1
    u = ok.KmerSet.union(odd)
    assert np.array_equal(u.to_array(), oracle.set_union([osets[3], osets[7], osets[0], last]))
    for x in sets + [u, odd[3]]:
        x.close()


@pytest.mark.parametrize("mode", ["1", "0"])
def test_query_by_merge_matches_hashed_probe_and_oracle(oracle, monkeypatch, mode):
    """query.rs:83-107 in its two forms: ORION_PROBE_MERGE=1 forces the probe by merge that sets of >= 2^26 keys take
    (distinct k-mers of the batch -> k_member_tiled against the sorted set -> hashed table of the matches -> per-read
    probe), 0 the hashed table of the whole set.  Same integers either way, and the oracle's."""
    import torch
    monkeypatch.setenv("ORION_PROBE_MERGE", mode)
    k = 31
    g = synth.genome(50, 500_000)
    other = synth.genome(51, 500_000)
    kset = ok.KmerSet.from_fastx(k, synth.fasta_text(b"g", g))
    oset = oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64))
    n = 20_000
    bases = np.concatenate([synth.reads(g, 52, n), synth.reads(other, 53, n)])      # 6 M bases: the one-shot count path
    off = synth.read_offsets(2 * n)
    want = oracle.query_hits(oset, k, bases, off, 8)
    assert np.array_equal(kset.probe_reads(bases, off, ok.RAW).astype(np.uint64), want)
    # the device-resident entry (bench `value` leg, sharded query)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    d_h = torch.full((2 * n,), 77, dtype=torch.int32, device="cuda")
    kset.probe_reads_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), 2 * n, d_h.data_ptr(), ok.RAW)
    assert np.array_equal(d_h.cpu().numpy().astype(np.uint64), want)
    # a small batch (table path of the counter), ragged reads incl. shorter than k and empty ones, lower case and U
    rng = np.random.default_rng(4)
    b2, o2 = random_batch(rng, 200_000, 40)
    b2[:100_000] = g[:100_000]
    b2[5000:5100] = np.frombuffer(bytes(g[5000:5100]).lower(), np.uint8)
    b2[np.flatnonzero(b2[:100_000] == ord("T"))[::50]] = ord("U")              # RAW: not a base; NORMALIZED: T
    assert np.array_equal(kset.probe_reads(b2, o2, ok.RAW).astype(np.uint64), oracle.query_hits(oset, k, b2, o2, 4))
    b2t = np.where((b2 == ord("U")) | (b2 == ord("u")), np.uint8(ord("T")), b2)
    assert np.array_equal(kset.probe_reads(b2, o2, ok.NORMALIZED).astype(np.uint64), oracle.query_hits(oset, k, b2t, o2, 4))
    # an empty set / reads shorter than k only
    empty = ok.KmerSet.from_sorted(k, np.zeros(0, np.uint64))
    assert not empty.probe_reads(bases[:15000], off[:101], ok.RAW).any()
    short = np.frombuffer(b"ACGTACGTAC" * 10, np.uint8)
    assert not kset.probe_reads(short, np.arange(0, 101, 10, dtype=np.uint64), ok.RAW).any()
    # a foreign k = 32 set holding u64::MAX: poly-A reads hit key 0 only
    w = ok.KmerSet.from_sorted(32, np.array([0, 5, 2 ** 64 - 1], np.uint64))
    polya = np.frombuffer(b"A" * 40 + b"T" * 40 + b"ACGT" * 10, np.uint8)
    assert list(w.probe_reads(polya, np.array([0, 40, 80, 120], np.uint64), ok.RAW)) == [9, 9, 0]
    for x in (kset, empty, w):
        x.close()


@pytest.mark.parametrize("threads", ["4", "1"])
def test_build_many_matches_one_by_one_and_oracle(oracle, monkeypatch, threads):
    """build.rs:93-116 over many files at once: several host threads, each with its own pooled builder, must give the
    sets the one-file-at-a-time calls give (and the oracle's).  Files of very different sizes (one-shot path, table
    path, empty, shorter than k), multi-record files, host and device entry."""
    import torch
    monkeypatch.setenv("ORION_BUILD_THREADS", threads)
    k = 31
    rng = np.random.default_rng(12)
    files = []
    for i in range(14):
        n = int([2_000_000, 40_000, 1_300_000, 0, 20, 700_000, 5000][i % 7] * (1 + 0.1 * (i // 7)))
        g = synth.genome(300 + i, max(n, 1))[:n]
        cuts = np.unique(rng.integers(0, n + 1, size=3)) if n else np.zeros(0, np.int64)
        off = np.concatenate([[0], cuts, [n]]).astype(np.uint64)
        files.append((np.ascontiguousarray(g, dtype=np.uint8), off))
    want = [oracle.kmer_set_batch(k, b, o) for b, o in files]
    host = ok.KmerSet.build_many(k, files)
    for s, w in zip(host, want):
        assert len(s) == len(w) and np.array_equal(s.to_array(), w)
    # device entry: every file padded to a 16-byte boundary inside one buffer
    starts, at = [], 0
    for b, _ in files:
        starts.append(at)
        at += (len(b) + 15) // 16 * 16
    buf = np.zeros(max(at, 16), np.uint8)
    for (b, _), st in zip(files, starts):
        buf[st:st + len(b)] = b
    d_buf = torch.from_numpy(buf).cuda()
    d_offs = [torch.from_numpy(o.view(np.int64).copy()).cuda() for _, o in files]
    dev = ok.KmerSet.build_many_device(k, [d_buf.data_ptr() + st for st in starts], [len(b) for b, _ in files],
                                       [t.data_ptr() for t in d_offs], [len(o) - 1 for _, o in files])
    for s, w in zip(dev, want):
        assert np.array_equal(s.to_array(), w)
    one = ok.KmerSet.build(k)
    one.add_batch(*files[2])
    assert np.array_equal(one.to_array(), want[2])
    with pytest.raises(ok.OrionError, match="Invalid K-mer size"):
        ok.KmerSet.build_many(33, files[:2])
    for x in host + dev + [one]:
        x.close()


@pytest.mark.parametrize("keyed", ["1", "0"])
def test_all_vs_all_keyed_matches_rows_and_numpy(oracle, monkeypatch, keyed):
    """compare.rs:51-60 for every pair, in its two forms: ORION_AVA_KEYED=1 forces the one-pass keyed form (setops.cuh:
    tiles of the key space, a hashed shared-memory table per tile, bit rows + popcount per pair block), 0 the row-by-row
    merges.  Families of related genomes (keys shared by many sets), an identical copy, a tiny set, an empty set, an
    unrelated genome; 13 sets (not a multiple of the 8 x 8 pair blocks), then key-range slices (a multi-GPU shard:
    a narrow slice of the position space), then 2 sets."""
    monkeypatch.setenv("ORION_AVA_KEYED", keyed)
    k = 21
    b1, b2 = synth.genome(71, 300_000), synth.genome(72, 200_000)
    gens = [b1, synth.mutate(b1, 1, 300), synth.mutate(b1, 2, 3_000), synth.mutate(b1, 3, 15_000), b2, synth.mutate(b2, 4, 2_000),
            synth.mutate(b2, 5, 10_000), b1, b1[:1000], synth.genome(73, 150_000), b2[50_000:120_000], synth.mutate(b1, 6, 100)]
    sets = [ok.KmerSet.from_fastx(k, synth.fasta_text(b"g%d" % i, g)) for i, g in enumerate(gens)]
    sets.insert(6, ok.KmerSet.from_sorted(k, np.zeros(0, np.uint64)))
    arrs = [s.to_array() for s in sets]
    assert np.array_equal(arrs[0], oracle.kmer_set_batch(k, b1, np.array([0, len(b1)], np.uint64)))
    n = len(sets)
    assert n == 13

    def check(handles, arrays):
        sizes, inter = ok.all_vs_all(handles)
        for i in range(len(handles)):
            assert sizes[i] == len(arrays[i]) == inter[i, i]
            for j in range(i + 1, len(handles)):
                want = len(np.intersect1d(arrays[i], arrays[j], assume_unique=True))
                assert inter[i, j] == want == inter[j, i], (keyed, i, j)
    check(sets, arrs)
    # key-range slices: rank 5 of 8 owners, every set cut at the same boundaries
    shards = []
    for s in sets:
        b = s.shard_bounds(8)
        ptr, _ = s.keys_device()
        shards.append(ok.KmerSet.from_sorted_device(k, ptr + 8 * int(b[5]), int(b[6] - b[5])))
    check(shards, [a[int(s.shard_bounds(8)[5]):int(s.shard_bounds(8)[6])] for s, a in zip(sets, arrs)])
    check([sets[0], sets[3]], [arrs[0], arrs[3]])
    # the part form a sharded job uses (one part = everything): upper triangle only, the rest zero
    sz, up = ok.all_vs_all_part(sets, 0, 1)
    assert np.array_equal(ok.finish_all_vs_all(sz, up), ok.all_vs_all(sets)[1]) and not np.tril(up).any()
    for x in sets + shards:
        x.close()


def test_all_vs_all_keyed_falls_back_on_clustered_keys(monkeypatch):
    """keys sharing their first 16 bases share one position: the tile holding them outgrows the shared-memory table,
    the keyed pass reports it and the matrix is computed row by row -- same integers."""
    monkeypatch.setenv("ORION_AVA_KEYED", "1")
    rng = np.random.default_rng(9)
    k = 31
    prefix = np.uint64(0x1234567) << np.uint64(30 + 2)           # the top 32 of the 62 key bits fixed
    pool = np.unique(rng.integers(0, 1 << 30, size=30_000, dtype=np.uint64)) | prefix
    spread = np.unique(np.minimum(rng.integers(0, 1 << 62, size=40_000, dtype=np.uint64), rng.integers(0, 1 << 62, size=40_000, dtype=np.uint64)))
    arrs = []
    for i in range(10):
        pick = pool[rng.random(len(pool)) < 0.5]
        arrs.append(np.unique(np.concatenate([pick, spread[rng.random(len(spread)) < 0.3]])))
    sets = [ok.KmerSet.from_sorted(k, a) for a in arrs]
    sizes, inter = ok.all_vs_all(sets)
    for i in range(10):
        assert sizes[i] == len(arrs[i])
        for j in range(i + 1, 10):
            assert inter[i, j] == len(np.intersect1d(arrs[i], arrs[j], assume_unique=True))
    for x in sets:
        x.close()


def test_intersection_tiled_kernel_shapes():
    """k_intersect_bounds + k_intersect_tiled against numpy on shapes that stress the tiling: equal sets, disjoint
    ranges, a small set inside a large one (one tile of A against hundreds of chunks of B), interleaved keys, tile
    and chunk boundaries, and a k = 32 set that holds the sentinel value u64::MAX."""
    rng = np.random.default_rng(91)

    def uniq(n, lo=0, hi=2 ** 62):
        return np.unique(rng.integers(lo, hi, n, dtype=np.uint64))

    big = uniq(3_000_000)
    cases = [
        (big, big),
        (big[::2].copy(), big[1::2].copy()),                       # interleaved, nothing shared
        (big[:100_000], big),                                       # A is a prefix of B
        (np.sort(rng.choice(big, 20_000, replace=False)), big),     # small A scattered over a large B
        (uniq(50_000, 0, 2 ** 40), uniq(50_000, 2 ** 41, 2 ** 42)),  # disjoint ranges
        (big[:2048 * 5], big[2048 * 3:2048 * 9]),                   # overlaps that start and end on tile boundaries
        (np.concatenate([big[:40_000], np.array([2 ** 64 - 1], np.uint64)]),
         np.concatenate([big[20_000:90_000], np.array([2 ** 64 - 1], np.uint64)])),
    ]
    for a, b in cases:
        k = 32 if (a[-1] == 2 ** 64 - 1 or b[-1] == 2 ** 64 - 1) else 31
        sa, sb = ok.KmerSet.from_sorted(k, a), ok.KmerSet.from_sorted(k, b)
        want = len(np.intersect1d(a, b, assume_unique=True))
        assert sa.intersection_size(sb) == want
        assert sb.intersection_size(sa) == want
        sizes, inter = ok.all_vs_all([sa, sb])
        assert list(sizes) == [len(a), len(b)] and inter[0, 1] == want and inter[1, 0] == want
        sa.close(); sb.close()


# ------------------------------------------------------------------------ query / classify --
def test_query_golden(golden):
    g = golden["query"]
    db = ok.run_build(g["k"], {"db.fa": g["db"].encode()})
    for mh, want in g["ids_by_min_hits"].items():
        ids, hits = ok.run_query(db, g["reads"].encode(), int(mh))
        assert list(hits) == g["hits"]
        assert [i.decode() for i in ids] == want


def test_query_random_against_oracle(oracle):
    k = 31
    g = synth.genome(50, 500_000)
    other = synth.genome(51, 500_000)
    kset = ok.KmerSet.from_fastx(k, synth.fasta_text(b"g", g))
    oset = oracle.kmer_set_batch(k, g, np.array([0, len(g)], np.uint64))
    n = 20_000
    bases = np.concatenate([synth.reads(g, 52, n), synth.reads(other, 53, n)])
    off = synth.read_offsets(2 * n)
    hits = kset.probe_reads(bases, off, ok.RAW)
    want = oracle.query_hits(oset, k, bases, off, 8)
    assert np.array_equal(hits.astype(np.uint64), want)
    # ragged reads incl. shorter than k and empty ones
    rng = np.random.default_rng(4)
    b2, o2 = random_batch(rng, 200_000, 40)
    b2[:100_000] = g[:100_000]
    assert np.array_equal(kset.probe_reads(b2, o2, ok.RAW).astype(np.uint64), oracle.query_hits(oset, k, b2, o2, 4))


def test_classify_golden(golden):
    for case in golden["classify"]["cases"]:
        k = case["k"]
        keys, counts = ok.count_fastx(k, [case["input"].encode()], case["min_kmer_frequency"])
        assert len(keys) == case["total_unique_kmers_in_input"]
        for dbspec in case["databases"]:
            db = ok.run_build(k, {n: c.encode() for n, c in dbspec["refs"].items()})
            union = db.get_all_kmers_unified()
            assert len(union) == dbspec["total_unique_kmers_in_db"]
            assert union.probe_counts(keys, counts) == (dbspec["overall_matched"], dbspec["overall_sum_depth"])
            for name, want in dbspec["per_ref"].items():
                ref = db.references[name]
                assert len(ref) == want["total"]
                assert ref.probe_counts(keys, counts) == (want["matched"], want["sum_depth"]), (case["name"], name)


# ----------------------------------------------- the two count paths give the same table --
@pytest.mark.parametrize("k", [3, 11, 21, 31, 32])
def test_partitioned_path_matches_oracle(oracle, k):
    """path 2 = partition -> shared-memory tables -> sorted runs; path 1 = device-wide table"""
    rng = np.random.default_rng(500 + k)
    for n_bases, mean_len in ((1, 1), (40, 7), (70_000, 150), (3_000_000, 150), (2_500_000, 2_500_000)):
        bases, off = random_batch(rng, n_bases, mean_len)
        wk, wc = oracle.count_batch(k, bases, off)
        for path in (1, 2):
            c = ok.KmerCounter(k)
            c.set_path(path)
            c.add_batch(bases, off)
            st = c.stats()
            assert st["partitioned"] == (1 if path == 2 else 0)
            for mc in (1, 2, 4):
                gk, gc = c.finish(mc)
                sel = wc >= mc
                assert np.array_equal(gk, wk[sel]) and np.array_equal(gc, wc[sel]), (k, n_bases, path, mc)
            assert st["n_windows"] == int(wc.sum())
            c.close()


def test_partitioned_two_batches_merge(oracle):
    """count.rs:48,52-79: ONE table across all input files.  A second (third, ...) large batch is counted on its own
    by the one-shot path and its sorted run merged into the accumulated one (merge.cuh) -- no fold into the
    device-wide table.  Host and device entries, several k, min_count filters, a batch that adds nothing new,
    a batch without a single countable window, and a small batch at the end (that one does go through the table)."""
    import torch
    for k in (31, 21, 32, 11):
        g = synth.genome(61 + k, 800_000)
        n = 40_000
        batches = [synth.reads(g, 62, n), synth.reads(g, 63, n), synth.reads(synth.genome(99, 300_000), 64, n)]
        batches.append(batches[0])                                  # nothing new: every key is a pair in the merge
        batches.append(np.full(n * 150, ord("N"), np.uint8))        # no countable window at all
        off = synth.read_offsets(n)
        c = ok.KmerCounter(k)
        o = oracle.Counter(k)
        for i, b in enumerate(batches):
            if i % 2:
                d_b, d_o = torch.from_numpy(b).cuda(), torch.from_numpy(off.view(np.int64)).cuda()
                c.add_batch_device(d_b.data_ptr(), len(b), d_o.data_ptr(), n)
            else:
                c.add_batch(b, off)
            o.add_batch(b, off)
            st = c.stats()
            assert st["partitioned"] == 1 and st["n_merges"] == i and st["n_slots"] == 0, (k, i, st)
            if i in (1, 3, 4):
                for mc in (1, 2, 5):
                    gk, gc = c.finish(mc)
                    wk, wc = o.finish(mc)
                    assert np.array_equal(gk, wk) and np.array_equal(gc, wc), (k, i, mc)
        assert c.stats()["n_windows"] == int(o.finish(1)[1].sum())
        small = synth.reads(g, 65, 500)                             # below the one-shot threshold: run -> table
        off_s = synth.read_offsets(500)
        c.add_batch(small, off_s)
        o.add_batch(small, off_s)
        assert c.stats()["partitioned"] == 0
        gk, gc = c.finish(1)
        wk, wc = o.finish(1)
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), k
        c.close(); o.close()


def test_large_batch_is_cut_into_sub_batches(oracle, monkeypatch):
    """batches beyond ~1.8 G bases are cut into sub-batches by tile range (32-bit offsets inside a pass), each counted
    on its own and merged.  ORION_MAX_BATCH_BASES shrinks the limit so that the oracle can check it: reads, and one
    long multi-megabase record whose windows straddle every cut (the walk takes its halo from the tile before)."""
    import torch
    monkeypatch.setenv("ORION_MAX_BATCH_BASES", "700000")
    k = 31
    g = synth.genome(66, 3_000_000)
    g[1_000_000:1_000_100] = ord("N")
    n = 30_000
    reads, off = synth.reads(g, 67, n), synth.read_offsets(n)
    cases = [(reads, off), (g, np.array([0, len(g)], np.uint64)),
             (np.concatenate([g[:1_500_000], reads]), np.concatenate([[0], 1_500_000 + off]).astype(np.uint64))]
    for i, (b, o) in enumerate(cases):
        wk, wc = oracle.count_batch(k, b, o)
        for host in (True, False):
            c = ok.KmerCounter(k, capacity_hint=0 if i else 1_000_000)
            if host:
                c.add_batch(b, o)
            else:
                d_b, d_o = torch.from_numpy(b).cuda(), torch.from_numpy(o.view(np.int64)).cuda()
                c.add_batch_device(d_b.data_ptr(), len(b), d_o.data_ptr(), len(o) - 1)
            st = c.stats()
            gk, gc = c.finish(1)
            c.close()
            assert st["partitioned"] == 1 and st["n_merges"] >= 4, st
            assert np.array_equal(gk, wk) and np.array_equal(gc, wc), (i, host)


def test_partitioned_run_then_more_batches(oracle):
    """a second batch on the forced one-shot path is merged into the run; a later batch through the table path
    (set_path(1)) folds the run into the table; results stay exact"""
    g = synth.genome(61, 800_000)
    n = 40_000
    b1, b2 = synth.reads(g, 62, n), synth.reads(g, 63, n)
    off = synth.read_offsets(n)
    c = ok.KmerCounter(31)
    c.set_path(2)
    c.add_batch(b1, off)
    assert c.stats()["partitioned"] == 1
    c.add_batch(b2[:n * 75], off[:n // 2 + 1])
    assert c.stats()["partitioned"] == 1 and c.stats()["n_merges"] == 1
    c.set_path(1)
    c.add_batch(b2[n * 75:], off[:n // 2 + 1])
    assert c.stats()["partitioned"] == 0
    o = oracle.Counter(31)
    o.add_batch(b1, off)
    o.add_batch(b2, off)
    for mc in (1, 3):
        gk, gc = c.finish(mc)
        wk, wc = o.finish(mc)
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    c.set_path(2)
    c.clear()
    c.add_batch(b2, off)
    gk, gc = c.finish()
    wk, wc = oracle.count_batch(31, b2, off)
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    c.close()


def test_partitioned_heavy_duplicates_and_low_complexity(oracle):
    """poly-A / dinucleotide reads: millions of windows on a handful of keys"""
    n = 20_000
    reads = np.tile(np.frombuffer(b"A" * 150, np.uint8), n // 2)
    reads2 = np.tile(np.frombuffer(b"AC" * 75, np.uint8), n // 2)
    bases = np.concatenate([reads, reads2, synth.reads(synth.genome(1, 100_000), 2, 5000)])
    off = synth.read_offsets(n + 5000)
    wk, wc = oracle.count_batch(21, bases, off)
    for path in (1, 2):
        c = ok.KmerCounter(21)
        c.set_path(path)
        c.add_batch(bases, off)
        gk, gc = c.finish()
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), path
        c.close()


def test_microsatellite_reads_with_errors_stay_exact(oracle):
    """2 % of a genome in microsatellites ((A)n, (AC)n, (AAT)n ... runs of 200-2000 bases) and 0.5 % substitution
    errors: a handful of sub-partitions receive hundreds of thousands of windows whose distinct k-mers share their
    first 16 bases.  They are deferred to the generic count kernel (hashed table + bitonic sort); round 1's monotone
    table spilled every one of those windows and the spill list overflowed."""
    rng = np.random.default_rng(8)
    g = synth.genome(8, 4_000_000)
    motifs = [b"A", b"AC", b"AAT", b"AG", b"T"]
    done = 0
    while done < len(g) * 0.02:
        ln = int(rng.integers(200, 2000))
        p = int(rng.integers(0, len(g) - ln))
        m = motifs[int(rng.integers(0, len(motifs)))]
        g[p:p + ln] = np.frombuffer((m * (ln // len(m) + 1))[:ln], np.uint8)
        done += ln
    n = 800_000
    bases, off = synth.reads(g, 9, n), synth.read_offsets(n)
    wk, wc = oracle.count_batch_mt(31, bases, off, 8)
    for hint in (int(len(bases) * 0.17), 0):
        c = ok.KmerCounter(31, capacity_hint=hint)
        c.add_batch(bases, off)
        st = c.stats()
        gk, gc = c.finish()
        c.close()
        assert st["partitioned"] == 1 and st["n_spilled"] == 0, st
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), hint
    assert int(wc.max()) > 10_000           # the hot k-mers are there
    # deeper coverage: the poly-A sub-partition now holds more distinct error variants than one shared-memory table
    # takes (6144): the excess is spilled, counted as a batch of its own and MERGED into the run (round 1 folded the
    # run into the device-wide ordered table, which cannot hold thousands of keys on one home slot, and failed)
    import os
    n = 2_400_000
    bases, off = synth.reads(g, 10, n), synth.read_offsets(n)
    wk, wc = oracle.count_batch_ranged_mt(31, bases, off, min(os.cpu_count() or 1, 16))
    c = ok.KmerCounter(31, capacity_hint=int(len(bases) * 0.17))
    c.add_batch(bases, off)
    st = c.stats()
    gk, gc = c.finish()
    c.close()
    assert st["partitioned"] == 1, st
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc), st


def test_capacity_hint_sizes_sub_partitions_and_a_wrong_hint_stays_exact(oracle):
    """With a capacity hint the sub-partitions are sized for their expected DISTINCT keys (several windows per
    table slot).  A hint far below the truth makes every shared-memory table overflow: the sub-partitions are
    deferred / spilled / recounted, slower but with the same table."""
    rng = np.random.default_rng(81)
    bases, off = random_batch(rng, 6_000_000, 150)            # random bases: nearly every window is distinct
    wk, wc = oracle.count_batch(31, bases, off)
    for hint in (50_000, 3_000_000, 10_000_000):
        c = ok.KmerCounter(31, capacity_hint=hint)
        c.set_path(2)
        c.add_batch(bases, off)
        gk, gc = c.finish()
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), hint
        c.close()
    # 40x coverage with a good hint: ~10 windows per distinct key, large sub-partitions, nothing deferred
    g = synth.genome(82, 300_000)
    n_reads = 80_000
    b2, o2 = synth.reads(g, 83, n_reads), synth.read_offsets(n_reads)
    wk, wc = oracle.count_batch(31, b2, o2)
    c = ok.KmerCounter(31, capacity_hint=int(len(wk) * 1.1))
    c.set_path(2)
    c.add_batch(b2, o2)
    gk, gc = c.finish()
    st = c.stats()
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    assert st["partitioned"] == 1 and st["n_deferred"] == 0
    c.close()


# ------------------------------------------- deferred host batches: the sliced result pipeline --
def _sliced_workload():
    g = synth.genome(71, 4_000_000)
    n_reads = 160_000                      # 24 M bases: 8192 sub-partitions, 8 result slices
    return synth.reads(g, 72, n_reads), synth.read_offsets(n_reads)


def test_sliced_result_pipeline_matches_oracle(oracle, monkeypatch):
    """ok_counter_add_batch defers level 2 + count of a large host batch; ok_counter_finish runs them slice
    by slice under the D2H copy.  Every way out of the deferred state gives the oracle's table."""
    bases, off = _sliced_workload()
    wk, wc = oracle.count_batch(31, bases, off)
    windows = int(wc.sum())
    # (a) straight through the pipeline
    c = ok.KmerCounter(31)
    c.add_batch(bases, off)
    gk, gc = c.finish()
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    st = c.stats()
    assert st["partitioned"] == 1 and st["n_windows"] == windows and st["n_distinct"] == len(wk)
    # the counter still answers afterwards: device result, filtered result
    _, _, n = c.finish_device(1)
    assert n == len(wk)
    gk, gc = c.finish(3)
    assert np.array_equal(gk, wk[wc >= 3]) and np.array_equal(gc, wc[wc >= 3])
    # (b) clear and reuse: stats first (settles the deferred batch), then finish
    c.clear()
    c.add_batch(bases, off)
    assert c.stats()["n_windows"] == windows
    gk, gc = c.finish()
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    # (c) min_count > 1 straight from the deferred state
    c.clear()
    c.add_batch(bases, off)
    gk, gc = c.finish(2)
    assert np.array_equal(gk, wk[wc >= 2]) and np.array_equal(gc, wc[wc >= 2])
    # (d) a deferred batch that is dropped
    c.clear()
    c.add_batch(bases, off)
    c.clear()
    assert c.finish()[0].size == 0
    # (e) a second batch arrives while the first is deferred: both are counted
    c.add_batch(bases, off)
    c.add_batch(bases[:150 * 1000], off[:1001])
    gk, gc = c.finish()
    o = oracle.Counter(31)
    o.add_batch(bases, off)
    o.add_batch(bases[:150 * 1000], off[:1001])
    ek, ec = o.finish(1)
    assert np.array_equal(gk, ek) and np.array_equal(gc, ec)
    c.close()
    # (f) the undeferred path (what add_batch did before) agrees
    monkeypatch.setenv("ORION_NO_DEFER", "1")
    c = ok.KmerCounter(31)
    c.add_batch(bases, off)
    gk, gc = c.finish()
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    c.close()


def test_sliced_result_pipeline_misjudged_size_falls_back(oracle):
    """The result buffers are sized from the first key-range slice.  A batch with (almost) nothing in that
    slice -- no AA / TT dinucleotides, so no canonical k-mer starts with AA -- overflows the estimate; the
    pipeline must notice and ship the exact-size result instead."""
    rng = np.random.default_rng(73)
    n = 24_000_000
    seq = rng.integers(0, 4, n).astype(np.uint8)
    for _ in range(40):                       # break up AA (0,0) and TT (3,3) pairs
        bad = np.flatnonzero((seq[1:] == seq[:-1]) & ((seq[1:] == 0) | (seq[1:] == 3))) + 1
        if bad.size == 0:
            break
        seq[bad] = rng.integers(1, 3, bad.size)
    bases = np.frombuffer(b"ACGT", np.uint8)[seq]
    off = np.arange(0, n + 1, 150, dtype=np.uint64)
    wk, wc = oracle.count_batch(31, bases, off)
    c = ok.KmerCounter(31)
    c.add_batch(bases, off)
    gk, gc = c.finish()
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
    c.close()


# ------------------------------------------------- multi-GPU route / shard logic on one GPU --
@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_route_and_sharded_counters_concatenate_to_the_global_table(oracle, n_ranks):
    import torch
    k = 31
    g = synth.genome(70, 1_000_000)
    n = 60_000
    bases = synth.reads(g, 71, n)
    off = synth.read_offsets(n)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    d_out = torch.empty(len(bases), dtype=torch.int64, device="cuda")
    router = ok.KmerCounter(k)
    counts = router.route_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), n, n_ranks, d_out.data_ptr())
    router.close()
    wk, wc = oracle.count_batch(k, bases, off)
    assert int(counts.sum()) == int(wc.sum())
    routed = d_out[:int(counts.sum())].cpu().numpy().view(np.uint64)
    owners = np.zeros(len(routed), np.int32)
    ok._check(ok.lib().okx_owner_of(ok._ptr(routed), len(routed), k, n_ranks, ok._ptr(owners)))
    bounds = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pieces_k, pieces_c = [], []
    for r in range(n_ranks):
        assert np.all(owners[bounds[r]:bounds[r + 1]] == r)
        for path in (1, 2):
            c = ok.KmerCounter(k)
            c.set_shard(r, n_ranks)
            c.set_path(path)
            c.add_kmers_device(d_out.data_ptr() + 8 * int(bounds[r]), int(counts[r]))
            gk, gc = c.finish()
            c.close()
            if path == 1:
                pieces_k.append(gk); pieces_c.append(gc)
            else:
                assert np.array_equal(gk, pieces_k[-1]) and np.array_equal(gc, pieces_c[-1]), (r, "paths differ")
    # disjoint key ranges in rank order: plain concatenation is the sorted global table
    assert np.array_equal(np.concatenate(pieces_k), wk)
    assert np.array_equal(np.concatenate(pieces_c), wc)


# --------------------------- fused multi-GPU exchange (sharded scatter), all ranks emulated on one GPU --
def _sharded_dance(counters, batches, n):
    """one sharded count step of every emulated rank: sample, 'exchange' the histograms (torch ops stand in for
    the reduce-scatter / all-gathers), scatter into the owners' level-1 regions, count what arrived"""
    import torch
    n_ranks = len(counters)
    dev = [(torch.from_numpy(b).cuda(), torch.from_numpy(o.view(np.int64)).cuda()) for b, o in batches]
    nmax = max(len(b) for b, _ in batches)
    geoms = [c.shard_geometry(nmax) for c in counters]
    assert len(set(geoms)) == 1
    sub_bits, l1_bits, cap = geoms[0]
    bufs = [ok.PeerBuffer(cap * 8) for _ in range(n_ranks)]
    try:
        for c in counters:
            c.shard_set_buffers([b.ptr for b in bufs], cap)
        i32 = dict(dtype=torch.int32, device="cuda")
        hist_fine = [torch.empty(n_ranks << sub_bits, **i32) for _ in range(n_ranks)]
        hist_l1 = [torch.empty(n_ranks << l1_bits, **i32) for _ in range(n_ranks)]
        for r, c in enumerate(counters):
            c.shard_sample_device(dev[r][0].data_ptr(), len(batches[r][0]), dev[r][1].data_ptr(), n,
                                  hist_fine[r].data_ptr(), hist_l1[r].data_ptr())
        hist_sum = torch.stack(hist_fine).sum(0, dtype=torch.int32)
        l1_all = torch.cat(hist_l1).contiguous()
        cursors = [torch.empty(n_ranks << l1_bits, **i32) for _ in range(n_ranks)]
        for r, c in enumerate(counters):
            mine = hist_sum[r << sub_bits:(r + 1) << sub_bits].contiguous()
            c.shard_scatter_device(dev[r][0].data_ptr(), len(batches[r][0]), dev[r][1].data_ptr(), n,
                                   mine.data_ptr(), l1_all.data_ptr(), cursors[r].data_ptr())
        cur_all = torch.cat(cursors).contiguous()
        keys, counts = [], []
        for r, c in enumerate(counters):
            c.shard_count_device(cur_all.data_ptr())
            c.commit_batch()
            gk, gc = c.finish()
            assert c.stats()["n_spilled"] == 0
            keys.append(gk); counts.append(gc)
        return np.concatenate(keys), np.concatenate(counts), sub_bits
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            b.destroy()


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
@pytest.mark.parametrize("hint", [0, 600_000])
def test_sharded_scatter_matches_oracle(oracle, n_ranks, hint):
    """ok_shard_*: every rank samples, the histograms are exchanged, every sender scatters into the owners'
    level-1 regions, every owner counts what arrived.  The ranks' outputs must concatenate to the oracle's
    table of all batches.  hint: sub-partitions sized for their expected distinct keys."""
    k = 31
    g = synth.genome(80, 400_000)
    n = 12_000                                   # reads per rank: 1.8 M bases -> partitioned geometry
    batches = [(synth.reads(g, 81, n, first_read=r * n), synth.read_offsets(n)) for r in range(n_ranks)]
    counters = [ok.KmerCounter(k, capacity_hint=hint) for _ in range(n_ranks)]
    for r, c in enumerate(counters):
        c.set_shard(r, n_ranks)
    gk, gc, _ = _sharded_dance(counters, batches, n)
    for c in counters:
        c.close()
    all_bases = np.concatenate([b for b, _ in batches])
    all_off = synth.read_offsets(n * n_ranks)
    wk, wc = oracle.count_batch(k, all_bases, all_off)
    assert np.array_equal(gk, wk)
    assert np.array_equal(gc, wc)


def _xchg_dance(counters, batches, n, stepwise=False, abort=False):
    """one step of the chunked exchange (ok_xchg_*) for every emulated rank; torch ops stand in for the collectives"""
    import torch
    n_ranks = len(counters)
    dev = [(torch.from_numpy(b).cuda(), torch.from_numpy(o.view(np.int64)).cuda()) for b, o in batches]
    nmax = max(len(b) for b, _ in batches)
    geoms = [c.xchg_geometry(nmax) for c in counters]
    assert len(set(geoms)) == 1
    sub_bits, l1_bits, n_chunks, cap = geoms[0]
    bufs = [ok.PeerBuffer(cap * 8) for _ in range(n_ranks)]
    try:
        for c in counters:
            c.shard_set_buffers([b.ptr for b in bufs], cap)
        i32 = dict(dtype=torch.int32, device="cuda")
        hist_fine = [torch.empty(n_ranks << sub_bits, **i32) for _ in range(n_ranks)]
        hist_l1c = [torch.empty(n_chunks * (n_ranks << l1_bits), **i32) for _ in range(n_ranks)]
        for r, c in enumerate(counters):
            c.xchg_sample_device(dev[r][0].data_ptr(), len(batches[r][0]), dev[r][1].data_ptr(), n,
                                 hist_fine[r].data_ptr(), hist_l1c[r].data_ptr())
        hist_sum = torch.stack(hist_fine).sum(0, dtype=torch.int32)
        l1c_all = torch.cat(hist_l1c).cpu().numpy().view(np.uint32)
        mine = [hist_sum[r << sub_bits:(r + 1) << sub_bits].contiguous() for r in range(n_ranks)]
        if stepwise:        # the receive pipeline of every chunk as soon as all ranks have sent it (a loop stands in for the host barrier)
            for r, c in enumerate(counters):
                c.xchg_scatter_begin(dev[r][0].data_ptr(), len(batches[r][0]), dev[r][1].data_ptr(), n, mine[r].data_ptr(), l1c_all)
            for ch in range(n_chunks):
                for c in counters:
                    c.xchg_chunk_sent(ch)
                for c in counters:
                    c.xchg_chunk_recv(ch)
            for c in counters:
                c.xchg_scatter_end()
        else:
            for r, c in enumerate(counters):
                c.xchg_scatter_device(dev[r][0].data_ptr(), len(batches[r][0]), dev[r][1].data_ptr(), n, mine[r].data_ptr(), l1c_all)
        keys, counts = [], []
        for r, c in enumerate(counters):
            c.xchg_count_device()
        for r, c in enumerate(counters):
            if abort:       # some rank failed (here: pretend): every rank drops its share of THIS batch, earlier batches stay
                c.abort_batch()
            else:
                c.commit_batch()
            gk, gc = c.finish()
            assert c.stats()["n_spilled"] == 0
            keys.append(gk); counts.append(gc)
        return np.concatenate(keys), np.concatenate(counts), n_chunks
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            b.destroy()


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
@pytest.mark.parametrize("hint", [0, 600_000])
@pytest.mark.parametrize("chunks", [8, 1, 3])
@pytest.mark.parametrize("levels", [2, 3])
def test_chunked_exchange_matches_oracle(oracle, monkeypatch, n_ranks, hint, chunks, levels):
    """ok_xchg_*: per-chunk sub-blocks moved by plain peer copies, fills carried in the sub-block headers; once in one
    call, once chunk by chunk with the receive work issued as the chunks land.  levels = 2: the sender splits by (owner,
    level-1 bin), the owner runs level 2; levels = 3: the sender splits by owner only, the owner runs both levels."""
    monkeypatch.setenv("ORION_XCHG_CHUNKS", str(chunks))
    monkeypatch.setenv("ORION_XCHG_LEVELS", str(levels))
    k = 31
    g = synth.genome(85, 400_000)
    n = 12_000
    batches = [(synth.reads(g, 86, n, first_read=r * n), synth.read_offsets(n)) for r in range(n_ranks)]
    counters = [ok.KmerCounter(k, capacity_hint=hint) for _ in range(n_ranks)]
    for r, c in enumerate(counters):
        c.set_shard(r, n_ranks)
    wk, wc = oracle.count_batch(k, np.concatenate([b for b, _ in batches]), synth.read_offsets(n * n_ranks))
    for rep in range(2):
        for c in counters:
            c.clear()
        gk, gc, used = _xchg_dance(counters, batches, n, stepwise=rep == 1)
        assert used == chunks
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), rep
    for c in counters:
        c.close()


def test_chunked_exchange_second_batch_merges_into_the_shards(oracle):
    """config 3 at 2 / 4 GPUs is several batches per rank: the second sharded batch is merged into every rank's shard"""
    k, n_ranks, n = 31, 4, 12_000
    g = synth.genome(87, 400_000)
    rounds = [[(synth.reads(g, 88 + i, n, first_read=r * n), synth.read_offsets(n)) for r in range(n_ranks)] for i in range(3)]
    counters = [ok.KmerCounter(k) for _ in range(n_ranks)]
    for r, c in enumerate(counters):
        c.set_shard(r, n_ranks)
    for i, batches in enumerate(rounds):
        if i == 1:          # a batch that is counted and then aborted leaves the table of the earlier batches
            gk, gc, _ = _xchg_dance(counters, rounds[2], n, abort=True)
            wk, wc = oracle.count_batch(k, np.concatenate([b for b, _ in rounds[0]]), synth.read_offsets(n * n_ranks))
            assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
            assert sum(c.stats()["n_windows"] for c in counters) == int(wc.sum())
        gk, gc, _ = _xchg_dance(counters, batches, n, stepwise=i == 1)
        assert all(c.stats()["n_merges"] == i for c in counters)
        all_bases = np.concatenate([b for rd in rounds[:i + 1] for b, _ in rd])
        wk, wc = oracle.count_batch(k, all_bases, synth.read_offsets(n * n_ranks * (i + 1)))
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), i
    for c in counters:
        c.close()


def test_sharded_count_with_a_hint_far_too_low_asks_for_a_recount(oracle):
    """the shared-memory tables overflow, the keys came from the peers: the rank reports it, the caller clears,
    drops the hint on every rank and counts the batch again"""
    k, n_ranks, n = 31, 2, 12_000
    rng = np.random.default_rng(84)           # random bases: every window distinct, ~14 K keys per sub-partition
    batches = [(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n * 150)], synth.read_offsets(n)) for r in range(n_ranks)]
    counters = [ok.KmerCounter(k, capacity_hint=2_000) for _ in range(n_ranks)]
    for r, c in enumerate(counters):
        c.set_shard(r, n_ranks)
    with pytest.raises(ok.OrionError, match="capacity hint too low"):
        _sharded_dance(counters, batches, n)
    for c in counters:
        c.clear()
        c.set_capacity_hint(0)
    gk, gc, _ = _sharded_dance(counters, batches, n)
    for c in counters:
        c.close()
    wk, wc = oracle.count_batch(k, np.concatenate([b for b, _ in batches]), synth.read_offsets(n * n_ranks))
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)


def test_large_table_count_kernel_variant(oracle, monkeypatch):
    """k_part_count<14> (16384 slots, sub-partitions up to 12288 keys, one CTA per SM) is what 8-GPU routing and
    batches beyond ~1.5 G bases use; force it on a batch the oracle can check."""
    monkeypatch.setenv("ORION_BIG_COUNT", "1")
    g = synth.genome(95, 300_000)
    n = 40_000
    bases, off = synth.reads(g, 96, n), synth.read_offsets(n)
    wk, wc = oracle.count_batch(31, bases, off)
    c = ok.KmerCounter(31)
    c.set_path(2)
    c.add_batch(bases, off)
    st = c.stats()
    gk, gc = c.finish()
    c.close()
    assert st["partitioned"] == 1 and st["n_spilled"] == 0
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)
