"""The reference's integration tests (tests/{count,build,compare,query,classify}_tests.rs spawn the
binary through assert_cmd) replayed against `orion-kmer-b200`, the same command line over
liborion_gpu.so: same flags, same output files, same error texts and exit codes."""
import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import orion_kmer_b200 as ok

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "fixtures")


@pytest.fixture(scope="module")
def exe():
    ok.build_host()
    p = ok.cli_path()
    if not os.path.exists(p):
        pytest.fail("orion-kmer-b200 is not built (ok.build_host())")
    return p


def run(exe, *args, ok_exit=True):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)
    if ok_exit:
        assert r.returncode == 0, r.stderr
    return r


def tsv(path):
    text = ok.read_file(str(path)).decode()
    rows = [ln.split("\t") for ln in text.splitlines()]
    return {k: int(c) for k, c in rows}, [k for k, _ in rows]


def kset(strings, k):
    """canonical_u64(seq_to_u64(s)) of every string, the way build_tests.rs:116-120 builds its expectation"""
    out = set()
    for s in strings:
        assert len(s) == k
        v = rc = 0
        for i, ch in enumerate(s):
            c = "ACGT".index(ch)
            v = v * 4 + c
            rc |= (3 - c) << (2 * i)
        out.add(min(v, rc))
    return out


def write(tmp, name, text):
    p = tmp / name
    p.write_text(text)
    return p


# ------------------------------------------------------------------------------------- count --
def test_count_cases(exe, golden, tmp_path):
    files = {n: write(tmp_path, n, t) for n, t in golden["count"]["files"].items()}
    for case in golden["count"]["cases"]:
        out = tmp_path / f"{case['name']}.tsv"
        run(exe, "count", "-k", case["k"], "-i", *[files[n] for n in case["inputs"]], "-o", out, "-m", case["min_count"])
        got, order = tsv(out)
        assert got == case["expected"], case["name"]
        assert order == sorted(order)                       # count.rs:119: ascending k-mer value == ACGT order


def test_count_fixture_files_every_codec_and_compressed_output(exe, golden, tmp_path):
    d = golden["derived_from_src"]
    for ext in ("gz", "xz", "zst"):
        out = tmp_path / f"k7.{ext}.tsv.{ext}"               # output codec by extension too (utils.rs:167-199)
        run(exe, "count", "--kmer-size", 7, "--input-files", os.path.join(FIX, f"test_input1.fasta.{ext}"), "--output-file", out)
        assert tsv(out)[0] == d["input1_k7"]
        out6 = tmp_path / f"k6.{ext}.tsv"
        run(exe, "-v", "count", "-k6", "-i", os.path.join(FIX, f"test_input2.fastq.{ext}"), "-o", out6, "-t", "2")
        assert tsv(out6)[0] == d["input2_k6"]
    assert gzip.open(tmp_path / "k7.gz.tsv.gz").read().decode().splitlines()[0] == "ACGTACG\t4"


def test_count_errors(exe, tmp_path):
    f = write(tmp_path, "a.fa", ">s\nACGT\n")
    for k in (0, 33):                                        # count_tests.rs:296-331
        r = run(exe, "count", "-k", k, "-i", f, "-o", tmp_path / "o.tsv", ok_exit=False)
        assert r.returncode == 1 and f"Invalid K-mer size: {k}. Must be between 1 and 32." in r.stderr
    r = run(exe, "count", "-k", 3, "-i", tmp_path / "missing.fa", "-o", tmp_path / "o.tsv", ok_exit=False)
    assert r.returncode == 1 and "Failed to get input reader for file" in r.stderr
    bad = write(tmp_path, "bad.fa", "this is not fasta\n")
    r = run(exe, "count", "-k", 3, "-i", bad, "-o", tmp_path / "o.tsv", ok_exit=False)
    assert r.returncode == 1 and "Failed to parse FASTA/Q content from" in r.stderr
    r = run(exe, "count", "-k", 3, "-o", tmp_path / "o.tsv", ok_exit=False)       # clap: missing required argument
    assert r.returncode == 2 and "required" in r.stderr
    r = run(exe, "count", "-k", 300, "-i", f, "-o", tmp_path / "o.tsv", ok_exit=False)   # -k is a u8
    assert r.returncode == 2


# ------------------------------------------------------------------------------------- build --
def test_build_cases(exe, golden, tmp_path):
    for case in golden["build"]["cases"]:
        d = tmp_path / case["name"]
        d.mkdir()
        paths = [write(d, n, t) for n, t in case["files"].items()]
        db = d / "out.db"
        run(exe, "build", "-k", case["k"], "-g", *paths, "-o", db)
        k, refs = ok.read_kmer_db(str(db))
        assert k == case["k"] and set(refs) == set(case["expected"])          # reference name = file basename (build.rs:106-109)
        for name, want in case["expected"].items():
            assert set(refs[name].tolist()) == kset(want, case["k"]), (case["name"], name)
        assert len(set().union(*[set(v.tolist()) for v in refs.values()])) == case["total_unique"]


def test_build_sniffs_gzip_and_xz_but_not_zstd(exe, golden, tmp_path):
    want = kset(golden["derived_from_src"]["input1_k7"].keys(), 7)
    db = tmp_path / "mix.db.gz"
    run(exe, "build", "-k", 7, "-g", os.path.join(FIX, "test_input1.fasta.gz"), os.path.join(FIX, "test_input1.fasta.xz"), "-o", db)
    k, refs = ok.read_kmer_db(str(db))
    assert k == 7 and set(refs) == {"test_input1.fasta.gz", "test_input1.fasta.xz"}
    assert all(set(v.tolist()) == want for v in refs.values())
    # build reads RAW bytes and lets needletail sniff; needletail 0.5.1 has no zstd -> not a FASTA/Q start byte
    r = run(exe, "build", "-k", 7, "-g", os.path.join(FIX, "test_input1.fasta.zst"), "-o", tmp_path / "z.db", ok_exit=False)
    assert r.returncode == 1 and "Failed to parse FASTA/Q content from" in r.stderr


# ----------------------------------------------------------------------------------- compare --
def test_compare_cases(exe, golden, tmp_path):
    for case in golden["compare"]["cases"]:
        d = tmp_path / case["name"]
        d.mkdir()
        run(exe, "build", "-k", case["k"], "-g", write(d, "one.fa", case["db1"]), "-o", d / "one.db")
        run(exe, "build", "-k", case.get("k2", case["k"]), "-g", write(d, "two.fa", case["db2"]), "-o", d / "two.db.xz")
        out = d / "cmp.json"
        r = run(exe, "compare", "--db1", d / "one.db", "--db2", d / "two.db.xz", "-o", out, ok_exit=False)
        if "error" in case:
            assert r.returncode == 1 and case["error"] in r.stderr
            continue
        assert r.returncode == 0, r.stderr
        text = out.read_text()
        j = json.loads(text)
        assert list(j) == ["db1_path", "db2_path", "kmer_size", "db1_total_unique_kmers_across_references",
                           "db2_total_unique_kmers_across_references", "intersection_size", "union_size", "jaccard_index"]
        assert (j["db1_total_unique_kmers_across_references"], j["db2_total_unique_kmers_across_references"],
                j["intersection_size"], j["union_size"]) == (case["db1_size"], case["db2_size"], case["intersection_size"], case["union_size"])
        assert abs(j["jaccard_index"] - case["jaccard"]) < 1e-6            # compare_tests.rs:98-108
        assert text.startswith('{\n  "db1_path": ') and isinstance(j["jaccard_index"], float)
        assert '"jaccard_index": 1.0' in text or j["jaccard_index"] != 1.0   # serde_json writes 1.0, not 1


def test_compare_k_mismatch_message(exe, tmp_path):
    fa = write(tmp_path, "g.fa", ">s\nACGTACGTAA\n")
    run(exe, "build", "-k", 3, "-g", fa, "-o", tmp_path / "k3.db")
    run(exe, "build", "-k", 4, "-g", fa, "-o", tmp_path / "k4.db")
    r = run(exe, "compare", "--db1", tmp_path / "k3.db", "--db2", tmp_path / "k4.db", "-o", tmp_path / "o.json", ok_exit=False)
    assert r.returncode == 1
    assert "K-mer databases have incompatible k-mer sizes (overall comparison): 3 vs 4" in r.stderr   # compare_tests.rs:216-218


# ------------------------------------------------------------------------------------- query --
def test_query_golden(exe, golden, tmp_path):
    q = golden["query"]
    run(exe, "build", "-k", q["k"], "-g", write(tmp_path, "ref.fa", q["db"]), "-o", tmp_path / "ref.db")
    reads = write(tmp_path, "reads.fq", q["reads"])
    for min_hits, ids in q["ids_by_min_hits"].items():
        out = tmp_path / f"hits_{min_hits}.txt"
        run(exe, "query", "-d", tmp_path / "ref.db", "-r", reads, "-o", out, "-c", min_hits)
        assert out.read_text().splitlines() == ids, min_hits          # input order, id without the '@' (query_tests.rs:141-143)
    out = tmp_path / "default.txt.gz"
    run(exe, "query", "--database", tmp_path / "ref.db", "--reads", reads, "--output-file", out)
    assert gzip.open(out).read().decode().splitlines() == q["ids_by_min_hits"]["1"]


# ---------------------------------------------------------------------------------- classify --
def test_classify_cases(exe, golden, tmp_path):
    for case in golden["classify"]["cases"]:
        d = tmp_path / case["name"]
        d.mkdir()
        dbs = []
        for i, dbc in enumerate(case["databases"]):
            refs = [write(d, n, t) for n, t in dbc["refs"].items()]
            run(exe, "build", "-k", case["k"], "-g", *refs, "-o", d / f"db{i}.db")
            dbs.append(d / f"db{i}.db")
        inp = write(d, "input.fx", case["input"])
        out, out_tsv = d / "cls.json", d / "cls.tsv"
        run(exe, "classify", "-i", inp, "-d", *dbs, "-o", out, "--min-kmer-frequency", case["min_kmer_frequency"],
            "--output-tsv", out_tsv)
        j = json.loads(out.read_text())
        assert j["total_unique_kmers_in_input"] == case["total_unique_kmers_in_input"]
        assert j["min_kmer_frequency_filter"] == case["min_kmer_frequency"]
        rows = [ln.split("\t") for ln in out_tsv.read_text().splitlines()]
        assert rows[0][:3] == ["InputFile", "Database", "Reference"]
        for dbc, got in zip(case["databases"], j["databases_analyzed"]):
            assert got["total_unique_kmers_in_db_across_references"] == dbc["total_unique_kmers_in_db"]
            assert got["overall_input_kmers_matched_in_db"] == dbc["overall_matched"]
            assert got["overall_sum_depth_of_matched_kmers_in_input"] == dbc["overall_sum_depth"]
            by_name = {r["reference_name"]: r for r in got["references"]}      # located by name (classify_tests.rs:224-229)
            assert set(by_name) == set(dbc["per_ref"])
            for name, want in dbc["per_ref"].items():
                r = by_name[name]
                assert (r["total_kmers_in_reference"], r["input_kmers_hitting_reference"], r["sum_depth_of_matched_kmers_in_input"]) == \
                       (want["total"], want["matched"], want["sum_depth"])
                if want["total"]:
                    assert abs(r["reference_breadth_of_coverage"] - want["matched"] / want["total"]) < 1e-12
                row = [x for x in rows if x[2] == name][0]
                assert row[3:6] == [str(want["total"]), str(want["matched"]), str(want["sum_depth"])]
                assert row[8] == f"{r['reference_breadth_of_coverage']:.4f}"     # classify.rs:373-375


def test_classify_k_validation_and_min_coverage(exe, tmp_path):
    fa = write(tmp_path, "r.fa", ">s\nACGTACGTAACC\n")
    run(exe, "build", "-k", 4, "-g", fa, "-o", tmp_path / "k4.db")
    run(exe, "build", "-k", 5, "-g", fa, "-o", tmp_path / "k5.db")
    r = run(exe, "classify", "-i", fa, "-d", tmp_path / "k4.db", "-o", tmp_path / "o.json", "-k", 5, ok_exit=False)
    assert r.returncode == 1 and "User-provided k-mer size 5 does not match k-mer size 4 from database" in r.stderr
    r = run(exe, "classify", "-i", fa, "-d", tmp_path / "k4.db", tmp_path / "k5.db", "-o", tmp_path / "o.json", ok_exit=False)
    assert r.returncode == 1 and "Effective k-mer size 4 (from first database) does not match k-mer size 5" in r.stderr
    other = write(tmp_path, "other.fa", ">t\nGGGGGGGGCCCCAAAT\n")
    run(exe, "build", "-k", 4, "-g", fa, other, "-o", tmp_path / "two.db")
    run(exe, "classify", "-i", fa, "-d", tmp_path / "two.db", "-o", tmp_path / "cov.json", "--min-coverage", 0.9)
    names = [r["reference_name"] for r in json.loads((tmp_path / "cov.json").read_text())["databases_analyzed"][0]["references"]]
    assert names == ["r.fa"]                                                      # classify.rs:247 filters on breadth
