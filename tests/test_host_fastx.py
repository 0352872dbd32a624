"""The product's host framing (liborion_host.so) against the oracle's needletail restatement."""
import numpy as np
import pytest

import orion_kmer_b200 as ok

CASES = [
    b">seq1\nACGTACGTACGT\n>seq2\nTTTTCCCCGGGGAAAA\n>seq3\nAgCtAgCtNaCcGgTt",
    b"@read1\nGATTACA\n+\n!!!!!!!\n@read2\nTACATACA\n+\n!!!!!!!!\n@read3\natatatNnN\n+\n!!!!!!!!!",
    b">seq1\nACGTACGTACGT\n>seq2\nTGCATGCATGCANNNACGT\n>seq3\nGATTACA\nNNNNN\nGATTACA\n",
    b"@read1\nCGTACGTACG\n+\nFFFFFFFFJJ\n@read3 NNN\nGATTACANNN\n+\nFFFFFFF###\n",
    b">header1\n>header2\n",
    b">h desc\r\nAC\r\nGT\r\n>x\r\n\r\nAC GT\tU\r\n\r\n",
    b">only_header",
    b"@r\r\nACGT\r\n+\r\n!!!!\r\n\r\n",
]


@pytest.mark.parametrize("content", CASES)
def test_framing_matches_oracle(oracle, content):
    recs = oracle.parse_fastx(content)
    raw = ok.parse_fastx(content, ok.RAW)
    assert raw.ids == [i for i, _ in recs]
    assert [bytes(raw.bases[int(raw.offsets[i]):int(raw.offsets[i + 1])]) for i in range(raw.n_records)] == [s for _, s in recs]
    norm = ok.parse_fastx(content, ok.NORMALIZED)
    got = [bytes(norm.bases[int(norm.offsets[i]):int(norm.offsets[i + 1])]) for i in range(norm.n_records)]
    # whitespace removed exactly where needletail's normalize(false) removes it
    want = [bytes(c for c in s if c not in b" \t\r\n") for _, s in recs]
    assert got == want
    assert [len(oracle.normalize(s)) for _, s in recs] == [len(w) for w in want]


@pytest.mark.parametrize("bad", [b"", b"This is not fasta content\nACGT", b"@r\nACGT\n+\n!!!\n", b"@r\nACGT\n"])
def test_framing_errors(oracle, bad):
    with pytest.raises(ok.FastxError):
        ok.parse_fastx(bad)
    with pytest.raises(oracle.FastxError):
        oracle.parse_fastx(bad)


def test_format_counts(oracle):
    keys = np.array([0, 27, 255], dtype=np.uint64)
    counts = np.array([1, 12345678901, 3], dtype=np.uint64)
    assert ok.format_counts(keys, counts, 4) == oracle.format_counts(keys, counts, 4) == b"AAAA\t1\nACGT\t12345678901\nTTTT\t3\n"


def test_synth_is_deterministic_and_thread_independent():
    H = ok.host_lib()
    g = np.zeros(5000, np.uint8)
    H.okh_synth_genome(3, len(g), ok._ptr(g))
    assert set(np.unique(g)) <= set(b"ACGT")
    a = np.zeros(200 * 150, np.uint8)
    b = np.zeros(200 * 150, np.uint8)
    H.okh_synth_reads(ok._ptr(g), len(g), 7, 0, 200, 150, 5000, 1000, ok._ptr(a), 1)
    H.okh_synth_reads(ok._ptr(g), len(g), 7, 0, 200, 150, 5000, 1000, ok._ptr(b), 4)
    assert np.array_equal(a, b)
    c = np.zeros(100 * 150, np.uint8)
    H.okh_synth_reads(ok._ptr(g), len(g), 7, 100, 100, 150, 5000, 1000, ok._ptr(c), 2)
    assert np.array_equal(a[100 * 150:], c)


def test_parser_fast_path_and_stripping_agree_on_large_inputs(oracle):
    """the parser copies lines without whitespace with one memcpy and strips the others byte by byte: both paths, CRLF
    line ends, wrapped FASTA and a FASTQ with 20,000 records give the oracle's batch"""
    rng = np.random.default_rng(23)
    seq = np.frombuffer(b"ACGTNacgtn", np.uint8)[rng.integers(0, 10, 300_000)]
    from orion_kmer_b200 import synth
    fasta = synth.fasta_text(b"chr1 some description", seq, width=61)
    fasta_crlf = fasta.replace(b"\n", b"\r\n")
    fasta_ws = fasta.replace(b"ACG", b"A C\tG", 2000)
    n = 20_000
    reads = seq[:n * 15].copy()
    fastq = b"".join(b"@r%d extra\n%s\n+\n%s\n" % (i, bytes(reads[i * 15:(i + 1) * 15]), b"I" * 15) for i in range(n))
    for text in (fasta, fasta_crlf, fasta_ws, fastq, fastq.replace(b"\n", b"\r\n")):
        got = ok.parse_fastx(text)
        want = oracle.parse_fastx(text)                 # [(id, raw sequence bytes)]: ids and record framing
        assert got.n_records == (1 if text[:1] == b">" else n)
        assert not np.isin(got.bases, np.frombuffer(b" \t\r\n", np.uint8)).any()
        if text[:1] == b">":
            assert np.array_equal(got.bases, seq) and got.ids == [b"chr1 some description"]
        else:
            assert np.array_equal(got.bases, reads) and got.ids[n - 1] == b"r%d extra" % (n - 1)
            assert np.array_equal(got.offsets, np.arange(n + 1, dtype=np.uint64) * np.uint64(15))
        assert len(want) == got.n_records and [w[0] for w in want[:50]] == got.ids[:50]
        ws = bytes.maketrans(b"", b"")
        assert bytes(got.bases[:int(got.offsets[1])]) == want[0][1].translate(ws, b" \t\r\n")
