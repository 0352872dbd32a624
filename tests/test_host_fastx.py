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


def _same_batch(a, b):
    return (np.array_equal(a.bases, b.bases) and np.array_equal(a.offsets, b.offsets) and a.ids == b.ids)


def test_parallel_framing_equals_the_sequential_parser():
    """the text is cut at record starts found from arbitrary byte offsets ('>' lines; '@' lines whose second successor
    begins with '+') and the pieces are parsed by several threads: same batch as the sequential parser, for quality
    lines that begin with '@' or '+', blank lines, CRLF, multi-line FASTA records, fewer records than threads"""
    rng = np.random.default_rng(29)
    alphabet = np.frombuffer(b"ACGTNacgt", np.uint8)
    qual = np.frombuffer(b"@+I#5", np.uint8)

    def fastq(n, nl=b"\n", blank_every=0):
        out = []
        for i in range(n):
            m = int(rng.integers(1, 40))
            seq = bytes(alphabet[rng.integers(0, len(alphabet), m)])
            q = bytes(qual[rng.integers(0, len(qual), m)])             # '@' and '+' as first quality characters too
            out.append(b"@r%d some text" % i + nl + seq + nl + b"+" + (b"r%d" % i if i % 3 == 0 else b"") + nl + q + nl)
            if blank_every and i % blank_every == 0:
                out.append(nl)
        return b"".join(out)

    def fasta(n, nl=b"\n"):
        out = []
        for i in range(n):
            m = int(rng.integers(0, 400))
            seq = bytes(alphabet[rng.integers(0, len(alphabet), m)])
            out.append(b">g%d desc" % i + nl + nl.join(seq[j:j + 37] for j in range(0, m, 37)) + (nl if m else b""))
        return b"".join(out)

    texts = [fastq(2000), fastq(500, b"\r\n"), fastq(700, blank_every=5), fastq(3), fastq(1),
             fasta(1500), fasta(300, b"\r\n"), fasta(1), fasta(2)]
    for text in texts:
        for mode in (ok.NORMALIZED, ok.RAW):
            want = ok.parse_fastx(text, mode, threads=1)
            for t in (2, 3, 7, 16):
                assert _same_batch(ok.parse_fastx(text, mode, threads=t), want), (text[:20], mode, t)
    # errors survive the split: a record cut short in the middle of the text
    bad = fastq(400)
    bad = bad[:len(bad) // 2] + b"@broken\nACGT\n" + bad[len(bad) // 2:]
    for t in (1, 4):
        with pytest.raises(ok.FastxError):
            ok.parse_fastx(bad, threads=t)


def test_framing_fuzz_against_the_oracle(oracle):
    """4,000 random texts glued from the tokens that matter to the framing rules (record markers, '+' lines, LF / CRLF /
    bare CR, blank lines, whitespace inside sequences, markers at odd places): the product's parser and the oracle's
    needletail restatement must accept the same texts and frame the same records, sequentially and in pieces."""
    import random
    rnd = random.Random(20261019)
    toks = [b">", b"@", b"+", b"\n", b"\n", b"\n", b"\r\n", b"ACGT", b"acgtn", b" ", b"\t", b"id1", b"GATTACA", b"!!!!", b"IIIIIII",
            b"N", b"U", b"\n+\n", b"\n>", b"\n@", b"", b"\r"]
    strip = bytes.maketrans(b"", b"")

    def records(b):
        return [(i, bytes(b.bases[int(b.offsets[j]):int(b.offsets[j + 1])])) for j, i in enumerate(b.ids)]

    accepted = 0
    for _ in range(4000):
        text = rnd.choice([b">", b"@", b">", b"@", b""]) + b"".join(rnd.choice(toks) for _ in range(rnd.randint(0, 14)))
        try:
            want = oracle.parse_fastx(text)
        except oracle.FastxError:
            want = None
        for threads in (1, 3):
            try:
                got = records(ok.parse_fastx(text, ok.RAW, threads=threads))
                norm = records(ok.parse_fastx(text, ok.NORMALIZED, threads=threads))
            except ok.FastxError:
                got = norm = None
            assert (got is None) == (want is None), (text, threads)
            if want is not None:
                assert got == want, (text, threads)
                assert norm == [(i, s.translate(strip, b" \t\r\n")) for i, s in want], (text, threads)
        accepted += want is not None
    assert 400 < accepted < 3600          # both outcomes are exercised
