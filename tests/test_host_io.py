"""Host I/O rows of SURVEY.md 8(f): codecs by extension / magic (utils.rs:115-199) and the bincode
KmerDbV2 .db format (db_types.rs:8-14, build.rs:141, utils.rs:37-55).  CPU only."""
import gzip
import lzma
import os
import struct

import numpy as np
import pytest

import orion_kmer_b200 as ok

FIX = os.path.join(os.path.dirname(__file__), "golden", "fixtures")
INPUT1 = b">seq1\nACGTACGTACGT\n>seq2\nTGCATGCATGCANNNACGT\n>seq3\nGATTACA\nNNNNN\nGATTACA\n"


def test_reference_fixtures_decode_identically_in_all_three_codecs():
    for stem in ("test_input1.fasta", "test_input2.fastq"):
        texts = {ext: ok.read_file(os.path.join(FIX, f"{stem}.{ext}")) for ext in ("gz", "xz", "zst")}
        assert texts["gz"] == texts["xz"] == texts["zst"]
        assert texts["gz"][:1] in (b">", b"@")
    assert ok.read_file(os.path.join(FIX, "test_input1.fasta.gz")) == gzip.open(os.path.join(FIX, "test_input1.fasta.gz")).read()
    assert ok.read_file(os.path.join(FIX, "test_input2.fastq.xz")) == lzma.open(os.path.join(FIX, "test_input2.fastq.xz")).read()


@pytest.mark.parametrize("ext", ["txt", "gz", "xz", "zst", "zstd", "GZ"])
def test_writer_reader_round_trip_by_extension(tmp_path, ext):
    rng = np.random.default_rng(5)
    data = bytes(rng.integers(65, 70, 300_000, dtype=np.uint8)) + b"\n" + bytes(range(256))
    p = str(tmp_path / f"blob.{ext}")
    ok.write_file(p, data)
    raw = open(p, "rb").read()
    if ext == "txt":
        assert raw == data
    else:
        assert raw != data and len(raw) < len(data)
    if ext.lower() == "gz":
        assert gzip.decompress(raw) == data          # a real gzip member, not our own container
    if ext == "xz":
        assert lzma.decompress(raw) == data
    assert ok.read_file(p) == data
    ok.write_file(p, b"")                            # empty payloads survive every codec
    assert ok.read_file(p) == b""


def test_multi_member_gzip_and_magic_sniffing(tmp_path):
    p = str(tmp_path / "two_members.fa.gz")
    open(p, "wb").write(gzip.compress(b">a\nACGT\n") + gzip.compress(b">b\nTTTT\n"))
    assert ok.read_file(p) == b">a\nACGT\n>b\nTTTT\n"            # flate2 MultiGzDecoder (utils.rs:131)
    # build / classify hand the raw file to needletail, which sniffs gzip and xz magic -- the name is irrelevant
    q = str(tmp_path / "looks_plain.fasta")
    open(q, "wb").write(gzip.compress(INPUT1))
    assert ok.read_file(q, by_magic=True) == INPUT1
    assert ok.read_file(q) != INPUT1                               # by extension it is "plain"
    r = str(tmp_path / "x.bin")
    open(r, "wb").write(lzma.compress(INPUT1))
    assert ok.read_file(r, by_magic=True) == INPUT1
    z = os.path.join(FIX, "test_input1.fasta.zst")                 # needletail 0.5.1 has no zstd: bytes come back as they are
    assert ok.read_file(z, by_magic=True) == open(z, "rb").read()


def test_io_errors_are_reported(tmp_path):
    with pytest.raises(ok.IoError, match="Failed to open input file"):
        ok.read_file(str(tmp_path / "missing.fa"))
    bad = str(tmp_path / "bad.gz")
    open(bad, "wb").write(b"this is not gzip")
    with pytest.raises(ok.IoError, match="gzip"):
        ok.read_file(bad)
    with pytest.raises(ok.IoError, match="Failed to create output file"):
        ok.write_file(str(tmp_path / "no_such_dir" / "out.txt"), b"x")


def test_db_bytes_follow_bincode_1_3_fixint_little_endian(tmp_path):
    """k:u8 | n_refs:u64 | { name_len:u64 | utf-8 | n_kmers:u64 | n_kmers x u64 } ..."""
    p = str(tmp_path / "tiny.db")
    ok.write_kmer_db(p, 5, {"a.fa": np.array([1, 2, 515], np.uint64), "böb": np.array([], np.uint64)})
    want = struct.pack("<BQ", 5, 2)
    want += struct.pack("<Q", 4) + b"a.fa" + struct.pack("<Q3Q", 3, 1, 2, 515)
    want += struct.pack("<Q", len("böb".encode())) + "böb".encode() + struct.pack("<Q", 0)
    assert open(p, "rb").read() == want
    k, refs = ok.read_kmer_db(p)
    assert k == 5 and list(refs) == ["a.fa", "böb"]
    assert refs["a.fa"].tolist() == [1, 2, 515] and len(refs["böb"]) == 0


@pytest.mark.parametrize("ext", ["db", "db.gz", "db.xz", "db.zst"])
def test_db_round_trip_compressed_and_in_any_key_order(tmp_path, ext):
    rng = np.random.default_rng(11)
    refs = {f"genome_{i}.fasta.gz": rng.permutation(rng.integers(0, 1 << 62, 5000, dtype=np.uint64)) for i in range(3)}
    p = str(tmp_path / f"refs.{ext}")
    ok.write_kmer_db(p, 31, refs)
    k, got = ok.read_kmer_db(p)
    assert k == 31 and list(got) == list(refs)
    for name in refs:                      # hash order in the reference's files: order is preserved as stored
        assert np.array_equal(got[name], refs[name])


def test_db_same_name_overwrites_and_truncation_is_detected(tmp_path):
    p = str(tmp_path / "dup.db")
    blob = struct.pack("<BQ", 4, 2)
    blob += struct.pack("<Q", 1) + b"x" + struct.pack("<Q2Q", 2, 7, 9)
    blob += struct.pack("<Q", 1) + b"x" + struct.pack("<Q1Q", 1, 11)      # HashMap: the later entry wins (db_types.rs:38-40)
    open(p, "wb").write(blob)
    k, refs = ok.read_kmer_db(p)
    assert k == 4 and list(refs) == ["x"] and refs["x"].tolist() == [11]
    open(p, "wb").write(blob[:-3])
    with pytest.raises(ok.IoError, match="Failed to deserialize KmerDbV2"):
        ok.read_kmer_db(p)
    open(p, "wb").write(struct.pack("<BQ", 4, 1) + struct.pack("<Q", 1 << 60))   # absurd length must not allocate
    with pytest.raises(ok.IoError, match="Failed to deserialize KmerDbV2"):
        ok.read_kmer_db(p)


def test_bzip2_and_the_sniff_after_the_extension_codec(tmp_path):
    """needletail 0.5.1 (Cargo.lock:580-591: bzip2, flate2, xz2) sniffs the magic bytes of whatever reader it is given.
    build / classify give it the raw file; count / query give it the output of the extension codec (count.rs:59-63), so
    a .bz2 input, or a gzip file under another name, is decoded for them as well.  .db files are never sniffed."""
    import bz2
    p = str(tmp_path / "genome.fa.bz2")
    open(p, "wb").write(bz2.compress(INPUT1))
    assert ok.read_file(p, by_magic=True) == INPUT1                 # build / classify
    assert ok.read_file(p, fastx=True) == INPUT1                    # count / query: no codec for ".bz2", then the sniff
    assert ok.read_file(p) == open(p, "rb").read()                  # a .db path: extension only
    multi = str(tmp_path / "two_streams.bz2")                       # pbzip2 writes one stream per block
    open(multi, "wb").write(bz2.compress(b">a\nACGT\n") + bz2.compress(b">b\nTTTT\n") + b"\0\0\0")
    assert ok.read_file(multi, by_magic=True) == b">a\nACGT\n>b\nTTTT\n"
    big = bytes(np.random.default_rng(3).integers(65, 70, 3_000_000, dtype=np.uint8))      # several 900 kB blocks, > one output buffer
    open(p, "wb").write(bz2.compress(big))
    assert ok.read_file(p, by_magic=True) == big
    q = str(tmp_path / "reads.fastq")                               # gzip content, no .gz in the name
    open(q, "wb").write(gzip.compress(INPUT1))
    assert ok.read_file(q, fastx=True) == INPUT1 and ok.read_file(q) != INPUT1
    twice = str(tmp_path / "twice.fa.gz")                           # the extension codec, then the sniff finds xz
    open(twice, "wb").write(gzip.compress(lzma.compress(INPUT1)))
    assert ok.read_file(twice, fastx=True) == INPUT1 and ok.read_file(twice) == lzma.compress(INPUT1)
    for name in ("test_input1.fasta.gz", "test_input1.fasta.xz", "test_input1.fasta.zst"):       # the reference's fixtures: unchanged
        assert ok.read_file(os.path.join(FIX, name), fastx=True) == INPUT1
    bad = str(tmp_path / "bad.fa")
    open(bad, "wb").write(b"BZh9 this is not bzip2 at all")
    with pytest.raises(ok.IoError, match="bzip2"):
        ok.read_file(bad, by_magic=True)
    cut = str(tmp_path / "cut.bz2")
    open(cut, "wb").write(bz2.compress(big)[:100_000])
    with pytest.raises(ok.IoError, match="bzip2"):
        ok.read_file(cut, fastx=True)
