"""ctypes binding of the CPU oracle (oracle/orion_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborion_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def build(force=False):
    src = os.path.join(_HERE, "orion_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborion_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_seq_to_u64.argtypes = [C.c_char_p, C.c_uint64, C.c_uint, u64p]
        L.orc_u64_to_seq.argtypes = [C.c_uint64, C.c_uint, C.c_char_p]
        L.orc_reverse_complement_u64.argtypes = [C.c_uint64, C.c_uint, u64p]
        L.orc_canonical_u64.argtypes = [C.c_uint64, C.c_uint, u64p]
        L.orc_normalize.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
        L.orc_normalize.restype = C.c_uint64
        L.orc_fastx_parse.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_fastx_parse.restype = C.c_void_p
        L.orc_fastx_error.argtypes = [C.c_void_p]
        L.orc_fastx_n.argtypes = [C.c_void_p]
        L.orc_fastx_n.restype = C.c_uint64
        L.orc_fastx_record.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p, u64p, u64p]
        L.orc_fastx_free.argtypes = [C.c_void_p]
        L.orc_counter_create.argtypes = [C.c_uint]
        L.orc_counter_create.restype = C.c_void_p
        L.orc_counter_destroy.argtypes = [C.c_void_p]
        L.orc_counter_add_seq.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.orc_counter_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_counter_add_fastx.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.orc_counter_distinct.argtypes = [C.c_void_p]
        L.orc_counter_distinct.restype = C.c_uint64
        L.orc_counter_finish.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(u64p), C.POINTER(u64p), u64p]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_format_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_char_p]
        L.orc_format_counts.restype = C.c_uint64
        L.orc_set_union.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_set_union.restype = C.c_uint64
        L.orc_compare.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_compare.restype = C.c_double
        L.orc_query_hits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p, C.c_void_p,
                                     C.c_uint64, C.c_void_p, C.c_int]
        L.orc_classify_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                       u64p, u64p]
        L.orc_count_batch_mt.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int,
                                         C.c_uint64, C.POINTER(u64p), C.POINTER(u64p), u64p]
        L.orc_count_batch_slice_mt.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                               C.c_int, C.POINTER(u64p), C.POINTER(u64p), u64p]
        L.orc_count_batch_ranged_mt.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int,
                                                C.c_uint64, C.POINTER(u64p), C.POINTER(u64p), u64p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _take(pk, pc, n):
    L = lib()
    n = int(n)
    keys = np.ctypeslib.as_array(pk, shape=(max(n, 1),))[:n].copy()
    counts = np.ctypeslib.as_array(pc, shape=(max(n, 1),))[:n].copy()
    L.orc_free(pk)
    L.orc_free(pc)
    return keys, counts


# ---- src/kmer.rs --------------------------------------------------------------
def seq_to_u64(seq: bytes, k: int):
    out = C.c_uint64()
    return out.value if lib().orc_seq_to_u64(seq, len(seq), k, C.byref(out)) else None


def u64_to_seq(v: int, k: int) -> bytes:
    buf = C.create_string_buffer(33)
    if not lib().orc_u64_to_seq(v, k, buf):
        raise ValueError(f"Invalid k-mer length for decoding: {k}")
    return buf.raw[:k]


def reverse_complement_u64(v: int, k: int) -> int:
    out = C.c_uint64()
    if not lib().orc_reverse_complement_u64(v, k, C.byref(out)):
        raise ValueError(f"Invalid k-mer length for reverse complement: {k}")
    return out.value


def canonical_u64(v: int, k: int) -> int:
    out = C.c_uint64()
    if not lib().orc_canonical_u64(v, k, C.byref(out)):
        raise ValueError(f"Invalid k-mer length: {k}")
    return out.value


def normalize(seq: bytes) -> bytes:
    buf = C.create_string_buffer(len(seq) + 1)
    n = lib().orc_normalize(seq, len(seq), buf)
    return buf.raw[:n]


# ---- needletail framing ---------------------------------------------------------
class FastxError(ValueError):
    pass


def parse_fastx(content: bytes):
    """-> list of (id_bytes, raw_sequence_bytes)"""
    L = lib()
    h = L.orc_fastx_parse(content, len(content))
    try:
        err = L.orc_fastx_error(h)
        if err:
            raise FastxError({1: "empty file", 2: "invalid start byte", 3: "malformed FASTQ"}[err])
        out = []
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        for i in range(L.orc_fastx_n(h)):
            L.orc_fastx_record(h, i, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
            out.append((content[a.value:a.value + b.value], content[c.value:c.value + d.value]))
        return out
    finally:
        L.orc_fastx_free(h)


def batch_from_records(seqs):
    """list of bytes -> (bases u8 array, offsets u64 array) in the C-ABI batch layout"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
    return bases, off


# ---- count.rs ---------------------------------------------------------------------
class InvalidKmerSize(ValueError):
    def __init__(self, k):
        super().__init__(f"Invalid K-mer size: {k}. Must be between 1 and 32.")  # errors.rs:6-7


class Counter:
    """run_count's table (count.rs:48) + finalisation (count.rs:106-119)."""

    def __init__(self, k: int):
        self._h = lib().orc_counter_create(k) if 0 <= k < 2 ** 31 else None
        if not self._h:
            raise InvalidKmerSize(k)
        self.k = k

    def add_seq(self, normalized: bytes):
        lib().orc_counter_add_seq(self._h, normalized, len(normalized))

    def add_batch(self, bases: np.ndarray, offsets: np.ndarray, normalize=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        lib().orc_counter_add_batch(self._h, _ptr(bases), _ptr(offsets), len(offsets) - 1,
                                    1 if normalize else 0)

    def add_fastx(self, content: bytes):
        err = lib().orc_counter_add_fastx(self._h, content, len(content))
        if err:
            raise FastxError(str(err))

    def distinct(self):
        return lib().orc_counter_distinct(self._h)

    def finish(self, min_count=1):
        pk, pc, n = u64p(), u64p(), C.c_uint64()
        lib().orc_counter_finish(self._h, min_count, C.byref(pk), C.byref(pc), C.byref(n))
        return _take(pk, pc, n.value)

    def close(self):
        if self._h:
            lib().orc_counter_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def count_fastx(k, contents, min_count=1):
    """count.rs:40-141 over in-memory file contents -> (sorted keys, counts)"""
    c = Counter(k)
    for content in contents:
        c.add_fastx(content)
    out = c.finish(min_count)
    c.close()
    return out


def count_batch(k, bases, offsets, min_count=1, normalize=True):
    c = Counter(k)
    c.add_batch(bases, offsets, normalize)
    out = c.finish(min_count)
    c.close()
    return out


def count_batch_mt(k, bases, offsets, n_threads, min_count=1):
    """NOT reference behaviour (the reference's count loop is sequential)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    pk, pc, n = u64p(), u64p(), C.c_uint64()
    lib().orc_count_batch_mt(k, _ptr(bases), _ptr(offsets), len(offsets) - 1, n_threads, min_count,
                             C.byref(pk), C.byref(pc), C.byref(n))
    return _take(pk, pc, n.value)


def count_batch_slice_mt(k, bases, offsets, lo, hi, n_threads):
    """Count table restricted to canonical k-mers in [lo, hi] (inclusive); raw bases (no normalize pass)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    pk, pc, n = u64p(), u64p(), C.c_uint64()
    lib().orc_count_batch_slice_mt(k, _ptr(bases), _ptr(offsets), len(offsets) - 1, int(lo), int(hi), n_threads,
                                   C.byref(pk), C.byref(pc), C.byref(n))
    return _take(pk, pc, n.value)


def count_batch_ranged_mt(k, bases, offsets, n_threads, min_count=1):
    """Whole count table of a batch too large for per-thread maps (8 bytes per window of scratch memory)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    pk, pc, n = u64p(), u64p(), C.c_uint64()
    lib().orc_count_batch_ranged_mt(k, _ptr(bases), _ptr(offsets), len(offsets) - 1, n_threads, min_count,
                                    C.byref(pk), C.byref(pc), C.byref(n))
    return _take(pk, pc, n.value)


def format_counts(keys, counts, k) -> bytes:
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    buf = C.create_string_buffer(len(keys) * (k + 22) + 1)
    n = lib().orc_format_counts(_ptr(keys), _ptr(counts), len(keys), k, buf)
    return buf.raw[:n]


# ---- build.rs / db_types.rs -----------------------------------------------------------
def kmer_set_fastx(k, content: bytes) -> np.ndarray:
    """build.rs:23-78,104: distinct canonical k-mers of one genome file (sorted)."""
    return count_fastx(k, [content])[0]


def kmer_set_batch(k, bases, offsets, normalize=True) -> np.ndarray:
    return count_batch(k, bases, offsets, 1, normalize)[0]


def set_union(sets) -> np.ndarray:
    """db_types.rs:43-48 get_all_kmers_unified"""
    sets = [np.ascontiguousarray(s, dtype=np.uint64) for s in sets]
    ptrs = (C.c_void_p * max(len(sets), 1))(*[s.ctypes.data for s in sets])
    ns = np.array([len(s) for s in sets], dtype=np.uint64)
    out = np.zeros(int(ns.sum()) + 1, dtype=np.uint64)
    n = lib().orc_set_union(ptrs, _ptr(ns), len(sets), _ptr(out))
    return out[:n].copy()


def compare(a, b):
    """compare.rs:51-66 -> dict(|A|, |B|, intersection, union, jaccard)"""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    j = lib().orc_compare(_ptr(a), len(a), _ptr(b), len(b), _ptr(out))
    return dict(db1=int(out[0]), db2=int(out[1]), intersection_size=int(out[2]),
                union_size=int(out[3]), jaccard_index=float(j))


def query_hits(kset, k, bases, offsets, n_threads=1) -> np.ndarray:
    """query.rs:79-108 per-read hit counts (raw sequence)."""
    kset = np.ascontiguousarray(kset, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    hits = np.zeros(len(offsets) - 1, dtype=np.uint64)
    lib().orc_query_hits(_ptr(kset), len(kset), k, _ptr(bases), _ptr(offsets), len(offsets) - 1,
                         _ptr(hits), n_threads)
    return hits


def classify_ref(in_keys, in_counts, ref):
    """classify.rs:224-236 -> (matched, sum_depth)"""
    in_keys = np.ascontiguousarray(in_keys, dtype=np.uint64)
    in_counts = np.ascontiguousarray(in_counts, dtype=np.uint64)
    ref = np.ascontiguousarray(ref, dtype=np.uint64)
    m, d = C.c_uint64(), C.c_uint64()
    lib().orc_classify_ref(_ptr(in_keys), _ptr(in_counts), len(in_keys), _ptr(ref), len(ref),
                           C.byref(m), C.byref(d))
    return m.value, d.value
