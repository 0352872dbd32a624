"""orion_kmer_b200 -- B200-native k-mer hot path of orion-kmer behind a C ABI.

This module is the thin host mirror used by tests and bench.py: ctypes over
include/orion_gpu.h (liborion_gpu.so, CUDA sm_100a) and over the host helpers
(liborion_host.so: FASTA/FASTQ framing, TSV formatting, synthetic workloads), shaped like the
reference's drivers (count.rs, build.rs, compare.rs, query.rs, classify.rs, db_types.rs).

There is no CPU fallback: if liborion_gpu.so is missing or no CUDA device is present, compute
calls raise.
"""
import ctypes as C
import os

import numpy as np

from . import _build

_HERE = os.path.dirname(os.path.abspath(__file__))
NORMALIZED, RAW = 0, 1

OK_SUCCESS = 0
OK_ERR_INVALID_KMER_SIZE = 1
OK_ERR_KMER_SIZE_MISMATCH = 2
OK_ERR_INVALID_ARGUMENT = 3
OK_ERR_OUT_OF_MEMORY = 4
OK_ERR_CUDA = 5
OK_ERR_NO_DEVICE = 6
OK_ERR_INTERNAL = 7

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class OrionError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class InvalidKmerSize(OrionError, ValueError):   # errors.rs:6-7
    pass


class KmerSizeMismatch(OrionError, ValueError):  # errors.rs:24-25
    pass


class FastxError(ValueError):
    pass


class CounterStats(C.Structure):
    _fields_ = [("n_slots", C.c_uint64), ("n_distinct", C.c_uint64), ("n_windows", C.c_uint64),
                ("n_bases", C.c_uint64), ("max_displacement", C.c_uint64), ("n_spilled", C.c_uint64),
                ("n_grows", C.c_uint64), ("ms_insert", C.c_float), ("ms_readout", C.c_float),
                ("ms_fill", C.c_float), ("ms_route", C.c_float), ("ms_sample", C.c_float),
                ("ms_scatter1", C.c_float), ("ms_scatter2", C.c_float), ("ms_count", C.c_float),
                ("ms_compact", C.c_float), ("partitioned", C.c_int), ("ms_push", C.c_float), ("n_deferred", C.c_uint64),
                ("ms_merge", C.c_float), ("n_merges", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/orion_gpu.h declares
ABI = {
    "ok_init": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "ok_shutdown": (C.c_int, []),
    "ok_last_error": (C.c_char_p, []),
    "ok_version": (C.c_char_p, []),
    "ok_launch_count": (C.c_uint64, []),
    "ok_synchronize": (C.c_int, []),
    "ok_host_alloc": (C.c_int, [C.POINTER(vp), C.c_uint64]),
    "ok_host_free": (C.c_int, [vp]),
    "ok_free": (C.c_int, [vp]),
    "ok_seq_to_u64": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint8, u64p]),
    "ok_u64_to_seq": (C.c_int, [C.c_uint64, C.c_uint8, C.c_char_p]),
    "ok_reverse_complement_u64": (C.c_int, [C.c_uint64, C.c_uint8, u64p]),
    "ok_canonical_u64": (C.c_int, [C.c_uint64, C.c_uint8, u64p]),
    "ok_counter_create": (C.c_int, [C.c_uint8, C.c_int, C.c_uint64, C.POINTER(vp)]),
    "ok_counter_add_batch": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "ok_counter_add_batch_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64]),
    "ok_counter_set_shard": (C.c_int, [vp, C.c_int, C.c_int]),
    "ok_counter_add_kmers_device": (C.c_int, [vp, vp, C.c_uint64]),
    "ok_counter_route_batch_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_int, vp, vp]),
    "ok_peer_buffer_create": (C.c_int, [C.c_uint64, C.POINTER(vp), vp]),
    "ok_peer_buffer_open": (C.c_int, [vp, C.POINTER(vp)]),
    "ok_peer_buffer_close": (C.c_int, [vp]),
    "ok_peer_buffer_destroy": (C.c_int, [vp]),
    "ok_counter_route_count_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_int, vp]),
    "ok_counter_route_scatter_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_int, vp, vp]),
    "ok_shard_geometry": (C.c_int, [vp, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p]),
    "ok_shard_set_buffers": (C.c_int, [vp, vp, C.c_uint64]),
    "ok_shard_set_margin": (C.c_int, [vp, C.c_double]),
    "ok_shard_sample_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "ok_shard_scatter_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp, vp]),
    "ok_shard_count_device": (C.c_int, [vp, vp]),
    "ok_xchg_geometry": (C.c_int, [vp, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), u64p]),
    "ok_xchg_sample_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "ok_xchg_scatter_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "ok_xchg_count_device": (C.c_int, [vp]),
    "ok_xchg_scatter_begin": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "ok_xchg_chunk_sent": (C.c_int, [vp, C.c_uint32]),
    "ok_xchg_chunk_recv": (C.c_int, [vp, C.c_uint32]),
    "ok_xchg_scatter_end": (C.c_int, [vp]),
    "ok_counter_finish": (C.c_int, [vp, C.c_uint64, C.POINTER(u64p), C.POINTER(u64p), u64p]),
    "ok_counter_finish_device": (C.c_int, [vp, C.c_uint64, C.POINTER(vp), C.POINTER(vp), u64p]),
    "ok_counter_set_path": (C.c_int, [vp, C.c_int]),
    "ok_counter_set_capacity_hint": (C.c_int, [vp, C.c_uint64]),
    "ok_counter_abort_batch": (C.c_int, [vp]),
    "ok_counter_commit_batch": (C.c_int, [vp]),
    "ok_counter_clear": (C.c_int, [vp]),
    "ok_counter_destroy": (C.c_int, [vp]),
    "ok_counter_get_stats": (C.c_int, [vp, C.POINTER(CounterStats)]),
    "ok_set_create": (C.c_int, [C.c_uint8, C.c_int, C.c_uint64, C.POINTER(vp)]),
    "ok_set_add_batch": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "ok_set_add_batch_device": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint64]),
    "ok_sets_build_many_device": (C.c_int, [C.c_uint8, C.c_int, C.c_uint64, vp, vp, vp, vp, C.POINTER(vp)]),
    "ok_sets_build_many": (C.c_int, [C.c_uint8, C.c_int, C.c_uint64, vp, vp, vp, C.POINTER(vp)]),
    "ok_set_from_sorted": (C.c_int, [C.c_uint8, vp, C.c_uint64, C.POINTER(vp)]),
    "ok_set_from_sorted_device": (C.c_int, [C.c_uint8, vp, C.c_uint64, C.POINTER(vp)]),
    "ok_set_keys_device": (C.c_int, [vp, C.POINTER(vp), u64p]),
    "ok_set_shard_bounds": (C.c_int, [vp, C.c_int, vp]),
    "ok_set_copy_keys_device": (C.c_int, [vp, C.c_uint64, C.c_uint64, vp]),
    "ok_probe_reads_device": (C.c_int, [vp, C.c_int, vp, C.c_uint64, vp, C.c_uint64, vp]),
    "ok_set_size": (C.c_int, [vp, u64p]),
    "ok_set_k": (C.c_int, [vp, u8p]),
    "ok_set_export": (C.c_int, [vp, C.POINTER(u64p), u64p]),
    "ok_set_union": (C.c_int, [C.POINTER(vp), C.c_uint64, C.POINTER(vp)]),
    "ok_set_destroy": (C.c_int, [vp]),
    "ok_set_intersection_size": (C.c_int, [vp, vp, u64p]),
    "ok_sets_all_vs_all": (C.c_int, [C.POINTER(vp), C.c_uint64, vp, vp]),
    "ok_sets_all_vs_all_part": (C.c_int, [C.POINTER(vp), C.c_uint64, C.c_uint64, C.c_uint64, vp, vp]),
    "ok_probe_reads": (C.c_int, [vp, C.c_int, vp, vp, C.c_uint64, vp]),
    "ok_probe_counts": (C.c_int, [vp, vp, vp, C.c_uint64, u64p, u64p]),
    "ok_probe_counts_many": (C.c_int, [C.POINTER(vp), C.c_uint64, vp, vp, C.c_uint64, vp, vp]),
    "ok_pack_2bit_device": (C.c_int, [vp, C.c_uint64, C.c_int, vp, vp]),
}
# internal test hooks (not in include/)
_HOOKS = {
    "okx_emulate_extract": (C.c_int, [vp, C.c_uint64, vp, C.c_uint64, C.c_uint, C.c_int, vp, C.c_uint64, u64p]),
    "okx_emulate_table": (C.c_int, [vp, C.c_uint64, C.c_uint, C.c_int, C.c_uint64, C.c_uint, C.c_uint64, vp, vp, u64p, u64p]),
    "okx_device_extract": (C.c_int, [vp, vp, C.c_uint64, C.c_uint, C.c_int, vp, C.c_uint64, u64p]),
    "okx_owner_of": (C.c_int, [vp, C.c_uint64, C.c_uint, C.c_int, vp]),
    "okx_plan_bits": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint, vp]),
    "okx_strided_order": (C.c_int, [C.c_uint64, vp, C.c_uint64, u64p]),
    "okx_ava_geometry": (C.c_int, [C.c_uint, vp, vp, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp]),
    "okx_ava_block": (C.c_int, [C.c_uint, C.c_uint, vp]),
}

_gpu = None
_host = None


def gpu_library_path():
    return _build.SO


def lib():
    """liborion_gpu.so; raises when it is not built (no CPU fallback exists)."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(_build.SO):
            raise OrionError(OK_ERR_NO_DEVICE,
                             f"{_build.SO} is missing: build it with __graft_entry__.build() "
                             "(there is no CPU fallback)")
        L = C.CDLL(_build.SO)
        for name, (res, args) in {**ABI, **_HOOKS}.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _gpu = L
    return _gpu


def host_lib():
    global _host
    if _host is None:
        so = os.path.join(_HERE, "liborion_host.so")
        if not os.path.exists(so):
            build_host()
        L = C.CDLL(so)
        L.okh_fastx_parse.restype = vp
        L.okh_fastx_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_int)]
        for n in ("okh_batch_n_records", "okh_batch_n_bases"):
            getattr(L, n).restype = C.c_uint64
            getattr(L, n).argtypes = [vp]
        for n in ("okh_batch_bases", "okh_batch_offsets", "okh_batch_ids", "okh_batch_id_offsets"):
            getattr(L, n).restype = vp
            getattr(L, n).argtypes = [vp]
        L.okh_batch_free.argtypes = [vp]
        L.okh_fastx_parse_mt.restype = vp
        L.okh_fastx_parse_mt.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.okh_format_counts.restype = C.c_uint64
        L.okh_format_counts.argtypes = [vp, vp, C.c_uint64, C.c_uint, vp]
        L.okh_synth_genome.argtypes = [C.c_uint64, C.c_uint64, vp]
        L.okh_json_f64.restype = C.c_int; L.okh_json_f64.argtypes = [C.c_double, C.c_char_p]
        L.okh_synth_reads.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                      C.c_uint32, C.c_uint32, vp, C.c_int]
        L.okh_synth_mutate.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, vp]
        # orion_io.cpp: codecs + .db
        L.okh_io_last_error.restype = C.c_char_p
        L.okh_read_file.restype = vp; L.okh_read_file.argtypes = [C.c_char_p, C.c_int]
        L.okh_file_data.restype = vp; L.okh_file_data.argtypes = [vp]
        L.okh_file_size.restype = C.c_uint64; L.okh_file_size.argtypes = [vp]
        L.okh_file_free.argtypes = [vp]
        L.okh_write_file.restype = C.c_int; L.okh_write_file.argtypes = [C.c_char_p, vp, C.c_uint64, C.c_int]
        L.okh_db_new.restype = vp; L.okh_db_new.argtypes = [C.c_uint8]
        L.okh_db_add_reference.argtypes = [vp, C.c_char_p, vp, C.c_uint64]
        L.okh_db_k.restype = C.c_uint8; L.okh_db_k.argtypes = [vp]
        L.okh_db_n_references.restype = C.c_uint64; L.okh_db_n_references.argtypes = [vp]
        L.okh_db_name.restype = C.c_char_p; L.okh_db_name.argtypes = [vp, C.c_uint64]
        L.okh_db_n_kmers.restype = C.c_uint64; L.okh_db_n_kmers.argtypes = [vp, C.c_uint64]
        L.okh_db_kmers.restype = vp; L.okh_db_kmers.argtypes = [vp, C.c_uint64]
        L.okh_db_free.argtypes = [vp]
        L.okh_db_write.restype = C.c_int; L.okh_db_write.argtypes = [vp, C.c_char_p]
        L.okh_db_read.restype = vp; L.okh_db_read.argtypes = [C.c_char_p]
        _host = L
    return _host


def build_host(force=False):
    """g++: liborion_host.so (framing, TSV, codecs, .db format, synthetic data) and the CLI binary
    orion-kmer-b200 (the reference's command line over liborion_gpu.so)."""
    import subprocess
    hd = os.path.join(_HERE, "csrc", "host")
    srcs = [os.path.join(hd, "orion_host.cpp"), os.path.join(hd, "orion_io.cpp")]
    so = os.path.join(_HERE, "liborion_host.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread",
                               "-shared", "-o", so] + srcs + ["-lz", "-ldl"])
    cli_src = os.path.join(hd, "orion_cli.cpp")
    exe = cli_path()
    gpu_so = os.path.join(_HERE, "liborion_gpu.so")
    if os.path.exists(gpu_so) and (force or not os.path.exists(exe) or
                                   os.path.getmtime(exe) < max(os.path.getmtime(cli_src), os.path.getmtime(so))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(os.path.dirname(_HERE), "include"),
                               "-o", exe, cli_src, "-L", _HERE, "-lorion_gpu", "-lorion_host",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + _HERE])
    return so


def cli_path():
    return os.path.join(_HERE, "orion-kmer-b200")


class IoError(OSError):
    pass


def read_file(path, by_magic=False, fastx=False):
    """utils.rs:125-152 (codec by extension: the .db files); by_magic: the raw read + needletail's magic sniff
    (gzip / bzip2 / xz) of build and classify; fastx: the extension codec and THEN the sniff, the way count and query
    read their FASTA/FASTQ inputs (count.rs:59-63)"""
    H = host_lib()
    h = H.okh_read_file(os.fsencode(path), 1 if by_magic else (2 if fastx else 0))
    if not h:
        raise IoError(H.okh_io_last_error().decode())
    try:
        n = H.okh_file_size(h)
        return C.string_at(H.okh_file_data(h), n) if n else b""
    finally:
        H.okh_file_free(h)


def json_f64(v):
    """serde_json's text of an f64 (the ratios in the compare / classify reports): ryu's shortest round-trip digits"""
    b = C.create_string_buffer(48)
    n = host_lib().okh_json_f64(float(v), b)
    return b.raw[:n].decode()


def write_file(path, data, by_extension=True):
    """utils.rs:167-199 get_output_writer"""
    H = host_lib()
    buf = (C.c_uint8 * max(len(data), 1)).from_buffer_copy(data if data else b"\0")
    if H.okh_write_file(os.fsencode(path), buf, len(data), 1 if by_extension else 0):
        raise IoError(H.okh_io_last_error().decode())


def write_kmer_db(path, k, references):
    """db_types.rs:8-14 KmerDbV2 as bincode 1.3.3 (build.rs:141).  references: {name: uint64 array}"""
    H = host_lib()
    h = H.okh_db_new(k)
    try:
        for name, kmers in references.items():
            a = np.ascontiguousarray(kmers, dtype=np.uint64)
            H.okh_db_add_reference(h, name.encode(), _ptr(a), len(a))
        if H.okh_db_write(h, os.fsencode(path)):
            raise IoError(H.okh_io_last_error().decode())
    finally:
        H.okh_db_free(h)


def read_kmer_db(path):
    """utils.rs:37-55 load_kmer_db_v2 -> (k, {name: uint64 array in file order})"""
    H = host_lib()
    h = H.okh_db_read(os.fsencode(path))
    if not h:
        raise IoError(H.okh_io_last_error().decode())
    try:
        refs = {}
        for i in range(H.okh_db_n_references(h)):
            refs[H.okh_db_name(h, i).decode()] = _copy_out(H.okh_db_kmers(h, i), H.okh_db_n_kmers(h, i), C.c_uint64, np.uint64)
        return H.okh_db_k(h), refs
    finally:
        H.okh_db_free(h)


def _check(rc):
    if rc == OK_SUCCESS:
        return
    msg = lib().ok_last_error().decode()
    if rc == OK_ERR_INVALID_KMER_SIZE:
        raise InvalidKmerSize(rc, msg)
    if rc == OK_ERR_KMER_SIZE_MISMATCH:
        raise KmerSizeMismatch(rc, msg)
    raise OrionError(rc, msg)


def _ptr(a):
    return a.ctypes.data_as(vp)


def _copy_out(ptr, n, ctype, dtype):
    """numpy copy of n elements at a C pointer (which may be NULL when n == 0)"""
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(n,)).copy()


def init(device=0):
    dev = (C.c_int * 1)(device)
    _check(lib().ok_init(dev, 1))


def launch_count():
    return int(lib().ok_launch_count())


# ---- src/kmer.rs -----------------------------------------------------------------------------
def seq_to_u64(seq: bytes, k: int):
    out = C.c_uint64()
    if not 0 <= k <= 255:
        return None
    return out.value if lib().ok_seq_to_u64(seq, len(seq), k, C.byref(out)) else None


def u64_to_seq(v: int, k: int) -> bytes:
    if not 0 <= k <= 255:
        raise InvalidKmerSize(OK_ERR_INVALID_KMER_SIZE, f"Invalid k-mer length for decoding: {k}")
    buf = C.create_string_buffer(33)
    _check(lib().ok_u64_to_seq(v, k, buf))
    return buf.raw[:k]


def reverse_complement_u64(v: int, k: int) -> int:
    out = C.c_uint64()
    _check(lib().ok_reverse_complement_u64(v, k if 0 <= k <= 255 else 0, C.byref(out)))
    return out.value


def canonical_u64(v: int, k: int) -> int:
    out = C.c_uint64()
    _check(lib().ok_canonical_u64(v, k if 0 <= k <= 255 else 0, C.byref(out)))
    return out.value


# ---- needletail-shaped framing (host) -------------------------------------------------------------
class Batch:
    """C-ABI batch: concatenated bases + n+1 offsets (+ record ids)."""

    def __init__(self, bases, offsets, ids=None, id_blob=None, id_offsets=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._ids, self._id_blob, self._id_offsets = ids, id_blob, id_offsets

    @property
    def ids(self):
        """record ids as a list of bytes; sliced out of the id blob on first use (count / build never look at them,
        and a list of 10 M bytes objects costs more than parsing the file)"""
        if self._ids is None and self._id_blob is not None:
            o = self._id_offsets
            self._ids = [self._id_blob[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]
        return self._ids

    @property
    def n_records(self):
        return len(self.offsets) - 1


def parse_fastx(content: bytes, norm_mode=NORMALIZED, threads=-1) -> Batch:
    """parse_fastx_reader (+ whitespace removal of normalize(false) when NORMALIZED).  threads: -1 = as many as the host
    offers for large texts (the text is cut at record starts), 1 = the sequential parser, n = exactly n pieces"""
    H = host_lib()
    st = C.c_int()
    h = H.okh_fastx_parse_mt(content, len(content), 1 if norm_mode == NORMALIZED else 0, threads, C.byref(st))
    try:
        if st.value:
            raise FastxError({1: "empty file", 2: "invalid start byte", 3: "malformed FASTQ"}[st.value])
        n, nb = H.okh_batch_n_records(h), H.okh_batch_n_bases(h)
        bases = _copy_out(H.okh_batch_bases(h), nb, C.c_uint8, np.uint8)
        off = _copy_out(H.okh_batch_offsets(h), n + 1, C.c_uint64, np.uint64)
        ido = _copy_out(H.okh_batch_id_offsets(h), n + 1, C.c_uint64, np.uint64)
        blob = bytes(_copy_out(H.okh_batch_ids(h), int(ido[-1]), C.c_uint8, np.uint8))
        return Batch(bases, off, id_blob=blob, id_offsets=ido)
    finally:
        H.okh_batch_free(h)


def format_counts(kmers, counts, k) -> bytes:
    """count.rs:127-135"""
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    buf = np.empty(len(kmers) * (k + 22) + 1, dtype=np.uint8)        # k bases + tab + <= 20 digits + newline per line
    n = host_lib().okh_format_counts(_ptr(kmers), _ptr(counts), len(kmers), k, _ptr(buf))
    return buf[:n].tobytes()


# ---- counter (count.rs) ------------------------------------------------------------------------------
def _take(pk, pc, n):
    n = int(n)
    L = lib()
    keys = _copy_out(pk, n, C.c_uint64, np.uint64)
    counts = _copy_out(pc, n, C.c_uint64, np.uint64) if pc is not None else None
    _check(L.ok_free(C.cast(pk, vp)))
    if pc is not None:
        _check(L.ok_free(C.cast(pc, vp)))
    return keys, counts


class KmerCounter:
    """DashMap<u64, AtomicUsize> of run_count (count.rs:48) on the device."""

    def __init__(self, k, norm_mode=NORMALIZED, capacity_hint=0):
        h = vp()
        if not 0 <= k <= 255:
            raise InvalidKmerSize(OK_ERR_INVALID_KMER_SIZE, f"Invalid K-mer size: {k}. Must be between 1 and 32.")
        _check(lib().ok_counter_create(k, norm_mode, capacity_hint, C.byref(h)))
        self._h, self.k, self.norm_mode = h, k, norm_mode

    def add_batch(self, bases, offsets):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().ok_counter_add_batch(self._h, _ptr(bases), _ptr(offsets), len(offsets) - 1))

    def add_batch_ptr(self, bases_ptr, offsets_ptr, n_records):
        """raw host pointers (e.g. pinned torch tensors)"""
        _check(lib().ok_counter_add_batch(self._h, bases_ptr, offsets_ptr, n_records))

    def add_batch_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records):
        _check(lib().ok_counter_add_batch_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records))

    def set_shard(self, rank, n_ranks):
        _check(lib().ok_counter_set_shard(self._h, rank, n_ranks))

    def route_batch_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, n_ranks, d_out_ptr):
        """-> per-rank k-mer counts; d_out holds the k-mers rank after rank"""
        counts = np.zeros(n_ranks, dtype=np.uint64)
        _check(lib().ok_counter_route_batch_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records,
                                                   n_ranks, d_out_ptr, _ptr(counts)))
        return counts

    def route_count_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, n_ranks):
        counts = np.zeros(n_ranks, dtype=np.uint64)
        _check(lib().ok_counter_route_count_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, n_ranks,
                                                   _ptr(counts)))
        return counts

    def route_scatter_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, dst_ptrs, counts):
        n_ranks = len(dst_ptrs)
        arr = (vp * n_ranks)(*[int(p) for p in dst_ptrs])
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        _check(lib().ok_counter_route_scatter_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, n_ranks,
                                                     arr, _ptr(counts)))

    # ---- fused multi-GPU exchange (sharded scatter); collectives stay with the caller ----
    def shard_geometry(self, n_bases_max):
        sb, lb, cap = C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().ok_shard_geometry(self._h, n_bases_max, C.byref(sb), C.byref(lb), C.byref(cap)))
        return sb.value, lb.value, cap.value

    def shard_set_margin(self, margin):
        _check(lib().ok_shard_set_margin(self._h, margin))

    def shard_set_buffers(self, peer_ptrs, cap_keys):
        arr = (vp * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        _check(lib().ok_shard_set_buffers(self._h, arr, cap_keys))

    def shard_sample_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_fine_ptr, d_hist_l1_ptr):
        _check(lib().ok_shard_sample_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_fine_ptr, d_hist_l1_ptr))

    def shard_scatter_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr, d_hist_l1_all_ptr, d_cursors_ptr):
        _check(lib().ok_shard_scatter_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr,
                                             d_hist_l1_all_ptr, d_cursors_ptr))

    def shard_count_device(self, d_cursors_all_ptr):
        _check(lib().ok_shard_count_device(self._h, d_cursors_all_ptr))

    # ---- chunked exchange over the copy engines (the default multi-GPU form); collectives stay with the caller ----
    def xchg_geometry(self, n_bases_max):
        sb, lb, nc, cap = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(lib().ok_xchg_geometry(self._h, n_bases_max, C.byref(sb), C.byref(lb), C.byref(nc), C.byref(cap)))
        return sb.value, lb.value, nc.value, cap.value

    def xchg_sample_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_fine_ptr, d_hist_l1c_ptr):
        _check(lib().ok_xchg_sample_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_fine_ptr, d_hist_l1c_ptr))

    def xchg_scatter_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr, h_l1c_all):
        h = np.ascontiguousarray(h_l1c_all, dtype=np.uint32)
        _check(lib().ok_xchg_scatter_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr, _ptr(h)))

    def xchg_scatter_begin(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr, h_l1c_all):
        h = np.ascontiguousarray(h_l1c_all, dtype=np.uint32)
        _check(lib().ok_xchg_scatter_begin(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hist_mine_ptr, _ptr(h)))

    def xchg_chunk_sent(self, chunk):
        _check(lib().ok_xchg_chunk_sent(self._h, chunk))

    def xchg_chunk_recv(self, chunk):
        _check(lib().ok_xchg_chunk_recv(self._h, chunk))

    def xchg_scatter_end(self):
        _check(lib().ok_xchg_scatter_end(self._h))

    def xchg_count_device(self):
        _check(lib().ok_xchg_count_device(self._h))

    def add_kmers_device(self, d_kmers_ptr, n):
        _check(lib().ok_counter_add_kmers_device(self._h, d_kmers_ptr, n))

    def add_fastx(self, content: bytes):
        b = parse_fastx(content, self.norm_mode)
        self.add_batch(b.bases, b.offsets)

    def finish(self, min_count=1):
        pk, pc, n = u64p(), u64p(), C.c_uint64()
        _check(lib().ok_counter_finish(self._h, min_count, C.byref(pk), C.byref(pc), C.byref(n)))
        return _take(pk, pc, n.value)

    def finish_raw(self, min_count=1):
        """-> (kmers_ptr, counts_ptr, n): page-locked host arrays, release with free_result()"""
        pk, pc, n = u64p(), u64p(), C.c_uint64()
        _check(lib().ok_counter_finish(self._h, min_count, C.byref(pk), C.byref(pc), C.byref(n)))
        return pk, pc, n.value

    @staticmethod
    def free_result(pk, pc):
        _check(lib().ok_free(C.cast(pk, vp)))
        _check(lib().ok_free(C.cast(pc, vp)))

    def finish_device(self, min_count=1):
        dk, dc, n = vp(), vp(), C.c_uint64()
        _check(lib().ok_counter_finish_device(self._h, min_count, C.byref(dk), C.byref(dc), C.byref(n)))
        return dk.value, dc.value, n.value

    def set_capacity_hint(self, capacity_hint):
        """expected distinct k-mers (0 = none): a speed knob, see include/orion_gpu.h"""
        _check(lib().ok_counter_set_capacity_hint(self._h, capacity_hint))

    def set_path(self, mode):
        """0 automatic, 1 table only, 2 partitioned whenever the counter is empty"""
        _check(lib().ok_counter_set_path(self._h, mode))

    def commit_batch(self):
        _check(lib().ok_counter_commit_batch(self._h))

    def abort_batch(self):
        _check(lib().ok_counter_abort_batch(self._h))

    def clear(self):
        _check(lib().ok_counter_clear(self._h))

    def stats(self):
        s = CounterStats()
        _check(lib().ok_counter_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def close(self):
        if getattr(self, "_h", None):
            lib().ok_counter_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def count_fastx(k, contents, min_count=1, capacity_hint=0):
    """run_count (count.rs:40-141) over in-memory files -> (sorted kmers, counts)"""
    c = KmerCounter(k, NORMALIZED, capacity_hint)
    try:
        for content in contents:
            c.add_fastx(content)
        return c.finish(min_count)
    finally:
        c.close()


def run_count(k, contents, min_count=1) -> bytes:
    """the TSV text run_count writes (count.rs:127-135)"""
    kmers, counts = count_fastx(k, contents, min_count)
    return format_counts(kmers, counts, k)


# ---- sets (build.rs, db_types.rs, compare.rs, query.rs, classify.rs) ----------------------------------
class KmerSet:
    """HashSet<u64> of one reference (db_types.rs:13), device resident and sorted."""

    def __init__(self, handle, k):
        self._h, self.k = handle, k

    @classmethod
    def build(cls, k, norm_mode=NORMALIZED, capacity_hint=0):
        h = vp()
        if not 0 <= k <= 255:
            raise InvalidKmerSize(OK_ERR_INVALID_KMER_SIZE, f"Invalid K-mer size: {k}. Must be between 1 and 32.")
        _check(lib().ok_set_create(k, norm_mode, capacity_hint, C.byref(h)))
        return cls(h, k)

    @classmethod
    def from_fastx(cls, k, content: bytes):
        """process_sequences_for_file (build.rs:23-78)"""
        s = cls.build(k)
        b = parse_fastx(content, NORMALIZED)
        s.add_batch(b.bases, b.offsets)
        return s

    @classmethod
    def from_sorted(cls, k, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        h = vp()
        _check(lib().ok_set_from_sorted(k, _ptr(kmers), len(kmers), C.byref(h)))
        return cls(h, k)

    @classmethod
    def from_sorted_device(cls, k, d_kmers_ptr, n):
        h = vp()
        _check(lib().ok_set_from_sorted_device(k, d_kmers_ptr, n, C.byref(h)))
        return cls(h, k)

    @classmethod
    def build_many_device(cls, k, d_bases_ptrs, n_bases, d_offsets_ptrs, n_records, norm_mode=NORMALIZED):
        """build.rs:93-116 for many files at once (ok_sets_build_many_device): file i is the device-resident batch
        (d_bases_ptrs[i], n_bases[i], d_offsets_ptrs[i], n_records[i]) -> list of sealed sets"""
        n = len(d_bases_ptrs)
        pb = np.ascontiguousarray(d_bases_ptrs, dtype=np.uint64)
        po = np.ascontiguousarray(d_offsets_ptrs, dtype=np.uint64)
        nb = np.ascontiguousarray(n_bases, dtype=np.uint64)
        nr = np.ascontiguousarray(n_records, dtype=np.uint64)
        assert len(po) == len(nb) == len(nr) == n
        out = (vp * n)()
        _check(lib().ok_sets_build_many_device(k, norm_mode, n, _ptr(pb), _ptr(nb), _ptr(po), _ptr(nr), out))
        return [cls(vp(h), k) for h in out]

    @classmethod
    def build_many(cls, k, batches, norm_mode=NORMALIZED):
        """build.rs:93-116 for many files at once (ok_sets_build_many): batches = [(bases, offsets), ...] host arrays"""
        keep = [(np.ascontiguousarray(b, dtype=np.uint8), np.ascontiguousarray(o, dtype=np.uint64)) for b, o in batches]
        n = len(keep)
        pb = np.array([b.ctypes.data for b, _ in keep], dtype=np.uint64)
        po = np.array([o.ctypes.data for _, o in keep], dtype=np.uint64)
        nr = np.array([len(o) - 1 for _, o in keep], dtype=np.uint64)
        out = (vp * n)()
        _check(lib().ok_sets_build_many(k, norm_mode, n, _ptr(pb), _ptr(po), _ptr(nr), out))
        return [cls(vp(h), k) for h in out]

    def add_batch(self, bases, offsets):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().ok_set_add_batch(self._h, _ptr(bases), _ptr(offsets), len(offsets) - 1))

    def add_batch_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records):
        _check(lib().ok_set_add_batch_device(self._h, d_bases_ptr, n_bases, d_offsets_ptr, n_records))

    def keys_device(self):
        """-> (device pointer, n): the sorted keys, valid while the set lives"""
        p, n = vp(), C.c_uint64()
        _check(lib().ok_set_keys_device(self._h, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def copy_keys_device(self, first, n, d_out_ptr):
        _check(lib().ok_set_copy_keys_device(self._h, first, n, d_out_ptr))

    def shard_bounds(self, n_ranks):
        """multi-GPU: index where each owner's key range begins (n_ranks + 1 entries)"""
        b = np.zeros(n_ranks + 1, dtype=np.uint64)
        _check(lib().ok_set_shard_bounds(self._h, n_ranks, _ptr(b)))
        return b

    def probe_reads_device(self, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hits_ptr, norm_mode=RAW):
        _check(lib().ok_probe_reads_device(self._h, norm_mode, d_bases_ptr, n_bases, d_offsets_ptr, n_records, d_hits_ptr))

    def __len__(self):
        n = C.c_uint64()
        _check(lib().ok_set_size(self._h, C.byref(n)))
        return n.value

    def to_array(self):
        pk, n = u64p(), C.c_uint64()
        _check(lib().ok_set_export(self._h, C.byref(pk), C.byref(n)))
        return _take(pk, None, n.value)[0]

    @staticmethod
    def union(sets):
        """get_all_kmers_unified (db_types.rs:43-48)"""
        arr = (vp * len(sets))(*[s._h for s in sets])
        h = vp()
        _check(lib().ok_set_union(arr, len(sets), C.byref(h)))
        return KmerSet(h, sets[0].k)

    def intersection_size(self, other):
        out = C.c_uint64()
        _check(lib().ok_set_intersection_size(self._h, other._h, C.byref(out)))
        return out.value

    def probe_reads(self, bases, offsets, norm_mode=RAW):
        """query.rs:79-108 per-read hit counts"""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        hits = np.zeros(len(offsets) - 1, dtype=np.uint32)
        _check(lib().ok_probe_reads(self._h, norm_mode, _ptr(bases), _ptr(offsets), len(offsets) - 1, _ptr(hits)))
        return hits

    def probe_counts(self, kmers, counts):
        """classify.rs:224-236 -> (matched, sum_depth)"""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        m, d = C.c_uint64(), C.c_uint64()
        _check(lib().ok_probe_counts(self._h, _ptr(kmers), _ptr(counts), len(kmers), C.byref(m), C.byref(d)))
        return m.value, d.value

    def close(self):
        if getattr(self, "_h", None):
            lib().ok_set_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class PeerBuffer:
    """device buffer that other ranks (processes) can map over NVLink"""

    def __init__(self, nbytes):
        p = vp()
        self.handle = (C.c_uint8 * 64)()
        _check(lib().ok_peer_buffer_create(nbytes, C.byref(p), self.handle))
        self.ptr, self.nbytes, self._peers = p.value, nbytes, []

    def handle_bytes(self):
        return bytes(self.handle)

    @staticmethod
    def open(handle_bytes):
        p = vp()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        _check(lib().ok_peer_buffer_open(buf, C.byref(p)))
        return p.value

    @staticmethod
    def close_peer(ptr):
        _check(lib().ok_peer_buffer_close(ptr))

    def destroy(self):
        if self.ptr:
            lib().ok_peer_buffer_destroy(self.ptr)
            self.ptr = None


def all_vs_all(sets):
    """sizes[n], intersection matrix[n,n] (compare.rs:51-60 for every pair)"""
    n = len(sets)
    arr = (vp * n)(*[s._h for s in sets])
    sizes = np.zeros(n, dtype=np.uint64)
    inter = np.zeros((n, n), dtype=np.uint64)
    _check(lib().ok_sets_all_vs_all(arr, n, _ptr(sizes), _ptr(inter)))
    return sizes, inter


def probe_counts_many(refs, kmers, counts):
    """classify.rs:224-277: one input count map against every reference -> (matched[n_refs], sum_depth[n_refs])"""
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    n = len(refs)
    arr = (vp * n)(*[r._h for r in refs])
    matched = np.zeros(n, dtype=np.uint64)
    depth = np.zeros(n, dtype=np.uint64)
    _check(lib().ok_probe_counts_many(arr, n, _ptr(kmers), _ptr(counts), len(kmers), _ptr(matched), _ptr(depth)))
    return matched, depth


def all_vs_all_part(sets, part, n_parts):
    """this part's share of the pairs (ok_sets_all_vs_all_part): sizes[n], upper-triangle entries of its pairs"""
    n = len(sets)
    arr = (vp * n)(*[s._h for s in sets])
    sizes = np.zeros(n, dtype=np.uint64)
    inter = np.zeros((n, n), dtype=np.uint64)
    _check(lib().ok_sets_all_vs_all_part(arr, n, part, n_parts, _ptr(sizes), _ptr(inter)))
    return sizes, inter


def finish_all_vs_all(sizes, upper):
    """summed upper-triangle matrix -> the full symmetric matrix with the set sizes on the diagonal"""
    full = upper + upper.T
    full[np.diag_indices(len(sizes))] = sizes
    return full


def compare(a: KmerSet, b: KmerSet):
    """run_compare core (compare.rs:51-66); a and b are the unified sets of the two DBs"""
    inter = a.intersection_size(b)
    na, nb = len(a), len(b)
    union = na + nb - inter
    return dict(db1=na, db2=nb, intersection_size=inter, union_size=union,
                jaccard_index=0.0 if union == 0 else inter / union)


class KmerDbV2:
    """db_types.rs:8-59"""

    def __init__(self, k):
        self.k = k
        self.references = {}

    def add_reference(self, name, kset):   # db_types.rs:38-40 (same name overwrites)
        self.references[name] = kset

    def get_all_kmers_unified(self):       # db_types.rs:43-48
        if not self.references:
            return KmerSet.from_sorted(self.k, np.zeros(0, np.uint64))
        return KmerSet.union(list(self.references.values()))

    def total_unique_kmers(self):          # db_types.rs:51-53
        return len(self.get_all_kmers_unified())

    def num_references(self):
        return len(self.references)


def run_build(k, files):
    """run_build (build.rs:80-160) over {basename: content} -> KmerDbV2"""
    if k == 0 or k > 32:
        raise InvalidKmerSize(OK_ERR_INVALID_KMER_SIZE, f"Invalid K-mer size: {k}. Must be between 1 and 32.")
    db = KmerDbV2(k)
    for name, content in files.items():
        db.add_reference(name, KmerSet.from_fastx(k, content))
    return db


def run_query(db: KmerDbV2, reads_content: bytes, min_hits=1):
    """run_query (query.rs:24-134) -> (ids in input order, per-read hits)"""
    unified = db.get_all_kmers_unified()
    b = parse_fastx(reads_content, RAW)
    hits = unified.probe_reads(b.bases, b.offsets, RAW)
    return [i for i, h in zip(b.ids, hits) if h >= min_hits], hits
