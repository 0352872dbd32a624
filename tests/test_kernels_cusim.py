"""The set-operation kernels' OWN SOURCE executed on the CPU, thread for thread (tests/cusim/cuda_runtime.h: a block's
threads are OS threads, __syncthreads a barrier, warp shuffles an exchange, atomics the GCC builtins), against numpy --
and under ThreadSanitizer / AddressSanitizer, which turn a missing barrier or an out-of-range index in the kernel code
into a report.  This does not replace the GPU parity tests (tests/test_gpu_parity.py runs the same shapes through the
C ABI on a B200); it checks indexing, barrier placement and arithmetic of setops.cuh (k_ava_bounds, k_ava_tiles) and
of the tiled intersection / membership kernels of kernels.cuh where no GPU is at hand.

The kernel headers are compiled from a copy in which only `extern __shared__` is rewritten to `extern` (dynamic shared
memory becomes one global buffer); nothing else of the source is touched."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import orion_kmer_b200 as ok

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM = os.path.join(ROOT, "tests", "cusim")
CSRC = os.path.join(ROOT, "orion_kmer_b200", "csrc")
CXX = ["g++", "-std=c++17", "-O1", "-g", "-D__CUDACC__", "-pthread", "-w"]


@pytest.fixture(scope="module")
def simdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("cusim")
    for f in ("kernels.cuh", "kmer_math.cuh", "setops.cuh"):
        text = open(os.path.join(CSRC, f)).read()
        (d / f).write_text(re.sub(r"extern\s+__shared__", "extern", text))
    return d


@pytest.fixture(scope="module")
def sim(simdir):
    so = simdir / "libsim_setops.so"
    subprocess.check_call(CXX + ["-fPIC", "-shared", "-I", SIM, "-I", str(simdir), "-o", str(so), os.path.join(SIM, "sim_setops.cpp")])
    L = C.CDLL(str(so))
    L.sim_ava_keyed.restype = C.c_int
    L.sim_ava_keyed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint32, C.c_uint64, C.c_uint, C.c_uint, C.c_void_p]
    L.sim_member.restype = C.c_uint64
    L.sim_member.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint, C.c_void_p]
    L.sim_intersect.restype = C.c_uint64
    L.sim_intersect.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint]
    return L


def keyed(sim, sets, k, grid=3):
    """ok_sets_all_vs_all's keyed form as orion_gpu.cu drives it: geometry from the library's own host code
    (okx_ava_geometry), then k_ava_bounds + k_ava_tiles in the simulator -> (failed flag, n x n matrix, n_tiles)"""
    sets = [np.ascontiguousarray(s, dtype=np.uint64) for s in sets]
    n = len(sets)
    ns = np.array([len(s) for s in sets], np.uint64)
    ends = np.zeros(2 * n, np.uint64)
    for i, s in enumerate(sets):
        if len(s):
            ends[2 * i], ends[2 * i + 1] = s[0], s[-1]
    geo, none, tiles = np.zeros(3, np.uint64), np.zeros(1, np.uint64), np.zeros(1, np.uint32)
    assert ok.lib().okx_ava_geometry(k, ok._ptr(ends), ok._ptr(ns), n, int(ns.sum()), ok._ptr(none), 0, ok._ptr(geo), ok._ptr(tiles)) == 0
    ptrs = np.array([s.ctypes.data if len(s) else 0 for s in sets], np.uint64)
    out = np.zeros((n, n), np.uint64)
    failed = sim.sim_ava_keyed(ptrs.ctypes.data, ns.ctypes.data, n, 64 - 2 * k, int(geo[1]), int(geo[2]), int(geo[0]), grid, out.ctypes.data)
    return failed, out, int(geo[0])


def check_matrix(sets, out):
    n = len(sets)
    for i in range(n):
        for j in range(n):
            want = len(np.intersect1d(sets[i], sets[j], assume_unique=True)) if i < j else 0      # entries above the diagonal only
            assert out[i, j] == want, (i, j)


def canonical_like(rng, n, bits=42):
    return np.unique(np.minimum(rng.integers(0, 1 << bits, n, dtype=np.uint64), rng.integers(0, 1 << bits, n, dtype=np.uint64)))


def test_keyed_all_vs_all_families_and_odd_shapes(sim):
    """13 sets (not a multiple of the 8 x 8 pair blocks): related families, an identical copy, a tiny set, an empty set,
    keys nobody else holds; several tiles per CTA (grid 3) and one CTA for all tiles (grid 1)"""
    rng = np.random.default_rng(5)
    pool = canonical_like(rng, 30_000)
    sets = [pool[rng.random(len(pool)) < f] for f in (0.6, 0.5, 0.9, 0.0, 0.2, 1.0, 0.05, 0.5, 0.7, 0.3, 0.8)]
    sets.append(sets[2].copy())
    sets.append(np.unique(np.concatenate([pool[:40], canonical_like(rng, 3000)])))
    for grid in (3, 1):
        failed, out, n_tiles = keyed(sim, sets, 21, grid)
        assert failed == 0 and n_tiles > 20
        check_matrix(sets, out)


def test_keyed_all_vs_all_256_sets_and_a_narrow_key_range(sim):
    """256 sets: all 528 pair blocks are owned by a thread; then a multi-GPU shard (1/8 of the position space), 2 sets"""
    rng = np.random.default_rng(6)
    pool = canonical_like(rng, 6000)
    sets = [pool[rng.random(len(pool)) < 0.02 + 0.1 * (i % 7) / 7] for i in range(256)]
    failed, out, n_tiles = keyed(sim, sets, 21, 2)
    assert failed == 0 and n_tiles >= 8
    check_matrix(sets, out)
    lo, hi = np.uint64(1 << 39), np.uint64(3 << 38)
    shard = [s[(s >= lo) & (s < hi)] for s in (canonical_like(rng, 40_000), canonical_like(rng, 40_000), pool, pool[::2])]
    failed, out, n_tiles = keyed(sim, shard, 21, 2)
    assert failed == 0
    check_matrix(shard, out)
    failed, out, _ = keyed(sim, [pool, pool[::3]], 21, 1)
    assert failed == 0 and out[0, 1] == len(pool[::3])


def test_keyed_all_vs_all_reports_a_tile_that_does_not_fit(sim):
    """keys sharing their first 16 bases share one position: their tile outgrows the 6144-entry table, the pass reports it
    (the library then computes the matrix row by row) and never writes out of bounds doing so"""
    rng = np.random.default_rng(9)
    prefix = np.uint64(0x1234567) << np.uint64(30 + 2)
    pool = np.unique(rng.integers(0, 1 << 30, size=20_000, dtype=np.uint64)) | prefix
    spread = canonical_like(rng, 20_000, 62)
    sets = [np.unique(np.concatenate([pool[rng.random(len(pool)) < 0.5], spread[rng.random(len(spread)) < 0.3]])) for _ in range(4)]
    failed, _, _ = keyed(sim, sets, 31, 2)
    assert failed == 1


def test_tiled_intersection_and_membership_shapes(sim):
    """k_intersect_bounds + k_intersect_tiled / k_member_tiled: equal sets, disjoint ranges, a small set inside a large one
    (one tile of A against many chunks of B), interleaved keys, sizes around the tile size, empty B"""
    rng = np.random.default_rng(3)

    def uniq(n, bits=50):
        return np.unique(rng.integers(0, 1 << bits, n, dtype=np.uint64))
    big = uniq(60_000)
    shapes = [(big, big), (big[:20_000], big[30_000:]), (big[::50], big), (big, big[::50]), (big[::2], big[1::2]),
              (uniq(2047), uniq(2049)), (uniq(2048), big), (uniq(1), big), (big[:5000], np.zeros(0, np.uint64)),
              (np.sort(np.concatenate([big[::7], uniq(3000)])), big)]
    for a, b in shapes:
        a, b = np.ascontiguousarray(np.unique(a)), np.ascontiguousarray(np.unique(b))
        want = np.intersect1d(a, b, assume_unique=True)
        for grid in (1, 3):
            assert sim.sim_intersect(a.ctypes.data, len(a), b.ctypes.data if len(b) else 0, len(b), grid) == len(want)
            out = np.zeros(len(a) + 1, np.uint64)
            m = sim.sim_member(a.ctypes.data, len(a), b.ctypes.data if len(b) else 0, len(b), grid, out.ctypes.data)
            assert m == len(want) and np.array_equal(np.sort(out[:m]), want)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_kernels_are_clean_under_the_sanitizers(simdir, sanitizer):
    """the same kernels, self-checking (tests/cusim/sim_main.cpp), under TSan (a data race between barriers = a missing
    __syncthreads or a non-atomic update) and ASan + UBSan (an index out of range in shared or global memory)"""
    exe = simdir / ("sim_" + sanitizer.split(",")[0])
    r = subprocess.run(CXX + ["-fsanitize=" + sanitizer, "-I", SIM, "-I", str(simdir), "-o", str(exe), os.path.join(SIM, "sim_main.cpp")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("this compiler cannot build with -fsanitize=" + sanitizer)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([str(exe), "9", "8000"], capture_output=True, text=True, timeout=900, env=env)
    if "unexpected memory mapping" in r.stderr or "ThreadSanitizer: CHECK failed" in r.stderr:
        pytest.skip("the sanitizer runtime does not start in this environment")
    assert r.returncode == 0 and "mismatches 0" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
    assert "ThreadSanitizer" not in r.stderr and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
