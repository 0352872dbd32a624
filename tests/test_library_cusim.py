"""The WHOLE device library -- the host runtime of orion_gpu.cu and every kernel it launches -- compiled against the CPU
stand-in of tests/cusim (see cuda_runtime.h there) and driven through the same C ABI and ctypes mirror as on a B200:
  * the golden-vector tests of tests/test_gpu_parity.py, verbatim, in a process whose ORION_GPU_LIB is the stand-in;
  * tests/cusim/sim_cases.py: the partitioned count path (incl. the TMA-fed level-2 scatter and the dense look-back
    output) with a second batch merged, query by merge, many files built side by side, union / compare / classify.
This is a check of LOGIC (indexing, barriers, host orchestration, parity with the oracle) where no GPU is at hand; it
says nothing about speed and does not replace `-m gpu`.  CUSIM_FULL=1 adds the cases that take minutes (the strided
union; `-k` subsets of test_gpu_parity.py can be run the same way by hand -- profiles/r2_cusim_runs.txt)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cusim"))


@pytest.fixture(scope="module")
def simlib(tmp_path_factory):
    import build_sim
    return build_sim.build_library(str(tmp_path_factory.mktemp("cusimlib")))


def _pytest(simlib, *args, timeout=1500):
    env = dict(os.environ, ORION_GPU_LIB=simlib, CUSIM_SMS="2")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", *args], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


def test_golden_gpu_tests_pass_on_the_simulated_library(simlib):
    out = _pytest(simlib, "tests/test_gpu_parity.py", "-m", "gpu", "-k",
                  "golden or edge_cases or invalid_k or fixture_files or clustered_keys")
    assert " passed" in out and "failed" not in out


def test_simulated_library_cases(simlib):
    out = _pytest(simlib, os.path.join("tests", "cusim", "sim_cases.py"))
    assert " passed" in out and "failed" not in out


@pytest.mark.skipif(not os.environ.get("CUSIM_FULL"), reason="several minutes (two sanitizer builds of the whole library): CUSIM_FULL=1")
@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_whole_library_under_the_sanitizers(simlib, sanitizer):
    """tests/cusim/sim_lib_main.cpp: the whole simulated library in one program, driven through the C ABI (table path,
    partitioned path + merge, sets, keyed all-vs-all, union, query by merge) and checked against the oracle's counter.
    ASan + UBSan: no report at all.  TSan: the only conflicting accesses allowed are the optimistic first probes of the
    hash tables -- a plain 8-byte READ of a key slot against another thread's atomicCAS claim of it, which the CAS that
    follows arbitrates (kernels.cuh ok_ld_key, partition.cuh k_part_count / ok_c2_insert_slow); anything else is a race."""
    import re
    d = os.path.dirname(simlib)
    exe = os.path.join(d, "sim_lib_" + sanitizer.split(",")[0])
    import build_sim
    r = subprocess.run(build_sim.CXX + ["-fsanitize=" + sanitizer, "-I", build_sim.SIM, "-I", d, "-I", os.path.join(ROOT, "include"), "-o", exe,
                                        os.path.join(build_sim.SIM, "sim_lib_main.cpp"), "-L", os.path.join(ROOT, "oracle"), "-lorion_oracle",
                                        "-Wl,-rpath," + os.path.join(ROOT, "oracle")], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("this compiler cannot build with -fsanitize=" + sanitizer)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0", ASAN_OPTIONS="detect_leaks=0", CUSIM_SMS="2")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=3000, env=env)
    assert "mismatches 0" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    reports = r.stderr.split("WARNING: ThreadSanitizer: data race")[1:]
    for rep in reports:
        head = [ln.strip() for ln in rep.splitlines() if re.match(r"\s+(Read|Write|Previous|Atomic)", ln)]
        assert len(head) == 2 and head[0].startswith("Read of size 8") and head[1].startswith("Previous atomic write of size 8"), rep[:1500]
        assert "atomicCAS" in rep.split("Previous atomic write")[1].split("\n\n")[0], rep[:1500]
