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
