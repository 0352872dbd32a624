"""Builds liborion_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.environ.get("ORION_GPU_LIB") or os.path.join(HERE, "liborion_gpu.so")   # ORION_GPU_LIB: A/B builds of the kernels (tuning runs)
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("orion_gpu.cu", "kernels.cuh", "kmer_math.cuh", "partition.cuh", "merge.cuh", "setops.cuh")]
HEADER = os.path.join(ROOT, "include", "orion_gpu.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "20013", "-shared",
]


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liborion_gpu.so cannot be built (there is no CPU fallback)")


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in SOURCES + [HEADER])


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", SO, SOURCES[0]]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
