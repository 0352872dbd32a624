"""TEST INFRASTRUCTURE: builds CPU stand-ins of the CUDA code from a rewritten COPY of the sources (the product sources are
never touched).  What is rewritten, and nothing else:
  * `extern __shared__ [__align__(n)]`      -> `extern`            dynamic shared memory = global buffers (sim_globals)
  * the LAUNCH macro's `kern<<<...>>>(...)`  -> cusim::launch(...)  blocks one after the other, threads side by side
  * the inline-PTX helpers of partition.cuh (mbarrier, TMA bulk copies) -> tests/cusim/ptx_sim.h
  * three stray asm statements (streaming load, named barrier, proxy fence) -> their plain C++ meaning
Every rewrite asserts that its pattern is still there, so a change of the sources fails loudly here."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SIM = os.path.join(ROOT, "tests", "cusim")
CSRC = os.path.join(ROOT, "orion_kmer_b200", "csrc")
CXX = ["g++", "-std=c++17", "-O1", "-g", "-D__CUDACC__", "-pthread", "-w"]
HEADERS = ("kernels.cuh", "kmer_math.cuh", "setops.cuh", "partition.cuh", "merge.cuh")


def _sub(pattern, repl, text, what, count=0, flags=0):
    out, n = re.subn(pattern, repl, text, count=count, flags=flags)
    assert n > 0, f"cusim rewrite '{what}' no longer matches the sources"
    return out


def rewrite_headers(dst):
    for f in HEADERS:
        text = open(os.path.join(CSRC, f)).read()
        if "__shared__" in text:
            text = re.sub(r"extern\s+__shared__\s+(__align__\(\d+\)\s+)?", "extern ", text)
        if f == "kernels.cuh":
            text = _sub(r'asm volatile\("ld\.global\.nc\.L1::no_allocate\.v4\.u32[^;]*;\s*"\s*:[^;]*;',
                        "r = *reinterpret_cast<const uint4*>(p);", text, "streaming 128-bit load", flags=re.S)
        if f == "partition.cuh":
            a = text.index("__device__ __forceinline__ uint32_t ok_smem_u32")
            b = text.index("ok_tma_store_wait_all()")
            b = text.index("\n", b) + 1
            text = text[:a] + '#include "ptx_sim.h"\n' + text[b:]
            text = _sub(r'asm volatile\("bar\.sync 1, %0;"[^;]*;', "cusim::named_barrier(OK_SB_THREADS);", text, "named barrier")
            text = _sub(r'asm volatile\("fence\.proxy\.async\.shared::cta;"[^;]*;', "__threadfence();", text, "proxy fence")
            assert "asm volatile" not in text and "asm(" not in text, "inline PTX the stand-in does not know"
        open(os.path.join(dst, f), "w").write(text)


def build_setops(dst):
    """libsim_setops.so: sim_setops.cpp over the rewritten headers (the kernels alone, driven by the test)"""
    rewrite_headers(dst)
    so = os.path.join(dst, "libsim_setops.so")
    subprocess.check_call(CXX + ["-fPIC", "-shared", "-I", SIM, "-I", dst, "-o", so, os.path.join(SIM, "sim_setops.cpp")])
    return so


def build_library(dst, opt="-O1"):
    """liborion_gpu_sim.so: the WHOLE of orion_gpu.cu -- host runtime and every kernel -- against the stand-in"""
    rewrite_headers(dst)
    text = open(os.path.join(CSRC, "orion_gpu.cu")).read()
    text = _sub(r"kern<<<\(grid\), \(block\), \(smem\), \(stream\)>>>\(__VA_ARGS__\);",
                "cusim::launch((grid), (block), [&] { kern(__VA_ARGS__); }, #kern);", text, "LAUNCH macro")
    assert "<<<" not in text
    text = _sub(r'#include "setops.cuh"\n', '#include "setops.cuh"\n#include "sim_globals.h"\n', text, "include list", count=1)
    src = os.path.join(dst, "orion_gpu_sim.cpp")
    open(src, "w").write(text)
    so = os.path.join(dst, "liborion_gpu_sim.so")
    subprocess.check_call([c if c != "-O1" else opt for c in CXX] +
                          ["-fPIC", "-shared", "-fvisibility=hidden", "-I", SIM, "-I", dst, "-I", os.path.join(ROOT, "include"), "-o", so, src])
    return so
