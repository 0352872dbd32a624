#!/usr/bin/env python
"""Host-side rows of SURVEY 8(f) on the CPU (no GPU needed): FASTA/FASTQ framing and the count table's TSV text.
   python tools/bench_host.py [--reads 400000]"""
import argparse, ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orion_kmer_b200 as ok
from orion_kmer_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=400_000)
args = ap.parse_args()
H = ok.host_lib()
g = synth.genome(3, 5_000_000)
bases = synth.reads(g, 3, args.reads)
texts = {"FASTQ (150 bp reads)": synth.fastq_text(bases, args.reads), "FASTA (one record, 80 columns)": synth.fasta_text(b"chr1", g)}
for name, text in texts.items():
    for threads, label in ((1, "sequential"), (-1, f"automatic: up to {min(os.cpu_count() or 1, 16)} pieces cut at record starts")):
        best = None
        for _ in range(3):
            st = C.c_int()
            t0 = time.perf_counter()
            h = H.okh_fastx_parse_mt(text, len(text), 1, threads, C.byref(st))
            dt = time.perf_counter() - t0
            H.okh_batch_free(h)
            best = dt if best is None else min(best, dt)
        print(f"okh_fastx_parse {name}, {label}: {len(text) / 1e6:.0f} MB in {best * 1e3:.0f} ms = {len(text) / best / 1e9:.2f} GB/s")
rng = np.random.default_rng(1)
n = 4_000_000
keys = np.sort(rng.integers(0, 2 ** 62, n, dtype=np.uint64))
counts = rng.integers(1, 60, n).astype(np.uint64)
buf = np.empty(n * 53 + 1, np.uint8)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    m = H.okh_format_counts(ok._ptr(keys), ok._ptr(counts), n, 31, ok._ptr(buf))
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
print(f"okh_format_counts k=31: {n} lines, {m / 1e6:.0f} MB in {best * 1e3:.0f} ms = {n / best / 1e6:.1f} M lines/s, {m / best / 1e9:.2f} GB/s "
      f"({os.cpu_count()} host threads available)")
