#!/usr/bin/env python
"""A/B of the set build (build.rs:93-116 over many files): one file at a time from Python, ok_sets_build_many_device
with 1 / 2 / 4 / 8 host threads, and what closing the sets costs.  `python tools/build_ab.py [--genomes 64]`."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                            # noqa: E402
import orion_kmer_b200 as ok            # noqa: E402
from orion_kmer_b200 import synth       # noqa: E402
import bench_sets                       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=64)
ap.add_argument("--k", type=int, default=31)
a = ap.parse_args()
ok.init(0)
L = bench_sets.GENOME_LEN
h = np.empty(a.genomes * L, np.uint8)
for i in range(a.genomes):
    h[i * L:(i + 1) * L] = bench_sets.genome(synth, i)
d = torch.from_numpy(h).cuda()
d_off = torch.tensor([0, L], dtype=torch.int64, device="cuda")
ptrs = [d.data_ptr() + i * L for i in range(a.genomes)]
out = {}


def one_by_one():
    sets = []
    for p in ptrs:
        s = ok.KmerSet.build(a.k)
        s.add_batch_device(p, L, d_off.data_ptr(), 1)
        len(s)
        sets.append(s)
    return sets


def many(threads):
    os.environ["ORION_BUILD_THREADS"] = str(threads)
    return ok.KmerSet.build_many_device(a.k, ptrs, [L] * a.genomes, [d_off.data_ptr()] * a.genomes, [1] * a.genomes)


for name, fn in [("python_loop", one_by_one)] + [(f"build_many_{t}", (lambda t=t: many(t))) for t in (1, 2, 4, 8)] + [("python_loop_again", one_by_one)]:
    best, close_ms = None, None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sets = fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for s in sets:
            s.close()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if rep:                                  # the first repetition warms the pooled builders up
            best = min(best or 1e9, (t1 - t0) * 1e3 / a.genomes)
            close_ms = (t2 - t1) * 1e3 / a.genomes
    out[name] = {"ms_per_genome": round(best, 4), "close_ms_per_set": round(close_ms, 4)}
print(json.dumps({"genomes": a.genomes, "k": a.k, "genome_len": L, "launches": ok.launch_count(), "modes": out}))
