#!/usr/bin/env python
"""A bare count step for the profiler: config 2 (10 M x 150 bp, k = 31, bench hint), batch resident in HBM, no CPU
baseline.  `ncu ... python tools/prof_step.py [--steps 2] [--merge] [--sets 6]`:
  --merge   a second batch, so that the run merge (k_merge_*) shows up
  --sets N  also build N genome sets (k = 21) and run their all-vs-all (k_intersect_row_*)"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                            # noqa: E402
import orion_kmer_b200 as ok            # noqa: E402
from orion_kmer_b200 import synth       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--reads", type=int, default=10_000_000)
ap.add_argument("--merge", action="store_true")
ap.add_argument("--sets", type=int, default=0)
a = ap.parse_args()
ok.init(0)
g = synth.genome(3, a.reads * 5)
bases, off = synth.reads(g, 3, a.reads), synth.read_offsets(a.reads)
d_b, d_o = torch.from_numpy(bases).cuda(), torch.from_numpy(off.view(np.int64)).cuda()
c = ok.KmerCounter(31, ok.NORMALIZED, int(len(bases) * 0.17))
for _ in range(a.steps):
    c.clear()
    c.add_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), a.reads)
    if a.merge:
        c.add_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), a.reads)
    _, _, n = c.finish_device(1)
print("distinct", n, c.stats())
c.close()
if a.sets:
    import bench_sets
    sets = []
    for i in range(a.sets):
        s = ok.KmerSet.build(21)
        gi = bench_sets.genome(synth, i * 25)
        s.add_batch(gi, np.array([0, len(gi)], np.uint64))
        sets.append(s)
    sizes, inter = ok.all_vs_all(sets)
    print("all-vs-all", sizes[:3], inter[0, :3])
    for s_ in sets:
        s_.close()
