#!/usr/bin/env python
"""The set operations for the profiler: build N genome sets (k = 21, the configs[3]/[4] collection), their keyed
all-vs-all (k_ava_*), their union (strided level-1 gather) and a query by merge (k_member_tiled) of a read sample.
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ... python tools/prof_sets.py`"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("ORION_PROBE_MERGE", "1")       # (a set of this size would be hashed otherwise)
import torch                            # noqa: E402,F401
import orion_kmer_b200 as ok            # noqa: E402
from orion_kmer_b200 import synth       # noqa: E402
import bench_sets                       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sets", type=int, default=32)
ap.add_argument("--reads", type=int, default=200_000)
ap.add_argument("--k", type=int, default=21)
a = ap.parse_args()
ok.init(0)
L = bench_sets.GENOME_LEN
# every 8th genome of the collection: four genomes per ancestor family, 0.1 % .. 5 % diverged
gens = [bench_sets.genome(synth, 8 * i + 3) for i in range(a.sets)]
off1 = np.array([0, L], np.uint64)
sets = ok.KmerSet.build_many(a.k, [(g, off1) for g in gens])
sizes, inter = ok.all_vs_all(sets)
union = ok.KmerSet.union(sets)
reads = np.concatenate([synth.reads(gens[0], 7, a.reads // 2), synth.reads(synth.genome(999, L), 8, a.reads - a.reads // 2)])
hits = union.probe_reads(reads, synth.read_offsets(a.reads), ok.RAW)
print("sets", a.sets, "keys", int(sizes.sum()), "union", len(union), "inter[0,1]", int(inter[0, 1]), "inter[0,4]", int(inter[0, 4]),
      "reads hit", int((hits > 0).sum()), "launches", ok.launch_count())
