#!/usr/bin/env python
"""all-vs-all intersection sizes of N synthetic k-mer sets (BASELINE.json configs[4] shape, scaled by --sets/--keys):
   python tools/bench_all_vs_all.py [--sets 32] [--keys 5000000]      (ORION_INTERSECT_PLAIN=1: the per-key search)"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orion_kmer_b200 as ok

ap = argparse.ArgumentParser()
ap.add_argument("--sets", type=int, default=32)
ap.add_argument("--keys", type=int, default=5_000_000)
args = ap.parse_args()
ok.init(0)
rng = np.random.default_rng(5)
pool = np.unique(rng.integers(0, 2 ** 42, int(args.keys * 2.2), dtype=np.uint64))     # k = 21 key space
sets = []
for i in range(args.sets):           # every set = a random half of a common pool: ~half of any two sets is shared
    sel = pool[rng.random(len(pool)) < args.keys / len(pool)]
    sets.append(ok.KmerSet.from_sorted(21, sel))
ok.all_vs_all(sets[:2])
best = None
for _ in range(3):
    t0 = time.perf_counter()
    sizes, inter = ok.all_vs_all(sets)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
dt = best
pairs = args.sets * (args.sets - 1) // 2
keys = float(sum(int(sizes[i]) + int(sizes[j]) for i in range(args.sets) for j in range(i + 1, args.sets)))
a, b = sets[0].to_array(), sets[1].to_array()
assert inter[0, 1] == len(np.intersect1d(a, b, assume_unique=True))
print(f"{args.sets} sets x ~{int(sizes.mean())} keys: {pairs} pairs in {dt * 1e3:.1f} ms = {dt / pairs * 1e6:.1f} us/pair, "
      f"{keys * 8 / dt / 1e9:.0f} GB/s of the sorted-merge model (8 B x (|A| + |B|) per pair); "
      f"256 sets would take {dt / pairs * 32640:.2f} s")
