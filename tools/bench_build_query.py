#!/usr/bin/env python
"""BASELINE.json configs[3] shape, scaled by --genomes: build a k-mer database from synthetic 5 Mbp genomes
(20 ancestors x descendants at 0.1-5 % divergence), then query / classify a 1M-read sample.

   python tools/bench_build_query.py [--genomes 100] [--reads 1000000] [--k 31]

Reports the device-side rows of SURVEY 8a through the C ABI: A9 set build per genome, A10 union, A12 per-read hit
counts (ok_probe_reads), A13 per-reference matched / depth (ok_probe_counts), with host buffers in and out."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orion_kmer_b200 as ok
from orion_kmer_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=100)
ap.add_argument("--genome-len", type=int, default=5_000_000)
ap.add_argument("--reads", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=31)
args = ap.parse_args()
ok.init(0)
K, G, L = args.k, args.genomes, args.genome_len
n_anc = max(1, min(20, G // 5))
ancestors = [synth.genome(100 + a, L) for a in range(n_anc)]
off1 = np.array([0, L], np.uint64)

t0 = time.perf_counter()
sets, t_dev, t_add, t_seal = [], 0.0, 0.0, 0.0
for i in range(G):
    anc = ancestors[i % n_anc]
    g = anc if i < n_anc else synth.mutate(anc, 1000 + i, 1000 + (i * 7919) % 49000)      # 0.1 % .. 5 % substitutions
    t1 = time.perf_counter()
    s = ok.KmerSet.build(K, capacity_hint=L)
    s.add_batch(g, off1)
    t2 = time.perf_counter()
    len(s)                                                                                   # seals the set (sorted, distinct)
    t3 = time.perf_counter()
    t_dev += t3 - t1; t_add += t2 - t1; t_seal += t3 - t2
    sets.append(s)
t_build = time.perf_counter() - t0
total_keys = sum(len(s) for s in sets)
print(f"A9  build: {G} genomes x {L} bases, k={K}: {t_dev * 1e3:.0f} ms through the C ABI ({G * L / t_dev / 1e9:.2f} G bases/s, "
      f"{t_dev / G * 1e3:.2f} ms per genome = {t_add / G * 1e3:.2f} create + add_batch, {t_seal / G * 1e3:.2f} seal; "
      f"with synthetic-genome generation {t_build:.1f} s), {total_keys} keys in all sets")

t_union = None
for _ in range(3):                      # best of 3: the first call also pays for the device buffers
    t0 = time.perf_counter()
    union = ok.KmerSet.union(sets)
    n_union = len(union)
    dt = time.perf_counter() - t0
    t_union = dt if t_union is None else min(t_union, dt)
print(f"A10 union: {n_union} distinct of {total_keys} keys in {t_union * 1e3:.0f} ms ({total_keys * 8 / t_union / 1e9:.0f} GB/s of input keys)")

# sample: reads from 10 of the genomes + 10 % from an unrelated genome (config-2 error recipe)
n_rel = args.reads * 9 // 10
per = n_rel // 10
chunks = [synth.reads(ancestors[a % n_anc], 500 + a, per) for a in range(10)]
chunks.append(synth.reads(synth.genome(999, L), 600, args.reads - per * 10))
bases = np.concatenate(chunks)
n_reads = len(bases) // 150
off = synth.read_offsets(n_reads)
union.probe_reads(bases[:150 * 1000], off[:1001])          # builds the hashed membership table (one-off per set)
best = None
for _ in range(3):
    t0 = time.perf_counter()
    hits = union.probe_reads(bases, off)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
W = n_reads * (150 - K + 1)
print(f"A12 query: {n_reads} reads x 150 bp against the union: {best * 1e3:.1f} ms incl. H2D of the reads and D2H of the hit counts "
      f"({len(bases) / best / 1e9:.2f} G bases/s, {(len(bases) * 1.5 + W * 8) / best / 1e9:.0f} GB/s of the probe model B*1.5 + W*8); "
      f"reads with >= 1 hit: {int((hits > 0).sum())}")

c = ok.KmerCounter(K, capacity_hint=int(len(bases) * 0.2))
c.add_batch(bases, off)
kmers, counts = c.finish(1)
c.close()
ok.probe_counts_many(sets, kmers[:1000], counts[:1000])      # builds every reference's hashed table (one-off per set)
t0 = time.perf_counter()
matched, depth = ok.probe_counts_many(sets, kmers, counts)
t_cls = time.perf_counter() - t0
m1, d1 = sets[1].probe_counts(kmers, counts)
assert (m1, d1) == (int(matched[1]), int(depth[1]))
print(f"A13 classify: {len(kmers)} input k-mers against {G} references: {t_cls * 1e3:.0f} ms incl. one H2D of the input "
      f"({t_cls / G * 1e3:.2f} ms per reference, {len(kmers) * G / t_cls / 1e9:.2f} G probes/s); best containment "
      f"{int(matched.max()) / len(kmers):.3f}")
