#!/usr/bin/env python
"""Skew robustness of the placement.  Every placement decision is a prefix of the canonical-prior
position x(key) = 1 - (1 - u)^2, which is exact for uniform base composition only.  This tool measures what other
compositions cost: owner balance at 2/4/8 ranks (host evaluation of the owner rule on extracted k-mers, no GPU needed)
and -- with a GPU -- the step time, deferred / spilled sub-partitions and parity against the oracle on one B200.

  python tools/skew.py [--reads 2000000] [--gpu]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import orion_kmer_b200 as ok            # noqa: E402
from orion_kmer_b200 import synth       # noqa: E402

K = 31


def genome_gc(seed, n, gc):
    rng = np.random.default_rng(seed)
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    return np.frombuffer(b"ACGT", np.uint8)[rng.choice(4, size=n, p=p)].copy()


def genome_repeats(seed, n, frac=0.10, unit=1000, families=100):
    g = synth.genome(seed, n)
    rng = np.random.default_rng(seed + 1)
    fam = [synth.genome(seed + 10 + f, unit) for f in range(families)]
    for _ in range(int(n * frac / unit)):
        p = int(rng.integers(0, n - unit))
        g[p:p + unit] = fam[int(rng.integers(0, families))]
    return g


def genome_lowcomplexity(seed, n, frac=0.02):
    """2 % of the genome in microsatellites ((AC)n, (A)n, (AAT)n runs of 200-2000 bases)"""
    g = synth.genome(seed, n)
    rng = np.random.default_rng(seed + 2)
    motifs = [b"A", b"AC", b"AAT", b"AG", b"T"]
    done = 0
    while done < n * frac:
        ln = int(rng.integers(200, 2000))
        p = int(rng.integers(0, n - ln))
        m = motifs[int(rng.integers(0, len(motifs)))]
        g[p:p + ln] = np.frombuffer((m * (ln // len(m) + 1))[:ln], np.uint8)
        done += ln
    return g


def owner_shares(bases, off, n_reads, world):
    keys = np.zeros(len(bases), dtype=np.uint64)
    n = C.c_uint64()
    ok._check(ok.lib().okx_emulate_extract(ok._ptr(bases), len(bases), ok._ptr(off), n_reads, K, ok.NORMALIZED, ok._ptr(keys), len(keys), C.byref(n)))
    keys = keys[:n.value]
    owners = np.zeros(len(keys), dtype=np.int32)
    ok._check(ok.lib().okx_owner_of(ok._ptr(keys), len(keys), K, world, ok._ptr(owners)))
    share = np.bincount(owners, minlength=world) / len(keys)
    return share


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--genome", type=int, default=10_000_000)
    args = ap.parse_args()
    sets = {
        "uniform (GC 50 %)": synth.genome(3, args.genome),
        "GC 35 %": genome_gc(5, args.genome, 0.35),
        "GC 65 %": genome_gc(6, args.genome, 0.65),
        "10 % in 1 kb repeats (100 families)": genome_repeats(7, args.genome),
        "2 % microsatellites": genome_lowcomplexity(8, args.genome),
    }
    out = {}
    for name, g in sets.items():
        rec = {}
        small_n = 20_000
        sb, so = synth.reads(g, 9, small_n), synth.read_offsets(small_n)
        for w in (2, 4, 8):
            sh = owner_shares(sb, so, small_n, w)
            rec[f"owner_share_max_over_mean_{w}"] = float(sh.max() * w)
        if args.gpu:
            import oracle
            oracle.build()
            ok.init(0)
            n = args.reads
            bases, off = synth.reads(g, 9, n), synth.read_offsets(n)
            import torch
            d_b, d_o = torch.from_numpy(bases).cuda(), torch.from_numpy(off.view(np.int64)).cuda()
            keys = counts = None
            for hint in (int(len(bases) * 0.17), 0):
                c = ok.KmerCounter(K, ok.NORMALIZED, hint)
                ts = []
                try:
                    for _ in range(4):
                        c.clear()
                        t0 = time.perf_counter()
                        c.add_batch_device(d_b.data_ptr(), len(bases), d_o.data_ptr(), n)
                        c.finish_device(1)
                        ts.append(time.perf_counter() - t0)
                except ok.OrionError as e:
                    rec["hint" if hint else "no_hint"] = {"error": str(e)}
                    c.close()
                    continue
                st = c.stats()
                keys, counts = c.finish(1)
                c.close()
                tag = "hint" if hint else "no_hint"
                rec[tag] = {"ms_per_step": min(ts) * 1e3, "G_bases_per_s": len(bases) / min(ts) / 1e9, "partitioned": st["partitioned"],
                            "n_deferred": st["n_deferred"], "n_spilled": st["n_spilled"], "distinct": st["n_distinct"],
                            "phases_ms": {p: round(st[p], 3) for p in ("ms_sample", "ms_scatter1", "ms_scatter2", "ms_count", "ms_compact")}}
            nt = min(os.cpu_count() or 1, 32)
            wk, wc = oracle.count_batch_ranged_mt(K, bases, off, nt)
            rec["parity_full_table_ok"] = bool(keys is not None and np.array_equal(keys, wk) and np.array_equal(counts, wc))
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "skew.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
