#!/usr/bin/env python
"""What boundaries taken from the sample's quantiles would buy on skewed inputs -- a CPU analysis,
no GPU needed.  Every placement of the count path is a prefix of the position x(key) = 1 - (1 - u)^2, cut at EQUAL steps
of x: exact for uniform base composition only.  Here the k-mers of a read set are extracted with the host replay of the
extraction kernel (okx_emulate_extract: the same __host__ __device__ code) and the share of the fullest owner / level-1
bin is computed for (a) equal steps of x and (b) boundaries at the quantiles of x over a 1/16 sample of the reads (what
the planner already samples), still a monotone map of the key, so sub-partition order stays key order.

  python tools/skew_quantiles.py [--reads 40000]      -> profiles/r2_skew_quantile_analysis.json
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import orion_kmer_b200 as ok            # noqa: E402
from orion_kmer_b200 import synth       # noqa: E402
import skew                             # noqa: E402

K = 31


def kmers(bases, off, n_reads):
    keys = np.zeros(len(bases), dtype=np.uint64)
    n = C.c_uint64()
    ok._check(ok.lib().okx_emulate_extract(ok._ptr(bases), len(bases), ok._ptr(off), n_reads, K, ok.NORMALIZED, ok._ptr(keys), len(keys), C.byref(n)))
    return keys[:n.value]


def phi32(keys):
    """top 32 bits of ok_canon_pos (kmer_math.cuh): x = 1 - (1 - u)^2 on the top 32 bits of u = key << (64 - 2k)"""
    u = keys << np.uint64(64 - 2 * K)
    w = (~u) >> np.uint64(32)
    return ((~(w * w)) >> np.uint64(32)).astype(np.uint64)


def fullest(x, cuts):
    """max share / mean share over the parts [cuts[i], cuts[i+1])"""
    part = np.searchsorted(cuts, x, side="right")
    cnt = np.bincount(part, minlength=len(cuts) + 1)
    return float(cnt.max() * (len(cuts) + 1) / len(x))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=40_000)
    ap.add_argument("--genome", type=int, default=10_000_000)
    args = ap.parse_args()
    sets = {
        "uniform (GC 50 %)": synth.genome(3, args.genome),
        "GC 35 %": skew.genome_gc(5, args.genome, 0.35),
        "GC 65 %": skew.genome_gc(6, args.genome, 0.65),
        "10 % in 1 kb repeats (100 families)": skew.genome_repeats(7, args.genome),
        "2 % microsatellites": skew.genome_lowcomplexity(8, args.genome),
    }
    out = {}
    n = args.reads
    for name, g in sets.items():
        bases, off = synth.reads(g, 9, n), synth.read_offsets(n)
        keys = kmers(bases, off, n)
        x = phi32(keys)
        # the sample: every 16th read (the planner samples 1/16 of the tiles)
        xs = np.sort(phi32(kmers(bases[:(n // 16) * 150], off[:n // 16 + 1], n // 16)))
        owners = np.zeros(len(keys), dtype=np.int32)
        rec = {"windows": int(len(keys)), "sample_windows": int(len(xs))}
        for parts in (2, 4, 8, 256, 4096):
            equal = (np.arange(1, parts, dtype=np.uint64) << np.uint64(32)) // np.uint64(parts)
            quant = xs[(np.arange(1, parts) * len(xs)) // parts]
            rec[f"{parts}_parts"] = {"equal_steps_max_over_mean": round(fullest(x, equal), 3),
                                     "sample_quantiles_max_over_mean": round(fullest(x, quant), 3)}
            if parts <= 8:      # cross-check the equal steps against the library's own owner rule
                ok._check(ok.lib().okx_owner_of(ok._ptr(keys), len(keys), K, parts, ok._ptr(owners)))
                lib_share = float(np.bincount(owners, minlength=parts).max() * parts / len(keys))
                assert abs(lib_share - rec[f"{parts}_parts"]["equal_steps_max_over_mean"]) < 2e-3, (lib_share, rec)
        out[name] = rec
        print(name, json.dumps(rec), flush=True)
    with open(os.path.join(ROOT, "profiles", "r2_skew_quantile_analysis.json"), "w") as f:
        json.dump({"what": "fullest part / mean part of the k-mer stream (k = 31) under equal steps of the closed-form position "
                           "(the current rule) and under boundaries at the quantiles of a 1/16 read sample; CPU analysis "
                           "(tools/skew_quantiles.py), not a GPU measurement", "reads": n, "genome": args.genome, "cases": out}, f, indent=1)


if __name__ == "__main__":
    main()
