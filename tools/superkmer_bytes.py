#!/usr/bin/env python
"""How many bytes the multi-GPU exchange would carry if it routed 2-bit super-k-mers (maximal runs of consecutive windows
of a read sharing one canonical minimizer, owner = minimizer's position) instead of 8-byte canonical k-mers (DESIGN.md
12.2).  CPU analysis on the configs[1] read recipe; no GPU needed.

  python tools/superkmer_bytes.py [--reads 20000] [--m 11 13 15 17]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orion_kmer_b200 import synth       # noqa: E402

K, L = 31, 150


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20_000)
    ap.add_argument("--m", type=int, nargs="+", default=[11, 13, 15, 17, 19])
    a = ap.parse_args()
    g = synth.genome(3, a.reads * 5 * 50)
    reads = synth.reads(g, 3, a.reads).reshape(a.reads, L)
    code = np.full(256, 4, np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
        code[c + 32] = i
    c = code[reads].astype(np.uint64)                        # 4 = N
    valid_base = c < 4
    c = np.where(valid_base, c, 0)
    rows = []
    for m in a.m:
        # canonical m-mers, hashed (a minimizer ordering that is not lexicographic)
        nm = L - m + 1
        fw = np.zeros((a.reads, nm), np.uint64)
        rc = np.zeros((a.reads, nm), np.uint64)
        for j in range(m):
            fw = (fw << np.uint64(2)) | c[:, j:j + nm]
            rc = rc | ((np.uint64(3) - c[:, j:j + nm]) << np.uint64(2 * j))
        can = np.minimum(fw, rc)
        h = (can * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(20)
        ok_m = np.ones((a.reads, nm), bool)
        for j in range(m):
            ok_m &= valid_base[:, j:j + nm]
        h = np.where(ok_m, h, np.uint64(2 ** 63))            # an m-mer with an N never wins
        # minimizer of window w (k-mer at w): argmin of h over the k - m + 1 m-mers inside it
        nw = L - K + 1
        span = K - m + 1
        win = np.lib.stride_tricks.sliding_window_view(h, span, axis=1)          # (reads, nw, span)
        pos = win.argmin(axis=2) + np.arange(nw)[None, :]                          # absolute position of the minimizer
        ok_w = np.ones((a.reads, nw), bool)
        for j in range(K):
            ok_w &= valid_base[:, j:j + nw]
        # a super-k-mer breaks where the minimizer position changes or a window is invalid
        brk = np.ones((a.reads, nw), bool)
        brk[:, 1:] = (pos[:, 1:] != pos[:, :-1]) | ~ok_w[:, :-1]
        brk &= ok_w
        n_super = int(brk.sum())
        n_win = int(ok_w.sum())
        bases = n_win + n_super * (K - 1)                                          # a run of r windows holds r + k - 1 bases
        payload = bases / 4.0 + n_super * 2.0                                      # 2 bits per base + a 2-byte length header
        rows.append((m, n_win, n_super, n_win / n_super, payload / n_win, 8.0 * n_win / payload))
    print(f"{a.reads} reads x {L} bp, k = {K}, substitution 0.5 %, N 0.1 % (bench.py recipe)")
    print("  m   windows  super-k-mers  windows/super-k-mer  bytes/window  reduction vs 8 B per k-mer")
    for r in rows:
        print("%3d %9d %13d %20.2f %13.3f %10.2fx" % r)


if __name__ == "__main__":
    main()
