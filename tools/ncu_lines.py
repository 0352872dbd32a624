#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo + --import-source)."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; iex = hdr.index('Instructions Executed'); isa = hdr.index('# Samples'); ith = hdr.index('Thread Instructions Executed'); continue
    if hdr is None or len(r) <= iex:
        continue
    if r[0] != '':
        line = (cur_file, r[0], r[1].strip()[:100])
        agg.setdefault(line, [0, 0, 0])
        continue
    try:
        agg[line][0] += int(r[iex]); agg[line][1] += int(r[isa]); agg[line][2] += int(r[ith])
    except (ValueError, NameError):
        pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[1] for v in agg.values()) or 1
print(f'total warp-inst {tot}  samples {tots}')
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for (f, ln, src), (ex, sa, th) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100*sa/tots:5.1f}% samp {100*ex/tot:5.1f}% inst  act {th/max(ex,1):4.1f}  {f}:{ln:>4s}  {src}')
