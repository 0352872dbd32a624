#!/usr/bin/env python
"""Per-kernel, per-source-line instruction counts and stall samples from an .ncu-rep
(needs -lineinfo + --import-source on).

  python tools/ncu_lines.py report.ncu-rep [top_lines_per_kernel] [kernel_substring]
"""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ''
rows = list(csv.reader(out.splitlines()))
cur_file, cur_fn, hdr, line = None, None, None, None
agg = {}   # fn -> {(file, line, src): [warp_inst, samples, thread_inst]}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        cur_fn = r[1].split('(')[0].replace('void ', '')
        continue
    if r[0] == 'Line No':
        hdr = r
        iex = hdr.index('Instructions Executed'); isa = hdr.index('# Samples'); ith = hdr.index('Thread Instructions Executed')
        continue
    if hdr is None or len(r) <= iex:
        continue
    if r[0] != '':
        line = (cur_file, r[0], r[1].strip()[:110])
        agg.setdefault(cur_fn, {}).setdefault(line, [0, 0, 0])
        continue
    try:
        a = agg[cur_fn][line]
        a[0] += int(r[iex]); a[1] += int(r[isa]); a[2] += int(r[ith])
    except (ValueError, KeyError, TypeError):
        pass
for fn, lines in agg.items():
    if want not in fn:
        continue
    tot = sum(v[0] for v in lines.values()) or 1
    tots = sum(v[1] for v in lines.values()) or 1
    print(f'== {fn}: warp-inst {tot}  samples {tots}')
    for (f, ln, src), (ex, sa, th) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'  {100*ex/tot:5.1f}% inst {100*sa/tots:5.1f}% samp  act {th/max(ex,1):4.1f}  {f}:{ln:>4s}  {src}')
