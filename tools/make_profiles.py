#!/usr/bin/env python
"""gpurun_out/*.ncu-rep + launch lists -> tracked summaries under profiles/ (round 1)."""
import collections, csv, json, subprocess, sys

def launches(path, out):
    rows = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(rows):
        n = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        unit = row.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'usecond': 1e3, 'msecond': 1e6, 'nsecond': 1}.get(unit, 1)
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ns
    tot = sum(a[1] for a in agg.values())
    with open(out, 'w') as f:
        f.write(f'# {path}: every launch of `bench.py --steps 1 --warmup 1` (cold-cache, serialised under ncu: compare SHARES)\n')
        f.write(f'{"launches":>8s} {"total ms":>10s} {"share":>7s}  kernel\n')
        for n, (c, t) in agg.items():
            f.write(f'{c:8d} {t/1e6:10.3f} {100*t/tot:6.1f}%  {n}\n')
    return agg

def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        res.append({h: (r[i], units[i]) for i, h in enumerate(hdr)})
    return res

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'lts__t_sectors_srcunit_tex_op_atom.sum', 'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']

def to_bytes(v, u):
    x = float(v.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)

def summary(path, out, note):
    ks = raw(path)
    traffic = {}
    with open(out, 'w') as f:
        f.write(f'# {path}\n# {note}\n')
        for k in ks:
            name = k['Kernel Name'][0].split('(')[0]
            f.write(f'\n== {name}\n')
            for m in KEEP:
                if m in k:
                    f.write(f'  {m:82s} {k[m][0]:>18s} {k[m][1]}\n')
            if 'dram__bytes_read.sum' in k:
                b = to_bytes(*k['dram__bytes_read.sum']) + to_bytes(*k['dram__bytes_write.sum'])
                f.write(f'  {"dram bytes read+write per launch":82s} {b:18.0f} byte\n')
                traffic.setdefault(name.replace('void ', ''), b)
    return traffic

if __name__ == '__main__':
    launches('gpurun_out/launches_r1.csv', 'profiles/r1_launches_table_path_2Mreads.txt')
    launches('gpurun_out/launches_r1_full.csv', 'profiles/r1_launches_partitioned_path_10Mreads.txt')
    t0 = summary('gpurun_out/prof_extract_r1.ncu-rep', 'profiles/r1_ncu_table_path_k_extract_2Mreads.txt',
                 'ncu --set full, table path (k_extract<SinkCount>), 2M reads x 150 bp, k=31; first two launches of a step')
    t1 = summary('gpurun_out/prof_part_r1_full.ncu-rep', 'profiles/r1_ncu_partitioned_path_10Mreads.txt',
                 'ncu --set full, partitioned path, FULL config 2 (10M reads x 150 bp, k=31), first step')
    json.dump({'source': 'profiles/r1_ncu_partitioned_path_10Mreads.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch, config 2)',
               **{k.split('<')[0]: v for k, v in t1.items()}}, open('profiles/ncu_traffic.json', 'w'), indent=1)
    print(json.dumps(t1, indent=1))
