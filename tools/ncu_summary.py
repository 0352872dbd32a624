#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters the design notes cite."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
only = sys.argv[2] if len(sys.argv) > 2 else None
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if only and only not in name:
        continue
    print('----', name[:90])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'  {w:82s} {r[i]:>18s} {units[i]}')
