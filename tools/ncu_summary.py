#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the metrics we track under profiles/. Usage: ncu_summary.py rep [rep...]"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio']
if len(sys.argv) > 2 and sys.argv[1] == '--traffic':
    # ncu_summary.py --traffic out.json rep [rep...]: mean DRAM bytes (read+write) per launch over the reports
    import json
    vals, per = [], {}
    for rep in sys.argv[3:]:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            def val(name):
                i = hdr.index(name); v = float(r[i]); u = units[i].lower()
                return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
            b = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
            vals.append(b); per[rep.split('/')[-1]] = {'dram_bytes': b, 'gpu_time_us': float(r[hdr.index('gpu__time_duration.sum')]) * (1e3 if units[hdr.index('gpu__time_duration.sum')] == 'ms' else 1)}
    json.dump({'dram_bytes_per_launch_mean': sum(vals) / len(vals), 'launches': per,
               'source': 'ncu --set full --clock-control none, icp_iter_kernel launches (cold L2: ncu flushes caches between replays)'},
              open(sys.argv[2], 'w'), indent=1)
    sys.exit(0)
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('== %s :: %s' % (rep, name[:90]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('  %-82s %s %s' % (w, r[i], units[i]))
