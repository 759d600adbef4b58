#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full` captures of icp_iter_kernel launches (tools/gpu_roofline.sh):
DRAM bytes per launch, launch duration and CANDIDATE EVALUATIONS per launch (thread-level executions of the fp32
squared-distance evaluation `dist32`, nn_search.cuh, from the source-level counters: 3 arithmetic instructions per
evaluation).  usage: ncu_traffic.py out.json rep [rep ...]"""
import csv, json, subprocess, sys
out_path, reps = sys.argv[1], sys.argv[2:]
per, dram, evals, times = {}, [], [], []
for rep in reps:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name); v = float(r[i]); u = units[i].lower()
        return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'us': 1, 'ms': 1e3, 'ns': 1e-3, 's': 1e6}.get(u, 1)
    b = val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
    t_us = val('gpu__time_duration.sum')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                         capture_output=True, text=True).stdout
    ev, h = 0.0, None
    for row in csv.reader(src.splitlines()):
        if not row: continue
        if row[0] == 'Line No': h = row; continue
        if h is None or len(row) < len(h) or not row[0].isdigit() or row[2] != '-': continue
        if 'return fmaf(dx, dx, fmaf(dy, dy, dz * dz));' in row[1]:
            ev += float(row[h.index('Thread Instructions Executed')]) / 3.0
    name = rep.split('/')[-1]
    per[name] = {'dram_bytes': b, 'gpu_time_us': t_us, 'candidate_evaluations': ev,
                 'kernel': r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '?'}
    dram.append(b); evals.append(ev); times.append(t_us)
json.dump({'dram_bytes_per_launch_mean': sum(dram) / len(dram),
           'candidate_evaluations_per_launch_mean': sum(evals) / len(evals),
           'candidate_evaluations_per_s_mean': sum(evals) / (sum(times) * 1e-6),
           'launches': per,
           'source': 'ncu --set full --clock-control none --import-source on, icp_iter_kernel launches of iterations '
                     '1, 15 and 40 of the bench match (cold L2: ncu flushes caches between replays; durations are '
                     'ncu\'s serialised ones)', 'tool': 'tools/gpu_roofline.sh + tools/ncu_traffic.py'},
          open(out_path, 'w'), indent=1)
print(open(out_path).read())
