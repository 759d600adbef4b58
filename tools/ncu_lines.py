#!/usr/bin/env python
"""Aggregate an ncu report's source page (cuda,sass view) per CUDA source line. usage: ncu_lines.py rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = '?', None, {}
def f(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; ix = {h: i for i, h in enumerate(hdr)}; i_inst = hdr.index('Instructions Executed'); i_thr = hdr.index('Thread Instructions Executed'); i_s = hdr.index('# Samples'); i_lsb = hdr.index('stall_long_sb'); continue
    if hdr is None or len(r) <= i_lsb or not r[0].isdigit(): continue
    if r[2] != '-':   # sass row belonging to the previous cuda line
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, 0, r[1]])
    a[0] += f(r[i_inst]); a[1] += f(r[i_thr]); a[2] += f(r[i_s]); a[3] += f(r[i_lsb])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print("total warp-inst %.0f  samples %.0f" % (tot, tots))
for (fn, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print("%5.1f%%smp %5.1f%%inst lanes=%4.1f longsb=%5.0f  %s:%d  %s" % (100 * a[2] / max(tots, 1), 100 * a[0] / max(tot, 1), a[1] / max(a[0], 1), a[3], fn, ln, a[4].strip()[:90]))
