timeout 900 python tools/bench_rows.py > gpurun_out/rows_bench.json 2> gpurun_out/rows_bench.err; tail -3 gpurun_out/rows_bench.err; python -c "
import json
for l in open('gpurun_out/rows_bench.json'):
    r=json.loads(l); print('%-70s gpu %8.2f ms  %10.3g %s/s  %6.1f GB/s (%.3f of peak)  cpu %s  x%.0f' % (r['row'][:70], r['gpu_ms'], r['gpu_units_per_s'], r['unit'], r['achieved_gbs'], r['frac_of_hbm_peak'], ('%.3g/s'%r['cpu']['units_per_s']) if r['cpu']['units_per_s'] else '-', r['speedup_vs_cpu'] or 0))
"
