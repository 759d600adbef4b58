#!/bin/bash
# Roofline evidence of the final tree: launch list of a short bench run + ncu --set full captures of three iterations
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-parity --no-cpu-baseline --no-extra > gpurun_out/r02_ncu_bench.log 2>&1
tail -c 300 gpurun_out/r02_ncu_bench.log
for s in 1 15 40; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_iter --launch-skip $s --launch-count 1 -f \
    -o gpurun_out/r02_iter$s python tools/prof_iter.py --ppc 4 --repeat 1 > gpurun_out/r02_iter$s.log 2>&1
  tail -1 gpurun_out/r02_iter$s.log
done
