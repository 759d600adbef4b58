#!/bin/bash
# usage: tools/gpu_ncu.sh LIBNAME TAG SKIP [SKIP...]  -- one `ncu --set full` capture of the SKIP-th icp_iter launch each
lib=$1; tag=$2; shift 2
mkdir -p gpurun_out
for s in "$@"; do
  B200ICP_LIB=$PWD/3dtk_b200/lib/$lib.so timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:icp_ --launch-skip $s --launch-count 1 -f -o gpurun_out/${tag}_it$s \
    python tools/prof_iter.py --ppc 4 --repeat 1 > gpurun_out/${tag}_it$s.log 2>&1
  tail -2 gpurun_out/${tag}_it$s.log
done
