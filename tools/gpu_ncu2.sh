#!/bin/bash
# usage: tools/gpu_ncu2.sh LIB TAG MAXITER COUNT -- ncu --set full of the first COUNT match-kernel launches with --max-iter MAXITER
lib=$1; tag=$2; mi=$3; cnt=$4
mkdir -p gpurun_out
B200ICP_LIB=$PWD/3dtk_b200/lib/$lib.so timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:icp_ --launch-skip 0 --launch-count $cnt -f -o gpurun_out/$tag \
  python tools/prof_iter.py --ppc 4 --repeat 1 --max-iter $mi > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log
