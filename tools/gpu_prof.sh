#!/bin/bash
# usage: tools/gpu_prof.sh LIB...  -- full per-iteration profile JSON of each variant into gpurun_out/prof_LIB.json
mkdir -p gpurun_out
for v in "$@"; do B200ICP_LIB=$PWD/3dtk_b200/lib/$v.so timeout 300 python tools/prof_iter.py --ppc 4 > gpurun_out/prof_$v.json 2>gpurun_out/prof_$v.err; tail -c 300 gpurun_out/prof_$v.err; done
