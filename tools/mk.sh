#!/bin/bash
# usage: tools/mk.sh NAME [-DMACRO=VALUE ...]  -> 3dtk_b200/lib/NAME.so (cross-compiles here; A/B builds for tools/ab.sh)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 -shared "$@" \
  -o 3dtk_b200/lib/$name.so 3dtk_b200/csrc/b200icp.cu 3dtk_b200/csrc/host_util.cpp 3dtk_b200/csrc/lum_graph.cpp 3dtk_b200/csrc/do_icp.cpp 3dtk_b200/csrc/scan_files.cpp
