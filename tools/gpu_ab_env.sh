#!/bin/bash
# usage: tools/gpu_ab_env.sh LIB "ENV=.. ENV=.." "ENV=.." ...   per-phase iteration times of one library under several environments
lib=$1; shift
for e in "$@"; do echo "== $lib $e"; env $e B200ICP_LIB=$PWD/3dtk_b200/lib/$lib.so timeout 300 python tools/prof_iter.py | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=[a+b for a,b in zip(d['nn_ms'],d['stream_ms'])]
print('it0 %.3f it1 %.3f'%(n[0],n[1]), end='  '); print('early(0-9) %.3f  mid(10-28) %.3f  late(29-47) %.3f  total %.2f ms  iters %d rms %.12f'%(sum(n[:10]), sum(n[10:29]), sum(n[29:]), sum(n), d['iters'], d['rms_last']))
print('searches', d['searches'][::4]); print('ms', d['nn_ms'][10::4])"; done
