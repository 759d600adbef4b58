#!/bin/bash
# usage: tools/ab.sh lib1 lib2 ...  (names under 3dtk_b200/lib without .so) -- per-phase iteration-kernel times
for v in "$@"; do echo "== $v"; B200ICP_LIB=$PWD/3dtk_b200/lib/$v.so timeout 120 python tools/prof_iter.py --ppc ${PPC:-3} | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['nn_ms']
print('mean %.4f  early(0-9) %.3f  mid(10-28) %.3f  late(29-47) %.3f  total %.2f ms  iters %d rms %.9f'%(d['nn_ms_mean'], sum(n[:10]), sum(n[10:29]), sum(n[29:]), sum(n), d['iters'], d['rms_last']))"; done
