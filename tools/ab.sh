#!/bin/bash
# usage: tools/ab.sh lib1 lib2 ...  (names under 3dtk_b200/lib without .so) -- per-phase iteration-kernel times
for v in "$@"; do echo "== $v"; B200ICP_LIB=$PWD/3dtk_b200/lib/$v.so timeout 300 python tools/prof_iter.py --ppc ${PPC:-3} | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=[a+b for a,b in zip(d['nn_ms'],d['stream_ms'])]
print('it0 %.3f it1 %.3f stream_mean %.4f'%(n[0],n[1],sum(d['stream_ms'])/len(n)), end='  '); print('mean %.4f  early(0-9) %.3f  mid(10-28) %.3f  late(29-47) %.3f  total %.2f ms  iters %d rms %.9f'%(sum(n)/len(n), sum(n[:10]), sum(n[10:29]), sum(n[29:]), sum(n), d['iters'], d['rms_last']))"; done
