#!/usr/bin/env python
"""Per-block timing of the fused iteration kernel's LAST launch of a match (needs a -DB200_TIMING build:
tools/build_variant.sh timing -DB200_TIMING).  usage: block_times.py [max_iter ...]
Shows how evenly the blocks / SMs finish: the kernel lasts as long as its slowest SM."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200ICP_LIB", os.path.join(ROOT, "3dtk_b200", "lib", "timing.so"))
icp = importlib.import_module("3dtk_b200")
n = 1_000_000
ctx = icp.Context(0)
model = icp.synth_scene(7, 42, n, 0.5); data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
m = icp.Scan(ctx, model, max_dist_hint=25.0); d = icp.Scan(ctx, data, max_dist_hint=25.0)
icp.lib.b200icp_debug_blocks.argtypes = [C.c_void_p, C.c_int]
NB = 444
for iters in [int(x) for x in sys.argv[1:]] or [1, 3, 4, 15, 40]:
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=iters, epsilon_icp=1e-5)
    for _ in range(2):
        d.set_pose(np.eye(4).reshape(16), np.eye(4).reshape(16)); r = eng.match(m, d)
    buf = (C.c_ulonglong * (NB * 4))()
    icp.lib.b200icp_debug_blocks(buf, NB)
    a = np.array(buf, dtype=np.uint64).reshape(NB, 4).astype(np.int64)
    t0 = a[:, 0].min()
    start, loop, end, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2] - t0, a[:, 3]
    print("launches %d (note: the last launch may be a no-op after convergence); block end us: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f | start spread %.1f | loop-done p50 %.1f max %.1f"
          % (iters, end.min() / 1e3, np.percentile(end, 10) / 1e3, np.percentile(end, 50) / 1e3, np.percentile(end, 90) / 1e3,
             end.max() / 1e3, start.max() / 1e3, np.percentile(loop, 50) / 1e3, loop.max() / 1e3))
    per_sm = {}
    for b in range(NB):
        per_sm.setdefault(int(sm[b]), []).append(end[b] / 1e3)
    sm_end = np.array([max(v) for v in per_sm.values()])
    order = np.argsort(end)
    print("   SMs %d, blocks/SM %s; per-SM finish us: min %.1f p50 %.1f p90 %.1f max %.1f; slowest blocks %s (SM %s)"
          % (len(per_sm), sorted(set(len(v) for v in per_sm.values())), sm_end.min(), np.percentile(sm_end, 50),
             np.percentile(sm_end, 90), sm_end.max(), list(order[-4:]), [int(sm[b]) for b in order[-4:]]))
