#!/usr/bin/env python
"""How many searches end fp32-ambiguous ("unsettled": exact rescan, no motion budget) or with a tiny budget, per
iteration range of the bench pair (needs tools/mk.sh NAME -DB200_COUNT_UNSETTLED; B200ICP_LIB=.../NAME.so)."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
icp = importlib.import_module("3dtk_b200")
n = 1_000_000
ctx = icp.Context(0)
model = icp.synth_scene(7, 42, n, 0.5); data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
m = icp.Scan(ctx, model, max_dist_hint=25.0); d = icp.Scan(ctx, data, max_dist_hint=25.0)
icp.lib.b200icp_debug_tile_stats.argtypes = [C.c_void_p]
t = (C.c_ulonglong * 8)()
prev = np.zeros(5)
for iters in [int(x) for x in sys.argv[1:]] or [1, 2, 5, 10, 15, 20, 25, 30, 35, 40, 45, 48]:
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=iters, epsilon_icp=1e-5, profile=True)
    d.set_pose(np.eye(4).reshape(16), np.eye(4).reshape(16))
    icp.lib.b200icp_debug_tile_stats(t)
    r = eng.match(m, d)
    icp.lib.b200icp_debug_tile_stats(t)
    s = float(np.sum(r["profile"]["searches"]))
    cur = np.array([t[0], t[1], t[2], t[3], s])
    dl = cur - prev
    print("iterations %2d..%2d: searches %9d | unsettled stencil %7d ball %7d | settled with budget < 2e-3: %8d, < 2e-4: %7d"
          % (0 if not prev[4] else len(r["profile"]["searches"]) - 0, iters, dl[4], dl[0], dl[1], dl[2], dl[3]))
    prev = cur
