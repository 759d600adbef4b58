#!/usr/bin/env python
"""Counters of the cooperative tile search over one match of the bench pair (needs a -DB200_TILE_STATS build:
tools/build_variant.sh NAME -DB200_TILE_STATS ...; B200ICP_LIB=.../NAME.so python tools/tile_stats.py [max_iter ...])."""
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
prev = np.zeros(3)
for iters in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 5, 10, 15, 20, 25, 30, 40, 50]:
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=iters, epsilon_icp=1e-5, profile=True)
    d.set_pose(np.eye(4).reshape(16), np.eye(4).reshape(16))
    icp.lib.b200icp_debug_tile_stats(t)
    r = eng.match(m, d)
    icp.lib.b200icp_debug_tile_stats(t)
    cur = np.array([t[0], t[1], t[2]], dtype=float)
    dl = cur - prev
    s = int(np.sum(r["profile"]["searches"]))
    print("max_iter %2d: searches %9d (batches %8d)  tiled %8d  refused %7d  mean staged %.1f | since previous row: tiled %8d refused %7d mean staged %.1f"
          % (iters, s, s // 32, t[0], t[2], t[1] / max(t[0], 1), dl[0], dl[2], dl[1] / max(dl[0], 1)))
    prev = cur
