#!/usr/bin/env python
"""Per-warp timing of the fused iteration kernel's LAST launch (needs tools/mk.sh timing -DB200_TIMING).
usage: warp_times.py [max_iter ...]"""
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
icp.lib.b200icp_debug_warps.argtypes = [C.c_void_p, C.c_int]
NB = 444
for iters in [int(x) for x in sys.argv[1:]] or [45]:
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=iters, epsilon_icp=1e-5)
    for _ in range(2):
        d.set_pose(np.eye(4).reshape(16), np.eye(4).reshape(16)); r = eng.match(m, d)
    buf = (C.c_ulonglong * (NB * 4))(); icp.lib.b200icp_debug_blocks(buf, NB)
    a = np.array(buf, dtype=np.uint64).reshape(NB, 4).astype(np.int64)
    wb = (C.c_ulonglong * (NB * 48))(); icp.lib.b200icp_debug_warps(wb, NB)
    w = np.array(wb, dtype=np.uint64).reshape(NB, 8, 6).astype(np.int64)
    t0 = a[:, 0].min()
    us = lambda x: (x - t0) / 1e3
    ws, we, ls, le = us(w[:, :, 0]), us(w[:, :, 1]), us(w[:, :, 2]), us(w[:, :, 3])
    pc = lambda x, q: np.percentile(x, q)
    walk = we - ws
    left = le - ls
    busy = left[left > 0.2]
    print("launches %d: walk start p50 %.1f | per-warp walk us p10 %.1f p50 %.1f p90 %.1f max %.1f | block barrier passed p50 %.1f | leftover batches: %d of %d warps busy, their time p10 %.1f p50 %.1f p90 %.1f max %.1f | search batches per warp mean %.2f"
          % (iters, pc(ws, 50), pc(walk, 10), pc(walk, 50), pc(walk, 90), walk.max(), pc(ls, 50), busy.size, left.size,
             pc(busy, 10) if busy.size else 0, pc(busy, 50) if busy.size else 0, pc(busy, 90) if busy.size else 0, busy.max() if busy.size else 0, w[:, :, 4].mean()))
    print("   block: loop done p50 %.1f, partials stored p50 %.1f max %.1f" % (pc(a[:, 1] - t0, 50) / 1e3, pc(a[:, 2] - t0, 50) / 1e3, (a[:, 2] - t0).max() / 1e3))
