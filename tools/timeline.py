#!/usr/bin/env python
"""In-kernel timeline of the last ICP iteration (needs a -DB200_TIMING build: tools/build_variant.sh timing -DB200_TIMING)."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200ICP_LIB"] = os.path.join(ROOT, "3dtk_b200", "lib", "timing.so")
icp = importlib.import_module("3dtk_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = icp.Context(0)
model = icp.synth_scene(7, 42, n, 0.5); data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
m = icp.Scan(ctx, model, max_dist_hint=25.0); d = icp.Scan(ctx, data, max_dist_hint=25.0)
eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=iters, epsilon_icp=1e-5)
for _ in range(2):
    d.set_pose(np.eye(4).reshape(16), np.eye(4).reshape(16)); r = eng.match(m, d)
t = (C.c_ulonglong * 32)()
icp.lib.b200icp_debug_timing.argtypes = [C.c_void_p]
icp.lib.b200icp_debug_timing(t)
names = {16: "stream: block0 start", 17: "stream: after pdl_wait", 18: "stream: block0 tiles done", 19: "stream: block0 reduced",
         0: "search: block0 start", 1: "search: after pdl_wait", 2: "search: prologue done", 3: "search: block0 batches done",
         4: "search: block0 partials stored", 7: "last block: ticket won", 8: "solve: start", 9: "solve: state staged",
         10: "solve: moments reduced", 20: "serial: start", 21: "serial: counters exchanged", 22: "serial: solve_any done",
         23: "serial: logs written", 24: "serial: poses composed", 11: "solve: serial solve done", 12: "solve: state written"}
order = [8, 9, 10, 20, 21, 22, 23, 24, 11, 12]
t0 = t[8]
print("n=%d iterations_run=%d  (ns since the stream kernel's block 0 started, last iteration)" % (n, r["iterations_run"]))
for k in order:
    print("%8d  %s" % (t[k] - t0, names[k]))
