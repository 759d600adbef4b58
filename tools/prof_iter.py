#!/usr/bin/env python
"""Per-iteration profile of one fused match on the bench workload (used alone and under ncu).
Prints one JSON line: per-iteration correspondence-kernel ms, stage-2 (ring search) query counts, grid."""
import argparse, importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--ppc", type=float, default=0.0)
ap.add_argument("--cell-edge", type=float, default=0.0)
ap.add_argument("--exact", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=50)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--algo", type=int, default=1)
a = ap.parse_args()
if a.ppc > 0:
    os.environ["B200ICP_TARGET_PPC"] = str(a.ppc)
icp = importlib.import_module("3dtk_b200")
ctx = icp.Context(0)
n = a.points
model = icp.synth_scene(7, 42, n, 0.5)
data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
m = icp.Scan(ctx, model, cell_edge=a.cell_edge, max_dist_hint=25.0)
d = icp.Scan(ctx, data, cell_edge=a.cell_edge, max_dist_hint=25.0)
eng = icp.icp6D(ctx, algo=a.algo, max_dist_match=25.0, max_num_iterations=a.max_iter, epsilon_icp=1e-5,
                exact=bool(a.exact), profile=True)
ident = np.eye(4).reshape(16)
for _ in range(a.repeat):
    d.set_pose(ident, ident)
    r = eng.match(m, d)
p = r["profile"]
print(json.dumps({"points": n, "exact": a.exact, "grid": m.grid_info(), "iters": r["iterations_run"],
                  "nn_ms": [round(x, 4) for x in p["nn_ms"]], "stage2": [int(x) for x in p["stage2"]], "searches": [int(x) for x in p["searches"]],
                  "stream_ms": [round(x, 4) for x in p["solve_ms"]], "nn_ms_mean": float(np.mean(p["nn_ms"])), "solve_ms_mean": float(np.mean(p["solve_ms"])),
                  "rms_last": float(r["rms"][-1])}))
