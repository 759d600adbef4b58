#!/usr/bin/env python
"""Where the end-to-end step of bench.py goes: upload + grid build per scan, match, read-back (wall clock,
pinned host buffers), plus the per-iteration kernel times of a match that starts at the converged pose
(iterations >= 2 of it are skip-only: the floor of one iteration)."""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
icp = importlib.import_module("3dtk_b200")
n = int(os.environ.get("N", 1_000_000))
ctx = icp.Context(0, stream=torch.cuda.current_stream().cuda_stream) if os.environ.get("STREAM") == "torch" else icp.Context(0)
model = icp.synth_scene(7, 42, n, 0.5)
data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
hm = torch.from_numpy(model).pin_memory(); hd = torch.from_numpy(data).pin_memory()
eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5)
out = {"scan_create_model_ms": [], "scan_create_data_ms": [], "match_ms": [], "get_pose_ms": [], "total_ms": []}
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = icp.Scan.from_host_pointers(ctx, hm.data_ptr(), None, n, 0.0, 25.0)
    t1 = time.perf_counter()
    ds = icp.Scan.from_host_pointers(ctx, hd.data_ptr(), None, n, 0.0, 25.0)
    t2 = time.perf_counter()
    r = eng.match(ms, ds)
    t3 = time.perf_counter()
    pose = ds.get_pose()[0]
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    if rep:
        for k, v in zip(out, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
            out[k].append(round(1e3 * v, 3))
    if rep < 5:
        ms.destroy(); ds.destroy()
# raw H2D of one scan
dev = torch.empty(n * 3, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): dev.copy_(hm.view(-1), non_blocking=True)
torch.cuda.synchronize(); out["h2d_24MB_ms"] = round(1e3 * (time.perf_counter() - t0) / 5, 3)
# match from the converged pose: skip-only iterations
engp = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=12, epsilon_icp=0.0, profile=True)
T, D = ds.get_pose()
r = engp.match(ms, ds)
out["converged_start_nn_ms"] = [round(x, 4) for x in r["profile"]["nn_ms"]]
out["converged_start_searches"] = [int(x) for x in r["profile"]["searches"]]
print(json.dumps(out))
