#!/usr/bin/env python
"""configs[2] end to end on one GPU: 1M-point pair, point-to-plane NAPX, normals by k-NN PCA (k = 10) on the GPU.
Wall clock from pinned host arrays to the final pose, two ways of getting the normals:
  (a) host round trip: b200icp_normals_knn(host xyz) -> host normals -> scan_create(xyz, normals)
  (b) on device:       scan_create(xyz) -> b200icp_scan_calc_normals(scan)
Prints one JSON line."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
icp = importlib.import_module("3dtk_b200")
ctx = icp.Context(0)
n = 1_000_000
model = icp.synth_scene(7, 42, n, 0.5); data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
hm, hd = torch.from_numpy(model).pin_memory(), torch.from_numpy(data).pin_memory()
rpos = np.array([0.0, 150.0, 0.0])
eng = icp.icp6D(ctx, algo=icp.ALGO_NAPX, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5, napx_weighted=True)


def run(on_device):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = icp.Scan.from_host_pointers(ctx, hm.data_ptr(), None, n, 0.0, 25.0)
    if on_device:
        d = icp.Scan.from_host_pointers(ctx, hd.data_ptr(), None, n, 0.0, 25.0)
        d.calc_normals(10, rpos)
    else:
        nrm = icp.normals_knn(ctx, hd.numpy(), 10, rpos)
        d = icp.Scan(ctx, hd.numpy(), normals=nrm, max_dist_hint=25.0)
    t1 = time.perf_counter()
    r = eng.match(m, d, icp.CLOSEST_PLANE_SIMPLE)
    T = d.get_pose()[0]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    m.destroy(); d.destroy()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), r["iterations_run"], T


out = {}
for name, flag in (("host_round_trip", False), ("on_device", True)):
    run(flag)
    best = min((run(flag) for _ in range(3)), key=lambda x: x[0] + x[1])
    out[name] = {"staging_and_normals_ms": best[0], "match_ms": best[1], "iterations": best[2],
                 "total_ms": best[0] + best[1], "pose_rel_frobenius_vs_truth": float(np.linalg.norm(best[3] - P) / np.linalg.norm(P))}
print(json.dumps({"workload": "configs[2]: 1M x 1M, NAPX (least-squares form), CLOSEST_PLANE_SIMPLE, k=10 normals on GPU", **out}))
ctx.close()
