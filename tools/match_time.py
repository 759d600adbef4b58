#!/usr/bin/env python
"""Whole-match device time of the bench pair without per-launch events (what bench.py's `value` times).
usage: [B200ICP_LIB=...] [env switches] python tools/match_time.py [repeats]"""
import importlib, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
icp = importlib.import_module("3dtk_b200")
n = int(os.environ.get("N", 1_000_000))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = icp.Context(0, stream=torch.cuda.current_stream().cuda_stream)
model = icp.synth_scene(7, 42, n, 0.5)
data = icp.synth_scene(7, 43, n, 0.5)
P = icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8]))
data = icp.transform_points(icp.m4inv(P)[0], data)
m = icp.Scan(ctx, model, max_dist_hint=25.0); d = icp.Scan(ctx, data, max_dist_hint=25.0)
eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5)
ident = np.eye(4).reshape(16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for r in range(reps + 2):
    d.set_pose(ident, ident)
    flush.fill_(r & 255)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    res = eng.match(m, d)
    b.record()
    torch.cuda.synchronize()
    if r >= 2:
        ts.append(a.elapsed_time(b))
print(json.dumps({"ms_per_match_mean": float(np.mean(ts)), "min": float(np.min(ts)), "max": float(np.max(ts)),
                  "iterations": int(res["iterations_run"]), "rms_last": float(res["rms"][-1])}))
