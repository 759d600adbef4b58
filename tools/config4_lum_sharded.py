#!/usr/bin/env python
"""configs[3] global relaxation with the graph links sharded over GPUs: every rank holds all 65 scans,
fills G / B from ITS links, ONE NCCL all-reduce of [G|B] per LUM iteration, replicated solve.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/config4_lum_sharded.py
(or plain `python tools/config4_lum_sharded.py` for N = 1).  Rank 0 prints one JSON line; time = max over ranks."""
import importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
icp = importlib.import_module("3dtk_b200")
par = importlib.import_module("3dtk_b200.parallel")
ctx = icp.Context(local)
n_scans, n_pts = int(os.environ.get("N_SCANS", 65)), int(os.environ.get("N_PTS", 300_000))
rng = np.random.default_rng(4)
dev, T = [], []
for i in range(n_scans):          # registered sequence with a small residual error per scan (what ICP leaves behind)
    P = icp.euler_to_matrix4(rng.normal(0, 0.5, 3), np.deg2rad(rng.normal(0, 0.05, 3))) if i else np.eye(4).reshape(16)
    s = icp.Scan(ctx, icp.transform_points(P, icp.synth_scene(7, 1400 + i, n_pts, 0.5)), max_dist_hint=25.0)
    s.set_pose(P, None)
    dev.append(s); T.append(P)
rpos = np.array([icp.matrix4_to_euler(t)[0] for t in T])
graph = icp.Graph.from_poses(rpos, 750.0 ** 2, 20)
lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0, epsilon_lum=-1.0)
device = torch.device("cuda", local)
par.graph_slam_sharded(lum, graph, dev, 1, rank, world, device=device)      # warm-up iteration
ctx.synchronize()
if world > 1:
    dist.barrier()
iters = 3
t0 = time.perf_counter()
ret, it = par.graph_slam_sharded(lum, graph, dev, iters, rank, world, device=device)
ctx.synchronize()
el = time.perf_counter() - t0
tmax, links = par.reduce_timing(el, len(par.shard_units(graph.get_nr_links(), rank, world)) * iters, device=device)
poses = np.array([d.get_pose()[0] for d in dev])
chk = torch.tensor(poses.reshape(-1), dtype=torch.float64, device=device)
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
else:
    identical = True
if rank == 0:
    print(json.dumps({"workload": "configs[3] LUM: %d scans x %d points, %d links, -D 25" % (n_scans, n_pts, graph.get_nr_links()),
                      "n_gpus": world, "lum_iterations": iters, "s_per_lum_iteration": tmax / iters,
                      "link_evaluations_per_s": links / tmax, "ret": ret,
                      "poses_bit_identical_across_ranks": identical,
                      "allreduce_bytes_per_iteration": 8 * ((6 * (n_scans - 1)) ** 2 + 6 * (n_scans - 1))}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
ctx.close()
