#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] at FULL size on one GPU (parity at these sizes is checked through
properties; the scaled pipelines are compared with the oracle in tests/test_configs.py):
  config 4  hannover1-shaped: 65 scans x 300k points, odometry drift, `-i 100 -d 75 --metascan`, then
            Graph(cldist 750, loopsize 20) + `-G 1 -I 50 -D 250` (here I = 10 iterations, epsSLAM 0.5)
  config 5  bremen_city-shaped: 13 scans x 10M points, `-r 10` octree reduction, `-i 0`, LUM over a given graph,
            `-D 100 -I 50` (here I = 10)
Prints one JSON object per config: wall times of the stages and the properties checked.
    python tools/config_scale.py [4] [5] > profiles/rNN_configs_full_scale.json"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
icp = importlib.import_module("3dtk_b200")
LOCAL = int(os.environ.get("LOCAL_RANK", 0))
ctx = icp.Context(LOCAL)     # one process per GPU under torchrun
which = [int(a) for a in sys.argv[1:]] or [4, 5]


def drift_sequence(n_scans, n_pts, seed, step_cm, step_deg, geom=7):
    rng = np.random.default_rng(seed)
    E = np.eye(4)
    scans, org = [], []
    for i in range(n_scans):
        if i:
            E = E @ icp.euler_to_matrix4(rng.normal(0, step_cm, 3), np.deg2rad(rng.normal(0, step_deg, 3))).reshape(4, 4).T
        M = E.T.reshape(16).copy()
        scans.append(icp.transform_points(M, icp.synth_scene(geom, 1000 + 100 * seed + i, n_pts, 0.5)))
        org.append(M)
    return scans, np.array(org)


def pose_err(dev):
    """all scans sample one scene at the identity: the final transMat IS the remaining pose error"""
    t = np.array([np.linalg.norm(d.get_pose()[0][12:15]) for d in dev])
    return float(t.max()), float(t.mean())


if 4 in which:
    n_scans, n_pts = 65, 300_000
    t0 = time.perf_counter(); scans, org = drift_sequence(n_scans, n_pts, 4, 1.5, 0.12); t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    dev = [icp.Scan(ctx, s, max_dist_hint=75.0) for s in scans]
    for d, t in zip(dev, org):
        d.set_pose(t, None)
    ctx.synchronize(); t_up = time.perf_counter() - t0
    drift_max = float(max(np.linalg.norm(t[12:15]) for t in org))
    eng = icp.icp6D(ctx, algo=1, max_dist_match=75.0, max_num_iterations=100, epsilon_icp=1e-5)
    frames = icp.Frames(n_scans)
    t0 = time.perf_counter(); its = eng.doICP(dev, extrapolate_pose=True, meta=True, transmat_org=org, frames=frames)
    ctx.synchronize(); t_icp = time.perf_counter() - t0
    e_icp = pose_err(dev)
    rpos = np.array([icp.matrix4_to_euler(d.get_pose()[0])[0] for d in dev])
    graph = icp.Graph.from_poses(rpos, 750.0 ** 2, 20)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=250.0 ** 0.5 * 250.0 ** 0.5, epsilon_lum=0.5)
    t0 = time.perf_counter(); ret, it = lum.doGraphSlam6D(graph, dev, 10, frames=frames); ctx.synchronize()
    t_lum = time.perf_counter() - t0
    e_lum = pose_err(dev)
    t0 = time.perf_counter(); ms = icp.Scan.metascan(ctx, dev[:64], max_dist_hint=75.0); ctx.synchronize()
    t_meta = time.perf_counter() - t0
    ms.destroy()
    print(json.dumps({"config": "configs[3] hannover1-shaped: 65 x 300k, sequential ICP + metascan, then LUM", "n_scans": n_scans,
                      "points_per_scan": n_pts, "generate_s": t_gen, "upload_and_grids_s": t_up,
                      "doICP_metascan_s": t_icp, "icp_iterations_total": int(np.sum(its)), "icp_iterations_max": int(np.max(its)),
                      "metascan_points_last": int((n_scans - 1) * n_pts), "metascan_rebuild_64_members_s": t_meta, "graph_links": int(graph.get_nr_links()),
                      "lum_s": t_lum, "lum_iterations": it, "lum_ret": ret,
                      "odometry_drift_max_cm": drift_max, "pose_error_after_icp_cm(max,mean)": e_icp,
                      "pose_error_after_lum_cm(max,mean)": e_lum,
                      "frames_per_scan": int(len(frames.get(0))),
                      "properties": {"drift_removed": e_icp[0] < 0.1 * drift_max, "lum_keeps_registration": e_lum[0] < 2.0 * max(e_icp[0], 1.0),
                                     "loop_closures_found": graph.get_nr_links() > n_scans - 1}}), flush=True)
    del dev, scans

if 5 in which:
    # bremen_city is a city block, not a room: the scene generator's room is scaled by CONFIG5_SCALE (default 5.7:
    # 114 x 17 x 57 m, ~2.3e4 m^2 of surfaces), so that `-r 10` takes 10 M raw points to ~1 M (the config's figure).
    # Under torchrun the scans are sharded one-per-GPU for generation + reduction (scan i on rank i mod N), the
    # reduced clouds are exchanged, and the relaxation runs link-sharded with one all-reduce of [G|B] per iteration.
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch, torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    par = importlib.import_module("3dtk_b200.parallel")
    import torch                                   # (graph_slam_sharded imports it: keep the import out of the timed region)
    import torch.distributed  # noqa: F401
    n_scans, n_raw = 13, int(os.environ.get("CONFIG5_RAW", 10_000_000))
    scale = float(os.environ.get("CONFIG5_SCALE", "5.7"))
    voxel = float(os.environ.get("CONFIG5_VOXEL", "10.0"))
    nr_it = int(os.environ.get("CONFIG5_LUM_ITERS", "50"))
    rng = np.random.default_rng(5)
    E, org = np.eye(4), []
    for i in range(n_scans):
        if i:
            E = E @ icp.euler_to_matrix4(rng.normal(0, 1.5, 3), np.deg2rad(rng.normal(0, 0.05, 3))).reshape(4, 4).T
        org.append(E.T.reshape(16).copy())
    org = np.array(org)
    t_gen = t_red = 0.0
    reduced = [None] * n_scans
    for i in range(rank, n_scans, world):
        t0 = time.perf_counter()
        raw = icp.transform_points(org[i], icp.synth_scene(7, 1500 + i, n_raw, 0.5 / scale) * scale)
        t_gen += time.perf_counter() - t0
        t0 = time.perf_counter(); reduced[i] = icp.reduce_octree_center(ctx, raw, voxel); ctx.synchronize()
        t_red += time.perf_counter() - t0
        del raw
    if dist is not None:      # exchange the reduced clouds (scan i lives on rank i mod N)
        import torch
        for i in range(n_scans):
            src = i % world
            n_i = torch.tensor([len(reduced[i]) if rank == src else 0], device="cuda:%d" % LOCAL)
            dist.broadcast(n_i, src)
            buf = torch.from_numpy(reduced[i]).to("cuda:%d" % LOCAL) if rank == src else torch.empty((int(n_i[0]), 3), dtype=torch.float64, device="cuda:%d" % LOCAL)
            dist.broadcast(buf, src)
            reduced[i] = buf.cpu().numpy()
        tt = torch.tensor([t_gen, t_red], dtype=torch.float64, device="cuda:%d" % LOCAL)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_gen, t_red = float(tt[0]), float(tt[1])
    t0 = time.perf_counter()
    dev = [icp.Scan(ctx, r, max_dist_hint=100.0) for r in reduced]
    for d, t in zip(dev, org):
        d.set_pose(t, None)
    ctx.synchronize(); t_up = time.perf_counter() - t0
    links = np.array([[i, i + 1] for i in range(n_scans - 1)] + [[0, n_scans - 1]] + [[i, i + 2] for i in range(n_scans - 2)], dtype=np.int32)
    e0 = pose_err(dev)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=100.0, epsilon_lum=0.0)
    device = None
    if dist is not None:
        import torch
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        dist.barrier()
    t0 = time.perf_counter()
    ret, it = par.graph_slam_sharded(lum, icp.Graph(links, n_scans), dev, nr_it, rank, world, device=device)
    ctx.synchronize()
    t_lum = time.perf_counter() - t0
    if dist is not None:
        import torch
        tt = torch.tensor([t_lum], dtype=torch.float64, device="cuda:%d" % LOCAL); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t_lum = float(tt[0])
    e1 = pose_err(dev)
    if rank == 0:
        print(json.dumps({"config": "configs[4] bremen_city-shaped: 13 x %d raw points on a %.0fx-scaled scene, -r %g, -i 0, "
                                    "LUM over a given graph (-G 1 -I %d -D 100)" % (n_raw, scale, voxel, nr_it),
                          "n_gpus": world, "n_scans": n_scans, "raw_points_per_scan": n_raw, "voxel": voxel,
                          "reduced_points(min,max)": [int(min(map(len, reduced))), int(max(map(len, reduced)))],
                          "generate_s(max over ranks)": t_gen, "octree_reduction_s(max over ranks, incl. host->device copies)": t_red,
                          "upload_and_grids_s": t_up, "graph_links": int(len(links)),
                          "lum_s": t_lum, "lum_iterations": it, "s_per_lum_iteration": t_lum / max(it, 1), "lum_ret": ret,
                          "pose_error_before_cm(max,mean)": e0, "pose_error_after_lum_cm(max,mean)": e1,
                          "properties": {"lum_reduces_error": e1[1] < 0.5 * e0[1]}}), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
ctx.close()
