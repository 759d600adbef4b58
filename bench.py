#!/usr/bin/env python
"""bench.py -- ICP correspondences/sec (and iterations/sec) on a 1M-point synthetic scan pair.

Workload (BASELINE.json configs[1]): synthetic 2-scan pair, 1M points each, known SE(3) offset,
point-to-point icp6D_QUAT, max_dist 25 cm, <= 50 iterations, epsICP 1e-5, no subsampling.
One "step" = one full icp6D::match of the pair (all iterations until the reference's convergence test
fires).  Metric = data points searched per second = N_d * iterations / time.

  value      scans + grids already resident in HBM when the timed region starts (CUDA events on the
             stream the kernels run on, one event pair per step, L2 flushed between steps)
  e2e        same metric through the public C ABI with HOST (pinned) buffers: upload + grid build of both
             scans, match, pose read-back, all inside the timed region
  roofline   correspondence kernel: algorithmic bytes / mean launch time (CUDA events inside the library)
  cpu_baseline / --impl reference
             the reference's own CPU path (oracle/_ref: unmodified 3DTK kd.cc / searchTree.cc /
             icp6Dquat.cc, OpenMP pICP arm, all host cores) on the SAME pair and the SAME match (all iterations
             to the reference's convergence test); the sample is bounded by the number of matches (<= 2), not
             by the iterations.  The reference arm never loads the product library: its inputs come from
             oracle/_build/libscenegen.so (same generator source, bit-identical arrays).
  parity     (GPU line) the final pose and the pair count of every iteration against a live run of the compiled
             reference on the identical arrays (serial-arm arithmetic, neighbour search spread over the host
             cores: oracle/ref_harness.cc) -- the BASELINE.md section 3 gate (< 1e-4 relative Frobenius).

Multi-GPU (--gpus N under torchrun): the path shards by scan pair -- every rank matches its own pair,
no data-path collective ("scaling": "weak"); time = max over ranks.
"""
import argparse
import ctypes as C
import importlib
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "icp_correspondences_per_sec"
UNIT = "correspondences/s"
POSE_POS = (12.0, -7.0, 5.0)
POSE_THETA_DEG = (0.5, -1.0, 0.8)
FP32_PEAK_FLOPS = 148 * 128 * 2 * 1.965e9   # B200 non-tensor fp32 (B200_PROFILING.md: 148 SMs, 1965 MHz)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--exact", type=int, default=1)
    ap.add_argument("--max-iter", type=int, default=50)
    ap.add_argument("--ref-matches", type=int, default=2, help="upper bound of timed matches of the reference arm")
    ap.add_argument("--no-parity", action="store_true", help="skip the live parity check against oracle/_ref")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cell-edge", type=float, default=0.0)
    ap.add_argument("--shard", default="pairs", choices=["pairs", "queries"],
                    help="pairs: one scan pair per GPU (weak scaling, default, north_star); queries: ONE pair, data\n"
                         "points split across GPUs, moments all-reduced inside the kernel over NVLink (strong scaling)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the multi-GPU objects of the line (strong: query-sharded pair; lum_link_sharded: configs[3])")
    ap.add_argument("--rows", action="store_true",
                    help="instead of the configs[1] line: measure the SURVEY 8f rows (octree reduction, normals, LUM link,\n"
                         "graph relaxation, metascan, uos reader), one JSON object per row (tests/rows_bench.py)")
    return ap.parse_args()


def make_pair(icp, n, rank):
    model = icp.synth_scene(7, 42 + 2 * rank, n, 0.5)
    data = icp.synth_scene(7, 43 + 2 * rank, n, 0.5)
    Pm = icp.euler_to_matrix4(np.array(POSE_POS), np.deg2rad(np.array(POSE_THETA_DEG)))
    Pinv, _ = icp.m4inv(Pm)
    return model, icp.transform_points(Pinv, data), Pm


def make_pair_standalone(n, rank):
    """The same pair (bit-identical arrays) from oracle/_build/libscenegen.so -- for the reference arm, which must
    not load the product library."""
    so = os.path.join(ROOT, "oracle", "_build", "libscenegen.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    L = C.CDLL(so)
    L.scenegen_pair.restype = C.c_int
    L.scenegen_pair.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, C.c_double] + [C.c_void_p] * 5
    model, data, P = np.empty((n, 3)), np.empty((n, 3)), np.empty(16)
    pos, th = np.array(POSE_POS), np.array(POSE_THETA_DEG)
    rc = L.scenegen_pair(7, 42 + 2 * rank, 43 + 2 * rank, n, 0.5, pos.ctypes.data, th.ctypes.data,
                         model.ctypes.data, data.ctypes.data, P.ctypes.data)
    if rc != 0:
        raise RuntimeError("scenegen_pair failed: %d" % rc)
    return model, data, P


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).

    Sampled in-process through NVML (pynvml) from a thread, every 20 ms: an `nvidia-smi -lms 50` child polling
    its full query set was measured to DOUBLE the wall-clock time of the end-to-end steps (20.1 ms vs 10.1 ms per
    step) by contending for the driver, so nvidia-smi is only the fallback (at a 500 ms period)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, uuid=None):
        self.rows, self.proc, self.idx, self.uuid = [], None, gpu_index, uuid
        self.sm, self.mx, self.reasons, self.how = [], [], set(), None
        self._stop = threading.Event()
        self.th = None

    def _nvml_loop(self, nv, h):
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for nm, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.how = "nvml"
            self.th = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.th.start()
            return
        except Exception:
            self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.how == "nvml":
            self._stop.set()
            self.th.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml, 20 ms period, sampled during the timed regions"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 500"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """DRAM bytes and candidate evaluations per launch of the correspondence kernel from the committed
    `ncu --set full` captures of THIS tree (profiles/ncu_traffic.json, written by tools/gpu_roofline.sh +
    tools/ncu_traffic.py); (None, None) when absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_per_launch_mean"]), d
    except Exception:
        return None, None


def host_threads():
    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1); the harness passes
    # the count in an explicit num_threads clause
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        threads = os.cpu_count() or 1
    return max(1, min(threads, 256))


def cpu_reference_run(model, data, max_iter, matches, serial_semantics=False):
    """`matches` full matches of the compiled reference (oracle/_ref) on all host cores.
    serial_semantics=False: the reference's OpenMP pICP arm (icp6D.cc:129-222: getPtPairsParallel + Align_Parallel).
    serial_semantics=True : the serial arm's arithmetic (icp6D.cc:224-244) with the k-d tree searches spread over
                            the cores -- bit-identical to the serial reference, used as the parity oracle."""
    import orclib
    from orclib import P
    L = orclib.ref(omp=True)
    if L is None:
        return None
    threads = host_threads()
    m = np.ascontiguousarray(model)
    t0 = time.perf_counter()
    tree = L.ref_tree_create(P(m), len(m), 0, 20)
    build_s = time.perf_counter() - t0
    times, iters, last = [], [], None
    for s_ in range(matches):
        d = np.ascontiguousarray(data).copy()
        T, D, S = orclib.identity(), orclib.identity(), orclib.identity()
        rms = np.zeros(max_iter); npairs = np.zeros(max_iter, dtype=np.int64)
        done, ms = C.c_int(0), C.c_double(0)
        t0 = time.perf_counter()
        ret = L.ref_match(tree, P(S), P(d), None, len(d), P(T), P(D), 1, 0, 25.0, max_iter, 1e-5, 1,
                          -threads if serial_semantics else threads,
                          P(rms), P(npairs), C.byref(done), C.byref(ms))
        dt = time.perf_counter() - t0
        times.append(dt); iters.append(done.value)
        last = {"transmat": T.copy(), "npairs": npairs[:done.value].copy(), "rms": rms[:done.value].copy(),
                "iterations": ret, "iterations_run": done.value}
    L.ref_tree_free(tree)
    total_t, total_it = sum(times), sum(iters)
    arm = ("serial-arm arithmetic (Scan::getPtPairs + icp6D_QUAT::Align), k-d tree searches on %d threads" % threads
           if serial_semantics else "icp6D_QUAT OpenMP pICP arm (getPtPairsParallel + Align_Parallel)")
    return {"value": len(data) * total_it / total_t, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "full %d x %d pair, %s, %d full match(es) of %d iterations each (all iterations to the "
                      "reference's convergence test); k-d tree build (%.2f s) excluded as the reference does "
                      "(icp6D.cc:127)" % (len(model), len(data), arm, matches, iters[-1], build_s),
            "iters_per_sec": total_it / total_t, "ms_per_step": 1e3 * total_t / max(len(times), 1),
            "iterations": total_it, "matches": matches, "last": last}


def strong_section(icp, torch, dist, local_rank, rank, world, a, t1_ms):
    """SURVEY 8e-A under the driver: ONE 1M pair, the data scan split across the ranks (the split of
    Scan::getPtPairsParallel, scan.cc:1335-1342), pair moments all-reduced INSIDE the iteration kernel over NVLink
    peer memory.  Returns the `strong` object of the JSON line (rank 0) or None."""
    stream = torch.cuda.current_stream()
    ctx = icp.Context(local_rank, stream=stream.cuda_stream)
    model, data, _ = make_pair(icp, a.points, 0)
    step_q = -(-len(data) // world)
    mine = np.ascontiguousarray(data[rank * step_q:min((rank + 1) * step_q, len(data))])
    handle = ctx.comm_create(rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ctx.comm_connect_ipc(handles)
    dist.barrier()
    m_scan = icp.Scan(ctx, model, max_dist_hint=25.0)
    d_scan = icp.Scan(ctx, mine, max_dist_hint=25.0)
    eng = icp.icp6D(ctx, algo=icp.ALGO_QUAT, max_dist_match=25.0, max_num_iterations=a.max_iter, epsilon_icp=1e-5,
                    exact=bool(a.exact), sharded=True)
    ident = np.eye(4).reshape(16).copy()
    steps = max(3, min(a.steps, 10))
    tot, iters, last = 0.0, 0, None
    for k in range(3 + steps):
        d_scan.set_pose(ident, ident)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        last = eng.match(m_scan, d_scan)
        e1.record(stream); e1.synchronize()
        if k >= 3:
            tot += e0.elapsed_time(e1); iters += last["iterations_run"]
    v = torch.tensor([tot], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    T = d_scan.get_pose()[0]
    chk = torch.tensor(T, dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ms = float(v[0]) / steps
    it = iters / steps
    out = {"workload": "ONE %d x %d pair, data points split over %d GPUs (scan.cc:1335-1342), pair moments "
                       "all-reduced in the iteration kernel through NVLink peer mailboxes (no NCCL on the path)"
                       % (a.points, a.points, world),
           "ms_per_match": ms, "iterations_per_match": it, "n1_ms_per_match": t1_ms,
           "speedup_vs_n1": t1_ms / ms, "efficiency": t1_ms / (world * ms),
           "value": a.points * it / (ms * 1e-3), "unit": UNIT, "steps": steps,
           "pose_bit_identical_across_ranks": bool(torch.equal(lo, hi)),
           "fixed_us_per_iteration": 1e3 * (ms - t1_ms / world) / max(it, 1),
           "limiter": "per-iteration fixed cost (kernel launch + last-block reduction + serial 6-DoF solve + mailbox "
                      "flag round trip over NVLink) does not shrink with N: fixed_us_per_iteration = (t_N - t_1 / N) "
                      "/ iterations"}
    m_scan.destroy(); d_scan.destroy()
    dist.barrier()
    ctx.close()
    return out if rank == 0 else None


def lum_section(icp, torch, dist, local_rank, rank, world):
    """SURVEY 8e-B exchange step under the driver: BASELINE configs[3]'s global relaxation (65 scans x 300k points,
    replicated per GPU), graph links sharded round-robin, ONE all-reduce of the packed [G|B] per LUM iteration
    (b200icp_lum_graph_slam_sharded; NCCL through torch.distributed).  Seeding of the link searches from the previous
    LUM iteration is on (default).  Returns the `lum_link_sharded` object (rank 0) or None."""
    par = importlib.import_module("3dtk_b200.parallel")
    n_scans, n_pts = int(os.environ.get("B200_BENCH_LUM_SCANS", 65)), int(os.environ.get("B200_BENCH_LUM_PTS", 300_000))
    ctx = icp.Context(local_rank)
    rng = np.random.default_rng(4)
    dev, T = [], []
    for i in range(n_scans):      # registered sequence with a small residual error per scan (what ICP leaves behind)
        Pm = icp.euler_to_matrix4(rng.normal(0, 0.5, 3), np.deg2rad(rng.normal(0, 0.05, 3))) if i else np.eye(4).reshape(16)
        sc = icp.Scan(ctx, icp.transform_points(Pm, icp.synth_scene(7, 1400 + i, n_pts, 0.5)), max_dist_hint=25.0)
        sc.set_pose(Pm, None)
        dev.append(sc); T.append(Pm)
    rpos = np.array([icp.matrix4_to_euler(t)[0] for t in T])
    graph = icp.Graph.from_poses(rpos, 750.0 ** 2, 20)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0, epsilon_lum=-1.0)
    device = torch.device("cuda", local_rank)
    par.graph_slam_sharded(lum, graph, dev, 1, rank, world, device=device)      # warm-up iteration (unseeded)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    iters = 3
    t0 = time.perf_counter()
    ret, it = par.graph_slam_sharded(lum, graph, dev, iters, rank, world, device=device)
    ctx.synchronize()
    el = time.perf_counter() - t0
    tmax, links = par.reduce_timing(el, len(par.shard_units(graph.get_nr_links(), rank, world)) * iters, device=device)
    poses = torch.tensor(np.array([d.get_pose()[0] for d in dev]).reshape(-1), dtype=torch.float64, device=device)
    identical = True
    if dist is not None:
        lo, hi = poses.clone(), poses.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
    for d in dev:
        d.destroy()
    ctx.close()
    if rank != 0:
        return None
    dim = 6 * (n_scans - 1)
    return {"workload": "BASELINE configs[3] relaxation: %d scans x %d points replicated per GPU, %d links (Graph(n, 750^2, "
                        "20)), -D 25, links round-robin over %d rank(s)" % (n_scans, n_pts, graph.get_nr_links(), world),
            "s_per_lum_iteration": tmax / iters, "link_evaluations_per_s": links / tmax, "lum_iterations": iters,
            "seeded": True, "allreduce_bytes_per_iteration": 8 * (dim * dim + dim) if world > 1 else 0,
            "collective": "one torch.distributed all_reduce (NCCL) of the packed fp64 [G|B] per LUM iteration" if world > 1 else "none (1 rank)",
            "poses_bit_identical_across_ranks": identical, "ret": ret,
            "timing": "wall clock around %d LUM iterations after one warm-up iteration, max over ranks" % iters}


def main():
    a = parse()
    if a.rows:   # single GPU, rank 0 only; the CPU arms of these rows execute the oracle, hence under tests/
        if int(os.environ.get("RANK", "0")) == 0:
            import runpy
            runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "rows_bench.py"),
                           run_name="__main__")
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl != "reference":
        if rank == 0:   # by path: the package __init__ raises while the library is still missing
            spec = importlib.util.spec_from_file_location(
                "_b200icp_build", os.path.join(os.path.dirname(os.path.abspath(__file__)), "3dtk_b200", "build.py"))
            build = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(build)
            build.build()
        icp = importlib.import_module("3dtk_b200")

    config = {"workload": "synthetic 2-scan pair (BASELINE configs[1]): %d pts each, known SE(3) offset, "
                          "point-to-point icp6D_QUAT, d=25 i=%d epsICP=1e-5" % (a.points, a.max_iter),
              "points_model": a.points, "points_data": a.points, "algo": "icp6D_QUAT", "max_dist": 25.0,
              "max_iter": a.max_iter, "eps_icp": 1e-5, "exact_nn": bool(a.exact),
              "sharding": "one scan pair per GPU, no data-path collective" if a.shard == "pairs" else
                          "ONE pair; data points split across GPUs; moments all-reduced in-kernel over NVLink peer memory",
              "l2": "flushed between steps (256 MiB write)",
              "iteration_form": "two-kernel (B200ICP_SPLIT=1)" if os.environ.get("B200ICP_SPLIT", "0")[:1] == "1"
                                else "fused kernel (default)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return 0
        model, data, _ = make_pair_standalone(a.points, 0)
        matches = max(1, min(a.steps, a.ref_matches))
        r = cpu_reference_run(model, data, a.max_iter, matches)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref3dtk_omp.so not built"}))
            return 0
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "iters_per_sec": r["iters_per_sec"],
                "iterations_per_match": r["iterations"] / r["matches"], "matches_timed": r["matches"],
                "note": "one step = one full match (same config as the GPU arm); the run is bounded to %d timed "
                        "match(es) and no warm-up matches (CPU code; the k-d tree is built once, outside the "
                        "timed region, as the reference's own timer does)" % r["matches"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner to stdout; rank 0's stdout must carry exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    stream = torch.cuda.current_stream()
    ctx = icp.Context(local_rank, stream=stream.cuda_stream)

    shard_q = a.shard == "queries" and world > 1
    model, data, Ptrue = make_pair(icp, a.points, 0 if shard_q else rank)
    n_total = a.points
    if shard_q:
        # whole model on every rank, contiguous slice of the data scan (scan.cc:1335-1342)
        step_q = -(-len(data) // world)
        data = np.ascontiguousarray(data[rank * step_q:min((rank + 1) * step_q, len(data))])
        handle = ctx.comm_create(rank, world)
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        ctx.comm_connect_ipc(handles)
        dist.barrier()
    n = a.points
    n_data = len(data)
    # pinned host staging (e2e path)
    h_model = torch.from_numpy(model).pin_memory()
    h_data = torch.from_numpy(data).pin_memory()
    m_scan = icp.Scan.from_host_pointers(ctx, h_model.data_ptr(), None, n, a.cell_edge, 25.0)
    d_scan = icp.Scan.from_host_pointers(ctx, h_data.data_ptr(), None, n_data, a.cell_edge, 25.0)
    ginfo = m_scan.grid_info()
    eng = icp.icp6D(ctx, algo=icp.ALGO_QUAT, max_dist_match=25.0, max_num_iterations=a.max_iter,
                    epsilon_icp=1e-5, exact=bool(a.exact), sharded=shard_q)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ident = np.eye(4).reshape(16).copy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        d_scan.set_pose(ident, ident)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = eng.match(m_scan, d_scan)
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1), r

    for _ in range(a.warmup):
        resident_step()
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    barrier()
    if not os.environ.get("B200_BENCH_NO_SAMPLER"):
        sampler.start()
    ms_total, iters_total, launches = 0.0, 0, 0
    wall0 = time.perf_counter()
    last = None
    for _ in range(a.steps):
        ms, r = resident_step()
        ms_total += ms
        iters_total += r["iterations_run"]
        launches += int(r["result"].kernel_launches)
        last = r
    barrier()
    wall = time.perf_counter() - wall0
    T_final, _ = d_scan.get_pose()
    pose_err = float(np.linalg.norm(T_final - Ptrue) / np.linalg.norm(Ptrue))

    # ---- kernel-level timing for the roofline (same workload, CUDA events inside the library)
    eng_prof = icp.icp6D(ctx, algo=icp.ALGO_QUAT, max_dist_match=25.0, max_num_iterations=a.max_iter,
                         epsilon_icp=1e-5, exact=bool(a.exact), profile=True, sharded=shard_q)
    d_scan.set_pose(ident, ident)
    flush.fill_(1)
    rp = eng_prof.match(m_scan, d_scan)
    nn_ms = rp["result"].nn_kernel_ms
    solve_ms = rp["result"].solve_kernel_ms

    # ---- e2e: host buffers -> upload + grid build -> match -> pose back
    def e2e_step():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ms_ = icp.Scan.from_host_pointers(ctx, h_model.data_ptr(), None, n, a.cell_edge, 25.0)
        ds_ = icp.Scan.from_host_pointers(ctx, h_data.data_ptr(), None, n_data, a.cell_edge, 25.0)
        r_ = eng.match(ms_, ds_)
        pose = ds_.get_pose()[0]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ms_.destroy(); ds_.destroy()
        return dt, r_, pose
    e2e_steps = max(2, min(a.steps, 10))
    e2e_step()
    barrier()
    e2e_t, e2e_it = 0.0, 0
    for _ in range(e2e_steps):
        dt, r_, _ = e2e_step()
        e2e_t += dt
        e2e_it += r_["iterations_run"]
    barrier()
    clocks = sampler.stop()   # sampled across both timed regions (resident steps and e2e steps)

    # ---- reduce over ranks: max time, summed work
    ms_total_local = ms_total
    vals = torch.tensor([ms_total, float(iters_total), e2e_t, float(e2e_it), float(launches)],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_t = float(mx[0]), float(mx[2])
        iters_sum, e2e_it_sum, launches = float(sm[1]), float(sm[3]), int(sm[4])
    else:
        iters_sum, e2e_it_sum = float(iters_total), float(e2e_it)
    strong = lum_obj = None
    if not a.no_extra and not shard_q:
        if world > 1:
            # t_1 of the same pair: rank 0's own (weak) pair IS pair 0
            t1 = torch.tensor([ms_total_local / a.steps if rank == 0 else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
            try:
                strong = strong_section(icp, torch, dist, local_rank, rank, world, a, float(t1[0]))
            except Exception as ex:
                strong = {"error": repr(ex)}
        try:
            lum_obj = lum_section(icp, torch, dist, local_rank, rank, world)
        except Exception as ex:
            lum_obj = {"error": repr(ex)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    if shard_q:   # one pair: every rank ran the same iterations; the job's work is N_d_total x iterations
        iters_sum, e2e_it_sum = iters_sum / world, e2e_it_sum / world
    value = n_total * iters_sum / (ms_total * 1e-3)
    e2e_value = n_total * e2e_it_sum / e2e_t
    hbm_peak, peak_src = peaks()
    n_occ = ginfo["n_occupied"]
    b_alg = 16.0 * n + 16.0 * n + 8.0 * n_occ + 512.0
    split = os.environ.get("B200ICP_SPLIT", "0")[:1] == "1"
    if split:
        # opt-in two-kernel form: one ICP iteration = icp_stream_kernel (every point, TMA-streamed) + icp_search_kernel
        # (queued exact searches + solve).  B_alg is per iteration, so it is divided by both launches.
        stream_ms = solve_ms
        iter_ms = nn_ms + stream_ms
        kname = "icp_search_kernel<POINT,%s> (+ icp_stream_kernel: one ICP iteration = 2 launches)"
    else:
        stream_ms = None
        iter_ms = nn_ms
        kname = "icp_iter_kernel<P2P,POINT,%s>"
    achieved = b_alg / (iter_ms * 1e-3) / 1e9 if iter_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": kname % ("EXACT" if a.exact else "FP32"),
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": (achieved / hbm_peak) if achieved else None, "traffic": ncu_traffic()[0],
                "traffic_source": (ncu_traffic()[1] or {}).get("source"),
                "algorithmic_bytes_per_launch": b_alg, "kernel_ms": nn_ms, "iteration_ms": iter_ms,
                "peak_source": peak_src,
                "candidate_evals_per_s": (ncu_traffic()[1] or {}).get("candidate_evaluations_per_s_mean"),
                "candidate_evals_fp32_frac": ((ncu_traffic()[1] or {}).get("candidate_evaluations_per_s_mean") or 0.0) * 8.0 / FP32_PEAK_FLOPS
                                             if ncu_traffic()[1] else None,
                "candidate_evals_note": "SURVEY 8d second figure: fp32 squared-distance evaluations per second (ncu source "
                                        "counters of iterations 1/15/40, profiles/ncu_traffic.json) x 8 flop-equivalents "
                                        "against the B200's non-tensor fp32 peak (148 SMs x 128 lanes x 2 x 1.965 GHz = "
                                        "74.4 TFLOP/s): the distance arithmetic is ~15 % of the kernel's instructions, the "
                                        "rest walks cells and ranges (profiles/r02_ncu_lines_iter1.txt)",
                "note": "B_alg = 16*N_d + 16*N_m + 8*N_occupied_cells + 512 per ICP iteration (SURVEY 8d); the "
                        "kernel is bound by L1/L2 latency and instruction issue of the candidate scan, not by HBM "
                        "(model + tables fit the 126 MB L2), see DESIGN.md section 4"}
    if split:
        stream_bytes = 72.0 * n_data   # d0 32 B + cached neighbour 32 B + budget 4 B read, budget 4 B written
        roofline["stream_kernel"] = {
            "bound": "hbm", "kernel_ms": stream_ms, "bytes_per_launch": stream_bytes,
            "achieved": stream_bytes / (stream_ms * 1e-3) / 1e9 if stream_ms > 0 else None,
            "frac": stream_bytes / (stream_ms * 1e-3) / 1e9 / hbm_peak if stream_ms > 0 else None,
            "note": "mean over all iterations of the match, launch overhead included"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong" if shard_q else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "iters_per_sec": iters_sum / (ms_total * 1e-3), "iterations_per_match": iters_total / a.steps,
            "pose_rel_frobenius_vs_truth": pose_err, "grid": ginfo, "wall_s_timed_region": wall,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (n + n_data) * 24,
                    "d2h_bytes_per_step": 2 * 640 * max(1, (a.max_iter + 3) // 4) + 16 * a.max_iter,
                    "ms_per_step": 1e3 * e2e_t / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches, "roofline": roofline}
    if strong is not None:
        line["strong"] = strong
    if lum_obj is not None:
        line["lum_link_sharded"] = lum_obj
    line["timing_note"] = ("value / e2e count every iteration of the match incl. iteration 0 (the reference's own "
                           "timer starts at iteration 1, icp6D.cc:127); grid build is outside `value`, inside `e2e`")
    if world == 1 and not a.no_parity:
        # ---- BASELINE.md section 3 gate at the bench size: same arrays through the compiled reference
        try:
            pr = cpu_reference_run(model, data, a.max_iter, 1, serial_semantics=True)
        except Exception as ex:
            pr = None
            line["parity_error"] = repr(ex)
        if pr:
            import orclib
            ref_T = pr["last"]["transmat"]
            got_np = np.asarray(last["npairs"], dtype=np.int64)
            line["pose_rel_frobenius_vs_reference"] = float(orclib.rel_frobenius(T_final, ref_T))
            line["npairs_per_iteration_equal"] = bool(len(got_np) == len(pr["last"]["npairs"]) and
                                                      np.array_equal(got_np, pr["last"]["npairs"]))
            line["iterations_equal"] = bool(last["iterations"] == pr["last"]["iterations"])
            line["parity"] = {"oracle": "oracle/_ref (compiled 3DTK kd.cc / searchTree.cc / icp6Dquat.cc), " + pr["sample"],
                              "gate": "pose_rel_frobenius_vs_reference < 1e-4 (BASELINE.md section 3)",
                              "passed": bool(line["pose_rel_frobenius_vs_reference"] < 1e-4),
                              "reference_iterations": int(pr["last"]["iterations_run"]),
                              "reference_seconds": pr["ms_per_step"] * 1e-3}
    if world == 1 and not a.no_cpu_baseline:
        try:
            cb = cpu_reference_run(model, data, a.max_iter, 1)
        except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
            cb = None
            line["cpu_baseline_error"] = repr(ex)
        if cb:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["iters_per_sec"] = cb["iters_per_sec"]
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
