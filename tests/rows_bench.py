#!/usr/bin/env python
"""Measurement of the SURVEY 8f rows (the callers and formats either side of the hot path), one JSON object per row:
GPU wall time through the C ABI with HOST buffers (copies included, after a warm-up call), algorithmic bytes and the
bandwidth they imply against MEASURED_PEAKS.json, and the CPU arm (compiled reference `oracle/_ref` where the
reference function compiles, otherwise the oracle port) timed on a bounded sample on this box's host cores.
    python bench.py --rows > profiles/rNN_rows_bench.json        (bench.py runs this file)
Lives under tests/ because its CPU arm executes the oracle libraries (test infrastructure); not part of bench.py's
one-line contract (that is the configs[1] match) -- these are the per-row numbers DESIGN.md cites."""
import importlib, json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.dirname(os.path.abspath(__file__))]
import orclib
icp = importlib.import_module("3dtk_b200")
ctx = icp.Context(0)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(peaks.get("hbm_gbs", 6555.0))
rows = []


def timed(fn, reps=3):
    fn()                                   # warm-up (allocations, first-touch)
    ctx.synchronize()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter(); fn(); ctx.synchronize(); best = min(best, time.perf_counter() - t)
    return best


def row(name, ref, units, unit_name, gpu_s, alg_bytes, cpu_s, cpu_units, cpu_kind, note=""):
    r = {"row": name, "reference": ref, "units": units, "unit": unit_name, "gpu_ms": 1e3 * gpu_s,
         "gpu_units_per_s": units / gpu_s, "algorithmic_bytes": alg_bytes,
         "achieved_gbs": alg_bytes / gpu_s / 1e9, "hbm_peak_gbs": HBM, "frac_of_hbm_peak": alg_bytes / gpu_s / 1e9 / HBM,
         "cpu": {"kind": cpu_kind, "sample_units": cpu_units, "seconds": cpu_s, "units_per_s": cpu_units / cpu_s if cpu_s else None},
         "speedup_vs_cpu": (units / gpu_s) / (cpu_units / cpu_s) if cpu_s else None, "note": note}
    rows.append(r)
    print(json.dumps(r), flush=True)


# ---- f1 octree reduction: 10M points -> voxel 10
n = 10_000_000
big = icp.synth_scene(7, 900, n, 0.5)
out = {}
g = timed(lambda: out.__setitem__("r", icp.reduce_octree_center(ctx, big, 10.0)), reps=2)
nred = len(out["r"])
sample = big[:300_000]
t = time.perf_counter(); orclib.octree_centres(sample, 10.0); c = time.perf_counter() - t
row("f1 octree voxel-centre reduction (-r 10)", "scan.cc:560-601 / Boctree.h", n, "points", g, 24 * n + 24 * nred, c, len(sample),
    "port (numpy restatement, 1 core)", "host->device 240 MB + result download inside the time; %d -> %d points" % (n, nred))

# ---- a11 normals (k = 10) on 1M points
n = 1_000_000
pts = icp.synth_scene(7, 901, n, 0.5)
rpos = np.array([0.0, 150.0, 0.0])
g = timed(lambda: icp.normals_knn(ctx, pts, 10, rpos), reps=2)
L = orclib.ref()
if L is not None:
    s = np.ascontiguousarray(pts[:100_000]); o = np.empty_like(s)
    t = time.perf_counter(); L.ref_normals_knn(orclib.P(s), len(s), 10, orclib.P(rpos), orclib.P(o)); c = time.perf_counter() - t
    kind = "reference (calculateNormalsKNN, 1 core)"
else:
    s = np.ascontiguousarray(pts[:100_000]); o = np.empty_like(s)
    t = time.perf_counter(); orclib.port().orc_normals_knn(orclib.P(s), len(s), 10, orclib.P(rpos), orclib.P(o)); c = time.perf_counter() - t
    kind = "port (1 core)"
row("a11 k-NN PCA normals (k=10)", "normals.cc:220-295,518-558", n, "points", g, (24 + 16 + 24) * n, c, len(s), kind,
    "upload + grid build + kernel + download")

# ---- a12 / f2 LUM: one link of 300k x 300k, and one doGraphSlam6D iteration over 8 scans / 10 links
ns, npts = 8, 300_000
rng = np.random.default_rng(1)
scans, T = [], []
for i in range(ns):
    P = icp.euler_to_matrix4(rng.normal(0, 2.0, 3), np.deg2rad(rng.normal(0, 0.3, 3))) if i else np.eye(4).reshape(16)
    scans.append(icp.transform_points(P, icp.synth_scene(7, 910 + i, npts, 0.5))); T.append(P)
dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
for d, t_ in zip(dev, T):
    d.set_pose(t_, None)
g = timed(lambda: icp.lum_link(ctx, dev[0], dev[1], 625.0))
if L is not None:
    m = np.ascontiguousarray(scans[0]); tree = L.ref_tree_create(orclib.P(m), len(m), 0, 20)
    d1 = np.ascontiguousarray(scans[1][:100_000]); Cm, CD = np.zeros(36), np.zeros(6)
    t = time.perf_counter(); L.ref_lum_link(tree, orclib.P(orclib.identity()), orclib.P(d1), len(d1), 625.0, orclib.P(Cm), orclib.P(CD)); c = time.perf_counter() - t
    L.ref_tree_free(tree); kind = "reference getPtPairs + restated sums (1 core, tree build excluded)"
else:
    t = time.perf_counter(); orclib.port_lum_link(scans[0], scans[1][:100_000], 625.0); c = time.perf_counter() - t
    kind = "port (1 core, tree build included)"
row("a12 LUM link covariance (covarianceEuler), scans resident", "lum6Deuler.cc:94-260", npts, "queries", g,
    2 * (32 * npts) + 16 * npts + 4 * npts * 2, c, 100_000, kind, "two passes (sums, residual) + 6x6 solve on host")
links = np.array([[i, i + 1] for i in range(ns - 1)] + [[0, ns - 1], [1, ns - 2], [2, ns - 3]], dtype=np.int32)
lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0, epsilon_lum=-1.0)
graph = icp.Graph(links, ns)
g = timed(lambda: lum.doGraphSlam6D(graph, dev, 1), reps=2)
row("f2 one doGraphSlam6D iteration (FillGB3D over %d links + solve + pose update)" % len(links), "lum6Deuler.cc:265-479",
    len(links) * npts, "link queries", g, len(links) * (2 * 32 * npts + 16 * npts + 8 * npts), c * len(links) * npts / 100_000,
    len(links) * npts, "extrapolated from the link row", "dense Cholesky of the %d x %d system included" % (6 * (ns - 1), 6 * (ns - 1)))

# ---- f3 metascan build: 8 x 300k
g = timed(lambda: icp.Scan.metascan(ctx, dev, max_dist_hint=25.0).destroy(), reps=2)
if L is not None:
    allp = np.ascontiguousarray(np.concatenate(scans[:2], axis=0))
    t = time.perf_counter(); tree = L.ref_tree_create(orclib.P(allp), len(allp), 0, 20); c = time.perf_counter() - t; L.ref_tree_free(tree)
    kind = "reference KDtree build (1 core) over 600k points"
    cu = len(allp)
else:
    c, cu, kind = 0.0, 0, "unavailable"
row("f3 MetaScan search structure over %d x %d points" % (ns, npts), "metaScan.cc:27-69 / kdMeta.cc:34-72", ns * npts, "points", g,
    (32 + 24) * ns * npts + 48 * ns * npts, c, cu, kind, "export of every member through dalignxf + grid build, all on device")

# ---- b drop-in `-t gpu`: the reference's OWN serial icp6D::match loop (compiled unmodified: scan.cc, icp6D.cc,
# searchTree.cc, icp6Dquat.cc) over a GpuSearchTree as the scan's tree, next to the same loop over its k-d tree, and
# icp6D_gpu::match (the fused path behind the same virtual) -- 1M x 1M pair, oracle/_ref/libadapter3dtk_full.so
AL = orclib.adapter_full()
if AL is not None:
    n = 1_000_000
    model = icp.synth_scene(7, 42, n, 0.5)
    data = icp.transform_points(icp.m4inv(icp.euler_to_matrix4(np.array([12.0, -7.0, 5.0]), np.deg2rad([0.5, -1.0, 0.8])))[0],
                                icp.synth_scene(7, 43, n, 0.5))
    Zz = np.zeros(3)

    def loop(gpu_tree, iters):
        with orclib.FullRefScans([model, data], [Zz, Zz], [Zz, Zz], lib=AL, gpu_tree=gpu_tree) as fr:
            fr.get(0, "xyz reduced"); fr.get(1, "xyz reduced")           # on-demand copies outside the time
            AL.reff_match(fr.h[0], fr.h[1], 1, 0, 25.0, 1, 1e-5, 1, 0)   # builds the tree (first getPtPairs), 1 iteration
            t = time.perf_counter()
            it = AL.reff_match(fr.h[0], fr.h[1], 1, 0, 25.0, iters, 1e-5, 1, 0)
            return (time.perf_counter() - t) / (it + 1)

    s_gpu_tree = loop(True, 6)
    s_kd = loop(False, 2)
    with orclib.FullRefScans([model, data], [Zz, Zz], [Zz, Zz], lib=AL) as fr:
        fr.get(0, "xyz reduced"); fr.get(1, "xyz reduced")
        t = time.perf_counter(); it, ran, _, _ = fr.match_gpu(0, 1, algo=1, max_iter=50); s_fused = time.perf_counter() - t
    row("b drop-in -t gpu: reference icp6D::match loop (serial, PtPair vector, Align walk) over GpuSearchTree",
        "icp6D.cc:104-285 / searchTree.h:38-113", n, "queries per iteration", s_gpu_tree, 24 * n + 4 * n, s_kd, n,
        "the same unmodified loop over the reference KDtree (1 core)",
        "per ICP iteration; the GPU tree answers one b200icp_nn_batch per getPtPairs call, PtPairs (208 B each) are rebuilt "
        "on the host; icp6D_gpu::match (fused path, whole %d-iteration match incl. upload, grid build and the host replay "
        "of every Scan::transform) takes %.3f s = %.1f ms per iteration" % (ran, s_fused, 1e3 * s_fused / max(ran, 1)))

# ---- f4 uos reader: 1M lines
pts = icp.synth_scene(7, 920, 1_000_000, 0.5)
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "scan000.3d")
    np.savetxt(p, pts, fmt="%.6g")
    size = os.path.getsize(p)
    t = time.perf_counter(); a = icp.read_uos(p); g = time.perf_counter() - t
    t = time.perf_counter(); b = np.loadtxt(p); c = time.perf_counter() - t
    assert np.array_equal(a, b)
row("f4 uos text reader (host, parallel chunks)", "scanio/helper.cc:577-835", len(pts), "points", g, size, c, len(pts),
    "numpy.loadtxt (1 core) as the CPU arm", "%d MB file; host-only row: 'gbs' is file bytes per second, not HBM" % (size // 1_000_000))
ctx.close()
