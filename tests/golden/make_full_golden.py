#!/usr/bin/env python
"""Golden vectors from the reference's OWN classes (oracle/_ref/libref3dtk_full.so = scan.cc, basicScan.cc,
icp6D.cc, lum6Deuler.cc, lum6Dquat.cc, graph.cc, Boctree.h ... compiled unmodified; oracle/full_harness.cc):
  oct_*     Scan::calcReducedPoints + BOctTree::GetOctTreeCenter   (row f1)
  doicp_*   icp6D::doICP incl. metascans, pose extrapolation, frames (row f3; same sequence as doicp_vectors.npz)
  cov_*     lum6DEuler::covarianceEuler, lum6DQuat::covarianceQuat   (rows a12, f2)
  lum_*     lum6DEuler::doGraphSlam6D                                (row f2)
  graph_*   Graph(nodes, cldist2, loopsize)
Run in the build container:  python tests/golden/make_full_golden.py -> full_vectors.npz"""
import hashlib, importlib, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]
import orclib, doicp_case, full_case

icp = importlib.import_module("3dtk_b200")     # host helpers only (scene generator, EulerToMatrix4 ...)
assert orclib.full() is not None, "build oracle/_ref first (make -C oracle ref)"
Z = np.zeros(3)
out = {}

# ---- octree reduction
cloud = icp.synth_scene(7, 11, 30000, 0.5)
for v in full_case.OCT_VOXELS:
    with orclib.FullRefScans([cloud], [Z], [Z], voxel=v) as fr:
        out["oct_centres_v%g" % v] = fr.get(0, "xyz reduced")
    print("octree voxel", v, "->", len(out["oct_centres_v%g" % v]))
# average / random extraction, normals carried through the reduction (PointType::USE_NORMAL)
cloud_s = icp.synth_scene(7, 12, 8000, 0.5)          # a smaller cloud keeps the fixture small
rngn = np.random.default_rng(3)
cloud_n = rngn.normal(size=cloud_s.shape)
cloud_n /= np.linalg.norm(cloud_n, axis=1)[:, None]
for nrpts, tag in ((-1, "avg"), (1, "rnd")):
    orclib.full().reff_srand(1)                      # the stream a fresh process sees
    with orclib.FullRefScans([cloud_s], [Z], [Z], voxel=10.0, nrpts=nrpts, normals=[cloud_n]) as fr:
        out["oct_%s_xyz" % tag] = fr.get(0, "xyz reduced")
        out["oct_%s_nrm" % tag] = fr.get(0, "normal reduced")

    print("octree -O", nrpts, "->", len(out["oct_%s_xyz" % tag]))
with orclib.FullRefScans([cloud_s], [Z], [Z], voxel=10.0) as fr:
    out["oct_small_centres"] = fr.get(0, "xyz reduced")
with orclib.FullRefScans([full_case.OCT_PLANE_KAT], [Z], [Z], voxel=0.4) as fr:
    out["oct_plane_kat"] = fr.get(0, "xyz reduced")
dat = os.path.join("/root/reference", "dat", "scan001.3d")
if os.path.exists(dat):
    pts = np.loadtxt(dat, skiprows=0)[:, :3] if open(dat).readline().count(" ") >= 2 else np.loadtxt(dat, skiprows=1)[:, :3]
    with orclib.FullRefScans([pts], [Z], [Z], voxel=10.0) as fr:
        red = fr.get(0, "xyz reduced")
    out["oct_dat001_count"] = np.array([len(pts), len(red)])
    out["oct_dat001_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(red).tobytes()).digest(), dtype=np.uint8)
    print("dat/scan001.3d -r 10:", len(pts), "->", len(red))

# ---- doICP (sequence of tests/doicp_case.py: global clouds + poses -> local clouds + Euler poses)
scans, org = doicp_case.make_sequence(icp)
locals_, rpos, rtheta = [], [], []
for s, M in zip(scans, org):
    p, t = icp.matrix4_to_euler(M)
    Minv, _ = icp.m4inv(icp.euler_to_matrix4(p, t))
    locals_.append(icp.transform_points(Minv, s)); rpos.append(p); rtheta.append(t)
for eP, meta, mx in doicp_case.VARIANTS:
    with orclib.FullRefScans(locals_, rpos, rtheta) as fr:
        fr.do_icp(meta=meta, extrapolate_pose=eP, max_num_metascans=mx if mx > 0 else -1, **doicp_case.PARAMS)
        key = "doicp_eP%d_meta%d_max%d" % (eP, meta, mx)
        out[key + "_transmats"] = np.array([fr.pose(i)["transmat"] for i in range(fr.n)])
        fm, ft = fr.frames(fr.n - 1)
        out[key + "_last_frames"] = fm
        out[key + "_last_frame_types"] = ft
        print(key, "frames of the last scan:", len(ft))

# ---- link covariances
model, data, dp, dt = full_case.cov_pair(icp)
with orclib.FullRefScans([model, data], [Z, dp], [Z, dt]) as fr:
    out["cov_euler_C"], out["cov_euler_CD"] = fr.covariance(0, 1, 625.0, quat=False)
    out["cov_quat_C"], out["cov_quat_CD"] = fr.covariance(0, 1, 625.0, quat=True)

# ---- LUM
locs, rp, rt = full_case.lum_sequence(icp)
with orclib.FullRefScans(locs, rp, rt) as fr:
    out["graph_links"] = fr.graph_from_poses(60.0 ** 2, 1)
    ret = fr.lum_euler(full_case.LUM_LINKS, **full_case.LUM_PARAMS)
    out["lum_ret"] = np.array([ret])
    out["lum_transmats"] = np.array([fr.pose(i)["transmat"] for i in range(fr.n)])
    out["lum_rpos"] = np.array([fr.pose(i)["rpos"] for i in range(fr.n)])
    out["lum_rpostheta"] = np.array([fr.pose(i)["rpostheta"] for i in range(fr.n)])
    print("lum ret", ret, "links from poses", len(out["graph_links"]))
np.savez_compressed(os.path.join(HERE, "full_vectors.npz"), **out)
print("wrote", os.path.join(HERE, "full_vectors.npz"), os.path.getsize(os.path.join(HERE, "full_vectors.npz")), "bytes")
