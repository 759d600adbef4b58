"""Rows f1 / f2 / f3 / a12 against the reference's OWN classes.

oracle/_ref/libref3dtk_full.so holds scan.cc, basicScan.cc, icp6D.cc, lum6Deuler.cc, lum6Dquat.cc, graph.cc,
Boctree.h ... compiled UNMODIFIED (oracle/full_harness.cc, oracle/shim/).  tests/golden/full_vectors.npz stores what
those classes produce (tests/golden/make_full_golden.py); where the library is present the same comparisons also
run live.  CPU tests pin the oracle's restatements (orclib / oracle_icp.cpp); GPU tests compare the product."""
import hashlib
import importlib
import os

import numpy as np
import pytest

import doicp_case
import full_case
import orclib

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.zeros(3)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "full_vectors.npz"))


def _lum_inputs(icp):
    locs, rp, rt = full_case.lum_sequence(icp)
    T = np.array([icp.euler_to_matrix4(p, t) for p, t in zip(rp, rt)])
    scans = [icp.transform_points(M, x) for M, x in zip(T, locs)]     # Scan::transformReduced(transMatOrg)
    return locs, rp, rt, scans, T


# ------------------------------------------------------------------ f1: octree reduction
def test_octree_port_equals_reference(icp, gold):
    cloud = icp.synth_scene(7, 11, 30000, 0.5)
    for v in full_case.OCT_VOXELS:
        want = gold["oct_centres_v%g" % v]
        got = orclib.octree_centres(cloud, v)
        assert got.shape == want.shape and np.array_equal(got, want)      # same centres, same (depth-first) order


def test_octree_points_on_splitting_planes_go_up(icp, gold):
    """Known answer from the compiled reference: Scan::calcReducedPoints builds its tree with the T** constructor, whose
    partition keeps `p < split` below (Boctree.h:268,1784-1815), so a point ON a plane belongs to the upper child."""
    assert np.array_equal(orclib.octree_centres(full_case.OCT_PLANE_KAT, 0.4), gold["oct_plane_kat"])


def test_octree_live_reference_on_bundled_scan(gold):
    """dat/scan001.3d -r 10 through Scan::calcReducedPoints: count + hash stored in the golden file."""
    dat = "/root/reference/dat/scan001.3d"
    if orclib.full() is None or not os.path.exists(dat):
        pytest.skip("needs /root/reference and oracle/_ref/libref3dtk_full.so")
    pts = np.loadtxt(dat)[:, :3]
    with orclib.FullRefScans([pts], [Z], [Z], voxel=10.0) as fr:
        red = fr.get(0, "xyz reduced")
    assert [len(pts), len(red)] == list(gold["oct_dat001_count"])
    assert np.array_equal(np.frombuffer(hashlib.sha256(np.ascontiguousarray(red).tobytes()).digest(), dtype=np.uint8),
                          gold["oct_dat001_sha256"])
    assert np.array_equal(orclib.octree_centres(pts, 10.0), red)


@pytest.mark.gpu
def test_gpu_octree_equals_reference(icp, ctx, gold):
    cloud = icp.synth_scene(7, 11, 30000, 0.5)
    for v in full_case.OCT_VOXELS:
        got = icp.reduce_octree_center(ctx, cloud, v)
        assert np.array_equal(got, gold["oct_centres_v%g" % v])


@pytest.mark.gpu
def test_gpu_octree_average_and_random_with_normals_equal_reference(icp, ctx, gold):
    """-O -1 (GetOctTreeAvg) and -O 1 (GetOctTreeRandom, glibc rand() stream of a fresh process) with the PointType
    that carries normals through the reduction: bit-identical to Scan::calcReducedPoints of the compiled reference."""
    cloud = icp.synth_scene(7, 12, 8000, 0.5)
    rngn = np.random.default_rng(3)
    nrm = rngn.normal(size=cloud.shape)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    centres = gold["oct_small_centres"]
    # -O -1: per-voxel means of xyz and normals, voxels in the reference's depth-first order.  Floating point: the
    # reference sums a leaf in the order its unstable partitions left the points -> agreement to rounding (tolerance
    # 1e-12 relative, observed ~1e-16)
    xyz, n = icp.reduce_octree(ctx, cloud, 10.0, nrpts=-1, normals=nrm)
    assert xyz.shape == gold["oct_avg_xyz"].shape
    np.testing.assert_allclose(xyz, gold["oct_avg_xyz"], rtol=1e-12, atol=1e-12)
    # (the reference's averaged NORMALS are not comparable: GetOctTreeAvg accumulates into `new T[POINTDIM]` without
    #  zeroing it -- `avgp[k] += 0`, Boctree.h:960-963 -- and the attribute slots come back holding stale heap contents,
    #  coordinates of freed points in the stored golden.  Here: true per-voxel means.)
    assert np.all(np.linalg.norm(n, axis=1) <= 1.0 + 1e-12)
    key = {tuple(p): i for i, p in enumerate(cloud)}
    single = [(k, key[tuple(p)]) for k, p in enumerate(xyz) if tuple(p) in key]       # one-point voxels
    assert len(single) > 100 and all(np.array_equal(n[k], nrm[r]) for k, r in single)
    np.testing.assert_allclose(icp.reduce_octree(ctx, cloud, 10.0, nrpts=-1), gold["oct_avg_xyz"], rtol=1e-12, atol=1e-12)
    # -O 1: one input point per voxel, voxel k in depth-first order, with its own normal.  The reference picks the same
    # NUMBER (glibc rand stream) out of a differently ordered leaf, so the point may differ -- it lies in the same voxel
    xyz, n = icp.reduce_octree(ctx, cloud, 10.0, nrpts=1, normals=nrm, rand_seed=1, rand_skip=0)
    ref_xyz = gold["oct_rnd_xyz"]
    assert xyz.shape == ref_xyz.shape == centres.shape
    half = np.abs(ref_xyz - centres).max() * 1.0000001
    assert np.all(np.abs(xyz - centres) <= half)                          # inside voxel k
    rows = np.array([key[tuple(p)] for p in xyz])                         # every output IS an input point ...
    assert np.array_equal(n, nrm[rows])                                   # ... carrying its own normal
    assert (xyz == ref_xyz).all(axis=1).mean() > 0.3                      # single-point voxels must agree exactly
    # the centre mode through the general entry point
    assert np.array_equal(icp.reduce_octree(ctx, cloud, 10.0, nrpts=0), centres)
    with pytest.raises(icp.B200ICPError):
        icp.reduce_octree(ctx, cloud, 10.0, nrpts=3)


def test_glibc_rand_restatement(icp):
    import ctypes as C
    libc = C.CDLL("libc.so.6")
    for seed in (1, 2024):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(1000)], dtype=np.int32)
        assert np.array_equal(icp.glibc_rand(seed, 1000), want)
        assert np.array_equal(icp.glibc_rand(seed, 10, skip=990), want[990:])


# ------------------------------------------------------------------ f3: doICP
def _doicp_key(eP, meta, mx):
    return "doicp_eP%d_meta%d_max%d" % (eP, meta, mx)


@pytest.mark.parametrize("variant", doicp_case.VARIANTS)
def test_doicp_port_equals_reference(icp, gold, variant):
    eP, meta, mx = variant
    scans, org = doicp_case.make_sequence(icp)
    r = orclib.do_icp(orclib.port_match, scans, org, extrapolate_pose=eP, meta=meta, max_num_metascans=mx,
                      **doicp_case.PARAMS)
    want = gold[_doicp_key(eP, meta, mx) + "_transmats"]
    for i in range(len(scans)):
        assert orclib.rel_frobenius(r["transmats"][i], want[i]) < 1e-9, i


@pytest.mark.gpu
@pytest.mark.parametrize("variant", doicp_case.VARIANTS)
def test_gpu_doicp_equals_reference(icp, ctx, gold, variant):
    eP, meta, mx = variant
    scans, org = doicp_case.make_sequence(icp)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, org):
        d.set_pose(t, None)
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=30, epsilon_icp=1e-5)
    frames = icp.Frames(len(scans))
    eng.doICP(dev, extrapolate_pose=eP, meta=meta, max_num_metascans=mx, transmat_org=org, frames=frames)
    want = gold[_doicp_key(eP, meta, mx) + "_transmats"]
    for i, d in enumerate(dev):
        assert orclib.rel_frobenius(d.get_pose()[0], want[i]) < 1e-8, i
    # frames of the last scan: as many, same types, same matrices as Scan::transform appended in the reference
    wf, wt = gold[_doicp_key(eP, meta, mx) + "_last_frames"], gold[_doicp_key(eP, meta, mx) + "_last_frame_types"]
    got = frames.get(len(scans) - 1)
    assert [t for _, t in got] == list(wt)
    for (m, _), w in zip(got, wf):
        assert orclib.rel_frobenius(m, w) < 1e-8


# ------------------------------------------------------------------ a12 / f2: link covariances
def test_lum_link_port_equals_reference(icp, gold):
    model, data, dp, dt = full_case.cov_pair(icp)
    data_g = icp.transform_points(icp.euler_to_matrix4(dp, dt), data)
    C, CD, m = orclib.port_lum_link(model, data_g, 625.0)
    np.testing.assert_allclose(C.reshape(6, 6), gold["cov_euler_C"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_euler_C"]).max())
    np.testing.assert_allclose(CD, gold["cov_euler_CD"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_euler_CD"]).max())


@pytest.mark.gpu
def test_gpu_link_covariances_equal_reference(icp, ctx, gold):
    model, data, dp, dt = full_case.cov_pair(icp)
    data_g = icp.transform_points(icp.euler_to_matrix4(dp, dt), data)
    first, second = icp.Scan(ctx, model, max_dist_hint=25.0), icp.Scan(ctx, data_g, max_dist_hint=25.0)
    C, CD, m = icp.lum_link(ctx, first, second, 625.0)
    np.testing.assert_allclose(C.reshape(6, 6), gold["cov_euler_C"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_euler_C"]).max())
    np.testing.assert_allclose(CD, gold["cov_euler_CD"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_euler_CD"]).max())
    Cq, CDq, mq = icp.lum_link_quat(ctx, first, second, 625.0)                 # lum6DQuat::covarianceQuat
    assert mq == m
    np.testing.assert_allclose(Cq.reshape(7, 7), gold["cov_quat_C"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_quat_C"]).max())
    np.testing.assert_allclose(CDq, gold["cov_quat_CD"], rtol=1e-9, atol=1e-9 * np.abs(gold["cov_quat_CD"]).max())


# ------------------------------------------------------------------ f2: graph + relaxation
def test_graph_from_poses_port_equals_reference(icp, gold):
    _, rp, _, _, _ = _lum_inputs(icp)
    assert np.array_equal(orclib.port_graph_from_poses(rp, 60.0 ** 2, 1), gold["graph_links"])
    assert np.array_equal(icp.Graph.from_poses(rp, cldist2=60.0 ** 2, loopsize=1).links, gold["graph_links"])


def test_lum_port_equals_reference(icp, gold):
    """orc_lum_graph_slam (restatement of lum6DEuler::doGraphSlam6D / FillGB3D, LU solve) against the compiled
    lum6Deuler.cc + graphSlam6D.cc (CXSparse stand-in: dense Cholesky): poses agree far below the 1e-4 gate."""
    _, _, _, scans, T = _lum_inputs(icp)
    p = full_case.LUM_PARAMS
    r = orclib.port_lum_graph_slam(scans, full_case.LUM_LINKS, p["max_dist_lum"] ** 2, p["nr_it"], p["eps_lum"], T)
    for i in range(len(scans)):
        assert orclib.rel_frobenius(r["transmats"][i], gold["lum_transmats"][i]) < 1e-7, i
    assert abs(r["ret"] - float(gold["lum_ret"][0])) < 1e-6


@pytest.mark.gpu
def test_gpu_lum_equals_reference(icp, ctx, gold):
    _, _, _, scans, T = _lum_inputs(icp)
    p = full_case.LUM_PARAMS
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, T):
        d.set_pose(t, None)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=p["max_dist_lum"], epsilon_lum=p["eps_lum"])
    ret, it = lum.doGraphSlam6D(icp.Graph(full_case.LUM_LINKS, len(scans)), dev, p["nr_it"])
    for i, d in enumerate(dev):
        assert orclib.rel_frobenius(d.get_pose()[0], gold["lum_transmats"][i]) < 1e-7, i
    assert abs(ret - float(gold["lum_ret"][0])) < 1e-6


# ------------------------------------------------------------------ live cross-checks of the harnesses
def test_restated_match_glue_equals_the_reference_own_match(icp):
    """oracle/ref_harness.cc::ref_match restates the loop of icp6D::match around compiled reference objects; the real
    icp6D::match of libref3dtk_full.so must leave the same pose bit for bit."""
    if orclib.full() is None or orclib.ref() is None:
        pytest.skip("needs oracle/_ref (built from /root/reference)")
    from conftest import make_pair
    model, data, _ = make_pair(icp, 20000, 15000)
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z]) as fr:
        it = fr.match(0, 1, algo=1)
        T = fr.pose(1)["transmat"]
    want = orclib.ref_match(model, data, algo=1)
    assert it == want["iterations"] and np.array_equal(T, want["transmat"])
