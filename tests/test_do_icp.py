"""icp6D::doICP (reference src/slam6d/icp6D.cc:374-437): sequential matching with odometry extrapolation
(Scan::mergeCoordinatesWithRoboterPosition) and the metascan option (MetaScan / KDtreeMetaManaged)."""
import os

import numpy as np
import pytest

import doicp_case
import orclib

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "doicp_vectors.npz"))


def _key(eP, meta, mx):
    return "eP%d_meta%d_max%d" % (eP, meta, mx)


@pytest.mark.parametrize("eP,meta,mx", doicp_case.VARIANTS[:2])
def test_oracle_do_icp_matches_reference_golden(icp, eP, meta, mx):
    scans, org = doicp_case.make_sequence(icp)
    r = orclib.do_icp(orclib.port_match, scans, org, extrapolate_pose=eP, meta=meta, max_num_metascans=mx,
                      **doicp_case.PARAMS)
    assert list(r["iterations"]) == list(GOLD[_key(eP, meta, mx) + "_iterations"])
    for a, b in zip(r["transmats"], GOLD[_key(eP, meta, mx) + "_transmats"]):
        assert orclib.rel_frobenius(a, b) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("eP,meta,mx", doicp_case.VARIANTS)
def test_gpu_do_icp_matches_reference_golden(icp, ctx, eP, meta, mx):
    scans, org = doicp_case.make_sequence(icp)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, org):
        d.set_pose(t, None)
    p = doicp_case.PARAMS
    eng = icp.icp6D(ctx, algo=p["algo"], max_dist_match=p["max_dist"], max_num_iterations=p["max_iter"],
                    epsilon_icp=p["eps"])
    its = eng.doICP(dev, extrapolate_pose=eP, meta=meta, max_num_metascans=mx, transmat_org=org)
    assert list(its) == list(GOLD[_key(eP, meta, mx) + "_iterations"])
    for i, d in enumerate(dev):
        T, _ = d.get_pose()
        assert orclib.rel_frobenius(T, GOLD[_key(eP, meta, mx) + "_transmats"][i]) < 1e-8, i
    # the drift is removed: the scans were loaded at pose = accumulated drift (truth: identity), so the final
    # transMat is the remaining error; pairwise ICP of independently sampled noisy clouds leaves ~0.5 cm per step
    for i in range(1, len(dev)):
        T, _ = dev[i].get_pose()
        assert np.linalg.norm(T[12:15]) < 2.0 and np.linalg.norm(org[i][12:15]) > np.linalg.norm(T[12:15]), (i, T[12:15])


@pytest.mark.gpu
def test_metascan_is_the_union_of_current_positions(icp, ctx):
    scans, org = doicp_case.make_sequence(icp)
    dev = [icp.Scan(ctx, s) for s in scans[:3]]
    dev[1].transform(icp.euler_to_matrix4(np.array([4.0, -1.0, 2.0]), np.deg2rad([0.3, 0.1, -0.2])))
    meta = icp.Scan.metascan(ctx, dev, max_dist_hint=25.0)
    assert len(meta) == sum(len(s) for s in scans[:3])
    want = np.concatenate([d.download() for d in dev], axis=0)
    got = meta.download()
    assert np.array_equal(got, want)
    T, D = meta.get_pose()
    assert np.array_equal(T, orclib.identity()) and np.array_equal(D, orclib.identity())
    # nearest neighbours in the metascan == brute force over the union
    q = want[::997] + 0.3
    idx, d2, _ = meta.nn_batch(q, 625.0)
    for k in range(len(q)):
        bf = np.sum((want - q[k]) ** 2, axis=1)
        assert idx[k] == int(np.argmin(bf))
    with pytest.raises(icp.B200ICPError):
        icp.Scan.metascan(ctx, [])


@pytest.mark.gpu
def test_match_pose_log(icp, ctx):
    scans, org = doicp_case.make_sequence(icp)
    m, d = icp.Scan(ctx, scans[0], max_dist_hint=25.0), icp.Scan(ctx, scans[1], max_dist_hint=25.0)
    d.set_pose(org[1], None)
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=30, epsilon_icp=1e-5)
    r = eng.match(m, d)
    poses = eng.last_poses()
    assert len(poses) == r["iterations_run"]
    assert np.array_equal(poses[-1], d.get_pose()[0])
    for k in (1, 3):          # transMat after k iterations == a k-iteration match of the oracle
        want = orclib.port_match(scans[0], scans[1], algo=1, max_dist=25.0, max_iter=k, eps=1e-5)
        T = np.zeros(16)
        orclib.port().orc_mmult(orclib.P(want["transmat"]), orclib.P(np.ascontiguousarray(org[1])), orclib.P(T))
        assert orclib.rel_frobenius(poses[k - 1], T) < 1e-9


@pytest.mark.gpu
def test_gpu_do_icp_frames(icp, ctx, tmp_path):
    scans, org = doicp_case.make_sequence(icp)
    n = len(scans)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, org):
        d.set_pose(t, None)
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=30, epsilon_icp=1e-5)
    frames = icp.Frames(n)
    eng.doICP(dev, extrapolate_pose=True, transmat_org=org, frames=frames)
    A, I, X = icp.FRAME_ICP, icp.FRAME_ICPINACTIVE, icp.FRAME_INVALID
    for k in range(n):
        fr = frames.get(k)
        assert len(fr) == 3 * (n - 1)            # per match: start pose, after iteration 0, end pose -- on every scan
        for i in range(1, n):                    # match of scan i
            want_type = A if k == i else (I if k < i else X)
            assert [t for _, t in fr[3 * (i - 1):3 * i]] == [want_type] * 3
        assert np.array_equal(fr[-1][0], dev[k].get_pose()[0])
    # scan 1: start pose = loaded pose (scan 0 did not move, so the extrapolation is the identity);
    # second frame = pose after ONE iteration
    f1 = frames.get(1)
    np.testing.assert_allclose(f1[0][0], org[1], rtol=0, atol=1e-12)
    one = orclib.port_match(scans[0], scans[1], algo=1, max_dist=25.0, max_iter=1, eps=1e-5)
    T = np.zeros(16)
    orclib.port().orc_mmult(orclib.P(one["transmat"]), orclib.P(np.ascontiguousarray(org[1])), orclib.P(T))
    assert orclib.rel_frobenius(f1[1][0], T) < 1e-9
    # frames of scan 2 while scan 1 is matched: untouched pose, INVALID
    assert np.array_equal(frames.get(2)[0][0], org[2])
    p = tmp_path / "scan001.frames"
    frames.save(1, p)
    rows = np.loadtxt(p)
    assert rows.shape == (3 * (n - 1), 17) and list(rows[:3, 16]) == [1, 1, 1]
    np.testing.assert_allclose(rows[2, :16], f1[2][0], rtol=2e-6, atol=1e-12)      # 6 significant digits
    # max_num_iterations = 0: only the identity frame per match (icp6D.cc:109-114)
    fr0 = icp.Frames(n)
    icp.icp6D(ctx, algo=1, max_num_iterations=0).doICP(dev, frames=fr0)
    assert [len(fr0.get(k)) for k in range(n)] == [n - 1] * n
