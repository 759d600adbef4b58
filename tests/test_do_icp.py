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
