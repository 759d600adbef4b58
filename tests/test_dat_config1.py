"""BASELINE configs[0]: the reference's bundled real scans (dat/scan000..002, 81 360 points each), octree-reduced
with `-r 10` (tests/golden/make_dat_fixture.py restates the voxel-centre reduction and stores the reduced clouds
plus the compiled reference's results), matched sequentially with icp6D_QUAT, -i 20, max_dist 25.
Real data: partial overlap, ~30 % of the points have no partner within 25 cm."""
import os

import numpy as np
import pytest

import orclib

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dat_reduced.npz"))
SCANS = [np.ascontiguousarray(GOLD["scan%d_xyz_reduced" % k]) for k in range(3)]


def test_fixture_shape():
    assert [int(GOLD["scan%d_raw_count" % k][0]) for k in range(3)] == [81360] * 3      # wc -l dat/scan00*.3d
    assert [len(s) for s in SCANS] == [8849, 8569, 7446]


def test_oracle_reproduces_reference_on_real_scans():
    dal = orclib.identity()
    for k in (1, 2):
        r = orclib.port_match(SCANS[k - 1], SCANS[k], algo=1, max_dist=25.0, max_iter=20, eps=1e-5, model_dalignxf=dal)
        assert r["iterations"] == int(GOLD["match%d_iterations" % k][0])
        assert np.array_equal(r["npairs"], GOLD["match%d_npairs" % k])
        assert orclib.rel_frobenius(r["transmat"], GOLD["match%d_transmat" % k]) < 1e-10
        dal = r["dalignxf"]


@pytest.mark.gpu
@pytest.mark.parametrize("exact", [True, False])
def test_gpu_sequential_icp_on_real_scans(icp, ctx, exact):
    scans = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in SCANS]
    eng = icp.icp6D(ctx, algo=icp.ALGO_QUAT, max_dist_match=25.0, max_num_iterations=20, epsilon_icp=1e-5, exact=exact)
    for k in (1, 2):
        r = eng.match(scans[k - 1], scans[k])
        T, _ = scans[k].get_pose()
        want = GOLD["match%d_transmat" % k]
        if exact:
            assert r["iterations"] == int(GOLD["match%d_iterations" % k][0])
            assert np.array_equal(r["npairs"], GOLD["match%d_npairs" % k].astype(np.uint64))   # same pairs every iteration
            np.testing.assert_allclose(r["rms"], GOLD["match%d_rms" % k], rtol=1e-9)
            assert orclib.rel_frobenius(T, want) < 1e-8
        else:
            assert orclib.rel_frobenius(T, want) < 1e-4        # north-star tolerance for fp32 decisions
