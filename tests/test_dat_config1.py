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


def test_reduction_oracle_counts_on_raw_subset():
    raw = GOLD["scan1_raw_first30000"]
    red = orclib.octree_centres(raw, 10.0)
    assert len(raw) == 30000 and 1000 < len(red) < 30000
    # every raw point lies inside the cube (edge <= 2*voxel, Boctree.h:1168) of exactly one output centre
    d = np.abs(raw[:, None, :] - red[None, :200, :]).max(-1)
    assert (d.min(0) <= 10.0 + 1e-9).all()


@pytest.mark.gpu
def test_gpu_octree_reduction_matches_oracle(icp, ctx):
    raw = np.ascontiguousarray(GOLD["scan1_raw_first30000"])
    for voxel in (10.0, 3.0, 37.5):
        got = icp.reduce_octree_center(ctx, raw, voxel)
        want = orclib.octree_centres(raw, voxel)
        assert got.shape == want.shape and np.array_equal(got, want)      # same centres, same depth-first order, bit-exact
    # points exactly on splitting planes (they go to the upper child: Boctree.h:268,1784-1815; known answer from the
    # compiled reference in tests/test_full_reference.py) and a larger synthetic cloud
    grid = np.stack(np.meshgrid(*[np.arange(-8.0, 9.0, 2.0)] * 3, indexing="ij"), -1).reshape(-1, 3)
    assert np.array_equal(icp.reduce_octree_center(ctx, grid, 1.0), orclib.octree_centres(grid, 1.0))
    big = icp.synth_scene(7, 61, 200000, 0.5)
    got, want = icp.reduce_octree_center(ctx, big, 10.0), orclib.octree_centres(big, 10.0)
    assert np.array_equal(got, want) and len(got) < len(big)
    one = icp.reduce_octree_center(ctx, np.array([[1.0, 2.0, 3.0]]), 10.0)
    assert np.array_equal(one, orclib.octree_centres(np.array([[1.0, 2.0, 3.0]]), 10.0))
