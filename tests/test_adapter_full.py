"""The reference-side adapters inside the reference's OWN classes (oracle/_ref/libadapter3dtk_full.so =
unmodified scan.cc / basicScan.cc / icp6D.cc ... + 3dtk_b200/host/{icp6d_gpu,gpu_search_tree}.cc + the product).

  icp6D_gpu : public icp6D       match() on the device, driven through an icp6D* -- alone and under the base class's
                                 unmodified doICP (metascans, pose extrapolation, frames)
  GpuSearchTree : public SearchTree   as the tree of a BasicScan, searched by the reference's unmodified match loop
"""
import os

import numpy as np
import pytest

import doicp_case
import orclib
from conftest import make_pair

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.zeros(3)


@pytest.fixture(scope="module")
def AL():
    L = orclib.adapter_full()
    if L is None:
        pytest.skip("oracle/_ref/libadapter3dtk_full.so not built (needs /root/reference)")
    return L


def _state(fr, i):
    return fr.pose(i), fr.get(i, "xyz reduced"), fr.frames(i)


@pytest.mark.parametrize("algo,anim", [(1, -1), (2, -1), (6, -1), (1, 3)])
def test_icp6d_gpu_match_equals_reference_match(icp, AL, algo, anim):
    model, data, _ = make_pair(icp, 30000, 25001)
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z], lib=AL) as fr:      # the reference's own loop
        it_ref = AL.reff_match(fr.h[0], fr.h[1], algo, 0, 25.0, 50, 1e-5, 1, 0) if anim == -1 else None
        ref_state = [_state(fr, i) for i in range(2)]
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z], lib=AL) as fr:      # icp6D_gpu through an icp6D*
        it, ran, npairs, launches = fr.match_gpu(0, 1, algo=algo, anim=anim)
        got_state = [_state(fr, i) for i in range(2)]
    assert launches > 0 and npairs > 10000
    if anim != -1:
        # every anim'th iteration writes a frame on every scan (icp6D.cc:258-264): start, iterations 0,3,6,..., end
        n_frames = 1 + len([k for k in range(ran) if k == 0 or k % anim == 0]) + 1
        assert len(got_state[1][2][1]) == n_frames and len(got_state[0][2][1]) == n_frames
        return
    assert it == it_ref
    for (pr, xr, (fmr, ftr)), (pg, xg, (fmg, ftg)) in zip(ref_state, got_state):
        assert orclib.rel_frobenius(pg["transmat"], pr["transmat"]) < 1e-8
        assert orclib.rel_frobenius(pg["dalignxf"], pr["dalignxf"]) < 1e-8
        np.testing.assert_allclose(pg["rpos"], pr["rpos"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(xg, xr, rtol=0, atol=1e-7)                  # the scan's points end where the CPU loop leaves them
        assert list(ftg) == list(ftr)                                          # same frames on both scans
        for a, b in zip(fmg, fmr):
            assert orclib.rel_frobenius(a, b) < 1e-8


@pytest.mark.parametrize("variant", doicp_case.VARIANTS)
def test_base_class_doicp_over_icp6d_gpu_equals_reference(icp, AL, variant):
    eP, meta, mx = variant
    gold = np.load(os.path.join(HERE, "golden", "full_vectors.npz"))
    scans, org = doicp_case.make_sequence(icp)
    locals_, rpos, rtheta = [], [], []
    for s, M in zip(scans, org):
        p, t = icp.matrix4_to_euler(M)
        Minv, _ = icp.m4inv(icp.euler_to_matrix4(p, t))
        locals_.append(icp.transform_points(Minv, s)); rpos.append(p); rtheta.append(t)
    key = "doicp_eP%d_meta%d_max%d" % (eP, meta, mx)
    with orclib.FullRefScans(locals_, rpos, rtheta, lib=AL) as fr:
        fr.do_icp_gpu(meta=meta, extrapolate_pose=eP, max_num_metascans=mx if mx > 0 else -1, **doicp_case.PARAMS)
        for i in range(fr.n):
            assert orclib.rel_frobenius(fr.pose(i)["transmat"], gold[key + "_transmats"][i]) < 1e-8, i
        fm, ft = fr.frames(fr.n - 1)
        assert list(ft) == list(gold[key + "_last_frame_types"])
        for a, b in zip(fm, gold[key + "_last_frames"]):
            assert orclib.rel_frobenius(a, b) < 1e-8


def test_reference_match_loop_over_gpu_search_tree_is_bit_identical(icp, AL):
    """`-t gpu` of INTEGRATION.md: BasicScan::createSearchTreePrivate builds a GpuSearchTree; the unmodified
    icp6D::match / Scan::getPtPairs / icp6D_QUAT::Align run on top of it.  Same PtPairs -> same bits."""
    model, data, _ = make_pair(icp, 20000, 15001)
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z], lib=AL) as fr:
        it_kd = fr.match(0, 1, algo=1)
        T_kd, X_kd = fr.pose(1)["transmat"], fr.get(1, "xyz reduced")
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z], lib=AL, gpu_tree=True) as fr:
        it_gpu = fr.match(0, 1, algo=1)
        T_gpu, X_gpu = fr.pose(1)["transmat"], fr.get(1, "xyz reduced")
    assert it_gpu == it_kd
    assert np.array_equal(T_gpu, T_kd) and np.array_equal(X_gpu, X_kd)


def test_icp6d_gpu_rejects_minimizers_off_the_path(icp, AL):
    model, data, _ = make_pair(icp, 3000, 3000)
    with orclib.FullRefScans([model, data], [Z, Z], [Z, Z], lib=AL) as fr:
        with pytest.raises(RuntimeError):
            fr.match_gpu(0, 1, algo=1, mode=1)       # CLOSEST_POINT_ALONG_NORMAL_SIMPLE: not on the accelerated path
