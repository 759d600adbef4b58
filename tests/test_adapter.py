"""The reference-side adapter (3dtk_b200/host/gpu_search_tree.cc, class GpuSearchTree : public SearchTree),
compiled against the reference's own headers and linked with its compiled searchTree.o (oracle/_ref/
libadapter3dtk.so, built by oracle/Makefile).  Pairs must be bit-identical to the reference KDtree's."""
import ctypes as C
import os

import numpy as np
import pytest

import orclib
from orclib import P, vp, cl, ci, cd

ADP_SO = os.path.join(orclib.ORACLE_DIR, "_ref", "libadapter3dtk.so")


@pytest.fixture(scope="module")
def adp(icp):
    if not os.path.exists(ADP_SO):
        pytest.skip("oracle/_ref/libadapter3dtk.so not built (needs /root/reference headers)")
    L = C.CDLL(ADP_SO)
    L.adp_tree_create.restype = vp; L.adp_tree_create.argtypes = [vp, cl, cd, C.c_char_p, ci]
    L.adp_tree_free.restype = None; L.adp_tree_free.argtypes = [vp]
    L.adp_find_closest.restype = cl; L.adp_find_closest.argtypes = [vp, vp, cd, ci]
    L.adp_get_pt_pairs.restype = cl
    L.adp_get_pt_pairs.argtypes = [vp, vp, vp, vp, cl, cl, ci, ci, cd, ci, ci, vp, vp, vp, vp, vp, vp]
    return L


def test_adapter_without_gpu_raises_like_the_reference_factory(adp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pts = np.random.default_rng(0).uniform(0, 1, (10, 3))
    err = C.create_string_buffer(512)
    h = adp.adp_tree_create(P(pts), 10, 0.0, err, 512)      # std::runtime_error, as basicScan.cc:723-726 throws
    assert not h and b"no CPU fallback" in err.value


def _ref_pairs(ref, rt, S, data, nrm, maxd2, mode):
    n = len(data)
    p1, p2, pn = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
    sm, cm, cdv = np.zeros(1), np.zeros(3), np.zeros(3)
    k = ref.ref_get_pt_pairs(rt, P(S), P(data), P(nrm), 0, n, 0, 1, maxd2, mode, P(p1), P(p2), P(pn), P(sm), P(cm), P(cdv))
    return k, p1[:k], p2[:k], pn[:k], sm[0], cm, cdv


def _adp_pairs(adp, h, S, data, nrm, maxd2, mode, base_loop):
    n = len(data)
    p1, p2, pn = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
    sm, cm, cdv = np.zeros(1), np.zeros(3), np.zeros(3)
    k = adp.adp_get_pt_pairs(h, P(S), P(data), P(nrm), 0, n, 0, 1, maxd2, mode, base_loop, P(p1), P(p2), P(pn),
                             P(sm), P(cm), P(cdv))
    return k, p1[:k], p2[:k], pn[:k], sm[0], cm, cdv


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 2])
def test_adapter_pairs_equal_reference_kdtree(icp, adp, ref, mode):
    rng = np.random.default_rng(5)
    model = icp.synth_scene(7, 51, 30000, 0.5)
    S = icp.euler_to_matrix4(np.array([3.0, -2.0, 1.0]), np.deg2rad([0.5, -0.3, 0.8]))
    data = icp.transform_points(S, icp.synth_scene(7, 52, 20000, 0.5))
    nrm = np.ascontiguousarray(rng.normal(size=data.shape))
    err = C.create_string_buffer(512)
    h = adp.adp_tree_create(P(model), len(model), 25.0, err, 512)
    assert h, err.value
    rt = ref.ref_tree_create(P(model), len(model), 0, 20)
    want = _ref_pairs(ref, rt, S, data, nrm, 625.0, mode)
    got = _adp_pairs(adp, h, S, data, nrm, 625.0, mode, base_loop=0)          # batched override
    assert got[0] == want[0] and got[0] > 10000
    for a, b in zip(got[1:4] if mode else got[1:3], want[1:4] if mode else want[1:3]):
        assert np.array_equal(a, b)                                              # PtPairs bit-identical
    assert got[4] == want[4] and np.array_equal(got[5], want[5]) and np.array_equal(got[6], want[6])
    # the reference's own batch loop (compiled searchTree.cc) on top of the adapter's virtual FindClosest
    small = np.ascontiguousarray(data[:300]); nsmall = np.ascontiguousarray(nrm[:300])
    want_s = _ref_pairs(ref, rt, S, small, nsmall, 625.0, mode)
    got_s = _adp_pairs(adp, h, S, small, nsmall, 625.0, mode, base_loop=1)
    assert got_s[0] == want_s[0] and np.array_equal(got_s[1], want_s[1]) and np.array_equal(got_s[2], want_s[2])
    # single queries: the KATs through the SearchTree interface
    assert adp.adp_find_closest(h, P(model[17] + 1e-3), 1.0, 0) == 17
    assert adp.adp_find_closest(h, P(np.array([5000.0, 5000.0, 5000.0])), 625.0, 3) == -1   # another thread_num
    ref.ref_tree_free(rt)
    adp.adp_tree_free(h)
