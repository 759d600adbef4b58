#!/usr/bin/env python
"""Generates tests/golden/ref_vectors.npz from the UNMODIFIED reference (oracle/_ref/libref3dtk.so, built
by oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Inputs are stored next to the outputs so the fixtures do not depend on any generator staying stable.
Every array name says which reference entry point produced it.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib  # noqa: E402
from orclib import P  # noqa: E402


def scene(rng, n):
    """small asymmetric indoor-like cloud: floor + two walls + a box, 0.3 cm noise"""
    parts = []
    k = n // 4
    parts.append(np.c_[rng.uniform(-300, 300, k), np.zeros(k), rng.uniform(-200, 200, k)])
    parts.append(np.c_[rng.uniform(-300, 300, k), rng.uniform(0, 250, k), np.full(k, -200.0)])
    parts.append(np.c_[np.full(k, 300.0), rng.uniform(0, 250, k), rng.uniform(-200, 200, k)])
    m = n - 3 * k
    parts.append(np.c_[rng.uniform(40, 120, m), np.full(m, 80.0), rng.uniform(-60, 30, m)])
    return np.ascontiguousarray(np.vstack(parts) + rng.normal(0, 0.3, (n, 3)))


def main():
    L = orclib.ref()
    assert L is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    rng = np.random.default_rng(20240517)

    # --- KDtree::FindClosest, seed-42-style differential set (testing/kdtree/kdtree_indexed_random.cc)
    pts = rng.uniform(-10, 10, (10000, 3))
    q = rng.uniform(-10, 10, (1000, 3))
    tree = L.ref_tree_create(P(pts), len(pts), 0, 20)
    out["nn_points"], out["nn_queries"] = pts, q
    radii = np.arange(0.5, 5.01, 0.5)
    out["nn_maxdist2"] = radii
    idx_all = np.empty((len(radii), len(q)), np.int32)
    for r, md2 in enumerate(radii):
        L.ref_find_closest_batch(tree, P(q), len(q), float(md2), P(idx_all[r]), 1)
    out["nn_idx_KDtree_FindClosest"] = idx_all
    L.ref_tree_free(tree)

    # --- SearchTree::getPtPairs + the four Align functions on one pair list
    model, data = scene(rng, 4000), scene(rng, 3000)
    pose = np.empty(16)
    L.ref_euler_to_matrix4(P(np.array([3.0, -2.0, 1.5])), P(np.deg2rad([0.8, -0.6, 1.0])), P(pose))
    pinv = np.empty(16)
    L.ref_m4inv(P(pose), P(pinv))
    M = pinv.reshape(4, 4).T
    data = np.ascontiguousarray(data @ M[:3, :3].T + M[:3, 3])
    nrm = rng.normal(size=data.shape)
    S = np.empty(16)
    L.ref_euler_to_matrix4(P(np.array([1.0, 0.5, -0.7])), P(np.deg2rad([0.2, 0.3, -0.1])), P(S))
    out["pair_model"], out["pair_data"], out["pair_data_normals"], out["pair_source_alignxf"] = model, data, nrm, S
    tree = L.ref_tree_create(P(model), len(model), 0, 20)
    for mode in (0, 2):
        n = len(data)
        p1, p2, pn = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        sm, cm, cd = np.zeros(1), np.zeros(3), np.zeros(3)
        k = L.ref_get_pt_pairs(tree, P(S), P(data), P(nrm), 0, n, 0, 1, 400.0, mode, P(p1), P(p2), P(pn), P(sm),
                               P(cm), P(cd))
        out["pairs%d_p1" % mode], out["pairs%d_p2" % mode], out["pairs%d_n" % mode] = p1[:k], p2[:k], pn[:k]
        out["pairs%d_sum_cm_cd" % mode] = np.r_[sm, cm, cd]
        cmn, cdn = cm / k, cd / k
        for algo in ((1, 2, 3, 4, 5, 6) if mode == 0 else (1, 10)):
            xf = np.zeros(16)
            rms = L.ref_align(algo, k, P(p1), P(p2), P(pn), P(cmn), P(cdn), P(xf))
            out["align_mode%d_algo%d" % (mode, algo)] = np.r_[xf, rms]
    L.ref_tree_free(tree)

    # --- whole matches (harness loop around the reference's getPtPairs / Align / transform3)
    out["match_maxdist_iters_eps"] = np.array([20.0, 30, 1e-5])
    for algo, mode in ((1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (10, 2), (1, 2)):
        r = orclib.ref_match(model, data, nrm if mode else None, algo=algo, mode=mode, max_dist=20.0,
                             max_iter=30, eps=1e-5)
        out["match_algo%d_mode%d_transmat" % (algo, mode)] = r["transmat"]
        out["match_algo%d_mode%d_rms" % (algo, mode)] = r["rms"]
        out["match_algo%d_mode%d_npairs" % (algo, mode)] = r["npairs"]
        out["match_algo%d_mode%d_iterations" % (algo, mode)] = np.array([r["iterations"]])

    # --- calculateNormalsKNN (k = 10)
    npts = scene(rng, 3000)
    rpos = np.array([0.0, 120.0, 0.0])
    nout = np.empty_like(npts)
    L.ref_normals_knn(P(npts), len(npts), 10, P(rpos), P(nout))
    out["normals_points"], out["normals_rpos"], out["normals_calculateNormalsKNN_k10"] = npts, rpos, nout

    path = os.path.join(HERE, "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
