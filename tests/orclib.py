"""ctypes loaders for the oracle libraries (test infrastructure).

  port()  -> oracle/_build/liboracle_icp.so   our CPU restatement (built on demand with g++)
  ref()   -> oracle/_ref/libref3dtk.so        the compiled, unmodified reference (None if absent)
  full()  -> oracle/_ref/libref3dtk_full.so   the reference's own Scan / icp6D / lum6DEuler / BOctTree classes
                                              (oracle/full_harness.cc; None if absent)
Both are checkers only; nothing in the product imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "liboracle_icp.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref3dtk.so")
REF_OMP_SO = os.path.join(ORACLE_DIR, "_ref", "libref3dtk_omp.so")
REF_FULL_SO = os.path.join(ORACLE_DIR, "_ref", "libref3dtk_full.so")
ADAPTER_FULL_SO = os.path.join(ORACLE_DIR, "_ref", "libadapter3dtk_full.so")

vp, cl, ci, cd = C.c_void_p, C.c_long, C.c_int, C.c_double
_port = None
_ref = {}


def build_port():
    src = os.path.join(ORACLE_DIR, "oracle_icp.cpp")
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)
    return PORT_SO


def build_ref():
    """Compile oracle/_ref from /root/reference when it is present (this container only)."""
    subprocess.run(["make", "-s", "-j8", "-C", ORACLE_DIR, "ref"], check=True)
    return os.path.exists(REF_SO)


def port():
    global _port
    if _port is None:
        L = C.CDLL(build_port())
        L.orc_m4inv.restype = ci; L.orc_m4inv.argtypes = [vp, vp]
        L.orc_mmult.restype = None; L.orc_mmult.argtypes = [vp, vp, vp]
        L.orc_euler_to_matrix4.restype = None; L.orc_euler_to_matrix4.argtypes = [vp, vp, vp]
        L.orc_tree_create.restype = vp; L.orc_tree_create.argtypes = [vp, cl, ci]
        L.orc_tree_free.restype = None; L.orc_tree_free.argtypes = [vp]
        L.orc_find_closest.restype = cl; L.orc_find_closest.argtypes = [vp, vp, cd]
        L.orc_find_closest_batch.restype = None
        L.orc_find_closest_batch.argtypes = [vp, vp, cl, cd, vp, vp]
        L.orc_brute_closest.restype = cl; L.orc_brute_closest.argtypes = [vp, cl, vp, cd]
        L.orc_get_pt_pairs.restype = cl
        L.orc_get_pt_pairs.argtypes = [vp, vp, vp, vp, cl, cl, cd, ci, vp, vp, vp, vp, vp, vp, vp]
        L.orc_align.restype = cd; L.orc_align.argtypes = [ci, cl, vp, vp, vp, vp, vp, ci, vp]
        L.orc_match.restype = ci
        L.orc_match.argtypes = [vp, vp, vp, vp, cl, vp, vp, ci, ci, cd, ci, cd, ci, vp, vp, vp]
        L.orc_normals_knn.restype = None; L.orc_normals_knn.argtypes = [vp, cl, ci, vp, vp]
        L.orc_lum_link.restype = cl; L.orc_lum_link.argtypes = [vp, vp, vp, cl, cd, vp, vp]
        L.orc_graph_from_poses.restype = ci; L.orc_graph_from_poses.argtypes = [vp, ci, cd, ci, vp, ci]
        L.orc_matrix4_to_euler.restype = None; L.orc_matrix4_to_euler.argtypes = [vp, vp, vp]
        L.orc_lum_graph_slam.restype = ci
        L.orc_lum_graph_slam.argtypes = [ci, vp, vp, vp, ci, cd, ci, cd, vp, vp, vp, vp, vp]
        _port = L
    return _port


def ref(omp=False):
    path = REF_OMP_SO if omp else REF_SO
    if path not in _ref:
        if not os.path.exists(path):
            _ref[path] = None
        else:
            L = C.CDLL(path)
            L.ref_max_threads.restype = ci
            L.ref_tree_create.restype = vp; L.ref_tree_create.argtypes = [vp, cl, ci, ci]
            L.ref_tree_free.restype = None; L.ref_tree_free.argtypes = [vp]
            L.ref_find_closest.restype = cl; L.ref_find_closest.argtypes = [vp, vp, cd, ci]
            L.ref_find_closest_batch.restype = None
            L.ref_find_closest_batch.argtypes = [vp, vp, cl, cd, vp, ci]
            L.ref_knn.restype = ci; L.ref_knn.argtypes = [vp, vp, ci, vp]
            L.ref_get_pt_pairs.restype = cl
            L.ref_get_pt_pairs.argtypes = [vp, vp, vp, vp, cl, cl, ci, ci, cd, ci, vp, vp, vp, vp, vp, vp]
            L.ref_align.restype = cd; L.ref_align.argtypes = [ci, cl, vp, vp, vp, vp, vp, vp]
            L.ref_m4inv.restype = ci; L.ref_m4inv.argtypes = [vp, vp]
            L.ref_mmult.restype = None; L.ref_mmult.argtypes = [vp, vp, vp]
            L.ref_euler_to_matrix4.restype = None; L.ref_euler_to_matrix4.argtypes = [vp, vp, vp]
            L.ref_matrix4_to_euler.restype = None; L.ref_matrix4_to_euler.argtypes = [vp, vp, vp]
            L.ref_match.restype = ci
            L.ref_match.argtypes = [vp, vp, vp, vp, cl, vp, vp, ci, ci, cd, ci, cd, ci, ci, vp, vp, vp, vp]
            L.ref_normals_knn.restype = None; L.ref_normals_knn.argtypes = [vp, cl, ci, vp, vp]
            L.ref_lum_link.restype = cl; L.ref_lum_link.argtypes = [vp, vp, vp, cl, cd, vp, vp]
            _ref[path] = L
    return _ref[path]


_full = []
_adapter_full = []


def _full_sigs(L):
    if True:
        if True:
            L.reff_scan_create.restype = vp
            L.reff_scan_create.argtypes = [vp, cl, vp, vp, cd, ci, ci, ci]
            L.reff_scan_create_normals.restype = vp
            L.reff_scan_create_normals.argtypes = [vp, vp, cl, vp, vp, cd, ci]
            L.reff_srand.restype = None; L.reff_srand.argtypes = [C.c_uint]
            L.reff_scan_free_all.restype = None; L.reff_scan_free_all.argtypes = [vp, ci]
            L.reff_scan_get.restype = cl; L.reff_scan_get.argtypes = [vp, C.c_char_p, vp, cl]
            L.reff_scan_pose.restype = None; L.reff_scan_pose.argtypes = [vp] * 5
            L.reff_scan_frames.restype = cl; L.reff_scan_frames.argtypes = [vp, vp, vp, cl]
            L.reff_scan_transform.restype = None; L.reff_scan_transform.argtypes = [vp, vp, ci, ci]
            L.reff_match.restype = ci; L.reff_match.argtypes = [vp, vp, ci, ci, cd, ci, cd, ci, ci]
            L.reff_do_icp.restype = ci; L.reff_do_icp.argtypes = [vp, ci, ci, ci, cd, ci, cd, ci, ci, ci, ci, ci]
            L.reff_covariance.restype = ci; L.reff_covariance.argtypes = [vp, vp, ci, ci, ci, cd, vp, vp]
            L.reff_graph_from_poses.restype = ci; L.reff_graph_from_poses.argtypes = [ci, cd, ci, vp, ci]
            L.reff_lum_euler.restype = cd; L.reff_lum_euler.argtypes = [vp, ci, vp, ci, ci, cd, cd, ci]
    return L


def full():
    """oracle/_ref/libref3dtk_full.so (the reference's own classes, see oracle/full_harness.cc) or None."""
    if not _full:
        _full.append(_full_sigs(C.CDLL(REF_FULL_SO)) if os.path.exists(REF_FULL_SO) else None)
    return _full[0]


def adapter_full():
    """oracle/_ref/libadapter3dtk_full.so: the same classes + the reference-side adapters (icp6D_gpu, GpuSearchTree)
    + the product library (oracle/adapter_full_harness.cc); None if absent.  Needs a GPU to do anything."""
    if not _adapter_full:
        if not os.path.exists(ADAPTER_FULL_SO):
            _adapter_full.append(None)
        else:
            L = _full_sigs(C.CDLL(ADAPTER_FULL_SO))
            L.reffa_last_error.restype = C.c_char_p
            L.reffa_scan_create_gputree.restype = vp
            L.reffa_scan_create_gputree.argtypes = [vp, cl, vp, vp, cd, ci, cd]
            L.reffa_match_gpu.restype = ci; L.reffa_match_gpu.argtypes = [vp, vp, ci, ci, cd, ci, cd, ci, ci, vp]
            L.reffa_do_icp_gpu.restype = ci; L.reffa_do_icp_gpu.argtypes = [vp, ci, ci, ci, cd, ci, cd, ci, ci, ci, ci]
            _adapter_full.append(L)
    return _adapter_full[0]


class FullRefScans:
    """A set of the reference's own in-memory BasicScans (scan-local points + pose), freed together.
    `with FullRefScans(locals_xyz, rpos, rpostheta, voxel=-1) as fr: ...`"""

    def __init__(self, locals_xyz, rpos, rpostheta, voxel=-1.0, nrpts=0, nns=0, bucket=20, lib=None, gpu_tree=False,
                 max_dist_hint=25.0, normals=None):
        """lib: full() (default) or adapter_full(); gpu_tree (adapter_full only): the scans' search tree is a
        GpuSearchTree instead of the k-d tree."""
        self.L = lib if lib is not None else full()
        self.h = (vp * len(locals_xyz))()
        for i, (x, p, t) in enumerate(zip(locals_xyz, rpos, rpostheta)):
            x = np.ascontiguousarray(x, dtype=np.float64)
            p = np.ascontiguousarray(p, dtype=np.float64); t = np.ascontiguousarray(t, dtype=np.float64)
            if normals is not None:
                nm = np.ascontiguousarray(normals[i], dtype=np.float64)
                self.h[i] = self.L.reff_scan_create_normals(P(x), P(nm), len(x), P(p), P(t), float(voxel), nrpts)
            elif gpu_tree:
                self.h[i] = self.L.reffa_scan_create_gputree(P(x), len(x), P(p), P(t), float(voxel), nrpts, max_dist_hint)
            else:
                self.h[i] = self.L.reff_scan_create(P(x), len(x), P(p), P(t), float(voxel), nrpts, nns, bucket)
        self.n = len(locals_xyz)

    def match_gpu(self, prev, cur, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5, rnd=1, anim=-1):
        """icp6D_gpu::match through an icp6D* -> (return value, iterations_run, npairs_last, kernel launches)"""
        out = (C.c_long * 3)()
        it = self.L.reffa_match_gpu(self.h[prev], self.h[cur], algo, mode, max_dist, max_iter, eps, rnd, anim, out)
        if it == -1000:
            raise RuntimeError(self.L.reffa_last_error().decode())
        return it, out[0], out[1], out[2]

    def do_icp_gpu(self, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5, meta=False, extrapolate_pose=True,
                   max_num_metascans=-1):
        rc = self.L.reffa_do_icp_gpu(self.h, self.n, algo, mode, max_dist, max_iter, eps, 1, int(meta),
                                     int(extrapolate_pose), max_num_metascans)
        if rc == -1000:
            raise RuntimeError(self.L.reffa_last_error().decode())
        return rc

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.L.reff_scan_free_all(self.h, self.n)

    def get(self, i, field="xyz reduced"):
        n = self.L.reff_scan_get(self.h[i], field.encode(), None, 0)
        out = np.zeros((n, 3))
        self.L.reff_scan_get(self.h[i], field.encode(), P(out), n)
        return out

    def pose(self, i):
        T, D, p, t = np.zeros(16), np.zeros(16), np.zeros(3), np.zeros(3)
        self.L.reff_scan_pose(self.h[i], P(T), P(D), P(p), P(t))
        return {"transmat": T, "dalignxf": D, "rpos": p, "rpostheta": t}

    def frames(self, i):
        n = self.L.reff_scan_frames(self.h[i], None, None, 0)
        mats, types = np.zeros((n, 16)), np.zeros(n, dtype=np.int32)
        self.L.reff_scan_frames(self.h[i], P(mats), types.ctypes.data, n)
        return mats, types

    def match(self, prev, cur, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5, rnd=1):
        return self.L.reff_match(self.h[prev], self.h[cur], algo, mode, max_dist, max_iter, eps, rnd, 0)

    def do_icp(self, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5, meta=False, extrapolate_pose=True,
               max_num_metascans=-1):
        return self.L.reff_do_icp(self.h, self.n, algo, mode, max_dist, max_iter, eps, 1, int(meta),
                                  int(extrapolate_pose), max_num_metascans, 0)

    def covariance(self, first, second, maxdist2, quat=False):
        dim = 7 if quat else 6
        Cm, CD = np.zeros((dim, dim)), np.zeros(dim)
        self.L.reff_covariance(self.h[first], self.h[second], int(quat), 0, 1, maxdist2, P(Cm), P(CD))
        return Cm, CD

    def graph_from_poses(self, cldist2, loopsize):
        cap = self.n * self.n + 1
        links = np.zeros((cap, 2), dtype=np.int32)
        m = self.L.reff_graph_from_poses(self.n, cldist2, loopsize, links.ctypes.data, cap)
        return links[:m].copy()

    def lum_euler(self, links, nr_it, max_dist_lum=25.0, eps_lum=0.5):
        links = np.ascontiguousarray(links, dtype=np.int32)
        return self.L.reff_lum_euler(self.h, self.n, links.ctypes.data, len(links), nr_it, max_dist_lum, eps_lum, 0)


def P(a):
    return None if a is None else a.ctypes.data


def identity():
    return np.eye(4).T.reshape(16).copy()


# ---- convenience wrappers used by several test files -------------------------------------------
class PortTree:
    def __init__(self, xyz, bucket=20):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self.h = port().orc_tree_create(P(self.xyz), len(self.xyz), bucket)

    def find_closest_batch(self, q, maxdist2):
        q = np.ascontiguousarray(q, dtype=np.float64)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q))
        port().orc_find_closest_batch(self.h, P(q), len(q), maxdist2, P(idx), P(d2))
        return idx, d2

    def get_pt_pairs(self, src_xf, data_xyz, data_nrm, maxdist2, mode):
        n = len(data_xyz)
        p1, p2, pn = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        idx = np.empty(n, np.int32)
        s, cm, cdv = np.zeros(1), np.zeros(3), np.zeros(3)
        k = port().orc_get_pt_pairs(self.h, P(src_xf), P(data_xyz), P(data_nrm), 0, n, maxdist2, mode,
                                    P(p1), P(p2), P(pn), P(idx), P(s), P(cm), P(cdv))
        return k, p1[:k], p2[:k], pn[:k], idx[:k], s[0], cm, cdv

    def __del__(self):
        try:
            port().orc_tree_free(self.h)
        except Exception:
            pass


def port_match(model_xyz, data_xyz, data_nrm=None, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5,
               model_dalignxf=None, napx_weighted=0):
    """icp6D::match serial arm on the restatement. Returns dict with transmat, rms, npairs, iterations."""
    tree = PortTree(model_xyz)
    d = np.ascontiguousarray(data_xyz, dtype=np.float64).copy()
    nrm = None if data_nrm is None else np.ascontiguousarray(data_nrm, dtype=np.float64).copy()
    T, D = identity(), identity()
    S = identity() if model_dalignxf is None else np.ascontiguousarray(model_dalignxf, dtype=np.float64)
    rms = np.zeros(max(max_iter, 1)); npairs = np.zeros(max(max_iter, 1), dtype=np.int64)
    done = C.c_int(0)
    it = port().orc_match(tree.h, P(S), P(d), P(nrm), len(d), P(T), P(D), algo, mode, max_dist, max_iter,
                          eps, napx_weighted, P(rms), P(npairs), C.byref(done))
    k = done.value
    return {"iterations": it, "iterations_run": k, "transmat": T, "dalignxf": D, "rms": rms[:k],
            "npairs": npairs[:k], "xyz": d, "nrm": nrm}


def ref_match(model_xyz, data_xyz, data_nrm=None, algo=1, mode=0, max_dist=25.0, max_iter=50, eps=1e-5,
              model_dalignxf=None, threads=0, omp=False):
    L = ref(omp)
    m = np.ascontiguousarray(model_xyz, dtype=np.float64)
    tree = L.ref_tree_create(P(m), len(m), 0, 20)
    d = np.ascontiguousarray(data_xyz, dtype=np.float64).copy()
    nrm = None if data_nrm is None else np.ascontiguousarray(data_nrm, dtype=np.float64).copy()
    T, D = identity(), identity()
    S = identity() if model_dalignxf is None else np.ascontiguousarray(model_dalignxf, dtype=np.float64)
    rms = np.zeros(max(max_iter, 1)); npairs = np.zeros(max(max_iter, 1), dtype=np.int64)
    done = C.c_int(0); ms = C.c_double(0)
    it = L.ref_match(tree, P(S), P(d), P(nrm), len(d), P(T), P(D), algo, mode, max_dist, max_iter, eps, 1,
                     threads, P(rms), P(npairs), C.byref(done), C.byref(ms))
    L.ref_tree_free(tree)
    k = done.value
    return {"iterations": it, "iterations_run": k, "transmat": T, "dalignxf": D, "rms": rms[:k],
            "npairs": npairs[:k], "xyz": d, "nrm": nrm, "ms_after_first": ms.value}


def port_lum_link(model_xyz, data_xyz, maxdist2, model_dalignxf=None):
    tree = PortTree(model_xyz)
    d = np.ascontiguousarray(data_xyz, dtype=np.float64)
    S = identity() if model_dalignxf is None else np.ascontiguousarray(model_dalignxf, dtype=np.float64)
    Cm, CD = np.zeros(36), np.zeros(6)
    m = port().orc_lum_link(tree.h, P(S), P(d), len(d), maxdist2, P(Cm), P(CD))
    return Cm.reshape(6, 6), CD, m


def do_icp(match_fn, scans_xyz, transmats_org, extrapolate_pose=True, meta=False, max_num_metascans=0, **kw):
    """icp6D::doICP (icp6D.cc:374-437) over in-memory scans, every match done by `match_fn` (port_match or
    ref_match -- the compiled reference).  scans_xyz: points in the global frame at load ("xyz reduced original");
    transmats_org: the poses they were loaded with.  Returns dict(transmats, dalignxfs, iterations, xyz)."""
    L = port()
    n = len(scans_xyz)
    org = [np.ascontiguousarray(t, dtype=np.float64).reshape(16) for t in transmats_org]
    cur = [np.ascontiguousarray(s, dtype=np.float64).copy() for s in scans_xyz]
    T = [t.copy() for t in org]
    D = [identity() for _ in range(n)]
    its = [0] * n

    def apply(M, i):          # Scan::transform: points, transMat, dalignxf
        R = M.reshape(4, 4).T
        cur[i] = cur[i] @ R[:3, :3].T + R[:3, 3]
        for arr in (T, D):
            out = np.zeros(16)
            L.orc_mmult(P(M), P(arr[i]), P(out))
            arr[i] = out

    meta_list = []
    for i in range(n):
        if i > 0:
            if extrapolate_pose:   # Scan::mergeCoordinatesWithRoboterPosition, scan.cc:826-833
                inv, delta = np.zeros(16), np.zeros(16)
                L.orc_m4inv(P(org[i - 1]), P(inv))
                L.orc_mmult(P(T[i - 1]), P(inv), P(delta))
                apply(delta, i)
            if meta:
                model, mdal = np.concatenate([cur[j] for j in meta_list], axis=0), identity()
            else:
                model, mdal = scans_xyz[i - 1], D[i - 1]
            r = match_fn(model, cur[i], model_dalignxf=mdal, **kw)
            its[i] = r["iterations"]
            cur[i] = r["xyz"]
            for arr in (T, D):
                out = np.zeros(16)
                L.orc_mmult(P(r["transmat"]), P(arr[i]), P(out))
                arr[i] = out
        if meta and i != n - 1:
            meta_list.append(i)
            if max_num_metascans > 0:
                meta_list = meta_list[-max_num_metascans:]
    return {"transmats": np.array(T), "dalignxfs": np.array(D), "iterations": its, "xyz": cur}


def port_graph_from_poses(rpos, cldist2, loopsize):
    rpos = np.ascontiguousarray(rpos, dtype=np.float64)
    cap = rpos.shape[0] * rpos.shape[0] + 1
    links = np.zeros((cap, 2), dtype=np.int32)
    m = port().orc_graph_from_poses(P(rpos), rpos.shape[0], cldist2, loopsize, links.ctypes.data, cap)
    return links[:m].copy()


def port_lum_graph_slam(scans_xyz, links, maxdist2, nr_it, eps, transmats, dalignxfs=None):
    """orc_lum_graph_slam over in-memory scans -> dict(iterations, ret, transmats, dalignxfs, G, B)"""
    n = len(scans_xyz)
    xyz = np.ascontiguousarray(np.concatenate(scans_xyz, axis=0), dtype=np.float64)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in scans_xyz])
    links = np.ascontiguousarray(links, dtype=np.int32)
    T = np.ascontiguousarray(transmats, dtype=np.float64).reshape(n, 16).copy()
    D = np.tile(identity(), (n, 1)) if dalignxfs is None else np.ascontiguousarray(dalignxfs, dtype=np.float64).reshape(n, 16).copy()
    dim = 6 * (n - 1)
    G, B, ret = np.zeros((dim, dim)), np.zeros(dim), np.zeros(1)
    it = port().orc_lum_graph_slam(n, P(xyz), off.ctypes.data, links.ctypes.data, links.shape[0], maxdist2, nr_it, eps,
                                   P(T), P(D), P(ret), P(G), P(B))
    return {"iterations": it, "ret": float(ret[0]), "transmats": T, "dalignxfs": D, "G": G, "B": B}


def octree_centres(pts, voxel):
    """Oracle for the `-r voxel` reduction (numpy restatement of include/slam6d/Boctree.h:224-270, :612-656,
    :928-949, :1164-1195, :268, :1784-1815): root cube = bbox centre, half-size = max half-extent + 1.0; child index bit
    k set iff not p[k] < centre[k] (the T** constructor's partition: a point ON a splitting plane goes up); a child is a leaf when ITS half-size <= voxel; output = leaf-cube centres, depth
    first, children 0..7.  Pinned: bit-identical to Scan::calcReducedPoints + BOctTree of the compiled reference
    (oracle/_ref/libref3dtk_full.so) -- tests/test_full_reference.py, golden tests/golden/full_vectors.npz."""
    import sys
    pts = np.asarray(pts, dtype=np.float64)
    mins, maxs = pts.min(0), pts.max(0)
    centre = 0.5 * (mins + maxs)
    size = float(np.max(0.5 * (maxs - mins))) + 1.0
    out = []

    def rec(idx, c, s):            # node with centre c, half-size s, points idx
        p = pts[idx]
        child = (~(p[:, 0] < c[0])).astype(np.int64) | ((~(p[:, 1] < c[1])).astype(np.int64) << 1) | \
                ((~(p[:, 2] < c[2])).astype(np.int64) << 2)
        for i in range(8):
            sel = idx[child == i]
            if len(sel) == 0:
                continue
            cc = c + (s / 2.0) * np.array([1 if i & 1 else -1, 1 if i & 2 else -1, 1 if i & 4 else -1], dtype=np.float64)
            if s / 2.0 <= voxel:
                out.append(cc)
            else:
                rec(sel, cc, s / 2.0)

    sys.setrecursionlimit(10000)
    rec(np.arange(len(pts)), centre, size)
    return np.array(out)


def rel_frobenius(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
