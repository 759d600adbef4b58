"""3dtk_b200 -- ctypes binding over the C ABI in include/b200icp.h (lib/libb200icp.so).

The package name starts with a digit, so import it with
    icp = importlib.import_module("3dtk_b200")
This binding is harness glue for tests/ and bench.py; the product is the shared library.  There is no
CPU fallback anywhere: if the library is missing the import raises, and without a CUDA device
`Context()` raises `B200ICPError` (ENODEV).

Names mirror the reference's interface for this path (3DTK, paths relative to its tree):
  Scan  <-> Scan + its SearchTree   (src/slam6d/scan.cc:285-306, basicScan.cc:702-728)
  Scan.find_closest / get_pt_pairs  <-> SearchTree::FindClosest / getPtPairs (searchTree.h:81-112)
  icp6D.match                        <-> icp6D::match (include/slam6d/icp6D.h:51-53)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ICP_LIB") or os.path.join(_HERE, "lib", "libb200icp.so")  # override: A/B builds

ALGO_QUAT, ALGO_SVD, ALGO_ORTHO, ALGO_DUAL, ALGO_HELIX, ALGO_APX, ALGO_NAPX = 1, 2, 3, 4, 5, 6, 10
CLOSEST_POINT, CLOSEST_PLANE_SIMPLE = 0, 2

E_NAMES = {0: "OK", -1: "EINVAL", -2: "ENODEV", -3: "ECUDA", -4: "ENOMEM", -5: "EEMPTY", -6: "ESTATE"}


class B200ICPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200icp %s (%d): %s" % (E_NAMES.get(code, "?"), code, msg))
        self.code = code


class MatchParams(C.Structure):
    _fields_ = [("algo", C.c_int), ("pairing_mode", C.c_int), ("max_dist_match", C.c_double),
                ("max_num_iterations", C.c_int), ("epsilon_icp", C.c_double), ("rnd", C.c_int),
                ("exact", C.c_int), ("profile", C.c_int), ("napx_weighted", C.c_int),
                ("sharded", C.c_int), ("reserved", C.c_int * 2)]


class MatchResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("iterations_run", C.c_int), ("rms_last", C.c_double),
                ("npairs_last", C.c_uint64), ("queries", C.c_uint64), ("nn_kernel_ms", C.c_double),
                ("solve_kernel_ms", C.c_double), ("kernel_launches", C.c_uint32),
                ("stage2_queries_last", C.c_uint32)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_size_t, C.c_void_p)   # b200icp_allreduce_fn


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "3dtk_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no pure-Python or CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, dp, sz, i32, f64 = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double
    sig = {
        "b200icp_create": (i32, [i32, C.POINTER(vp)]),
        "b200icp_destroy": (None, [vp]),
        "b200icp_set_stream": (i32, [vp, vp]),
        "b200icp_synchronize": (i32, [vp]),
        "b200icp_last_error": (C.c_char_p, []),
        "b200icp_version": (C.c_char_p, []),
        "b200icp_scan_create": (i32, [vp, dp, dp, sz, f64, f64, C.POINTER(vp)]),
        "b200icp_scan_create_device": (i32, [vp, dp, dp, sz, f64, f64, C.POINTER(vp)]),
        "b200icp_scan_destroy": (None, [vp, vp]),
        "b200icp_scan_size": (sz, [vp]),
        "b200icp_scan_grid_info": (i32, [vp, C.POINTER(i32 * 3), C.POINTER(f64), C.POINTER(C.c_uint64),
                                         C.POINTER(C.c_uint64)]),
        "b200icp_scan_get_pose": (i32, [vp, dp, dp]),
        "b200icp_scan_set_pose": (i32, [vp, dp, dp]),
        "b200icp_scan_transform": (i32, [vp, dp]),
        "b200icp_metascan_create": (i32, [vp, dp, i32, f64, f64, C.POINTER(vp)]),
        "b200icp_do_icp": (i32, [vp, dp, i32, C.POINTER(MatchParams), i32, i32, i32, dp, dp, vp]),
        "b200icp_last_poses": (i32, [vp, i32, dp]),
        "b200icp_read_uos": (i32, [C.c_char_p, C.POINTER(vp), C.POINTER(sz)]),
        "b200icp_read_pose": (i32, [C.c_char_p, dp, dp]),
        "b200icp_write_uos": (i32, [C.c_char_p, dp, sz, f64, i32]),
        "b200icp_free": (None, [vp]),
        "b200icp_frames_create": (vp, [i32]),
        "b200icp_frames_destroy": (None, [vp]),
        "b200icp_frames_add": (i32, [vp, i32, dp, i32]),
        "b200icp_frames_transform": (i32, [vp, i32, dp, i32, i32]),
        "b200icp_frames_count": (i32, [vp, i32]),
        "b200icp_frames_get": (i32, [vp, i32, i32, dp, C.POINTER(i32)]),
        "b200icp_frames_save": (i32, [vp, i32, C.c_char_p, i32]),
        "b200icp_frames_load": (i32, [vp, i32, C.c_char_p]),
        "b200icp_graph_read_net": (i32, [C.c_char_p, dp, i32, C.POINTER(i32), C.POINTER(i32)]),
        "b200icp_scan_download": (i32, [vp, vp, dp, dp]),
        "b200icp_find_closest": (i32, [vp, vp, dp, f64, C.POINTER(C.c_int64)]),
        "b200icp_nn_batch": (i32, [vp, vp, dp, dp, sz, dp, f64, i32, dp, dp, dp]),
        "b200icp_nn_batch_device": (i32, [vp, vp, dp, dp, sz, dp, f64, i32, dp, dp, dp]),
        "b200icp_align_pairs": (i32, [i32, sz, dp, dp, dp, dp, dp, dp, C.POINTER(f64)]),
        "b200icp_match": (i32, [vp, vp, vp, C.POINTER(MatchParams), dp, dp, C.POINTER(MatchResult)]),
        "b200icp_comm_create": (i32, [vp, i32, i32, dp]),
        "b200icp_comm_connect_ipc": (i32, [vp, i32, dp]),
        "b200icp_comm_connect_local": (i32, [vp, i32, C.POINTER(vp)]),
        "b200icp_comm_mailbox": (vp, [vp]),
        "b200icp_comm_destroy": (i32, [vp]),
        "b200icp_last_profile": (i32, [vp, i32, dp, dp, dp, dp]),
        "b200icp_lum_link": (i32, [vp, vp, vp, f64, dp, dp, C.POINTER(C.c_uint64)]),
        "b200icp_lum_link_quat": (i32, [vp, vp, vp, f64, dp, dp, C.POINTER(C.c_uint64)]),
        "b200icp_lum_seed_cache": (i32, [vp, sz]),
        "b200icp_graph_from_poses": (i32, [dp, i32, f64, i32, dp, i32, C.POINTER(i32)]),
        "b200icp_graph_chain": (i32, [i32, i32, dp, i32, C.POINTER(i32)]),
        "b200icp_lum_fill_gb": (i32, [vp, dp, i32, dp, i32, f64, dp, dp, dp]),
        "b200icp_lum_solve_update": (i32, [dp, i32, dp, dp, C.POINTER(f64), vp]),
        "b200icp_lum_graph_slam": (i32, [vp, dp, i32, dp, i32, f64, i32, f64, C.POINTER(f64), C.POINTER(i32), vp]),
        "b200icp_lum_graph_slam_sharded": (i32, [vp, dp, i32, dp, i32, f64, i32, f64, i32, i32, ALLREDUCE_FN, vp,
                                                 C.POINTER(f64), C.POINTER(i32), vp]),
        "b200icp_matrix4_to_euler": (None, [dp, dp, dp]),
        "b200icp_scan_calc_normals": (i32, [vp, vp, i32, dp]),
        "b200icp_normals_knn": (i32, [vp, dp, sz, i32, dp, dp]),
        "b200icp_reduce_octree_center": (i32, [vp, dp, sz, f64, dp, C.POINTER(sz)]),
        "b200icp_reduce_octree": (i32, [vp, dp, dp, sz, f64, i32, C.c_uint, sz, dp, dp, C.POINTER(sz)]),
        "b200icp_glibc_rand": (i32, [C.c_uint, sz, sz, dp]),
        "b200icp_synth_scene": (i32, [C.c_uint64, C.c_uint64, sz, f64, dp]),
        "b200icp_euler_to_matrix4": (None, [dp, dp, dp]),
        "b200icp_m4inv": (i32, [dp, dp]),
        "b200icp_mmult": (None, [dp, dp, dp]),
        "b200icp_transform_points": (None, [dp, dp, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()


def _check(rc):
    if rc != 0:
        raise B200ICPError(rc, lib.b200icp_last_error().decode())


def _f64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and (a.ndim == 0 or a.shape[-1] != shape_last):
        raise ValueError("expected trailing dimension %d" % shape_last)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data


# ---- host-only helpers (no GPU) -----------------------------------------------------------------
def synth_scene(geom_seed, sample_seed, n, noise_sigma=0.5):
    out = np.empty((n, 3), dtype=np.float64)
    _check(lib.b200icp_synth_scene(geom_seed, sample_seed, n, noise_sigma, _ptr(out)))
    return out


def euler_to_matrix4(rpos, rpostheta):
    m = np.empty(16)
    lib.b200icp_euler_to_matrix4(_ptr(_f64(rpos)), _ptr(_f64(rpostheta)), _ptr(m))
    return m


def m4inv(m):
    out = np.empty(16)
    ok = lib.b200icp_m4inv(_ptr(_f64(m).reshape(16)), _ptr(out))
    return out, ok


def mmult(a, b):
    out = np.empty(16)
    lib.b200icp_mmult(_ptr(_f64(a).reshape(16)), _ptr(_f64(b).reshape(16)), _ptr(out))
    return out


def transform_points(xf, xyz):
    out = _f64(xyz, 3).copy()
    lib.b200icp_transform_points(_ptr(_f64(xf).reshape(16)), _ptr(out), out.shape[0])
    return out


def align_pairs(algo, p1, p2, nrm=None, centroid_m=None, centroid_d=None):
    """icp6Dminimizer::Align on an explicit pair list (host arithmetic). -> (alignxf[16], rms)"""
    p1, p2 = _f64(p1, 3), _f64(p2, 3)
    nrm = None if nrm is None else _f64(nrm, 3)
    cm = _f64(p1.mean(0) if centroid_m is None else centroid_m)
    cd = _f64(p2.mean(0) if centroid_d is None else centroid_d)
    xf, rms = np.empty(16), C.c_double(0)
    _check(lib.b200icp_align_pairs(algo, p1.shape[0], _ptr(p1), _ptr(p2), _ptr(nrm), _ptr(cm), _ptr(cd),
                                   _ptr(xf), C.byref(rms)))
    return xf, rms.value


# ---- device objects ----------------------------------------------------------------------------
class Context:
    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        _check(lib.b200icp_create(device, C.byref(h)))
        self._h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream):
        _check(lib.b200icp_set_stream(self._h, C.c_void_p(cuda_stream)))

    def lum_seed_cache(self, limit_bytes):
        """drop the remembered LUM link neighbours and set their memory limit (0 disables seeding)"""
        _check(lib.b200icp_lum_seed_cache(self._h, int(limit_bytes)))

    def synchronize(self):
        _check(lib.b200icp_synchronize(self._h))

    # ---- query-sharded match
    def comm_create(self, rank, world):
        """-> 64-byte cudaIpc handle of this rank's mailbox (bytes)"""
        buf = C.create_string_buffer(64)
        _check(lib.b200icp_comm_create(self._h, rank, world, buf))
        return buf.raw

    def comm_connect_ipc(self, handles):
        blob = b"".join(handles)
        _check(lib.b200icp_comm_connect_ipc(self._h, len(handles), blob))

    @staticmethod
    def comm_connect_local(contexts):
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        for c in contexts:
            _check(lib.b200icp_comm_connect_local(c._h, len(contexts), arr))

    def close(self):
        if self._h:
            lib.b200icp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scan:
    """A scan resident on the device with its search grid (model and/or data role)."""

    def __init__(self, ctx, xyz=None, normals=None, cell_edge=0.0, max_dist_hint=0.0,
                 device_ptrs=None, n=None):
        self.ctx = ctx
        h = C.c_void_p()
        if device_ptrs is not None:
            dx, dn = device_ptrs
            _check(lib.b200icp_scan_create_device(ctx._h, C.c_void_p(dx), C.c_void_p(dn) if dn else None,
                                                  n, cell_edge, max_dist_hint, C.byref(h)))
        else:
            xyz = _f64(xyz, 3)
            normals = None if normals is None else _f64(normals, 3)
            _check(lib.b200icp_scan_create(ctx._h, _ptr(xyz), _ptr(normals), xyz.shape[0], cell_edge,
                                           max_dist_hint, C.byref(h)))
        self._h = h

    @classmethod
    def from_host_pointers(cls, ctx, xyz_ptr, nrm_ptr, n, cell_edge=0.0, max_dist_hint=0.0):
        """Raw host pointers (e.g. pinned torch tensors): the e2e path of bench.py."""
        self = cls.__new__(cls)
        self.ctx = ctx
        h = C.c_void_p()
        _check(lib.b200icp_scan_create(ctx._h, C.c_void_p(xyz_ptr), C.c_void_p(nrm_ptr) if nrm_ptr else None,
                                       n, cell_edge, max_dist_hint, C.byref(h)))
        self._h = h
        return self

    def __len__(self):
        return lib.b200icp_scan_size(self._h)

    def grid_info(self):
        dims, h, nc, no = (C.c_int * 3)(), C.c_double(), C.c_uint64(), C.c_uint64()
        _check(lib.b200icp_scan_grid_info(self._h, C.byref(dims), C.byref(h), C.byref(nc), C.byref(no)))
        return {"dims": tuple(dims), "cell_edge": h.value, "n_cells": nc.value, "n_occupied": no.value}

    def get_pose(self):
        t, d = np.empty(16), np.empty(16)
        _check(lib.b200icp_scan_get_pose(self._h, _ptr(t), _ptr(d)))
        return t, d

    def set_pose(self, transmat=None, dalignxf=None):
        t = None if transmat is None else _f64(transmat).reshape(16)
        d = None if dalignxf is None else _f64(dalignxf).reshape(16)
        _check(lib.b200icp_scan_set_pose(self._h, _ptr(t), _ptr(d)))

    def calc_normals(self, k, rpos):
        """Scan::calcNormals on the resident scan (k-NN PCA, oriented towards rpos); no host round trip"""
        r = _f64(rpos).reshape(3)
        _check(lib.b200icp_scan_calc_normals(self.ctx._h, self._h, int(k), _ptr(r)))

    def transform(self, alignxf):
        """Scan::transform bookkeeping (scan.cc:851-898): pose matrices only, points stay where they are"""
        a = _f64(alignxf).reshape(16)
        _check(lib.b200icp_scan_transform(self._h, _ptr(a)))

    @classmethod
    def metascan(cls, ctx, scans, cell_edge=0.0, max_dist_hint=0.0):
        """MetaScan(scans) (metaScan.cc:27-69): one search structure over the members' current positions"""
        self = cls.__new__(cls)
        self.ctx = ctx
        h = C.c_void_p()
        arr = (C.c_void_p * len(scans))(*[s._h for s in scans])
        _check(lib.b200icp_metascan_create(ctx._h, arr, len(scans), cell_edge, max_dist_hint, C.byref(h)))
        self._h = h
        return self

    def download(self, with_normals=False):
        n = len(self)
        xyz = np.empty((n, 3))
        nrm = np.empty((n, 3)) if with_normals else None
        _check(lib.b200icp_scan_download(self.ctx._h, self._h, _ptr(xyz), _ptr(nrm)))
        return (xyz, nrm) if with_normals else xyz

    # SearchTree::FindClosest
    def find_closest(self, p, maxdist2):
        p = _f64(p)
        idx = C.c_int64(-1)
        _check(lib.b200icp_find_closest(self.ctx._h, self._h, _ptr(p), maxdist2, C.byref(idx)))
        return idx.value

    # SearchTree::getPtPairs (index form)
    def nn_batch(self, q_xyz, maxdist2, source_alignxf=None, q_nrm=None, pairing_mode=CLOSEST_POINT,
                 want_d2=True):
        q = _f64(q_xyz, 3)
        nq = q.shape[0]
        nrm = None if q_nrm is None else _f64(q_nrm, 3)
        xf = None if source_alignxf is None else _f64(source_alignxf).reshape(16)
        idx = np.empty(nq, dtype=np.int32)
        d2 = np.empty(nq) if want_d2 else None
        sums = np.zeros(8)
        _check(lib.b200icp_nn_batch(self.ctx._h, self._h, _ptr(q), _ptr(nrm), nq, _ptr(xf), maxdist2,
                                    pairing_mode, _ptr(idx), _ptr(d2), _ptr(sums)))
        return idx, d2, sums

    def destroy(self):
        if getattr(self, "_h", None):
            lib.b200icp_scan_destroy(self.ctx._h if self.ctx._h else None, self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class icp6D:
    """Mirror of the reference's icp6D (include/slam6d/icp6D.h:27-156) for the accelerated path."""

    def __init__(self, ctx, algo=ALGO_QUAT, max_dist_match=25.0, max_num_iterations=50,
                 epsilon_icp=1e-7, rnd=1, exact=True, napx_weighted=False, profile=False, sharded=False):
        if max_dist_match < 0.0:
            raise ValueError("ERROR [ICP6D]: first parameter (max_dist_match) has to be >= 0")
        if max_num_iterations < 0:
            raise ValueError("ERROR [ICP6D]: second parameter (max_num_iterations) has to be >= 0")
        self.ctx = ctx
        self.params = MatchParams(algo=algo, pairing_mode=CLOSEST_POINT, max_dist_match=max_dist_match,
                                  max_num_iterations=max_num_iterations, epsilon_icp=epsilon_icp, rnd=rnd,
                                  exact=1 if exact else 0, profile=1 if profile else 0,
                                  napx_weighted=1 if napx_weighted else 0, sharded=1 if sharded else 0)

    def match(self, previous_scan, current_scan, pairing_mode=CLOSEST_POINT):
        """-> dict(iterations, rms[], npairs[], result struct).  Updates current_scan's pose."""
        p = self.params
        p.pairing_mode = pairing_mode
        n = max(p.max_num_iterations, 1)
        rms = np.zeros(n)
        npairs = np.zeros(n, dtype=np.uint64)
        res = MatchResult()
        _check(lib.b200icp_match(self.ctx._h, previous_scan._h, current_scan._h, C.byref(p), _ptr(rms),
                                 _ptr(npairs), C.byref(res)))
        k = res.iterations_run
        prof = None
        if k > 0:
            nn, sv, s2, se = np.zeros(k), np.zeros(k), np.zeros(k, dtype=np.uint32), np.zeros(k, dtype=np.uint32)
            lib.b200icp_last_profile(self.ctx._h, k, _ptr(nn), _ptr(sv), _ptr(s2), _ptr(se))
            prof = {"nn_ms": nn, "solve_ms": sv, "stage2": s2, "searches": se}
        return {"profile": prof, "iterations": res.iterations, "iterations_run": k, "rms": rms[:k].copy(),
                "npairs": npairs[:k].copy(), "result": res}


    def last_poses(self):
        """transMat of the data scan after every iteration of the last match (what Scan::transform saw)"""
        n = lib.b200icp_last_poses(self.ctx._h, 0, None)
        out = np.zeros((max(n, 1), 16))
        lib.b200icp_last_poses(self.ctx._h, n, _ptr(out))
        return out[:n]

    def doICP(self, all_scans, pairing_mode=CLOSEST_POINT, extrapolate_pose=True, meta=False, max_num_metascans=0,
              transmat_org=None, frames=None):
        """icp6D::doICP (icp6D.cc:374-437) -> iterations per scan.  Updates the scans' poses."""
        p = self.params
        p.pairing_mode = pairing_mode
        arr = (C.c_void_p * len(all_scans))(*[s._h for s in all_scans])
        org = None if transmat_org is None else _f64(transmat_org).reshape(-1)
        its = np.zeros(max(len(all_scans), 1), dtype=np.int32)
        _check(lib.b200icp_do_icp(self.ctx._h, arr, len(all_scans), C.byref(p), 1 if extrapolate_pose else 0,
                                  1 if meta else 0, int(max_num_metascans), _ptr(org), _ptr(its),
                                  frames._h if frames is not None else None))
        return its[:len(all_scans)]


def lum_link(ctx, first, second, max_dist_match2):
    """lum6DEuler::covarianceEuler for one graph link -> (C[6,6], CD[6], npairs)"""
    Cm, CD, n = np.zeros(36), np.zeros(6), C.c_uint64(0)
    _check(lib.b200icp_lum_link(ctx._h, first._h, second._h, max_dist_match2, _ptr(Cm), _ptr(CD), C.byref(n)))
    return Cm.reshape(6, 6), CD, n.value


def lum_link_quat(ctx, first, second, max_dist_match2):
    """lum6DQuat::covarianceQuat for one graph link -> (C[7,7], CD[7], npairs)"""
    Cm, CD, n = np.zeros(49), np.zeros(7), C.c_uint64(0)
    _check(lib.b200icp_lum_link_quat(ctx._h, first._h, second._h, max_dist_match2, _ptr(Cm), _ptr(CD), C.byref(n)))
    return Cm.reshape(7, 7), CD, n.value


FRAME_INVALID, FRAME_ICP, FRAME_ICPINACTIVE, FRAME_LUM, FRAME_ELCH = 0, 1, 2, 3, 4   # Scan::AlgoType


def read_uos(path):
    """ScanIO_uos point file -> (n, 3) float64"""
    p, n = C.c_void_p(), C.c_size_t(0)
    _check(lib.b200icp_read_uos(os.fsencode(path), C.byref(p), C.byref(n)))
    try:
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(max(n.value, 1) * 3,))[:3 * n.value]
        return out.reshape(-1, 3).copy()
    finally:
        lib.b200icp_free(p)


def write_uos(path, xyz, scale=1.0, fmt=0):
    """write_uos (scanio/writer.cc:146-178): fmt 0 "%lf", 1 high precision, 2 hex floats"""
    xyz = _f64(xyz, 3)
    _check(lib.b200icp_write_uos(os.fsencode(path), _ptr(xyz), xyz.shape[0], float(scale), int(fmt)))


def read_pose(path):
    """scanNNN.pose -> (rPos[3], rPosTheta[3] in radians)"""
    pos, th = np.zeros(3), np.zeros(3)
    _check(lib.b200icp_read_pose(os.fsencode(path), _ptr(pos), _ptr(th)))
    return pos, th


class Frames:
    """The m_frames lists of a set of scans (BasicScan::addFrame / saveFrames, basicScan.cc:902-936)."""

    def __init__(self, n_scans):
        self._h = C.c_void_p(lib.b200icp_frames_create(n_scans))
        self.n_scans = n_scans

    def add(self, scan, transmat, type_):
        t = _f64(transmat).reshape(16)
        _check(lib.b200icp_frames_add(self._h, scan, _ptr(t), type_))

    def transform(self, scan, transmats, type_, islum):
        t = _f64(transmats).reshape(-1)
        _check(lib.b200icp_frames_transform(self._h, scan, _ptr(t), type_, islum))

    def get(self, scan):
        """-> list of (transMat[16], type)"""
        out = []
        for k in range(lib.b200icp_frames_count(self._h, scan)):
            m, t = np.zeros(16), C.c_int(0)
            _check(lib.b200icp_frames_get(self._h, scan, k, _ptr(m), C.byref(t)))
            out.append((m, t.value))
        return out

    def save(self, scan, path, append=False):
        _check(lib.b200icp_frames_save(self._h, scan, os.fsencode(path), 1 if append else 0))

    def load(self, scan, path):
        """BasicScan::readFrames: replaces the scan's frames by the file's"""
        _check(lib.b200icp_frames_load(self._h, scan, os.fsencode(path)))

    def __del__(self):
        try:
            lib.b200icp_frames_destroy(self._h)
        except Exception:
            pass


def matrix4_to_euler(m):
    """Matrix4ToEuler (globals.icc:540-578) -> (rPos[3], rPosTheta[3])"""
    m = _f64(m).reshape(16)
    th, pos = np.empty(3), np.empty(3)
    lib.b200icp_matrix4_to_euler(_ptr(m), _ptr(th), _ptr(pos))
    return pos, th


def _scan_array(scans):
    return (C.c_void_p * len(scans))(*[s._h for s in scans])


class Graph:
    """Graph (src/slam6d/graph.cc): list of (from, to) links between scan numbers."""

    def __init__(self, links, n_scans):
        self.links = np.ascontiguousarray(links, dtype=np.int32).reshape(-1, 2)
        self.n_scans = int(n_scans)

    @classmethod
    def from_poses(cls, rpos, cldist2, loopsize):
        """Graph(int nodes, double cldist2, int loopsize), graph.cc:108-127"""
        rpos = _f64(rpos, 3)
        n = C.c_int(0)
        _check(lib.b200icp_graph_from_poses(_ptr(rpos), rpos.shape[0], cldist2, loopsize, None, 0, C.byref(n)))
        links = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        _check(lib.b200icp_graph_from_poses(_ptr(rpos), rpos.shape[0], cldist2, loopsize, _ptr(links), n.value,
                                            C.byref(n)))
        return cls(links[:n.value], rpos.shape[0])

    @classmethod
    def chain(cls, n_scans, loop=False):
        """Graph(int nScans, bool loop), graph.cc:76-105"""
        n = C.c_int(0)
        _check(lib.b200icp_graph_chain(n_scans, 1 if loop else 0, None, 0, C.byref(n)))
        links = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        _check(lib.b200icp_graph_chain(n_scans, 1 if loop else 0, _ptr(links), n.value, C.byref(n)))
        return cls(links[:n.value], n_scans)

    @classmethod
    def from_net_file(cls, path):
        """Graph(const string& netfile), graph.cc:52-74"""
        n, ns = C.c_int(0), C.c_int(0)
        _check(lib.b200icp_graph_read_net(os.fsencode(path), None, 0, C.byref(n), C.byref(ns)))
        links = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        _check(lib.b200icp_graph_read_net(os.fsencode(path), _ptr(links), n.value, C.byref(n), C.byref(ns)))
        return cls(links[:n.value], ns.value)

    def get_nr_links(self):
        return self.links.shape[0]


class lum6DEuler:
    """lum6DEuler (src/slam6d/lum6Deuler.cc): global relaxation over a graph of scans resident on the device."""

    def __init__(self, ctx, max_dist_match_lum=25.0, epsilon_lum=0.5):
        self.ctx = ctx
        self.max_dist_match2_lum = float(max_dist_match_lum) ** 2     # graphSlam6D.cc:60: squared once
        self.epsilon_lum = float(epsilon_lum)

    def fill_gb(self, graph, scans, G=None, B=None, link_subset=None):
        """FillGB3D over `link_subset` (indices into graph.links; default all) -> (G, B, npairs per link)"""
        dim = 6 * (len(scans) - 1)
        G = np.zeros((dim, dim)) if G is None else G
        B = np.zeros(dim) if B is None else B
        links = graph.links if link_subset is None else graph.links[list(link_subset)]
        links = np.ascontiguousarray(links, dtype=np.int32)
        npairs = np.zeros(max(links.shape[0], 1), dtype=np.uint64)
        _check(lib.b200icp_lum_fill_gb(self.ctx._h, _scan_array(scans), len(scans), _ptr(links), links.shape[0],
                                       self.max_dist_match2_lum, _ptr(G), _ptr(B), _ptr(npairs)))
        return G, B, npairs[:links.shape[0]]

    @staticmethod
    def solve_update(scans, G, B, frames=None):
        """X = G^-1 B, pose corrections, Scan::transformToEuler -> sum of position differences"""
        s = C.c_double(0.0)
        G, B = _f64(G), _f64(B)
        _check(lib.b200icp_lum_solve_update(_scan_array(scans), len(scans), _ptr(G), _ptr(B), C.byref(s),
                                            frames._h if frames is not None else None))
        return s.value

    def doGraphSlam6D(self, graph, scans, nr_it, frames=None):
        """lum6DEuler::doGraphSlam6D -> (ret, iterations run)"""
        ret, it = C.c_double(0.0), C.c_int(0)
        links = np.ascontiguousarray(graph.links, dtype=np.int32)
        _check(lib.b200icp_lum_graph_slam(self.ctx._h, _scan_array(scans), len(scans), _ptr(links), links.shape[0],
                                          self.max_dist_match2_lum, int(nr_it), self.epsilon_lum, C.byref(ret),
                                          C.byref(it), frames._h if frames is not None else None))
        return ret.value, it.value

    def doGraphSlam6D_sharded(self, graph, scans, nr_it, rank, world, allreduce, frames=None):
        """b200icp_lum_graph_slam_sharded: links round-robin over `world` ranks, `allreduce(buf)` sums the packed
        [G|B] numpy view in place over the ranks (called once per LUM iteration) -> (ret, iterations run)"""
        ret, it = C.c_double(0.0), C.c_int(0)
        links = np.ascontiguousarray(graph.links, dtype=np.int32)
        err = []

        def _cb(ptr, count, _user):
            try:
                allreduce(np.ctypeslib.as_array(ptr, shape=(count,)))
                return 0
            except Exception as ex:      # never let an exception cross the C frame
                err.append(ex)
                return 1
        cb = ALLREDUCE_FN(_cb)
        rc = lib.b200icp_lum_graph_slam_sharded(self.ctx._h, _scan_array(scans), len(scans), _ptr(links), links.shape[0],
                                                self.max_dist_match2_lum, int(nr_it), self.epsilon_lum, int(rank),
                                                int(world), cb, None, C.byref(ret), C.byref(it),
                                                frames._h if frames is not None else None)
        if err:
            raise err[0]
        _check(rc)
        return ret.value, it.value


def reduce_octree_center(ctx, xyz, voxel_size):
    """Scan::calcReducedPoints for `-r voxel_size` (octree voxel-centre reduction) -> reduced xyz"""
    xyz = _f64(xyz, 3)
    out = np.empty_like(xyz)
    m = C.c_size_t(0)
    _check(lib.b200icp_reduce_octree_center(ctx._h, _ptr(xyz), xyz.shape[0], voxel_size, _ptr(out), C.byref(m)))
    return out[:m.value].copy()


def reduce_octree(ctx, xyz, voxel_size, nrpts=0, normals=None, rand_seed=1, rand_skip=0):
    """Scan::calcReducedPoints with -O nrpts (0 centre, -1 average, 1 one random point per voxel) -> xyz[, normals]"""
    xyz = _f64(xyz, 3)
    nrm = None if normals is None else _f64(normals, 3)
    out = np.empty_like(xyz)
    out_n = None if nrm is None else np.empty_like(xyz)
    m = C.c_size_t(0)
    _check(lib.b200icp_reduce_octree(ctx._h, _ptr(xyz), _ptr(nrm), len(xyz), float(voxel_size), int(nrpts),
                                     int(rand_seed), int(rand_skip), _ptr(out), _ptr(out_n), C.byref(m)))
    return out[:m.value].copy() if nrm is None else (out[:m.value].copy(), out_n[:m.value].copy())


def glibc_rand(seed, count, skip=0):
    out = np.empty(count, dtype=np.int32)
    _check(lib.b200icp_glibc_rand(int(seed), int(skip), int(count), out.ctypes.data))
    return out


def normals_knn(ctx, xyz, k, rpos):
    xyz = _f64(xyz, 3)
    out = np.empty_like(xyz)
    _check(lib.b200icp_normals_knn(ctx._h, _ptr(xyz), xyz.shape[0], k, _ptr(_f64(rpos)), _ptr(out)))
    return out
