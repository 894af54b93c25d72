"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/_build/liboracle.so (the CPU restatement of the reference's scan-to-map
path, oracle.cpp) and, when present, oracle/_ref/libnanoflann_ref.so (the reference's own vendored
nanoflann compiled from /root/reference in place). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(lego-loam-sr_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libnanoflann_ref.so")

KNN_BRUTE, KNN_GRID, KNN_NANOFLANN = 0, 1, 2


class Params(C.Structure):
    _fields_ = [
        ("corner_leaf", C.c_float), ("surf_leaf", C.c_float), ("max_sq_dist", C.c_float),
        ("eig_ratio", C.c_float), ("plane_tol", C.c_double), ("weight_k", C.c_double),
        ("min_weight", C.c_double), ("min_residuals", C.c_int), ("eig_thresh", C.c_float),
        ("iter_count_thres", C.c_int), ("step_size", C.c_float), ("stop_thres", C.c_float),
        ("apply_update", C.c_int), ("eigen_mode", C.c_int), ("map_downsample", C.c_int),
        ("min_corner_map", C.c_int), ("min_surf_map", C.c_int),
    ]


class State(C.Structure):
    _fields_ = [("is_degenerate", C.c_int), ("matP", C.c_float * 36), ("min_lambda", C.c_float),
                ("cf_mean", C.c_float)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iter_num", C.c_int), ("min_lambda", C.c_float),
                ("cf_mean", C.c_float), ("is_degenerate", C.c_int), ("converged", C.c_int),
                ("sel_num", C.c_int), ("pose", C.c_float * 6), ("last_x", C.c_float * 6)]


def build(force=False):
    """Compile the restatement (always) and the nanoflann shim (when /root/reference exists)."""
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in ("oracle.cpp", "oracle.h", "eigen_restate.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/LeGO-LOAM/include/lego_loam") and (
            force or not os.path.exists(_REF) or os.path.getmtime(os.path.join(_HERE, "nanoflann_ref.cpp")) > os.path.getmtime(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None
_keep = []


def _p4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, a.shape
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib, _ref
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_replay_create.restype = C.c_void_p
        _lib.orc_replay_create2.restype = C.c_void_p
        _lib.orc_lm_step.restype = C.c_int
        _lib.orc_voxel_downsample.restype = C.c_int
        if os.path.exists(_REF):
            _ref = C.CDLL(_REF)
            _ref.ref_kdtree_build.restype = C.c_void_p
            _lib.orc_set_ext_knn(C.cast(_ref.ref_kdtree_build, C.c_void_p),
                                 C.cast(_ref.ref_kdtree_knn5, C.c_void_p),
                                 C.cast(_ref.ref_kdtree_free, C.c_void_p))
    return _lib


def have_ref():
    lib()
    return _ref is not None


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        assert hasattr(p, k), k
        setattr(p, k, v)
    return p


def new_state():
    s = State()
    lib().orc_init_state(C.byref(s))
    return s


def set_knn_backend(kind):
    if kind == KNN_NANOFLANN and not have_ref():
        raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not built")
    lib().orc_set_knn_backend(C.c_int(kind))


def set_radius_backend(use_nanoflann):
    """Radius search of the default local-map branch (mapOptmization.cpp:1157): brute force (ties by index) or the
    reference's own nanoflann radiusSearch (oracle/_ref)."""
    if use_nanoflann:
        if not have_ref() or not hasattr(_ref, "ref_kdtree_radius"):
            raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not built (or predates ref_kdtree_radius)")
        lib().orc_set_ext_radius(C.cast(_ref.ref_kdtree_radius, C.c_void_p))
    else:
        lib().orc_set_ext_radius(C.c_void_p(0))


def select_surrounding(key_poses, cur_pos, radius, existing, use_nanoflann=False):
    """The id selection of extractSurroundingKeyFrames' default branch (:1152-1216) -> new surroundingExistingKeyPosesID."""
    kp = _p4(key_poses)
    cap = len(kp) + len(existing) + 1
    ex = np.zeros(cap, np.int32)
    ex[:len(existing)] = existing
    cur = np.ascontiguousarray(cur_pos, np.float32)
    if use_nanoflann:
        set_radius_backend(True)
    try:
        n = lib().orc_select_surrounding(_ptr(kp), C.c_int(len(kp)), _ptr(cur), C.c_float(radius), _ptr(ex), C.c_int(len(existing)),
                                         C.c_int(cap), C.c_int(1 if use_nanoflann else 0))
    finally:
        if use_nanoflann:
            set_radius_backend(False)
    return ex[:n].copy()


def transform_cloud(pts, pose6):
    pts = _p4(pts)
    out = np.empty_like(pts)
    pose = np.ascontiguousarray(pose6, dtype=np.float32)
    lib().orc_transform_cloud(_ptr(pts), C.c_int(len(pts)), _ptr(pose), _ptr(out))
    return out


def voxel_downsample(pts, leaf):
    pts = _p4(pts)
    n = len(pts)
    out = np.empty((max(n, 1), 4), np.float32)
    member = np.empty(max(n, 1), np.int32)
    n_out = C.c_int(0)
    flag = lib().orc_voxel_downsample(_ptr(pts), C.c_int(n), C.c_float(leaf), _ptr(out), C.byref(n_out),
                                      _ptr(member))
    return out[:n_out.value].copy(), member[:n], flag


def knn5(map_pts, q, backend=KNN_GRID):
    map_pts, q = _p4(map_pts), _p4(q)
    if backend == KNN_NANOFLANN and not have_ref():
        raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not built")
    idx = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    lib().orc_knn5(_ptr(map_pts), C.c_int(len(map_pts)), _ptr(q), C.c_int(len(q)), _ptr(idx), _ptr(d2),
                   C.c_int(backend))
    return idx, d2


def _residuals(fn, map_pts, q, pose6, prm):
    map_pts, q = _p4(map_pts), _p4(q)
    nq = len(q)
    pose = np.ascontiguousarray(pose6, dtype=np.float32)
    coeff = np.zeros((nq, 4), np.float32)
    mask = np.zeros(nq, np.int32)
    accept = np.zeros(nq, np.int32)
    idx = np.zeros((nq, 5), np.int32)
    d2 = np.zeros((nq, 5), np.float32)
    fn(_ptr(map_pts), C.c_int(len(map_pts)), _ptr(q), C.c_int(nq), _ptr(pose), C.byref(prm), _ptr(coeff),
       _ptr(mask), _ptr(accept), _ptr(idx), _ptr(d2))
    return dict(coeff=coeff, mask=mask, accept=accept, knn_idx=idx, knn_d2=d2)


def corner_residuals(map_pts, q, pose6, prm):
    return _residuals(lib().orc_corner_residuals, map_pts, q, pose6, prm)


def surf_residuals(map_pts, q, pose6, prm):
    return _residuals(lib().orc_surf_residuals, map_pts, q, pose6, prm)


def lm_step(ori, coeff, it, pose6, prm, st):
    ori = _p4(ori)
    coeff = _p4(coeff)
    pose = np.array(pose6, dtype=np.float32)
    AtA = np.zeros((6, 6), np.float32)
    Atb = np.zeros(6, np.float32)
    x = np.zeros(6, np.float32)
    dR, dT = C.c_float(0), C.c_float(0)
    r = lib().orc_lm_step(_ptr(ori), _ptr(coeff), C.c_int(len(ori)), C.c_int(it), _ptr(pose), C.byref(prm),
                          C.byref(st), _ptr(AtA), _ptr(Atb), _ptr(x), C.byref(dR), C.byref(dT))
    return dict(ret=r, pose=pose, AtA=AtA, Atb=Atb, x=x, deltaR=dR.value, deltaT=dT.value)


def scan2map(corner_map, surf_map, corner_q, surf_q, pose6, prm, st=None):
    corner_map, surf_map, corner_q, surf_q = map(_p4, (corner_map, surf_map, corner_q, surf_q))
    pose = np.array(pose6, dtype=np.float32)
    st = st if st is not None else new_state()
    res = Result()
    lib().orc_scan2map(_ptr(corner_map), C.c_int(len(corner_map)), _ptr(surf_map), C.c_int(len(surf_map)),
                       _ptr(corner_q), C.c_int(len(corner_q)), _ptr(surf_q), C.c_int(len(surf_q)), _ptr(pose),
                       C.byref(prm), C.byref(st), C.byref(res))
    return res, pose, st


def _pack_clouds(clouds):
    off = np.zeros(len(clouds) + 1, np.int32)
    for i, c in enumerate(clouds):
        off[i + 1] = off[i] + len(c)
    if off[-1] == 0:
        return np.zeros((1, 4), np.float32), off
    return np.ascontiguousarray(np.concatenate([_p4(c) for c in clouds if len(c)]), np.float32), off


def build_local_map(poses6, corner, surf, outlier, prm):
    """poses6: [n_kf, 6]; corner/surf/outlier: lists of [n,4] clouds in the sensor frame."""
    n_kf = len(poses6)
    poses = np.ascontiguousarray(poses6, np.float32).reshape(-1)
    c, co = _pack_clouds(corner)
    s, so = _pack_clouds(surf)
    o, oo = _pack_clouds(outlier)
    cap_c = max(int(co[-1]), 1)
    cap_s = max(int(so[-1] + oo[-1]), 1)
    cds = np.empty((cap_c, 4), np.float32)
    sds = np.empty((cap_s, 4), np.float32)
    ncd, nsd, ncr, nsr = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    lib().orc_build_local_map(C.c_int(n_kf), _ptr(poses), _ptr(c), _ptr(co), _ptr(s), _ptr(so), _ptr(o), _ptr(oo),
                              C.byref(prm), _ptr(cds), C.byref(ncd), _ptr(sds), C.byref(nsd), C.byref(ncr),
                              C.byref(nsr))
    return cds[:ncd.value].copy(), sds[:nsd.value].copy(), ncr.value, nsr.value


def transform_associate_to_map(tsum, bef, aft):
    a = [np.ascontiguousarray(v, np.float32) for v in (tsum, bef, aft)]
    out = np.zeros(6, np.float32)
    lib().orc_transform_associate_to_map(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(out))
    return out


def self_adjoint_eig(A):
    A = np.ascontiguousarray(A, np.float32)
    n = A.shape[0]
    assert A.shape == (n, n) and n in (3, 6)
    ev = np.zeros(n, np.float32)
    V = np.zeros((n, n), np.float32)
    lib().orc_test_self_adjoint_eig(_ptr(A), C.c_int(n), _ptr(ev), _ptr(V))
    return ev, V


def colpiv_qr_solve(A, b):
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert A.shape in ((5, 3), (6, 6))
    x = np.zeros(A.shape[1], np.float32)
    lib().orc_test_colpiv_qr_solve(_ptr(A), _ptr(b), C.c_int(A.shape[0]), C.c_int(A.shape[1]), _ptr(x))
    return x


def inverse6(A):
    A = np.ascontiguousarray(A, np.float32)
    out = np.zeros((6, 6), np.float32)
    lib().orc_test_inverse6(_ptr(A), _ptr(out))
    return out


def solve_normal_eq(rows6, b, dof_mask, it, eig_thresh, eigen_mode, matP=None, is_degenerate=0):
    rows = np.ascontiguousarray(rows6, np.float32).reshape(-1, 6)
    bb = np.ascontiguousarray(b, np.float32).reshape(-1)
    m = bin(dof_mask & 63).count("1")
    x = np.zeros(m, np.float32)
    P = np.zeros((m, m), np.float32) if matP is None else np.ascontiguousarray(matP, np.float32).copy()
    deg = C.c_int(is_degenerate)
    lib().orc_solve_normal_eq.restype = C.c_int
    rc = lib().orc_solve_normal_eq(_ptr(rows), _ptr(bb), C.c_int(len(rows)), C.c_int(dof_mask), C.c_int(it),
                                   C.c_float(eig_thresh), C.c_int(eigen_mode), _ptr(x), _ptr(P), C.byref(deg))
    assert rc == 0, "oracle supports 3 or 6 unknowns"
    return x, P, deg.value


class Replay:
    def __init__(self, prm, window=50, loop_closure_enabled=True, search_radius=50.0):
        self._h = C.c_void_p(lib().orc_replay_create2(C.byref(prm), C.c_int(window), C.c_int(1 if loop_closure_enabled else 0),
                                                      C.c_float(search_radius)))

    def set_params(self, prm):
        lib().orc_replay_set_params(self._h, C.byref(prm))

    def map_ids(self, cap=100000):
        ids = np.zeros(cap, np.int32)
        n = lib().orc_replay_get_map_ids(self._h, _ptr(ids), C.c_int(cap))
        return ids[:n].copy()

    def step(self, odom6, corner_scan, corner_scan_ds, surf_last_ds, outlier_last_ds, surf_total_ds):
        a = [_p4(v) for v in (corner_scan, corner_scan_ds, surf_last_ds, outlier_last_ds, surf_total_ds)]
        odom = np.ascontiguousarray(odom6, np.float32)
        res = Result()
        nc, ns = C.c_int(0), C.c_int(0)
        lib().orc_replay_step(self._h, _ptr(odom), _ptr(a[0]), C.c_int(len(a[0])), _ptr(a[1]), C.c_int(len(a[1])),
                              _ptr(a[2]), C.c_int(len(a[2])), _ptr(a[3]), C.c_int(len(a[3])), _ptr(a[4]),
                              C.c_int(len(a[4])), C.byref(res), C.byref(nc), C.byref(ns))
        return res, nc.value, ns.value

    def get_map(self, which, cap=4_000_000):
        out = np.empty((cap, 4), np.float32)
        n = C.c_int(0)
        lib().orc_replay_get_map(self._h, C.c_int(which), _ptr(out), C.c_int(cap), C.byref(n))
        return out[:n.value].copy()

    def close(self):
        if self._h:
            lib().orc_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fa_correspondences(which, q, last, transform_cur, it, nearest_feature_dist_sqr=25.0, ind=None, use_nanoflann=False):
    """f-3: FeatureAssociation::findCorrespondingCornerFeatures (which = 0) / SurfFeatures (which = 1).
    ind = (ind1, ind2, ind3) from an earlier iteration (needed when it % 5 != 0). -> (ind1, ind2, ind3, coeff, keep)."""
    q, last = _p4(q), _p4(last)
    nq = len(q)
    m = max(nq, 1)
    if ind is None:
        ind = [np.full(m, -1, np.int32) for _ in range(3)]
    ind = [np.ascontiguousarray(a, np.int32).copy() for a in ind]
    coeff = np.zeros((m, 4), np.float32)
    keep = np.zeros(m, np.int32)
    tc = np.ascontiguousarray(transform_cur, np.float32)
    if use_nanoflann and not have_ref():
        raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not built")
    lib().orc_fa_correspondences(C.c_int(which), _ptr(q), C.c_int(nq), _ptr(last), C.c_int(len(last)), _ptr(tc), C.c_int(it),
                                 C.c_float(nearest_feature_dist_sqr), _ptr(ind[0]), _ptr(ind[1]), _ptr(ind[2]), _ptr(coeff),
                                 _ptr(keep), C.c_int(1 if use_nanoflann else 0))
    return ind[0][:nq], ind[1][:nq], ind[2][:nq], coeff[:nq], keep[:nq]


def fa_transform_to_start(pts, transform_cur):
    pts = _p4(pts)
    out = np.empty_like(pts)
    tc = np.ascontiguousarray(transform_cur, np.float32)
    lib().orc_fa_transform_to_start(_ptr(pts), C.c_int(len(pts)), _ptr(tc), _ptr(out))
    return out


class FaState(C.Structure):
    _fields_ = [("is_degenerate", C.c_int), ("matP_surf", C.c_float * 9), ("matP_corner", C.c_float * 9)]


def fa_update_transformation(corner_sharp, surf_flat, corner_last, surf_last, transform_cur, nearest_feature_dist_sqr=25.0,
                             eigen_mode=0, use_nanoflann=False, state=None):
    """f-1 + f-3: FeatureAssociation::updateTransformation (featureAssociation.cpp:2505-2535). -> dict(transform_cur,
    iters_surf, iters_corner, skipped, is_degenerate, trace_surf [iters][4], trace_corner [iters][4], state)."""
    cs, sf, cl, sl = _p4(corner_sharp), _p4(surf_flat), _p4(corner_last), _p4(surf_last)
    tc = np.ascontiguousarray(transform_cur, np.float32).copy()
    st = state if state is not None else FaState()
    if state is None:
        lib().orc_fa_init_state(C.byref(st))
    iters = (C.c_int * 2)(0, 0)
    tr_s = np.zeros((100, 4), np.float32)
    tr_c = np.zeros((100, 4), np.float32)
    if use_nanoflann and not have_ref():
        raise RuntimeError("oracle/_ref/libnanoflann_ref.so is not built")
    lib().orc_fa_update_transformation.restype = C.c_int
    skipped = lib().orc_fa_update_transformation(_ptr(cs), C.c_int(len(cs)), _ptr(sf), C.c_int(len(sf)), _ptr(cl), C.c_int(len(cl)),
                                                 _ptr(sl), C.c_int(len(sl)), _ptr(tc), C.c_float(nearest_feature_dist_sqr),
                                                 C.c_int(eigen_mode), C.c_int(1 if use_nanoflann else 0), C.byref(st), iters,
                                                 _ptr(tr_s), _ptr(tr_c))
    return dict(transform_cur=tc, iters_surf=iters[0], iters_corner=iters[1], skipped=int(skipped),
                is_degenerate=int(st.is_degenerate), trace_surf=tr_s[:min(iters[0] + 1, 100)],
                trace_corner=tr_c[:min(iters[1] + 1, 100)], state=st)
