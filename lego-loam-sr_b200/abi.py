"""ctypes binding of the C ABI in include/legoloam_b200.h (lib/liblegoloam_b200.so).

This is the call path tests and bench.py use; it is the same boundary the reference's C++ node binds
(INTEGRATION.md). There is NO fallback: if the CUDA library is missing or no GPU is usable the calls
raise — nothing here computes on the CPU.

Point clouds are numpy float32 arrays of shape [n, 4] (packed x,y,z,intensity; stride 16) or [n, 8]
(pcl::PointXYZI layout: x,y,z,pad,intensity,pad,pad,pad; stride 32). Poses are 6 floats
[rx, ry, rz, tx, ty, tz].
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblegoloam_b200.so")

PROFILE_CLASSES = ("transform_minmax", "voxel_keys", "voxel_sort_pass", "voxel_reduce", "grid_keys", "grid_sort_pass",
                   "grid_finalize", "lm_begin", "lm_iter")

LL_OK = 0
STATUS_NAMES = {0: "LL_OK", -1: "LL_ERR_INVALID_ARG", -2: "LL_ERR_CUDA", -3: "LL_ERR_NOMEM", -4: "LL_ERR_NOT_FOUND",
                -5: "LL_ERR_CAPACITY", -6: "LL_ERR_GRID_TOO_LARGE", -7: "LL_ERR_STATE"}


class Params(C.Structure):
    _fields_ = [
        ("corner_leaf", C.c_float), ("surf_leaf", C.c_float), ("max_sq_dist", C.c_float),
        ("eig_ratio", C.c_float), ("plane_tol", C.c_double), ("weight_k", C.c_double),
        ("min_weight", C.c_double), ("min_residuals", C.c_int), ("eig_thresh", C.c_float),
        ("iter_count_thres", C.c_int), ("step_size", C.c_float), ("stop_thres", C.c_float),
        ("apply_update", C.c_int), ("eigen_mode", C.c_int), ("map_downsample", C.c_int),
        ("min_corner_map", C.c_int), ("min_surf_map", C.c_int), ("lm_solver", C.c_int), ("persistent_lm", C.c_int), ("lm_max_blocks", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("iter_num", C.c_int), ("min_lambda", C.c_float), ("cf_mean", C.c_float),
                ("is_degenerate", C.c_int), ("converged", C.c_int), ("sel_num", C.c_int), ("flags", C.c_int),
                ("corner_map_ds", C.c_int), ("surf_map_ds", C.c_int), ("pose", C.c_float * 6),
                ("last_x", C.c_float * 6)]


class FaResult(C.Structure):
    """ll_fa_result (include/legoloam_b200.h)."""
    _fields_ = [("transform_cur", C.c_float * 6), ("iters_surf", C.c_int), ("iters_corner", C.c_int), ("is_degenerate", C.c_int),
                ("skipped", C.c_int), ("n_sel_surf", C.c_int), ("n_sel_corner", C.c_int)]


class ScanClouds(C.Structure):
    """ll_scan_clouds: device copies of downsampleCurrentScan's inputs and results (valid until the next call)."""
    _fields_ = [("d_raw", C.c_void_p * 5), ("n_raw", C.c_size_t * 5), ("d_ds", C.c_void_p * 6), ("n_ds", C.c_size_t * 6),
                ("flags", C.c_int * 6)]


class LegoLoamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (STATUS_NAMES.get(code, "?"), code, msg))
        self.code = code


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.ll_last_error_string.restype = C.c_char_p
        lib.ll_last_error_string.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def default_params(**kw):
    p = Params()
    load().ll_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _cloud(a):
    """-> (contiguous float32 array, n, stride_bytes)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] not in (4, 8):
        raise ValueError("cloud must be [n, 4] or [n, 8] float32, got %s" % (a.shape,))
    return a, a.shape[0], a.shape[1] * 4


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _pose(p):
    p = np.ascontiguousarray(p, dtype=np.float32).reshape(6)
    return p


class Handle:
    def __init__(self, params=None, device=0):
        self._lib = load()
        self._h = C.c_void_p()
        self.params = params if params is not None else default_params()
        rc = self._lib.ll_create(C.byref(self.params), C.c_int(device), C.byref(self._h))
        if rc != LL_OK:
            raise LegoLoamError(rc, "ll_create failed (is a CUDA device visible?)")

    @classmethod
    def from_raw(cls, raw_handle, params=None):
        """Non-owning view of an ll_handle created elsewhere (e.g. inside the C++ host mirror)."""
        self = cls.__new__(cls)
        self._lib = load()
        self._h = C.c_void_p(raw_handle)
        self.params = params
        self._borrowed = True
        return self

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != LL_OK:
            raise LegoLoamError(rc, self._lib.ll_last_error_string(self._h).decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                self._lib.ll_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, params):
        self._check(self._lib.ll_set_params(self._h, C.byref(params)))
        self.params = params

    @property
    def stream(self):
        s = C.c_void_p()
        self._check(self._lib.ll_get_stream(self._h, C.byref(s)))
        return s.value or 0

    @property
    def launch_count(self):
        n = C.c_uint64(0)
        self._check(self._lib.ll_get_launch_count(self._h, C.byref(n)))
        return n.value

    def profile_enable(self, on=True):
        self._check(self._lib.ll_profile_enable(self._h, C.c_int(1 if on else 0)))

    def profile_read(self):
        """-> {class name: (total ms, launches)} since the last read (CUDA events on the handle's stream)."""
        ms = (C.c_double * len(PROFILE_CLASSES))()
        cnt = (C.c_uint64 * len(PROFILE_CLASSES))()
        self._check(self._lib.ll_profile_read(self._h, ms, cnt, C.c_int(len(PROFILE_CLASSES))))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(PROFILE_CLASSES)}

    # -- keyframes / local map -------------------------------------------------------------------
    def add_keyframe(self, kf_id, corner, surf, outlier, pose6):
        c, nc, s1 = _cloud(corner)
        s, ns, s2 = _cloud(surf)
        o, no, s3 = _cloud(outlier)
        strides = {st for st, n in ((s1, nc), (s2, ns), (s3, no)) if n}
        if len(strides) > 1:
            raise ValueError("the three clouds of a keyframe must share one layout")
        stride = strides.pop() if strides else 16
        self._check(self._lib.ll_add_keyframe(self._h, C.c_int(kf_id), _vp(c), C.c_size_t(nc), _vp(s), C.c_size_t(ns),
                                              _vp(o), C.c_size_t(no), C.c_size_t(stride), _vp(_pose(pose6))))

    def add_keyframe_dev(self, kf_id, d_corner, nc, d_surf, ns, d_outlier, no, pose6):
        self._check(self._lib.ll_add_keyframe_dev(self._h, C.c_int(kf_id), C.c_void_p(d_corner), C.c_size_t(nc),
                                                  C.c_void_p(d_surf), C.c_size_t(ns), C.c_void_p(d_outlier),
                                                  C.c_size_t(no), _vp(_pose(pose6))))

    def set_keyframe_pose(self, kf_id, pose6):
        self._check(self._lib.ll_set_keyframe_pose(self._h, C.c_int(kf_id), _vp(_pose(pose6))))

    def build_local_map(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        self._check(self._lib.ll_build_local_map(self._h, _vp(ids), C.c_int(len(ids))))

    def set_local_map_raw(self, corner, surf):
        c, nc, s1 = _cloud(corner)
        s, ns, s2 = _cloud(surf)
        if nc and ns and s1 != s2:
            raise ValueError("corner and surf clouds must share one layout")
        stride = s1 if nc else s2
        self._check(self._lib.ll_set_local_map_raw(self._h, _vp(c), C.c_size_t(nc), _vp(s), C.c_size_t(ns),
                                                   C.c_size_t(stride)))

    def get_map_ds(self, which):
        n = C.c_size_t(0)
        self._check(self._lib.ll_get_map_ds(self._h, C.c_int(which), None, C.c_size_t(0), C.c_size_t(16), C.byref(n)))
        out = np.empty((max(n.value, 1), 4), np.float32)
        self._check(self._lib.ll_get_map_ds(self._h, C.c_int(which), _vp(out), C.c_size_t(len(out)), C.c_size_t(16),
                                            C.byref(n)))
        return out[:n.value].copy()

    # -- a10 -----------------------------------------------------------------------------------------
    def scan2map(self, corner_q, surf_q, pose6):
        cq, nc, s1 = _cloud(corner_q)
        sq, ns, s2 = _cloud(surf_q)
        stride = s1 if nc else s2
        pose = _pose(pose6).copy()
        res = Result()
        self._check(self._lib.ll_scan2map(self._h, _vp(cq), C.c_size_t(nc), _vp(sq), C.c_size_t(ns), C.c_size_t(stride),
                                          _vp(pose), C.byref(res)))
        return pose, res

    def scan2map_async(self, corner_q, surf_q, pose6):
        cq, nc, s1 = _cloud(corner_q)
        sq, ns, s2 = _cloud(surf_q)
        stride = s1 if nc else s2
        self._check(self._lib.ll_scan2map_async(self._h, _vp(cq), C.c_size_t(nc), _vp(sq), C.c_size_t(ns),
                                                C.c_size_t(stride), _vp(_pose(pose6))))

    def scan2map_dev(self, d_cq, nc, d_sq, ns, pose6):
        pose = _pose(pose6).copy()
        res = Result()
        self._check(self._lib.ll_scan2map_dev(self._h, C.c_void_p(d_cq), C.c_size_t(nc), C.c_void_p(d_sq),
                                              C.c_size_t(ns), _vp(pose), C.byref(res)))
        return pose, res

    def scan2map_dev_async(self, d_cq, nc, d_sq, ns, pose6):
        self._check(self._lib.ll_scan2map_dev_async(self._h, C.c_void_p(d_cq), C.c_size_t(nc), C.c_void_p(d_sq),
                                                    C.c_size_t(ns), _vp(_pose(pose6))))

    def scan2map_wait(self):
        pose = np.zeros(6, np.float32)
        res = Result()
        self._check(self._lib.ll_scan2map_wait(self._h, _vp(pose), C.byref(res)))
        return pose, res

    # -- stage-level entry points -----------------------------------------------------------------
    def voxel_downsample(self, pts, leaf, want_member=True):
        p, n, stride = _cloud(pts)
        out = np.empty((max(n, 1), 4), np.float32)
        member = np.full(max(n, 1), -1, np.int32)
        n_out = C.c_size_t(0)
        flags = C.c_int(0)
        self._check(self._lib.ll_voxel_downsample(self._h, _vp(p), C.c_size_t(n), C.c_size_t(stride), C.c_float(leaf),
                                                  _vp(out), C.c_size_t(len(out)), C.c_size_t(16), C.byref(n_out),
                                                  _vp(member) if want_member else None, C.byref(flags)))
        return out[:n_out.value].copy(), member[:n], flags.value

    def downsample_scan(self, clouds, want_host=True):
        """downsampleCurrentScan in one call. clouds = (corner_last, surf_last, corner_scan, surf_scan, outlier_last);
        -> (six downsampled clouds or None, ScanClouds with the device pointers / counts / flags)."""
        packed = [_cloud(c) for c in clouds]
        assert len(packed) == 5 and len({st for _, _, st in packed}) == 1
        ptrs = (C.c_void_p * 5)(*[p.ctypes.data if n else None for p, n, _ in packed])
        ns = (C.c_size_t * 5)(*[n for _, n, _ in packed])
        dev = ScanClouds()
        outs = None
        out_ptrs, caps = None, None
        if want_host:
            cap = [packed[0][1], packed[1][1], packed[2][1], packed[3][1], packed[4][1], packed[1][1] + packed[4][1]]
            outs = [np.empty((max(c, 1), 4), np.float32) for c in cap]
            out_ptrs = (C.c_void_p * 6)(*[o.ctypes.data for o in outs])
            caps = (C.c_size_t * 6)(*[len(o) for o in outs])
        self._check(self._lib.ll_downsample_scan(self._h, ptrs, ns, C.c_size_t(packed[0][2]), C.byref(dev), out_ptrs, caps,
                                                 C.c_size_t(16)))
        if want_host:
            outs = [o[:dev.n_ds[k]].copy() for k, o in enumerate(outs)]
        return outs, dev

    def knn5(self, which, queries):
        q, nq, stride = _cloud(queries)
        idx = np.full((max(nq, 1), 5), -1, np.int32)
        d2 = np.zeros((max(nq, 1), 5), np.float32)
        accept = np.zeros(max(nq, 1), np.int32)
        self._check(self._lib.ll_knn5(self._h, C.c_int(which), _vp(q), C.c_size_t(nq), C.c_size_t(stride), _vp(idx),
                                      _vp(d2), _vp(accept)))
        return idx[:nq], d2[:nq], accept[:nq]

    # -- f-3: FeatureAssociation correspondence search ---------------------------------------------
    def fa_set_last(self, which, cloud):
        c, n, stride = _cloud(cloud)
        self._check(self._lib.ll_fa_set_last(self._h, C.c_int(which), _vp(c), C.c_size_t(n), C.c_size_t(stride)))

    def fa_find_correspondences(self, which, queries, transform_cur, it, nearest_feature_dist_sqr=25.0, ind=None):
        """-> (ind1, ind2, ind3, coeff [nq,4], keep [nq]); ind = the index arrays of an earlier call (it % 5 != 0)."""
        q, nq, stride = _cloud(queries)
        m = max(nq, 1)
        if ind is None:
            ind = [np.full(m, -1, np.int32) for _ in range(3)]
        ind = [np.ascontiguousarray(a, np.int32).copy() for a in ind]
        coeff = np.zeros((m, 4), np.float32)
        keep = np.zeros(m, np.int32)
        tc = np.ascontiguousarray(transform_cur, np.float32)
        self._check(self._lib.ll_fa_find_correspondences(self._h, C.c_int(which), _vp(q), C.c_size_t(nq), C.c_size_t(stride),
                                                         _vp(tc), C.c_int(it), C.c_float(nearest_feature_dist_sqr),
                                                         _vp(ind[0]), _vp(ind[1]), _vp(ind[2]), _vp(coeff), _vp(keep)))
        return ind[0][:nq], ind[1][:nq], ind[2][:nq], coeff[:nq], keep[:nq]

    def fa_update_transformation(self, corner_sharp, surf_flat, transform_cur, nearest_feature_dist_sqr=25.0, want_trace=False):
        """f-1 + f-3 in one launch: FeatureAssociation::updateTransformation. -> dict(transform_cur, iters_surf,
        iters_corner, is_degenerate, skipped, n_sel_surf, n_sel_corner[, trace_surf, trace_corner])."""
        cs = np.ascontiguousarray(corner_sharp, np.float32).reshape(-1, 4)
        ss = np.ascontiguousarray(surf_flat, np.float32).reshape(-1, 4)
        tc = np.ascontiguousarray(transform_cur, np.float32).copy()
        res = FaResult()
        trace = np.zeros((2, 100, 4), np.float32) if want_trace else None
        self._check(self._lib.ll_fa_update_transformation(self._h, _vp(cs), C.c_size_t(len(cs)), _vp(ss), C.c_size_t(len(ss)),
                                                          C.c_size_t(16), _vp(tc), C.c_float(nearest_feature_dist_sqr),
                                                          C.byref(res), _vp(trace) if want_trace else None))
        out = dict(transform_cur=tc, iters_surf=res.iters_surf, iters_corner=res.iters_corner, is_degenerate=res.is_degenerate,
                   skipped=res.skipped, n_sel_surf=res.n_sel_surf, n_sel_corner=res.n_sel_corner)
        if want_trace:
            out["trace_surf"] = trace[0, :min(res.iters_surf + 1, 100)]
            out["trace_corner"] = trace[1, :min(res.iters_corner + 1, 100)]
        return out

    def fa_reset_state(self):
        self._check(self._lib.ll_fa_reset_state(self._h))

    def _residuals(self, fn, queries, pose6):
        q, nq, stride = _cloud(queries)
        m = max(nq, 1)
        coeff = np.zeros((m, 4), np.float32)
        mask = np.zeros(m, np.int32)
        accept = np.zeros(m, np.int32)
        idx = np.full((m, 5), -1, np.int32)
        d2 = np.zeros((m, 5), np.float32)
        self._check(fn(self._h, _vp(q), C.c_size_t(nq), C.c_size_t(stride), _vp(_pose(pose6)), _vp(coeff), _vp(mask),
                       _vp(accept), _vp(idx), _vp(d2)))
        return dict(coeff=coeff[:nq], mask=mask[:nq], accept=accept[:nq], knn_idx=idx[:nq], knn_d2=d2[:nq])

    def corner_residuals(self, queries, pose6):
        return self._residuals(self._lib.ll_corner_residuals, queries, pose6)

    def surf_residuals(self, queries, pose6):
        return self._residuals(self._lib.ll_surf_residuals, queries, pose6)

    def lm_step(self, ori, coeff, it, pose6):
        o, n, stride = _cloud(ori)
        c = np.ascontiguousarray(coeff, np.float32).reshape(-1, 4)
        assert len(c) == n
        pose = _pose(pose6).copy()
        AtA = np.zeros((6, 6), np.float32)
        Atb = np.zeros(6, np.float32)
        x = np.zeros(6, np.float32)
        dR, dT, ret = C.c_float(0), C.c_float(0), C.c_int(0)
        self._check(self._lib.ll_lm_step(self._h, _vp(o), _vp(c), C.c_size_t(n), C.c_size_t(stride), C.c_int(it),
                                         _vp(pose), _vp(AtA), _vp(Atb), _vp(x), C.byref(dR), C.byref(dT), C.byref(ret)))
        return dict(ret=ret.value, pose=pose, AtA=AtA, Atb=Atb, x=x, deltaR=dR.value, deltaT=dT.value)

    def get_degeneracy(self):
        d = C.c_int(0)
        P = np.zeros((6, 6), np.float32)
        self._check(self._lib.ll_get_degeneracy(self._h, C.byref(d), _vp(P)))
        return d.value, P

    def reset_state(self):
        self._check(self._lib.ll_reset_state(self._h))

    def solve_normal_eq(self, rows6, b, dof_mask, it, eig_thresh, matP=None, is_degenerate=0):
        rows = np.ascontiguousarray(rows6, np.float32).reshape(-1, 6)
        bb = np.ascontiguousarray(b, np.float32).reshape(-1)
        m = bin(dof_mask & 63).count("1")
        x = np.zeros(m, np.float32)
        P = np.zeros((m, m), np.float32) if matP is None else np.ascontiguousarray(matP, np.float32).copy()
        deg = C.c_int(is_degenerate)
        self._check(self._lib.ll_solve_normal_eq(self._h, _vp(rows), _vp(bb), C.c_size_t(len(rows)), C.c_int(dof_mask),
                                                 C.c_int(it), C.c_float(eig_thresh), _vp(x), _vp(P), C.byref(deg)))
        return x, P, deg.value
