"""ctypes driver of the C++ MapOptimization mirror (host/map_optimization.cpp, lib/liblegoloam_host.so).

The mirror keeps the reference's member names and per-frame call order and talks to the GPU through the
C ABI only; this module lets Python tests and bench.py feed it frames (trajectory replay, BASELINE
config 2). No CPU fallback: construction raises if the CUDA library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "liblegoloam_host.so")

_lib = None


def load():
    global _lib
    if _lib is None:
        abi.load()  # fails loudly if the CUDA library is missing
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("%s is missing: run __graft_entry__.build()" % HOST_LIB_PATH)
        lib = C.CDLL(HOST_LIB_PATH)
        lib.llh_create.restype = C.c_void_p
        lib.llh_abi_handle.restype = C.c_void_p
        _lib = lib
    return _lib


def _p4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("cloud must be [n, 4] float32")
    return a


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def transform_associate_to_map(tsum, bef, aft):
    a = [np.ascontiguousarray(v, np.float32) for v in (tsum, bef, aft)]
    out = np.zeros(6, np.float32)
    load().llh_transform_associate_to_map(_vp(a[0]), _vp(a[1]), _vp(a[2]), _vp(out))
    return out


class MapOptimization:
    """One MapOptimization node's hot path: feed frames with process_frame()."""

    def __init__(self, params=None, device=0, loop_closure_enabled=True, search_num=50, search_radius=50.0):
        self._lib = load()
        self.params = params if params is not None else abi.default_params()
        err = C.create_string_buffer(512)
        self._mo = C.c_void_p(self._lib.llh_create(C.byref(self.params), C.c_int(device),
                                                   C.c_int(1 if loop_closure_enabled else 0), C.c_int(search_num),
                                                   C.c_float(search_radius), err, C.c_size_t(512)))
        if not self._mo:
            raise RuntimeError(err.value.decode("utf-8", "replace") or "llh_create failed")

    def process_frame(self, odom6, frame, pre_downsampled=True):
        """frame: dict from synth.frame_inputs (pre_downsampled=True) or the raw clouds of synth.scan."""
        odom = np.ascontiguousarray(odom6, np.float32)
        z = np.zeros((0, 4), np.float32)
        if pre_downsampled:
            cs, csd = _p4(frame["corner_scan"]), _p4(frame["corner_scan_ds"])
            sl, ol, st = _p4(frame["surf_last_ds"]), _p4(frame["outlier_last_ds"]), _p4(frame["surf_total_ds"])
        else:
            cs, csd = _p4(frame["corner_scan"]), z
            sl, ol, st = _p4(frame["surf_last"]), _p4(frame["outlier_last"]), z
        res = abi.Result()
        pose = np.zeros(6, np.float32)
        err = C.create_string_buffer(512)
        rc = self._lib.llh_process_frame(self._mo, _vp(odom), _vp(cs), C.c_size_t(len(cs)), _vp(csd), C.c_size_t(len(csd)),
                                         _vp(sl), C.c_size_t(len(sl)), _vp(ol), C.c_size_t(len(ol)), _vp(st),
                                         C.c_size_t(len(st)), C.c_int(1 if pre_downsampled else 0), C.byref(res),
                                         _vp(pose), err, C.c_size_t(512))
        if rc != 0:
            raise RuntimeError(err.value.decode("utf-8", "replace"))
        return pose, res

    def process_frame_dev(self, odom6, ptrs):
        """ptrs: five (device_pointer, count) pairs — corner_scan, corner_scan_ds, surf_last_ds, outlier_last_ds,
        surf_total_ds — packed float4 clouds already resident on the device."""
        return self._process_view(odom6, ptrs, device=1, stride=16)

    def process_frame_host(self, odom6, clouds):
        """clouds: five numpy arrays in pcl::PointXYZI layout ([n, 8] float32, stride 32) or packed ([n, 4]),
        read in place by the library (no intermediate PointCloud copies) — the ROS node's situation."""
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        stride = arrs[0].shape[1] * 4 if arrs[0].ndim == 2 else 16
        ptrs = [(a.ctypes.data if a.size else 0, a.shape[0]) for a in arrs]
        return self._process_view(odom6, ptrs, device=0, stride=stride)

    def _process_view(self, odom6, ptrs, device, stride):
        odom = np.ascontiguousarray(odom6, np.float32)
        res = abi.Result()
        pose = np.zeros(6, np.float32)
        err = C.create_string_buffer(512)
        args = []
        for p, n in ptrs:
            args += [C.c_void_p(p), C.c_size_t(n)]
        rc = self._lib.llh_process_frame_view(self._mo, _vp(odom), *args, C.c_int(device), C.c_size_t(stride),
                                              C.byref(res), _vp(pose), err, C.c_size_t(512))
        if rc != 0:
            raise RuntimeError(err.value.decode("utf-8", "replace"))
        return pose, res

    @property
    def num_keyframes(self):
        return self._lib.llh_num_keyframes(self._mo)

    def key_pose(self, i):
        p = np.zeros(6, np.float32)
        self._lib.llh_get_key_pose(self._mo, C.c_int(i), _vp(p))
        return p

    @property
    def last_prior(self):
        p = np.zeros(6, np.float32)
        self._lib.llh_get_last_prior(self._mo, _vp(p))
        return p

    def abi_handle(self):
        """Non-owning abi.Handle view of the ll_handle the mirror drives (for inspection in tests)."""
        return abi.Handle.from_raw(self._lib.llh_abi_handle(self._mo), self.params)

    @property
    def launch_count(self):
        n = C.c_uint64(0)
        abi.load().ll_get_launch_count(C.c_void_p(self._lib.llh_abi_handle(self._mo)), C.byref(n))
        return n.value

    @property
    def stream(self):
        s = C.c_void_p()
        abi.load().ll_get_stream(C.c_void_p(self._lib.llh_abi_handle(self._mo)), C.byref(s))
        return s.value or 0

    def close(self):
        if getattr(self, "_mo", None):
            self._lib.llh_destroy(self._mo)
            self._mo = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
