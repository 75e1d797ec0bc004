"""ctypes driver of the PatchMatch CPU oracle (oracle/libpm_oracle.so).  Test infrastructure."""
import ctypes
import os
import subprocess

import numpy as np

from colmap_b200.patch_match import _COptions, _CProblem, _f32p, _u8p, marshal

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "libpm_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "libpm_oracle.so"])
        L = ctypes.CDLL(path)
        L.pm_oracle_run_partial.argtypes = [ctypes.POINTER(_COptions), ctypes.POINTER(_CProblem), ctypes.c_int,
                                            _f32p, _f32p, _f32p, _u8p, _f32p]
        L.pm_oracle_expf.argtypes = [ctypes.c_float]
        L.pm_oracle_expf.restype = ctypes.c_float
        L.pm_oracle_sincosf.argtypes = [ctypes.c_float, _f32p, _f32p]
        L.pm_oracle_rng_stream.argtypes = [ctypes.c_uint64, ctypes.c_int, _f32p]
        L.pm_oracle_poses.argtypes = [ctypes.POINTER(_CProblem), ctypes.c_int, _f32p, _f32p, _f32p]
        L.pm_oracle_rotate_f32.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        for name in ("pm_oracle_compose_projection_matrix", "pm_oracle_compose_inverse_projection_matrix"):
            getattr(L, name).argtypes = [_f32p, _f32p, _f32p, _f32p]
        L.pm_oracle_projection_center.argtypes = [_f32p, _f32p, _f32p]
        _LIB = L
    return _LIB


def run(options, problem, stop_after_sweeps=-1):
    """Returns dict(depth (H,W), normal (3,H,W), sel_prob (N,H,W), mask (N,H,W), cost (N,H,W))."""
    co, cp, keep = marshal(options, problem)
    h, w, n = cp.ref_height, cp.ref_width, cp.num_src
    depth = np.empty((h, w), np.float32)
    normal = np.empty((3, h, w), np.float32)
    sel = np.empty((n, h, w), np.float32)
    mask = np.empty((n, h, w), np.uint8)
    cost = np.empty((n, h, w), np.float32)
    rc = lib().pm_oracle_run_partial(ctypes.byref(co), ctypes.byref(cp), stop_after_sweeps,
                                     depth.ctypes.data_as(_f32p), normal.ctypes.data_as(_f32p),
                                     sel.ctypes.data_as(_f32p), mask.ctypes.data_as(_u8p),
                                     cost.ctypes.data_as(_f32p))
    if rc != 0:
        raise RuntimeError(f"pm_oracle_run_partial -> {rc}")
    return dict(depth=depth, normal=normal, sel_prob=sel, mask=mask, cost=cost)
