"""ctypes driver of the BA CPU oracle (oracle/libba_oracle.so).  Test infrastructure."""
import ctypes
import os
import subprocess

import numpy as np

from colmap_b200.bundle_adjustment import (BundleAdjustmentSummary, _COptions, _CProblem, _CSummary, _f64p)

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "libba_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "libba_oracle.so"])
        L = ctypes.CDLL(path)
        L.ba_oracle_solve.argtypes = [ctypes.POINTER(_COptions), ctypes.POINTER(_CProblem), ctypes.POINTER(_CSummary)]
        L.ba_oracle_reproj.argtypes = [ctypes.c_int, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.ba_oracle_quat_plus.argtypes = [_f64p, _f64p, _f64p]
        L.ba_oracle_reproj_rig.argtypes = [ctypes.c_int, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        _LIB = L
    return _LIB


def solve(options, flat):
    """Runs the oracle on `flat` in place; returns BundleAdjustmentSummary."""
    co, cp, cs = options.to_c(), flat.to_c(), _CSummary()
    rc = lib().ba_oracle_solve(ctypes.byref(co), ctypes.byref(cp), ctypes.byref(cs))
    if rc != 0:
        raise RuntimeError(f"ba_oracle_solve -> {rc}")
    return BundleAdjustmentSummary.from_c(cs)


def reproj(model_id, point, pose, params, xy):
    d = lambda a: np.ascontiguousarray(a, np.float64)
    point, pose, params, xy = d(point), d(pose), d(params), d(xy)
    res = np.zeros(2); Jpt = np.zeros((2, 3)); Jps = np.zeros((2, 7)); Jpr = np.zeros((2, len(params)))
    ok = lib().ba_oracle_reproj(model_id, point.ctypes.data_as(_f64p), pose.ctypes.data_as(_f64p),
                                params.ctypes.data_as(_f64p), xy.ctypes.data_as(_f64p), res.ctypes.data_as(_f64p),
                                Jpt.ctypes.data_as(_f64p), Jps.ctypes.data_as(_f64p), Jpr.ctypes.data_as(_f64p))
    return ok, res, Jpt, Jps, Jpr


def reproj_rig(model_id, point, rig, sensor, params, xy):
    """ba_oracle_reproj_rig: residual + Jacobians of cam_from_world = sensor_from_rig * rig_from_world (sensor None = identity)."""
    d = lambda a: np.ascontiguousarray(a, np.float64)
    point, rig, params, xy = d(point), d(rig), d(params), d(xy)
    sp = d(sensor).ctypes.data_as(_f64p) if sensor is not None else None
    res = np.zeros(2); Jpt = np.zeros((2, 3)); Jr = np.zeros((2, 7)); Js = np.zeros((2, 7)); Jpr = np.zeros((2, len(params)))
    ok = lib().ba_oracle_reproj_rig(model_id, point.ctypes.data_as(_f64p), rig.ctypes.data_as(_f64p), sp, params.ctypes.data_as(_f64p),
                                    xy.ctypes.data_as(_f64p), res.ctypes.data_as(_f64p), Jpt.ctypes.data_as(_f64p),
                                    Jr.ctypes.data_as(_f64p), Js.ctypes.data_as(_f64p), Jpr.ctypes.data_as(_f64p))
    return ok, res, Jpt, Jr, Js, Jpr
