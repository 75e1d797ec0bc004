"""ctypes driver of oracle/_ref/libpm_ref.so: the UNMODIFIED reference PatchMatchCuda, compiled from
/root/reference against stub headers (oracle/build_ref.sh).  Test / measurement infrastructure only."""
import ctypes
import os

import numpy as np

from colmap_b200.patch_match import _COptions, _CProblem, _f32p, marshal

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(_ROOT, "oracle", "_ref", "libpm_ref.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(PATH)
        L.pm_ref_run.argtypes = [ctypes.POINTER(_COptions), ctypes.POINTER(_CProblem), _f32p, _f32p, _f32p,
                                 ctypes.POINTER(ctypes.c_double), ctypes.c_char_p, ctypes.c_int]
        _LIB = L
    return _LIB


def run(options, problem):
    """Returns dict(depth, normal, sel_prob, ms) — ms = wall time of constructor + Run + GetDepthMap/GetNormalMap."""
    co, cp, keep = marshal(options, problem)
    h, w, n = cp.ref_height, cp.ref_width, cp.num_src
    depth = np.empty((h, w), np.float32); normal = np.empty((3, h, w), np.float32); sel = np.empty((n, h, w), np.float32)
    ms = ctypes.c_double()
    err = ctypes.create_string_buffer(512)
    rc = lib().pm_ref_run(ctypes.byref(co), ctypes.byref(cp), depth.ctypes.data_as(_f32p), normal.ctypes.data_as(_f32p),
                          sel.ctypes.data_as(_f32p), ctypes.byref(ms), err, 512)
    if rc != 0:
        raise RuntimeError("reference PatchMatchCuda failed: " + err.value.decode())
    return dict(depth=depth, normal=normal, sel_prob=sel, ms=ms.value)
