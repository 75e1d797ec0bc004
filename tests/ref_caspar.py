"""ctypes driver of oracle/_ref/libcaspar_ref.so: the reference's own GPU bundle-adjustment backend (the generated
Caspar solver, compiled from /root/reference by oracle/build_caspar.sh) behind oracle/caspar_harness.cu.
Test / measurement infrastructure only."""
import ctypes
import os

from colmap_b200.bundle_adjustment import BundleAdjustmentOptions, _COptions, _CProblem

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(_ROOT, "oracle", "_ref", "libcaspar_ref.so")
_LIB = None


class _CResult(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int), ("exit_reason", ctypes.c_int), ("initial_cost", ctypes.c_double),
                ("final_cost", ctypes.c_double), ("solve_ms", ctypes.c_double), ("setup_ms", ctypes.c_double),
                ("num_residuals", ctypes.c_int)]


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(PATH)
        L.caspar_ref_solve.argtypes = [ctypes.POINTER(_COptions), ctypes.POINTER(_CProblem), ctypes.c_int,
                                       ctypes.POINTER(_CResult), ctypes.c_char_p, ctypes.c_int]
        _LIB = L
    return _LIB


def solve(flat, options=None, solver_iter_max=0):
    """Runs Caspar on `flat` in place (fp32 inside); returns dict(iterations, initial_cost, final_cost, solve_ms, ...).
    solver_iter_max = 0 -> CasparBundleAdjustmentOptions default (200)."""
    options = options or BundleAdjustmentOptions()
    co, cp, cr = options.to_c(), flat.to_c(), _CResult()
    err = ctypes.create_string_buffer(512)
    rc = lib().caspar_ref_solve(ctypes.byref(co), ctypes.byref(cp), int(solver_iter_max), ctypes.byref(cr), err, 512)
    if rc != 0:
        raise RuntimeError("Caspar reference failed: " + err.value.decode())
    return {f[0]: getattr(cr, f[0]) for f in _CResult._fields_}
