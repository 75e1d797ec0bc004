"""Loader of the C-ABI library.  Fails loudly when the CUDA extension is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path():
    return os.path.join(_HERE, "lib", "libcolmap_b200.so")


def load_library():
    """dlopen libcolmap_b200.so (built by __graft_entry__.build()).  No fallback of any kind."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  colmap_b200 has no CPU fallback.")
        _LIB = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    return _LIB
