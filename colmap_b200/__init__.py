"""colmap_b200 — B200-native (sm_100a) implementation of COLMAP's two numeric hot paths.

* PatchMatch MVS sweep  (reference: src/colmap/mvs/patch_match_cuda.{h,cu})
* Bundle-adjustment LM normal-equation build + solve
  (reference: src/colmap/estimators/bundle_adjustment*.{h,cc})

The product is the C-ABI shared library ``colmap_b200/lib/libcolmap_b200.so`` declared in
``include/*.h``; this package is the thin host-side mirror of the reference's interfaces over that
ABI (ctypes).  There is no CPU fallback: importing the ops without the built CUDA library raises.
"""
from ._lib import load_library, library_path  # noqa: F401

__all__ = ["load_library", "library_path"]
