"""Mirror of the reference's stereo fusion (src/colmap/mvs/fusion.{h,cc}) over include/b200_mvs_fusion.h: the consumer of
the sweep's depth / normal maps.  Names follow the reference: StereoFusionOptions,
StereoFusion(options, workspace_path, ..., input_type).Run(), GetFusedPoints(), GetFusedPointsVisibility(),
WritePointsVisibility / ReadPointsVisibility."""
import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from ._lib import load_library
from .mvs_workspace import Model, WorkspaceError, read_depth_map, read_normal_map

_f32p = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
FLT_MAX = float(np.finfo(np.float32).max)


class _COptions(ctypes.Structure):
    _fields_ = [("min_num_pixels", ctypes.c_int), ("max_num_pixels", ctypes.c_int), ("max_traversal_depth", ctypes.c_int),
                ("max_reproj_error", ctypes.c_double), ("max_depth_error", ctypes.c_double), ("max_normal_error", ctypes.c_double),
                ("bbox_min", ctypes.c_float * 3), ("bbox_max", ctypes.c_float * 3)]


class _CImage(ctypes.Structure):
    _fields_ = [("used", ctypes.c_int), ("image_width", ctypes.c_int), ("image_height", ctypes.c_int), ("K", _f32p), ("R", _f32p),
                ("T", _f32p), ("map_width", ctypes.c_int), ("map_height", ctypes.c_int), ("depth", _f32p), ("normal", _f32p),
                ("bitmap_width", ctypes.c_int), ("bitmap_height", ctypes.c_int), ("rgb", _u8p), ("mask", _u8p)]


_BOUND = None


def _lib():
    global _BOUND
    if _BOUND is None:
        L = load_library()
        L.b200fuse_last_error.restype = ctypes.c_char_p
        L.b200fuse_options_init.argtypes = [ctypes.POINTER(_COptions)]
        L.b200fuse_options_check.argtypes = [ctypes.POINTER(_COptions)]
        L.b200fuse_run.argtypes = [ctypes.POINTER(_COptions), ctypes.c_int, ctypes.POINTER(_CImage), _i32p, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_void_p)]
        L.b200fuse_num_points.argtypes = [ctypes.c_void_p]; L.b200fuse_num_points.restype = ctypes.c_int64
        L.b200fuse_num_visibility.argtypes = [ctypes.c_void_p]; L.b200fuse_num_visibility.restype = ctypes.c_int64
        L.b200fuse_get.argtypes = [ctypes.c_void_p, _f32p, _f32p, _u8p, _i64p, _i32p]
        L.b200fuse_free.argtypes = [ctypes.c_void_p]
        L.b200fuse_write_visibility.argtypes = [ctypes.c_char_p, ctypes.c_int64, _i64p, _i32p]
        L.b200fuse_read_visibility.argtypes = [ctypes.c_char_p, ctypes.c_int64, _i64p, _i32p, _i64p]
        _BOUND = L
    return _BOUND


def _check(rc, what):
    if rc != 0:
        raise WorkspaceError(f"{what} failed ({rc}): {_lib().b200fuse_last_error().decode()}")


@dataclass
class StereoFusionOptions:
    """StereoFusionOptions (fusion.h:47-95); mask_path / cache knobs belong to the workspace and are not used here."""
    mask_path: str = ""
    num_threads: int = -1
    max_image_size: int = -1
    min_num_pixels: int = 5
    max_num_pixels: int = 10000
    max_traversal_depth: int = 100
    max_reproj_error: float = 2.0
    max_depth_error: float = 0.01
    max_normal_error: float = 10.0
    check_num_images: int = 50
    use_cache: bool = False
    cache_size: float = 32.0
    bounding_box: tuple = ((-FLT_MAX,) * 3, (FLT_MAX,) * 3)

    def to_c(self) -> _COptions:
        c = _COptions()
        _lib().b200fuse_options_init(ctypes.byref(c))
        c.min_num_pixels, c.max_num_pixels, c.max_traversal_depth = self.min_num_pixels, self.max_num_pixels, self.max_traversal_depth
        c.max_reproj_error, c.max_depth_error, c.max_normal_error = self.max_reproj_error, self.max_depth_error, self.max_normal_error
        for k in range(3):
            c.bbox_min[k] = self.bounding_box[0][k]; c.bbox_max[k] = self.bounding_box[1][k]
        return c

    def Check(self) -> bool:
        return bool(_lib().b200fuse_options_check(ctypes.byref(self.to_c()))) and self.check_num_images > 0 and self.cache_size > 0


@dataclass
class FusionImage:
    """Inputs of one view: calibration (for an image of image_width x image_height), maps, RGB bitmap, optional mask."""
    K: np.ndarray
    R: np.ndarray
    T: np.ndarray
    image_width: int
    image_height: int
    depth_map: Optional[np.ndarray] = None      # (h, w) float32
    normal_map: Optional[np.ndarray] = None     # (3, h, w) float32, camera frame
    bitmap: Optional[np.ndarray] = None         # (H, W, 3) uint8
    mask: Optional[np.ndarray] = None           # (h, w) non-zero = skip
    used: bool = True


@dataclass
class FusedPoints:
    xyz: np.ndarray       # (n, 3) float32
    normal: np.ndarray    # (n, 3) float32
    rgb: np.ndarray       # (n, 3) uint8
    visibility: List[np.ndarray] = field(default_factory=list)


def fuse(options: StereoFusionOptions, images: Sequence[FusionImage], overlapping_images: Sequence[Sequence[int]]) -> FusedPoints:
    """StereoFusion::Run on in-memory inputs (single-threaded schedule of the reference)."""
    L = _lib()
    n = len(images)
    keep = []
    arr = (_CImage * max(n, 1))()
    for i, im in enumerate(images):
        c = arr[i]
        ok = bool(im.used and im.depth_map is not None and im.normal_map is not None and im.bitmap is not None)
        c.used = 1 if ok else 0
        c.image_width, c.image_height = int(im.image_width), int(im.image_height)
        K = np.ascontiguousarray(im.K, np.float32).reshape(9); R = np.ascontiguousarray(im.R, np.float32).reshape(9)
        T = np.ascontiguousarray(im.T, np.float32).reshape(3)
        keep += [K, R, T]
        c.K, c.R, c.T = K.ctypes.data_as(_f32p), R.ctypes.data_as(_f32p), T.ctypes.data_as(_f32p)
        if ok:
            d = np.ascontiguousarray(im.depth_map, np.float32); nm = np.ascontiguousarray(im.normal_map, np.float32)
            bm = np.ascontiguousarray(im.bitmap, np.uint8)
            if nm.shape != (3,) + d.shape or bm.ndim != 3 or bm.shape[2] != 3:
                raise WorkspaceError("fusion input shapes: depth (h,w), normal (3,h,w), bitmap (H,W,3)")
            keep += [d, nm, bm]
            c.map_height, c.map_width = d.shape
            c.depth, c.normal = d.ctypes.data_as(_f32p), nm.ctypes.data_as(_f32p)
            c.bitmap_height, c.bitmap_width = bm.shape[:2]
            c.rgb = bm.ctypes.data_as(_u8p)
            if im.mask is not None:
                mk = np.ascontiguousarray(im.mask != 0, np.uint8)
                keep.append(mk)
                c.mask = mk.ctypes.data_as(_u8p)
    mo = max([len(o) for o in overlapping_images] + [1])
    ov = np.full((max(n, 1), mo), -1, np.int32)
    for i, o in enumerate(overlapping_images):
        ov[i, :len(o)] = o
    co = options.to_c()
    h = ctypes.c_void_p()
    _check(L.b200fuse_run(ctypes.byref(co), n, arr, ov.ctypes.data_as(_i32p), mo, ctypes.byref(h)), "StereoFusion")
    try:
        npts, nvis = L.b200fuse_num_points(h), L.b200fuse_num_visibility(h)
        xyz = np.empty((npts, 3), np.float32); nrm = np.empty((npts, 3), np.float32); rgb = np.empty((npts, 3), np.uint8)
        off = np.empty(npts + 1, np.int64); vis = np.empty(max(nvis, 1), np.int32)
        _check(L.b200fuse_get(h, xyz.ctypes.data_as(_f32p), nrm.ctypes.data_as(_f32p), rgb.ctypes.data_as(_u8p),
                              off.ctypes.data_as(_i64p), vis.ctypes.data_as(_i32p)), "fusion result")
    finally:
        L.b200fuse_free(h)
    return FusedPoints(xyz, nrm, rgb, [vis[off[i]:off[i + 1]].copy() for i in range(npts)])


def WritePointsVisibility(path: str, points_visibility: Sequence[Sequence[int]]) -> None:
    off = np.zeros(len(points_visibility) + 1, np.int64)
    off[1:] = np.cumsum([len(v) for v in points_visibility])
    vis = np.ascontiguousarray([i for v in points_visibility for i in v], np.int32)
    if vis.size == 0:
        vis = np.zeros(1, np.int32)
    _check(_lib().b200fuse_write_visibility(os.fsencode(path), len(points_visibility), off.ctypes.data_as(_i64p),
                                            vis.ctypes.data_as(_i32p)), "WritePointsVisibility")


def ReadPointsVisibility(path: str, num_points: int) -> List[np.ndarray]:
    L = _lib()
    total = ctypes.c_int64()
    _check(L.b200fuse_read_visibility(os.fsencode(path), num_points, None, None, ctypes.byref(total)), "ReadPointsVisibility")
    off = np.empty(num_points + 1, np.int64); vis = np.empty(max(total.value, 1), np.int32)
    _check(L.b200fuse_read_visibility(os.fsencode(path), num_points, off.ctypes.data_as(_i64p), vis.ctypes.data_as(_i32p),
                                      ctypes.byref(total)), "ReadPointsVisibility")
    return [vis[off[i]:off[i + 1]].copy() for i in range(num_points)]


def write_ply(path: str, pts: FusedPoints) -> None:
    """Binary little-endian PLY with x y z nx ny nz red green blue (WriteBinaryPlyPoints, util/ply.cc)."""
    n = len(pts.xyz)
    rec = np.empty(n, np.dtype([("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)]))
    rec["p"], rec["n"], rec["c"] = pts.xyz, pts.normal, pts.rgb
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\n"
                 "property uchar blue\nend_header\n").encode())
        f.write(rec.tobytes())


class StereoFusion:
    """StereoFusion (fusion.h:98-163) on a COLMAP workspace: `stereo/fusion.cfg` lists the images, the maps come from
    `stereo/{depth_maps,normal_maps}/<image>.<input_type>.bin`, colours from `images/`."""

    def __init__(self, options: StereoFusionOptions, workspace_path: str, workspace_format: str = "COLMAP",
                 pmvs_option_name: str = "", input_type: str = "geometric", stereo_folder: str = "stereo"):
        if not options.Check():
            raise WorkspaceError("StereoFusionOptions::Check failed")
        if workspace_format.lower() != "colmap":
            raise WorkspaceError("only COLMAP workspaces are supported")
        self.options, self.workspace_path, self.input_type, self.stereo_folder = options, workspace_path, input_type, stereo_folder
        self._points: Optional[FusedPoints] = None

    def Run(self) -> None:
        from PIL import Image as PILImage
        base = os.path.join(self.workspace_path, self.stereo_folder)
        with open(os.path.join(base, "fusion.cfg")) as f:
            names = [ln.strip() for ln in f if ln.strip() and not ln.strip().startswith("#")]
        model = Model.ReadFromCOLMAP(self.workspace_path)
        overlapping = model.GetMaxOverlappingImages(self.options.check_num_images, 0.0)
        images = [FusionImage(mi.K, mi.R, mi.T, mi.width, mi.height, used=False) for mi in model.images]
        for name in names:
            idx = model.GetImageIdx(name)
            dp = os.path.join(base, "depth_maps", f"{name}.{self.input_type}.bin")
            npth = os.path.join(base, "normal_maps", f"{name}.{self.input_type}.bin")
            if not (os.path.exists(dp) and os.path.exists(npth) and os.path.exists(model.images[idx].path)):
                continue   # "Ignoring image ..., because input does not exist."
            im = images[idx]
            im.depth_map, im.normal_map = read_depth_map(dp), read_normal_map(npth)
            bm = PILImage.open(model.images[idx].path).convert("RGB")
            im.bitmap = np.asarray(bm, np.uint8)
            im.image_width, im.image_height = im.bitmap.shape[1], im.bitmap.shape[0]
            if (im.image_width, im.image_height) != (model.images[idx].width, model.images[idx].height):   # Image::Rescale
                sx, sy = im.image_width / model.images[idx].width, im.image_height / model.images[idx].height
                K = np.asarray(im.K, np.float64).copy(); K[0, 0] *= sx; K[0, 2] *= sx; K[1, 1] *= sy; K[1, 2] *= sy
                im.K = K.astype(np.float32)
            if self.options.mask_path:
                mp = os.path.join(self.options.mask_path, name + ".png")
                if not os.path.exists(mp) and name.lower().endswith(".png"):
                    mp = os.path.join(self.options.mask_path, name)
                if os.path.exists(mp):
                    mk = PILImage.open(mp).convert("L").resize((im.depth_map.shape[1], im.depth_map.shape[0]), PILImage.BOX)
                    im.mask = (np.asarray(mk) == 0)
            im.used = True
        self._points = fuse(self.options, images, overlapping)

    def GetFusedPoints(self) -> FusedPoints:
        return self._points

    def GetFusedPointsVisibility(self) -> List[np.ndarray]:
        return self._points.visibility
