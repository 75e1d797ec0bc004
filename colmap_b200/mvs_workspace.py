"""Host-side mirror of the reference's MVS workspace pieces either side of the sweep:

  Mat / DepthMap / NormalMap files          src/colmap/mvs/mat.cc:41-66, depth_map.cc, normal_map.cc
  ConsistencyGraph                          src/colmap/mvs/consistency_graph.{h,cc}
  Model (sparse model statistics)           src/colmap/mvs/model.{h,cc}
  patch-match.cfg                           src/colmap/mvs/patch_match.cc:240-372
  PatchMatchController (problem schedule)   src/colmap/mvs/patch_match.cc:156-536

The arithmetic lives in C++ behind include/b200_mvs_workspace.h (colmap_b200/csrc/mvs_workspace.cu); this module is the
ctypes mirror with the reference's names, plus the readers the controller needs (COLMAP sparse model, bitmaps)."""
import ctypes
import os
import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import load_library
from .patch_match import Image, PatchMatch, PatchMatchOptions, Problem, consistency_list_from_mask

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class WorkspaceError(RuntimeError):
    pass


class _CModel(ctypes.Structure):
    _fields_ = [("num_images", ctypes.c_int), ("R", _f32p), ("T", _f32p), ("num_points", ctypes.c_int64), ("xyz", _f32p),
                ("track_offset", _i64p), ("track", _i32p)]


_BOUND = None


def _lib():
    global _BOUND
    if _BOUND is None:
        L = load_library()
        cp, ip, sz = ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t
        L.b200ws_last_error.restype = ctypes.c_char_p
        L.b200ws_mat_read_header.argtypes = [cp, ip, ip, ip]
        L.b200ws_mat_read.argtypes = [cp, _f32p, sz]
        L.b200ws_mat_write.argtypes = [cp, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.b200ws_graph_write.argtypes = [cp, ctypes.c_int, ctypes.c_int, _i32p, sz]
        L.b200ws_graph_read.argtypes = [cp, ip, ip, _i32p, sz, ctypes.POINTER(sz)]
        L.b200ws_graph_build_map.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, sz, _i32p]
        mp = ctypes.POINTER(_CModel)
        L.b200ws_compute_depth_ranges.argtypes = [mp, _f32p]
        L.b200ws_compute_shared_points.argtypes = [mp, _i32p]
        L.b200ws_compute_triangulation_angles.argtypes = [mp, ctypes.c_float, _f32p]
        L.b200ws_max_overlapping_images.argtypes = [mp, ctypes.c_int, ctypes.c_double, _i32p, _i32p]
        L.b200ws_read_problems.argtypes = [cp, mp, ctypes.POINTER(cp), ctypes.c_double, _i32p, _i64p, _i32p, sz, sz,
                                           ctypes.POINTER(sz), ctypes.POINTER(sz)]
        _BOUND = L
    return _BOUND


def _check(rc, what):
    if rc != 0:
        raise WorkspaceError(f"{what} failed ({rc}): {_lib().b200ws_last_error().decode()}")


# ---------------------------------------------------------------------------------------------------------------------
# map files
# ---------------------------------------------------------------------------------------------------------------------
def read_mat(path: str) -> np.ndarray:
    """Mat<float>::Read: returns (depth, height, width) float32 (slice-major, then row-major)."""
    L = _lib()
    w, h, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(L.b200ws_mat_read_header(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), ctypes.byref(d)), "read map header")
    out = np.empty((d.value, h.value, w.value), np.float32)
    _check(L.b200ws_mat_read(os.fsencode(path), out.ctypes.data_as(_f32p), out.size), "read map")
    return out


def write_mat(path: str, data: np.ndarray) -> None:
    """Mat<float>::Write of a (height, width) or (depth, height, width) array."""
    a = np.ascontiguousarray(data, np.float32)
    if a.ndim == 2:
        a = a[None]
    d, h, w = a.shape
    _check(_lib().b200ws_mat_write(os.fsencode(path), a.ctypes.data_as(_f32p), w, h, d), "write map")


def read_depth_map(path: str) -> np.ndarray:
    m = read_mat(path)
    if m.shape[0] != 1:
        raise WorkspaceError(f"{path}: a depth map has one slice, found {m.shape[0]}")
    return m[0]


def read_normal_map(path: str) -> np.ndarray:
    m = read_mat(path)
    if m.shape[0] != 3:
        raise WorkspaceError(f"{path}: a normal map has three slices, found {m.shape[0]}")
    return m


class ConsistencyGraph:
    """mvs::ConsistencyGraph: records [col, row, n, idx_1..idx_n] + a per-pixel offset map."""

    kNoConsistentImageIds = -1

    def __init__(self, width: int = 0, height: int = 0, data: Sequence[int] = ()):
        self.width, self.height = int(width), int(height)
        self.data = np.ascontiguousarray(data, np.int32)
        self.map = np.zeros((0, 0), np.int32)
        if width and height:
            self._initialize_map()

    def _initialize_map(self):
        self.map = np.empty((self.height, self.width), np.int32)
        _check(_lib().b200ws_graph_build_map(self.width, self.height, self.data.ctypes.data_as(_i32p), self.data.size,
                                             self.map.ctypes.data_as(_i32p)), "consistency graph")

    def GetNumBytes(self) -> int:
        return (self.data.size + self.map.size) * 4

    def GetImageIdxs(self, row: int, col: int) -> np.ndarray:
        index = int(self.map[row, col])
        if index == self.kNoConsistentImageIds:
            return np.zeros(0, np.int32)
        n = int(self.data[index])
        return self.data[index + 1:index + 1 + n]

    def Write(self, path: str) -> None:
        _check(_lib().b200ws_graph_write(os.fsencode(path), self.width, self.height, self.data.ctypes.data_as(_i32p),
                                         self.data.size), "write consistency graph")

    def Read(self, path: str) -> None:
        L = _lib()
        w, h, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _check(L.b200ws_graph_read(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), None, 0, ctypes.byref(n)), "read consistency graph")
        self.data = np.empty(n.value, np.int32)
        _check(L.b200ws_graph_read(os.fsencode(path), ctypes.byref(w), ctypes.byref(h), self.data.ctypes.data_as(_i32p),
                                   self.data.size, ctypes.byref(n)), "read consistency graph")
        self.width, self.height = w.value, h.value
        self._initialize_map()


# ---------------------------------------------------------------------------------------------------------------------
# sparse model
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class ModelImage:
    path: str
    width: int
    height: int
    K: np.ndarray
    R: np.ndarray
    T: np.ndarray


@dataclass
class ModelPoint:
    x: float
    y: float
    z: float
    track: List[int] = field(default_factory=list)


class Model:
    """mvs::Model (model.h:44-118): images with K/R/T and points with image-index tracks."""

    def __init__(self):
        self.images: List[ModelImage] = []
        self.points: List[ModelPoint] = []
        self._names: List[str] = []

    # -- construction
    def add_image(self, name: str, width: int, height: int, K, R, T, path: Optional[str] = None) -> int:
        self.images.append(ModelImage(path or name, int(width), int(height), np.asarray(K, np.float32).reshape(3, 3),
                                      np.asarray(R, np.float32).reshape(3, 3), np.asarray(T, np.float32).reshape(3)))
        self._names.append(name)
        return len(self.images) - 1

    def GetImageIdx(self, name: str) -> int:
        try:
            return self._names.index(name)
        except ValueError:
            raise WorkspaceError(f"Image with name `{name}` does not exist")

    def GetImageName(self, image_idx: int) -> str:
        if image_idx < 0 or image_idx >= len(self._names):
            raise WorkspaceError("image index out of range")
        return self._names[image_idx]

    # -- C view
    def _c(self):
        n = len(self.images)
        R = np.ascontiguousarray(np.stack([im.R.reshape(9) for im in self.images]) if n else np.zeros((0, 9)), np.float32)
        T = np.ascontiguousarray(np.stack([im.T for im in self.images]) if n else np.zeros((0, 3)), np.float32)
        xyz = np.ascontiguousarray([[p.x, p.y, p.z] for p in self.points], np.float32).reshape(-1, 3)
        off = np.zeros(len(self.points) + 1, np.int64)
        off[1:] = np.cumsum([len(p.track) for p in self.points])
        track = np.ascontiguousarray([i for p in self.points for i in p.track], np.int32)
        c = _CModel(n, R.ctypes.data_as(_f32p), T.ctypes.data_as(_f32p), len(self.points), xyz.ctypes.data_as(_f32p),
                    off.ctypes.data_as(_i64p), track.ctypes.data_as(_i32p))
        return c, (R, T, xyz, off, track)

    # -- statistics (model.cc:120-283)
    def ComputeDepthRanges(self) -> List[Tuple[float, float]]:
        c, keep = self._c()
        out = np.empty(2 * len(self.images), np.float32)
        _check(_lib().b200ws_compute_depth_ranges(ctypes.byref(c), out.ctypes.data_as(_f32p)), "ComputeDepthRanges")
        return [(float(out[2 * i]), float(out[2 * i + 1])) for i in range(len(self.images))]

    def ComputeSharedPoints(self) -> List[Dict[int, int]]:
        c, keep = self._c()
        n = len(self.images)
        out = np.zeros((n, n), np.int32)
        _check(_lib().b200ws_compute_shared_points(ctypes.byref(c), out.ctypes.data_as(_i32p)), "ComputeSharedPoints")
        return [{int(j): int(out[i, j]) for j in np.nonzero(out[i])[0]} for i in range(n)]

    def ComputeTriangulationAngles(self, percentile: float = 50.0) -> List[Dict[int, float]]:
        c, keep = self._c()
        n = len(self.images)
        out = np.zeros((n, n), np.float32)
        _check(_lib().b200ws_compute_triangulation_angles(ctypes.byref(c), percentile, out.ctypes.data_as(_f32p)),
               "ComputeTriangulationAngles")
        return [{int(j): float(out[i, j]) for j in np.nonzero(out[i] >= 0)[0]} for i in range(n)]

    def GetMaxOverlappingImages(self, num_images: int, min_triangulation_angle: float) -> List[List[int]]:
        c, keep = self._c()
        n = len(self.images)
        out = np.full((n, max(num_images, 1)), -1, np.int32)
        cnt = np.zeros(n, np.int32)
        _check(_lib().b200ws_max_overlapping_images(ctypes.byref(c), num_images, min_triangulation_angle,
                                                    out.ctypes.data_as(_i32p), cnt.ctypes.data_as(_i32p)), "GetMaxOverlappingImages")
        return [[int(v) for v in out[i, :cnt[i]]] for i in range(n)]

    # -- COLMAP sparse model (scene/reconstruction_io_binary.cc:107-291; legacy trio cameras / images / points3D .bin)
    @staticmethod
    def ReadFromCOLMAP(path: str, sparse_path: str = "sparse", images_path: str = "images") -> "Model":
        cams = read_cameras_binary(os.path.join(path, sparse_path, "cameras.bin"))
        imgs = read_images_binary(os.path.join(path, sparse_path, "images.bin"))
        pts = read_points3D_binary(os.path.join(path, sparse_path, "points3D.bin"))
        model = Model()
        image_id_to_idx = {}
        for image_id in sorted(imgs):                      # RegImageIds(): ascending image id
            im = imgs[image_id]
            cam = cams[im["camera_id"]]
            K = calibration_matrix(cam["model_id"], cam["params"])
            R = quat_wxyz_to_R(im["qvec"])
            image_id_to_idx[image_id] = model.add_image(im["name"], cam["width"], cam["height"], K, R, im["tvec"],
                                                        path=os.path.join(path, images_path, im["name"]))
        for pid in pts:
            p = pts[pid]
            model.points.append(ModelPoint(float(np.float32(p["xyz"][0])), float(np.float32(p["xyz"][1])), float(np.float32(p["xyz"][2])),
                                           [image_id_to_idx[i] for i, _ in p["track"]]))
        return model


_MODEL_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}


def calibration_matrix(model_id: int, params) -> np.ndarray:
    """Camera::CalibrationMatrix: focal length(s) and principal point of the model, distortion ignored."""
    p = np.asarray(params, np.float64)
    single_focal = model_id in (0, 2, 3, 8, 9)
    fx, fy = (p[0], p[0]) if single_focal else (p[0], p[1])
    cx, cy = (p[1], p[2]) if single_focal else (p[2], p[3])
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)


def quat_wxyz_to_R(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    n = (w * w + x * x + y * y + z * z) ** 0.5
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)


def read_cameras_binary(path: str) -> Dict[int, dict]:
    out = {}
    with open(path, "rb") as f:
        n, = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            cam_id, model_id, w, h = struct.unpack("<IiQQ", f.read(24))
            if model_id not in _MODEL_PARAMS:
                raise WorkspaceError(f"{path}: camera model {model_id} is not known to this reader")
            k = _MODEL_PARAMS[model_id]
            out[cam_id] = dict(model_id=model_id, width=w, height=h, params=np.frombuffer(f.read(8 * k), "<f8").copy())
    return out


def read_images_binary(path: str) -> Dict[int, dict]:
    out = {}
    with open(path, "rb") as f:
        n, = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            image_id, = struct.unpack("<I", f.read(4))
            qt = struct.unpack("<7d", f.read(56))
            cam_id, = struct.unpack("<I", f.read(4))
            name = bytearray()
            while True:
                c = f.read(1)
                if c in (b"\x00", b""):
                    break
                name += c
            m, = struct.unpack("<Q", f.read(8))
            rec = np.frombuffer(f.read(24 * m), np.dtype([("x", "<f8"), ("y", "<f8"), ("id", "<u8")]))
            out[image_id] = dict(qvec=np.array(qt[:4]), tvec=np.array(qt[4:]), camera_id=cam_id, name=name.decode(),
                                 xys=np.stack([rec["x"], rec["y"]], 1) if m else np.zeros((0, 2)),
                                 point3D_ids=rec["id"].astype(np.int64) if m else np.zeros(0, np.int64))
    return out


def read_points3D_binary(path: str) -> Dict[int, dict]:
    out = {}
    with open(path, "rb") as f:
        n, = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            pid, = struct.unpack("<Q", f.read(8))
            xyz = struct.unpack("<3d", f.read(24))
            rgb = struct.unpack("<3B", f.read(3))
            err, = struct.unpack("<d", f.read(8))
            tl, = struct.unpack("<Q", f.read(8))
            tr = np.frombuffer(f.read(8 * tl), "<u4").reshape(-1, 2)
            out[pid] = dict(xyz=np.array(xyz), rgb=rgb, error=err, track=[(int(a), int(b)) for a, b in tr])
    return out


def write_model_binary(sparse_dir: str, cameras: Dict[int, dict], images: Dict[int, dict], points3D: Dict[int, dict]) -> None:
    """Writer of the same three files (for tests and synthetic workspaces)."""
    os.makedirs(sparse_dir, exist_ok=True)
    with open(os.path.join(sparse_dir, "cameras.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for cid in sorted(cameras):
            c = cameras[cid]
            f.write(struct.pack("<IiQQ", cid, c["model_id"], c["width"], c["height"]))
            f.write(np.asarray(c["params"], "<f8").tobytes())
    with open(os.path.join(sparse_dir, "images.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for iid in sorted(images):
            im = images[iid]
            f.write(struct.pack("<I", iid))
            f.write(struct.pack("<7d", *im["qvec"], *im["tvec"]))
            f.write(struct.pack("<I", im["camera_id"]))
            f.write(im["name"].encode() + b"\x00")
            xys = np.asarray(im.get("xys", np.zeros((0, 2))), np.float64).reshape(-1, 2)
            ids = np.asarray(im.get("point3D_ids", np.zeros(0)), np.int64)
            f.write(struct.pack("<Q", len(xys)))
            for (x, y), pid in zip(xys, ids):
                f.write(struct.pack("<ddQ", x, y, int(pid) & 0xFFFFFFFFFFFFFFFF))
    with open(os.path.join(sparse_dir, "points3D.bin"), "wb") as f:
        f.write(struct.pack("<Q", len(points3D)))
        for pid in sorted(points3D):
            p = points3D[pid]
            f.write(struct.pack("<Q", pid))
            f.write(struct.pack("<3d", *p["xyz"]))
            f.write(struct.pack("<3B", *p.get("rgb", (128, 128, 128))))
            f.write(struct.pack("<d", p.get("error", 0.0)))
            f.write(struct.pack("<Q", len(p["track"])))
            for a, b in p["track"]:
                f.write(struct.pack("<II", a, b))


# ---------------------------------------------------------------------------------------------------------------------
# patch-match.cfg
# ---------------------------------------------------------------------------------------------------------------------
def read_problems(config_text: str, model: Model, min_triangulation_angle: float = 1.0) -> List[Tuple[int, List[int]]]:
    """PatchMatchController::ReadProblems: [(ref_image_idx, [src_image_idx ...]) ...]."""
    L = _lib()
    c, keep = model._c()
    names = (ctypes.c_char_p * max(len(model._names), 1))(*[n.encode() for n in model._names])
    np_, ns = ctypes.c_size_t(), ctypes.c_size_t()
    txt = config_text.encode()
    _check(L.b200ws_read_problems(txt, ctypes.byref(c), names, min_triangulation_angle, None, None, None, 0, 0,
                                  ctypes.byref(np_), ctypes.byref(ns)), "read patch-match.cfg")
    ref = np.empty(max(np_.value, 1), np.int32); off = np.empty(np_.value + 1, np.int64); src = np.empty(max(ns.value, 1), np.int32)
    _check(L.b200ws_read_problems(txt, ctypes.byref(c), names, min_triangulation_angle, ref.ctypes.data_as(_i32p),
                                  off.ctypes.data_as(_i64p), src.ctypes.data_as(_i32p), max(np_.value, 1), max(ns.value, 1),
                                  ctypes.byref(np_), ctypes.byref(ns)), "read patch-match.cfg")
    return [(int(ref[k]), [int(v) for v in src[off[k]:off[k + 1]]]) for k in range(np_.value)]


# ---------------------------------------------------------------------------------------------------------------------
# controller
# ---------------------------------------------------------------------------------------------------------------------
def _read_gray(path: str, max_image_size: int = -1) -> np.ndarray:
    from PIL import Image as PILImage
    im = PILImage.open(path).convert("L")
    if max_image_size > 0 and max(im.size) > max_image_size:
        s = max_image_size / max(im.size)
        im = im.resize((max(1, int(round(im.size[0] * s))), max(1, int(round(im.size[1] * s)))), PILImage.BILINEAR)
    return np.ascontiguousarray(np.asarray(im, np.uint8))


Runner = Callable[[PatchMatchOptions, Problem], dict]


def _default_runner(options: PatchMatchOptions, problem: Problem) -> dict:
    pm = PatchMatch(options, problem)
    pm.Run()
    out = dict(depth=pm.GetDepthMap(), normal=pm.GetNormalMap())
    if options.write_consistency_graph:
        out["consistency"] = pm.GetConsistentImageIdxs()
    pm.close()
    return out


class PatchMatchController:
    """PatchMatchController (patch_match.cc:156-536): reads the workspace (`sparse/`, `images/`, `stereo/patch-match.cfg`),
    runs the photometric problems, then - with geom_consistency - the geometric ones on the photometric maps, and writes
    `stereo/{depth_maps,normal_maps,consistency_graphs}/<image>.<photometric|geometric>.bin`.  Existing outputs are
    skipped like in the reference.  `runner(options, problem) -> dict(depth, normal[, consistency])` defaults to the
    CUDA path; tests inject a CPU stand-in."""

    def __init__(self, options: PatchMatchOptions, workspace_path: str, config_path: str = "", stereo_folder: str = "stereo"):
        self.options, self.workspace_path, self.config_path, self.stereo_folder = options, workspace_path, config_path, stereo_folder
        self.model: Optional[Model] = None
        self.problems: List[Tuple[int, List[int]]] = []
        self.depth_ranges: List[Tuple[float, float]] = []
        self._bitmaps: Dict[int, np.ndarray] = {}
        self.gpu_indices: List[int] = []

    # -- ReadWorkspace / ReadProblems
    def ReadWorkspace(self):
        self.model = Model.ReadFromCOLMAP(self.workspace_path)
        self.depth_ranges = self.model.ComputeDepthRanges()

    def ReadProblems(self):
        path = self.config_path or os.path.join(self.workspace_path, self.stereo_folder, "patch-match.cfg")
        with open(path) as f:
            self.problems = read_problems(f.read(), self.model, self.options.min_triangulation_angle)

    def ReadGpuIndices(self, num_devices: Optional[int] = None) -> List[int]:
        """PatchMatchController::ReadGpuIndices (patch_match.cc:375-384): one worker thread per entry of
        `options.gpu_index`; "-1" = every visible device.  Listing a device twice ("0,0") runs two problems on it at a
        time - the latency-bound serial pass of one overlaps the pixel pass of the other."""
        idx = [int(v) for v in str(self.options.gpu_index).replace(";", ",").split(",") if v.strip()]
        if len(idx) == 1 and idx[0] == -1:
            if num_devices is None:
                try:
                    import torch
                    num_devices = torch.cuda.device_count()
                except Exception:
                    num_devices = 0
            idx = list(range(max(int(num_devices), 1)))
        self.gpu_indices = idx
        return idx

    def _out_paths(self, image_name: str, output_type: str):
        base = os.path.join(self.workspace_path, self.stereo_folder)
        fn = f"{image_name}.{output_type}.bin"
        return (os.path.join(base, "depth_maps", fn), os.path.join(base, "normal_maps", fn), os.path.join(base, "consistency_graphs", fn))

    def _bitmap(self, idx: int) -> np.ndarray:
        if idx not in self._bitmaps:
            self._bitmaps[idx] = _read_gray(self.model.images[idx].path, self.options.max_image_size)
        return self._bitmaps[idx]

    def _image(self, idx: int) -> Image:
        mi = self.model.images[idx]
        bm = self._bitmap(idx)
        K = mi.K.astype(np.float64).copy()
        if bm.shape[1] != mi.width or bm.shape[0] != mi.height:   # Image::Rescale (mvs/image.cc:66-95)
            sx, sy = bm.shape[1] / mi.width, bm.shape[0] / mi.height
            K[0, 0] *= sx; K[0, 2] *= sx; K[1, 1] *= sy; K[1, 2] *= sy
        return Image(bitmap=bm, K=K.astype(np.float32), R=mi.R, T=mi.T)

    def _process(self, options: PatchMatchOptions, problem_idx: int, runner: Runner, gpu_index: int = -1):
        ref, srcs = self.problems[problem_idx]
        output_type = "geometric" if options.geom_consistency else "photometric"
        name = self.model.GetImageName(ref)
        dp, npth, gp = self._out_paths(name, output_type)
        if os.path.exists(dp) and os.path.exists(npth) and (not options.write_consistency_graph or os.path.exists(gp)):
            return False
        o = PatchMatchOptions(**{**options.__dict__})
        if o.depth_min < 0 or o.depth_max < 0:
            o.depth_min, o.depth_max = self.depth_ranges[ref]
            if not (o.depth_min > 0 and o.depth_max > 0):
                raise WorkspaceError("You must manually set the minimum and maximum depth, since no sparse model is provided in the workspace.")
        if o.sigma_spatial <= 0:
            o.sigma_spatial = float(o.window_radius)
        o.gpu_index = str(gpu_index)
        # used_image_idxs is a set in the reference (patch_match.cc:451-458): duplicates collapse, the reference image is
        # not a source of itself, and the consistency threshold cannot exceed the number of sources (a single-source
        # problem with the default of 2 would otherwise filter every pixel).  Sources keep their cfg order (the
        # reference iterates a hash set, i.e. an unspecified order).
        srcs = [g for g in dict.fromkeys(srcs) if g != ref]
        o.filter_min_num_consistent = min(len(srcs), o.filter_min_num_consistent)
        used = [ref] + srcs
        local = {g: k for k, g in enumerate(used)}
        images = [self._image(g) for g in used]
        problem = Problem(ref_image_idx=0, src_image_idxs=[local[g] for g in srcs], images=images)
        if options.geom_consistency:   # the photometric maps of the reference and of every source image (workspace.cc:208-240)
            problem.depth_maps, problem.normal_maps = [], []
            for g in used:
                pdp, pnp, _ = self._out_paths(self.model.GetImageName(g), "photometric")
                problem.depth_maps.append(read_depth_map(pdp))
                problem.normal_maps.append(read_normal_map(pnp))
        out = runner(o, problem)
        for d in (os.path.dirname(dp), os.path.dirname(npth), os.path.dirname(gp)):
            os.makedirs(d, exist_ok=True)
        write_mat(dp, out["depth"])
        write_mat(npth, out["normal"])
        if options.write_consistency_graph:
            # the sweep reports local source indices; the file stores global image indices (patch_match_cuda.cu:1381-1386)
            rec = np.asarray(out.get("consistency", np.zeros(0, np.int32)), np.int32).copy()
            i = 0
            while i < rec.size:
                n = int(rec[i + 2])
                rec[i + 3:i + 3 + n] = [used[int(v)] for v in rec[i + 3:i + 3 + n]]
                i += 3 + n
            ConsistencyGraph(out["depth"].shape[1], out["depth"].shape[0], rec).Write(gp)
        return True

    def Run(self, runner: Optional[Runner] = None, rank: int = 0, world: int = 1) -> int:
        """Returns the number of problems THIS rank processed (skipped outputs not counted).

        Multi-GPU: one process per GPU; every rank reads the workspace, takes its share of the
        problems (largest reference image first onto the least loaded rank, `assign_problems`) and writes its maps.
        The photometric problems need no communication; the geometric phase reads the photometric maps of its source
        images, which other ranks may have produced, so the phases are separated by one barrier - the same hand-over
        through the workspace files the reference uses between its two thread-pool passes (patch_match.cc:182-205).
        (`colmap_b200/workspace.py` is the variant that keeps the maps on the GPUs and all-gathers them over NCCL.)"""
        import threading
        from concurrent.futures import ThreadPoolExecutor
        from .sharding import assign_problems
        runner = runner or _default_runner
        self.ReadWorkspace()
        self.ReadProblems()
        gpu_indices = getattr(self, "gpu_indices", None) or self.ReadGpuIndices()
        tls = threading.local()
        free = list(gpu_indices)
        lock = threading.Lock()

        def run_phase(options, indices):   # the reference's thread pool: one thread per GPU-index entry
            def work(k):
                if not hasattr(tls, "gpu"):
                    with lock:
                        tls.gpu = free.pop(0)
                return bool(self._process(options, k, runner, tls.gpu))
            if len(gpu_indices) == 1:
                return sum(bool(self._process(options, k, runner, gpu_indices[0])) for k in indices)
            with lock:
                free[:] = list(gpu_indices)
            with ThreadPoolExecutor(max_workers=len(gpu_indices)) as pool:
                for t in list(vars(tls)):
                    delattr(tls, t)
                return sum(pool.map(work, indices))
        costs = [float(self.model.images[ref].width * self.model.images[ref].height * max(len(srcs), 1)) for ref, srcs in self.problems]
        mine = assign_problems(costs, world)[rank]

        def barrier():
            if world > 1:
                import torch.distributed as dist
                dist.barrier()

        done = 0
        if self.options.geom_consistency:
            photo = PatchMatchOptions(**{**self.options.__dict__})
            photo.geom_consistency = False
            photo.filter = False
            done += run_phase(photo, mine)
            barrier()
        done += run_phase(self.options, mine)
        barrier()
        return done


# ---------------------------------------------------------------------------------------------------------------------
# b200pm_run_workspace: the controller in C++ (include/b200_patch_match.h), image decoding through a callback
# ---------------------------------------------------------------------------------------------------------------------
_LOAD_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                            ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8)))


class _CWorkspace(ctypes.Structure):
    _fields_ = [("workspace_path", ctypes.c_char_p), ("stereo_folder", ctypes.c_char_p), ("config_path", ctypes.c_char_p),
                ("gpu_indices", ctypes.POINTER(ctypes.c_int)), ("num_gpu_indices", ctypes.c_int),
                ("write_consistency_graph", ctypes.c_int), ("load_gray", _LOAD_FN), ("load_gray_user", ctypes.c_void_p),
                ("rank", ctypes.c_int), ("world_size", ctypes.c_int), ("phase", ctypes.c_int)]


def run_workspace(options: PatchMatchOptions, workspace_path: str, config_path: str = "", stereo_folder: str = "stereo",
                  gpu_indices: Optional[Sequence[int]] = None, rank: int = 0, world: int = 1, phase: int = 0) -> int:
    """PatchMatchController::Run in C++ (b200pm_run_workspace): reads the workspace, runs every problem on the GPU(s),
    writes the map files, skips outputs that exist.  Images are decoded by PIL through the entry's callback - the place
    where COLMAP's adapter calls Bitmap::Read.  Returns the number of problems processed."""
    from ._lib import load_library
    from .patch_match import _COptions
    lib = load_library()
    lib.b200pm_run_workspace.argtypes = [ctypes.POINTER(_COptions), ctypes.POINTER(_CWorkspace), ctypes.POINTER(ctypes.c_int)]
    lib.b200pm_last_error.restype = ctypes.c_char_p
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p
    libc.malloc.argtypes = [ctypes.c_size_t]

    def load(_user, path, pw, ph, pdata):
        try:
            bm = _read_gray(path.decode(), options.max_image_size)
        except Exception:
            return -1
        buf = libc.malloc(bm.size)                       # released by the library with free()
        ctypes.memmove(buf, bm.ctypes.data, bm.size)
        pw[0], ph[0] = int(bm.shape[1]), int(bm.shape[0])
        pdata[0] = ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8))
        return 0
    cb = _LOAD_FN(load)
    gi = list(gpu_indices) if gpu_indices is not None else []
    arr = (ctypes.c_int * max(len(gi), 1))(*gi)
    w = _CWorkspace(workspace_path.encode(), stereo_folder.encode(), config_path.encode() if config_path else None,
                    arr if gi else None, len(gi), int(options.write_consistency_graph), cb, None, rank, world, phase)
    co = options.to_c()
    n = ctypes.c_int(0)
    rc = lib.b200pm_run_workspace(ctypes.byref(co), ctypes.byref(w), ctypes.byref(n))
    if rc != 0:
        raise WorkspaceError(f"b200pm_run_workspace failed ({rc}): {lib.b200pm_last_error().decode()}")
    return n.value
