"""Synthetic inputs for the two hot paths.

PatchMatch: an analytic textured height field rendered into 1 + N pinhole views (ground-truth depth
and normal are exact), cameras on an arc with 5-15 degree triangulation angles.

Bundle adjustment: the semantics of colmap::SynthesizeDataset / SynthesizeNoise
(src/colmap/scene/synthetic.cc:341-672, benchmark/runtime/bundle_adjustment.cc:62-83) with tracks
sampled directly (points on the unit sphere, cameras at radius 5 looking at the origin).
"""
import numpy as np

from .patch_match import Image, Problem


# --------------------------------------------------------------------------------------------------
# PatchMatch scene
# --------------------------------------------------------------------------------------------------
class _HeightField:
    """z = h(x, y) in world coordinates, cameras look along +z from z ~ 0."""

    def __init__(self, rng, z0=5.0):
        self.z0 = z0
        self.amp = 0.18
        self.kx, self.ky = 1.1, 0.9
        self.sx, self.sy = 0.12, -0.07  # global slant
        self.px, self.py = rng.uniform(0, 2 * np.pi, 2)

    def h(self, x, y):
        return (self.z0 + self.sx * x + self.sy * y
                + self.amp * np.sin(self.kx * x + self.px) * np.cos(self.ky * y + self.py))

    def grad(self, x, y):
        hx = self.sx + self.amp * self.kx * np.cos(self.kx * x + self.px) * np.cos(self.ky * y + self.py)
        hy = self.sy - self.amp * self.ky * np.sin(self.kx * x + self.px) * np.sin(self.ky * y + self.py)
        return hx, hy


class _Texture:
    """Band-limited procedural texture on the (x, y) surface parameterisation."""

    def __init__(self, rng, min_wavelength, num_waves=28):
        lam = min_wavelength * np.exp(rng.uniform(0.0, np.log(24.0), num_waves))
        ang = rng.uniform(0, np.pi, num_waves)
        self.fx = (2 * np.pi / lam * np.cos(ang)).astype(np.float32)
        self.fy = (2 * np.pi / lam * np.sin(ang)).astype(np.float32)
        self.ph = rng.uniform(0, 2 * np.pi, num_waves).astype(np.float32)
        self.amp = (lam / lam.max()) ** 0.35
        self.amp = (self.amp / np.sqrt(0.5 * np.sum(self.amp ** 2))).astype(np.float32)

    def __call__(self, x, y):
        acc = np.zeros_like(x, dtype=np.float32)
        for fx, fy, ph, a in zip(self.fx, self.fy, self.ph, self.amp):
            acc += a * np.sin(fx * x + fy * y + ph)
        return acc  # unit variance


def _look_at_R(center, target):
    """World->camera rotation for a camera at `center` looking at `target`, x right, y down."""
    z = target - center
    z = z / np.linalg.norm(z)
    up = np.array([0.0, -1.0, 0.0])
    x = np.cross(-up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z])  # rows


def _render(hf, tex, K, R, T, width, height, contrast=45.0):
    """Exact per-pixel ray / height-field intersection.  Returns (uint8 image, depth, normal_cam)."""
    C = -R.T @ T
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    dcx = (u - K[0, 2]) / K[0, 0]
    dcy = (v - K[1, 2]) / K[1, 1]
    # d_world = R^T (dcx, dcy, 1)
    dx = R[0, 0] * dcx + R[1, 0] * dcy + R[2, 0]
    dy = R[0, 1] * dcx + R[1, 1] * dcy + R[2, 1]
    dz = R[0, 2] * dcx + R[1, 2] * dcy + R[2, 2]
    t = (hf.z0 - C[2]) / dz
    for _ in range(12):
        t = (hf.h(C[0] + t * dx, C[1] + t * dy) - C[2]) / dz
    x = C[0] + t * dx
    y = C[1] + t * dy
    img = 128.0 + contrast * tex(x.astype(np.float32), y.astype(np.float32))
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    hx, hy = hf.grad(x, y)
    nw = np.stack([hx, hy, -np.ones_like(hx)])
    nw /= np.linalg.norm(nw, axis=0, keepdims=True)
    nc = np.einsum("ij,jhw->ihw", R, nw)
    return img, t.astype(np.float32), nc.astype(np.float32)


def make_patch_match_scene(width=1920, height=1080, num_src=8, seed=0, with_gt_maps=False):
    """Returns dict(images=[Image] (index 0 = reference), problem=Problem, depth_gt, normal_gt,
    depth_min, depth_max[, depth_maps, normal_maps for every view])."""
    rng = np.random.default_rng(seed)
    hf = _HeightField(rng)
    f = 0.9 * width
    K = np.array([[f, 0, (width - 1) / 2.0], [0, f, (height - 1) / 2.0], [0, 0, 1]], np.float64)
    pixel = hf.z0 / f
    tex = _Texture(rng, min_wavelength=5.0 * pixel)
    target = np.array([0.0, 0.0, hf.z0])
    # world frame deliberately not aligned with the reference camera
    centers = [np.array([0.05, -0.03, 0.0])]
    for k in range(num_src):
        ang = 2 * np.pi * (k + 0.37) / max(num_src, 1)
        tri = np.deg2rad(rng.uniform(5.0, 15.0))
        rad = hf.z0 * np.tan(tri)
        centers.append(centers[0] + np.array([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-0.15, 0.15)]))
    images, depths, normals = [], [], []
    for k, c in enumerate(centers):
        tgt = target + (0 if k == 0 else rng.uniform(-0.15, 0.15, 3) * np.array([1, 1, 0]))
        R = _look_at_R(c, tgt)
        T = -R @ c
        if k == 0 or with_gt_maps:
            img, d, n = _render(hf, tex, K, R, T, width, height)
        else:
            img, d, n = _render(hf, tex, K, R, T, width, height)
        images.append(Image(bitmap=img, K=K.astype(np.float32), R=R.astype(np.float32), T=T.astype(np.float32)))
        depths.append(d)
        normals.append(n)
    problem = Problem(ref_image_idx=0, src_image_idxs=list(range(1, num_src + 1)), images=images)
    dmin, dmax = float(depths[0].min()), float(depths[0].max())
    out = dict(images=images, problem=problem, depth_gt=depths[0], normal_gt=normals[0],
               depth_min=0.75 * dmin, depth_max=1.25 * dmax)
    if with_gt_maps:
        out["depth_maps"] = depths
        out["normal_maps"] = normals
    return out


def make_workspace_scene(width, height, num_images, seed=0, device=None):
    """A workspace-shaped scene (BASELINE config C4): `num_images` views of one textured height field, cameras on a ring
    (mutual triangulation angles of a few degrees), every view a reference image.  Rendered with torch on `device` when
    given (4K frames take seconds per view in numpy).  Returns dict(images=[Image], depth_gt=[...], depth_min, depth_max,
    centers)."""
    rng = np.random.default_rng(seed)
    hf = _HeightField(rng)
    f = 0.9 * width
    K = np.array([[f, 0, (width - 1) / 2.0], [0, f, (height - 1) / 2.0], [0, 0, 1]], np.float64)
    tex = _Texture(rng, min_wavelength=5.0 * hf.z0 / f)
    target = np.array([0.0, 0.0, hf.z0])
    centers = []
    for k in range(num_images):
        ang = 2 * np.pi * k / num_images
        rad = hf.z0 * np.tan(np.deg2rad(rng.uniform(4.0, 9.0)))
        centers.append(np.array([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-0.15, 0.15)]))
    images, depths = [], []
    for c in centers:
        R = _look_at_R(c, target + rng.uniform(-0.1, 0.1, 3) * np.array([1, 1, 0]))
        T = -R @ c
        if device is None:
            img, d, _ = _render(hf, tex, K, R, T, width, height)
        else:
            img, d = _render_torch(hf, tex, K, R, T, width, height, device)
        images.append(Image(bitmap=img, K=K.astype(np.float32), R=R.astype(np.float32), T=T.astype(np.float32)))
        depths.append(d)
    dmin, dmax = min(float(d.min()) for d in depths), max(float(d.max()) for d in depths)
    return dict(images=images, depth_gt=depths, depth_min=0.75 * dmin, depth_max=1.25 * dmax, centers=np.stack(centers))


def _render_torch(hf, tex, K, R, T, width, height, device, contrast=45.0):
    """_render on a torch device (same ray / height-field intersection and texture; bench inputs, not bit-pinned)."""
    import torch
    C = -R.T @ T
    dd = torch.float64
    u = torch.arange(width, dtype=dd, device=device)[None, :].expand(height, width)
    v = torch.arange(height, dtype=dd, device=device)[:, None].expand(height, width)
    dcx = (u - K[0, 2]) / K[0, 0]
    dcy = (v - K[1, 2]) / K[1, 1]
    dx = R[0, 0] * dcx + R[1, 0] * dcy + R[2, 0]
    dy = R[0, 1] * dcx + R[1, 1] * dcy + R[2, 1]
    dz = R[0, 2] * dcx + R[1, 2] * dcy + R[2, 2]

    def h(x, y):
        return hf.z0 + hf.sx * x + hf.sy * y + hf.amp * torch.sin(hf.kx * x + hf.px) * torch.cos(hf.ky * y + hf.py)
    t = (hf.z0 - C[2]) / dz
    for _ in range(12):
        t = (h(C[0] + t * dx, C[1] + t * dy) - C[2]) / dz
    x = (C[0] + t * dx).to(torch.float32)
    y = (C[1] + t * dy).to(torch.float32)
    acc = torch.zeros_like(x)
    for fx, fy, ph, a in zip(tex.fx, tex.fy, tex.ph, tex.amp):
        acc += float(a) * torch.sin(float(fx) * x + float(fy) * y + float(ph))
    img = torch.clamp(torch.round(128.0 + contrast * acc), 0, 255).to(torch.uint8)
    return img.cpu().numpy(), t.to(torch.float32).cpu().numpy()


# --------------------------------------------------------------------------------------------------
# Bundle adjustment scenes (colmap::SynthesizeDataset / SynthesizeNoise semantics, tracks sampled directly)
# --------------------------------------------------------------------------------------------------
def _quat_from_two_vectors(a, b):
    """Eigen::Quaterniond::FromTwoVectors(a, b) for unit vectors, (x, y, z, w); vectorised over rows of a."""
    a = a / np.linalg.norm(a, axis=-1, keepdims=True)
    b = np.broadcast_to(b / np.linalg.norm(b), a.shape)
    c = np.sum(a * b, axis=-1, keepdims=True)
    axis = np.cross(a, b)
    s = np.sqrt((1.0 + c) * 2.0)
    q = np.concatenate([axis / s, s * 0.5], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _quat_rotate(q, p):
    v = q[..., :3]
    w = q[..., 3:4]
    t = 2.0 * np.cross(v, p)
    return p + w * t + np.cross(v, t)


_MODEL_DEFAULTS = {0: [1280.0, 512.0, 384.0], 1: [1280.0, 1280.0, 512.0, 384.0], 2: [1280.0, 512.0, 384.0, 0.05],
                   3: [1280.0, 512.0, 384.0, 0.05, 0.01], 8: [1280.0, 512.0, 384.0, 0.05], 9: [1280.0, 512.0, 384.0, 0.05, 0.01],
                   # the twelve models beyond the radial pinhole family (sensor/models.h:533-800): mild distortion
                   4: [1280.0, 1270.0, 512.0, 384.0, 0.04, -0.01, 1e-3, -8e-4],                               # OPENCV
                   5: [1280.0, 1270.0, 512.0, 384.0, 0.03, -0.01, 2e-3, -1e-3],                               # OPENCV_FISHEYE
                   6: [1280.0, 1270.0, 512.0, 384.0, 0.04, -0.01, 1e-3, -8e-4, 2e-3, 0.01, -3e-3, 1e-3],      # FULL_OPENCV
                   7: [1280.0, 1270.0, 512.0, 384.0, 0.3],                                                    # FOV
                   10: [1280.0, 1270.0, 512.0, 384.0, 0.03, -0.01, 1e-3, -8e-4, 2e-3, -1e-3, 5e-4, -4e-4],    # THIN_PRISM_FISHEYE
                   11: [1280.0, 1270.0, 512.0, 384.0, 0.03, -0.01, 2e-3, -1e-3, 5e-4, -2e-4, 1e-3, -8e-4, 5e-4, -4e-4, 3e-4, -2e-4],
                   12: [1280.0, 512.0, 384.0, -0.02], 13: [1280.0, 1270.0, 512.0, 384.0, -0.02],              # (SIMPLE_)DIVISION
                   14: [1280.0, 512.0, 384.0], 15: [1280.0, 1270.0, 512.0, 384.0],                            # (SIMPLE_)FISHEYE
                   16: [1280.0, 1270.0, 512.0, 384.0, 0.5, 1.1],                                              # EUCM
                   17: [2048.0, 1024.0]}                                                                      # EQUIRECTANGULAR


def _project(model, params, pc):
    if model not in (0, 1, 2, 3, 8, 9):   # wide models: the product's own host formulas (b200ba_test_project_wide), point by point
        import ctypes
        from ._lib import load_library
        lib = load_library()
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.b200ba_test_project_wide.argtypes = [ctypes.c_int, f64p, f64p, f64p, f64p, f64p]
        out = np.empty((len(pc), 2))
        P = params.shape[-1]
        xy = np.zeros(2); Ju = np.zeros(6); Jp = np.zeros(2 * P)
        pcc = np.ascontiguousarray(pc, np.float64); prm = np.ascontiguousarray(params, np.float64)
        for k in range(len(pc)):
            ok = lib.b200ba_test_project_wide(int(model), prm[k].ctypes.data_as(f64p), pcc[k].ctypes.data_as(f64p), xy.ctypes.data_as(f64p),
                                              Ju.ctypes.data_as(f64p), Jp.ctypes.data_as(f64p))
            assert ok == 1, (model, pcc[k])
            out[k] = xy
        return out
    uu = pc[..., 0] / pc[..., 2]
    vv = pc[..., 1] / pc[..., 2]
    if model == 0:
        return np.stack([params[..., 0] * uu + params[..., 1], params[..., 0] * vv + params[..., 2]], -1)
    if model == 1:
        return np.stack([params[..., 0] * uu + params[..., 2], params[..., 1] * vv + params[..., 3]], -1)
    if model in (8, 9):   # equidistant fisheye: (uu, vv) <- (atan r / r)(uu, vv), then the radial polynomial
        r = np.sqrt(uu * uu + vv * vv)
        s = np.where(r > 1e-12, np.arctan(r) / np.maximum(r, 1e-300), 1.0)
        uu, vv = s * uu, s * vv
    r2 = uu * uu + vv * vv
    rad = params[..., 3] * r2 if model in (2, 8) else params[..., 3] * r2 + params[..., 4] * r2 * r2
    return np.stack([params[..., 0] * uu * (1 + rad) + params[..., 1], params[..., 0] * vv * (1 + rad) + params[..., 2]], -1)


def synthesize_ba_problem(num_images, num_points, track_length, models=(2,), shared_camera=False, seed=42,
                          point2D_stddev=1.0, point3D_stddev=0.05, translation_stddev=0.01, rotation_stddev_deg=1.0,
                          num_obs=None, track_lengths=None):
    """Ground truth + noisy flat BA problems.  `models`: camera model ids cycled over the images (one camera per
    image) or a single shared camera.  `num_obs` (optional) = exact observation count (tracks of length
    floor/ceil(num_obs/num_points)).  Noise defaults = benchmark/runtime/bundle_adjustment.cc:76-80."""
    from .bundle_adjustment import MODEL_NUM_PARAMS, FlatProblem
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1, 1, (num_points, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    view = -rng.uniform(-1, 1, (num_images, 3))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    q = _quat_from_two_vectors(view, np.array([0.0, 0.0, 1.0]))
    t = _quat_rotate(q, 5.0 * view)
    poses = np.concatenate([q, t], axis=1)
    if shared_camera:
        cam_model = [models[0]]
        img_cam = np.zeros(num_images, np.int32)
    else:
        cam_model = [models[i % len(models)] for i in range(num_images)]
        img_cam = np.arange(num_images, dtype=np.int32)
    cam_off, off = [], 0
    for m in cam_model:
        cam_off.append(off); off += MODEL_NUM_PARAMS[m]
    cam_params = np.concatenate([np.asarray(_MODEL_DEFAULTS[m]) for m in cam_model])
    # tracks
    if track_lengths is not None:
        lens = np.asarray(track_lengths, np.int64)
    elif num_obs is None:
        lens = np.full(num_points, track_length, np.int64)
    else:
        base = num_obs // num_points
        lens = np.full(num_points, base, np.int64)
        lens[: num_obs - base * num_points] += 1
    obs_point = np.repeat(np.arange(num_points, dtype=np.int32), lens)
    # distinct random images per point: random keys, take the `len` smallest per point (vectorised for equal lens)
    maxlen = int(lens.max())
    keys = rng.random((num_points, num_images)) if num_points * num_images <= 5e7 else None
    if keys is not None:
        order = np.argsort(keys, axis=1)[:, :maxlen]
        sel = order[np.arange(maxlen)[None, :] < lens[:, None]]
    else:
        # large problems: sample with replacement then de-duplicate by re-drawing collisions a few times
        cand = rng.integers(0, num_images, (num_points, maxlen))
        for _ in range(8):
            cs = np.sort(cand, axis=1)
            dup = np.zeros_like(cand, bool)
            dup[:, 1:] = cs[:, 1:] == cs[:, :-1]
            if not dup.any():
                break
            cand = cs
            cand[dup] = rng.integers(0, num_images, int(dup.sum()))
        sel = cand[np.arange(maxlen)[None, :] < lens[:, None]]
    obs_pose = sel.astype(np.int32)
    obs_cam = img_cam[obs_pose]
    # exact projections
    pc = _quat_rotate(poses[obs_pose, :4], pts[obs_point]) + poses[obs_pose, 4:]
    xy = np.empty((len(obs_pose), 2))
    cm = np.asarray(cam_model)[obs_cam]
    for m in set(cam_model):
        msk = cm == m
        offs = np.asarray(cam_off)[obs_cam[msk]]
        prm = cam_params[offs[:, None] + np.arange(MODEL_NUM_PARAMS[m])[None, :]]
        xy[msk] = _project(m, prm, pc[msk])
    n_img, n_cam = num_images, len(cam_model)
    gt = FlatProblem(poses, np.zeros(n_img, np.uint8), -np.ones(n_img, np.int8), cam_model, cam_off, cam_params,
                     np.zeros(n_cam, np.uint8), pts, np.zeros(num_points, np.uint8), obs_pose, obs_cam, obs_point, xy)
    noisy = gt.copy()
    # SynthesizeNoise (synthetic.cc:675-733)
    if rotation_stddev_deg > 0:
        ang = np.deg2rad(np.clip(rng.normal(0, rotation_stddev_deg, num_images), -180, 180))
        qz = np.stack([np.zeros(num_images), np.zeros(num_images), np.sin(ang / 2), np.cos(ang / 2)], 1)
        noisy.poses[:, :4] = _quat_mul(noisy.poses[:, :4], qz)
    if translation_stddev > 0:
        noisy.poses[:, 4:] += rng.normal(0, translation_stddev, (num_images, 3))
    if point2D_stddev > 0:
        noisy.obs_xy = noisy.obs_xy + rng.normal(0, point2D_stddev, noisy.obs_xy.shape)
    if point3D_stddev > 0:
        noisy.points += rng.normal(0, point3D_stddev, noisy.points.shape)
    return gt, noisy


def flat_to_reconstruction(flat):
    """FlatProblem -> minimal Reconstruction (ids are 1-based like COLMAP's)."""
    from .bundle_adjustment import MODEL_NUM_PARAMS, Camera, Image, Point2D, Point3D, Reconstruction
    rec = Reconstruction()
    for c in range(len(flat.cam_model)):
        n = MODEL_NUM_PARAMS[int(flat.cam_model[c])]
        rec.cameras[c + 1] = Camera(c + 1, int(flat.cam_model[c]), flat.cam_params[flat.cam_off[c]:flat.cam_off[c] + n].copy())
    img_cam = {}
    for o in range(len(flat.obs_pose)):
        img_cam[int(flat.obs_pose[o])] = int(flat.obs_cam[o])
    for i in range(len(flat.poses)):
        rec.images[i + 1] = Image(i + 1, img_cam.get(i, 0) + 1, flat.poses[i].copy())
    for p in range(len(flat.points)):
        rec.points3D[p + 1] = Point3D(flat.points[p].copy())
    for o in range(len(flat.obs_pose)):
        im = rec.images[int(flat.obs_pose[o]) + 1]
        pid = int(flat.obs_point[o]) + 1
        im.points2D.append(Point2D(flat.obs_xy[o].copy(), pid))
        rec.points3D[pid].track.append((im.image_id, len(im.points2D) - 1))
    return rec


def synthesize_rig_problem(num_frames, num_sensors, num_points, track_length, models=(2,), seed=42, point2D_stddev=1.0,
                           point3D_stddev=0.05, translation_stddev=0.01, rotation_stddev_deg=1.0, sensor_translation_stddev=0.005,
                           sensor_rotation_stddev_deg=0.3):
    """Rig scene (SynthesizeDataset with num_cameras_per_rig > 1, scene/synthetic.cc:341-672): ONE rig of `num_sensors`
    cameras (sensor 0 = reference sensor, the others carry a sensor_from_rig pose), `num_frames` frames (one
    rig_from_world pose each), every point seen from `track_length` random (frame, sensor) pairs.  In the flat problem
    obs_pose_idx is the FRAME and obs_camera_idx the sensor's camera.  Returns (ground truth, noisy) FlatProblems."""
    from .bundle_adjustment import MODEL_NUM_PARAMS, FlatProblem
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1, 1, (num_points, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    view = -rng.uniform(-1, 1, (num_frames, 3))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    q = _quat_from_two_vectors(view, np.array([0.0, 0.0, 1.0]))
    poses = np.concatenate([q, _quat_rotate(q, 5.0 * view)], axis=1)                 # rig_from_world per frame
    # sensor_from_rig of the non-reference sensors: a few degrees of rotation, a 0.1-0.3 baseline
    ns = num_sensors - 1
    ax = rng.normal(size=(ns, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = np.deg2rad(rng.uniform(2.0, 6.0, ns))
    sq = np.concatenate([ax * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    st = rng.uniform(-0.3, 0.3, (ns, 3))
    sensors = np.concatenate([sq, st], axis=1)
    cam_model = [models[i % len(models)] for i in range(num_sensors)]
    cam_off, off = [], 0
    for m in cam_model:
        cam_off.append(off); off += MODEL_NUM_PARAMS[m]
    cam_params = np.concatenate([np.asarray(_MODEL_DEFAULTS[m]) for m in cam_model])
    cam_sensor = np.array([-1] + list(range(ns)), np.int32)
    # tracks over (frame, sensor) slots
    slots = num_frames * num_sensors
    obs_point = np.repeat(np.arange(num_points, dtype=np.int32), track_length)
    keys = rng.random((num_points, slots))
    sel = np.argsort(keys, axis=1)[:, :track_length].reshape(-1)
    obs_pose = (sel // num_sensors).astype(np.int32)
    obs_cam = (sel % num_sensors).astype(np.int32)
    pr = _quat_rotate(poses[obs_pose, :4], pts[obs_point]) + poses[obs_pose, 4:]
    pc = pr.copy()
    nr = obs_cam > 0
    pc[nr] = _quat_rotate(sensors[obs_cam[nr] - 1, :4], pr[nr]) + sensors[obs_cam[nr] - 1, 4:]
    xy = np.empty((len(obs_pose), 2))
    for c in range(num_sensors):
        msk = obs_cam == c
        m = cam_model[c]
        prm = np.broadcast_to(cam_params[cam_off[c]:cam_off[c] + MODEL_NUM_PARAMS[m]], (int(msk.sum()), MODEL_NUM_PARAMS[m]))
        xy[msk] = _project(m, prm, pc[msk])
    gt = FlatProblem(poses, np.zeros(num_frames, np.uint8), -np.ones(num_frames, np.int8), cam_model, cam_off, cam_params,
                     np.zeros(num_sensors, np.uint8), pts, np.zeros(num_points, np.uint8), obs_pose, obs_cam, obs_point, xy)
    gt.set_sensors(sensors, np.zeros(ns, np.uint8), cam_sensor)
    noisy = gt.copy()
    if rotation_stddev_deg > 0:
        a = np.deg2rad(rng.normal(0, rotation_stddev_deg, num_frames))
        noisy.poses[:, :4] = _quat_mul(noisy.poses[:, :4], np.stack([0 * a, 0 * a, np.sin(a / 2), np.cos(a / 2)], 1))
    if translation_stddev > 0:
        noisy.poses[:, 4:] += rng.normal(0, translation_stddev, (num_frames, 3))
    if ns and sensor_rotation_stddev_deg > 0:
        a = np.deg2rad(rng.normal(0, sensor_rotation_stddev_deg, ns))
        noisy.sensors[:, :4] = _quat_mul(noisy.sensors[:, :4], np.stack([0 * a, 0 * a, np.sin(a / 2), np.cos(a / 2)], 1))
    if ns and sensor_translation_stddev > 0:
        noisy.sensors[:, 4:] += rng.normal(0, sensor_translation_stddev, (ns, 3))
    if point2D_stddev > 0:
        noisy.obs_xy = noisy.obs_xy + rng.normal(0, point2D_stddev, noisy.obs_xy.shape)
    if point3D_stddev > 0:
        noisy.points += rng.normal(0, point3D_stddev, noisy.points.shape)
    return gt, noisy


def synthesize_rig_reconstruction(num_rigs, num_cameras_per_rig, num_frames_per_rig, num_points3D, model=2, seed=0,
                                  point2D_stddev=0.0):
    """SynthesizeDataset with rigs (scene/synthetic.cc:341-672) as a Reconstruction of the Python mirror: `num_rigs` rigs of
    `num_cameras_per_rig` cameras each (own intrinsics; the first camera of a rig is its reference sensor), every rig
    with `num_frames_per_rig` frames, one image per (frame, sensor), image / camera / frame / rig ids ascending from 1 in
    that order, every point observed by every image (as in the reference's small test scenes)."""
    from .bundle_adjustment import Camera, Frame, Image, MODEL_NUM_PARAMS, Point2D, Point3D, Reconstruction, Rig, _rigid_compose
    rng = np.random.default_rng(seed)
    rec = Reconstruction()
    pts = rng.uniform(-1, 1, (num_points3D, 3))
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    for k in range(num_points3D):
        rec.points3D[k + 1] = Point3D(xyz=pts[k].copy())
    cam_id = img_id = frame_id = 0
    for r in range(1, num_rigs + 1):
        cams = []
        for _ in range(num_cameras_per_rig):
            cam_id += 1
            rec.cameras[cam_id] = Camera(cam_id, model, np.asarray(_MODEL_DEFAULTS[model], np.float64).copy())
            cams.append(cam_id)
        rig = Rig(r, cams[0])
        for c in cams[1:]:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            ang = np.deg2rad(rng.uniform(2.0, 6.0))
            rig.sensors[c] = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)], rng.uniform(-0.3, 0.3, 3)])
        rec.rigs[r] = rig
        for _ in range(num_frames_per_rig):
            frame_id += 1
            view = -rng.uniform(-1, 1, 3); view /= np.linalg.norm(view)
            q = _quat_from_two_vectors(view[None], np.array([0.0, 0.0, 1.0]))[0]
            rfw = np.concatenate([q, _quat_rotate(q[None], 5.0 * view[None])[0]])
            rec.frames[frame_id] = Frame(frame_id, r, rfw)
            for c in cams:
                img_id += 1
                cfw = rfw if c == rig.ref_sensor else _rigid_compose(rig.sensors[c], rfw)
                pc = _quat_rotate(np.broadcast_to(cfw[:4], (num_points3D, 4)), pts) + cfw[4:]
                prm = rec.cameras[c].params
                xy = _project(model, np.broadcast_to(prm, (num_points3D, MODEL_NUM_PARAMS[model])), pc)
                if point2D_stddev > 0:
                    xy = xy + rng.normal(0, point2D_stddev, xy.shape)
                im = Image(img_id, c, cfw.copy(), frame_id=frame_id)
                for k in range(num_points3D):
                    im.points2D.append(Point2D(xy[k].copy(), k + 1))
                    rec.points3D[k + 1].track.append((img_id, k))
                rec.images[img_id] = im
    return rec

