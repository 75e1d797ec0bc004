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
