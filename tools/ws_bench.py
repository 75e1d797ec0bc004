"""CPU timing of the MVS workspace helpers: map files and sparse-model statistics at workspace scale."""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colmap_b200.mvs_workspace import Model, ModelPoint, read_mat, write_mat

rng = np.random.default_rng(0)
normal = rng.random((3, 1080, 1920)).astype(np.float32)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "n.bin")
    t = time.time(); write_mat(p, normal); tw = time.time() - t
    t = time.time(); back = read_mat(p); tr = time.time() - t
    assert np.array_equal(back, normal)
    print(f"normal map 1920x1080x3 ({normal.nbytes/1e6:.1f} MB): write {tw*1e3:.1f} ms ({normal.nbytes/tw/1e9:.2f} GB/s), read {tr*1e3:.1f} ms ({normal.nbytes/tr/1e9:.2f} GB/s)")

n_img, n_pts = 200, 100000
m = Model()
for i in range(n_img):
    a = 2 * np.pi * i / n_img
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    m.add_image(f"{i}.jpg", 1920, 1080, np.eye(3), R, [0, 0, 5.0])
for _ in range(n_pts):
    xyz = rng.normal(size=3)
    first = int(rng.integers(0, n_img))
    m.points.append(ModelPoint(float(xyz[0]), float(xyz[1]), float(xyz[2]), [(first + k) % n_img for k in range(int(rng.integers(2, 9)))]))
t = time.time(); c, keep = m._c(); tm = time.time() - t
for name, fn in (("ComputeDepthRanges", m.ComputeDepthRanges), ("ComputeSharedPoints", m.ComputeSharedPoints),
                 ("ComputeTriangulationAngles(75)", lambda: m.ComputeTriangulationAngles(75.0)),
                 ("GetMaxOverlappingImages(20, 1 deg)", lambda: m.GetMaxOverlappingImages(20, 1.0))):
    t = time.time(); fn(); print(f"{name}: {1e3*(time.time()-t):.0f} ms  ({n_img} images, {n_pts} points; incl. {tm*1e3:.0f} ms Python->C marshalling)")

# ---- stereo fusion at a workspace-like size (ground-truth maps of the synthetic scene, 1 + 4 views)
from colmap_b200.mvs_fusion import FusionImage, StereoFusionOptions, fuse
from colmap_b200.synthetic import make_patch_match_scene
W, H = 640, 360
sc = make_patch_match_scene(W, H, 4, seed=0, with_gt_maps=True)
views = [FusionImage(im.K, im.R, im.T, W, H, sc["depth_maps"][k].astype(np.float32), sc["normal_maps"][k].astype(np.float32),
                     np.stack([im.bitmap] * 3, -1)) for k, im in enumerate(sc["images"])]
overlap = [[j for j in range(5) if j != i] for i in range(5)]
t = time.time(); pts = fuse(StereoFusionOptions(), views, overlap); dt = time.time() - t
print(f"stereo fusion: 5 views {W}x{H} ({5*W*H/1e6:.2f} Mpx of depth) -> {len(pts.xyz)} points in {dt*1e3:.0f} ms ({5*W*H/1e6/dt:.1f} Mpx/s, one host thread)")
