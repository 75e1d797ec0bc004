"""The MVS chain end to end on the CPU: a rendered synthetic workspace (COLMAP sparse model + PNG images +
patch-match.cfg / fusion.cfg) -> PatchMatchController with the PatchMatch ORACLE as its runner (test infrastructure:
the product runner is the CUDA path) -> depth / normal map files -> StereoFusion -> point cloud on the analytic surface."""
import os

import numpy as np

import oracle_pm
from colmap_b200.mvs_fusion import StereoFusion, StereoFusionOptions
from colmap_b200.mvs_workspace import PatchMatchController, read_depth_map, write_model_binary
from colmap_b200.patch_match import PatchMatchOptions
from colmap_b200.synthetic import make_patch_match_scene


def _R_to_quat_wxyz(R):
    R = np.asarray(R, np.float64)
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return [w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)]


def test_workspace_to_point_cloud(tmp_path):
    from PIL import Image as PILImage
    tmp = str(tmp_path)
    W, H, N = 80, 56, 3
    sc = make_patch_match_scene(W, H, N, seed=4, with_gt_maps=True)
    os.makedirs(os.path.join(tmp, "images")); os.makedirs(os.path.join(tmp, "stereo"))
    K = sc["images"][0].K
    cams = {1: dict(model_id=1, width=W, height=H, params=[float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])])}
    images, points = {}, {}
    for i, im in enumerate(sc["images"]):
        name = f"v{i}.png"
        images[i + 1] = dict(qvec=_R_to_quat_wxyz(im.R), tvec=[float(t) for t in im.T], camera_id=1, name=name)
        PILImage.fromarray(im.bitmap).save(os.path.join(tmp, "images", name))
    # sparse points: back-projected ground-truth depths of view 0, seen by every view (gives depth ranges + overlaps)
    rng = np.random.default_rng(0)
    R0, T0 = np.asarray(sc["images"][0].R, np.float64), np.asarray(sc["images"][0].T, np.float64)
    Kd = np.asarray(K, np.float64)
    for p in range(80):
        u, v = rng.integers(4, W - 4), rng.integers(4, H - 4)
        d = float(sc["depth_maps"][0][v, u])
        X = R0.T @ (d * np.linalg.inv(Kd) @ np.array([u, v, 1.0]) - T0)
        points[p + 1] = dict(xyz=[float(x) for x in X], track=[(i + 1, p) for i in range(N + 1)])
    write_model_binary(os.path.join(tmp, "sparse"), cams, images, points)
    with open(os.path.join(tmp, "stereo", "patch-match.cfg"), "w") as f:
        for i in range(N + 1):
            f.write(f"v{i}.png\n__auto__, {N}\n")
    open(os.path.join(tmp, "stereo", "fusion.cfg"), "w").write("".join(f"v{i}.png\n" for i in range(N + 1)))

    def runner(o, problem):     # the CPU oracle stands in for the CUDA sweep
        out = oracle_pm.run(o, problem)
        return dict(depth=out["depth"], normal=out["normal"])

    o = PatchMatchOptions(geom_consistency=False, num_iterations=3, window_radius=3, filter=True, filter_min_num_consistent=1)
    c = PatchMatchController(o, tmp)
    assert c.Run(runner) == N + 1
    lo, hi = c.depth_ranges[0]
    gt0 = sc["depth_maps"][0]
    assert lo < gt0.min() and hi > gt0.max()                                 # sparse-model depth range brackets the scene
    d0 = read_depth_map(os.path.join(tmp, "stereo", "depth_maps", "v0.png.photometric.bin"))
    valid = d0 > 0
    assert valid.mean() > 0.6 and np.median(np.abs(d0[valid] - gt0[valid]) / gt0[valid]) < 5e-3

    f = StereoFusion(StereoFusionOptions(min_num_pixels=2, max_reproj_error=2.0, max_depth_error=0.02), tmp, input_type="photometric")
    f.Run()
    pts = f.GetFusedPoints()
    assert len(pts.xyz) > 500 and any(len(v) >= 2 for v in f.GetFusedPointsVisibility())
    # fused points sit on the analytic surface as seen from view 0
    pc = (R0 @ pts.xyz.T.astype(np.float64)).T + T0
    uv = (Kd @ pc.T).T
    u, v = np.round(uv[:, 0] / uv[:, 2]).astype(int), np.round(uv[:, 1] / uv[:, 2]).astype(int)
    ok = (u >= 0) & (v >= 0) & (u < W) & (v < H)
    assert ok.mean() > 0.5
    assert np.median(np.abs(pc[ok, 2] - gt0[v[ok], u[ok]]) / pc[ok, 2]) < 1e-2
