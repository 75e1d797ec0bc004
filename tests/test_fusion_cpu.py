"""Stereo fusion (include/b200_mvs_fusion.h, colmap_b200/mvs_fusion.py): the reference's integration scenario
(mvs/fusion_test.cc:45-140), the visibility file tests (:142-177), option validation, and point-for-point parity with the
pure-Python oracle (oracle/ws_oracle.py) on rendered multi-view scenes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ws_oracle  # noqa: E402

from colmap_b200.mvs_fusion import (FusionImage, ReadPointsVisibility, StereoFusion, StereoFusionOptions,  # noqa: E402
                                    WritePointsVisibility, fuse, write_ply)
from colmap_b200.mvs_workspace import WorkspaceError, write_mat, write_model_binary  # noqa: E402
from colmap_b200.synthetic import make_patch_match_scene  # noqa: E402


def _oracle_inputs(o, images):
    opt = dict(min_num_pixels=o.min_num_pixels, max_num_pixels=o.max_num_pixels, max_traversal_depth=o.max_traversal_depth,
               max_reproj_error=o.max_reproj_error, max_depth_error=o.max_depth_error, max_normal_error=o.max_normal_error,
               bbox_min=o.bounding_box[0], bbox_max=o.bounding_box[1])
    ims = [dict(used=im.used, K=im.K, R=im.R, T=im.T, image_width=im.image_width, image_height=im.image_height,
                depth=im.depth_map, normal=im.normal_map, rgb=im.bitmap, mask=im.mask) for im in images]
    return opt, ims


def _scene(width, height, n_src, seed, noise=0.0):
    """Views with ground-truth depth / normal maps (camera frame) of the analytic height field + colour bitmaps."""
    sc = make_patch_match_scene(width, height, n_src, seed=seed, with_gt_maps=True)
    rng = np.random.default_rng(seed)
    images = []
    for k, im in enumerate(sc["images"]):
        d = sc["depth_maps"][k].astype(np.float32)
        if noise:
            d = (d * (1 + noise * rng.standard_normal(d.shape))).astype(np.float32)
        d[:2, :] = 0.0                                      # some filtered pixels
        rgb = np.stack([im.bitmap, 255 - im.bitmap, (im.bitmap // 2)], -1).astype(np.uint8)
        images.append(FusionImage(im.K, im.R, im.T, width, height, d, sc["normal_maps"][k].astype(np.float32), rgb))
    overlap = [[j for j in range(len(images)) if j != i] for i in range(len(images))]
    return images, overlap


@pytest.mark.parametrize("seed,noise,kw", [(0, 0.0, {}), (1, 0.002, dict(min_num_pixels=2, max_traversal_depth=3)),
                                           (2, 0.004, dict(min_num_pixels=2, max_num_pixels=4, max_normal_error=25.0, max_depth_error=0.02))])
def test_fusion_matches_the_oracle_point_for_point(seed, noise, kw):
    images, overlap = _scene(40, 30, 3, seed, noise)
    o = StereoFusionOptions(**kw)
    got = fuse(o, images, overlap)
    opt, ims = _oracle_inputs(o, images)
    xyz, nrm, rgb, vis = ws_oracle.fuse(opt, ims, overlap)
    assert len(got.xyz) == len(xyz) > 50
    assert np.array_equal(got.rgb, rgb) and [list(v) for v in got.visibility] == vis
    assert np.array_equal(got.xyz.view(np.uint32), xyz.view(np.uint32))          # same fp32 operations in the same order
    assert np.array_equal(got.normal.view(np.uint32), nrm.view(np.uint32))
    # the fused cloud lies on the analytic surface: every point re-projects into view 0 near that view's depth
    if noise == 0.0:
        im0 = images[0]
        pc = (np.asarray(im0.R, np.float64) @ got.xyz.T.astype(np.float64)).T + np.asarray(im0.T, np.float64)
        uv = (np.asarray(im0.K, np.float64) @ pc.T).T
        u, v = np.round(uv[:, 0] / uv[:, 2]).astype(int), np.round(uv[:, 1] / uv[:, 2]).astype(int)
        ok = (u >= 0) & (v >= 2) & (u < 40) & (v < 30)
        assert ok.mean() > 0.5
        assert np.median(np.abs(pc[ok, 2] - im0.depth_map[v[ok], u[ok]]) / pc[ok, 2]) < 5e-3
    assert np.allclose(np.linalg.norm(got.normal, axis=1), 1.0, atol=1e-6)


def test_masks_bounding_box_and_unused_images():
    images, overlap = _scene(32, 24, 2, 3)
    o = StereoFusionOptions(min_num_pixels=1)
    full = fuse(o, images, overlap)
    images[0].mask = np.ones((24, 32), bool)                       # reference view fully pre-masked
    images[2].used = False
    part = fuse(o, images, overlap)
    assert len(part.xyz) > 0 and all(set(v) <= {1} for v in part.visibility)     # only view 1 is left to contribute
    assert any(len(v) > 1 for v in full.visibility)
    opt, ims = _oracle_inputs(o, images)
    assert len(ws_oracle.fuse(opt, ims, overlap)[0]) == len(part.xyz)
    images[0].mask = None; images[2].used = True
    zs = full.xyz[:, 2]
    box = StereoFusionOptions(min_num_pixels=1, bounding_box=((-1e9, -1e9, float(np.median(zs))), (1e9, 1e9, 1e9)))
    half = fuse(box, images, overlap)
    assert 0 < len(half.xyz) < len(full.xyz) and np.all(half.xyz[:, 2] >= np.float32(np.median(zs)) - 1e-3)


def test_options_check_and_errors():
    assert StereoFusionOptions().Check()
    for bad in (dict(min_num_pixels=-1), dict(min_num_pixels=5, max_num_pixels=4), dict(max_traversal_depth=0), dict(max_reproj_error=-1.0),
                dict(max_depth_error=-0.1), dict(max_normal_error=-1.0), dict(check_num_images=0), dict(cache_size=0.0)):
        assert not StereoFusionOptions(**bad).Check(), bad
    images, overlap = _scene(16, 12, 1, 0)
    with pytest.raises(WorkspaceError):
        fuse(StereoFusionOptions(max_traversal_depth=0), images, overlap)
    with pytest.raises(WorkspaceError):
        fuse(StereoFusionOptions(), images, [[5], [0]])


def test_points_visibility_files(tmp_path):
    """fusion_test.cc:142-177 (RoundTrip, SizeMismatch) + the byte layout."""
    vis = [[0, 2, 5], [], [7], [1, 3]]
    path = str(tmp_path / "fused.ply.vis")
    WritePointsVisibility(path, vis)
    raw = open(path, "rb").read()
    expect = np.array([4], "<u8").tobytes() + b"".join(np.array([len(v)] + v, "<u4").tobytes() for v in vis)
    assert raw == expect
    back = ReadPointsVisibility(path, 4)
    assert [list(b) for b in back] == vis
    with pytest.raises(WorkspaceError):
        ReadPointsVisibility(path, 3)


def test_reference_integration_scenario(tmp_path):
    """mvs/fusion_test.cc:45-140: two overlapping 30x20 views, constant depth 5, normals +z, bitmap colour (0,64,128),
    min_num_pixels 1, max_num_pixels 100, max_traversal_depth 10, check_num_images 10: some points are fused, every one
    inside (-10, 10)^3, colour (0,64,128), unit normal, non-empty visibility."""
    from PIL import Image as PILImage
    tmp = str(tmp_path)
    w, h = 30, 20
    cams = {1: dict(model_id=0, width=w, height=h, params=[25.0, w / 2, h / 2])}
    images, points = {}, {}
    os.makedirs(os.path.join(tmp, "images")); os.makedirs(os.path.join(tmp, "stereo", "depth_maps")); os.makedirs(os.path.join(tmp, "stereo", "normal_maps"))
    for i in range(2):
        name = f"image{i}.png"
        images[i + 1] = dict(qvec=[1.0, 0, 0, 0], tvec=[-0.3 * i, 0.0, 0.0], camera_id=1, name=name)
        PILImage.fromarray(np.tile(np.array([0, 64, 128], np.uint8), (h, w, 1))).save(os.path.join(tmp, "images", name))
        write_mat(os.path.join(tmp, "stereo", "depth_maps", name + ".geometric.bin"), np.full((h, w), 5.0, np.float32))
        nm = np.zeros((3, h, w), np.float32); nm[2] = 1.0
        write_mat(os.path.join(tmp, "stereo", "normal_maps", name + ".geometric.bin"), nm)
    rng = np.random.default_rng(0)
    for p in range(30):
        points[p + 1] = dict(xyz=[rng.uniform(-1, 1), rng.uniform(-1, 1), 5.0], track=[(1, p), (2, p)])
    write_model_binary(os.path.join(tmp, "sparse"), cams, images, points)
    open(os.path.join(tmp, "stereo", "fusion.cfg"), "w").write("image0.png\nimage1.png\n")
    o = StereoFusionOptions(min_num_pixels=1, max_num_pixels=100, max_traversal_depth=10, check_num_images=10)
    f = StereoFusion(o, tmp, "COLMAP", "", "geometric")
    f.Run()
    pts, vis = f.GetFusedPoints(), f.GetFusedPointsVisibility()
    assert len(pts.xyz) > 0 and len(vis) == len(pts.xyz)
    assert np.all(np.abs(pts.xyz) < 10.0)
    assert np.all(pts.rgb == np.array([0, 64, 128], np.uint8))
    assert np.allclose((pts.normal ** 2).sum(1), 1.0, rtol=0, atol=4e-7)
    assert all(len(v) > 0 for v in vis) and any(len(v) == 2 for v in vis)      # the views overlap: some points seen by both
    ply = str(tmp_path / "fused.ply")
    write_ply(ply, pts)
    head = open(ply, "rb").read(64)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex ")
    assert os.path.getsize(ply) > 27 * len(pts.xyz)


def test_library_exports_every_declared_fusion_symbol():
    import re
    from colmap_b200 import load_library
    lib = load_library()
    hdr = open(os.path.join(ROOT, "include", "b200_mvs_fusion.h")).read()
    names = set(re.findall(r"\b(b200fuse_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
