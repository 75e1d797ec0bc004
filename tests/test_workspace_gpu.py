"""Workspace-level GPU tier (SURVEY 8a PatchMatchController::*, 8b b200pm_run_workspace, 8f-1): a rendered synthetic
workspace on disk (COLMAP sparse model + PNG images + patch-match.cfg) through
  (a) the C entry b200pm_run_workspace (controller in C++, CUDA sweep),
  (b) the Python mirror PatchMatchController with its default CUDA runner,
  (c) the same controller with the CPU ORACLE as runner (the checker),
two-phase geometric run with consistency graphs.  The three sets of output files must be byte-identical (the sweep is
bit-exact, the formats are fixed), and a second run must skip every problem (patch_match.cc:410-414)."""
import os
import shutil

import numpy as np
import pytest

import oracle_pm
from colmap_b200.mvs_workspace import PatchMatchController, run_workspace, write_model_binary
from colmap_b200.patch_match import PatchMatchOptions, consistency_list_from_mask
from colmap_b200.synthetic import make_patch_match_scene

pytestmark = pytest.mark.gpu


def _R_to_quat_wxyz(R):
    R = np.asarray(R, np.float64)
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    return [w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)]


def _write_workspace(tmp, W=72, H=54, N=3, seed=4):
    from PIL import Image as PILImage
    sc = make_patch_match_scene(W, H, N, seed=seed, with_gt_maps=True)
    os.makedirs(os.path.join(tmp, "images")); os.makedirs(os.path.join(tmp, "stereo"))
    K = sc["images"][0].K
    cams = {1: dict(model_id=1, width=W, height=H, params=[float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])])}
    images, points = {}, {}
    for i, im in enumerate(sc["images"]):
        name = f"v{i}.png"
        images[i + 1] = dict(qvec=_R_to_quat_wxyz(im.R), tvec=[float(t) for t in im.T], camera_id=1, name=name)
        PILImage.fromarray(im.bitmap).save(os.path.join(tmp, "images", name))
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
        f.write("# comment\nv0.png\n__auto__, 3\nv1.png\n__all__\nv2.png\nv0.png, v1.png, v3.png\nv3.png\nv2.png, v0.png\n")
    return sc


def _files(tmp):
    out = {}
    for kind in ("depth_maps", "normal_maps", "consistency_graphs"):
        d = os.path.join(tmp, "stereo", kind)
        for fn in sorted(os.listdir(d)) if os.path.isdir(d) else []:
            out[kind + "/" + fn] = open(os.path.join(d, fn), "rb").read()
    return out


def test_workspace_c_entry_python_controller_and_oracle_agree(tmp_path):
    base = str(tmp_path / "base")
    os.makedirs(base)
    _write_workspace(base)
    runs = {}
    o = PatchMatchOptions(geom_consistency=True, num_iterations=1, window_radius=3, filter=True, filter_min_num_consistent=1,
                          write_consistency_graph=True, gpu_index="0")
    for name in ("c_entry", "py_cuda", "py_oracle"):
        tmp = str(tmp_path / name)
        shutil.copytree(base, tmp)
        if name == "c_entry":
            n = run_workspace(o, tmp, gpu_indices=[0, 0])          # two problems in flight on device 0
            assert n == 8                                          # 4 photometric + 4 geometric problems
            assert run_workspace(o, tmp, gpu_indices=[0]) == 0     # everything exists: skipped
        elif name == "py_cuda":
            c = PatchMatchController(o, tmp)
            c.gpu_indices = [0]
            assert c.Run() == 8
            assert c.Run() == 0
        else:
            def runner(opt, problem):     # the CPU oracle stands in for the CUDA sweep
                out = oracle_pm.run(opt, problem)
                res = dict(depth=out["depth"], normal=out["normal"])
                if opt.write_consistency_graph:
                    res["consistency"] = consistency_list_from_mask(out["mask"], list(range(1, len(problem.src_image_idxs) + 1))) if opt.filter else np.zeros(0, np.int32)
                return res
            c = PatchMatchController(o, tmp)
            c.gpu_indices = [0]
            assert c.Run(runner) == 8
        runs[name] = _files(tmp)
    names = sorted(runs["py_oracle"])
    assert len(names) == 4 * 2 * 3                                  # 4 images x {photometric, geometric} x 3 kinds
    for other in ("c_entry", "py_cuda"):
        assert sorted(runs[other]) == names
        for fn in names:
            assert runs[other][fn] == runs["py_oracle"][fn], f"{other}: {fn} differs from the oracle pipeline"
    d = np.frombuffer(runs["c_entry"]["depth_maps/v0.png.geometric.bin"][len(b"72&54&1&"):], "<f4")
    assert (d > 0).mean() > 0.5
