"""Generator of the committed golden fixtures in tests/golden/ (test infrastructure).

  python tests/golden/make_golden.py              # oracle fixtures (CPU, anywhere)
  python tests/golden/make_golden.py --reference  # + the fixture produced by the UNMODIFIED reference PatchMatchCuda
                                                  #   (oracle/_ref/libpm_ref.so; needs a GPU, run under gpurun)

Fixtures
  reference_known_answers.json   vectors transcribed from the reference's own tests (file:line in every entry); this
                                 file is hand-written, the script only validates that it parses.
  pm_case_96x64.npz              a seeded 1 ref + 4 src PatchMatch problem (bitmaps, K/R/T, depth range, options) and
                                 the oracle's depth / normal / selection-probability maps for it.  The CUDA path must
                                 reproduce them bit for bit (tests/test_pm_gpu.py) and the oracle must keep producing
                                 them (tests/test_pm_cpu.py) - the fp32 contract is frozen by this file.
  pm_reference_cuda_160x120.npz  depth / normal maps the reference's own CUDA implementation produced for a seeded
                                 problem on a B200 (written by --reference).  Not bit-defined (texture filtering,
                                 --use_fast_math): tests compare statistics (tests/test_pm_gpu.py).
  ba_case_small.npz              a seeded BA problem (8 images, 120 points) and the oracle's solution / summary.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pm_problem_arrays(sc):
    imgs = sc["images"]
    return dict(bitmaps=np.stack([im.bitmap for im in imgs]), K=np.stack([im.K for im in imgs]),
                R=np.stack([im.R for im in imgs]), T=np.stack([im.T for im in imgs]),
                depth_min=np.float64(sc["depth_min"]), depth_max=np.float64(sc["depth_max"]),
                depth_gt=sc["depth_gt"].astype(np.float32))


def pm_problem_from_arrays(z):
    from colmap_b200.patch_match import Image, Problem
    images = [Image(bitmap=np.ascontiguousarray(z["bitmaps"][i]), K=z["K"][i], R=z["R"][i], T=z["T"][i])
              for i in range(len(z["bitmaps"]))]
    return Problem(ref_image_idx=0, src_image_idxs=list(range(1, len(images))), images=images)


def make_pm_oracle_case():
    import oracle_pm
    from colmap_b200.patch_match import PatchMatchOptions
    from colmap_b200.synthetic import make_patch_match_scene
    sc = make_patch_match_scene(96, 64, 4, seed=3)
    opts = dict(geom_consistency=False, num_iterations=2, filter=True)
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], **opts)
    out = oracle_pm.run(o, sc["problem"])
    np.savez_compressed(os.path.join(HERE, "pm_case_96x64.npz"), **pm_problem_arrays(sc),
                        options=json.dumps(opts), depth=out["depth"], normal=out["normal"], sel_prob=out["sel_prob"],
                        mask=out["mask"])
    print("pm_case_96x64.npz: valid", float((out["depth"] > 0).mean()))


def make_pm_reference_case():
    import ref_pm
    from colmap_b200.patch_match import PatchMatchOptions
    from colmap_b200.synthetic import make_patch_match_scene
    if not ref_pm.available():
        raise SystemExit("oracle/_ref/libpm_ref.so missing: run oracle/build_ref.sh where /root/reference exists")
    sc = make_patch_match_scene(160, 120, 4, seed=5)
    opts = dict(geom_consistency=False, num_iterations=5, filter=True)
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], **opts)
    out = ref_pm.run(o, sc["problem"])
    dst = os.path.join(ROOT, "gpurun_out") if os.environ.get("GRAFT_REPO_ROOT") else HERE
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "pm_reference_cuda_160x120.npz"), **pm_problem_arrays(sc),
                        options=json.dumps(opts), depth=out["depth"], normal=out["normal"])
    print("pm_reference_cuda_160x120.npz ->", dst, "valid", float((out["depth"] > 0).mean()))


def make_ba_case():
    import oracle_ba
    from colmap_b200.bundle_adjustment import BundleAdjustmentOptions, ITERATIVE_SCHUR, SIMPLE_RADIAL, PINHOLE
    from colmap_b200.synthetic import synthesize_ba_problem
    gt, noisy = synthesize_ba_problem(8, 120, 4, models=(SIMPLE_RADIAL, PINHOLE), seed=11)
    noisy.pose_constant = noisy.pose_constant.copy(); noisy.pose_fixed_dim = noisy.pose_fixed_dim.copy()
    noisy.pose_constant[0] = 1
    noisy.pose_fixed_dim[1] = int(np.argmax(np.abs(noisy.poses[1, 4:] - noisy.poses[0, 4:])))
    o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR, max_num_iterations=30)
    sol = noisy.copy(); sol.pose_constant, sol.pose_fixed_dim = noisy.pose_constant, noisy.pose_fixed_dim
    s = oracle_ba.solve(o, sol)
    np.savez_compressed(
        os.path.join(HERE, "ba_case_small.npz"),
        poses=noisy.poses, pose_constant=noisy.pose_constant, pose_fixed_dim=noisy.pose_fixed_dim,
        cam_model=noisy.cam_model, cam_off=noisy.cam_off, cam_params=noisy.cam_params, cam_constant=noisy.cam_constant,
        points=noisy.points, point_constant=noisy.point_constant, obs_pose=noisy.obs_pose, obs_cam=noisy.obs_cam,
        obs_point=noisy.obs_point, obs_xy=noisy.obs_xy,
        sol_poses=sol.poses, sol_cam_params=sol.cam_params, sol_points=sol.points,
        initial_cost=np.float64(s.initial_cost), final_cost=np.float64(s.final_cost),
        num_residuals=np.int64(s.num_residuals), num_effective_parameters=np.int64(s.num_effective_parameters),
        max_num_iterations=np.int64(30))
    print("ba_case_small.npz: cost", s.initial_cost, "->", s.final_cost, "steps", s.num_successful_steps, s.num_unsuccessful_steps)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    json.load(open(os.path.join(HERE, "reference_known_answers.json")))
    if a.reference:
        make_pm_reference_case()
    else:
        make_pm_oracle_case()
        make_ba_case()
