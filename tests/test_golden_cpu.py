"""The oracles against the committed golden fixtures (tests/golden/, generator: tests/golden/make_golden.py):
the reference's own known-answer vectors, and the frozen oracle outputs that the CUDA path is compared with."""
import json
import os

import numpy as np

import oracle_ba
import oracle_pm
from colmap_b200.bundle_adjustment import (BundleAdjustmentOptions, FlatProblem, ITERATIVE_SCHUR, SIMPLE_PINHOLE,
                                           PINHOLE)
from colmap_b200.patch_match import PatchMatchOptions, _f32p
from colmap_b200.synthetic import synthesize_ba_problem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KA = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _p(a):
    return a.ctypes.data_as(_f32p)


def test_known_answers_mvs_image():
    """mvs/image_test.cc vectors (see the fixture for file:line)."""
    L = oracle_pm.lib()
    c = KA["mvs_image"]["compose_projection_matrix"]
    K, R, T, P = _f32(c["K"]), _f32(c["R"]), _f32(c["T"]), np.empty(12, np.float32)
    L.pm_oracle_compose_projection_matrix(_p(K), _p(R), _p(T), _p(P))
    assert np.array_equal(P, _f32(c["P"]))
    c = KA["mvs_image"]["compute_projection_center"]
    R, T, C = _f32(c["R"]), _f32(c["T"]), np.empty(3, np.float32)
    L.pm_oracle_projection_center(_p(R), _p(T), _p(C))
    assert np.array_equal(C, _f32(c["C"]))
    L.pm_oracle_rotate_pose.argtypes = [_f32p, _f32p, _f32p]
    for c in KA["mvs_image"]["rotate_pose"]:
        RR, R, T = _f32(c["RR"]), _f32(c["R"]), _f32(c["T"])
        L.pm_oracle_rotate_pose(_p(RR), _p(R), _p(T))
        assert np.allclose(R, c["R_out"], atol=1e-7) and np.allclose(T, c["T_out"], atol=1e-7)


def test_known_answers_rotate_convention():
    """mvs/gpu_mat_test.cu: rotated(d, W-1-c, r) == original(d, r, c)."""
    L = oracle_pm.lib()
    rng = np.random.default_rng(0)
    for w, h, d in KA["gpu_mat_rotate"]["sizes"]:
        src = rng.random((d, h, w)).astype(np.float32)
        dst = np.empty((d, w, h), np.float32)
        L.pm_oracle_rotate_f32(_p(src), w, h, d, _p(dst))
        r, c = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        assert np.array_equal(dst[:, w - 1 - c, r], src[:, r, c])


def test_known_answers_reprojection_error():
    """cost_functions/reprojection_error_test.cc: residual = projection - observation."""
    c = KA["reprojection_error"]
    for case in c["cases"]:
        ok, res, *_ = oracle_ba.reproj(SIMPLE_PINHOLE, case["point"], c["pose"], case["params"], c["xy"])
        assert np.array_equal(res, np.asarray(case["residual"], np.float64)), case


def test_known_answers_bundle_adjustment_counts():
    """bundle_adjustment_test.cc MinimumTrackLength (594) and ConstantPoints3D (80, points bit-identical)."""
    from colmap_b200.bundle_adjustment import BundleAdjustmentConfig, flatten_reconstruction
    from colmap_b200.synthetic import flat_to_reconstruction
    cases = {c["name"]: c for c in KA["bundle_adjustment_counts"]["cases"]}
    c = cases["ConstantPoints3D"]
    gt, noisy = synthesize_ba_problem(c["num_images"], c["num_points"], c["track_length"], models=(PINHOLE,), seed=1)
    before = noisy.points.copy()
    s = oracle_ba.solve(BundleAdjustmentOptions(refine_points3D=False, max_num_iterations=5), noisy)
    assert s.num_residuals == c["num_residuals"]
    assert np.array_equal(noisy.points, before)
    c = cases["MinimumTrackLength"]
    gt, noisy = synthesize_ba_problem(c["num_images"], c["num_points"], c["track_length"], models=(PINHOLE,), seed=1)
    rec = flat_to_reconstruction(noisy)
    # delete one observation: its point is left with a two-observation track and drops out at min_track_length = 3
    pid = next(iter(rec.points3D))
    image_id, p2_idx = rec.points3D[pid].track.pop()
    rec.images[image_id].points2D[p2_idx].point3D_id = -1
    cfg = BundleAdjustmentConfig()
    for image_id in rec.images:
        cfg.AddImage(image_id)
    o = BundleAdjustmentOptions(min_track_length=c["min_track_length"], max_num_iterations=3)
    flat = flatten_reconstruction(o, cfg, rec)[0]
    s = oracle_ba.solve(o, flat)
    assert s.num_residuals == c["num_residuals"]
    from colmap_b200.bundle_adjustment import assemble_reconstruction
    flat_cpp = assemble_reconstruction(o, cfg, rec)[0]               # the C++ assembly gives the same count
    assert oracle_ba.solve(o, flat_cpp).num_residuals == c["num_residuals"]


def test_known_answers_option_defaults():
    d = KA["options_defaults"]
    o = PatchMatchOptions()
    for k, v in d["patch_match"].items():
        assert getattr(o, k) == v, k
    b = BundleAdjustmentOptions()
    for k, v in d["bundle_adjustment"].items():
        assert getattr(b, k) == v, k


def _pm_case():
    import make_golden  # noqa: F401  (tests/golden is put on sys.path by conftest)
    z = np.load(os.path.join(GOLDEN, "pm_case_96x64.npz"))
    problem = make_golden.pm_problem_from_arrays(z)
    o = PatchMatchOptions(depth_min=float(z["depth_min"]), depth_max=float(z["depth_max"]), **json.loads(str(z["options"])))
    return z, o, problem


def test_pm_oracle_reproduces_the_frozen_fixture():
    """The fp32 contract is frozen by pm_case_96x64.npz: any change of the oracle's arithmetic shows up here."""
    z, o, problem = _pm_case()
    out = oracle_pm.run(o, problem)
    for k in ("depth", "normal", "sel_prob"):
        assert np.array_equal(out[k].view(np.uint32), z[k].view(np.uint32)), k
    assert np.array_equal(out["mask"], z["mask"])
    assert (z["depth"] > 0).mean() > 0.5


def ba_case():
    z = np.load(os.path.join(GOLDEN, "ba_case_small.npz"))
    flat = FlatProblem(z["poses"], z["pose_constant"], z["pose_fixed_dim"], z["cam_model"], z["cam_off"], z["cam_params"],
                       z["cam_constant"], z["points"], z["point_constant"], z["obs_pose"], z["obs_cam"], z["obs_point"],
                       z["obs_xy"])
    o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR, max_num_iterations=int(z["max_num_iterations"]))
    return z, o, flat


def test_ba_oracle_reproduces_the_frozen_fixture():
    z, o, flat = ba_case()
    s = oracle_ba.solve(o, flat)
    assert s.num_residuals == int(z["num_residuals"]) and s.num_effective_parameters == int(z["num_effective_parameters"])
    assert abs(s.initial_cost - float(z["initial_cost"])) <= 1e-9 * float(z["initial_cost"])
    assert abs(s.final_cost - float(z["final_cost"])) <= 1e-7 * float(z["final_cost"])
    assert np.allclose(flat.poses, z["sol_poses"], rtol=0, atol=1e-7)
    assert np.allclose(flat.points, z["sol_points"], rtol=0, atol=1e-7)
    assert np.allclose(flat.cam_params, z["sol_cam_params"], rtol=1e-8, atol=1e-7)


def test_known_answers_partially_contained_tracks():
    """bundle_adjustment_ceres_test.cc:674-807: residual / parameter counts and which blocks stay constant when only part of
    a track lies inside the config, through the C++ assembly and the oracle."""
    from colmap_b200.bundle_adjustment import BundleAdjustmentConfig, SIMPLE_RADIAL, assemble_reconstruction
    from colmap_b200.synthetic import flat_to_reconstruction
    cases = {c["name"]: c for c in KA["bundle_adjustment_partial_tracks"]["cases"]}
    for name in ("PartiallyContainedTracks", "PartiallyContainedTracksForceToOptimizePoint"):
        gt, noisy = synthesize_ba_problem(3, 100, 3, models=(SIMPLE_RADIAL,), seed=3, point3D_stddev=0.0, translation_stddev=0.0,
                                          rotation_stddev_deg=0.0)
        rec = flat_to_reconstruction(noisy)
        im3 = rec.images[3]
        pid_var, pid_add_var, pid_add_const = im3.points2D[0].point3D_id, im3.points2D[1].point3D_id, im3.points2D[2].point3D_id
        rec.points3D[pid_var].track = [t for t in rec.points3D[pid_var].track if t != (3, 0)]      # DeleteObservation(3, 0)
        im3.points2D[0].point3D_id = -1
        cfg = BundleAdjustmentConfig()
        cfg.AddImage(1); cfg.AddImage(2)
        cfg.SetConstantRigFromWorldPose(1); cfg.SetConstantRigFromWorldPose(2)
        if name.endswith("ForceToOptimizePoint"):
            cfg.AddVariablePoint(pid_add_var); cfg.AddConstantPoint(pid_add_const)
        o = BundleAdjustmentOptions(max_num_iterations=20)
        flat, image_ids, camera_ids, point_ids = assemble_reconstruction(o, cfg, rec)
        before = flat.copy()
        s = oracle_ba.solve(o, flat)
        assert s.termination_type in (0, 1)
        assert s.num_residuals == cases[name]["num_residuals_reduced"]
        assert s.num_effective_parameters == cases[name]["num_effective_parameters_reduced"]
        # poses constant; cameras 1, 2 variable (f and k move, principal point fixed), camera 3 constant
        assert np.array_equal(flat.poses, before.poses)
        cp, cb = flat.cam_params.reshape(3, 4), before.cam_params.reshape(3, 4)
        assert np.array_equal(cp[2], cb[2]) and np.array_equal(cp[:2, 1:3], cb[:2, 1:3])
        assert np.all(cp[:2, 0] != cb[:2, 0]) and np.all(cp[:2, 3] != cb[:2, 3])
        moved = {point_ids[k] for k in np.nonzero(np.any(flat.points != before.points, axis=1))[0]}
        expect = {pid_var, pid_add_var} if name.endswith("ForceToOptimizePoint") else {pid_var}
        assert moved == expect


def test_known_answers_bundle_adjustment_scenarios():
    """The variable / constant matrix of bundle_adjustment_ceres_test.cc (twelve scenarios): residual and effective-parameter
    counts of the problem the C++ assembly builds, as counted by the oracle."""
    from colmap_b200.bundle_adjustment import (BundleAdjustmentConfig, SIMPLE_RADIAL, THREE_POINTS, TWO_CAMS_FROM_WORLD,
                                               assemble_reconstruction)
    from colmap_b200.synthetic import flat_to_reconstruction
    gauges = {"TWO_CAMS_FROM_WORLD": TWO_CAMS_FROM_WORLD, "THREE_POINTS": THREE_POINTS}
    for c in KA["bundle_adjustment_scenarios"]["cases"]:
        gt, noisy = synthesize_ba_problem(c["num_images"], 100, c["num_images"], models=(SIMPLE_RADIAL,), seed=5)
        rec = flat_to_reconstruction(noisy)
        cfg = BundleAdjustmentConfig()
        for i in range(1, c["num_images"] + 1):
            cfg.AddImage(i)
        for i in c.get("constant_poses", []): cfg.SetConstantRigFromWorldPose(i)
        for i in c.get("constant_cameras", []): cfg.SetConstantCamIntrinsics(i)
        for p in c.get("constant_points", []): cfg.AddConstantPoint(p)
        for p in c.get("ignored_points", []): cfg.IgnorePoint(p)
        if "gauge" in c:
            cfg.FixGauge(gauges[c["gauge"]])
        o = BundleAdjustmentOptions(max_num_iterations=2, **c.get("options", {}))
        flat = assemble_reconstruction(o, cfg, rec)[0]
        s = oracle_ba.solve(o, flat)
        assert (s.num_residuals, s.num_effective_parameters) == (c["num_residuals"], c["num_effective_parameters"]), c["name"]
        # the PRODUCT's own host flattening (what b200ba_solve reports in its summary) counts the same - no oracle involved
        from test_ba_cpu import _pack
        L = _pack(flat, o)
        assert (L["num_residuals"], L["num_effective_parameters"]) == (c["num_residuals"], c["num_effective_parameters"]), c["name"]


def _rig_scenario(c):
    from colmap_b200.bundle_adjustment import BundleAdjustmentConfig, SIMPLE_RADIAL, THREE_POINTS, TWO_CAMS_FROM_WORLD
    from colmap_b200.synthetic import synthesize_rig_reconstruction
    gauges = {"TWO_CAMS_FROM_WORLD": TWO_CAMS_FROM_WORLD, "THREE_POINTS": THREE_POINTS}
    rec = synthesize_rig_reconstruction(*c["dataset"], model=SIMPLE_RADIAL, seed=11, point2D_stddev=1.0)
    for i in c.get("delete_observations_of_images", []):        # Reconstruction::DeleteObservation for every point of the image
        for k, p2 in enumerate(rec.images[i].points2D):
            if p2.point3D_id >= 0:
                rec.points3D[p2.point3D_id].track = [t for t in rec.points3D[p2.point3D_id].track if t != (i, k)]
                p2.point3D_id = -1
    cfg = BundleAdjustmentConfig()
    for i in (sorted(rec.images) if c["images"] == "all" else c["images"]):
        cfg.AddImage(i)
    for f in c.get("constant_frames", []): cfg.SetConstantRigFromWorldPose(f)
    for s in c.get("constant_sensors", []): cfg.SetConstantSensorFromRigPose(s)
    if "gauge" in c:
        cfg.FixGauge(gauges[c["gauge"]])
    return rec, cfg, BundleAdjustmentOptions(max_num_iterations=2, **c.get("options", {}))


def test_known_answers_bundle_adjustment_rig_scenarios():
    """The rig scenarios of bundle_adjustment_ceres_test.cc (TwoViewRig, ManyViewRig*, the three FixGaugeWithTwoCamsFromWorld*
    sequences with rigs, the three-points fallback): residual / effective-parameter counts of the problem the C++ assembly
    builds, counted by the oracle and by the product's own host flattening; the C++ assembly equals the Python mirror."""
    from colmap_b200.bundle_adjustment import assemble_reconstruction, flatten_reconstruction
    from test_ba_cpu import _pack
    for c in KA["bundle_adjustment_rig_scenarios"]["cases"]:
        rec, cfg, o = _rig_scenario(c)
        flat_py = flatten_reconstruction(o, cfg, rec)[0]
        flat = assemble_reconstruction(o, cfg, rec)[0]
        for name in ("poses", "pose_constant", "pose_fixed_dim", "cam_constant", "point_constant", "obs_pose", "obs_cam", "obs_point",
                     "obs_xy", "sensors", "sensor_constant", "cam_sensor"):
            assert np.array_equal(getattr(flat, name), getattr(flat_py, name)), (c["name"], name)
        s = oracle_ba.solve(o, flat)
        assert s.num_effective_parameters == c["num_effective_parameters"], (c["name"], s.num_effective_parameters)
        if "num_residuals" in c:
            assert s.num_residuals == c["num_residuals"], c["name"]
        L = _pack(flat, o)
        assert L["num_effective_parameters"] == c["num_effective_parameters"], (c["name"], L["num_effective_parameters"])
