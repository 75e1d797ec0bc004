"""CPU tier for the PatchMatch path: the oracle against the reference's own known answers, the host
side of the product against the oracle (bit level), and the C-ABI surface.  No GPU compute."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_pm
from colmap_b200 import load_library
from colmap_b200.patch_match import (PatchMatch, PatchMatchError, PatchMatchOptions, _CProblem, _bind, _f32p,
                                     consistency_list_from_mask, marshal)
from colmap_b200.synthetic import make_patch_match_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---------------------------------------------------------------- reference known answers
def test_rotate_convention_matches_reference_gpu_mat_test():
    """mvs/gpu_mat_test.cu:188-206 / cuda_rotate.h:66-73: out(row=W-1-x, col=y) = in(row=y, col=x)."""
    w, h = 7, 5
    a = np.arange(2 * w * h, dtype=np.float32).reshape(2, h, w)
    out = np.empty((2, w, h), np.float32)
    oracle_pm.lib().pm_oracle_rotate_f32(a.ctypes.data_as(_f32p), w, h, 2, out.ctypes.data_as(_f32p))
    for s in range(2):
        for y in range(h):
            for x in range(w):
                assert out[s, w - 1 - x, y] == a[s, y, x]
    # numpy: counter-clockwise rotation
    assert np.array_equal(out[0], np.rot90(a[0], 1))


def test_pose_helpers_known_answers_image_test():
    """mvs/image_test.cc:149-212: ComposeProjectionMatrix {0,2,0,2, 2,0,0,4, 0,0,1,3}, projection centre,
    inverse projection consistent with projection."""
    L = oracle_pm.lib()
    K = np.array([2, 0, 0, 0, 2, 0, 0, 0, 1], np.float32)   # K = diag(2,2,1)
    R = np.array([0, 1, 0, 1, 0, 0, 0, 0, 1], np.float32)
    T = np.array([1, 2, 3], np.float32)
    P = np.empty(12, np.float32)
    L.pm_oracle_compose_projection_matrix(K.ctypes.data_as(_f32p), R.ctypes.data_as(_f32p), T.ctypes.data_as(_f32p),
                                          P.ctypes.data_as(_f32p))
    assert np.array_equal(P, np.array([0, 2, 0, 2, 2, 0, 0, 4, 0, 0, 1, 3], np.float32))
    C = np.empty(3, np.float32)
    L.pm_oracle_projection_center(R.ctypes.data_as(_f32p), T.ctypes.data_as(_f32p), C.ctypes.data_as(_f32p))
    assert np.allclose(C, -R.reshape(3, 3).T @ T)
    iP = np.empty(12, np.float32)
    L.pm_oracle_compose_inverse_projection_matrix(K.ctypes.data_as(_f32p), R.ctypes.data_as(_f32p),
                                                  T.ctypes.data_as(_f32p), iP.ctypes.data_as(_f32p))
    P4 = np.vstack([P.reshape(3, 4), [0, 0, 0, 1]]).astype(np.float64)
    assert np.allclose(np.linalg.inv(P4)[:3], iP.reshape(3, 4), atol=1e-6)


def test_xorwow_matches_curand_reference_values():
    """XORWOW with cuRAND seeding; seed 0 must give the generator's published first outputs
    (Marsaglia's xorwow with cuRAND's salt: state v = {123456789+t0, ...}); checked against an independent
    pure-python statement of curand_kernel.h."""
    def py_stream(seed, n):
        M = 0xFFFFFFFF
        s0 = (seed & M) ^ 0xaad26b49
        s1 = ((seed >> 32) & M) ^ 0xf7dcefdd
        t0 = (1099087573 * s0) & M
        t1 = (2591861531 * s1) & M
        d = (6615241 + t1 + t0) & M
        v = [(123456789 + t0) & M, 362436069 ^ t0, (521288629 + t1) & M, 88675123 ^ t1, (5783321 + t0) & M]
        out = []
        for _ in range(n):
            t = v[0] ^ (v[0] >> 2)
            v = v[1:] + [((v[4] ^ ((v[4] << 4) & M)) ^ (t ^ ((t << 1) & M))) & M]
            d = (d + 362437) & M
            x = (v[4] + d) & M
            out.append(np.float32(np.float32(x) * np.float32(2.3283064365386963e-10)) + np.float32(1.1641532182693481e-10))
        return np.array(out, np.float32)

    for seed in (0, 1, 511, 123456, 2 ** 33 + 5):
        got = np.empty(16, np.float32)
        oracle_pm.lib().pm_oracle_rng_stream(seed, 16, got.ctypes.data_as(_f32p))
        assert np.array_equal(_bits(got), _bits(py_stream(seed, 16)))
        assert np.all(got > 0) and np.all(got <= 1)


def test_oracle_math_accuracy():
    L = oracle_pm.lib()
    xs = np.concatenate([np.linspace(-87, 0, 4001), np.linspace(0, 10, 500)]).astype(np.float32)
    got = np.array([L.pm_oracle_expf(float(x)) for x in xs], np.float32)
    ref = np.exp(xs.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 4e-7
    s = ctypes.c_float(); c = ctypes.c_float()
    for a in np.linspace(-np.pi / 2, np.pi / 2, 2001).astype(np.float32):
        L.pm_oracle_sincosf(float(a), ctypes.byref(s), ctypes.byref(c))
        assert abs(s.value - np.sin(float(a))) < 2e-7 and abs(c.value - np.cos(float(a))) < 2e-7


# ---------------------------------------------------------------- oracle behaviour
def test_oracle_converges_to_ground_truth():
    sc = make_patch_match_scene(128, 96, 4, seed=3)
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False,
                          num_iterations=3)
    r = oracle_pm.run(o, sc["problem"])
    valid = r["depth"] > 0
    assert valid.mean() > 0.9
    rel = np.abs(r["depth"] - sc["depth_gt"])[valid] / sc["depth_gt"][valid]
    assert np.median(rel) < 2e-3
    assert (rel < 0.01).mean() > 0.8
    # filtered pixels are zero in depth, normal and mask
    assert np.all(r["normal"][:, ~valid] == 0)
    assert np.all(r["mask"][:, ~valid] == 0)
    # kept pixels have >= filter_min_num_consistent consistent source images
    assert np.all(r["mask"][:, valid].sum(axis=0) >= 2)
    # unit normals facing the camera
    nn = np.linalg.norm(r["normal"][:, valid], axis=0)
    assert np.allclose(nn, 1, atol=1e-4)


def test_oracle_is_deterministic_and_thread_count_independent():
    sc = make_patch_match_scene(48, 40, 3, seed=5)
    o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False,
                          num_iterations=1, window_radius=3)
    a = oracle_pm.run(o, sc["problem"])
    b = oracle_pm.run(o, sc["problem"])
    for k in ("depth", "normal", "sel_prob", "mask"):
        assert np.array_equal(a[k], b[k])


# ---------------------------------------------------------------- product host side vs oracle (bit level)
def test_library_exports_every_declared_symbol():
    lib = load_library()
    hdr = open(os.path.join(ROOT, "include", "b200_patch_match.h")).read()
    names = set(re.findall(r"\b(b200pm_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n


def test_product_host_math_is_bit_identical_to_oracle():
    lib = load_library()
    L = oracle_pm.lib()
    lib.b200pm_test_expf.argtypes = [ctypes.c_float]
    lib.b200pm_test_expf.restype = ctypes.c_float
    lib.b200pm_test_sincosf.argtypes = [ctypes.c_float, _f32p, _f32p]
    lib.b200pm_test_rng_stream.argtypes = [ctypes.c_uint64, ctypes.c_int, _f32p]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-100, 5, 5000), [-87.0, -87.5, 0.0, 88.0, 1e-8, np.nan]]).astype(np.float32)
    for x in xs:
        assert _bits(np.float32(lib.b200pm_test_expf(float(x)))) == _bits(np.float32(L.pm_oracle_expf(float(x))))
    s1 = np.empty(1, np.float32); c1 = np.empty(1, np.float32); s2 = np.empty(1, np.float32); c2 = np.empty(1, np.float32)
    for a in rng.uniform(-1.6, 1.6, 3000).astype(np.float32):
        lib.b200pm_test_sincosf(float(a), s1.ctypes.data_as(_f32p), c1.ctypes.data_as(_f32p))
        L.pm_oracle_sincosf(float(a), s2.ctypes.data_as(_f32p), c2.ctypes.data_as(_f32p))
        assert _bits(s1)[0] == _bits(s2)[0] and _bits(c1)[0] == _bits(c2)[0]
    for seed in (0, 7, 999999):
        a = np.empty(64, np.float32); b = np.empty(64, np.float32)
        lib.b200pm_test_rng_stream(seed, 64, a.ctypes.data_as(_f32p))
        L.pm_oracle_rng_stream(seed, 64, b.ctypes.data_as(_f32p))
        assert np.array_equal(_bits(a), _bits(b))


def test_product_pose_setup_is_bit_identical_to_oracle():
    lib = load_library()
    L = oracle_pm.lib()
    lib.b200pm_test_poses.argtypes = [ctypes.POINTER(_CProblem), ctypes.c_int, _f32p, _f32p, _f32p]
    sc = make_patch_match_scene(64, 48, 5, seed=1)
    o = PatchMatchOptions(depth_min=1, depth_max=2, geom_consistency=False)
    co, cp, keep = marshal(o, sc["problem"])
    for k in range(4):
        pa = np.empty(5 * 43, np.float32); pb = np.empty(5 * 43, np.float32)
        Ka = np.empty(4, np.float32); Kb = np.empty(4, np.float32); ia = np.empty(4, np.float32); ib = np.empty(4, np.float32)
        lib.b200pm_test_poses(ctypes.byref(cp), k, pa.ctypes.data_as(_f32p), Ka.ctypes.data_as(_f32p), ia.ctypes.data_as(_f32p))
        L.pm_oracle_poses(ctypes.byref(cp), k, pb.ctypes.data_as(_f32p), Kb.ctypes.data_as(_f32p), ib.ctypes.data_as(_f32p))
        assert np.array_equal(_bits(pa), _bits(pb))
        assert np.array_equal(_bits(Ka), _bits(Kb)) and np.array_equal(_bits(ia), _bits(ib))


def test_check_rejects_what_the_reference_rejects():
    """PatchMatch::Check (patch_match.cc:67-126) / PatchMatchOptions::Check error behaviour."""
    lib = _bind(load_library())
    sc = make_patch_match_scene(32, 24, 2, seed=0)
    good = PatchMatchOptions(depth_min=1.0, depth_max=2.0, geom_consistency=False)
    co, cp, keep = marshal(good, sc["problem"])
    assert lib.b200pm_check(ctypes.byref(co), ctypes.byref(cp)) == 0
    # geom_consistency without depth / normal maps
    bad = PatchMatchOptions(depth_min=1.0, depth_max=2.0, geom_consistency=True)
    co, cp, keep = marshal(bad, sc["problem"])
    assert lib.b200pm_check(ctypes.byref(co), ctypes.byref(cp)) != 0
    assert b"geom_consistency" in lib.b200pm_last_error()
    # window radius too large, bad step, unresolved depth range
    for kw in (dict(window_radius=21), dict(window_step=3), dict(depth_min=-1.0, depth_max=-1.0), dict(num_samples=0)):
        args = dict(depth_min=1.0, depth_max=2.0, geom_consistency=False)
        args.update(kw)
        co, cp, keep = marshal(PatchMatchOptions(**args), sc["problem"])
        assert lib.b200pm_check(ctypes.byref(co), ctypes.byref(cp)) != 0
    # skewed K
    sc["images"][1].K = sc["images"][1].K.copy()
    sc["images"][1].K[0, 1] = 0.1
    co, cp, keep = marshal(good, sc["problem"])
    assert lib.b200pm_check(ctypes.byref(co), ctypes.byref(cp)) != 0
    pm = PatchMatch(good, sc["problem"])
    with pytest.raises(PatchMatchError):
        pm.Check()


def test_consistency_list_layout():
    mask = np.zeros((3, 2, 4), np.uint8)
    mask[0, 1, 2] = 1; mask[2, 1, 2] = 1; mask[1, 0, 0] = 1
    lst = consistency_list_from_mask(mask, [10, 20, 30])
    assert lst.tolist() == [0, 0, 1, 20, 2, 1, 2, 10, 30]   # [col,row,n,idx...] row-major pixel order
