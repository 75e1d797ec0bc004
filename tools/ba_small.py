"""Small-problem truth: B1 (10 pinhole cameras, 1k points, 5k observations) and a local-BA-sized
problem (incremental_mapper.cc:991-1116: a handful of images inside a larger model), END TO END through b200ba_solve with
host arrays (set-up included), against the CPU oracle port at ONE thread (Ceres runs problems below 50k residuals
single-threaded, bundle_adjustment_ceres.h:64)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import json
import numpy as np
from colmap_b200.bundle_adjustment import PINHOLE, SIMPLE_RADIAL, BundleAdjustmentOptions, solve_flat
from colmap_b200.synthetic import synthesize_ba_problem


def gauge(f):
    f.pose_constant = f.pose_constant.copy(); f.pose_fixed_dim = f.pose_fixed_dim.copy()
    f.pose_constant[0] = 1
    f.pose_fixed_dim[1] = int(np.argmax(np.abs(f.poses[1, 4:] - f.poses[0, 4:])))
    return f


def fresh(n):
    f = n.copy(); f.pose_constant, f.pose_fixed_dim = n.pose_constant, n.pose_fixed_dim
    return f


out = {}
cases = {"B1 (10 shared-PINHOLE images, 1k points, 5k obs)": synthesize_ba_problem(10, 1000, 5, models=(PINHOLE,), shared_camera=True, seed=42)[1],
         "local BA (8 SIMPLE_RADIAL images, 2k points, 12k obs)": synthesize_ba_problem(8, 2000, 6, models=(SIMPLE_RADIAL,), seed=3)[1],
         "50 images (own SIMPLE_RADIAL), 10k points, 80k obs": synthesize_ba_problem(50, 10000, 8, models=(SIMPLE_RADIAL,), seed=4)[1]}
for name, noisy in cases.items():
    gauge(noisy)
    o = BundleAdjustmentOptions(gpu_index=0)
    for env in (() if "--cpu" in sys.argv else ({}, {"B200BA_NO_DENSE": "1"})):
        for k, v in env.items():
            os.environ[k] = v
        solve_flat(o, fresh(noisy))
        t, lm, dev = 0.0, 0, 0.0
        for _ in range(5):
            f = fresh(noisy)
            t0 = time.time(); s = solve_flat(o, f); t += time.time() - t0
            lm += s.num_successful_steps + s.num_unsuccessful_steps; dev += s.solve_ms
        for k in env:
            os.environ.pop(k, None)
        out.setdefault(name, {})["pcg_to_1e-12" if env else "dense_cholesky"] = dict(
            lm_it_per_s_e2e=lm / t, ms_per_solve_e2e=t / 5 * 1e3, ms_per_solve_device=dev / 5, setup_ms=s.setup_ms,
            lm_iterations=lm / 5, final_cost=s.final_cost, solver=s.linear_solver_type_used)
    if "--cpu" in sys.argv:      # run as: OMP_NUM_THREADS=1 python tools/ba_small.py --cpu
        import oracle_ba
        f = fresh(noisy)
        t0 = time.time(); s = oracle_ba.solve(BundleAdjustmentOptions(), f); dt = time.time() - t0
        out.setdefault(name, {})["cpu_oracle_%s_threads" % os.environ.get("OMP_NUM_THREADS", "all")] = dict(lm_it_per_s=(s.num_successful_steps + s.num_unsuccessful_steps) / dt, ms_per_solve=dt * 1e3,
                                                final_cost=s.final_cost)
print(json.dumps(out, indent=1))
