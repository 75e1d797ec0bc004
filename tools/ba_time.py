"""Timing of the BA configs (B1 / B3) on the GPU (the oracle comparison lives in tests/)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colmap_b200.bundle_adjustment import BundleAdjustmentOptions, ITERATIVE_SCHUR, SIMPLE_RADIAL, PINHOLE, solve_flat
from colmap_b200.synthetic import synthesize_ba_problem

cfg = sys.argv[1] if len(sys.argv) > 1 else "B3"
t = time.time()
if cfg == "B1":
    gt, noisy = synthesize_ba_problem(10, 1000, 5, models=(PINHOLE,), shared_camera=True, seed=42); o = BundleAdjustmentOptions()
elif cfg == "B3":
    gt, noisy = synthesize_ba_problem(500, 300000, 7, models=(SIMPLE_RADIAL,), seed=42, num_obs=2000000); o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR)
else:
    n = int(cfg); gt, noisy = synthesize_ba_problem(n, n * 600, 6, models=(SIMPLE_RADIAL,), seed=42); o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR)
noisy.pose_constant = noisy.pose_constant.copy(); noisy.pose_fixed_dim = noisy.pose_fixed_dim.copy()
noisy.pose_constant[0] = 1; noisy.pose_fixed_dim[1] = int(np.argmax(np.abs(noisy.poses[1, 4:] - noisy.poses[0, 4:])))
print("gen s", time.time() - t, "obs", len(noisy.obs_pose), flush=True)
for rep in range(3):
    f = noisy.copy(); f.pose_constant, f.pose_fixed_dim = noisy.pose_constant, noisy.pose_fixed_dim
    t = time.time(); s = solve_flat(o, f); wall = time.time() - t
    steps = s.num_successful_steps + s.num_unsuccessful_steps
    print(f"rep{rep}: term {s.termination_type} cost {s.initial_cost:.6g}->{s.final_cost:.6g} LM steps {steps} pcg {s.num_linear_solver_iterations} solve_ms {s.solve_ms:.1f} setup_ms {s.setup_ms:.1f} wall {wall*1e3:.1f} -> {steps/(s.solve_ms/1e3):.1f} LM it/s; spmv avg {s.spmv_ms_total/max(s.spmv_launches,1)*1e3:.1f} us over {s.spmv_launches} samples; launches {s.kernel_launches}", flush=True)
