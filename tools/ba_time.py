"""B3 timing of b200ba_solve (device LM iterations/s, per-solve wall) - regression check for the tuned narrow path."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from colmap_b200.bundle_adjustment import ITERATIVE_SCHUR, SIMPLE_RADIAL, BundleAdjustmentOptions, solve_flat
from colmap_b200.synthetic import synthesize_ba_problem
gt, noisy = synthesize_ba_problem(500, 300000, 7, models=(SIMPLE_RADIAL,), seed=42, num_obs=2000000)
noisy.pose_constant = noisy.pose_constant.copy(); noisy.pose_fixed_dim = noisy.pose_fixed_dim.copy()
noisy.pose_constant[0] = 1
noisy.pose_fixed_dim[1] = int(np.argmax(np.abs(noisy.poses[1, 4:] - noisy.poses[0, 4:])))
o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR, gpu_index=0)
def fresh():
    f = noisy.copy(); f.pose_constant, f.pose_fixed_dim = noisy.pose_constant, noisy.pose_fixed_dim
    return f
solve_flat(o, fresh())
for _ in range(3):
    f = fresh(); t = time.time(); s = solve_flat(o, f); dt = time.time() - t
    lm = s.num_successful_steps + s.num_unsuccessful_steps
    print(f"B3: {lm} LM its, device {s.solve_ms:.1f} ms = {lm / s.solve_ms * 1e3:.1f} LM it/s, wall {dt * 1e3:.1f} ms = {lm / dt:.1f} LM it/s, setup {s.setup_ms:.1f} ms, cost {s.final_cost:.6f}, spmv {s.spmv_ms_total / max(s.spmv_launches, 1) * 1e3:.1f} us", flush=True)
