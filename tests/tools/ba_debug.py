import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ba
from colmap_b200.bundle_adjustment import BundleAdjustmentOptions, ITERATIVE_SCHUR, SIMPLE_RADIAL, solve_flat
from colmap_b200.synthetic import synthesize_ba_problem
gt, noisy = synthesize_ba_problem(60, 12000, 6, models=(SIMPLE_RADIAL,), seed=21)
noisy.pose_constant = noisy.pose_constant.copy(); noisy.pose_fixed_dim = noisy.pose_fixed_dim.copy()
noisy.pose_constant[0] = 1; noisy.pose_fixed_dim[1] = int(np.argmax(np.abs(noisy.poses[1, 4:] - noisy.poses[0, 4:])))
o = BundleAdjustmentOptions(linear_solver_type=ITERATIVE_SCHUR)
a, b = noisy.copy(), noisy.copy()
for f in (a, b): f.pose_constant, f.pose_fixed_dim = noisy.pose_constant, noisy.pose_fixed_dim
if "gpu" in sys.argv: print(solve_flat(o, a))
if "cpu" in sys.argv: print(oracle_ba.solve(o, b))
