"""First-contact debugging: compare the CUDA path with the oracle after 0,1,2,... sweeps."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_pm
from colmap_b200.patch_match import PatchMatch, PatchMatchOptions, _f32p
from colmap_b200.synthetic import make_patch_match_scene

W, H, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sc = make_patch_match_scene(W, H, N, seed=0)
o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False, num_iterations=iters)
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
for sweeps in list(range(0, 4 * iters + 1)):
    os.environ["B200PM_MAX_SWEEPS"] = str(sweeps)
    pm = PatchMatch(o, sc["problem"]); pm.Run()
    d, nrm = pm.GetDepthMap(), pm.GetNormalMap()
    cost = np.empty((N, H, W), np.float32)
    pm._lib.b200pm_debug_get_cost.argtypes = [ctypes.c_void_p, _f32p]
    pm._lib.b200pm_debug_get_cost(pm._h, cost.ctypes.data_as(_f32p))
    ref = oracle_pm.run(o, sc["problem"], stop_after_sweeps=sweeps if sweeps < 4 * iters else -1)
    eq_d = (bits(d) == bits(ref["depth"])).mean()
    eq_n = (bits(nrm) == bits(ref["normal"])).mean()
    eq_c = (bits(cost) == bits(ref["cost"])).mean()
    print(f"sweeps={sweeps}: depth eq {eq_d:.6f} normal eq {eq_n:.6f} cost eq {eq_c:.6f} maxabs cost diff {np.nanmax(np.abs(cost-ref['cost'])):.3g} ms={pm.last_run_ms():.2f}", flush=True)
    if eq_d < 1 or eq_c < 1:
        bad = np.argwhere(bits(cost) != bits(ref["cost"]))[:5]
        print(" first cost mismatches (img,row,col):", bad.tolist())
        bad = np.argwhere(bits(d) != bits(ref["depth"]))[:5]
        print(" first depth mismatches (row,col):", bad.tolist(), [ (float(d[r,c]), float(ref['depth'][r,c])) for r,c in bad])
    pm.close()
