"""Statistical cross-check of this repo's PatchMatch against the real reference PatchMatchCuda (oracle/_ref) on the
same GPU and inputs, plus the reference's timing.  Both are stochastic and numerically different (fast-math, texture
filtering, PRNG use is identical in law but not in float detail), so agreement is judged per pixel against the
analytic ground truth and against each other."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_pm
from colmap_b200.patch_match import PatchMatch, PatchMatchOptions
from colmap_b200.synthetic import make_patch_match_scene

W, H, N = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 8)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
sc = make_patch_match_scene(W, H, N, seed=0)
o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False, num_iterations=iters, gpu_index="0")
gt = sc["depth_gt"]
out = {"config": f"{W}x{H}, {N} src, window 11, {iters} iterations, photometric + filter"}
t = time.time(); r = ref_pm.run(o, sc["problem"]); first = time.time() - t
ref_ms = [ref_pm.run(o, sc["problem"])["ms"] for _ in range(reps)]
out["reference"] = {"first_call_s_incl_jit": first, "ms": ref_ms, "mpix_per_s": W * H / 1e6 / (min(ref_ms) / 1e3)}
pm = PatchMatch(o, sc["problem"]); pm.Run(); ours = dict(depth=pm.GetDepthMap(), normal=pm.GetNormalMap())
ms = []
for _ in range(reps):
    t = time.time(); p2 = PatchMatch(o, sc["problem"]); p2.Run(); p2.GetDepthMap(); p2.GetNormalMap(); p2.close(); ms.append((time.time() - t) * 1e3)
out["ours"] = {"ms_e2e": ms, "run_ms": pm.last_run_ms(), "mpix_per_s_e2e": W * H / 1e6 / (min(ms) / 1e3)}
out["speedup_e2e"] = min(ref_ms) / min(ms)
def stats(d):
    v = d > 0
    rel = np.abs(d - gt)[v] / gt[v]
    return {"valid_frac": float(v.mean()), "median_rel_err": float(np.median(rel)), "frac_rel_err_lt_1e-3": float((rel < 1e-3).mean()),
            "frac_rel_err_lt_1e-2": float((rel < 1e-2).mean())}
out["reference"]["quality"] = stats(r["depth"]); out["ours"]["quality"] = stats(ours["depth"])
both = (r["depth"] > 0) & (ours["depth"] > 0)
dd = np.abs(r["depth"] - ours["depth"])[both]
nn = (r["normal"] * ours["normal"]).sum(0)[both]
out["agreement"] = {"both_valid_frac": float(both.mean()), "median_abs_depth_diff": float(np.median(dd)),
                    "frac_abs_diff_lt_1e-4": float((dd < 1e-4).mean()), "frac_abs_diff_lt_1e-3": float((dd < 1e-3).mean()),
                    "frac_abs_diff_lt_1e-2": float((dd < 1e-2).mean()), "median_normal_dot": float(np.median(nn)),
                    "depth_scale_m": float(gt.mean())}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"pm_ref_compare_{W}x{H}.json"), "w"), indent=1)
