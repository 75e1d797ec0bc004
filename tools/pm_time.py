"""Quick timing of the C2 workload (1 ref + 8 src, 1920x1080, window 11, 5 iterations) per WPC."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from colmap_b200.patch_match import PatchMatch, PatchMatchOptions
from colmap_b200.synthetic import make_patch_match_scene
W, H, N = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 8)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
t = time.time(); sc = make_patch_match_scene(W, H, N, seed=0); print("scene s", time.time() - t, flush=True)
o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False, num_iterations=iters)
for wpc in [int(x) for x in os.environ.get("WPCS", "0,1,2,4").split(",")]:   # 0 = automatic per-sweep schedule
    if wpc: os.environ["B200PM_WPC"] = str(wpc)
    else: os.environ.pop("B200PM_WPC", None)
    pm = PatchMatch(o, sc["problem"])
    t = time.time(); pm.Run(); e2e = time.time() - t
    ms = [pm.last_run_ms()]
    for _ in range(2):
        pm.RunOnly(); ms.append(pm.last_run_ms())
    d = pm.GetDepthMap(); valid = d > 0
    rel = np.abs(d - sc["depth_gt"])[valid] / sc["depth_gt"][valid]
    passes = [pm._lib.b200pm_last_pass_ms(pm._h, k) for k in range(3)] if hasattr(pm, "_lib") else []
    print(f"WPC={wpc}: passes R/P/S ms {passes} run ms {ms} sweep_ms {pm.last_sweep_ms():.1f} -> {W*H/1e6/(min(ms)/1e3):.2f} Mpx/s; first Run() incl create {e2e:.2f}s; valid {valid.mean():.3f} med rel err {np.median(rel):.2e}", flush=True)
    pm.close()
