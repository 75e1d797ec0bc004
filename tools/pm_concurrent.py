"""Aggregate PatchMatch throughput with K problems in flight on ONE GPU (K host threads, one handle each) - the
reference's `gpu_index = "0,0"` mode (patch_match.cc:375-384)."""
import os
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from colmap_b200.patch_match import PatchMatch, PatchMatchOptions
from colmap_b200.synthetic import make_patch_match_scene

w, h, n = (int(v) for v in sys.argv[1:4])
ks = [int(v) for v in sys.argv[4:]] or [1, 2, 3]
scs = [make_patch_match_scene(w, h, n, seed=i) for i in range(max(ks))]
for K in ks:
    pms = []
    for i in range(K):
        sc = scs[i]
        o = PatchMatchOptions(depth_min=sc["depth_min"], depth_max=sc["depth_max"], geom_consistency=False, gpu_index="0")
        pm = PatchMatch(o, sc["problem"]); pm.Run(); pms.append(pm)
    def work(pm):
        for _ in range(2):
            pm.RunOnly()
    t0 = time.time()
    th = [threading.Thread(target=work, args=(pm,)) for pm in pms]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.time() - t0
    print(f"{K} problems in flight: {2 * K * w * h / 1e6 / dt:.3f} Mpx/s aggregate ({dt * 1e3 / 2:.1f} ms per round of {K})", flush=True)
    [pm.close() for pm in pms]
