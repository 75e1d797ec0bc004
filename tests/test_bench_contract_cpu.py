"""bench.py's reference arm runs without a GPU (oracle port legs) and prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    # force the CPU legs even where a GPU exists; OMP_NUM_THREADS=1 is what torchrun exports to every rank - the CPU legs
    # must raise their OpenMP team to all host cores themselves (with one thread the B3 leg alone would take many minutes)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "patchmatch_mpixels_per_s" and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    ba = d["ba"]
    assert ba["impl"] == "reference" and ba["metric"] == "ba_lm_iterations_per_s" and ba["value"] > 0
    assert ba["cpu_baseline"]["kind"] == "port" and ba["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
