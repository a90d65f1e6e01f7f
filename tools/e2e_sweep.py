"""Sweeps gs_grid1d_step_host's chunk size (option host_chunk_mb) on the C2 shape: e2e ms per step."""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in (sys.argv[1:] or ["16", "32", "64", "128", "256", "512"]):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--configs", "none",
                          "--e2e-steps", "5", "--cpu-steps", "20", "--opt", f"host_chunk_mb={c}"], capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print("host_chunk_mb", c, "e2e ms/step", round(d["e2e"]["ms_per_step"], 2), "G upd/s", round(d["e2e"]["value"] / 1e9, 3), flush=True)
