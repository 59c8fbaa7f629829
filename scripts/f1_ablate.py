"""GPU: ablation timings of the fused conv1a+conv1b kernel (OSB_F1_ABLATE; results are wrong by construction, only the
times and cycle counters mean something).  One process per setting (the switch is read once)."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for ab in (sys.argv[1:] or ["0", "1", "2", "3", "4", "8", "16", "20", "28"]):
    env = dict(os.environ, OSB_F1_ABLATE=ab, F1_MODES="1", OSB_F1_DEBUG="1")
    out = subprocess.run([sys.executable, os.path.join(here, "f1_probe.py")], env=env, capture_output=True, text=True).stdout
    print("ablate", ab, "|", " ".join(l for l in out.splitlines() if "conv1" in l or "prod_" in l)[:900], flush=True)
