"""GPU box: ablation timings of the fused conv1a+conv1b kernel.  The switch is COMPILE-TIME (conv1_fused.cu, F1_ABLATE bits:
1 producers do not store, 2 no patch loads, 4 epilogue only drains TMEM, 8 no lo*hi MMA, 16 producers do not compute), so
every setting rebuilds conv1_fused.o and the library in place and restores the normal build at the end.  Results are wrong
by construction, only the times and cycle counters mean something.

    gpurun -- 'python scripts/f1_ablate.py 0 1 2 4 8 16 28 | tee gpurun_out/f1_ablate.txt'

(profiles/r02_f1_ablate.txt was taken with a run-time form of the same switch, before the elect.sync fix.)"""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
csrc = os.path.join(os.path.dirname(here), "omni-swarm_b200", "csrc")


def build(bits):
    os.utime(os.path.join(csrc, "conv1_fused.cu"))
    subprocess.run(["make", "-C", csrc, "-j8"] + ([f"EXTRA=-DF1_ABLATE={bits}"] if bits else []), check=True, capture_output=True)


try:
    for ab in (sys.argv[1:] or ["0", "1", "2", "3", "4", "8", "16", "20", "28"]):
        build(int(ab))
        env = dict(os.environ, F1_MODES="1", OSB_F1_DEBUG="1")
        out = subprocess.run([sys.executable, os.path.join(here, "f1_probe.py")], env=env, capture_output=True, text=True).stdout
        print("ablate", ab, "|", " ".join(l for l in out.splitlines() if "conv1" in l or "prod_" in l)[:900], flush=True)
finally:
    build(0)
