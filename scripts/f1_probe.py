"""GPU: time the fused conv1a+conv1b kernel and print its cycle counters (CTA 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from omniswarm_b200 import synth, host
comp, mean = synth.pca_matrices(0)
spw = synth.flatten_sp_weights(synth.superpoint_weights(0))
imgs = np.stack([synth.image(s) for s in range(8)])
os.environ.setdefault("OSB_F1_DEBUG", "1")          # cycle counters on (layer 1 = the fused first layers)
for fuse in (os.environ.get("F1_MODES", "1,0").split(",")):
    os.environ["OSB_SP_FUSE1"] = fuse
    sp = host.SuperPoint(spw, comp, mean, 640, 480, 0.015, 200, max_batch=8)
    sp.layer_ms(imgs)
    runs = [sp.layer_ms(imgs) for _ in range(5)]
    ms = {k: float(np.median([r[k] for r in runs])) for k in runs[0]}
    print("fuse", fuse, {k: round(v, 4) for k, v in ms.items()}, "sum", round(sum(ms.values()), 4))
    if fuse == "1":
        c = sp.read("fused1_cycles")
        n = max(c[9], 1)
        names = ["prod_wait_window", "prod_compute", "prod_wait_shared", "prod_total", "mma_wait_tmem", "mma_wait_tile",
                 "mma_issue", "epi_wait", "epi_work", "tiles"]
        print({k: round(float(v / n), 1) for k, v in zip(names, c[:9])}, "tiles", int(c[9]))
        print("producer warps busy+wait per tile:", [round(float(v / n), 1) for v in c[10:16]])
    sp.close()
