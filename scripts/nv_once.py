import sys
import numpy as np
sys.path.insert(0, ".")
from omniswarm_b200 import host, synth
nv = host.NetVLAD(synth.flatten_nv_weights(synth.netvlad_weights(0)), 640, 480, max_batch=4)
imgs = np.stack([synth.image(s) for s in range(4)])
for _ in range(3):
    nv.inference_batch(imgs)
