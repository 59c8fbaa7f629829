"""Stand-alone timing of the NetVLAD stand-in (4 images) and SuperPoint (8 images) through the host mirror --
used to attribute the cost of running them side by side in the keyframe front-end."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from omniswarm_b200 import host, synth

W, H = 640, 480
nv = host.NetVLAD(synth.flatten_nv_weights(synth.netvlad_weights(0)), W, H, max_batch=4)
imgs = np.stack([synth.image(s) for s in range(4)])
for _ in range(5):
    nv.inference_batch(imgs)
t0 = time.perf_counter()
for _ in range(50):
    nv.inference_batch(imgs)
print("NetVLAD x4 (host buffers): %.3f ms per batch" % ((time.perf_counter() - t0) / 50 * 1e3))
comp, mean = synth.pca_matrices(0)
sp = host.SuperPoint(synth.flatten_sp_weights(synth.superpoint_weights(0)), comp, mean, W, H, 0.015, 200, max_batch=8)
imgs8 = np.concatenate([imgs, imgs])
for _ in range(5):
    sp.inference_batch(imgs8)
t0 = time.perf_counter()
for _ in range(50):
    sp.inference_batch(imgs8)
print("SuperPoint x8 (host buffers): %.3f ms per batch" % ((time.perf_counter() - t0) / 50 * 1e3))
