"""Generates the committed fixtures in tests/golden/ from the ORACLE (oracle/*.py), which is itself pinned
against torch.grid_sample / cv2.BFMatcher / finite differences by tests/test_oracle_pins.py.

The reference ships no golden vectors and cannot be built or imported here, so these
fixtures pin the oracle's behaviour at commit time: the -m "not gpu" suite re-derives them from the oracle,
the -m gpu suite compares the CUDA path against them.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omniswarm_b200 import synth                     # noqa: E402
from oracle import frontend_ref as fr                # noqa: E402
from oracle import solver_ref as sr                  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def heatmap(seed, H, W, density=0.03, plateau=False):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.0, 0.014, (H, W)).astype(np.float32)
    m = rng.uniform(size=(H, W)) < density
    p[m] = rng.uniform(0.016, 0.9, m.sum()).astype(np.float32)
    if plateau:   # equal confidences never suppress each other
        p[H // 2:H // 2 + 6, W // 2:W // 2 + 12] = np.float32(0.5)
        p[0:3, W - 5:W] = np.float32(0.25)      # touches the right edge: flat-address wrap
    return p


def main():
    # ---- 1. keypoints + descriptors from a given heatmap / descriptor map (post-processing stage) ----
    H, W = 64, 96
    comp, mean = synth.pca_matrices(0)
    cases = {}
    for name, seed, dens, plat in [("a", 1, 0.03, False), ("b", 2, 0.15, True), ("c", 3, 0.002, False)]:
        semi = heatmap(seed, H, W, dens, plat)
        rng = np.random.default_rng(100 + seed)
        desc = rng.standard_normal((256, H // 8, W // 8)).astype(np.float32)
        desc /= np.linalg.norm(desc, axis=0, keepdims=True)
        k, c = fr.get_keypoints(semi, 0.015, 50)
        d = fr.compute_descriptors(desc, k, W, H, comp, mean)
        cases[name] = dict(semi=semi, desc=desc, kpts=k, conf=c, out=d)
    np.savez_compressed(os.path.join(OUT, "postproc.npz"),
                        **{f"{n}_{k}": v for n, c in cases.items() for k, v in c.items()})

    # ---- 2. SuperPoint network on a 64x96 image ----
    w = synth.superpoint_weights(0)
    img = synth.image(5, 64, 96)
    semi, desc = fr.superpoint_net(img, w)
    np.savez_compressed(os.path.join(OUT, "superpoint_net.npz"), img=img, semi=semi, desc=desc.astype(np.float16))

    # ---- 3. NetVLAD stand-in on a 64x96 image ----
    nvw = synth.netvlad_weights(0)
    v = fr.netvlad_net(img, nvw)
    np.savez_compressed(os.path.join(OUT, "netvlad.npz"), img=img, out=v)

    # ---- 4. matcher ----
    a = synth.local_descriptors(57, 1)
    b = synth.local_descriptors(43, 2, base=a)
    qi, ti, dist = fr.bf_crosscheck(a, b)
    np.savez_compressed(os.path.join(OUT, "matcher.npz"), q=a, t=b, qi=qi, ti=ti, dist=dist)

    # ---- 5. database search + acceptance rule ----
    db = synth.descriptor_db(300, 4096, 1)
    q = synth.noisy_queries(db, np.array([3, 150, 299, 7]))
    idx = fr.IndexFlatIP(4096); idx.add(db)
    D, I = idx.search(q, 10)
    np.savez_compressed(os.path.join(OUT, "db_search.npz"), rows=np.array([3, 150, 299, 7]), D=D, I=I)

    # ---- 6. pose graph ----
    g = synth.pose_graph(3, 12, n_uwb=20, n_loop=15, n_det=8, n_bearing=9, seed=3)
    res = sr.solve(g)
    r0 = [sr.factor_residual_jacobian(int(g["ftype"][f]), g["init"][g["ia"][f]], g["init"][g["ib"][f]], g["payload"][f])
          for f in range(len(g["ftype"]))]
    R = np.zeros((len(r0), 4)); JA = np.zeros((len(r0), 4, 4)); JB = np.zeros((len(r0), 4, 4))
    for f, (r, ja, jb) in enumerate(r0):
        R[f, :len(r)] = r; JA[f, :len(r)] = ja; JB[f, :len(r)] = jb
    np.savez_compressed(os.path.join(OUT, "graph_small.npz"), poses=res["poses"], final_cost=res["final_cost"],
                        initial_cost=res["initial_cost"], r=R, Ja=JA, Jb=JB)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
