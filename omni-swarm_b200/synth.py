"""Seeded synthetic inputs shared by the product, the tests, the oracle and bench.py.

Nothing here computes any part of the hot path: it only *generates data* (network weights,
PCA matrices, images, descriptor databases, pose graphs).  The reference ships no weights, bags
or fixtures: weights are a Dropbox download
(/root/reference/README.md:27), so every run of the path uses seeded stand-ins fed identically
to the oracle and to the CUDA kernels.

All generators use numpy's PCG64 (`np.random.default_rng(seed)`), which is bit-reproducible across
machines for a given numpy version (the GPU box runs this same image).
"""
from __future__ import annotations

import math
import numpy as np

# ----------------------------------------------------------------------------------------------
# SuperPoint (layer table follows /root/reference/swarm_loop/superpoint.ipynb:135-160)
# ----------------------------------------------------------------------------------------------
# (name, Cin, Cout, ksize)
SP_LAYERS = [
    ("conv1a", 1, 64, 3), ("conv1b", 64, 64, 3),
    ("conv2a", 64, 64, 3), ("conv2b", 64, 64, 3),
    ("conv3a", 64, 128, 3), ("conv3b", 128, 128, 3),
    ("conv4a", 128, 128, 3), ("conv4b", 128, 128, 3),
    ("convPa", 128, 256, 3), ("convPb", 256, 65, 1),
    ("convDa", 128, 256, 3), ("convDb", 256, 256, 1),
]


def sp_num_weights() -> int:
    return sum(co * ci * k * k + co for _, ci, co, k in SP_LAYERS)


def superpoint_weights(seed: int = 0, pb_gain: float = 8.0, dustbin_bias: float = 11.0) -> dict:
    """He-initialised SuperPoint weights, OIHW float32, plus biases.

    `pb_gain` / `dustbin_bias` shape the detector head so that the 65-way softmax is not
    degenerate (uniform 1/65 would put every pixel above the reference's 0.015 threshold):
    with these values a textured 640x480 image yields ~3500 candidates above
    thres=0.015 and ~320 above 0.2, i.e. the regime of SURVEY.md section 8d C2.
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, ci, co, k in SP_LAYERS:
        std = math.sqrt(2.0 / (ci * k * k))
        W = rng.standard_normal((co, ci, k, k)).astype(np.float32) * np.float32(std)
        b = (rng.standard_normal(co) * 0.05).astype(np.float32)
        if name == "convPb":
            W = W * np.float32(pb_gain)
            b[64] += np.float32(dustbin_bias)
        w[name + ".weight"] = np.ascontiguousarray(W)
        w[name + ".bias"] = b
    return w


def flatten_sp_weights(w: dict) -> np.ndarray:
    """The C-ABI weight blob: for each layer in SP_LAYERS order, weight (OIHW) then bias."""
    parts = []
    for name, *_ in SP_LAYERS:
        parts.append(w[name + ".weight"].reshape(-1))
        parts.append(w[name + ".bias"].reshape(-1))
    out = np.concatenate(parts).astype(np.float32)
    assert out.size == sp_num_weights()
    return out


def pca_matrices(seed: int = 0):
    """Stand-in for components_.csv (64x256) / mean_.csv (256) written by
    /root/reference/swarm_loop/pca.ipynb cells 2-3: orthonormal rows + small mean."""
    rng = np.random.default_rng(seed + 1000)
    q, _ = np.linalg.qr(rng.standard_normal((256, 64)))
    comp = np.ascontiguousarray(q.T.astype(np.float32))          # [64,256]
    mean = (rng.standard_normal(256) * 0.02).astype(np.float32)  # [256]
    return comp, mean


# ----------------------------------------------------------------------------------------------
# NetVLAD stand-in ("MobileNetVLAD-lite"; the real hfnet architecture is absent from the
# reference).  I/O contract: mobilenetvlad_tensorrt.h:9-15 (HxW f32 0..255
# in, 4096 f32 out).
# ----------------------------------------------------------------------------------------------
# backbone: conv0 3x3 s2 (1->32) + ReLU6, then depthwise-separable blocks (dw3x3 stride s, pw 1x1)
NV_BLOCKS = [  # (Cin, Cout, stride)
    (32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2),
    (512, 512, 1),
]
NV_K = 32     # clusters
NV_D = 128    # descriptor dim  -> K*D = 4096 = DEEP_DESC_SIZE (loop_defines.h:30)
NV_INPUT_SCALE = 1.0 / 255.0  # folded into conv0 (engine input is unscaled 0..255)


def netvlad_layer_table():
    """[(name, shape)] in blob order."""
    t = [("conv0.weight", (32, 1, 3, 3)), ("conv0.bias", (32,))]
    for i, (ci, co, s) in enumerate(NV_BLOCKS):
        t += [(f"b{i}.dw.weight", (ci, 1, 3, 3)), (f"b{i}.dw.bias", (ci,)),
              (f"b{i}.pw.weight", (co, ci, 1, 1)), (f"b{i}.pw.bias", (co,))]
    t += [("proj.weight", (NV_D, 512, 1, 1)), ("proj.bias", (NV_D,)),
          ("assign.weight", (NV_K, NV_D, 1, 1)), ("assign.bias", (NV_K,)),
          ("centroids", (NV_K, NV_D))]
    return t


def nv_num_weights() -> int:
    return sum(int(np.prod(s)) for _, s in netvlad_layer_table())


def netvlad_weights(seed: int = 0) -> dict:
    rng = np.random.default_rng(seed + 2000)
    w = {}
    for name, shape in netvlad_layer_table():
        if name.endswith(".bias"):
            w[name] = (rng.standard_normal(shape) * 0.05).astype(np.float32)
        elif name == "centroids":
            # small centroids: residuals x - c stay image-specific (real NetVLAD centroids are data cluster centres)
            w[name] = (rng.standard_normal(shape) * 0.01).astype(np.float32)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 8.0 if name.startswith("assign") else 1.0
            w[name] = (rng.standard_normal(shape) * math.sqrt(2.0 / fan_in) * gain).astype(np.float32)
    return w


def flatten_nv_weights(w: dict) -> np.ndarray:
    out = np.concatenate([w[n].reshape(-1) for n, _ in netvlad_layer_table()]).astype(np.float32)
    assert out.size == nv_num_weights()
    return out


# ----------------------------------------------------------------------------------------------
# Images, descriptor databases
# ----------------------------------------------------------------------------------------------
def image(seed: int, H: int = 480, W: int = 640, zero_bottom_quarter: bool = False) -> np.ndarray:
    """Textured uint8 image: smoothed noise at two scales + sharp blobs + pixel noise
    (SURVEY.md section 8d C2).  `zero_bottom_quarter` reproduces loop_cam.cpp:536-539."""
    from scipy import ndimage
    rng = np.random.default_rng(seed + 3000)
    a = ndimage.gaussian_filter(rng.standard_normal((H, W)), 6.0)
    b = ndimage.gaussian_filter(rng.standard_normal((H, W)), 1.5)
    img = a / a.std() * 40.0 + b / b.std() * 25.0 + 128.0
    nb = 120
    ys = rng.integers(0, H, nb)
    xs = rng.integers(0, W, nb)
    rad = rng.integers(2, 9, nb)
    val = rng.uniform(-90, 90, nb)
    yy, xx = np.mgrid[0:H, 0:W]
    for y, x, r, v in zip(ys, xs, rad, val):
        y0, y1, x0, x1 = max(0, y - r), min(H, y + r + 1), max(0, x - r), min(W, x + r + 1)
        m = (yy[y0:y1, x0:x1] - y) ** 2 + (xx[y0:y1, x0:x1] - x) ** 2 <= r * r
        img[y0:y1, x0:x1][m] += v
    img += rng.uniform(-6, 6, (H, W))
    img = np.clip(img, 0, 255).astype(np.uint8)
    if zero_bottom_quarter:
        img[H * 3 // 4:, :] = 0
    return img


def descriptor_db(n: int, dim: int = 4096, seed: int = 1) -> np.ndarray:
    """n unit-norm Gaussian rows, float32."""
    rng = np.random.default_rng(seed + 4000)
    out = np.empty((n, dim), np.float32)
    step = 4096
    for s in range(0, n, step):
        x = rng.standard_normal((min(step, n - s), dim)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[s:s + x.shape[0]] = x
    return out


def noisy_queries(db: np.ndarray, rows: np.ndarray, sigma: float = 0.5, seed: int = 2) -> np.ndarray:
    """Queries = chosen DB rows + Gaussian noise of total norm ~sigma, renormalised, so that the
    inner product with the source row is ~1/sqrt(1+sigma^2) (0.89 for 0.5): hits exist."""
    rng = np.random.default_rng(seed + 5000)
    dim = db.shape[1]
    noise = rng.standard_normal((len(rows), dim)).astype(np.float32) * np.float32(sigma / math.sqrt(dim))
    q = (db[rows] + noise).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(q)


def local_descriptors(n: int, seed: int, dim: int = 64, base: np.ndarray | None = None,
                      sigma: float = 0.15) -> np.ndarray:
    """n x dim float32 local descriptors; with `base`, a noisy permuted copy (so matches exist)."""
    rng = np.random.default_rng(seed + 6000)
    if base is None:
        return rng.standard_normal((n, dim)).astype(np.float32)
    idx = rng.permutation(base.shape[0])[:n]
    d = base[idx] + sigma * rng.standard_normal((len(idx), dim)).astype(np.float32)
    return np.ascontiguousarray(d.astype(np.float32))


# ----------------------------------------------------------------------------------------------
# Pose graphs (recipe: /root/reference/swarm_localization/test/swarm_local_sim.cpp)
# ----------------------------------------------------------------------------------------------
FACTOR_DISTANCE = 0    # DistanceMeasurementFactor        factors.hpp:203-224
FACTOR_RELPOSE = 1     # RelativePoseFactor4d             factors.hpp:226-271
FACTOR_DETECTION = 2   # DroneDetection4dFactor           factors.hpp:273-367
PAYLOAD_LEN = 24       # doubles per factor (layout: include/omniswarm_b200.h)


def _wrap(a):
    return a - 2.0 * math.pi * np.floor((a + math.pi) / (2.0 * math.pi))


def _delta_pose(pa, pb):
    """a^-1 . b  for 4-DoF poses [x,y,z,yaw]  (factors.hpp:139-149)."""
    c, s = math.cos(pa[3]), math.sin(pa[3])
    d = pb[:3] - pa[:3]
    return np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2], _wrap(pb[3] - pa[3])])


def pose_graph(n_drones: int = 5, n_frames: int = 100, n_uwb: int | None = None,
               n_loop: int | None = None, n_det: int | None = None, n_bearing: int = 0,
               seed: int = 0, init_sigma_pos: float = 0.3, init_sigma_yaw: float = 0.05,
               outlier_frac: float = 0.0) -> dict:
    """Synthetic swarm pose graph in the flat layout `osb_graph_solve` consumes.

    Ground truth follows swarm_local_sim.cpp:589-604 (`ParalCirc`: radius 5 m, z amplitude 2 m,
    period 50 s, drones offset on a grid), plus a slow yaw so that yaw is exercised.
    Node index = frame * n_drones + drone.  Factor mix: ego-motion
    edges between consecutive frames of each drone (no loss), UWB distances between drones in the
    same frame (Huber), loop edges from a keyframe to one of its 4 nearest stored keyframes (Huber) and
    detections-as-relative-pose between drones in the same frame (Huber).
    Noise model: simulator.launch:33-62 / swarm_local_sim.cpp:532-550,349-353.
    """
    rng = np.random.default_rng(seed + 7000)
    nd, nf = n_drones, n_frames
    n_nodes = nd * nf
    R, Rz, T = 5.0, 2.0, 50.0
    cols = 3
    d0 = 2.0
    t = np.arange(nf) * 0.5  # one swarm frame every 0.5 s
    gt = np.zeros((nf, nd, 4))
    for i in range(nd):
        gt[:, i, 0] = R * np.sin(2 * math.pi * t / T) + (i // cols) * d0
        gt[:, i, 1] = R * (1 - np.cos(2 * math.pi * t / T)) + (i % cols) * d0
        gt[:, i, 2] = Rz * np.sin(2 * math.pi * t / T) + 0.3 * i
        gt[:, i, 3] = _wrap(0.4 * np.sin(2 * math.pi * t / T + i))
    gt = gt.reshape(n_nodes, 4)

    n_ego = nd * (nf - 1)
    if n_uwb is None:
        n_uwb = 2 * nf
    if n_loop is None:
        n_loop = 2 * nf
    if n_det is None:
        n_det = nf

    types, ia, ib, huber = [], [], [], []
    payload = []

    def add(tp, a, b, pl, hub):
        types.append(tp); ia.append(a); ib.append(b); huber.append(hub)
        p = np.zeros(PAYLOAD_LEN); p[:len(pl)] = pl
        payload.append(p)

    # ego motion: RelativePoseFactor4d::CreateCov6d (factors.hpp:255-263): S = sqrt(|inv(cov4)|)
    vo_cov_pos, vo_cov_yaw = 1e-4, 1e-5
    for f in range(nf - 1):
        for i in range(nd):
            a, b = f * nd + i, (f + 1) * nd + i
            step = max(np.linalg.norm(gt[b, :3] - gt[a, :3]), 0.05)
            cov = np.diag([vo_cov_pos * step] * 3 + [vo_cov_yaw * step])
            meas = _delta_pose(gt[a], gt[b]) + rng.standard_normal(4) * np.sqrt(np.diag(cov))
            meas[3] = _wrap(meas[3])
            S = np.sqrt(np.abs(np.linalg.inv(cov)))
            add(FACTOR_RELPOSE, a, b, np.concatenate([meas, S.reshape(-1)]), 0)
    # UWB distances (solver.cpp:1136-1144): r = (|Ta-Tb| - d) / sqrt(cov)
    uwb_cov = 0.0014
    cnt = 0
    while cnt < n_uwb:
        f = int(rng.integers(0, nf)); i, j = rng.choice(nd, 2, replace=False)
        a, b = f * nd + int(max(i, j)), f * nd + int(min(i, j))   # pairs with idb < ida (:1131)
        d = np.linalg.norm(gt[a, :3] - gt[b, :3]) + rng.standard_normal() * math.sqrt(uwb_cov)
        add(FACTOR_DISTANCE, a, b, [d, 1.0 / math.sqrt(uwb_cov)], 1)
        cnt += 1
    # loops: relative pose with sqrt-information diag(1/sigma) (swarm_local_sim.cpp:459-465)
    loop_cov_pos, loop_cov_yaw = 0.003, 5.2e-4
    S_loop = np.diag([1 / math.sqrt(loop_cov_pos)] * 3 + [1 / math.sqrt(loop_cov_yaw)])
    # as the simulator does (swarm_local_sim.cpp:474-529): a keyframe is linked to one of its nearest stored
    # keyframe positions (any drone), excluding its own temporal neighbours
    frame_of = np.arange(n_nodes) // nd
    drone_of = np.arange(n_nodes) % nd
    for _ in range(n_loop):
        a = int(rng.integers(0, n_nodes))
        dist = np.linalg.norm(gt[:, :3] - gt[a, :3], axis=1)
        dist[(drone_of == drone_of[a]) & (np.abs(frame_of - frame_of[a]) <= 3)] = np.inf
        near = np.argsort(dist, kind="stable")[:4]
        b = int(near[int(rng.integers(0, 4))])
        meas = _delta_pose(gt[a], gt[b]) + rng.standard_normal(4) * np.sqrt([loop_cov_pos] * 3 + [loop_cov_yaw])
        if rng.uniform() < outlier_frac:
            meas[:3] += rng.uniform(-3, 3, 3)
        meas[3] = _wrap(meas[3])
        add(FACTOR_RELPOSE, a, b, np.concatenate([meas, S_loop.reshape(-1)]), 1)
    # detections-as-relative-pose (solver.cpp:542-551): same frame, two drones
    det_cov_pos, det_cov_yaw = 0.01, 0.01
    S_det = np.diag([1 / math.sqrt(det_cov_pos)] * 3 + [1 / math.sqrt(det_cov_yaw)])
    for _ in range(n_det):
        f = int(rng.integers(0, nf)); i, j = rng.choice(nd, 2, replace=False)
        a, b = f * nd + int(i), f * nd + int(j)
        meas = _delta_pose(gt[a], gt[b]) + rng.standard_normal(4) * np.sqrt([det_cov_pos] * 3 + [det_cov_yaw])
        meas[3] = _wrap(meas[3])
        add(FACTOR_RELPOSE, a, b, np.concatenate([meas, S_det.reshape(-1)]), 1)
    # bearing / inverse-depth detections (DroneDetection4dFactor, factors.hpp:273-367)
    for k in range(n_bearing):
        f = int(rng.integers(0, nf)); i, j = rng.choice(nd, 2, replace=False)
        a, b = f * nd + int(i), f * nd + int(j)
        mode = k % 3   # 0: extrinsic-z, no depth; 1: extrinsic-z + inverse depth; 2: dposes + depth
        ext_z = 0.05
        if mode == 2:
            dpa = np.array([0.02, -0.01, 0.03, 0.01]); dpb = np.array([-0.01, 0.02, 0.0, -0.02])
            def mul(p, q):
                c, s = math.cos(p[3]), math.sin(p[3])
                return np.array([p[0] + c * q[0] - s * q[1], p[1] + s * q[0] + c * q[1], p[2] + q[2], _wrap(p[3] + q[3])])
            pa, pb = mul(gt[a], dpa), mul(gt[b], dpb)
        else:
            dpa = dpb = np.zeros(4)
            pa = gt[a].copy(); pa[2] += ext_z; pb = gt[b]
        rel = _delta_pose(pa, pb)[:3]
        rho = 1.0 / np.linalg.norm(rel)
        dirv = rel * rho + rng.standard_normal(3) * 0.01
        dirv /= np.linalg.norm(dirv)
        # tangent base: two unit vectors orthogonal to dir
        tmp = np.array([0, 0, 1.0]) if abs(dirv[2]) < 0.9 else np.array([1.0, 0, 0])
        b0 = np.cross(dirv, tmp); b0 /= np.linalg.norm(b0); b1 = np.cross(dirv, b0)
        inv_dep = rho + rng.standard_normal() * 0.01
        flags = (1 if mode >= 1 else 0) | (2 if mode == 2 else 0)
        sphere_std, invdep_std = 0.03, 0.1
        pl = np.concatenate([dirv, b0, b1, [inv_dep, float(flags), ext_z], dpa, dpb, [sphere_std, invdep_std]])
        add(FACTOR_DETECTION, a, b, pl, 1)

    init = gt.copy()
    init[:, :3] += rng.standard_normal((n_nodes, 3)) * init_sigma_pos
    init[:, 3] = _wrap(init[:, 3] + rng.standard_normal(n_nodes) * init_sigma_yaw)
    fixed = np.zeros(n_nodes, np.uint8)
    fixed[0] = 1                     # first pose of self_id constant (solver.cpp:1196-1199)
    init[0] = gt[0]
    return dict(
        n_nodes=n_nodes, gt=gt, init=init, fixed=fixed,
        ftype=np.array(types, np.int32), ia=np.array(ia, np.int32), ib=np.array(ib, np.int32),
        huber=np.array(huber, np.uint8), payload=np.ascontiguousarray(np.array(payload, np.float64)),
    )


def pose_graph_c5(seed: int = 0) -> dict:
    """BASELINE.json C5 graph: 5 drones x 400 frames = 2000 nodes, 12 000 factors
    = 1995 ego + 4000 UWB + 4000 loop + 2005 detection-as-relative-pose."""
    return pose_graph(5, 400, n_uwb=4000, n_loop=4000, n_det=2005, seed=seed)


# ----------------------------------------------------------------------------------------------------------------
# loop edges of one drone pair for the PCM outlier rejection
# ----------------------------------------------------------------------------------------------------------------
def _quat_from_rotvec(rv):
    a = np.linalg.norm(rv)
    if a < 1e-12:
        return np.array([1.0, 0.5 * rv[0], 0.5 * rv[1], 0.5 * rv[2]])
    return np.concatenate([[np.cos(a / 2)], np.sin(a / 2) * rv / a])


class _PoseAlgebra:
    """(t, unit quaternion wxyz) poses: compose / invert (generator-side helper, independent of the oracle)"""

    @staticmethod
    def q_mul(a, b):
        aw, ax, ay, az = a; bw, bx, by, bz = b
        return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                         aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])

    @staticmethod
    def q_rot(q, v):
        u = q[1:]
        c = np.cross(u, v)
        return v + 2.0 * (q[0] * c + np.cross(u, c))

    @classmethod
    def pose_mul(cls, a, b):
        return np.concatenate([a[:3] + cls.q_rot(a[3:], b[:3]), cls.q_mul(a[3:], b[3:])])

    @classmethod
    def pose_inv(cls, a):
        qc = np.array([a[3], -a[4], -a[5], -a[6]])
        return np.concatenate([-cls.q_rot(qc, a[:3]), qc])


def pcm_edges(n: int = 60, outlier_frac: float = 0.3, seed: int = 0, flip_frac: float = 0.3, id_a: int = 1, id_b: int = 2,
              other_pair: int = 0):
    """n loop edges between drones id_a and id_b flying the circles of swarm_local_sim.cpp (:589-604) with yaw and small
    roll/pitch: inliers = true relative pose + noise of the labelled covariance, outliers = gross errors; a fraction of
    the edges is stored b -> a (LoopEdge::same_robot_pair == 2); `other_pair` extra edges belong to another drone pair.
    Edge dicts as oracle/pcm_ref.py and host.pcm_outlier_rejection take them; e["inlier"] is the ground truth."""
    rng = np.random.default_rng(seed + 7000)
    T, R = 50.0, 5.0
    pr = _PoseAlgebra

    def gt(drone, t):
        ph = 2 * np.pi * t / T
        pos = np.array([R * np.sin(ph) + 2.0 * drone, R * (1 - np.cos(ph)) - 1.5 * drone, 2.0 * np.sin(ph) + 0.3 * drone])
        q = _quat_from_rotvec(np.array([0.03 * np.sin(ph), 0.02 * np.cos(ph), 0.6 * ph + 0.2 * drone]))
        return np.concatenate([pos, q])

    def path_len(t):                          # arc length of the circle (constant speed), metres
        return np.hypot(2 * np.pi * R / T, 2 * np.pi * 2.0 / T * 0.7) * t

    sig_p, sig_a = 0.03, 0.01
    cov = np.diag([sig_p ** 2] * 3 + [sig_a ** 2] * 3)
    edges = []
    for i in range(n + other_pair):
        ta, tb = rng.uniform(0, 40), rng.uniform(0, 40)
        a, b = (id_a, id_b) if i < n else (id_a, id_b + 5)
        pa, pb = gt(a, ta), gt(b, tb)
        rel = pr.pose_mul(pr.pose_inv(pa), pb)
        inlier = rng.uniform() >= outlier_frac
        if inlier:
            noise = np.concatenate([rng.normal(0, sig_p, 3), _quat_from_rotvec(rng.normal(0, sig_a, 3))])
        else:
            noise = np.concatenate([rng.uniform(-3, 3, 3), _quat_from_rotvec(rng.uniform(-0.8, 0.8, 3))])
        rel = pr.pose_mul(rel, noise)
        rel[3:] /= np.linalg.norm(rel[3:])
        e = dict(id_a=a, id_b=b, rel=rel, cov=cov * rng.uniform(0.8, 1.3), odom_a=pa, odom_b=pb, len_a=path_len(ta),
                 len_b=path_len(tb), inlier=bool(inlier) and i < n)
        if i < n and rng.uniform() < flip_frac:                      # the same loop reported b -> a
            e = dict(id_a=b, id_b=a, rel=pr.pose_inv(rel), cov=e["cov"], odom_a=pb, odom_b=pa, len_a=e["len_b"],
                     len_b=e["len_a"], inlier=e["inlier"])
        edges.append(e)
    order = rng.permutation(len(edges))
    return [edges[i] for i in order]


# ----------------------------------------------------------------------------------------------------------------
# 3-D / 2-D correspondences of one loop candidate for the PnP stage (second half)
# ----------------------------------------------------------------------------------------------------------------
def pnp_case(n: int = 200, outlier_frac: float = 0.25, seed: int = 0, noise: float = 0.002, yaw: float = 0.3,
             prior_error: float = 0.15):
    """n landmarks in the NEW drone's odometry frame (matched_3d_now) seen by the OLD camera (matched_2d_norm_old, normalised
    coordinates, K = I): the old drone sits ~1.5 m away with a yaw offset, its camera looks along body +x (extrinsic
    rotation of the usual camera convention).  Returns dict(X [n,3] f32, uv [n,2] f32, inlier [n] bool, pose_true (t,q) with
    x_cam_old = R X + t, prior (t,q) = the truth disturbed by `prior_error` (what odometry would predict), extrinsic,
    drone_pose_now, drone_pose_old)."""
    pa = _PoseAlgebra
    rng = np.random.default_rng(seed + 9000)
    # camera-in-body extrinsic: camera z = body x, camera x = -body y, camera y = -body z
    Rcb = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])

    def quat_from_R(R):
        w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
        return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])

    extrinsic = np.concatenate([[0.05, 0.0, 0.03], quat_from_R(Rcb)])
    # everything the PnP sees lives in the NEW drone's odometry (world) frame Wn: its own pose, the landmarks
    # (matched_3d_now) and therefore the solved old-camera pose; the OLD drone reports its pose in its own gravity-aligned
    # frame Wo, which differs from Wn by a yaw and a translation
    drone_pose_now = np.concatenate([[2.0, -1.0, 1.2], _quat_from_rotvec(np.array([0.01, -0.02, 0.7]))])
    d_old_in_new = np.concatenate([[-1.2, 0.6, 0.1], _quat_from_rotvec(np.array([0.0, 0.0, yaw]))])
    old_in_wn = pa.pose_mul(drone_pose_now, d_old_in_new)
    wo_from_wn = np.concatenate([[4.0, 2.5, -0.4], _quat_from_rotvec(np.array([0.0, 0.0, -1.1]))])
    drone_pose_old = pa.pose_mul(wo_from_wn, old_in_wn)
    cam_old_in_new = pa.pose_mul(old_in_wn, extrinsic)                # old camera pose in Wn
    pose_true = pa.pose_inv(cam_old_in_new)                            # x_cam_old = R X + t
    # landmarks in front of the old camera, expressed in Wn
    Pc = np.stack([rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2.0, 9.0, n)], 1)
    X = np.stack([pa.q_rot(cam_old_in_new[3:], p) + cam_old_in_new[:3] for p in Pc])
    uv = Pc[:, :2] / Pc[:, 2:3] + rng.normal(0, noise, (n, 2))
    inlier = rng.uniform(size=n) >= outlier_frac
    uv[~inlier] = rng.uniform(-1.2, 1.2, ((~inlier).sum(), 2))
    far = np.linalg.norm(uv - Pc[:, :2] / Pc[:, 2:3], axis=1) < 0.2
    inlier = inlier | far                                              # a random "outlier" that happens to fit is an inlier
    prior = pa.pose_mul(np.concatenate([rng.normal(0, prior_error, 3), _quat_from_rotvec(rng.normal(0, prior_error * 0.5, 3))]),
                        pose_true)
    return dict(X=X.astype(np.float32), uv=uv.astype(np.float32), inlier=inlier, pose_true=pose_true, prior=prior,
                extrinsic=extrinsic, drone_pose_now=drone_pose_now, drone_pose_old=drone_pose_old)
