"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the PCM outlier rejection that runs immediately
upstream of the pose-graph solve.

Follows (paths relative to /root/reference):
  * SwarmLocalOutlierRejection::OutlierRejectionLoopEdgesPCM
      swarm_localization/src/swarm_outlier_rejection/swarm_outlier_rejection.cpp:173-297 -- pairwise consistency of the
      loop edges of one drone pair: err = odom_a * p_edge2 * odom_b^-1 * p_edge1^-1 (:227), 6-D log map (:228), squared
      Mahalanobis distance against cov_1 + cov_2 + cov(odom_a) + cov(odom_b) (:193,212,224,229), edge in the consistency
      graph iff smd < pcm_thres (:231-235); edge1 is the LATER-inserted loop of the pair (:180,190);
  * FMC::maxCliqueHeu (Josh Mangelson's variant that returns the clique)
      swarm_localization/src/swarm_outlier_rejection/third_party/fast_max-clique_finder/src/findCliqueHeu.cpp:120-244
      -- restated literally, prunings 1/3/5 and the "last element of S" choice (:185) included; PINNED against the
      reference's own library compiled from these very sources (oracle/fmc_ref.py, oracle/_ref/libfmc_ref.so):
      tests/test_oracle_pins.py::test_max_clique_restatement_equals_the_references_own_library.

Third-party arithmetic that is ABSENT from the reference tree (HKUST-Swarm/swarm_msgs, un-vendored, no commit pinned):
`Swarm::Pose` composition / inverse / `log_map`, `LoopEdge::get_covariance`, `computeSquaredMahalanobisDistance` and
`DroneTrajectory::get_relative_pose_by_ts`.  They are DEFINED here (parity unpinned at that boundary, as for the factors):
  * Pose = (t, unit quaternion wxyz); a*b = (t_a + R_a t_b, q_a q_b); inverse = (-R^T t, q*);
  * log_map(T) = [t ; rotation vector of q]  (6-vector, translation first -- the order of get_covariance's blocks);
  * smd(v, C) = v^T C^-1 v;
  * the covariances and the ego-motion poses are INPUTS: every loop edge carries its 6x6 covariance, the ego-motion
    (odometry-frame) pose of both drones at its two stamps and the accumulated trajectory length there; the odometry
    between two stamps of a drone is pose(ts1)^-1 * pose(ts2) with covariance |len(ts2) - len(ts1)| * diag(pos_cov_per_m x3,
    ang_cov_per_m x3) -- the per-metre model of the simulator's labels (swarm_local_sim.cpp:532-550).
"""
from __future__ import annotations

import numpy as np


# ---- Swarm::Pose algebra (defined here; swarm_msgs is not in the tree) ----------------------------------------
def q_mul(a, b):
    aw, ax, ay, az = a; bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw])


def q_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def q_rot(q, v):
    """R(q) v, expanded (no matrix): v + 2 w (u x v) + 2 u x (u x v)"""
    u = q[1:]
    c = np.cross(u, v)
    return v + 2.0 * (q[0] * c + np.cross(u, c))


def pose_mul(a, b):
    return np.concatenate([a[:3] + q_rot(a[3:], b[:3]), q_mul(a[3:], b[3:])])


def pose_inv(a):
    qc = q_conj(a[3:])
    return np.concatenate([-q_rot(qc, a[:3]), qc])


def log_map(p):
    """[translation ; rotation vector]"""
    q = p[3:]
    if q[0] < 0:
        q = -q
    n = np.sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    if n < 1e-12:
        rv = 2.0 * q[1:]
    else:
        rv = (2.0 * np.arctan2(n, q[0]) / n) * q[1:]
    return np.concatenate([p[:3], rv])


def cholesky_smd(v, C):
    """v^T C^-1 v through an unpivoted Cholesky factorisation, the arithmetic the kernel performs"""
    n = len(v)
    L = np.zeros((n, n))
    for j in range(n):
        s = C[j, j] - np.dot(L[j, :j], L[j, :j])
        if not s > 0:
            return np.inf
        L[j, j] = np.sqrt(s)
        for i in range(j + 1, n):
            L[i, j] = (C[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    y = np.zeros(n)
    for i in range(n):
        y[i] = (v[i] - np.dot(L[i, :i], y[:i])) / L[i, i]
    return float(np.dot(y, y))


def same_robot_pair(e1, e2):
    """LoopEdge::same_robot_pair: 1 same orientation, 2 swapped, 0 another pair"""
    if e1["id_a"] == e2["id_a"] and e1["id_b"] == e2["id_b"]:
        return 1
    if e1["id_a"] == e2["id_b"] and e1["id_b"] == e2["id_a"]:
        return 2
    return 0


def pair_smd(e1, e2, pos_cov_per_m, ang_cov_per_m):
    """squared Mahalanobis consistency error of (edge1 = later loop, edge2 = earlier loop), or None for another pair
    (swarm_outlier_rejection.cpp:190-229)"""
    srp = same_robot_pair(e1, e2)
    if srp == 0:
        return None
    if srp == 1:                                                   # :200-212
        p2 = e2["rel"]
        a2, la2, b2, lb2 = e2["odom_a"], e2["len_a"], e2["odom_b"], e2["len_b"]
    else:                                                          # :214-224 edge2 runs b -> a
        p2 = pose_inv(e2["rel"])
        a2, la2, b2, lb2 = e2["odom_b"], e2["len_b"], e2["odom_a"], e2["len_a"]
    odom_a = pose_mul(pose_inv(e1["odom_a"]), a2)                  # drone id_a: ts_a(edge1) -> its stamp on edge2
    odom_b = pose_mul(pose_inv(e1["odom_b"]), b2)
    da, db = abs(la2 - e1["len_a"]), abs(lb2 - e1["len_b"])
    C = e1["cov"] + e2["cov"] + np.diag([pos_cov_per_m] * 3 + [ang_cov_per_m] * 3) * (da + db)
    err = pose_mul(pose_mul(pose_mul(odom_a, p2), pose_inv(odom_b)), pose_inv(e1["rel"]))          # :227
    return cholesky_smd(log_map(err), C)


def consistency_matrix(edges, pcm_thres, pos_cov_per_m, ang_cov_per_m):
    """adj [n,n] uint8, symmetric, zero diagonal: 1 iff smd < pcm_thres (:231-235); smd [n,n] (inf where undefined)"""
    n = len(edges)
    adj = np.zeros((n, n), np.uint8)
    smd = np.full((n, n), np.inf)
    for i in range(n):
        for j in range(i):
            s = pair_smd(edges[i], edges[j], pos_cov_per_m, ang_cov_per_m)     # edge1 = the later one
            if s is None:
                continue
            smd[i, j] = smd[j, i] = s
            if s < pcm_thres:
                adj[i, j] = adj[j, i] = 1
    return adj, smd


def max_clique_heu(adj):
    """FMC::maxCliqueHeu (findCliqueHeu.cpp:120-244) on the adjacency matrix; neighbour lists ascending (the order in which
    OutlierRejectionLoopEdgesPCM fills pcm_graph).  -> (clique vertex list in the reference's order, maxClq)"""
    n = adj.shape[0]
    nbr = [np.nonzero(adj[v])[0].tolist() for v in range(n)]
    deg = [len(x) for x in nbr]
    max_clq, best = -1, []
    for v in range(n):
        if max_clq > deg[v]:                                        # pruning 1 (:149)
            continue
        S = [v] + [u for u in nbr[v] if max_clq <= deg[u]]          # :156-165 pruning 3
        inter = [v]                                                 # :174
        icc = 0
        while S:                                                    # :176
            icc += 1
            imdv = S[-1]                                            # :185
            row = adj[imdv]
            S1 = [u for u in S if row[u] and max_clq <= deg[u]]     # :190-203 pruning 5
            if S1:
                inter.append(imdv)                                  # :218-221
            S = S1
        if max_clq < icc:                                           # :236-239
            max_clq, best = icc, inter
    return best, max_clq


def pcm(edges, pcm_thres, pos_cov_per_m, ang_cov_per_m):
    """-> (indices of the loops kept (good_loops_set, :291-296), adjacency, clique size)"""
    adj, _ = consistency_matrix(edges, pcm_thres, pos_cov_per_m, ang_cov_per_m)
    clique, size = max_clique_heu(adj)
    return clique, adj, size


def max_clique_exact(adj):
    """brute force, for pinning the heuristic on small graphs"""
    n = adj.shape[0]
    best = 0
    for mask in range(1 << n):
        vs = [i for i in range(n) if mask >> i & 1]
        if len(vs) <= best:
            continue
        if all(adj[a, b] for k, a in enumerate(vs) for b in vs[k + 1:]):
            best = len(vs)
    return best
