"""Oracle (TEST INFRASTRUCTURE ONLY) for the geometric filter of the loop matcher: the homography-RANSAC inlier mask of
LoopDetector::compute_correspond_features (/root/reference/swarm_loop/src/loop_detector.cpp:569-598):

    keep the cross-check matches whose NEW landmark has a 3-D flag (:572-586); if at least 4 remain,
    cv::findHomography(old_2d, new_2d, CV_RANSAC, 3, mask) and keep the masked matches (:589-598); else reject the pair.

The reference's RANSAC is OpenCV's (3.4, swarm_loop/CMakeLists.txt:26), which draws its samples from cv::RNG: its mask is
not reproducible run to run or across OpenCV versions, so the library defines a DETERMINISTIC RANSAC with the same model,
error and threshold, restated here operation by operation (IEEE double, no fused multiply-add) so that the CUDA kernel can be
bit-exact against it:

  * hypothesis h = 0..n_hyp-1 draws 4 distinct match indices from a counter-based hash (lowbias32 of seed, h, slot, try);
  * H maps old -> new (h33 = 1): 8x8 linear system, Gaussian elimination with partial pivoting; singular -> skipped;
  * a match is an inlier iff |new - H old|^2 <= thresh^2 (OpenCV: findInliers, err <= thresh*thresh), evaluated in the
    division-free form (u w - px)^2 + (v w - py)^2 <= thresh^2 w^2;
  * the winner has the most inliers, ties to the smaller h (OpenCV keeps the first best); mask = its inliers.

Pinned against the real OpenCV (cv2.findHomography(..., cv2.RANSAC, 3.0)) in tests/test_oracle_pins.py: on correspondences
whose outliers are far from any consensus the two masks are identical; near the 3-pixel threshold they may differ because the
reference itself is randomised -- that part is "parity unpinned" by construction.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

N_HYP = 512
M32 = 0xFFFFFFFF


def lowbias32(x: int) -> int:
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def draw4(seed: int, h: int, n: int):
    """4 distinct indices in [0, n) for hypothesis h; None if 16 tries per slot do not give distinct ones."""
    idx = []
    for slot in range(4):
        ok = False
        for t in range(16):
            v = lowbias32(seed ^ lowbias32((h * 4 + slot) * 16 + t + 0x9E3779B9)) % n
            if v not in idx:
                idx.append(v); ok = True
                break
        if not ok:
            return None
    return idx


def solve_h(src4: np.ndarray, dst4: np.ndarray):
    """H (9 doubles, h33 = 1) with dst ~ H src from 4 correspondences; None if singular."""
    A = np.zeros((8, 9), np.float64)
    for i in range(4):
        x, y = np.float64(src4[i, 0]), np.float64(src4[i, 1])
        u, v = np.float64(dst4[i, 0]), np.float64(dst4[i, 1])
        A[2 * i] = [x, y, 1.0, 0.0, 0.0, 0.0, -(u * x), -(u * y), u]
        A[2 * i + 1] = [0.0, 0.0, 0.0, x, y, 1.0, -(v * x), -(v * y), v]
    for c in range(8):
        p = c
        best = abs(A[c, c])
        for r in range(c + 1, 8):
            if abs(A[r, c]) > best:
                best = abs(A[r, c]); p = r
        if not best > 1e-9:
            return None
        if p != c:
            A[[c, p]] = A[[p, c]]
        inv = np.float64(1.0) / A[c, c]
        for r in range(c + 1, 8):
            f = A[r, c] * inv
            for k in range(c, 9):
                A[r, k] = A[r, k] - f * A[c, k]
    h = np.zeros(9, np.float64)
    h[8] = 1.0
    for c in range(7, -1, -1):
        s = A[c, 8]
        for k in range(c + 1, 8):
            s = s - A[c, k] * h[k]
        h[c] = s / A[c, c]
    return h


def inliers(h: np.ndarray, src: np.ndarray, dst: np.ndarray, thresh: float) -> np.ndarray:
    """|new - H old|^2 <= thresh^2, stated without the division: with (px, py, w) = H (x, y, 1),
    (u w - px)^2 + (v w - py)^2 <= thresh^2 w^2."""
    x = src[:, 0].astype(np.float64); y = src[:, 1].astype(np.float64)
    w = (h[6] * x + h[7] * y) + 1.0
    px = (h[0] * x + h[1] * y) + h[2]
    py = (h[3] * x + h[4] * y) + h[5]
    ex = dst[:, 0].astype(np.float64) * w - px
    ey = dst[:, 1].astype(np.float64) * w - py
    t2 = np.float64(thresh) * np.float64(thresh)
    return (ex * ex + ey * ey) <= t2 * (w * w)


def homography_ransac_mask(src: np.ndarray, dst: np.ndarray, thresh: float = 3.0, seed: int = 0, n_hyp: int = N_HYP):
    """src = old_2d, dst = new_2d, [n, 2] float32 -> (mask uint8 [n], n_inliers, winning hypothesis or -1).
    n < 4: nothing passes (the reference rejects the pair, loop_detector.cpp:598-600)."""
    n = src.shape[0]
    mask = np.zeros(n, np.uint8)
    if n < 4:
        return mask, 0, -1
    best, best_h, best_m = -1, -1, None
    for hyp in range(n_hyp):
        idx = draw4(seed, hyp, n)
        if idx is None:
            continue
        h = solve_h(src[idx], dst[idx])
        if h is None:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            m = inliers(h, src, dst, thresh)
        c = int(m.sum())
        if c > best:
            best, best_h, best_m = c, hyp, m
    if best_h < 0:
        return mask, 0, -1
    return best_m.astype(np.uint8), best, best_h


def loop_pair_filter(match_new, match_old, flags_new, kpts_new, kpts_old, thresh: float = 3.0, seed: int = 0):
    """The whole filter of loop_detector.cpp:569-598 on one direction pair -> (kept new idx, kept old idx) or None when
    fewer than 4 flagged matches remain (the reference returns false)."""
    keep = [i for i, q in enumerate(match_new) if flags_new[q]]
    if len(keep) < 4:
        return None
    qn = np.asarray([match_new[i] for i in keep]); qo = np.asarray([match_old[i] for i in keep])
    m, _, _ = homography_ransac_mask(kpts_old[qo], kpts_new[qn], thresh, seed)
    return qn[m != 0], qo[m != 0]
