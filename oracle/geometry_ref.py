"""Oracle (TEST INFRASTRUCTURE ONLY) for the geometric filter of the loop matcher: the homography-RANSAC inlier mask of
LoopDetector::compute_correspond_features (/root/reference/swarm_loop/src/loop_detector.cpp:569-598):

    keep the cross-check matches whose NEW landmark has a 3-D flag (:572-586); if at least 4 remain,
    cv::findHomography(old_2d, new_2d, CV_RANSAC, 3, mask) and keep the masked matches (:589-598); else reject the pair.

The reference's RANSAC is OpenCV's (3.4, swarm_loop/CMakeLists.txt:26), which draws its samples from cv::RNG: its mask is
not reproducible run to run or across OpenCV versions, so the library defines a DETERMINISTIC RANSAC with the same model,
error and threshold, restated here operation by operation (IEEE double, no fused multiply-add) so that the CUDA kernel can be
bit-exact against it:

  * hypothesis h = 0..n_hyp-1 draws 4 distinct match indices from a counter-based hash (lowbias32 of seed, h, slot, try);
  * H maps old -> new: closed-form projective-basis construction H = frame(new) adj(frame(old)), unnormalised (the inlier
    test is homogeneous in H); a sample with a triangle of twice-area <= 1 px^2 is degenerate and skipped;
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


def _frame(p: np.ndarray):
    """[p1 p2 p3] diag(adj([p1 p2 p3]) p4) for 4 points (homogeneous, unnormalised) -> (3x3 row-major list, ok)."""
    f = np.float64
    x1, y1, x2, y2, x3, y3, x4, y4 = (f(p[0, 0]), f(p[0, 1]), f(p[1, 0]), f(p[1, 1]), f(p[2, 0]), f(p[2, 1]), f(p[3, 0]),
                                      f(p[3, 1]))
    a00 = y2 - y3; a01 = x3 - x2; a02 = x2 * y3 - x3 * y2
    a10 = y3 - y1; a11 = x1 - x3; a12 = x3 * y1 - x1 * y3
    a20 = y1 - y2; a21 = x2 - x1; a22 = x1 * y2 - x2 * y1
    det = (a02 + a12) + a22
    v0 = (a00 * x4 + a01 * y4) + a02
    v1 = (a10 * x4 + a11 * y4) + a12
    v2 = (a20 * x4 + a21 * y4) + a22
    ok = abs(det) > 1.0 and abs(v0) > 1.0 and abs(v1) > 1.0 and abs(v2) > 1.0     # no triangle of twice-area <= 1 px^2
    return [x1 * v0, x2 * v1, x3 * v2, y1 * v0, y2 * v1, y3 * v2, v0, v1, v2], ok


def solve_h(src4: np.ndarray, dst4: np.ndarray):
    """H (9 doubles, row-major, UNNORMALISED) with dst ~ H src from 4 correspondences by the projective-basis
    construction H = frame(dst) adj(frame(src)); None for a degenerate sample."""
    a, oka = _frame(src4)
    b, okb = _frame(dst4)
    if not (oka and okb):
        return None
    c = [a[4] * a[8] - a[5] * a[7], a[2] * a[7] - a[1] * a[8], a[1] * a[5] - a[2] * a[4],
         a[5] * a[6] - a[3] * a[8], a[0] * a[8] - a[2] * a[6], a[2] * a[3] - a[0] * a[5],
         a[3] * a[7] - a[4] * a[6], a[1] * a[6] - a[0] * a[7], a[0] * a[4] - a[1] * a[3]]
    h = np.zeros(9, np.float64)
    for i in range(3):
        for j in range(3):
            h[i * 3 + j] = (b[i * 3] * c[j] + b[i * 3 + 1] * c[3 + j]) + b[i * 3 + 2] * c[6 + j]
    return h


def inliers(h: np.ndarray, src: np.ndarray, dst: np.ndarray, thresh: float) -> np.ndarray:
    """|new - H old|^2 <= thresh^2, stated without the division: with (px, py, w) = H (x, y, 1),
    (u w - px)^2 + (v w - py)^2 <= thresh^2 w^2."""
    x = src[:, 0].astype(np.float64); y = src[:, 1].astype(np.float64)
    w = (h[6] * x + h[7] * y) + h[8]
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
