"""GPU parity of the geometric filter: homography-RANSAC masks bit-exact against oracle/geometry_ref.py
(which tests/test_oracle_pins.py pins against cv2.findHomography)."""
import numpy as np
import pytest

from omniswarm_b200 import host
from oracle import geometry_ref as gr
from test_oracle_pins import _homography_case

pytestmark = pytest.mark.gpu


def test_homography_masks_bit_exact(gpu):
    cases = [_homography_case(s) for s in range(4)]
    cases.append(_homography_case(7, n=37, n_out=20, noise=1.5))           # heavy noise: many borderline points
    cases.append(_homography_case(8, n=200, n_out=150))                    # 25 % inliers
    cases.append(tuple(a[:3] for a in _homography_case(9)))                # fewer than 4 points
    same = np.repeat(cases[0][0][:1], 12, 0)
    cases.append((same, same, np.zeros(12, np.uint8)))                     # degenerate: no valid hypothesis
    cases.append((cases[0][0][:4], cases[0][1][:4], None))                 # exactly 4
    for seed in (0, 12345):
        out = host.homography_ransac([c[0] for c in cases], [c[1] for c in cases], 3.0, seed)
        for (src, dst, _), (m, n_inl, win) in zip(cases, out):
            rm, rc, rw = gr.homography_ransac_mask(src, dst, 3.0, seed)
            assert np.array_equal(m, rm) and n_inl == rc and win == rw


def test_homography_equals_opencv_on_separated_data(gpu):
    import cv2
    src, dst, truth = _homography_case(11)
    (m, n_inl, _), = host.homography_ransac([src], [dst])
    _, mc = cv2.findHomography(src, dst, cv2.RANSAC, 3.0)
    assert np.array_equal(m, mc.ravel()) and np.array_equal(m, truth) and n_inl == int(truth.sum())
