"""GPU parity of the loop candidate's relative pose (osb_pnp_ransac: deterministic PnP-RANSAC, LM refinement, RPerror,
pnp_result_verify, odometry consistency) against oracle/pnp_ref.py."""
import numpy as np
import pytest

from omniswarm_b200 import synth, host
from oracle import pnp_ref as pn, pcm_ref as pr

pytestmark = pytest.mark.gpu


def case(seed, n=200, out=0.25, **kw):
    c = synth.pnp_case(n, out, seed)
    c.update(dict(iterations=100, thresh=0.03, seed=seed, is_4dof=1, min_loop_num=15, rperr_thres=0.1,
                  accept_loop_yaw_rad=0.8, max_loop_dis=5.0))
    c.update(kw)
    return c


def check(c, mask, r):
    ref = pn.pnp_ransac(c["X"], c["uv"], c["prior"], c["iterations"], c["thresh"], c["seed"])
    assert bool(r.pnp_success) == ref["success"]
    assert r.n_inliers == ref["n_inliers"] and r.winner == ref["winner"]
    assert np.array_equal(mask, ref["mask"])
    if not ref["success"]:
        assert r.verified == 0
        return
    assert np.abs(np.array(r.pose_cam) - ref["pose"]).max() < 1e-8
    v = pn.loop_from_pnp(ref, c)
    assert np.abs(np.array(r.dp_old_to_new) - v["dp"]).max() < 1e-8 and abs(r.rperr - v["rperr"]) < 1e-8
    assert bool(r.verified) == v["verified"] and bool(r.odometry_consistent) == v["odometry_consistent"]
    assert abs(r.md - v["md"]) <= 1e-7 * max(1.0, v["md"])


def test_pnp_matches_oracle(gpu):
    cases = [case(0), case(1, out=0.45), case(2, n=40, out=0.0), case(3, thresh=3.0),          # the reference's threshold
             case(4, is_4dof=0), case(5, iterations=1000, min_loop_num=8), case(6, n=800, out=0.3)]
    out = host.pnp_ransac(cases)
    for c, (mask, r) in zip(cases, out):
        check(c, mask, r)
    m0, r0 = out[0]
    assert r0.verified == 1 and np.array_equal(m0.astype(bool), cases[0]["inlier"])
    assert np.abs(np.array(r0.pose_cam)[:3] - cases[0]["pose_true"][:3]).max() < 0.02
    assert out[3][1].n_inliers == 200                                   # threshold 3 in normalised units: everything fits


def test_pnp_same_drone_odometry_check_and_rejections(gpu):
    base = case(7)
    ref = pn.pnp_ransac(base["X"], base["uv"], base["prior"], 100, 0.03, 7)
    d = np.concatenate([pn.loop_from_pnp(ref, base)["dp"][:3], pn.quat_from_rotvec(np.array([0, 0, pn.loop_from_pnp(ref, base)["dp"][3]]))])
    cov = np.diag([0.05 ** 2] * 3 + [0.02 ** 2] * 3)
    good = dict(base, same_drone=1, odom_rel=d, cov=cov, odometry_consistency_threshold=15.0)
    bad_odom = np.concatenate([d[:3] + np.array([1.0, -0.8, 0.3]), d[3:]])
    bad = dict(base, same_drone=1, odom_rel=bad_odom, cov=cov, odometry_consistency_threshold=15.0)
    strict = dict(base, min_loop_num=10_000)                            # too few inliers for MIN_LOOP_NUM
    near = dict(base, max_loop_dis=0.5)                                 # farther than MAX_LOOP_DIS
    few = dict(case(8, n=3, out=0.0))                                   # fewer than 4 points: no model
    out = host.pnp_ransac([good, bad, strict, near, few])
    for c, (mask, r) in zip([good, bad, strict, near, few], out):
        check(c, mask, r)
    assert out[0][1].verified == 1 and out[0][1].odometry_consistent == 1
    assert out[1][1].verified == 1 and out[1][1].odometry_consistent == 0 and out[1][1].md > 15.0
    assert out[2][1].verified == 0 and out[3][1].verified == 0
    assert out[4][1].pnp_success == 0 and out[4][1].n_inliers == 0
