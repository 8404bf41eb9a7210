"""CPU tests of the LBA oracle: g2o-semantics sanity on seeded problems (the reference ships no LBA vectors)."""
import numpy as np

import oracle_lib as O
from orb_slam3_modified_b200 import synth


def test_lba_converges_and_culls_outliers():
    p = synth.lba_problem(n_kf=8, n_pts=600, obs_per_pt=5, seed=1)
    r0 = O.lba_residuals(p, p['poses'], p['points'])
    out = O.lba_solve(p)
    r1 = O.lba_residuals(p, out['poses'], out['points'])
    assert out['iters'] >= 3 and out['stats'][1] < 0.5 * out['stats'][2]          # robust chi2 dropped
    assert np.median(np.abs(r1)) < 0.5 * np.median(np.abs(r0))
    frac_out = float((out['chi2'] > 5.991).mean())
    assert 0.02 < frac_out < 0.2                                                   # ~5% gross outliers injected
    assert out['depth_pos'].all()
    assert np.allclose(out['poses'][0], p['poses'][0], atol=1e-7)                  # fixed keyframe untouched (only re-normalised, SE3Quat ctor)
    assert np.allclose(np.linalg.norm(out['poses'][:, :4], axis=1), 1.0, atol=1e-12)


def test_lba_noise_free_reaches_ground_truth_residuals():
    p = synth.lba_problem(n_kf=6, n_pts=300, obs_per_pt=6, seed=2, outlier_frac=0.0, n_fixed=2)
    gt = O.lba_residuals(p, p['gt_poses'], p['gt_points'])
    out = O.lba_solve(p, iterations=20)
    r = O.lba_residuals(p, out['poses'], out['points'])
    assert np.sqrt((r ** 2).mean()) <= 1.05 * np.sqrt((gt ** 2).mean())


def test_lba_stop_flag_and_zero_iterations():
    p = synth.lba_problem(n_kf=5, n_pts=100, obs_per_pt=4, seed=3)
    out = O.lba_solve(p, stop_flag=np.ones(1, np.int32))
    assert out['iters'] == 0 and np.array_equal(out['points'], p['points'])
    out = O.lba_solve(p, iterations=0)
    assert out['iters'] == 0


def test_lba_user_lambda_init_changes_first_step():
    p = synth.lba_problem(n_kf=5, n_pts=100, obs_per_pt=4, seed=4)
    a = O.lba_solve(p, iterations=1)
    b = O.lba_solve(p, iterations=1, user_lambda_init=100.0)   # inertial maps (Optimizer.cc:1197-1198)
    assert not np.allclose(a['points'], b['points'])


def test_pose_optimization_oracle_recovers_pose_and_flags_outliers():
    f = synth.pose_opt_problem(n=500, seed=3, outlier_frac=0.2)
    out = O.pose_optimization(f)
    assert 350 <= out['inliers'] <= 420
    assert 0.1 < out['outlier'].mean() < 0.3
    assert np.abs(out['pose'][4:] - f['gt_pose'][4:]).max() < 0.03
