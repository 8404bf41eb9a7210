"""CPU: the oracle of Optimizer::LocalInertialBA (oracle/local_inertial_ba_oracle.cpp; reference src/Optimizer.cc:2383-2958) is pinned by an
independent derivation, since g2o / G2oTypes need Eigen and cannot be compiled into oracle/_ref:
  * its first Levenberg-Marquardt step (lambda = the reference's setUserLambdaInit value) equals the dense damped normal-equation step built
    from NUMERICAL derivatives of the stacked, whitened residuals (EdgeMono, EdgeInertial incl. the down-weighted link to the fixed keyframe,
    EdgeGyroRW, EdgeAccRW) under the reference's update rules (ImuCamPose::Update, additive velocity / bias / point updates), with all points
    eliminated by nothing but numpy's dense solve (no Schur complement);
  * on noisy data it recovers the window and flags the gross outliers; at the noise-free ground truth it stays put."""
import numpy as np

import oracle_lib as O
from orb_slam3_modified_b200 import synth


def _consistent(pr):
    """camera poses derived from the body poses in double (what ImuCamPose::Update computes), so that a zero update is a fixed point"""
    ex = pr['extr']; Rcb, tcb = ex[:9].reshape(3, 3), ex[9:12]
    for k in range(pr['n_kf']):
        R, p = pr['state'][k, :9].reshape(3, 3), pr['state'][k, 9:12]
        pr['tcw'][k, :9] = (Rcb @ R.T).reshape(9); pr['tcw'][k, 9:] = Rcb @ (-R.T @ p) + tcb
    return pr


def test_first_lm_step_is_the_dense_damped_step_from_numerical_derivatives():
    pr = _consistent(synth.local_inertial_ba_problem(n_opt=3, n_cov_fixed=1, n_pts=30, seed=2, outlier_frac=0.0, perturb=0.05, noise_px=0.2, float_inputs=False))
    P = O.liba_preints(pr)
    nO, nL, nE = pr['n_opt'], len(pr['points']), len(pr['e_pt'])
    ex = pr['extr']; Rcb, tcb = ex[:9].reshape(3, 3), ex[9:12]
    infos = [O.imu_information(P[i]) for i in range(nO)]
    chol = []
    for i, (i9, ig, ia) in enumerate(infos):
        chol.append((np.linalg.cholesky(i9 * pr['ie_info_scale'][i] + 1e-18 * np.eye(9)), np.linalg.cholesky(ig), np.linalg.cholesky(ia)))

    def apply(x):
        st = pr['state'].copy(); pts = pr['points'].copy()
        for k in range(nO):
            d = x[15 * k:15 * k + 15]
            R, t = O.imu_pose_update(st[k, :9].reshape(3, 3), st[k, 9:12], d[:6])
            st[k, :9] = R.reshape(9); st[k, 9:12] = t; st[k, 12:15] += d[6:9]; st[k, 15:18] += d[9:12]; st[k, 18:21] += d[12:15]
        pts += x[15 * nO:].reshape(-1, 3)
        return st, pts

    chi_inertial = []

    def residuals(x):
        st, pts = apply(x)
        out = []
        cams = []
        for k in range(pr['n_kf']):
            R, p = st[k, :9].reshape(3, 3), st[k, 9:12]
            cams.append((Rcb @ R.T, Rcb @ (-R.T @ p) + tcb))
        for e in range(nE):
            Rc, tc = cams[pr['e_kf'][e]]
            Xc = Rc @ pts[pr['e_pt'][e]] + tc
            c = pr['cam'][pr['e_kf'][e]].astype(np.float64)
            out.append(np.sqrt(float(pr['inv_sigma2'][e])) * (pr['obs'][e] - np.array([c[0] * Xc[0] / Xc[2] + c[2], c[1] * Xc[1] / Xc[2] + c[3]])))
        for i in range(nO):
            a, b = st[pr['ie_kf1'][i]], st[pr['ie_kf2'][i]]
            e9 = O.imu_edge_inertial(P[i], dict(Rwb1=a[:9].reshape(3, 3), twb1=a[9:12], v1=a[12:15], bg=a[15:18], ba=a[18:21], Rwb2=b[:9].reshape(3, 3), twb2=b[9:12], v2=b[12:15]),
                                     jac=False)[0]
            chi_inertial.append(float(e9 @ (infos[i][0] * pr['ie_info_scale'][i]) @ e9))
            out.append(chol[i][0].T @ e9)
            out.append(chol[i][1].T @ (b[15:18] - a[15:18]))
            out.append(chol[i][2].T @ (b[18:21] - a[18:21]))
        return np.concatenate(out)
    n = 15 * nO + 3 * nL
    r0 = residuals(np.zeros(n))
    # no Huber weight is active at the linearisation point (mono: sqrt(5.991); inertial link: sqrt(16.92))
    assert (r0[:2 * nE].reshape(-1, 2) ** 2).sum(1).max() < 5.99 and chi_inertial[nO - 1] < 16.9
    J = np.zeros((len(r0), n))
    for k in range(n):
        h = 2e-3 if (k < 15 * nO and k % 15 >= 9) else 1e-6     # bias steps must stand out of the float rounding of the preintegrated terms
        d = np.zeros(n); d[k] = h
        J[:, k] = (residuals(d) - residuals(-d)) / (2 * h)
    lam = pr['lambda_init']
    dx = np.linalg.solve(J.T @ J + lam * np.eye(n), -J.T @ r0)
    want_st, want_pts = apply(dx)
    got = O.local_inertial_ba(pr, P, iterations=1)
    assert got['trials'] == 1 and got['iters'] == 1 and not got['failed']
    assert np.allclose(got['state'], want_st, rtol=0, atol=3e-6), np.abs(got['state'] - want_st).max()
    assert np.allclose(got['points'], want_pts, rtol=0, atol=3e-5), np.abs(got['points'] - want_pts).max()
    # the robust cost the oracle reports before the step is the squared norm of the stacked residuals
    assert abs(got['err'] - float(r0 @ r0)) < 1e-3 * float(r0 @ r0)


def test_recovers_the_window_and_flags_gross_outliers():
    for seed, large in ((1, False), (3, True)):
        pr = synth.local_inertial_ba_problem(n_opt=10 if not large else 14, n_cov_fixed=3, n_pts=500, seed=seed, large=large)
        P = O.liba_preints(pr)
        r = O.local_inertial_ba(pr, P)
        nO = pr['n_opt']
        assert not r['failed'] and r['err_end'] < 0.05 * r['err'] and 1 <= r['iters'] <= pr['iterations']
        before = np.abs(pr['state'][:nO, 9:12] - pr['truth'][:nO, 9:12]).max(); after = np.abs(r['state'][:nO, 9:12] - pr['truth'][:nO, 9:12]).max()
        assert after < 0.5 * before, (before, after)
        assert np.abs(r['state'][:nO, 12:15] - pr['truth'][:nO, 12:15]).max() < 0.02
        # fixed keyframes do not move
        assert np.array_equal(r['state'][nO:], pr['state'][nO:]) and np.array_equal(r['tcw'][nO:], pr['tcw'][nO:])
        # the optimised camera pose is the one ImuCamPose::Update derives from the body pose
        ex = pr['extr']; Rcb, tcb = ex[:9].reshape(3, 3), ex[9:12]
        for k in range(nO):
            R, p = r['state'][k, :9].reshape(3, 3), r['state'][k, 9:12]
            assert np.allclose(r['tcw'][k, :9].reshape(3, 3), Rcb @ R.T, atol=1e-12) and np.allclose(r['tcw'][k, 9:], Rcb @ (-R.T @ p) + tcb, atol=1e-12)
            assert np.allclose(R.T @ R, np.eye(3), atol=1e-9)
        # erased observations: residual test of :2848-2862 on the final state
        res = O.local_inertial_ba_residuals(pr, r['tcw'], r['points'])
        chi = pr['inv_sigma2'] * (res ** 2).sum(1)
        close = pr['track_depth'][pr['e_pt']] < 10
        want = np.where(close, chi > 1.5 * 5.991, chi > 5.991)
        assert (want != r['erase'].astype(bool)).mean() < 0.01          # the edges hold the errors of the last evaluated state; depth is positive here
        assert 0.01 < r['erase'].mean() < 0.12


def test_noise_free_ground_truth_is_a_fixed_point():
    pr = _consistent(synth.local_inertial_ba_problem(n_opt=5, n_cov_fixed=2, n_pts=120, seed=4, outlier_frac=0.0, perturb=0.0, noise_px=0.0, float_inputs=False))
    pr['imu'] = [synth.imu_interval(pr['times'][i + 1], pr['times'][i], rate=2000.0, seed=i, bias=tuple(pr['bias6']), noise=False) for i in range(pr['n_opt'])]
    P = O.liba_preints(pr)
    r = O.local_inertial_ba(pr, P)
    assert not r['failed'] and r['erase'].sum() == 0
    assert np.abs(r['state'][:, 9:12] - pr['truth'][:, 9:12]).max() < 2e-3 and np.abs(r['points'] - pr['points_true']).max() < 2e-2
