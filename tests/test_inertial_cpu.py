"""CPU: the inertial oracle (oracle/inertial_oracle.cpp) is pinned by independent derivations, since G2oTypes / ImuTypes need Eigen
and cannot be compiled into oracle/_ref:
  * the analytic Jacobians of EdgeInertial::linearizeOplus and EdgeMono::linearizeOplus equal numerical derivatives of the residuals
    under the reference's own update rules (ImuCamPose::Update for poses, additive for velocity / biases / points);
  * preintegration of noise-free IMU data of an analytic trajectory reproduces the trajectory's relative motion, and the
    EdgeInertial residual vanishes at the ground truth;
  * NormalizeRotation returns the polar factor (checked against numpy's SVD), Exp / Log / Jr / Jr^-1 satisfy their identities;
  * the information matrix is the symmetric inverse of the covariance."""
import numpy as np

import oracle_lib as O
from orb_slam3_modified_b200 import synth


def test_so3_helpers():
    rng = np.random.default_rng(0)
    for _ in range(50):
        w = rng.normal(0, 0.8, 3)
        R = O.so3('exp', w)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1)
        assert np.allclose(O.so3('log', R), w, atol=1e-9)
        assert np.allclose(O.so3('Jr', w) @ O.so3('invJr', w), np.eye(3), atol=1e-9)
        M = R + rng.normal(0, 1e-3, (3, 3))
        U, _, Vt = np.linalg.svd(M)
        assert np.allclose(O.so3('normalize', M), U @ Vt, atol=1e-12)
        d = rng.normal(0, 1e-6, 3)     # Exp(w + d) ~ Exp(w) Exp(Jr(w) d)
        assert np.allclose(O.so3('exp', w + d), R @ O.so3('exp', O.so3('Jr', w) @ d), atol=1e-10)


def test_preintegration_reproduces_the_trajectory():
    t0, t1 = 1.0, 1.0 + 1.0 / 3
    bias = (0.02, -0.01, 0.03, 0.002, -0.001, 0.0015)
    acc, gyr, dts = synth.imu_interval(t0, t1, rate=3000.0, bias=bias, noise=False)
    P = O.imu_preintegrate(acc, gyr, dts, bias, synth.IMU_NOISE)
    R1, p1, v1, _, _ = synth.imu_trajectory(t0)
    R2, p2, v2, _, _ = synth.imu_trajectory(t1)
    dt = float(P[0])
    g = np.array([0, 0, -9.81])
    dR, dV, dP = O.imu_delta(P, bias[3:], bias[:3])
    assert np.isclose(dt, t1 - t0, atol=1e-5)
    assert np.allclose(dR, R1.T @ R2, atol=2e-5)
    assert np.allclose(dV, R1.T @ (v2 - v1 - g * dt), atol=2e-4)
    assert np.allclose(dP, R1.T @ (p2 - p1 - v1 * dt - 0.5 * g * dt * dt), atol=1e-4)
    s = synth.inertial_edge_state(t0, t1, perturb=0.0)
    err, _ = O.imu_edge_inertial(P, s)
    assert np.abs(err).max() < 5e-4
    info, ig, ia = O.imu_information(P)
    C9 = P[67:].reshape(15, 15)[:9, :9].astype(np.float64)
    ni = np.linalg.inv(C9); ni = (ni + ni.T) / 2
    assert np.allclose(info, info.T) and np.abs(info - ni).max() < 1e-6 * np.abs(ni).max() and np.linalg.eigvalsh(info).min() > 0
    assert np.allclose(ig @ P[67:].reshape(15, 15)[9:12, 9:12].astype(np.float64), np.eye(3), atol=1e-9)


def _numeric_jacobian_inertial(P, s, h_state=1e-6, h_bias=2e-3):
    J = np.zeros((9, 24))
    e0, _ = O.imu_edge_inertial(P, s, jac=False)

    def col(k, sp, sm, h):
        ep, _ = O.imu_edge_inertial(P, sp, jac=False)
        em, _ = O.imu_edge_inertial(P, sm, jac=False)
        J[:, k] = (ep - em) / (2 * h)
    for which, c0 in (('1', 0), ('2', 15)):
        for k in range(6):
            pu = np.zeros(6); pu[k] = h_state
            sp, sm = dict(s), dict(s)
            sp['Rwb' + which], sp['twb' + which] = O.imu_pose_update(s['Rwb' + which], s['twb' + which], pu)
            sm['Rwb' + which], sm['twb' + which] = O.imu_pose_update(s['Rwb' + which], s['twb' + which], -pu)
            col(c0 + k, sp, sm, h_state)
    for name, c0, h in (('v1', 6, h_state), ('bg', 9, h_bias), ('ba', 12, h_bias), ('v2', 21, h_state)):
        for k in range(3):
            d = np.zeros(3); d[k] = h
            sp, sm = dict(s), dict(s)
            sp[name] = s[name] + d; sm[name] = s[name] - d
            col(c0 + k, sp, sm, h)
    return J


def test_edge_inertial_jacobians_equal_numeric_derivatives():
    for seed in range(4):
        t0 = 0.5 + seed
        acc, gyr, dts = synth.imu_interval(t0, t0 + 0.4, seed=seed)
        P = O.imu_preintegrate(acc, gyr, dts, (0.02, -0.01, 0.03, 0.002, -0.001, 0.0015), synth.IMU_NOISE)
        s = synth.inertial_edge_state(t0, t0 + 0.4, seed=seed)
        err, J = O.imu_edge_inertial(P, s)
        Jn = _numeric_jacobian_inertial(P, s)
        exact = [c for c in range(24) if not 9 <= c < 15]
        assert np.abs(J[:, exact] - Jn[:, exact]).max() < 2e-6, np.abs(J[:, exact] - Jn[:, exact]).max()
        # bias columns: the preintegrated terms are float32 and first order in the bias change -> looser
        assert np.abs(J[:, 9:15] - Jn[:, 9:15]).max() < 5e-3 * max(1.0, np.abs(J[:, 9:15]).max()), np.abs(J[:, 9:15] - Jn[:, 9:15]).max()


def test_edge_mono_jacobians_equal_numeric_derivatives():
    rng = np.random.default_rng(3)
    for _ in range(5):
        Rwb = O.so3('exp', rng.normal(0, 0.4, 3)); twb = rng.normal(0, 1, 3)
        Rbc = O.so3('exp', rng.normal(0, 0.3, 3)); tbc = rng.normal(0, 0.1, 3)
        Rcb = Rbc.T; tcb = -Rcb @ tbc
        cam = np.array([458.0, 457.0, 367.0, 248.0], np.float32)
        Xc = np.array([rng.uniform(-1, 1), rng.uniform(-0.7, 0.7), rng.uniform(2, 8)])
        Xw = Rwb @ (Rbc @ Xc + tbc) + twb
        obs = np.array([cam[0] * Xc[0] / Xc[2] + cam[2], cam[1] * Xc[1] / Xc[2] + cam[3]]) + rng.normal(0, 1.0, 2)
        err, Jp, Jx, dp = O.imu_edge_mono(Rwb, twb, Rcb, tcb, Rbc, tbc, cam, Xw, obs)
        assert dp and np.abs(err).max() < 6
        h = 1e-6
        for k in range(3):
            d = np.zeros(3); d[k] = h
            ep = O.imu_edge_mono(Rwb, twb, Rcb, tcb, Rbc, tbc, cam, Xw + d, obs, jac=False)[0]
            em = O.imu_edge_mono(Rwb, twb, Rcb, tcb, Rbc, tbc, cam, Xw - d, obs, jac=False)[0]
            assert np.allclose((ep - em) / (2 * h), Jp[:, k], rtol=1e-5, atol=1e-5)
        for k in range(6):
            pu = np.zeros(6); pu[k] = h
            Rp, tp = O.imu_pose_update(Rwb, twb, pu)
            Rm, tm = O.imu_pose_update(Rwb, twb, -pu)
            ep = O.imu_edge_mono(Rp, tp, Rcb, tcb, Rbc, tbc, cam, Xw, obs, jac=False)[0]
            em = O.imu_edge_mono(Rm, tm, Rcb, tcb, Rbc, tbc, cam, Xw, obs, jac=False)[0]
            assert np.allclose((ep - em) / (2 * h), Jx[:, k], rtol=1e-5, atol=1e-4)
