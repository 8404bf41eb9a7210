"""GPU parity of the inertial edges against the CPU oracle (oracle/inertial_oracle.cpp, pinned by numerical
Jacobians in tests/test_inertial_cpu.py): IMU preintegration (float; tolerance stated per field: the SVD-based NormalizeRotation and the
float accumulations differ from the CPU in the last bits only through FMA contraction), information matrices, EdgeInertial and EdgeMono
residuals / Jacobians / chi2 (double: 1e-9 relative)."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

pytestmark = pytest.mark.gpu
BIAS = (0.02, -0.01, 0.03, 0.002, -0.001, 0.0015)


@pytest.fixture(scope='module')
def orb():
    import orb_slam3_modified_b200 as m
    m.lib()
    return m


def _intervals(n, maxm=48):
    acc = np.zeros((n, maxm, 3), np.float32); gyr = np.zeros((n, maxm, 3), np.float32); dt = np.zeros((n, maxm), np.float32); nm = np.zeros(n, np.int32)
    bias = np.tile(np.array(BIAS, np.float32), (n, 1))
    for e in range(n):
        t0 = 0.3 + 0.37 * e
        a, g, d = synth.imu_interval(t0, t0 + (1 / 30 if e % 3 else 0.1 + 0.01 * (e % 5)), seed=e)
        k = len(d)
        acc[e, :k], gyr[e, :k], dt[e, :k], nm[e] = a, g, d, k
        bias[e] += np.float32(1e-3) * np.float32(e % 4)
    return acc, gyr, dt, nm, bias


def test_preintegration_and_information(orb):
    acc, gyr, dt, nm, bias = _intervals(37)
    P = orb.imu_preintegrate(acc, gyr, dt, nm, bias, synth.IMU_NOISE)
    info, ig, ia = orb.imu_information(P)
    for e in range(len(nm)):
        k = nm[e]
        Po = O.imu_preintegrate(acc[e, :k], gyr[e, :k], dt[e, :k], bias[e], synth.IMU_NOISE)
        assert P[e, 0] == Po[0] and np.array_equal(P[e, 61:67], Po[61:67])
        assert np.abs(P[e, 1:61] - Po[1:61]).max() < 2e-6, (e, np.abs(P[e, 1:61] - Po[1:61]).max())       # dR dV dP and the bias Jacobians
        C, Co = P[e, 67:], Po[67:]
        assert np.abs(C - Co).max() <= 1e-5 * np.abs(Co).max()
        oi, og, oa = O.imu_information(Po)
        assert np.abs(info[e] - oi).max() < 2e-4 * np.abs(oi).max()          # inverse of an ill-conditioned float covariance
        assert np.allclose(ig[e], og, rtol=1e-5) and np.allclose(ia[e], oa, rtol=1e-5)
        gi, _, _ = O.imu_information(P[e])                                   # same covariance in -> same information out
        assert np.abs(info[e] - gi).max() < 1e-9 * np.abs(gi).max()


def test_edge_inertial_residuals_jacobians_chi2(orb):
    acc, gyr, dt, nm, bias = _intervals(24)
    P = np.stack([O.imu_preintegrate(acc[e, :nm[e]], gyr[e, :nm[e]], dt[e, :nm[e]], bias[e], synth.IMU_NOISE) for e in range(len(nm))])
    states, keys = [], ('Rwb1', 'twb1', 'v1', 'bg', 'ba', 'Rwb2', 'twb2', 'v2')
    sdicts = []
    for e in range(len(nm)):
        t0 = 0.3 + 0.37 * e
        s = synth.inertial_edge_state(t0, t0 + float(P[e, 0]), seed=e, perturb=0.0 if e % 3 == 0 else (1.0 if e % 2 else 4.0))   # every third edge at the ground truth: small chi2
        sdicts.append(s)
        states.append(np.concatenate([np.asarray(s[k], np.float64).reshape(-1) for k in keys]))
    states = np.stack(states)
    info = np.stack([O.imu_information(P[e])[0] for e in range(len(nm))])
    delta = float(np.sqrt(16.92))
    out = orb.imu_inertial_edges(P, states, info, delta)
    for e in range(len(nm)):
        err, J = O.imu_edge_inertial(P[e], sdicts[e])
        assert np.abs(out['err'][e] - err).max() < 1e-6, (e, np.abs(out['err'][e] - err).max())       # float delta terms inside
        assert np.abs(out['J'][e] - J).max() < 1e-6 * max(1.0, np.abs(J).max())
        c2 = float(err @ info[e] @ err)
        assert abs(out['chi2'][e] - float(out['err'][e] @ info[e] @ out['err'][e])) <= 1e-9 * max(1.0, c2)
        assert np.isclose(out['rho'][e], 1.0 if out['chi2'][e] <= delta * delta else delta / np.sqrt(out['chi2'][e]), rtol=1e-12)
    assert (out['rho'] < 1).any() and (out['rho'] == 1).any()
    only = orb.imu_inertial_edges(P, states, jac=False)
    assert np.array_equal(only['err'], out['err']) and only['J'] is None


def test_edge_mono_imu(orb):
    rng = np.random.default_rng(5)
    nP, nL = 9, 700
    poses = np.zeros((nP, 12)); cam = np.tile(np.array([458.0, 457.0, 367.0, 248.0], np.float32), (nP, 1))
    Rbc = O.so3('exp', np.array([0.02, -0.01, 1.55])); tbc = np.array([0.05, -0.02, 0.01])
    Rcb = Rbc.T; tcb = -Rcb @ tbc
    extr = np.concatenate([Rcb.reshape(-1), tcb, Rbc.reshape(-1), tbc])
    for i in range(nP):
        R, p, _, _, _ = synth.imu_trajectory(0.4 * i)
        poses[i, :9] = R.reshape(-1); poses[i, 9:] = p
    ep, ek, obs, pts = [], [], [], np.zeros((nL, 3))
    for j in range(nL):
        k0 = int(rng.integers(0, nP))
        R = poses[k0, :9].reshape(3, 3); p = poses[k0, 9:]
        Xc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(2, 9) * (-1 if j % 20 == 0 else 1)])   # every 20th point behind its cameras
        pts[j] = R @ (Rbc @ Xc + tbc) + p
        for k in range(max(0, k0 - 2), min(nP, k0 + 3)):
            Rk = poses[k, :9].reshape(3, 3); pk = poses[k, 9:]
            xc = Rcb @ (Rk.T @ (pts[j] - pk)) + tcb
            ep.append(j); ek.append(k)
            obs.append([458.0 * xc[0] / xc[2] + 367.0 + rng.normal(0, 1.5) + (25 if rng.random() < 0.05 else 0), 457.0 * xc[1] / xc[2] + 248.0 + rng.normal(0, 1.5)])
    ep = np.array(ep, np.int32); ek = np.array(ek, np.int32); obs = np.array(obs)
    isg = (1.0 / (np.float32(1.2) ** rng.integers(0, 8, len(ep)).astype(np.float32)) ** 2).astype(np.float32)
    delta = float(np.float32(np.sqrt(5.991)))
    out = orb.imu_mono_edges(poses, extr, cam, pts, ep, ek, obs, isg, delta)
    for e in range(0, len(ep), 7):
        P = poses[ek[e]]
        err, Jp, Jx, dp = O.imu_edge_mono(P[:9].reshape(3, 3), P[9:], Rcb, tcb, Rbc, tbc, cam[ek[e]], pts[ep[e]], obs[e])
        assert np.allclose(out['err'][e], err, rtol=1e-10, atol=1e-9)
        assert np.allclose(out['Jpoint'][e], Jp, rtol=1e-9, atol=1e-9) and np.allclose(out['Jpose'][e], Jx, rtol=1e-9, atol=1e-8)
        assert bool(out['depth_pos'][e]) == dp
        c2 = float(isg[e]) * float(err @ err)
        assert np.isclose(out['chi2'][e], c2, rtol=1e-9) and np.isclose(out['rho'][e], 1.0 if c2 <= delta * delta else delta / np.sqrt(c2), rtol=1e-9)
    assert (out['depth_pos'] == 0).any() and (out['rho'] < 1).any()
    with pytest.raises(orb.OrbError):
        orb.imu_mono_edges(poses, extr, cam, pts, np.array([nL], np.int32), np.array([0], np.int32), obs[:1], isg[:1])


def test_pose_inertial_optimization_last_keyframe(orb):
    """f1: Optimizer::PoseInertialOptimizationLastKeyFrame for a batch of frames (outliers, few points, rec-init flag) vs the oracle: same outlier
    flags and return value, reprojection residuals within 1e-4 px (states within 1e-6), prior Hessian within 1e-6 relative."""
    from orb_slam3_modified_b200 import synth
    cases = [dict(seed=0, n=300, outlier_frac=0.1), dict(seed=1, n=700, outlier_frac=0.2), dict(seed=2, n=40, outlier_frac=0.3), dict(seed=3, n=5, outlier_frac=0.0),
             dict(seed=4, n=200, outlier_frac=0.0, perturb=0.0, noise_px=0.0), dict(seed=5, n=1000, outlier_frac=0.05, perturb=2.0)]
    prs = [synth.pose_inertial_problem(**c) for c in cases]
    for rec in (False, True):
        frames = []
        for pr in prs:
            P = O.imu_preintegrate(pr['acc'], pr['gyr'], pr['dt'], pr['bias6'], synth.IMU_NOISE)
            frames.append(dict(pr, preint=P))
        got = orb.PoseInertialOptimizationLastKeyFrame(frames, prs[0]['extr'], rec_init=rec)
        for pr, fr, g in zip(prs, frames, got):
            want = O.pose_inertial_opt_last_kf(pr, fr['preint'], rec_init=rec)
            assert np.array_equal(g['outlier'], want['outlier']) and g['ret'] == want['ret'], (len(pr['Xw']), int((g['outlier'] != want['outlier']).sum()))
            # the device inverts the ill-conditioned float covariance of the preintegration with its own Jacobi sweeps (information matrices equal
            # to 1e-9 relative, test above): states agree to ~1e-8, i.e. the reprojection residuals to far below the 1e-4 px of the north star
            assert np.abs(g['state'] - want['state']).max() < 1e-6
            ex = pr['extr']
            res = lambda st: np.array([O.imu_edge_mono(st[:9].reshape(3, 3), st[9:12], ex[:9], ex[9:12], ex[12:21], ex[21:24], pr['cam'], pr['Xw'][i].astype(np.float64),
                                                         pr['obs'][i].astype(np.float64), jac=False)[0] for i in range(min(len(pr['Xw']), 60))])
            assert np.abs(res(g['state']) - res(want['state'])).max() < 1e-4          # px
            assert np.abs(g['H'] - want['H']).max() <= 1e-6 * np.abs(want['H']).max()


def test_pose_inertial_optimization_last_frame(orb):
    """f1: Optimizer::PoseInertialOptimizationLastFrame (30 unknowns, prior edge, marginalisation) for a batch of frames vs the oracle."""
    from orb_slam3_modified_b200 import synth
    cases = [dict(seed=0, n=300, outlier_frac=0.1), dict(seed=1, n=700, outlier_frac=0.2), dict(seed=2, n=40, outlier_frac=0.3), dict(seed=3, n=4, outlier_frac=0.0),
             dict(seed=5, n=1000, outlier_frac=0.05, perturb=2.0), dict(seed=7, n=250, outlier_frac=0.1, prior_sigma=(2e-2, 5e-2, 1e-1, 1e-3, 1e-2))]
    prs = [synth.pose_inertial_problem_last_frame(**c) for c in cases]
    frames = []
    for pr in prs:
        Pf = O.imu_preintegrate(pr['acc'], pr['gyr'], pr['dt'], pr['bias6'], synth.IMU_NOISE)
        Pk = O.imu_preintegrate(pr['acc_kf'], pr['gyr_kf'], pr['dt_kf'], pr['bias6'], synth.IMU_NOISE)
        frames.append(dict(pr, preint_frame=Pf, preint_kf=Pk))
    for rec in (False, True):
        got = orb.PoseInertialOptimizationLastFrame(frames, prs[0]['extr'], rec_init=rec)
        for pr, fr, g in zip(prs, frames, got):
            want = O.pose_inertial_opt_last_frame(pr, fr['preint_frame'], fr['preint_kf'], rec_init=rec)
            assert np.array_equal(g['outlier'], want['outlier']) and g['ret'] == want['ret'], (len(pr['Xw']), int((g['outlier'] != want['outlier']).sum()))
            assert np.abs(g['state'] - want['state']).max() < 1e-6 and np.abs(g['prev_state'] - want['prev_state']).max() < 1e-6
            ex = pr['extr']
            res = lambda st: np.array([O.imu_edge_mono(st[:9].reshape(3, 3), st[9:12], ex[:9], ex[9:12], ex[12:21], ex[21:24], pr['cam'], pr['Xw'][i].astype(np.float64),
                                                         pr['obs'][i].astype(np.float64), jac=False)[0] for i in range(min(len(pr['Xw']), 60))])
            assert np.abs(res(g['state']) - res(want['state'])).max() < 1e-4          # px
            assert np.abs(g['H'] - want['H']).max() <= 1e-5 * np.abs(want['H']).max()
