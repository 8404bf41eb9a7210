"""The Levenberg-Marquardt control flow of the LBA a second time: `SparseOptimizer::optimize`
(g2o/core/sparse_optimizer.cpp:354-418) around `OptimizationAlgorithmLevenberg::solve` (optimization_algorithm_levenberg.cpp:61-168:
lambda init, trial loop with push / solve / update / rho test / pop, lambda schedule, max 10 trials, Raul's stop rule) transcribed in
Python over a DENSE numpy system with NUMERICAL Jacobians -- against the C++ oracle (Schur complement, analytic Jacobians): same number
of iterations and LM trials, same final lambda and chi2, same final estimate."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth
from test_lba_dense_cpu import _R, _exp


def _residuals(p, Rs, ts, X):
    ic, ip = p['edge_pose'], p['edge_point']
    R = np.stack(Rs)[ic]; t = np.stack(ts)[ic]
    xc = np.einsum('eij,ej->ei', R, X[ip]) + t
    cam = p['cam'].astype(np.float64)[ic]
    return p['obs'] - np.stack([cam[:, 0] * xc[:, 0] / xc[:, 2] + cam[:, 2], cam[:, 1] * xc[:, 1] / xc[:, 2] + cam[:, 3]], 1)


def _robust(chi2, delta):
    big = chi2 > delta * delta
    s = np.sqrt(np.maximum(chi2, 1e-300))
    return np.where(big, 2 * s * delta - delta * delta, chi2), np.where(big, delta / s, 1.0)


def _lm(p, iterations=10):
    nP, nL = len(p['poses']), len(p['points'])
    q = p['poses'][:, :4] / np.linalg.norm(p['poses'][:, :4], axis=1)[:, None]
    Rs = [_R(q[i]) for i in range(nP)]; ts = [p['poses'][i, 4:].astype(np.float64) for i in range(nP)]
    X = p['points'].astype(np.float64).copy()
    free = [i for i in range(nP) if not p['fixed'][i]]
    col = {i: 6 * k for k, i in enumerate(free)}
    n = 6 * len(free) + 3 * nL
    is2 = p['inv_sigma2'].astype(np.float64); delta = float(p['huber_delta'])

    def chi(Rs_, ts_, X_):
        e = _residuals(p, Rs_, ts_, X_)
        return e, _robust(is2 * (e ** 2).sum(1), delta)

    def apply(dx):
        Rn, tn = list(Rs), list(ts)
        for i in free:
            Rd, td = _exp(dx[col[i]:col[i] + 6])
            Rn[i] = Rd @ Rs[i]; tn[i] = Rd @ ts[i] + td
        return Rn, tn, X + dx[6 * len(free):].reshape(nL, 3)

    lam = ni = None
    nBad = cj = trials = 0
    currentChi = 0.0
    ok = True
    it = 0
    while it < iterations and ok:
        e0, (rho0, rho1) = chi(Rs, ts, X)                       # computeActiveErrors + activeRobustChi2
        currentChi = float(rho0.sum()); iniChi = currentChi
        J = np.zeros((2 * len(e0), n)); h = 1e-6               # buildSystem (numerical Jacobians on the manifold)
        for c in range(n):
            d = np.zeros(n); d[c] = h
            J[:, c] = ((_residuals(p, *apply(d)) - _residuals(p, *apply(-d))) / (2 * h)).reshape(-1)
        w = np.repeat(rho1 * is2, 2)
        H = J.T @ (w[:, None] * J); b = -J.T @ (w * e0.reshape(-1))
        if it == 0:
            lam = 1e-5 * np.abs(np.diag(H)).max(); ni = 2.0; nBad = 0
        qmax = 0
        while True:
            try:
                x = np.linalg.solve(H + lam * np.eye(n), b); ok2 = True
            except np.linalg.LinAlgError:
                x = np.zeros(n); ok2 = False
            Rn, tn, Xn = apply(x)
            _, (r0n, _) = chi(Rn, tn, Xn)
            tempChi = float(r0n.sum()) if ok2 else np.finfo(float).max
            rho = (currentChi - tempChi) / (float((x * (lam * x + b)).sum()) + 1e-3)
            if rho > 0 and np.isfinite(tempChi):
                alpha = min(1. - (2 * rho - 1) ** 3, 2. / 3.)
                lam *= max(1. / 3., alpha); ni = 2.0
                currentChi = tempChi
                Rs, ts, X = Rn, tn, Xn                          # discardTop
            else:
                lam *= ni; ni *= 2                              # pop
            qmax += 1; trials += 1
            if not (rho < 0 and qmax < 10):
                break
        cj += 1
        if qmax == 10 or rho == 0:
            ok = False
        else:
            nBad = nBad + 1 if (iniChi - currentChi) * 1e3 < iniChi else 0
            if nBad >= 3:
                ok = False
        it += 1
    return dict(iters=cj, trials=trials, lam=lam, chi2=currentChi, Rs=Rs, ts=ts, X=X)


@pytest.mark.parametrize('kw,shape', [(dict(seed=3, pose_noise=(0.01, 0.02), point_noise=0.02), 'plain'),
                                      (dict(seed=4, pose_noise=(0.6, 15.0), point_noise=1.0), 'rejected'),    # rejected trials: lambda escalation + pop
                                      (dict(seed=7, pose_noise=(0.6, 15.0), point_noise=1.0), 'rejected'),
                                      (dict(seed=2, pose_noise=(0.3, 8.0), point_noise=0.5), 'early'),         # Raul's stop rule ends it before 10
                                      (dict(seed=0, pose_noise=(0.3, 8.0), point_noise=0.5), 'early')])
def test_lm_control_flow_second_transcription(kw, shape):
    p = synth.lba_problem(n_kf=4, n_pts=40, obs_per_pt=3, n_fixed=1, **kw)
    ref = O.lba_solve(p)
    got = _lm(p)
    assert got['iters'] == ref['iters'] and got['trials'] == int(ref['stats'][3]), (got['iters'], got['trials'], ref['iters'], ref['stats'][3])
    assert {'plain': got['trials'] == got['iters'] == 10, 'rejected': got['trials'] > got['iters'], 'early': got['iters'] < 10}[shape]
    assert abs(got['chi2'] - ref['stats'][1]) <= 1e-6 * max(1.0, ref['stats'][1])
    assert abs(got['lam'] - ref['stats'][0]) <= 1e-4 * ref['stats'][0]
    for i in range(len(p['poses'])):
        qo = ref['poses'][i, :4]
        # central differences (h = 1e-6) carry ~1e-7 relative error per linearisation; ten iterations from a bad start stay within 1e-5
        assert np.abs(_R(qo / np.linalg.norm(qo)) - got['Rs'][i]).max() < 1e-5 and np.abs(ref['poses'][i, 4:] - got['ts'][i]).max() < 1e-5
    assert np.abs(ref['points'] - got['X']).max() < 1e-4
