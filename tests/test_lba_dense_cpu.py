"""An independent check of the LBA oracle's linear algebra ("Schur solution = full-system solution"): one
Levenberg-Marquardt step of the oracle (Schur complement + landmark back-substitution, analytic Jacobians) against a dense
full-system step built here in numpy from NUMERICAL Jacobians of the reprojection residual on the manifold
(pose update T <- exp(d) T with d = (omega, upsilon), g2o se3quat.h:223-256; point update p <- p + d)."""
import numpy as np

import oracle_lib as O
from orb_slam3_modified_b200 import synth


def _R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _exp(d):
    om, up = d[:3], d[3:]
    th = np.linalg.norm(om)
    Om = _hat(om)
    if th < 1e-12:
        return np.eye(3) + Om, up.copy()
    Rm = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
    V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    return Rm, V @ up


def _residuals(p, Rs, ts, X):
    cam = p['cam'].astype(np.float64)
    e = np.zeros((len(p['edge_point']), 2))
    for k, (ip, ic) in enumerate(zip(p['edge_point'], p['edge_pose'])):
        xc = Rs[ic] @ X[ip] + ts[ic]
        fx, fy, cx, cy = cam[ic]
        e[k] = p['obs'][k] - np.array([fx * xc[0] / xc[2] + cx, fy * xc[1] / xc[2] + cy])
    return e


def test_one_lm_step_equals_dense_numeric_step():
    p = synth.lba_problem(n_kf=4, n_pts=40, obs_per_pt=3, seed=3, n_fixed=1, pose_noise=(0.01, 0.02), point_noise=0.02)
    ref = O.lba_solve(p, iterations=1)
    assert ref['iters'] == 1 and int(ref['stats'][3]) == 1            # one iteration, one (accepted) trial
    nP, nL = len(p['poses']), len(p['points'])
    q = p['poses'][:, :4] / np.linalg.norm(p['poses'][:, :4], axis=1)[:, None]
    Rs = [_R(q[i]) for i in range(nP)]
    ts = [p['poses'][i, 4:].astype(np.float64) for i in range(nP)]
    X = p['points'].astype(np.float64)
    free = [i for i in range(nP) if not p['fixed'][i]]
    col = {i: 6 * k for k, i in enumerate(free)}
    n = 6 * len(free) + 3 * nL
    e0 = _residuals(p, Rs, ts, X)
    nE = len(e0)
    # numerical Jacobian of the stacked residual, central differences on the manifold
    J = np.zeros((2 * nE, n))
    h = 1e-6
    for i in free:
        for a in range(6):
            d = np.zeros(6); d[a] = h
            out = []
            for s in (+1, -1):
                Rd, td = _exp(s * d)
                Rs2, ts2 = list(Rs), list(ts)
                Rs2[i] = Rd @ Rs[i]; ts2[i] = Rd @ ts[i] + td
                out.append(_residuals(p, Rs2, ts2, X))
            J[:, col[i] + a] = ((out[0] - out[1]) / (2 * h)).reshape(-1)
    for l in range(nL):
        for a in range(3):
            out = []
            for s in (+1, -1):
                X2 = X.copy(); X2[l, a] += s * h
                out.append(_residuals(p, Rs, ts, X2))
            J[:, 6 * len(free) + 3 * l + a] = ((out[0] - out[1]) / (2 * h)).reshape(-1)
    # g2o's robustified normal equations: H = J^T (rho' Omega) J, b = -J^T (rho' Omega) e   (base_binary_edge.hpp:55-120)
    is2 = p['inv_sigma2'].astype(np.float64)
    chi2 = is2 * (e0 ** 2).sum(1)
    delta = float(p['huber_delta'])
    rho1 = np.where(chi2 <= delta * delta, 1.0, delta / np.sqrt(np.maximum(chi2, 1e-300)))
    w = np.repeat(rho1 * is2, 2)
    H = J.T @ (w[:, None] * J)
    b = -J.T @ (w * e0.reshape(-1))
    lam = 1e-5 * np.abs(np.diag(H)).max()                             # computeLambdaInit
    dx = np.linalg.solve(H + lam * np.eye(n), b)
    # apply and compare with the oracle's state after its single accepted step
    for i in free:
        Rd, td = _exp(dx[col[i]:col[i] + 6])
        Rn, tn = Rd @ Rs[i], Rd @ ts[i] + td
        qo = ref['poses'][i, :4]
        assert np.abs(_R(qo / np.linalg.norm(qo)) - Rn).max() < 2e-7
        assert np.abs(ref['poses'][i, 4:] - tn).max() < 2e-7
    Xn = X + dx[6 * len(free):].reshape(nL, 3)
    assert np.abs(ref['points'] - Xn).max() < 2e-6
    # and the step is a real one (so the comparison is not vacuous)
    assert np.abs(ref['points'] - X).max() > 1e-3
