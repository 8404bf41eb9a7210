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


def test_converged_solution_is_the_minimum_an_independent_optimizer_finds():
    """The Schur machinery (a17-a19) checked at the level of its RESULT: g2o's Levenberg-Marquardt with Huber kernels is an iteratively re-weighted descent on
    sum_e rho(chi2_e); scipy's trust-region solver minimising that same robust cost directly (residuals scaled by sqrt(rho / chi2), numerical Jacobian, its own
    parametrisation T = exp(d) T0, p = p0 + d; gauge fixed by two fixed keyframes) must end at the same cost and the same reprojection residuals as the oracle."""
    from scipy.optimize import least_squares
    p = synth.lba_problem(n_kf=5, n_pts=40, obs_per_pt=4, n_fixed=2, seed=11, outlier_frac=0.05, pose_noise=(0.005, 0.2), point_noise=0.01)
    nP, nL = len(p['poses']), len(p['points'])
    free = [i for i in range(nP) if not p['fixed'][i]]
    R0 = [_R(q[:4] / np.linalg.norm(q[:4])) for q in p['poses']]; t0 = [q[4:].copy() for q in p['poses']]
    w = np.sqrt(p['inv_sigma2'].astype(np.float64))
    delta = p['huber_delta']

    def unpack(x):
        Rs, ts = list(R0), list(t0)
        for k, i in enumerate(free):
            Re, te = _exp(x[6 * k:6 * k + 6])
            Rs[i] = Re @ R0[i]; ts[i] = Re @ t0[i] + te
        return Rs, ts, p['points'] + x[6 * len(free):].reshape(-1, 3)

    def robust(e):                                                    # e: whitened residuals [nE, 2] -> scaled so that |.|^2 = rho(chi2)
        c = (e ** 2).sum(1)
        f = np.ones(len(c))
        out = c > delta * delta
        f[out] = np.sqrt((2 * delta * np.sqrt(c[out]) - delta * delta) / c[out])
        return e * f[:, None]

    def residuals(x):
        Rs, ts, X = unpack(x)
        return robust(_residuals(p, Rs, ts, X) * w[:, None]).reshape(-1)
    n = 6 * len(free) + 3 * nL
    sol = least_squares(residuals, np.zeros(n), method='trf', jac='2-point', xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=80)
    got = O.lba_solve(p, iterations=60)
    res_o = O.lba_residuals(p, got['poses'], got['points'])
    ro = robust(res_o * w[:, None]).reshape(-1)
    cost_o, cost_s = float(ro @ ro), float(sol.fun @ sol.fun)
    assert ((res_o * w[:, None]) ** 2).sum(1).max() > delta * delta    # the kernel is active at the optimum (gross outliers present)
    assert cost_s <= cost_o * (1 + 1e-9) and abs(cost_o - cost_s) < 1e-4 * cost_s, (cost_o, cost_s)   # measured 1.8e-5: g2o's stop rule (three iterations below 0.1 %) ends just short of the minimum
    Rs, ts, X = unpack(sol.x)
    assert np.abs(res_o - _residuals(p, Rs, ts, X)).max() < 0.25       # px (measured 0.1 on residuals of up to 20 px): two optimisers at the same minimum of the robust cost, one stopped by g2o's rule
