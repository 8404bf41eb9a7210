"""CPU: the inertial oracle's per-edge numerics equal the REFERENCE's own g2o types.  oracle/_ref/libref_inertial.so holds the bodies of
ImuCamPose::Project / isDepthPositive / Update, EdgeMono / EdgeMonoOnlyPose::linearizeOplus, EdgeInertial::computeError / linearizeOplus,
EdgePriorPoseImu::computeError / linearizeOplus, ExpSO3 / LogSO3 / RightJacobianSO3 / InverseRightJacobianSO3 / Skew (src/G2oTypes.cc) and
Pinhole::project / projectJac (src/CameraModels/Pinhole.cpp), cut out of /root/reference at build time and compiled verbatim against a small
fixed-size matrix type (oracle/ref_shim/mini_eigen.hpp; Eigen itself is not installed).  That type evaluates eagerly, so the agreement pinned here
is the reference's formulas to rounding (1e-12), not Eigen's last bit; the float preintegrated terms (Sophus::SO3f::exp, JacobiSVD) and
NormalizeRotation are the oracle's on both sides."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'oracle', '_ref', 'libref_inertial.so')
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason='oracle/_ref is not built here')


def _lib():
    O.lib()
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def test_so3_functions():
    L = _lib()
    rng = np.random.default_rng(0)
    vs = [rng.normal(0, s, 3) for s in (1e-7, 1e-4, 0.01, 0.5, 2.0, 3.0) for _ in range(4)] + [np.array([np.pi - 1e-4, 0, 0]), np.zeros(3)]
    for v in vs:
        v = np.ascontiguousarray(v)
        for what in (0, 2, 3):
            out = np.zeros((3, 3))
            L.ref_so3(what, _p(v), _p(out))
            assert _close(out, O.so3({0: 'exp', 2: 'Jr', 3: 'invJr'}[what], v)), (what, v)
        R = O.so3('exp', v)
        w = np.zeros(3)
        L.ref_so3(1, _p(np.ascontiguousarray(R)), _p(w))
        assert _close(w, O.so3('log', R))


def test_edge_inertial_error_and_jacobians():
    L = _lib()
    for seed in range(8):
        t0 = 0.4 + 0.7 * seed
        acc, gyr, dts = synth.imu_interval(t0, t0 + (0.05 if seed % 2 else 0.4), seed=seed)
        P = O.imu_preintegrate(acc, gyr, dts, (0.02, -0.01, 0.03, 0.002, -0.001, 0.0015), synth.IMU_NOISE)
        s = synth.inertial_edge_state(t0, t0 + 0.4, seed=seed, perturb=1.0 + seed)
        want_e, want_J = O.imu_edge_inertial(P, s)
        e = np.zeros(9); J = np.zeros((9, 24))
        a = [np.ascontiguousarray(s[k], np.float64) for k in ('Rwb1', 'twb1', 'v1', 'bg', 'ba', 'Rwb2', 'twb2', 'v2')]
        L.ref_edge_inertial(_p(P), *[_p(x) for x in a], _p(e), _p(J))
        assert _close(e, want_e) and _close(J, want_J), (seed, np.abs(e - want_e).max(), np.abs(J - want_J).max())


def test_edge_mono_and_pose_update():
    L = _lib()
    rng = np.random.default_rng(1)
    ex = np.ascontiguousarray(synth.imu_extrinsics())
    cam = synth.camera()
    for k in range(20):
        R, p, _, _, _ = synth.imu_trajectory(0.3 * k)
        R = np.ascontiguousarray(R); p = np.ascontiguousarray(p)
        Rcw = ex[:9].reshape(3, 3) @ R.T; tcw = ex[:9].reshape(3, 3) @ (-R.T @ p) + ex[9:12]
        Xc = np.array([rng.uniform(-2, 2), rng.uniform(-1.5, 1.5), rng.uniform(0.5, 20) * (1 if k % 5 else -1)])
        Xw = np.ascontiguousarray(Rcw.T @ (Xc - tcw)); obs = np.ascontiguousarray(rng.uniform(0, 600, 2))
        we, wJp, wJx, wd = O.imu_edge_mono(R, p, ex[:9], ex[9:12], ex[12:21], ex[21:24], cam, Xw, obs)
        e = np.zeros(2); Jp = np.zeros((2, 3)); Jx = np.zeros((2, 6)); Jo = np.zeros((2, 6)); d = C.c_int(0)
        L.ref_edge_mono(_p(R), _p(p), _p(ex), _p(np.ascontiguousarray(cam, np.float32)), _p(Xw), _p(obs), _p(e), _p(Jp), _p(Jx), _p(Jo), C.byref(d))
        assert _close(e, we) and _close(Jp, wJp) and _close(Jx, wJx) and _close(Jo, wJx) and bool(d.value) == wd
        # ImuCamPose::Update: one update, and five in a row (the reference never renormalises Rwb: its NormalizeRotation(Rwb) call discards the result)
        pu = np.ascontiguousarray(rng.normal(0, 0.02, 6))
        for times in (1, 5):
            Rr, tr = R.copy(), p.copy(); Rc = np.zeros((3, 3)); tc = np.zeros(3)
            L.ref_pose_update(_p(Rr), _p(tr), _p(ex), _p(pu), times, _p(Rc), _p(tc))
            Ro, to = R.copy(), p.copy()
            for _ in range(times):
                Ro, to = O.imu_pose_update(Ro, to, pu)
            assert _close(Rr, Ro) and _close(tr, to)
            assert _close(Rc, ex[:9].reshape(3, 3) @ Ro.T) and _close(tc, ex[:9].reshape(3, 3) @ (-Ro.T @ to) + ex[9:12])


def test_edge_prior_pose_imu():
    """EdgePriorPoseImu against a numpy restatement on the oracle's SO3 functions (the oracle's own prior edge is internal to the last-frame optimiser,
    whose first Gauss-Newton step is pinned by numerical derivatives in test_inertial_cpu.py)."""
    L = _lib()
    rng = np.random.default_rng(2)
    for k in range(6):
        pr = synth.pose_inertial_problem_last_frame(seed=k, n=10)
        prior = np.ascontiguousarray(pr['prior_state']); st = np.ascontiguousarray(pr['truth_prev'] + np.concatenate([np.zeros(9), rng.normal(0, 0.01, 12)]))
        st[:9] = (st[:9].reshape(3, 3) @ synth._rodrigues(rng.normal(0, 0.02, 3))).reshape(9)
        e = np.zeros(15); J = np.zeros((15, 15))
        L.ref_edge_prior(_p(prior), _p(st), _p(e), _p(J))
        Rp, R = prior[:9].reshape(3, 3), st[:9].reshape(3, 3)
        er = O.so3('log', Rp.T @ R)
        want = np.concatenate([er, Rp.T @ (st[9:12] - prior[9:12]), st[12:15] - prior[12:15], st[15:18] - prior[15:18], st[18:21] - prior[18:21]])
        Jw = np.zeros((15, 15)); Jw[:3, :3] = O.so3('invJr', er); Jw[3:6, 3:6] = Rp.T @ R; Jw[6:, 6:] = np.eye(9)
        assert _close(e, want) and _close(J, Jw)


def test_preintegration_equals_the_reference_text():
    """IMU::Preintegrated::Initialize + IntegrateNewMeasurement (src/ImuTypes.cc:147-166, :177-236) and IntegratedRotation (:84-105), compiled verbatim into
    oracle/_ref/libref_preint.so (float; NormalizeRotation is the oracle's polar factor on both sides): every field of the 292-float record -- dT, dR, dV, dP,
    the five bias Jacobians and the 15 x 15 covariance -- of the oracle's preintegration equals it."""
    so = os.path.join(ROOT, 'oracle', '_ref', 'libref_preint.so')
    if not os.path.exists(so):
        pytest.skip('oracle/_ref is not built here')
    O.lib()
    L = C.CDLL(so)
    worst = 0.0
    for seed in range(10):
        t0 = 0.3 + 0.9 * seed
        acc, gyr, dts = synth.imu_interval(t0, t0 + (0.04 if seed % 3 == 0 else 0.25 + 0.05 * seed), seed=seed, noise=seed % 2 == 0)
        if seed == 7:
            gyr[:] = np.float32(1e-7)                    # |w| dt below IMU::eps: the first-order branch of IntegratedRotation
        bias = np.array([0.02, -0.01, 0.03, 0.002, -0.001, 0.0015], np.float32) * np.float32(1 + seed)
        want = O.imu_preintegrate(acc, gyr, dts, bias, synth.IMU_NOISE)
        got = np.zeros(292, np.float32)
        a = [np.ascontiguousarray(x, np.float32) for x in (acc, gyr, dts, bias, np.array(synth.IMU_NOISE, np.float32))]
        L.ref_imu_preintegrate(len(dts), *[_p(x) for x in a], _p(got))
        for lo, hi in ((0, 1), (1, 10), (10, 13), (13, 16), (16, 25), (25, 34), (34, 43), (43, 52), (52, 61), (61, 67), (67, 292)):
            g, w = got[lo:hi].astype(np.float64), want[lo:hi].astype(np.float64)
            scale = max(np.abs(w).max(), 1e-30)
            worst = max(worst, np.abs(g - w).max() / scale)
            assert np.abs(g - w).max() <= 2e-6 * scale, (seed, lo, hi, np.abs(g - w).max(), scale)
    print('worst relative difference per field group: %.2e' % worst)
