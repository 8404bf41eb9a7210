"""CPU: the per-edge numerics of the LocalBundleAdjustment / PoseOptimization oracle equal the REFERENCE's own text: EdgeSE3ProjectXYZ::linearizeOplus and
EdgeSE3ProjectXYZOnlyPose::linearizeOplus (src/OptimizableTypes.cpp:139-160, :49-63), Pinhole::project / projectJac, and g2o's SE3Quat::exp / operator* /
map / normalizeRotation (Thirdparty/g2o/g2o/types/se3quat.h) behind VertexSE3Expmap::oplusImpl -- compiled verbatim against oracle/ref_shim/mini_eigen.hpp
into oracle/_ref/libref_inertial.so (Eigen::Quaternion's own algorithms are that header's stand-in) -- to 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'oracle', '_ref', 'libref_inertial.so')
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason='oracle/_ref is not built here')


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


def _edge(L, name, pose, cam, X, obs):
    e = np.zeros(2); Jp = np.zeros((2, 3)); Jx = np.zeros((2, 6)); Jo = np.zeros((2, 6)); d = C.c_int(0)
    getattr(L, name)(_p(pose), _p(cam), _p(X), _p(obs), _p(e), _p(Jp), _p(Jx), _p(Jo), C.byref(d))
    return e, Jp, Jx, Jo, d.value


def test_edge_errors_and_jacobians():
    Lo = O.lib(); Lr = C.CDLL(SO)
    rng = np.random.default_rng(0)
    cam = np.array([458.654, 457.296, 367.215, 248.375], np.float32)
    for k in range(60):
        q = rng.normal(0, 1, 4); q /= np.linalg.norm(q)
        if k % 3 == 0:
            q = -np.abs(q)                                         # w < 0: the SE3Quat constructor flips the sign
        pose = np.ascontiguousarray(np.concatenate([q, rng.normal(0, 2, 3)]))
        X = np.ascontiguousarray(rng.normal(0, 5, 3)); obs = np.ascontiguousarray(rng.uniform(0, 700, 2))
        got = _edge(Lo, 'orbo_lba_edge', pose, cam, X, obs)
        want = _edge(Lr, 'ref_lba_edge', pose, cam, X, obs)
        for a, b in zip(got[:4], want[:4]):
            assert _close(a, b), (k, np.abs(a - b).max())
        assert got[4] == want[4]


def test_pose_oplus():
    Lo = O.lib(); Lr = C.CDLL(SO)
    rng = np.random.default_rng(1)
    for k in range(60):
        q = rng.normal(0, 1, 4); q /= np.linalg.norm(q)
        pose = np.concatenate([q, rng.normal(0, 2, 3)])
        scale = (1e-7, 1e-4, 0.05, 1.0, 3.0)[k % 5]                # both branches of SE3Quat::exp (theta < 1e-5) and large rotations
        upd = np.ascontiguousarray(np.concatenate([rng.normal(0, scale, 3), rng.normal(0, 0.3, 3)]))
        a, b = np.ascontiguousarray(pose.copy()), np.ascontiguousarray(pose.copy())
        Lo.orbo_lba_pose_oplus(_p(a), _p(upd)); Lr.ref_lba_pose_oplus(_p(b), _p(upd))
        assert _close(a, b), (k, np.abs(a - b).max())
        assert abs(np.linalg.norm(a[:4]) - 1) < 1e-14 and a[0] >= 0


def test_huber_kernel():
    Lo = O.lib(); Lr = C.CDLL(SO)
    for f in (Lo.orbo_huber, Lr.ref_huber):
        f.argtypes = [C.c_double, C.c_double, C.c_void_p]
    for delta in (float(np.float32(np.sqrt(5.991))), np.sqrt(16.92), 5.0):
        for e in (0.0, 1e-9, delta * delta, delta * delta * (1 + 1e-15), 7.3, 1e4):
            a, b = np.zeros(3), np.zeros(3)
            Lo.orbo_huber(delta, e, _p(a)); Lr.ref_huber(delta, e, _p(b))
            # (the reference's flags fuse 2*sqrt(e)*delta - dsqr into one FMA, the oracle is built with -ffp-contract=off: the last bit may differ)
            assert np.allclose(a, b, rtol=4e-16, atol=1e-15), (delta, e, a, b)
