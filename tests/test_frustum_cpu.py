"""CPU: the oracle's Frame::isInFrustum behaves like the reference describes it, and logf_glibc (the device's std::log(float))
equals the host libm bit for bit."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import frustum_scenes
import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _fnmadd(a, b, c):
    """float32 fma(-a, b, c) element-wise through the host libm's fmaf (one rounding, like the FMA the reference build emits)."""
    libm = C.CDLL('libm.so.6'); libm.fmaf.restype = C.c_float; libm.fmaf.argtypes = [C.c_float] * 3
    with np.errstate(all='ignore'):
        return np.array([libm.fmaf(-float(a), float(bi), float(ci)) for bi, ci in zip(b, c)], np.float32)


def _bits(f):
    return struct.unpack('<I', struct.pack('<f', f))[0]


def _hc():
    L = C.CDLL(os.path.join(HERE, 'libhostcheck.so'))
    L.hc_logf_sweep.restype = C.c_long
    L.hc_logf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    return L


def test_logf_glibc_matches_libm_on_the_ratio_range():
    # MapPoint::PredictScale passes mfMaxDistance / dist: every float in [2^-6, 2^10] (134 M values)
    bad = C.c_uint32(0)
    assert _hc().hc_logf_sweep(_bits(2.0 ** -6), _bits(2.0 ** 10), C.byref(bad)) == 0, hex(bad.value)


@pytest.mark.slow
def test_logf_glibc_matches_libm_all_positive_normals():
    bad = C.c_uint32(0)
    assert _hc().hc_logf_sweep(0x00800000, 0x7f800000, C.byref(bad)) == 0, hex(bad.value)


def test_oracle_is_in_frustum_fields():
    sc = frustum_scenes.scene(4000, seed=3)
    out = O.is_in_frustum(**sc)
    P = sc['pts']['worldPos'].astype(np.float64)
    pc = P @ sc['Rcw'].astype(np.float64).T + sc['tcw'].astype(np.float64)
    behind = pc[:, 2] < 0
    assert 0.15 < out['inView'].mean() < 0.8
    assert not out['inView'][behind].any() and np.all(out['projX'][behind] == -1) and np.all(out['projY'][behind] == -1)
    iv = out['inView'] == 1
    u = sc['cam'][0] * pc[:, 0] / pc[:, 2] + sc['cam'][2]
    assert np.allclose(out['projX'][iv], u[iv], rtol=1e-5, atol=1e-3)
    assert np.all((out['projX'][iv] >= 0) & (out['projX'][iv] <= 640) & (out['projY'][iv] >= 0) & (out['projY'][iv] <= 480))
    assert np.all(out['viewCos'][iv] >= 0.5) and np.all((out['level'][iv] >= 0) & (out['level'][iv] < 8)) and np.all(out['level'][~iv] == -1)
    assert np.allclose(out['depth'][iv], np.linalg.norm(pc[iv], axis=1), rtol=1e-5)
    assert np.allclose(out['projXR'][iv], out['projX'][iv] - 40.0 / pc[iv, 2], rtol=1e-5, atol=1e-3)
    # the bounds test sets the projection even when a later test fails (src/Frame.cc:539-540)
    later = (~iv) & (out['projX'] != -1)
    assert later.any()
    dist = np.linalg.norm(P - sc['Ow'].astype(np.float64), axis=1)
    lvl = np.clip(np.ceil(np.log(sc['pts']['maxDistance'] / dist) / np.log(1.2)), 0, 7)
    assert (out['level'][iv] == lvl[iv]).mean() > 0.999      # float vs double log: equal except on exact level boundaries


def test_oracle_is_in_frustum_equals_float32_numpy_restatement():
    """Frame::isInFrustum + PredictScale once more, vectorised in numpy float32 (every operation rounds to float like the C++ float
    expressions; logf from the host libm): all seven output fields equal the oracle's bit for bit."""
    f32 = np.float32
    libm = C.CDLL('libm.so.6'); libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    for seed in (4, 8):
        sc = frustum_scenes.scene(5000, seed=seed)
        out = O.is_in_frustum(**sc)
        P, Nn = sc['pts']['worldPos'], sc['pts']['normal']
        R, t, Ow = sc['Rcw'].astype(f32), sc['tcw'].astype(f32), sc['Ow'].astype(f32)
        fx, fy, cx, cy = (f32(v) for v in sc['cam'])
        X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
        xc = ((R[0, 0] * X + R[0, 1] * Y) + R[0, 2] * Z) + t[0]
        yc = ((R[1, 0] * X + R[1, 1] * Y) + R[1, 2] * Z) + t[1]
        zc = ((R[2, 0] * X + R[2, 1] * Y) + R[2, 2] * Z) + t[2]
        pc = np.sqrt((xc * xc + yc * yc) + zc * zc)
        with np.errstate(divide='ignore', invalid='ignore'):
            invz = f32(1) / zc
            u, v = fx * xc / zc + cx, fy * yc / zc + cy
        a = ~(zc < 0)
        b = a & ~((u < f32(0)) | (u > f32(640))) & ~((v < f32(0)) | (v > f32(480)))
        ox, oy, oz = X - Ow[0], Y - Ow[1], Z - Ow[2]
        dist = np.sqrt((ox * ox + oy * oy) + oz * oz)
        c = b & ~((dist < sc['pts']['minDistInv']) | (dist > sc['pts']['maxDistInv']))
        with np.errstate(divide='ignore', invalid='ignore'):
            vc = ((ox * Nn[:, 0] + oy * Nn[:, 1]) + oz * Nn[:, 2]) / dist
            ratio = sc['pts']['maxDistance'] / dist
        d = c & ~(vc < f32(0.5))
        lvl = np.full(len(P), -1, np.int32)
        for i in np.flatnonzero(d):
            n = int(np.ceil(f32(libm.logf(float(ratio[i]))) / f32(sc['log_scale_factor'])))
            lvl[i] = min(max(n, 0), 7)
        assert np.array_equal(out['inView'], d.astype(np.uint8))
        assert np.array_equal(out['projX'], np.where(b, u, f32(-1))) and np.array_equal(out['projY'], np.where(b, v, f32(-1)))
        assert np.array_equal(out['projXR'], np.where(d, _fnmadd(f32(sc['mbf']), invz, u), f32(0))) and np.array_equal(out['depth'], np.where(d, pc, f32(0)))
        assert np.array_equal(out['viewCos'], np.where(d, vc, f32(0))) and np.array_equal(out['level'], lvl)
