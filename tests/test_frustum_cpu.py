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
