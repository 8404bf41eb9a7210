"""CPU tests of csrc/exact_math.h compiled for the host (tests/hostcheck.cpp): the very functions the
kernels use, against glibc sincosf, cv::fastAtan2 (via the cv2-pinned oracle), libstdc++ std::sort and the
oracle's std::list quadtree."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def hc():
    so = os.path.join(ROOT, 'tests', 'libhostcheck.so')
    src = os.path.join(ROOT, 'tests', 'hostcheck.cpp')
    hdr = os.path.join(ROOT, 'orb_slam3_modified_b200', 'csrc', 'exact_math.h')
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-ffp-contract=off', '-mfma', '-shared', '-o', so, src])
    L = C.CDLL(so)
    L.hc_sincosf_sweep.restype = C.c_long
    L.hc_sincosf_sweep.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f2u(x):
    return struct.unpack('<I', struct.pack('<f', x))[0]


def test_sincosf_sampled(hc):
    """Every 64th float of [0, 7.0] (17 M values) against the host glibc."""
    hi = _f2u(7.0)
    bad = 0
    for lo in range(0, hi, 1 << 22):
        bad += hc.hc_sincosf_sweep(lo, min(lo + (1 << 16), hi), None)
    assert bad == 0


@pytest.mark.slow
def test_sincosf_exhaustive(hc):
    """All 1.09e9 floats in [0, 7.0] (covers angle*pi/180 for angle in [0, 360])."""
    from concurrent.futures import ThreadPoolExecutor
    hi = _f2u(7.0)
    chunks = [(lo, min(lo + (1 << 24), hi)) for lo in range(0, hi, 1 << 24)]
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        bad = sum(ex.map(lambda c: hc.hc_sincosf_sweep(c[0], c[1], None), chunks))
    assert bad == 0


def test_fast_atan2_vs_oracle(hc):
    rng = np.random.default_rng(3)
    y = rng.integers(-2900000, 2900000, 300000).astype(np.float32)
    x = rng.integers(-2900000, 2900000, 300000).astype(np.float32)
    y[:50] = 0
    x[25:75] = 0
    out = np.zeros_like(y)
    hc.hc_fast_atan2_n(_p(y), _p(x), _p(out), len(y))
    assert np.array_equal(out, O.fast_atan2(y, x))


def test_sort_emulation_vs_libstdcxx(hc):
    rng = np.random.default_rng(0)
    bad = 0
    for it in range(1500):
        n = int(rng.integers(2, 1300))
        size = rng.integers(2, 2 + int(rng.integers(1, 12)), n).astype(np.int32)
        ulx = (rng.integers(0, 1 + int(rng.integers(1, 40)), n) * 7).astype(np.int32)
        bad += hc.hc_sort_check(_p(size), _p(ulx), n, None)
        bad += hc.hc_sort_check_rounds(_p(size), _p(ulx), n)
    for n in (17, 100, 1000, 5000):   # adversarial shapes incl. the heapsort fallback
        for arr in (np.arange(n), np.arange(n)[::-1], np.r_[np.arange(n // 2), np.arange(n - n // 2)[::-1]], np.zeros(n), np.arange(n) % 3):
            size = np.ascontiguousarray(arr, dtype=np.int32)
            ulx = np.ascontiguousarray((np.arange(n) * 7919) % 13, dtype=np.int32)
            bad += hc.hc_sort_check(_p(size), _p(ulx), n, None)
            bad += hc.hc_sort_check_rounds(_p(size), _p(ulx), n)
    assert bad == 0


def test_quadtree_formulation_vs_oracle(hc):
    """The list-rebuild formulation of DistributeOctTree used by the kernel == the oracle's std::list version."""
    e = O.OracleExtractor()
    rng = np.random.default_rng(1)
    cases = 0
    for it in range(250):
        W = int(rng.integers(40, 1300))
        H = int(rng.integers(max(30, W // 3), max(31, min(W * 2 - 1, 900))))
        if round(np.float32(W) / np.float32(H)) < 1:
            continue
        n, N = int(rng.integers(1, 5000)), int(rng.integers(1, 1200))
        if it % 3 == 0:
            cx, cy = rng.integers(3, W, 5), rng.integers(3, H, 5)
            xs = np.clip(cx[rng.integers(0, 5, n)] + rng.integers(-12, 13, n), 3, W - 1)
            ys = np.clip(cy[rng.integers(0, 5, n)] + rng.integers(-12, 13, n), 3, H - 1)
            u = np.unique(np.stack([ys, xs], 1), axis=0)
            ys, xs = u[:, 0], u[:, 1]
        else:
            pos = np.sort(rng.choice((W - 3) * (H - 3), size=min(n, (W - 3) * (H - 3)), replace=False))
            ys, xs = pos // (W - 3) + 3, pos % (W - 3) + 3
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        rs = np.ascontiguousarray(rng.integers(7, 60, len(xs)), np.int32)
        cands = np.zeros(len(xs), O.KP_DTYPE)
        cands['x'], cands['y'], cands['response'] = xs, ys, rs
        ref = e.distribute(cands, 16, 16 + W, 16, 16 + H, N)
        out = np.zeros(len(xs) + 8, np.int32)
        m = hc.hc_quadtree(_p(xs), _p(ys), _p(rs), len(xs), W, H, N, _p(out))
        assert m == len(ref), (W, H, len(xs), N)
        idx = out[:m]
        assert np.array_equal(xs[idx], ref['x']) and np.array_equal(ys[idx], ref['y']) and np.array_equal(rs[idx], ref['response'])
        cases += 1
    assert cases > 150
