"""Independent pin of computeOrbDescriptor.  ORB-SLAM3's steered BRIEF is OpenCV's (same 256-pair pattern, same
rotation formula, same 7x7 sigma-2 blur), so for level-0 keypoints `cv2.ORB.compute` fed with the oracle's keypoints (position + angle)
is an independent implementation of the same descriptor.  Checked here:
  * a numpy restatement (pattern table from csrc/brief_pattern.inc, glibc sincosf, cv2.GaussianBlur plane, round-half-even) equals the
    oracle bit for bit, with and without the FMA contraction of the reference build;
  * cv2.ORB agrees on > 99.8 % of the bits, and EVERY bit it disagrees on compares two samples that differ by at most one grey level in
    the blurred plane (cv2's ORB-internal blur is not bit-identical to a stand-alone cv2.GaussianBlur, which is what the reference
    calls on a clone of the level, src/ORBextractor.cc:1132-1133, and what the oracle is pinned to)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

cv2 = pytest.importorskip('cv2')
HERE = os.path.dirname(os.path.abspath(__file__))


def _pattern():
    txt = open(os.path.join(HERE, '..', 'orb_slam3_modified_b200', 'csrc', 'brief_pattern.inc')).read()
    v = np.array([int(t) for t in re.findall(r'-?\d+', re.sub(r'//.*', '', txt))], np.int32)
    assert len(v) == 1024
    return v.reshape(256, 2, 2)      # pair, point, (x, y)


def _sincosf(angles):
    libm = C.CDLL('libm.so.6')
    libm.sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    s, c = C.c_float(), C.c_float()
    out = []
    for a in angles:
        libm.sincosf(C.c_float(a), C.byref(s), C.byref(c))
        out.append((np.float32(s.value), np.float32(c.value)))
    return out


def _samples(blur, x, y, a, b, pair):
    cx, cy = int(np.rint(x)), int(np.rint(y))
    out = []
    for pt in range(2):
        px, py = np.float32(pair[pt, 0]), np.float32(pair[pt, 1])
        r = np.float32(np.float64(px) * np.float64(b) + np.float64(py * a))
        q = np.float32(np.float64(px) * np.float64(a) - np.float64(py * b))
        out.append(blur[cy + int(np.rint(r)), cx + int(np.rint(q))])
    return out


def _describe(blur, x, y, a, b, pat, fma):
    cx, cy = int(np.rint(x)), int(np.rint(y))
    px, py = pat[:, :, 0].astype(np.float32), pat[:, :, 1].astype(np.float32)
    if fma:   # fma(x, b, y*a), fma(x, a, -(y*b)): exact product + rounded product, one rounding
        r = (px.astype(np.float64) * np.float64(b) + np.float64(py * a)).astype(np.float32)
        q = (px.astype(np.float64) * np.float64(a) - np.float64(py * b)).astype(np.float32)
    else:
        r = px * b + py * a
        q = px * a - py * b
    ri, qi = np.rint(r).astype(np.int64), np.rint(q).astype(np.int64)      # cvRound: half to even
    v = blur[cy + ri, cx + qi]
    bits = (v[:, 0] < v[:, 1]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8)[:, ::-1], axis=1).reshape(32)


@pytest.mark.parametrize('t', [3, 11])
def test_level0_descriptors_against_cv2_orb(t):
    img = synth.frame(t)
    _, kps, desc = O.OracleExtractor(1000, 1.2, 8, 20, 7)(img, (0, 1000))
    sel = [i for i in np.flatnonzero(kps['octave'] == 0) if 31 <= kps['x'][i] < 640 - 31 and 31 <= kps['y'][i] < 480 - 31]
    assert len(sel) > 100
    cvk = [cv2.KeyPoint(float(kps['x'][i]), float(kps['y'][i]), 31.0, float(kps['angle'][i]), float(kps['response'][i]), 0, -1) for i in sel]
    k2, d2 = cv2.ORB_create(nfeatures=5000, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2, patchSize=31).compute(img, cvk)
    assert len(k2) == len(sel) and all(abs(k.pt[0] - kps['x'][i]) < 1e-4 and abs(k.pt[1] - kps['y'][i]) < 1e-4 for k, i in zip(k2, sel))
    blur = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    pat = _pattern()
    factor = np.float32(3.14159265358979323846 / np.float32(180.0))
    sc = _sincosf([np.float32(kps['angle'][i]) * factor for i in sel])
    differing = close = total = n_fma_differs = 0
    for j, i in enumerate(sel):
        b, a = sc[j]
        plain = _describe(blur, kps['x'][i], kps['y'][i], a, b, pat, fma=False)
        fused = _describe(blur, kps['x'][i], kps['y'][i], a, b, pat, fma=True)
        assert np.array_equal(fused, desc[i]), 'FMA restatement differs from the oracle'
        n_fma_differs += not np.array_equal(plain, fused)
        bits_o = np.unpackbits(desc[i].reshape(32, 1), axis=1)[:, ::-1].reshape(-1)
        bits_c = np.unpackbits(d2[j].reshape(32, 1), axis=1)[:, ::-1].reshape(-1)
        total += 256
        for pidx in np.flatnonzero(bits_o != bits_c):
            differing += 1
            v = _samples(blur, kps['x'][i], kps['y'][i], a, b, pat[pidx])
            close += abs(int(v[0]) - int(v[1])) <= 1
    assert differing <= 0.002 * total, (differing, total)
    assert close == differing, 'a bit that cv2.ORB disagrees on is not a near-tie of the two samples'
    assert n_fma_differs <= len(sel) // 10      # the contraction changes a rounding only rarely
