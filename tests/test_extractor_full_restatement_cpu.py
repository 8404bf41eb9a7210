"""ORBextractor::operator() end to end a second time (src/ORBextractor.cc:1086-1168 with ComputePyramid :1170-1195 and
ComputeKeyPointsOctTree :781-896): real cv2 primitives (`resize`, `FAST`, `GaussianBlur`, `fastAtan2`), the Python transcription of
the quadtree, numpy IC_Angle and steered BRIEF, and the output placement (scaling by mvScaleFactor, lapping area filled from the
back) -- against the C++ oracle's final keypoints and descriptors, bit for bit."""
import math

import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth
from test_brief_vs_cv2_cpu import _describe, _pattern, _sincosf
from test_orientation_cpu import HALF_PATCH, _umax
from test_quadtree_transcription_cpu import distribute

cv2 = pytest.importorskip('cv2')
EDGE_THRESHOLD, PATCH_SIZE = 19, 31


def _extract(img, nfeatures, scale_factor, nlevels, ini, mn, lap):
    ex = O.OracleExtractor(nfeatures, scale_factor, nlevels, ini, mn)      # only for the constructor's tables (row a1, checked elsewhere)
    tab = ex.tables()
    sf, inv, per_level = np.asarray(tab['scale'], np.float32), np.asarray(tab['inv_scale'], np.float32), tab['features_per_level']
    umax, pat = _umax(), _pattern()
    factor = np.float32(3.14159265358979323846 / np.float32(180.0))
    levels, all_k = [], []
    prev = img
    for level in range(nlevels):
        if level:
            sz = (int(np.rint(np.float32(img.shape[1]) * inv[level])), int(np.rint(np.float32(img.shape[0]) * inv[level])))
            prev = cv2.resize(prev, sz, interpolation=cv2.INTER_LINEAR)
        levels.append(prev)
        plane = prev
        minB = EDGE_THRESHOLD - 3
        maxBX, maxBY = plane.shape[1] - EDGE_THRESHOLD + 3, plane.shape[0] - EDGE_THRESHOLD + 3
        width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
        nCols, nRows = int(width / np.float32(35)), int(height / np.float32(35))
        wCell, hCell = int(math.ceil(width / nCols)), int(math.ceil(height / nRows))
        X, Y, Rp = [], [], []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                roi = np.ascontiguousarray(plane[iniY:maxY, iniX:maxX])
                k = cv2.FastFeatureDetector_create(ini, True).detect(roi) or cv2.FastFeatureDetector_create(mn, True).detect(roi)
                for p in k:
                    X.append(np.float32(p.pt[0] + j * wCell)); Y.append(np.float32(p.pt[1] + i * hCell)); Rp.append(np.float32(p.response))
        X, Y, Rp = np.array(X, np.float32), np.array(Y, np.float32), np.array(Rp, np.float32)
        idx = distribute(X, Y, Rp, minB, maxBX, minB, maxBY, int(per_level[level])) if len(X) else []
        size = np.float32(int(np.float32(PATCH_SIZE) * sf[level]))
        I = plane.astype(np.int64)
        kl = []
        for k in idx:
            x, y = np.float32(X[k] + np.float32(minB)), np.float32(Y[k] + np.float32(minB))
            xi, yi = int(np.rint(x)), int(np.rint(y))
            m10 = m01 = 0
            for v in range(-HALF_PATCH, HALF_PATCH + 1):
                d = umax[abs(v)]
                row = I[yi + v, xi - d:xi + d + 1]
                m10 += int((np.arange(-d, d + 1) * row).sum()); m01 += v * int(row.sum())
            kl.append([x, y, size, np.float32(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10)))), Rp[k], level])
        all_k.append(kl)
    n = sum(len(k) for k in all_k)
    kps = np.zeros(n, O.KP_DTYPE)
    desc = np.zeros((n, 32), np.uint8)
    mono, stereo = 0, n - 1
    for level in range(nlevels):
        if not all_k[level]:
            continue
        blur = cv2.GaussianBlur(levels[level], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        sc = _sincosf([k[3] * factor for k in all_k[level]])
        for (x, y, size, ang, rsp, lv), (b, a) in zip(all_k[level], sc):
            d = _describe(blur, x, y, a, b, pat, fma=True)
            if level:
                x, y = np.float32(x * sf[level]), np.float32(y * sf[level])
            if lap[0] <= x <= lap[1]:
                at = stereo; stereo -= 1
            else:
                at = mono; mono += 1
            kps[at] = (x, y, size, ang, rsp, lv, -1)
            desc[at] = d
    return mono, kps, desc


@pytest.mark.parametrize('t,w,h,nf,ini,mn,lap', [(2, 640, 480, 1000, 20, 7, (0, 1000)), (7, 640, 480, 500, 20, 7, (200, 400)),
                                                 (5, 1280, 720, 1000, 20, 7, (0, 1000))])
def test_whole_extractor_second_restatement(t, w, h, nf, ini, mn, lap):
    img = synth.frame(t, w, h)
    mono_o, kps_o, desc_o = O.OracleExtractor(nf, 1.2, 8, ini, mn)(img, lap)
    mono, kps, desc = _extract(img, nf, 1.2, 8, ini, mn, lap)
    assert mono == mono_o and len(kps) == len(kps_o) > 0.8 * nf
    for f in ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id'):
        assert np.array_equal(kps[f], kps_o[f]), f
    assert np.array_equal(desc, desc_o)
