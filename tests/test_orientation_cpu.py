"""Independent check of IC_Angle / computeOrientation (src/ORBextractor.cc:76-103,471-478): intensity-centroid
moments over the radius-15 disc written straight from the definition in numpy + `cv2.fastAtan2`, against the angles the oracle
assigns to its level-0 keypoints; and the `umax` table against its defining construction (:452-469)."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

cv2 = pytest.importorskip('cv2')
HALF_PATCH = 15


def _umax():
    # ORBextractor.cc:452-469: quarter circle by rounding, then made symmetric about the diagonal
    umax = np.zeros(HALF_PATCH + 1, np.int32)
    vmax = int(np.floor(HALF_PATCH * np.sqrt(2.0) / 2 + 1))
    vmin = int(np.ceil(HALF_PATCH * np.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(np.sqrt(float(HALF_PATCH * HALF_PATCH - v * v))))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


def test_umax_table_and_angles():
    ex = O.OracleExtractor(1000, 1.2, 8, 20, 7)
    umax = _umax()
    assert np.array_equal(np.asarray(ex.tables()['umax'])[:HALF_PATCH + 1], umax)
    for t in (2, 9):
        img = synth.frame(t)
        _, kps, _ = ex(img, (0, 1000))
        sel = np.flatnonzero(kps['octave'] == 0)
        assert len(sel) > 100
        I = img.astype(np.int64)
        for i in sel:
            x, y = int(np.rint(kps['x'][i])), int(np.rint(kps['y'][i]))
            m10 = m01 = 0
            for v in range(-HALF_PATCH, HALF_PATCH + 1):
                d = umax[abs(v)]
                u = np.arange(-d, d + 1)
                row = I[y + v, x - d:x + d + 1]
                m10 += int((u * row).sum())
                m01 += v * int(row.sum())
            ang = cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10)))
            assert np.float32(ang) == kps['angle'][i], (i, ang, kps['angle'][i])
