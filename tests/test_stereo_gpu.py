"""GPU parity of Frame::ComputeStereoMatches (src/Frame.cc:811-982) and of the bordered pyramid: mvuRight / mvDepth bit for
bit against the oracle (which equals the reference's own function body, tests/test_ref_pins_oracle_cpu.py)."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('w,h,nf,B', [(640, 480, 1200, 3), (752, 480, 2000, 1)])
def test_compute_stereo_matches(w, h, nf, B):
    import orb_slam3_modified_b200 as orb
    pairs = [synth.stereo_pair(4 + 3 * b, w, h, seed=b) for b in range(B)]
    exl = orb.ORBextractor(nf, 1.2, 8, 20, 7, w, h, B); exr = orb.ORBextractor(nf, 1.2, 8, 20, 7, w, h, B)
    _, kl, dl = exl.extract_batch(np.stack([p[0] for p in pairs]), (0, 0))
    _, kr, dr = exr.extract_batch(np.stack([p[1] for p in pairs]), (0, 0))
    mbf = 22.0 * 8.0; mb = mbf / 458.0
    ur, dep = exl.ComputeStereoMatches(exr, mb, mbf, B)
    ol, orr = O.OracleExtractor(nf, 1.2, 8, 20, 7), O.OracleExtractor(nf, 1.2, 8, 20, 7)
    tb = ol.tables()
    for b in range(B):
        _, okl, odl = ol(pairs[b][0], (0, 0)); _, okr, odr = orr(pairs[b][1], (0, 0))
        assert kl[b].tobytes() == okl.tobytes() and kr[b].tobytes() == okr.tobytes()
        our, odep = O.stereo_matches([ol.level(l) for l in range(8)], [orr.level(l) for l in range(8)], okl, odl, okr, odr, tb['scale'], tb['inv_scale'], mb, mbf)
        n = len(okl)
        assert ur[b, :n].tobytes() == our.tobytes() and dep[b, :n].tobytes() == odep.tobytes(), (b, int((ur[b, :n] != our).sum()))
        assert (our >= 0).sum() > 300
        if b == 0:
            for l in (0, 2, 7):    # the 19-px reflected frame of mvImagePyramid, materialised on request
                assert np.array_equal(exl.level_bordered(l, frame=b), np.pad(ol.level(l), 19, mode='reflect'))
