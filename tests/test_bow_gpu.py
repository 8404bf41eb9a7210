"""GPU parity of the DBoW2 transform against the oracle (which equals the reference's own DBoW2 sources compiled into
oracle/_ref, tests/test_ref_pins_oracle_cpu.py): BowVector word ids equal and values bit for bit, FeatureVector equal."""
import numpy as np
import pytest

import matcher_scenes as S
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('k,L,levelsup,weighting,norm', [(10, 4, 2, 0, 1), (10, 3, 4, 0, 1), (6, 5, 4, 0, 1), (10, 3, 1, 2, 2), (8, 3, 2, 1, 0), (9, 3, 1, 3, 1)])
def test_transform_batch(k, L, levelsup, weighting, norm):
    import orb_slam3_modified_b200 as orb
    voc = O.synthetic_vocabulary(k, L, seed=3 * L + k)
    v = orb.ORBVocabulary(voc['L'], voc['child_start'], voc['children'], voc['desc'], voc['weight'], voc['word_id'], weighting, norm)
    descs = [S.extract(t)[1] for t in (2, 3, 9)] + [S.extract(5, nf=5000)[1], np.zeros((0, 32), np.uint8), S.extract(4)[1][:1]]
    outs = v.transform(descs, levelsup)
    for d, got in zip(descs, outs):
        ref = O.bow_transform(voc, d, levelsup, weighting, norm)
        assert np.array_equal(got[0], ref[0]) and got[1].tobytes() == ref[1].tobytes(), len(d)
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3]), len(d)
    assert len(outs[0][0]) > 50 and len(outs[4][0]) == 0
    with pytest.raises(orb.OrbError):
        orb.ORBVocabulary(voc['L'], voc['child_start'], voc['children'] + 10 ** 6, voc['desc'], voc['weight'], voc['word_id'])
