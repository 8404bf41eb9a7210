"""GPU parity of orbm_frustum_project (Frame::isInFrustum + MapPoint::PredictScale): every output field bit-exact against the
oracle, and the chain frustum -> local-map SearchByProjection gives the oracle's matches."""
import numpy as np
import pytest

import frustum_scenes
import oracle_lib as O

pytestmark = pytest.mark.gpu
FIELDS = ('inView', 'projX', 'projY', 'projXR', 'depth', 'level', 'viewCos')


@pytest.fixture(scope='module')
def matcher():
    import orb_slam3_modified_b200 as m
    m.lib()
    return m.ORBmatcher(0.8, True, max_keypoints=2048, max_mappoints=20000)


@pytest.mark.parametrize('M,seed', [(1, 0), (257, 1), (3000, 2), (20000, 3)])
def test_frustum_bit_exact(matcher, M, seed):
    sc = frustum_scenes.scene(M, seed)
    ref = O.is_in_frustum(**sc)
    out = matcher.isInFrustum(sc['pts'], sc['Rcw'], sc['tcw'], sc['Ow'], sc['cam'], sc['bounds'], sc['log_scale_factor'], sc['n_levels'], 0.5, sc['mbf'])
    for k in FIELDS:
        assert out[k].tobytes() == ref[k].tobytes(), k


def test_frustum_edge_cases(matcher):
    import orb_slam3_modified_b200 as m
    sc = frustum_scenes.scene(64, 5)
    empty = {k: v[:0] for k, v in sc['pts'].items()}
    out = matcher.isInFrustum(empty, sc['Rcw'], sc['tcw'], sc['Ow'], sc['cam'], sc['bounds'], sc['log_scale_factor'], 8)
    assert all(len(out[k]) == 0 for k in FIELDS)
    # points exactly on the optical centre plane / at the camera centre: same IEEE special values as the oracle
    pts = {k: v.copy() for k, v in sc['pts'].items()}
    pts['worldPos'][0] = sc['Ow']
    pts['worldPos'][1] = sc['Ow'] + sc['Rcw'].T @ np.array([1.0, 0.5, 0.0], np.float32)
    ref = O.is_in_frustum(pts, sc['Rcw'], sc['tcw'], sc['Ow'], sc['cam'], sc['bounds'], sc['log_scale_factor'], 8, 0.5, sc['mbf'])
    out = matcher.isInFrustum(pts, sc['Rcw'], sc['tcw'], sc['Ow'], sc['cam'], sc['bounds'], sc['log_scale_factor'], 8, 0.5, sc['mbf'])
    for k in FIELDS:
        assert np.array_equal(out[k], ref[k], equal_nan=True) if out[k].dtype.kind == 'f' else out[k].tobytes() == ref[k].tobytes(), k
    with pytest.raises(m.OrbError):
        big = frustum_scenes.scene(30000, 6)
        matcher.isInFrustum(big['pts'], big['Rcw'], big['tcw'], big['Ow'], big['cam'], big['bounds'], big['log_scale_factor'], 8)


def test_frustum_feeds_local_map_search(matcher):
    """SearchLocalPoints (src/Tracking.cc:3346): isInFrustum per local map point, then SearchByProjection(F, points, th)."""
    import orb_slam3_modified_b200 as m
    rng = np.random.default_rng(11)
    sc = frustum_scenes.scene(2500, 7)
    fr = matcher.isInFrustum(sc['pts'], sc['Rcw'], sc['tcw'], sc['Ow'], sc['cam'], sc['bounds'], sc['log_scale_factor'], sc['n_levels'], 0.5, sc['mbf'])
    ref = O.is_in_frustum(**sc)
    # a frame whose keypoints sit near the projections of the points in view, with descriptors close to the points'
    iv = np.flatnonzero(ref['inView'])
    K = min(1500, len(iv))
    pick = rng.choice(iv, K, replace=False)
    mp_desc = rng.integers(0, 256, (len(ref['inView']), 32), dtype=np.uint8)
    kps, desc, sf = frustum_scenes.keypoints_near(ref['projX'][pick], ref['projY'][pick], ref['level'][pick], mp_desc[pick], rng)
    pts = lambda f: dict(inView=f['inView'], bad=np.zeros(len(f['inView']), np.uint8), depth=f['depth'], projX=f['projX'], projY=f['projY'],
                         level=f['level'], viewCos=f['viewCos'], hasObs=np.ones(len(f['inView']), np.uint8), descriptors=mp_desc)
    F = m.Frame(kps, desc, sc['bounds'], sf)
    n_gpu = matcher.SearchByProjection(F, pts(fr), 3.0, False, 50.0)
    match = np.full(len(kps), -1, np.int32); claimed = np.zeros(len(kps), np.uint8)
    n_ref = O.search_local_map(kps, desc, sc['bounds'], sf, pts(ref), 3.0, 0.8, False, 50.0, match, claimed)
    assert n_gpu == n_ref and n_ref > 200
    assert np.array_equal(F.match, match) and np.array_equal(F.claimed, claimed)
