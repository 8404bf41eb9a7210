"""CPU tests of the matcher oracle: golden vectors (cv2 BFMatcher), Hamming properties, grid queries, sanity of the
projection matchers on seeded scenes."""
import os

import numpy as np
import pytest

import oracle_lib as O
import matcher_scenes as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_bf_knn2_golden():
    g = np.load(os.path.join(G, 'primitives.npz'))
    idx, dist = O.bf_knn2(g['bf_q'], g['bf_t'])
    assert np.array_equal(idx, g['bf_idx']) and np.array_equal(dist, g['bf_dist'])


def test_bf_knn2_vs_cv2_live_with_ties():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (300, 32)).astype(np.uint8)
    t = rng.integers(0, 256, (400, 32)).astype(np.uint8)
    t[100:200] = t[:100]          # exact duplicates -> distance ties
    q[:50] = t[:50]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx, dist = O.bf_knn2(q, t)
    assert np.array_equal(idx, np.array([[a.trainIdx, b.trainIdx] for a, b in m]))
    assert np.array_equal(dist, np.array([[int(a.distance), int(b.distance)] for a, b in m]))


def test_descriptor_distance_properties():
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (64, 32)).astype(np.uint8)
    for i in range(64):
        d = O.descriptor_distance(a[i], b[i])
        assert d == int(np.unpackbits(a[i] ^ b[i]).sum()) == O.descriptor_distance(b[i], a[i])
        assert O.descriptor_distance(a[i], a[i]) == 0
    assert O.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_features_in_area_bruteforce():
    kps, _ = S.extract(3)
    rng = np.random.default_rng(4)
    b = (0.0, 0.0, 640.0, 480.0)
    for _ in range(200):
        x, y, r = rng.uniform(-20, 660), rng.uniform(-20, 500), rng.uniform(1, 60)
        lo, hi = int(rng.integers(-1, 6)), int(rng.integers(-1, 8))
        got = O.features_in_area(kps, b, x, y, r, lo, hi)
        x32, y32, r32 = np.float32(x), np.float32(y), np.float32(r)
        ok = (np.abs(kps['x'] - x32) < r32) & (np.abs(kps['y'] - y32) < r32)
        if lo > 0 or hi >= 0:
            ok &= kps['octave'] >= lo
            if hi >= 0:
                ok &= kps['octave'] <= hi
        assert sorted(got.tolist()) == np.nonzero(ok)[0].tolist()
        assert len(set(got.tolist())) == len(got)


def test_last_frame_scene_matches_well():
    s = S.last_frame_scene(5)
    K = len(s['kps'])
    match = np.full(K, -1, np.int32)
    claimed = np.zeros(K, np.uint8)
    n = O.search_last_frame(s['kps'], s['desc'], s['bounds'], s['sf'], s['Tcw'], s['cam'], s['last'], 15.0, True, match, claimed)
    assert n >= int((match >= 0).sum()) > 300      # nmatches also counts overwrites of keypoints held by points without observations
    m = np.nonzero(match >= 0)[0]
    d = [O.descriptor_distance(s['desc'][i], s['last']['descriptors'][match[i]]) for i in m]
    assert max(d) <= 100
    assert all(s['last']['valid'][match[i]] for i in m)


def test_local_map_scene_matches():
    s = S.local_map_scene(6)
    K = len(s['kps'])
    match = np.full(K, -1, np.int32)
    claimed = np.zeros(K, np.uint8)
    n = O.search_local_map(s['kps'], s['desc'], s['bounds'], s['sf'], s['pts'], 1.0, 0.8, False, 50.0, match, claimed)
    assert n >= int((match >= 0).sum()) > 200     # nmatches also counts overwrites of unclaimed keypoints
    assert all(s['pts']['inView'][match[i]] and not s['pts']['bad'][match[i]] for i in np.nonzero(match >= 0)[0])
