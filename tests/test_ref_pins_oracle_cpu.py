"""The oracle is pinned by the REFERENCE ITSELF: oracle/_ref/libref_orb.so is /root/reference/src/ORBextractor.cc (the whole
file: ctor tables, ComputePyramid, ComputeKeyPointsOctTree, DivideNode / compareNodes / DistributeOctTree, IC_Angle,
computeOrbDescriptor, operator()) and the matcher-side function bodies (both SearchByProjection overloads, SearchForInitialization,
ComputeThreeMaxima, DescriptorDistance, Frame::AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid / isInFrustum,
MapPoint::PredictScale, Pinhole::project) compiled verbatim with the reference's flags (-O3, FMA contraction on) against type
stand-ins; the five OpenCV image primitives underneath are the cv2-pinned ones (tests/test_oracle_cpu.py).  Everything is compared
bit for bit with the restated oracle that the GPU parity tests use."""
import numpy as np
import pytest

import frustum_scenes
import matcher_scenes
import oracle_lib as O
import ref_lib as R
from orb_slam3_modified_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built (needs /root/reference at build time)')


def _same_extraction(a, b):
    (m1, k1, d1), (m2, k2, d2) = a, b
    assert m1 == m2 and len(k1) == len(k2)
    assert k1.tobytes() == k2.tobytes()
    assert np.array_equal(d1, d2)


@pytest.mark.parametrize('cfg', [
    dict(w=640, h=480, nf=1000, lap=(0, 1000), ts=(0, 1, 7, 12, 40)),
    dict(w=640, h=480, nf=1000, lap=(200, 400), ts=(3,)),
    dict(w=1280, h=720, nf=1000, lap=(0, 1000), ts=(5,)),
    dict(w=640, h=480, nf=5000, lap=(0, 1000), ts=(2,)),          # the 5x extractor of the monocular initialiser (Tracking.cc:603)
    dict(w=600, h=800, nf=1500, lap=(0, 0), ts=(4,)),             # portrait
    dict(w=1241, h=376, nf=2000, lap=(0, 0), ts=(6,)),            # KITTI aspect: two initial quadtree nodes... (nIni = 3)
    dict(w=640, h=480, nf=1000, lap=(0, 1000), ts=(9,), scale=1.3, nlevels=6, ini=15, mn=5),
], ids=lambda c: '%dx%d_n%d' % (c['w'], c['h'], c['nf']))
def test_operator_call_equals_reference(cfg):
    args = (cfg['nf'], cfg.get('scale', 1.2), cfg.get('nlevels', 8), cfg.get('ini', 20), cfg.get('mn', 7))
    ref, ora = R.RefExtractor(*args), O.OracleExtractor(*args)
    tr, to = ref.tables(), ora.tables()
    for k in tr:
        assert np.array_equal(tr[k], to[k]), k
    for t in cfg['ts']:
        img = synth.frame(t, cfg['w'], cfg['h'])
        _same_extraction(ref(img, cfg['lap']), ora(img, cfg['lap']))
        for l in range(args[2]):
            assert np.array_equal(ref.level(l), ora.level(l)), ('pyramid', l)


def test_degenerate_images_equal_reference():
    ref, ora = R.RefExtractor(), O.OracleExtractor()
    rng = np.random.default_rng(5)
    flat = np.full((480, 640), 77, np.uint8)
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    grad = (np.add.outer(np.arange(480), np.arange(640)) // 5).astype(np.uint8)
    low = (128 + rng.integers(-6, 7, (480, 640))).astype(np.uint8)
    for img in (flat, noise, grad, low):
        _same_extraction(ref(img, (0, 1000)), ora(img, (0, 1000)))
    assert ref(np.zeros((0, 0), np.uint8))[0] == -1 == ora(np.zeros((0, 0), np.uint8))[0]


def test_distribute_oct_tree_equals_reference_on_random_candidates():
    # DistributeOctTree / DivideNode / compareNodes on candidate sets the images above do not produce: heavy ties in
    # (node size, UL.x), clustered points, fewer candidates than N
    ref, ora = R.RefExtractor(), O.OracleExtractor()
    rng = np.random.default_rng(11)
    for case in range(60):
        n = int(rng.integers(1, 4000))
        W, H = int(rng.integers(120, 1300)), int(rng.integers(100, 760))
        if W < H // 2 + 1:
            W = H
        c = np.zeros(n, O.KP_DTYPE)
        if case % 3 == 0:     # clusters
            cx, cy = rng.uniform(0, W, 6), rng.uniform(0, H, 6)
            k = rng.integers(0, 6, n)
            c['x'] = np.clip(np.round(cx[k] + rng.normal(0, 12, n)), 0, W - 1)
            c['y'] = np.clip(np.round(cy[k] + rng.normal(0, 12, n)), 0, H - 1)
        else:
            c['x'] = rng.integers(0, W, n)
            c['y'] = rng.integers(0, H, n)
        c['response'] = rng.integers(7, 60, n)    # few distinct responses: ties inside nodes too
        c['size'] = 7; c['angle'] = -1; c['class_id'] = -1
        N = int(rng.integers(1, 1200))
        a = ref.distribute(c, 0, W, 0, H, N)
        b = ora.distribute(c, 0, W, 0, H, N)
        assert a.tobytes() == b.tobytes(), case


def test_ic_angle_and_descriptor_equal_reference_on_keypoints():
    # file-static helpers of src/ORBextractor.cc called directly: angle on the level plane, descriptor on the blurred plane
    ref, ora = R.RefExtractor(), O.OracleExtractor()
    img = synth.frame(21)
    _, kps, desc = ora(img, (0, 0))          # lap (0,0): keypoints fill from the front, level order
    l0 = kps[kps['octave'] == 0]
    d0 = desc[kps['octave'] == 0]
    blurred = O.blur7(img)
    for i in range(0, len(l0), 7):
        k = l0[i]
        assert np.float32(ref.ic_angle(img, float(k['x']), float(k['y']))) == k['angle']
        assert np.array_equal(ref.descriptor(blurred, float(k['x']), float(k['y']), float(k['angle'])), d0[i])


def test_descriptor_distance_and_three_maxima_equal_reference():
    rng = np.random.default_rng(2)
    d = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    for i in range(0, 299):
        assert R.descriptor_distance(d[i], d[i + 1]) == O.descriptor_distance(d[i], d[i + 1]) == int(np.unpackbits(d[i] ^ d[i + 1]).sum())
    assert R.compute_three_maxima([0] * 30) == (-1, -1, -1)
    assert R.compute_three_maxima([5, 0, 100, 9, 50] + [0] * 25) == (2, 4, -1)
    assert R.compute_three_maxima([5, 0, 100, 10, 50] + [0] * 25) == (2, 4, 3)


def test_features_in_area_equals_reference():
    kps, _ = matcher_scenes.extract(4)
    rng = np.random.default_rng(8)
    b = (0.0, 0.0, 640.0, 480.0)
    for _ in range(300):
        x, y, r = float(rng.uniform(-30, 670)), float(rng.uniform(-30, 510)), float(rng.uniform(1, 60))
        lo, hi = int(rng.integers(-1, 7)), int(rng.integers(-1, 8))
        assert np.array_equal(R.features_in_area(kps, b, x, y, r, lo, hi), O.features_in_area(kps, b, x, y, r, lo, hi))


@pytest.mark.parametrize('t,check_ori,th', [(3, 1, 15.0), (8, 0, 7.0), (15, 1, 30.0), (22, 1, 15.0)])
def test_search_last_frame_equals_reference(t, check_ori, th):
    sc = matcher_scenes.last_frame_scene(t)
    K = len(sc['kps'])
    rng = np.random.default_rng(t)
    res = []
    for fn in (R.search_last_frame, O.search_last_frame):
        match = np.full(K, -1, np.int32); claimed = np.zeros(K, np.uint8)
        pre = np.random.default_rng(t + 100).random(K) < 0.05          # some keypoints already hold points from an earlier search
        match[pre] = 5000 + np.arange(pre.sum()); claimed[pre] = (np.random.default_rng(t + 101).random(pre.sum()) < 0.5)
        n = fn(sc['kps'], sc['desc'], sc['bounds'], sc['sf'], sc['Tcw'], sc['cam'], sc['last'], th, check_ori, match, claimed)
        res.append((n, match, claimed))
    assert res[0][0] == res[1][0] and res[0][0] > 100
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize('t,th,far', [(4, 1.0, 0), (9, 3.0, 0), (14, 3.0, 1), (20, 5.0, 0)])
def test_search_local_map_equals_reference(t, th, far):
    sc = matcher_scenes.local_map_scene(t)
    K = len(sc['kps'])
    res = []
    for fn in (R.search_local_map, O.search_local_map):
        match = np.full(K, -1, np.int32); claimed = np.zeros(K, np.uint8)
        pre = np.random.default_rng(t + 7).random(K) < 0.3             # as after TrackWithMotionModel: part of the frame is matched
        match[pre] = 9000 + np.arange(pre.sum()); claimed[pre] = (np.random.default_rng(t + 8).random(pre.sum()) < 0.8)
        n = fn(sc['kps'], sc['desc'], sc['bounds'], sc['sf'], sc['pts'], th, 0.8, far, 6.0, match, claimed)
        res.append((n, match, claimed))
    assert res[0][0] == res[1][0] and res[0][0] > 50
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])


@pytest.mark.parametrize('t,window,ori', [(2, 100, True), (6, 30, True), (11, 100, False)])
def test_search_for_initialization_equals_reference(t, window, ori):
    k1, d1 = matcher_scenes.extract(t)
    k2, d2 = matcher_scenes.extract(t + 1)
    sf = O.OracleExtractor().tables()['scale']
    prev = np.stack([k1['x'], k1['y']], 1)
    a = R.search_for_initialization(k1, d1, k2, d2, (0.0, 0.0, 640.0, 480.0), sf, prev, window, 0.9, ori)
    b = O.search_for_initialization(k1, d1, k2, d2, (0.0, 0.0, 640.0, 480.0), sf, prev, window, 0.9, ori)
    assert a[0] == b[0] and a[0] > 30
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize('seed', [0, 3, 9])
def test_is_in_frustum_equals_reference(seed):
    sc = frustum_scenes.scene(5000, seed=seed)
    a = R.is_in_frustum(**sc)
    b = O.is_in_frustum(**sc)
    assert 0.1 < a['inView'].mean() < 0.9
    for k in ('inView', 'projX', 'projY', 'depth', 'level', 'viewCos', 'projXR'):
        assert a[k].tobytes() == b[k].tobytes(), k


@pytest.mark.parametrize('t,w,h', [(3, 640, 480), (11, 752, 480)])
def test_compute_stereo_matches_equals_reference(t, w, h):
    """Frame::ComputeStereoMatches (src/Frame.cc:811-982), the stereo consumer of mvImagePyramid: the reference's own body on the reference's own
    extractors vs the oracle restatement on the oracle's pyramid planes -- mvuRight and mvDepth bit for bit."""
    left, right = synth.stereo_pair(t, w, h)
    rl, rr = R.RefExtractor(1200, 1.2, 8, 20, 7), R.RefExtractor(1200, 1.2, 8, 20, 7)
    ol, orr = O.OracleExtractor(1200, 1.2, 8, 20, 7), O.OracleExtractor(1200, 1.2, 8, 20, 7)
    _, kl, dl = rl(left, (0, 0)); _, kr, dr = rr(right, (0, 0))
    _same_extraction((0, kl, dl), (0,) + ol(left, (0, 0))[1:]); _same_extraction((0, kr, dr), (0,) + orr(right, (0, 0))[1:])
    tb = ol.tables()
    mbf, mb = 22.0 * 8.0, 22.0 * 8.0 / 458.0 * 1.0     # disparity 22 px at depth 8 (arbitrary units): maxD = mbf / mb = 458 px
    a = R.stereo_matches(rl, rr, kl, dl, kr, dr, tb['scale'], tb['inv_scale'], mb, mbf)
    b = O.stereo_matches([ol.level(l) for l in range(8)], [orr.level(l) for l in range(8)], kl, dl, kr, dr, tb['scale'], tb['inv_scale'], mb, mbf)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    ok = a[0] >= 0
    assert ok.sum() > 300 and np.abs((kl['x'][ok] - a[0][ok]) - 22.0).max() < 2.0      # the plane's disparity is recovered
    # the reflected 19-px frame ComputePyramid writes around every level (src/ORBextractor.cc:1185-1191): reflect-101 of the plane
    for l in (0, 3, 7):
        assert np.array_equal(rl.level(l, border=19), np.pad(ol.level(l), 19, mode='reflect'))


@pytest.mark.parametrize('k,L,levelsup,weighting,scoring', [(10, 3, 1, 0, 0), (10, 4, 2, 0, 0), (6, 5, 4, 0, 0), (10, 3, 4, 2, 0), (8, 3, 2, 1, 0), (9, 3, 1, 3, 0), (10, 3, 1, 2, 1), (10, 3, 2, 0, 1)])
def test_dbow2_transform_equals_reference(k, L, levelsup, weighting, scoring):
    """TemplatedVocabulary::transform (the reference's own DBoW2 sources in oracle/_ref) vs the oracle restatement on synthetic vocabulary trees:
    BowVector (word ids, values bit for bit), FeatureVector (node ids, feature lists) and the L1 score of two frames."""
    voc = O.synthetic_vocabulary(k, L, seed=L * 7 + k)
    ref = R.RefVocabulary(voc, weighting=weighting, scoring=scoring)    # 0 = L1_NORM as ORB-SLAM3 configures it, 1 = L2_NORM
    outs = []
    for t in (4, 5):
        _, desc = matcher_scenes.extract(t)
        a = ref.transform(desc, levelsup)
        b = O.bow_transform(voc, desc, levelsup, weighting, 1 + scoring)
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes() and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
        assert len(a[0]) > 50 and abs((a[1].sum() if scoring == 0 else (a[1] ** 2).sum()) - 1.0) < 1e-12 and len(a[3]) <= len(desc)
        outs.append(a)
    if scoring:
        return
    s = ref.score(outs[0], outs[1])
    assert s == O.bow_score_l1(outs[0], outs[1]) and 0 < s < 1 and abs(ref.score(outs[0], outs[0]) - 1.0) < 1e-12


def _bow_scene(t, voc, levelsup=2, seed=0):
    rng = np.random.default_rng(seed)
    kk, dk = matcher_scenes.extract(t); kf, df = matcher_scenes.extract(t + 1)
    fvk = O.bow_transform(voc, dk, levelsup)[2:]; fvf = O.bow_transform(voc, df, levelsup)[2:]
    kf_point = rng.choice([0, 1, 1, 1, 1, 2], len(kk)).astype(np.uint8)
    return kk, dk, kf_point, fvk, kf, df, fvf


@pytest.mark.parametrize('t,k,L,levelsup,ratio,ori', [(3, 10, 4, 2, 0.7, True), (8, 10, 3, 2, 0.75, True), (15, 6, 4, 3, 0.9, False), (21, 10, 3, 1, 0.7, True)])
def test_search_by_bow_equals_reference(t, k, L, levelsup, ratio, ori):
    voc = O.synthetic_vocabulary(k, L, seed=k + L)
    sc = _bow_scene(t, voc, levelsup, seed=t)
    a = R.search_by_bow(*sc, nnratio=ratio, check_ori=ori)
    b = O.search_by_bow(*sc, nnratio=ratio, check_ori=ori)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 20, (a[0], b[0])


def test_compute_distinctive_descriptors_equals_reference():
    rng = np.random.default_rng(6)
    _, desc = matcher_scenes.extract(5)
    for n in list(range(1, 12)) + [17, 32, 33, 64]:
        base = desc[rng.integers(0, len(desc))]
        obs = np.tile(base, (n, 1))
        for i in range(n):
            for bit in rng.integers(0, 256, rng.integers(0, 40)):
                obs[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
        if n > 3:
            obs[2] = obs[1]            # duplicates: ties between medians -> the first row wins
        a, b = R.distinctive_descriptor(obs), O.distinctive_descriptor(obs)
        assert np.array_equal(obs[a], obs[b]), (n, a, b)      # the reference keeps a clone of the chosen row: compare bytes


@pytest.mark.parametrize('t,th', [(6, 3.0), (14, 3.0), (23, 5.0)])
def test_fuse_search_equals_reference(t, th):
    """ORBmatcher::Fuse (src/ORBmatcher.cc:1148-1338): the reference's own body run on stand-in KeyFrame / MapPoint objects that log what Fuse did,
    against the oracle's per-map-point search result (best keypoint, best distance): every logged action names the oracle's best keypoint, and the
    return value is the number of oracle hits within TH_LOW."""
    sc = O.fuse_scene(t)
    n, act, idx = R.fuse(sc, th)
    bi, bd = O.fuse_search(sc, th)
    hit = bd <= 50
    assert n == int(hit.sum()) and n > 100, (n, int(hit.sum()))
    # every hit is one logged action at the oracle's best keypoint (a point the loop made bad earlier by Replace is still searched: the reference
    # tests isBad() once, before the search), and nothing else is logged
    assert np.array_equal(act > 0, hit) and np.array_equal(idx[hit], bi[hit])
    assert {1, 2, 3} <= set(act[hit].tolist())
    assert not hit[sc['state'] != 1].any()


@pytest.mark.parametrize('t,dt,coarse,ori', [(4, 3, False, True), (11, 2, False, True), (17, 5, False, False), (22, 3, True, True), (9, -4, False, True)])
def test_search_for_triangulation_equals_reference(t, dt, coarse, ori):
    """ORBmatcher::SearchForTriangulation + Pinhole::epipolarConstrain (src/ORBmatcher.cc:907-1146, src/CameraModels/Pinhole.cpp:107-129): the
    reference's own bodies against the oracle, given the epipole and fundamental matrix the bodies worked with."""
    sc = O.triangulation_scene(t, dt)
    n, m12, ep, F12 = R.search_for_triangulation(sc, coarse, ori)
    on, om = O.search_for_triangulation(sc, ep, F12, coarse, ori)
    assert n == on and np.array_equal(m12, om), (n, on, int((m12 != om).sum()))
    assert n > 30
    assert not sc['mp1'][m12 >= 0].any() and not sc['mp2'][m12[m12 >= 0]].any()


@pytest.mark.parametrize('t,dt,k,L,levelsup,ratio,ori', [(3, 1, 10, 4, 2, 0.8, True), (8, 2, 10, 3, 2, 0.75, True), (15, 1, 6, 4, 3, 0.9, False), (21, 3, 10, 3, 1, 0.8, True)])
def test_search_by_bow_kf_kf_equals_reference(t, dt, k, L, levelsup, ratio, ori):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (src/ORBmatcher.cc:765-905): the reference's own body against the oracle."""
    voc = O.synthetic_vocabulary(k, L, seed=k + L)
    rng = np.random.default_rng(t)
    k1, d1 = matcher_scenes.extract(t); k2, d2 = matcher_scenes.extract(t + dt)
    fv1 = O.bow_transform(voc, d1, levelsup)[2:]; fv2 = O.bow_transform(voc, d2, levelsup)[2:]
    p1 = rng.choice([0, 1, 1, 1, 1, 2], len(k1)).astype(np.uint8); p2 = rng.choice([0, 1, 1, 1, 1, 2], len(k2)).astype(np.uint8)
    n, m = R.search_by_bow_kf(k1, d1, p1, fv1, k2, d2, p2, fv2, nnratio=ratio, check_ori=ori)
    on, om = O.search_by_bow_kf(k1, d1, p1, fv1, k2, d2, p2, fv2, nnratio=ratio, check_ori=ori)
    assert n == on and np.array_equal(m, om) and n > 20, (n, on)
    assert (p1[m >= 0] == 1).all() and (p2[m[m >= 0]] == 1).all() and len(set(m[m >= 0].tolist())) == (m >= 0).sum()


@pytest.mark.parametrize('t,dt,th', [(5, 2, 7.5), (13, 3, 7.5), (20, 1, 4.0)])
def test_search_by_sim3_equals_reference(t, dt, th):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1457-1674): the reference's own body against the oracle, given the camera-frame points the body
    computed from S12 / S21."""
    sc = O.sim3_scene(t, dt)
    n, m12, p21, p12 = R.search_by_sim3(sc, th)
    on, om = O.search_by_sim3(sc, p21, p12, th)
    assert n == on and np.array_equal(m12, om), (n, on, int((m12 != om).sum()))
    assert n > 50 and (m12[sc['pre12'] >= 0] == sc['pre12'][sc['pre12'] >= 0]).all()


@pytest.mark.parametrize('t,th', [(6, 3.0), (14, 4.0)])
def test_fuse_sim3_search_equals_reference(t, th):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1340-1455): every action the reference body took names the oracle's
    best keypoint, the return value is the number of oracle hits within TH_LOW."""
    sc = O.fuse_scene(t)
    rng = np.random.default_rng(t)
    sc['state'] = np.where(sc['state'] == 0, 1, sc['state']).astype(np.uint8)          # this overload is never handed NULL points
    kf_state = rng.choice([0, 0, 1, 1, 2], len(sc['kps'])).astype(np.uint8)
    s = np.float32(1.0 + 0.02 * rng.normal())
    Scw = np.concatenate([[s], sc['Tcw'][:4], sc['Tcw'][4:] * s]).astype(np.float32)
    n, act, idx, T7, Ow = R.fuse_sim3(sc, Scw, kf_state, th)
    bi, bd = O.fuse_search_sim3(sc, T7, Ow, th)
    hit = bd <= 50
    assert n == int(hit.sum()) and n > 100, (n, int(hit.sum()))
    logged = act > 0
    assert hit[logged].all() and np.array_equal(idx[logged], bi[logged])
    # hits without a logged action: the keyframe's point at that keypoint is bad (counted, nothing done)
    silent = hit & ~logged
    assert (kf_state[bi[silent]] == 2).all() and {1, 2} <= set(act[hit].tolist())
    assert not hit[sc['state'] != 1].any()
