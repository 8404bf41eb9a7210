"""GPU parity tests of the matchers (through the C-ABI) vs the CPU oracle: identical match arrays."""
import numpy as np
import pytest

import oracle_lib as O
import matcher_scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def orb():
    import orb_slam3_modified_b200 as m
    m.lib()
    return m


@pytest.fixture(scope='module')
def matcher(orb):
    return orb.ORBmatcher(0.8, True, max_batch=4, max_keypoints=2048, max_mappoints=8192)


def test_descriptor_distance(matcher):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (1000, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (1000, 32)).astype(np.uint8)
    b[:10] = a[:10]
    got = matcher.DescriptorDistance(a, b)
    assert np.array_equal(got, np.unpackbits(a ^ b, axis=1).sum(1))
    assert matcher.DescriptorDistance(a[0], b[500]) == O.descriptor_distance(a[0], b[500])


def test_bf_knn2(matcher):
    rng = np.random.default_rng(1)
    for Q, T in ((300, 400), (1000, 1000), (5, 1), (7, 2), (1, 3000)):
        q = rng.integers(0, 256, (Q, 32)).astype(np.uint8)
        t = rng.integers(0, 256, (T, 32)).astype(np.uint8)
        if T > 200:
            t[100:200] = t[:100]
            q[:min(Q, 50)] = t[:min(Q, 50)]
        idx, dist = matcher.knnMatch2(q, t)
        oi, od = O.bf_knn2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), (Q, T)


def _run_last(orb, matcher, s, th, ori, init=None):
    F = orb.Frame(s['kps'], s['desc'], s['bounds'], s['sf'])
    om = np.full(len(s['kps']), -1, np.int32)
    oc = np.zeros(len(s['kps']), np.uint8)
    if init is not None:
        F.match[:], F.claimed[:] = init
        om[:], oc[:] = init
    matcher.mbCheckOrientation = ori
    n = matcher.SearchByProjection(F, s['last'], th, True, Tcw=s['Tcw'], cam=s['cam'])
    on = O.search_last_frame(s['kps'], s['desc'], s['bounds'], s['sf'], s['Tcw'], s['cam'], s['last'], th, ori, om, oc)
    assert n == on, (n, on)
    assert np.array_equal(F.match, om), int((F.match != om).sum())
    assert np.array_equal(F.claimed, oc)
    return n


@pytest.mark.parametrize('t', [5, 12, 30])
def test_search_last_frame(orb, matcher, t):
    s = S.last_frame_scene(t, seed=t % 2)
    n = _run_last(orb, matcher, s, 15.0, True)
    assert n > 200
    _run_last(orb, matcher, s, 30.0, True)      # the 2*th retry of Tracking.cc:2897: much more contention
    _run_last(orb, matcher, s, 15.0, False)


def test_search_last_frame_contention_and_initial_state(orb, matcher):
    """Heavy claim contention: every map point duplicated 3x + keypoints pre-assigned before the call."""
    s = S.last_frame_scene(8)
    L = s['last']
    rng = np.random.default_rng(3)
    rep = {k: np.concatenate([v, v, v]) for k, v in L.items()}
    rep['hasObs'] = (rng.random(len(rep['valid'])) > 0.3).astype(np.uint8)
    s2 = dict(s, last=rep)
    K = len(s['kps'])
    init = (np.where(rng.random(K) < 0.2, 7, -1).astype(np.int32), (rng.random(K) < 0.5).astype(np.uint8))
    _run_last(orb, matcher, s2, 15.0, True, init)
    _run_last(orb, matcher, s2, 40.0, True)


def test_search_last_frame_degenerate(orb, matcher):
    s = S.last_frame_scene(9)
    empty = {k: v[:0] for k, v in s['last'].items()}
    assert _run_last(orb, matcher, dict(s, last=empty), 15.0, True) == 0
    behind = dict(s['last'])
    behind['xyz'] = behind['xyz'].copy()
    behind['xyz'][:, 2] = -50.0        # behind the camera -> invzc < 0
    assert _run_last(orb, matcher, dict(s, last=behind), 15.0, True) == 0
    nokp = dict(s, kps=s['kps'][:0], desc=s['desc'][:0])
    assert _run_last(orb, matcher, nokp, 15.0, True) == 0


def _run_local(orb, matcher, s, th, nnratio, far=False, thfar=50.0, init=None):
    F = orb.Frame(s['kps'], s['desc'], s['bounds'], s['sf'])
    om = np.full(len(s['kps']), -1, np.int32)
    oc = np.zeros(len(s['kps']), np.uint8)
    if init is not None:
        F.match[:], F.claimed[:] = init
        om[:], oc[:] = init
    matcher.mfNNratio = nnratio
    n = matcher.SearchByProjection(F, s['pts'], th, far, thfar)
    on = O.search_local_map(s['kps'], s['desc'], s['bounds'], s['sf'], s['pts'], th, nnratio, far, thfar, om, oc)
    assert n == on, (n, on)
    assert np.array_equal(F.match, om), int((F.match != om).sum())
    assert np.array_equal(F.claimed, oc)
    return n


@pytest.mark.parametrize('t', [6, 20])
def test_search_local_map(orb, matcher, t):
    s = S.local_map_scene(t, seed=t % 3)
    assert _run_local(orb, matcher, s, 1.0, 0.8) > 100
    _run_local(orb, matcher, s, 3.0, 0.8)             # th=3 as after relocalisation (Tracking.cc:3409-3414)
    _run_local(orb, matcher, s, 5.0, 0.9)
    _run_local(orb, matcher, s, 1.0, 0.8, True, 3.4)  # bFarPoints
    rng = np.random.default_rng(t)
    K = len(s['kps'])
    _run_local(orb, matcher, s, 3.0, 0.8, init=(np.where(rng.random(K) < 0.3, 11, -1).astype(np.int32), (rng.random(K) < 0.6).astype(np.uint8)))


def test_batch_device_matches_host(orb, matcher):
    """Batched device entry point == per-stream host calls."""
    import torch
    B = 3
    scenes = [S.last_frame_scene(5 + 7 * b, seed=b % 2) for b in range(B)]
    kcap, mcap = 1100, 1100
    dev = torch.device('cuda')
    kps = np.zeros((B, kcap), orb.KP_DTYPE); desc = np.zeros((B, kcap, 32), np.uint8); nK = np.zeros(B, np.int32)
    nM = np.zeros(B, np.int32); valid = np.zeros((B, mcap), np.uint8); xyz = np.zeros((B, mcap, 3), np.float32)
    octv = np.zeros((B, mcap), np.int32); ang = np.zeros((B, mcap), np.float32); obs = np.zeros((B, mcap), np.uint8)
    mpd = np.zeros((B, mcap, 32), np.uint8); Tcw = np.zeros((B, 7), np.float32)
    for b, s in enumerate(scenes):
        k, m = len(s['kps']), len(s['last']['valid'])
        nK[b], nM[b] = k, m
        kps[b, :k], desc[b, :k] = s['kps'], s['desc']
        L = s['last']
        valid[b, :m], xyz[b, :m], octv[b, :m], ang[b, :m], obs[b, :m], mpd[b, :m] = L['valid'], L['xyz'], L['octave'], L['angle'], L['hasObs'], L['descriptors']
        Tcw[b] = s['Tcw']
    tt = lambda a: torch.from_numpy(a.view(np.uint8) if a.dtype == orb.KP_DTYPE else a).to(dev)
    d = dict(batch=B, kcap=kcap, mcap=mcap, nlevels=8, kps=tt(kps), desc=tt(desc), nK=tt(nK), scaleFactors=tt(scenes[0]['sf']), nM=tt(nM),
             valid=tt(valid), xyz=tt(xyz), octave=tt(octv), angle=tt(ang), hasObs=tt(obs), mpDesc=tt(mpd), Tcw7=tt(Tcw),
             bounds=(0.0, 0.0, 640.0, 480.0), cam=[float(c) for c in scenes[0]['cam']])
    d_match = torch.full((B, kcap), -1, dtype=torch.int32, device=dev)
    d_claimed = torch.zeros((B, kcap), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    matcher.mbCheckOrientation = True
    matcher.search_last_frame_batch_device(d, 15.0, d_match, d_claimed, d_n, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for b, s in enumerate(scenes):
        k = len(s['kps'])
        om = np.full(k, -1, np.int32); oc = np.zeros(k, np.uint8)
        on = O.search_last_frame(s['kps'], s['desc'], s['bounds'], s['sf'], s['Tcw'], s['cam'], s['last'], 15.0, True, om, oc)
        assert int(d_n[b]) == on
        assert np.array_equal(d_match[b, :k].cpu().numpy(), om) and np.array_equal(d_claimed[b, :k].cpu().numpy(), oc)


@pytest.mark.parametrize('t,window,ori,nf', [(2, 100, True, 1000), (6, 30, True, 1000), (11, 100, False, 1000), (4, 100, True, 5000), (9, 10, True, 1000)])
def test_search_for_initialization(orb, t, window, ori, nf):
    """a13, second call site: ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:648-763) vs the oracle (which equals the reference's
    own function body, tests/test_ref_pins_oracle_cpu.py): vnMatches12, the return value and the updated vbPrevMatched."""
    k1, d1 = S.extract(t, nf=nf)
    k2, d2 = S.extract(t + 1, nf=nf)
    sf = O.OracleExtractor().tables()['scale']
    b = (0.0, 0.0, 640.0, 480.0)
    m = orb.ORBmatcher(0.9, ori, max_batch=1, max_keypoints=8192, max_mappoints=8192)
    prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1), np.float32)
    rng = np.random.default_rng(t)
    prev += rng.normal(0, 3.0, prev.shape).astype(np.float32)            # vbPrevMatched drifts while the initialiser retries
    on, om, oprev = O.search_for_initialization(k1, d1, k2, d2, b, sf, prev, window, 0.9, ori)
    F1, F2 = orb.Frame(k1, d1, b, sf), orb.Frame(k2, d2, b, sf)
    gp = prev.copy()
    n, m12 = m.SearchForInitialization(F1, F2, gp, window)
    assert n == on and (window < 30 or n > 30), (n, on)
    assert np.array_equal(m12, om) and np.array_equal(gp, oprev)
    # second attempt with the updated vbPrevMatched (Tracking::MonocularInitialization keeps it across frames)
    on2, om2, oprev2 = O.search_for_initialization(k1, d1, k2, d2, b, sf, oprev, window, 0.9, ori)
    n2, m122 = m.SearchForInitialization(F1, F2, gp, window)
    assert n2 == on2 and np.array_equal(m122, om2) and np.array_equal(gp, oprev2)


def test_search_for_initialization_heavy_rematching(orb):
    """Many F1 keypoints compete for few F2 keypoints (descriptors duplicated): exercises the vMatchedDistance overwrite rule, the
    displacement of earlier matches and the rescan path."""
    k1, d1 = S.extract(13)
    k2, d2 = S.extract(14)
    sf = O.OracleExtractor().tables()['scale']
    b = (0.0, 0.0, 640.0, 480.0)
    rng = np.random.default_rng(3)
    l0 = np.flatnonzero(k2['octave'] == 0)
    d2 = d2.copy()
    src = rng.choice(l0, 25)
    for i in l0:                                   # level-0 descriptors of F2 collapse onto 25 prototypes (+ a few bit flips)
        d2[i] = d2[src[rng.integers(0, 25)]]
        for bit in rng.integers(0, 256, rng.integers(0, 4)):
            d2[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
    d1 = d1.copy()
    l01 = np.flatnonzero(k1['octave'] == 0)
    for i in l01:
        d1[i] = d2[src[rng.integers(0, 25)]]
        for bit in rng.integers(0, 256, rng.integers(0, 12)):
            d1[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
    prev = np.ascontiguousarray(np.stack([k1['x'], k1['y']], 1), np.float32)
    for ratio in (0.9, 1.5):
        m = orb.ORBmatcher(ratio, True, max_batch=1, max_keypoints=2048, max_mappoints=2048)
        on, om, oprev = O.search_for_initialization(k1, d1, k2, d2, b, sf, prev, 100, ratio, True)
        gp = prev.copy()
        n, m12 = m.SearchForInitialization(orb.Frame(k1, d1, b, sf), orb.Frame(k2, d2, b, sf), gp, 100)
        assert n == on and np.array_equal(m12, om) and np.array_equal(gp, oprev), (ratio, n, on)


@pytest.mark.parametrize('t,k,L,levelsup,ratio,ori', [(3, 10, 4, 2, 0.7, True), (8, 10, 3, 2, 0.75, True), (15, 6, 4, 3, 0.9, False), (21, 10, 3, 1, 0.7, True)])
def test_search_by_bow(orb, t, k, L, levelsup, ratio, ori):
    """f2: ORBmatcher::SearchByBoW(KeyFrame*, Frame&) vs the oracle (= the reference's own body, tests/test_ref_pins_oracle_cpu.py)."""
    voc = O.synthetic_vocabulary(k, L, seed=k + L)
    rng = np.random.default_rng(t)
    kk, dk = S.extract(t); kf, df = S.extract(t + 1)
    fvk = O.bow_transform(voc, dk, levelsup)[2:]; fvf = O.bow_transform(voc, df, levelsup)[2:]
    kf_point = rng.choice([0, 1, 1, 1, 1, 2], len(kk)).astype(np.uint8)
    m = orb.ORBmatcher(ratio, ori, max_batch=1, max_keypoints=2048, max_mappoints=2048)
    n, match = m.SearchByBoW(kk, dk, kf_point, fvk, kf, df, fvf)
    on, om = O.search_by_bow(kk, dk, kf_point, fvk, kf, df, fvf, nnratio=ratio, check_ori=ori)
    assert n == on and np.array_equal(match, om) and n > 20, (n, on)


def test_compute_distinctive_descriptors(orb, matcher):
    rng = np.random.default_rng(6)
    _, desc = S.extract(5)
    obs_list = []
    for n in list(range(0, 12)) + [17, 32, 33, 64, 100] + list(rng.integers(2, 30, 300)):
        base = desc[rng.integers(0, len(desc))]
        obs = np.tile(base, (int(n), 1))
        for i in range(int(n)):
            for bit in rng.integers(0, 256, rng.integers(0, 40)):
                obs[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
        if n > 3:
            obs[2] = obs[1]
        obs_list.append(obs)
    best = matcher.ComputeDistinctiveDescriptors(obs_list)
    for o, b in zip(obs_list, best):
        assert b == O.distinctive_descriptor(o), (len(o), b)


@pytest.mark.parametrize('t,th,M', [(6, 3.0, 3000), (14, 3.0, 9000), (23, 5.0, 300), (9, 3.0, 0)])
def test_fuse_search(orb, t, th, M):
    """f2: the search of ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>, th) vs the oracle, which tests/test_ref_pins_oracle_cpu.py holds against the
    reference's own Fuse body (every logged AddObservation / Replace names the oracle's best keypoint)."""
    sc = O.fuse_scene(t, M=max(M, 1))
    if M == 0:
        for k in ('state', 'xyz', 'normal', 'min_d', 'max_d', 'mp_desc'):
            sc[k] = sc[k][:0]
    m = orb.ORBmatcher(0.6, True, max_batch=1, max_keypoints=2048, max_mappoints=64)       # M is not bounded by the handle
    bi, bd = m.FuseSearch(sc['kps'], sc['desc'], sc['bounds'], sc['sf'], sc['isg'], sc['log_sf'], sc['Tcw'], sc['Ow'], sc['cam'], sc['state'], sc['xyz'],
                          sc['normal'], sc['min_d'], sc['max_d'], sc['mp_desc'], th)
    if M == 0:
        assert len(bi) == 0
        return
    obi, obd = O.fuse_search(sc, th)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd)
    assert (bd <= 50).sum() > M // 20


@pytest.mark.parametrize('t,dts,coarse,ori', [(4, (3, -2, 5), False, True), (11, (2,), False, False), (17, (5, 1), True, True), (22, (), False, True)])
def test_search_for_triangulation(orb, t, dts, coarse, ori):
    """f2: ORBmatcher::SearchForTriangulation for one keyframe against several neighbours in one launch vs the oracle (which
    tests/test_ref_pins_oracle_cpu.py holds against the reference's own body + Pinhole::epipolarConstrain)."""
    scs = [O.triangulation_scene(t, dt) for dt in dts] or [O.triangulation_scene(t, 1)]
    kf1 = dict(kps=scs[0]['k1'], desc=scs[0]['d1'], has_mp=scs[0]['mp1'], fv=scs[0]['fv1'])
    for sc in scs:
        sc['mp1'] = scs[0]['mp1']                     # one KF1 for all pairs
    geo = [O.triangulation_geometry(sc) for sc in scs] if dts else []
    kf2 = [dict(kps=sc['k2'], desc=sc['d2'], has_mp=sc['mp2'], fv=sc['fv2']) for sc in scs] if dts else []
    m = orb.ORBmatcher(0.6, ori, max_batch=1, max_keypoints=2048, max_mappoints=2048)
    nm, m12 = m.SearchForTriangulation(kf1, kf2, scs[0]['sf'], scs[0]['sigma2'], np.array([g[0] for g in geo], np.float32).reshape(-1, 2),
                                       np.array([g[1] for g in geo], np.float32).reshape(-1, 9), coarse)
    assert len(nm) == len(dts)
    for k, sc in enumerate(scs if dts else []):
        on, om = O.search_for_triangulation(sc, geo[k][0], geo[k][1], coarse, ori)
        assert nm[k] == on and np.array_equal(m12[k], om), (k, nm[k], on)
        assert on > 30


@pytest.mark.parametrize('t,dt,k,L,levelsup,ratio,ori', [(3, 1, 10, 4, 2, 0.8, True), (8, 2, 10, 3, 2, 0.75, True), (15, 1, 6, 4, 3, 0.9, False), (21, 3, 10, 3, 1, 0.8, True)])
def test_search_by_bow_kf_kf(orb, t, dt, k, L, levelsup, ratio, ori):
    """f2: ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*) vs the oracle (= the reference's own body, tests/test_ref_pins_oracle_cpu.py)."""
    voc = O.synthetic_vocabulary(k, L, seed=k + L)
    rng = np.random.default_rng(t)
    k1, d1 = S.extract(t); k2, d2 = S.extract(t + dt)
    fv1 = O.bow_transform(voc, d1, levelsup)[2:]; fv2 = O.bow_transform(voc, d2, levelsup)[2:]
    p1 = rng.choice([0, 1, 1, 1, 1, 2], len(k1)).astype(np.uint8); p2 = rng.choice([0, 1, 1, 1, 1, 2], len(k2)).astype(np.uint8)
    m = orb.ORBmatcher(ratio, ori, max_batch=1, max_keypoints=2048, max_mappoints=2048)
    n, m12 = m.SearchByBoWKF(k1, d1, p1, fv1, k2, d2, p2, fv2)
    on, om = O.search_by_bow_kf(k1, d1, p1, fv1, k2, d2, p2, fv2, nnratio=ratio, check_ori=ori)
    assert n == on and np.array_equal(m12, om) and n > 20, (n, on)


@pytest.mark.parametrize('t,dt,th', [(5, 2, 7.5), (13, 3, 7.5), (20, 1, 4.0)])
def test_search_by_sim3(orb, t, dt, th):
    """f2: ORBmatcher::SearchBySim3 (both projection searches + agreement) vs the oracle (= the reference's own body, tests/test_ref_pins_oracle_cpu.py)."""
    sc = O.sim3_scene(t, dt)
    p21, p12 = O.sim3_camera_points(sc)
    side = lambda k, pc: dict(kps=sc['k%d' % k], desc=sc['d%d' % k], state=sc['state%d' % k], pcam=pc, min_d=sc['min%d' % k], max_d=sc['max%d' % k], mp_desc=sc['mpd%d' % k])
    m = orb.ORBmatcher(0.75, True, max_batch=1, max_keypoints=2048, max_mappoints=64)
    n, m12 = m.SearchBySim3(side(1, p21), side(2, p12), sc['bounds'], sc['sf'], sc['log_sf'], sc['cam'], sc['pre12'], th)
    on, om = O.search_by_sim3(sc, p21, p12, th)
    assert n == on and np.array_equal(m12, om) and n > 50, (n, on)


@pytest.mark.parametrize('t,th', [(6, 3.0), (14, 4.0)])
def test_fuse_search_sim3(orb, t, th):
    """f2: the search of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) vs the oracle."""
    sc = O.fuse_scene(t)
    m = orb.ORBmatcher(0.6, True, max_batch=1, max_keypoints=2048, max_mappoints=64)
    bi, bd = m.FuseSearchSim3(sc['kps'], sc['desc'], sc['bounds'], sc['sf'], sc['log_sf'], sc['Tcw'], sc['Ow'], sc['cam'], sc['state'], sc['xyz'], sc['normal'], sc['min_d'],
                              sc['max_d'], sc['mp_desc'], th)
    obi, obd = O.fuse_search_sim3(sc, sc['Tcw'], sc['Ow'], th)
    assert np.array_equal(bi, obi) and np.array_equal(bd, obd) and (bd <= 50).sum() > 100
    gbi, gbd = O.fuse_search(sc, th)
    assert (bd <= gbd).all() and (bd < gbd).any()          # without the chi-square gate the best candidate can only get closer
