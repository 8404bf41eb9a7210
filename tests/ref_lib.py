"""ctypes loader for oracle/_ref/libref_orb.so: the REFERENCE's own sources (src/ORBextractor.cc whole; line ranges of
src/ORBmatcher.cc, src/Frame.cc, src/MapPoint.cc, src/CameraModels/Pinhole.cpp) compiled where they lie under /root/reference
against the type stand-ins of oracle/ref_shim/ (recipe: oracle/Makefile).  Test infrastructure; used to pin the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import KP_DTYPE, _p, _c

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, 'oracle', '_ref', 'libref_orb.so')
_lib = None


def available():
    if not os.path.exists(_SO) and os.path.exists('/root/reference/src/ORBextractor.cc'):
        subprocess.check_call(['make', '-s', '-C', os.path.join(_ROOT, 'oracle')])
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if os.path.exists('/root/reference/src/ORBextractor.cc'):
            subprocess.check_call(['make', '-s', '-C', os.path.join(_ROOT, 'oracle')])   # rebuilds when the shim or the recipe changed
        _lib = C.CDLL(_SO)
        _lib.ref_orbx_create.restype = C.c_void_p
        _lib.ref_orbx_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.ref_orbx_destroy.argtypes = [C.c_void_p]
        _lib.ref_ic_angle.restype = C.c_float
        _lib.ref_ic_angle.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        _lib.ref_orb_descriptor.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
    return _lib


class RefExtractor:
    """ORB_SLAM3::ORBextractor itself (reference include/ORBextractor.h:43-109)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.ref_orbx_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.ref_orbx_destroy(self.h)
            self.h = None

    def __call__(self, img, lap=(0, 0)):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 4096
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        rows, cols = (img.shape[0], img.shape[1]) if img.size else (0, 0)
        mono = self.L.ref_orbx_extract(self.h, _p(img), rows, cols, img.strides[0] if img.size else 0, lap[0], lap[1], _p(kps), _p(desc), cap, C.byref(n))
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def tables(self):
        nl = self.nlevels
        s, i, g, ig = (np.zeros(nl, np.float32) for _ in range(4))
        f = np.zeros(nl, np.int32)
        u = np.zeros(16, np.int32)
        self.L.ref_orbx_tables(self.h, _p(s), _p(i), _p(g), _p(ig), _p(f), _p(u))
        return dict(scale=s, inv_scale=i, sigma2=g, inv_sigma2=ig, features_per_level=f, umax=u)

    def level(self, l, border=0):
        w, h = C.c_int(), C.c_int()
        self.L.ref_orbx_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value + 2 * border, w.value + 2 * border), np.uint8)
        self.L.ref_orbx_level_copy(self.h, l, border, _p(out))
        return out

    def distribute(self, cands, minX, maxX, minY, maxY, N):
        cands = np.ascontiguousarray(cands, KP_DTYPE)
        out = np.zeros(len(cands) + 8, KP_DTYPE)
        n = self.L.ref_orbx_distribute(self.h, _p(cands), len(cands), minX, maxX, minY, maxY, N, _p(out), len(out))
        return out[:n].copy()

    def ic_angle(self, img, x, y):
        img = np.ascontiguousarray(img, np.uint8)
        return self.L.ref_ic_angle(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], x, y)

    def descriptor(self, img, x, y, angle):
        img = np.ascontiguousarray(img, np.uint8)
        d = np.zeros(32, np.uint8)
        self.L.ref_orb_descriptor(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], x, y, angle, _p(d))
        return d


def descriptor_distance(a, b):
    a = _c(a, np.uint8); b = _c(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


def compute_three_maxima(sizes):
    s = _c(sizes, np.int32)
    out = np.zeros(3, np.int32)
    lib().ref_compute_three_maxima(_p(s), len(s), _p(out))
    return tuple(int(v) for v in out)


def features_in_area(kps, bounds, x, y, r, min_level, max_level):
    kps = _c(kps, KP_DTYPE); b = _c(bounds, np.float32)
    out = np.zeros(len(kps) + 1, np.int32)
    L = lib()
    L.ref_features_in_area.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = L.ref_features_in_area(len(kps), _p(kps), _p(b), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n].copy()


def _pad_sf(sf):
    out = np.ones(64, np.float32)     # the wrapper copies 64 entries of mvScaleFactors
    out[:len(sf)] = sf
    return out


def search_local_map(kps, desc, bounds, scale_factors, pts, th, nnratio, b_far, th_far, match, claimed):
    kps = _c(kps, KP_DTYPE); desc = _c(desc, np.uint8); b = _c(bounds, np.float32); sf = _pad_sf(_c(scale_factors, np.float32))
    a = {k: _c(pts[k], dt) for k, dt in (('inView', np.uint8), ('bad', np.uint8), ('depth', np.float32), ('projX', np.float32),
                                          ('projY', np.float32), ('level', np.int32), ('viewCos', np.float32), ('hasObs', np.uint8),
                                          ('descriptors', np.uint8))}
    L = lib()
    L.ref_search_local_map.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 9 + [C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    return L.ref_search_local_map(len(kps), _p(kps), _p(desc), _p(b), _p(sf), len(a['projX']), _p(a['inView']), _p(a['bad']), _p(a['depth']),
                                  _p(a['projX']), _p(a['projY']), _p(a['level']), _p(a['viewCos']), _p(a['hasObs']), _p(a['descriptors']),
                                  th, nnratio, int(b_far), th_far, _p(match), _p(claimed))


def search_last_frame(kps, desc, bounds, scale_factors, Tcw, cam, last, th, check_ori, match, claimed):
    kps = _c(kps, KP_DTYPE); desc = _c(desc, np.uint8); b = _c(bounds, np.float32); sf = _pad_sf(_c(scale_factors, np.float32))
    T = _c(Tcw, np.float32); cm = _c(cam, np.float32)
    a = {k: _c(last[k], dt) for k, dt in (('valid', np.uint8), ('xyz', np.float32), ('octave', np.int32), ('angle', np.float32),
                                           ('hasObs', np.uint8), ('descriptors', np.uint8))}
    L = lib()
    L.ref_search_last_frame.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    return L.ref_search_last_frame(len(kps), _p(kps), _p(desc), _p(b), _p(sf), _p(T), _p(cm), len(a['valid']), _p(a['valid']), _p(a['xyz']),
                                   _p(a['octave']), _p(a['angle']), _p(a['hasObs']), _p(a['descriptors']), th, int(check_ori), _p(match), _p(claimed))


def search_for_initialization(kps1, desc1, kps2, desc2, bounds, scale_factors, prev_matched, window=100, nnratio=0.9, check_ori=True):
    kps1 = _c(kps1, KP_DTYPE); kps2 = _c(kps2, KP_DTYPE); d1 = _c(desc1, np.uint8); d2 = _c(desc2, np.uint8)
    b = _c(bounds, np.float32); sf = _pad_sf(_c(scale_factors, np.float32))
    pm = _c(prev_matched, np.float32).copy()
    m12 = np.full(len(kps1), -1, np.int32)
    L = lib()
    L.ref_search_for_initialization.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = L.ref_search_for_initialization(len(kps1), _p(kps1), _p(d1), len(kps2), _p(kps2), _p(d2), _p(b), _p(sf), _p(pm), int(window), nnratio,
                                        int(check_ori), _p(m12))
    return n, m12, pm


def is_in_frustum(pts, Rcw, tcw, Ow, cam, bounds, log_scale_factor, n_levels, viewing_cos_limit=0.5, mbf=0.0):
    a = {k: _c(pts[k], np.float32) for k in ('worldPos', 'normal', 'minDistance', 'maxDistance')}
    M = len(a['minDistance'])
    R = _c(np.asarray(Rcw, np.float32).reshape(9), np.float32); t = _c(tcw, np.float32); o = _c(Ow, np.float32)
    cm = _c(cam, np.float32); b = _c(bounds, np.float32)
    out = dict(inView=np.zeros(M, np.uint8), projX=np.zeros(M, np.float32), projY=np.zeros(M, np.float32), projXR=np.zeros(M, np.float32),
               depth=np.zeros(M, np.float32), level=np.zeros(M, np.int32), viewCos=np.zeros(M, np.float32))
    L = lib()
    L.ref_is_in_frustum.restype = None
    L.ref_is_in_frustum.argtypes = [C.c_int] + [C.c_void_p] * 9 + [C.c_float, C.c_float, C.c_int, C.c_float] + [C.c_void_p] * 7
    L.ref_is_in_frustum(M, _p(a['worldPos']), _p(a['normal']), _p(a['minDistance']), _p(a['maxDistance']), _p(R), _p(t), _p(o),
                        _p(cm), _p(b), float(mbf), float(np.float32(log_scale_factor)), int(n_levels), float(viewing_cos_limit),
                        *[_p(out[k]) for k in ('inView', 'projX', 'projY', 'projXR', 'depth', 'level', 'viewCos')])
    return out


def stereo_matches(ex_left, ex_right, kl, dl, kr, dr, scale, inv_scale, mb, mbf):
    """Frame::ComputeStereoMatches on two RefExtractor objects (each after its own call on the left / right image)."""
    kl = _c(kl, KP_DTYPE); kr = _c(kr, KP_DTYPE); dl = _c(dl, np.uint8); dr = _c(dr, np.uint8)
    sf = _c(scale, np.float32); isf = _c(inv_scale, np.float32)
    ur = np.zeros(len(kl), np.float32); dep = np.zeros(len(kl), np.float32)
    L = lib()
    L.ref_stereo_matches.restype = None
    L.ref_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    L.ref_stereo_matches(ex_left.h, ex_right.h, len(kl), _p(kl), _p(dl), len(kr), _p(kr), _p(dr), _p(sf), _p(isf), len(sf), mb, mbf, _p(ur), _p(dep))
    return ur, dep


class RefVocabulary:
    """DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> itself (the reference's ORBVocabulary), filled from flat arrays."""

    def __init__(self, voc, weighting=0, scoring=0):
        self.L = lib()
        self.L.ref_bow_create.restype = C.c_void_p
        self.L.ref_bow_create.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        self.h = C.c_void_p(self.L.ref_bow_create(voc['k'], voc['L'], weighting, scoring, len(voc['parent']), _p(voc['parent']), _p(_c(voc['desc'], np.uint8)),
                                                  _p(_c(voc['weight'], np.float64))))

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.ref_bow_destroy.argtypes = [C.c_void_p]
            self.L.ref_bow_destroy(self.h)
            self.h = None

    def transform(self, feats, levelsup=4):
        feats = _c(feats, np.uint8).reshape(-1, 32)
        N = len(feats)
        ow = np.zeros(max(N, 1), np.int32); ov = np.zeros(max(N, 1)); fn = np.zeros(max(N, 1), np.int32); ff = np.zeros(max(N, 1), np.int32); nf = C.c_int(0)
        self.L.ref_bow_transform.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        n = self.L.ref_bow_transform(self.h, N, _p(feats), levelsup, _p(ow), _p(ov), len(ow), _p(fn), _p(ff), C.byref(nf))
        return ow[:n].copy(), ov[:n].copy(), fn[:nf.value].copy(), ff[:nf.value].copy()

    def score(self, a, b):
        i1, v1 = _c(a[0], np.int32), _c(a[1], np.float64); i2, v2 = _c(b[0], np.int32), _c(b[1], np.float64)
        self.L.ref_bow_score.restype = C.c_double
        self.L.ref_bow_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        return self.L.ref_bow_score(self.h, len(i1), _p(i1), _p(v1), len(i2), _p(i2), _p(v2))


def search_by_bow(*a, **kw):
    import oracle_lib as O
    return O.search_by_bow(*a, _lib=lib(), _name='ref_search_by_bow', **kw)


def distinctive_descriptor(desc):
    import oracle_lib as O
    return O.distinctive_descriptor(desc, _lib=lib(), _name='ref_distinctive_descriptor')


def fuse(sc, th=3.0):
    """ORBmatcher::Fuse itself; returns (nFused, action [M], action_idx [M]) -- see ref_wrap_matcher.cpp: ref_fuse."""
    M = len(sc['state'])
    act = np.zeros(M, np.int32); idx = np.zeros(M, np.int32)
    L = lib()
    L.ref_fuse.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_float] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_void_p, C.c_void_p]
    a = [_c(sc['kps'], KP_DTYPE), _c(sc['desc'], np.uint8), _c(sc['bounds'], np.float32), _c(sc['sf'], np.float32), _c(sc['isg'], np.float32)]
    b = [_c(sc['Tcw'], np.float32), _c(sc['Ow'], np.float32), _c(sc['cam'], np.float32), _c(sc['kf_point'], np.int32), _c(sc['kf_point_bad'], np.uint8)]
    c = [_c(sc['state'], np.uint8), _c(sc['xyz'], np.float32), _c(sc['normal'], np.float32), _c(sc['min_d'], np.float32), _c(sc['max_d'], np.float32), _c(sc['mp_desc'], np.uint8),
         _c(sc['mp_obs'], np.int32)]
    n = L.ref_fuse(len(a[0]), *[_p(v) for v in a], len(a[3]), float(sc['log_sf']), *[_p(v) for v in b], M, *[_p(v) for v in c], th, _p(act), _p(idx))
    return n, act, idx


def search_for_triangulation(sc, coarse=False, check_ori=True):
    """ORBmatcher::SearchForTriangulation itself (+ Pinhole::epipolarConstrain); returns (nmatches, matches12, ep [2], F12 [9]) -- the epipole and
    fundamental matrix are the ones the body worked with (ref_wrap_matcher.cpp)."""
    a = [_c(sc['k1'], KP_DTYPE), _c(sc['d1'], np.uint8), _c(sc['mp1'], np.uint8), _c(sc['fv1'][0], np.int32), _c(sc['fv1'][1], np.int32)]
    b = [_c(sc['k2'], KP_DTYPE), _c(sc['d2'], np.uint8), _c(sc['mp2'], np.uint8), _c(sc['fv2'][0], np.int32), _c(sc['fv2'][1], np.int32)]
    c = [_c(sc['sf'], np.float32), _c(sc['sigma2'], np.float32)]
    d = [_c(sc['T1w'], np.float32), _c(sc['T2w'], np.float32), _c(sc['cam'], np.float32), _c(sc['cam'], np.float32)]
    m12 = np.full(len(a[0]), -1, np.int32); ep = np.zeros(2, np.float32); F12 = np.zeros(9, np.float32)
    L = lib()
    L.ref_search_for_triangulation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] * 2 + [C.c_void_p, C.c_void_p, C.c_int] + \
        [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.ref_search_for_triangulation(len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]),
                                       len(b[0]), _p(b[0]), _p(b[1]), _p(b[2]), len(b[3]), _p(b[3]), _p(b[4]), _p(c[0]), _p(c[1]), len(c[0]),
                                       *[_p(v) for v in d], int(coarse), int(check_ori), _p(m12), _p(ep), _p(F12))
    return n, m12, ep, F12


def search_by_bow_kf(*a, **kw):
    import oracle_lib as O
    return O.search_by_bow_kf(*a, _lib=lib(), _name='ref_search_by_bow_kf', **kw)


def search_by_sim3(sc, th=7.5):
    """ORBmatcher::SearchBySim3 itself: (nFound, match12, pc2of1, pc1of2) -- see ref_wrap_matcher.cpp."""
    L = lib()
    N1, N2 = len(sc['k1']), len(sc['k2'])
    m12 = np.zeros(N1, np.int32); p21 = np.zeros((N1, 3), np.float32); p12 = np.zeros((N2, 3), np.float32)
    L.ref_search_by_sim3.argtypes = [C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p] + ([C.c_int] + [C.c_void_p] * 8) * 2 + [C.c_void_p, C.c_float] + [C.c_void_p] * 4
    g = [_c(sc['sf'], np.float32), _c(sc['bounds'], np.float32), _c(sc['cam'], np.float32)]
    side = lambda k: [_c(sc['k%d' % k], KP_DTYPE), _c(sc['d%d' % k], np.uint8), _c(sc['T%dw' % k], np.float32), _c(sc['state%d' % k], np.uint8), _c(sc['xyz%d' % k], np.float32),
                      _c(sc['min%d' % k], np.float32), _c(sc['max%d' % k], np.float32), _c(sc['mpd%d' % k], np.uint8)]
    a, b = side(1), side(2)
    S = _c(sc['S12'], np.float32); pre = _c(sc['pre12'], np.int32)
    n = L.ref_search_by_sim3(len(g[0]), _p(g[0]), float(sc['log_sf']), _p(g[1]), _p(g[2]), N1, *[_p(v) for v in a], N2, *[_p(v) for v in b], _p(S), th, _p(pre), _p(m12), _p(p21), _p(p12))
    return n, m12, p21, p12


def fuse_sim3(sc, Scw, kf_state, th=3.0):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) itself: (nFused, action, action_idx, Tcw7, Ow)."""
    L = lib()
    M = len(sc['state'])
    act = np.zeros(M, np.int32); idx = np.zeros(M, np.int32); T7 = np.zeros(7, np.float32); Ow = np.zeros(3, np.float32)
    L.ref_fuse_sim3.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_float] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float] + [C.c_void_p] * 4
    a = [_c(sc['kps'], KP_DTYPE), _c(sc['desc'], np.uint8), _c(sc['bounds'], np.float32), _c(sc['sf'], np.float32)]
    b = [_c(sc['cam'], np.float32), _c(kf_state, np.uint8), _c(Scw, np.float32)]
    c = [_c(sc['state'], np.uint8), _c(sc['xyz'], np.float32), _c(sc['normal'], np.float32), _c(sc['min_d'], np.float32), _c(sc['max_d'], np.float32), _c(sc['mp_desc'], np.uint8)]
    n = L.ref_fuse_sim3(len(a[0]), *[_p(v) for v in a], len(a[3]), float(sc['log_sf']), *[_p(v) for v in b], M, *[_p(v) for v in c], th, _p(act), _p(idx), _p(T7), _p(Ow))
    return n, act, idx, T7, Ow
