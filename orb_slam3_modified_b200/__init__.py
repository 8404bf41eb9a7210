"""B200-native ORB-SLAM3 hot path: Python mirror of the reference class surfaces over the C-ABI.

The compute lives in ``liborb_b200.so`` (hand-written sm_100a CUDA, ``csrc/``); this module only
marshals numpy / torch buffers to the ``extern "C"`` entry points declared in ``include/orb_b200.h``.
There is no CPU fallback: importing works anywhere, but creating any object without the built
library or without a Blackwell GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ORB_B200_LIB', os.path.join(_HERE, 'liborb_b200.so'))   # override: kernel-variant experiments

ORB_OK, ORB_ERR_EMPTY, ORB_ERR_ARG, ORB_ERR_CAPACITY, ORB_ERR_CUDA, ORB_ERR_ASPECT = 0, -1, -2, -3, -4, -5

#: numpy view of OrbKeyPoint / cv::KeyPoint (28 bytes)
KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'),
                     ('octave', 'i4'), ('class_id', 'i4')])

_lib = None


class OrbError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().orb_last_error().decode() if _lib is not None else ''
        super().__init__('%s failed with status %d: %s' % (where, code, msg))


def lib():
    """Load liborb_b200.so (built by ``__graft_entry__.build()``); raise if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('liborb_b200.so is not built (run `python -c "import __graft_entry__ as g; g.build()"`); '
                               'this package has no fallback path')
        L = C.CDLL(LIB_PATH)
        L.orb_last_error.restype = C.c_char_p
        vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.orbx_create.argtypes = [C.POINTER(vp), i, f, i, i, i, i, i, i, i]
        L.orbx_destroy.argtypes = [vp]
        L.orbx_destroy.restype = None
        L.orbx_get_levels.argtypes = [vp]
        L.orbx_get_tables.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orbx_max_keypoints.argtypes = [vp]
        L.orbx_extract.argtypes = [vp, vp, i, i, sz, i, i, vp, vp, i, vp, vp]
        L.orbx_extract_batch.argtypes = [vp, vp, i, i, i, sz, sz, i, i, vp, vp, i, vp, vp]
        L.orbx_extract_batch_device.argtypes = [vp, vp, i, i, i, sz, sz, i, i, vp, vp, i, vp, vp, vp]
        L.orbx_get_level_size.argtypes = [vp, i, vp, vp]
        L.orbx_copy_level.argtypes = [vp, i, i, i, vp]
        L.orbx_copy_candidates.argtypes = [vp, i, i, vp, i]
        L.orbx_last_launch_count.argtypes = [vp]
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


class ORBextractor:
    """Mirror of ``ORB_SLAM3::ORBextractor`` (reference include/ORBextractor.h:43-109).

    ``ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)``; calling the object is
    ``operator()``: ``mono_index, keypoints, descriptors = ex(image, vLappingArea)``.
    """

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 max_width=640, max_height=480, max_batch=1, device=0):
        L = lib()
        self._h = C.c_void_p()
        self.nlevels = nlevels
        self.max_batch = max_batch
        rc = L.orbx_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                           max_width, max_height, max_batch, device)
        if rc != ORB_OK:
            self._h = None
            raise OrbError(rc, 'orbx_create')
        self.max_keypoints = L.orbx_max_keypoints(self._h)

    def close(self):
        if getattr(self, '_h', None):
            lib().orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    # --- getters (include/ORBextractor.h:60-82) ---
    def GetLevels(self):
        return lib().orbx_get_levels(self._h)

    def _tables(self):
        n = self.nlevels
        s, i, g, ig = (np.zeros(n, np.float32) for _ in range(4))
        f = np.zeros(n, np.int32)
        lib().orbx_get_tables(self._h, _ptr(s), _ptr(i), _ptr(g), _ptr(ig), _ptr(f))
        return s, i, g, ig, f

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetScaleFactor(self):
        return float(self._tables()[0][1]) if self.nlevels > 1 else 1.0

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    # --- operator() ---
    def __call__(self, image, vLappingArea=(0, 0)):
        """Returns (monoIndex, keypoints[K] as KP_DTYPE, descriptors[K,32] u8); monoIndex == -1 for an empty image."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, 'CV_8UC1 expected (reference src/ORBextractor.cc:1094)'
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        cap = self.max_keypoints
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = lib().orbx_extract(self._h, _ptr(image), image.shape[0], image.shape[1], image.strides[0],
                                int(vLappingArea[0]), int(vLappingArea[1]), _ptr(kps), _ptr(desc), cap,
                                C.byref(n), C.byref(mono))
        if rc == ORB_ERR_EMPTY:
            return -1, kps[:0], desc[:0]
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_extract')
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0)):
        """images: [B, rows, cols] u8 host array.  Returns lists (mono, kps, desc) per frame."""
        images = np.ascontiguousarray(images, np.uint8)
        B, rows, cols = images.shape
        cap = self.max_keypoints
        kps = np.zeros((B, cap), KP_DTYPE)
        desc = np.zeros((B, cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        mono = np.zeros(B, np.int32)
        rc = lib().orbx_extract_batch(self._h, _ptr(images), B, rows, cols, images.strides[1], images.strides[0],
                                      int(vLappingArea[0]), int(vLappingArea[1]), _ptr(kps), _ptr(desc), cap,
                                      _ptr(n), _ptr(mono))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_extract_batch')
        return ([int(m) for m in mono], [kps[b, :n[b]].copy() for b in range(B)], [desc[b, :n[b]].copy() for b in range(B)])

    def extract_batch_slabs(self, images, kps, desc, n, mono, vLappingArea=(0, 0)):
        """Batched operator() into caller-owned fixed-capacity slabs (numpy, ideally pinned): kps [B, cap] KP_DTYPE,
        desc [B, cap, 32] u8, n / mono [B] i32 with cap == max_keypoints.  No per-frame copies."""
        B, rows, cols = images.shape
        assert kps.shape[1] == self.max_keypoints and images.strides[2] == 1
        rc = lib().orbx_extract_batch(self._h, _ptr(images), B, rows, cols, images.strides[1], images.strides[0],
                                      int(vLappingArea[0]), int(vLappingArea[1]), _ptr(kps), _ptr(desc), self.max_keypoints,
                                      _ptr(n), _ptr(mono))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_extract_batch')

    def resident_slabs(self):
        """Device addresses (ints) of the slabs the last host-buffer call left in the handle: (kps, desc, n, cap)."""
        k, d, n, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int()
        lib().orbx_resident_slabs.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        rc = lib().orbx_resident_slabs(self._h, C.byref(k), C.byref(d), C.byref(n), C.byref(cap))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_resident_slabs')
        return k.value, d.value, n.value, cap.value

    def extract_batch_device(self, d_images, d_kps, d_desc, d_n, d_mono, vLappingArea=(0, 0), stream=0):
        """All-device variant (torch CUDA tensors): d_images [B, rows, cols] u8; slabs d_kps [B, cap, 7] (28-byte rows),
        d_desc [B, cap, 32] u8, d_n / d_mono [B] i32.  Enqueues on ``stream`` and returns immediately."""
        B, rows, cols = d_images.shape
        cap = d_desc.shape[1]
        rc = lib().orbx_extract_batch_device(self._h, _ptr(d_images), B, rows, cols, d_images.stride(1), d_images.stride(0),
                                             int(vLappingArea[0]), int(vLappingArea[1]), _ptr(d_kps), _ptr(d_desc), cap,
                                             _ptr(d_n), _ptr(d_mono), C.c_void_p(stream))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_extract_batch_device')

    # --- inspection taps (mvImagePyramid is public in the reference, include/ORBextractor.h:84) ---
    def level(self, level, frame=0, blurred=False):
        w, h = C.c_int(), C.c_int()
        rc = lib().orbx_get_level_size(self._h, level, C.byref(w), C.byref(h))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_get_level_size')
        out = np.zeros((h.value, w.value), np.uint8)
        rc = lib().orbx_copy_level(self._h, frame, level, int(blurred), _ptr(out))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_copy_level')
        return out

    def level_bordered(self, level, frame=0, border=19):
        """mvImagePyramid[level] with the reflected frame ComputePyramid keeps around it (src/ORBextractor.cc:1185-1191)."""
        w, h = C.c_int(), C.c_int()
        lib().orbx_get_level_size(self._h, level, C.byref(w), C.byref(h))
        out = np.zeros((h.value + 2 * border, w.value + 2 * border), np.uint8)
        lib().orbx_copy_level_bordered.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        rc = lib().orbx_copy_level_bordered(self._h, frame, level, border, _ptr(out))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_copy_level_bordered')
        return out

    def ComputeStereoMatches(self, right, mb, mbf, batch=1):
        """``Frame::ComputeStereoMatches`` (src/Frame.cc:811-982) for the pairs both extractors processed in their last host-buffer call
        (self = left).  Returns (mvuRight, mvDepth) as [batch, cap] float32, -1 = no match."""
        cap = self.max_keypoints
        ur = np.zeros((batch, cap), np.float32); dep = np.zeros((batch, cap), np.float32)
        lib().orbx_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        rc = lib().orbx_stereo_matches(self._h, right._h, batch, mb, mbf, _ptr(ur), _ptr(dep), cap)
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_stereo_matches')
        return ur, dep

    def candidates(self, level, frame=0, cap=400000):
        out = np.zeros((cap, 3), np.int32)
        n = lib().orbx_copy_candidates(self._h, frame, level, _ptr(out), cap)
        if n < 0:
            raise OrbError(n, 'orbx_copy_candidates')
        return out[:n].copy()

    def last_launch_count(self):
        return lib().orbx_last_launch_count(self._h)

    STAGES = ('pyramid', 'blur', 'fast_cells', 'quadtree_orient', 'assemble', 'brief')

    def set_profiling(self, on):
        lib().orbx_set_profiling.argtypes = [C.c_void_p, C.c_int]
        lib().orbx_set_profiling(self._h, int(on))

    def stage_ms(self):
        lib().orbx_get_stage_ms.argtypes = [C.c_void_p, C.c_void_p]
        ms = np.zeros(6, np.float32)
        rc = lib().orbx_get_stage_ms(self._h, _ptr(ms))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbx_get_stage_ms')
        return dict(zip(self.STAGES, (float(m) for m in ms)))


# =============================================================================================
# ORBmatcher (reference include/ORBmatcher.h:36-103)
# =============================================================================================
class _OrbmFrame(C.Structure):
    _fields_ = [('K', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('minX', C.c_float), ('minY', C.c_float),
                ('maxX', C.c_float), ('maxY', C.c_float), ('scaleFactors', C.c_void_p), ('nlevels', C.c_int)]


class _OrbmLocalPoints(C.Structure):
    _fields_ = [('M', C.c_int)] + [(n, C.c_void_p) for n in ('inView', 'bad', 'depth', 'projX', 'projY', 'level', 'viewCos', 'hasObs', 'descriptors')]


class _OrbmFrustumIn(C.Structure):
    _fields_ = ([('M', C.c_int)] + [(n, C.c_void_p) for n in ('worldPos', 'normal', 'minDistInv', 'maxDistInv', 'maxDistance')] +
                [('Rcw', C.c_float * 9), ('tcw', C.c_float * 3), ('Ow', C.c_float * 3), ('cam', C.c_float * 4),
                 ('minX', C.c_float), ('minY', C.c_float), ('maxX', C.c_float), ('maxY', C.c_float), ('mbf', C.c_float),
                 ('logScaleFactor', C.c_float), ('nScaleLevels', C.c_int), ('viewingCosLimit', C.c_float)])


class _OrbmLastFrame(C.Structure):
    _fields_ = [('M', C.c_int)] + [(n, C.c_void_p) for n in ('valid', 'xyz', 'octave', 'angle', 'hasObs', 'descriptors')]


class _OrbmBatchDevice(C.Structure):
    _fields_ = [('batch', C.c_int), ('kcap', C.c_int), ('mcap', C.c_int), ('nlevels', C.c_int),
                ('kps', C.c_void_p), ('desc', C.c_void_p), ('nK', C.c_void_p),
                ('minX', C.c_float), ('minY', C.c_float), ('maxX', C.c_float), ('maxY', C.c_float),
                ('scaleFactors', C.c_void_p), ('nM', C.c_void_p),
                ('valid', C.c_void_p), ('xyz', C.c_void_p), ('octave', C.c_void_p), ('angle', C.c_void_p), ('hasObs', C.c_void_p),
                ('mpDesc', C.c_void_p), ('Tcw7', C.c_void_p), ('cam', C.c_float * 4), ('resetState', C.c_int)]


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class Frame:
    """The slice of ``ORB_SLAM3::Frame`` the matchers read: mvKeysUn, mDescriptors, image bounds, mvScaleFactors,
    and the in/out ``mvpMapPoints`` state as (match index, claimed flag) per keypoint."""

    def __init__(self, keypoints, descriptors, bounds, scale_factors):
        self.keypoints = _c(keypoints, KP_DTYPE)
        self.descriptors = _c(descriptors, np.uint8).reshape(-1, 32)
        self.bounds = tuple(float(b) for b in bounds)   # mnMinX, mnMinY, mnMaxX, mnMaxY
        self.scale_factors = _c(scale_factors, np.float32)
        self.match = np.full(len(self.keypoints), -1, np.int32)     # mvpMapPoints[i] as an index, -1 = NULL
        self.claimed = np.zeros(len(self.keypoints), np.uint8)      # mvpMapPoints[i]->Observations() > 0

    def _struct(self):
        return _OrbmFrame(len(self.keypoints), _ptr(self.keypoints), _ptr(self.descriptors), *self.bounds,
                          _ptr(self.scale_factors), len(self.scale_factors))


class ORBmatcher:
    """Mirror of ``ORB_SLAM3::ORBmatcher`` (reference include/ORBmatcher.h:36-103) for the per-frame hot functions."""
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30   # src/ORBmatcher.cc:35-37

    def __init__(self, nnratio=0.6, checkOri=True, max_batch=1, max_keypoints=2048, max_mappoints=8192, device=0):
        L = lib()
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.orbm_create.argtypes = [C.POINTER(vp), i, i, i, i]
        L.orbm_destroy.argtypes = [vp]
        L.orbm_destroy.restype = None
        L.orbm_search_local_map.argtypes = [vp, vp, vp, f, f, i, f, vp, vp, vp]
        L.orbm_search_last_frame.argtypes = [vp, vp, vp, vp, vp, f, i, vp, vp, vp]
        L.orbm_search_last_frame_batch_device.argtypes = [vp, vp, f, i, vp, vp, vp, vp]
        L.orbm_bf_knn2.argtypes = [vp, vp, i, vp, i, vp, vp]
        L.orbm_descriptor_distance.argtypes = [vp, vp, vp, i, vp]
        L.orbm_last_launch_count.argtypes = [vp]
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)
        self._h = C.c_void_p()
        rc = L.orbm_create(C.byref(self._h), max_batch, max_keypoints, max_mappoints, device)
        if rc != ORB_OK:
            self._h = None
            raise OrbError(rc, 'orbm_create')

    def close(self):
        if getattr(self, '_h', None):
            lib().orbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def DescriptorDistance(self, a, b):
        """ORBmatcher::DescriptorDistance for one pair or for n pairs ([n,32] arrays)."""
        a = _c(a, np.uint8).reshape(-1, 32)
        b = _c(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        rc = lib().orbm_descriptor_distance(self._h, _ptr(a), _ptr(b), len(a), _ptr(out))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_descriptor_distance')
        return int(out[0]) if len(out) == 1 else out

    def isInFrustum(self, pts, Rcw, tcw, Ow, cam, bounds, log_scale_factor, n_levels, viewingCosLimit=0.5, mbf=0.0):
        """``Frame::isInFrustum`` + ``MapPoint::PredictScale`` (src/Frame.cc:512-574, src/MapPoint.cc:531-546) for all local map
        points of one frame. ``pts``: worldPos [M,3], normal [M,3], minDistInv, maxDistInv, maxDistance [M]. Returns the tracking
        fields (inView, projX, projY, projXR, depth, level, viewCos) that the local-map SearchByProjection reads."""
        keep = {k: _c(pts[k], np.float32) for k in ('worldPos', 'normal', 'minDistInv', 'maxDistInv', 'maxDistance')}
        M = len(keep['minDistInv'])
        s = _OrbmFrustumIn(M, *[_ptr(keep[k]) for k in ('worldPos', 'normal', 'minDistInv', 'maxDistInv', 'maxDistance')],
                           (C.c_float * 9)(*[float(v) for v in np.asarray(Rcw, np.float32).reshape(9)]),
                           (C.c_float * 3)(*[float(v) for v in np.asarray(tcw, np.float32).reshape(3)]),
                           (C.c_float * 3)(*[float(v) for v in np.asarray(Ow, np.float32).reshape(3)]),
                           (C.c_float * 4)(*[float(v) for v in cam]), *[float(b) for b in bounds], float(mbf),
                           float(np.float32(log_scale_factor)), int(n_levels), float(viewingCosLimit))
        out = dict(inView=np.zeros(M, np.uint8), projX=np.zeros(M, np.float32), projY=np.zeros(M, np.float32), projXR=np.zeros(M, np.float32),
                   depth=np.zeros(M, np.float32), level=np.zeros(M, np.int32), viewCos=np.zeros(M, np.float32))
        L = lib()
        L.orbm_frustum_project.argtypes = [C.c_void_p] * 9
        rc = L.orbm_frustum_project(self._h, C.byref(s), *[_ptr(out[k]) for k in ('inView', 'projX', 'projY', 'projXR', 'depth', 'level', 'viewCos')])
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_frustum_project')
        return out

    def SearchByProjection(self, F, *args, **kw):
        """Overloads by argument type as in the reference: (F, map_points: dict, th, bFarPoints=False, thFarPoints=50)
        or (CurrentFrame, last_frame: dict, th, bMono, Tcw=..., cam=...)."""
        pts = args[0]
        if 'projX' in pts:
            return self._search_local_map(F, pts, *args[1:], **kw)
        return self._search_last_frame(F, pts, *args[1:], **kw)

    def _search_local_map(self, F, pts, th=1.0, bFarPoints=False, thFarPoints=50.0):
        M = len(pts['projX'])
        keep = dict(inView=_c(pts['inView'], np.uint8), bad=_c(pts['bad'], np.uint8), depth=_c(pts['depth'], np.float32),
                    projX=_c(pts['projX'], np.float32), projY=_c(pts['projY'], np.float32), level=_c(pts['level'], np.int32),
                    viewCos=_c(pts['viewCos'], np.float32), hasObs=_c(pts['hasObs'], np.uint8),
                    descriptors=_c(pts['descriptors'], np.uint8))
        s = _OrbmLocalPoints(M, *[_ptr(keep[k]) for k in ('inView', 'bad', 'depth', 'projX', 'projY', 'level', 'viewCos', 'hasObs', 'descriptors')])
        fs = F._struct()
        n = C.c_int(0)
        rc = lib().orbm_search_local_map(self._h, C.byref(fs), C.byref(s), th, self.mfNNratio, int(bFarPoints), thFarPoints,
                                         _ptr(F.match), _ptr(F.claimed), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_local_map')
        return n.value

    def _search_last_frame(self, F, last, th, bMono=True, Tcw=None, cam=None):
        assert bMono, 'only the monocular branch is on the hot path (SURVEY.md 8a row a10)'
        M = len(last['valid'])
        keep = dict(valid=_c(last['valid'], np.uint8), xyz=_c(last['xyz'], np.float32), octave=_c(last['octave'], np.int32),
                    angle=_c(last['angle'], np.float32), hasObs=_c(last['hasObs'], np.uint8), descriptors=_c(last['descriptors'], np.uint8))
        s = _OrbmLastFrame(M, *[_ptr(keep[k]) for k in ('valid', 'xyz', 'octave', 'angle', 'hasObs', 'descriptors')])
        T = _c(Tcw, np.float32)
        cm = _c(cam, np.float32)
        fs = F._struct()
        n = C.c_int(0)
        rc = lib().orbm_search_last_frame(self._h, C.byref(fs), C.byref(s), _ptr(T), _ptr(cm), th, int(self.mbCheckOrientation),
                                          _ptr(F.match), _ptr(F.claimed), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_last_frame')
        return n.value

    def search_last_frame_batch_device(self, d, th, d_match, d_claimed, d_nmatches, stream=0):
        """Device-resident batch (torch CUDA tensors in dict ``d``; see OrbmBatchDevice in include/orb_b200.h)."""
        s = _OrbmBatchDevice()
        s.batch, s.kcap, s.mcap, s.nlevels = d['batch'], d['kcap'], d['mcap'], d['nlevels']
        for k in ('kps', 'desc', 'nK', 'scaleFactors', 'nM', 'valid', 'xyz', 'octave', 'angle', 'hasObs', 'mpDesc', 'Tcw7'):
            setattr(s, k, d[k].data_ptr())
        s.minX, s.minY, s.maxX, s.maxY = d['bounds']
        s.cam = (C.c_float * 4)(*d['cam'])
        s.resetState = int(d.get('reset', 0))
        rc = lib().orbm_search_last_frame_batch_device(self._h, C.byref(s), th, int(self.mbCheckOrientation), _ptr(d_match),
                                                       _ptr(d_claimed), _ptr(d_nmatches), C.c_void_p(stream))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_last_frame_batch_device')

    def search_last_frame_batch(self, d, th, match, claimed, nmatches, resident=None):
        """Host-buffer batch (numpy arrays in dict ``d``, same keys as the device variant); in/out arrays are numpy.
        ``resident`` = ``ORBextractor.resident_slabs()`` of the extractor that just produced the frames: the current frame is then read
        from the device slabs (d['kps'], d['desc'], d['nK'] are ignored) instead of being uploaded again."""
        s = _OrbmBatchDevice()
        s.batch, s.kcap, s.mcap, s.nlevels = d['batch'], d['kcap'], d['mcap'], d['nlevels']
        for k in ('scaleFactors', 'nM', 'valid', 'xyz', 'octave', 'angle', 'hasObs', 'mpDesc', 'Tcw7'):
            setattr(s, k, d[k].ctypes.data)
        if resident is not None:
            s.kps, s.desc, s.nK = resident[0], resident[1], resident[2]
            assert resident[3] == d['kcap']
        else:
            for k in ('kps', 'desc', 'nK'):
                setattr(s, k, d[k].ctypes.data)
        s.minX, s.minY, s.maxX, s.maxY = d['bounds']
        s.cam = (C.c_float * 4)(*d['cam'])
        s.resetState = int(d.get('reset', 0))
        fn = lib().orbm_search_last_frame_batch_resident if resident is not None else lib().orbm_search_last_frame_batch
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = fn(self._h, C.byref(s), th, int(self.mbCheckOrientation), _ptr(match), _ptr(claimed), _ptr(nmatches))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_last_frame_batch')

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10):
        """``int ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)`` (src/ORBmatcher.cc:648-763).
        ``vbPrevMatched`` [K1, 2] float32 is updated in place; returns (nmatches, vnMatches12 [K1] i32, -1 = none)."""
        assert vbPrevMatched.dtype == np.float32 and vbPrevMatched.flags.c_contiguous and vbPrevMatched.shape == (len(F1.keypoints), 2)
        m12 = np.full(len(F1.keypoints), -1, np.int32)
        n = C.c_int(0)
        f1, f2 = F1._struct(), F2._struct()
        L = lib()
        L.orbm_search_for_initialization.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        rc = L.orbm_search_for_initialization(self._h, C.byref(f1), C.byref(f2), _ptr(vbPrevMatched), int(windowSize), self.mfNNratio,
                                              int(self.mbCheckOrientation), _ptr(m12), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_for_initialization')
        return n.value, m12

    def SearchByBoW(self, kf_kps, kf_desc, kf_point, kf_fv, f_kps, f_desc, f_fv):
        """``int ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)`` (src/ORBmatcher.cc:223-425).  ``*_fv`` = (node ids, feature indices) of
        the DBoW2 FeatureVector; ``kf_point`` [nKF]: 0 none / 1 map point / 2 bad.  Returns (nmatches, match [nF] = keyframe feature index or -1)."""
        class _BF(C.Structure):
            _fields_ = [('N', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('nEntries', C.c_int), ('fvNode', C.c_void_p), ('fvFeature', C.c_void_p)]
        keep = [_c(kf_kps, KP_DTYPE), _c(kf_desc, np.uint8), _c(kf_fv[0], np.int32), _c(kf_fv[1], np.int32), _c(f_kps, KP_DTYPE), _c(f_desc, np.uint8),
                _c(f_fv[0], np.int32), _c(f_fv[1], np.int32), _c(kf_point, np.uint8)]
        a = _BF(len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, len(keep[2]), keep[2].ctypes.data, keep[3].ctypes.data)
        b = _BF(len(keep[4]), keep[4].ctypes.data, keep[5].ctypes.data, len(keep[6]), keep[6].ctypes.data, keep[7].ctypes.data)
        match = np.full(len(keep[4]), -1, np.int32)
        n = C.c_int(0)
        L = lib()
        L.orbm_search_by_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        rc = L.orbm_search_by_bow(self._h, C.byref(a), _ptr(keep[8]), C.byref(b), self.mfNNratio, int(self.mbCheckOrientation), _ptr(match), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_by_bow')
        return n.value, match

    def FuseSearch(self, kf_kps, kf_desc, bounds, scale_factors, inv_level_sigma2, log_scale_factor, Tcw7, Ow, cam4, state, xyz, normal, min_d, max_d,
                   mp_desc, th=3.0):
        """The search of ``int ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th)`` (src/ORBmatcher.cc:1148-1338): (bestIdx [M], bestDist [M]);
        Fuse's return value is ``(bestDist <= 50).sum()``; see include/orb_b200.h: orbm_fuse_search."""
        class _Fr(C.Structure):
            _fields_ = [('K', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('minX', C.c_float), ('minY', C.c_float), ('maxX', C.c_float),
                        ('maxY', C.c_float), ('scaleFactors', C.c_void_p), ('nlevels', C.c_int)]
        class _Pt(C.Structure):
            _fields_ = [('M', C.c_int)] + [(n, C.c_void_p) for n in ('state', 'worldPos', 'normal', 'minDistance', 'maxDistance', 'descriptors')]
        k = [_c(kf_kps, KP_DTYPE), _c(kf_desc, np.uint8), _c(scale_factors, np.float32), _c(inv_level_sigma2, np.float32)]
        m = [_c(state, np.uint8), _c(xyz, np.float32), _c(normal, np.float32), _c(min_d, np.float32), _c(max_d, np.float32), _c(mp_desc, np.uint8)]
        o = [_c(Tcw7, np.float32), _c(Ow, np.float32), _c(cam4, np.float32)]
        fr = _Fr(len(k[0]), k[0].ctypes.data, k[1].ctypes.data, *[float(b) for b in bounds], k[2].ctypes.data, len(k[2]))
        pt = _Pt(len(m[0]), *[a.ctypes.data for a in m])
        bi = np.zeros(len(m[0]), np.int32); bd = np.zeros(len(m[0]), np.int32)
        L = lib()
        L.orbm_fuse_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        rc = L.orbm_fuse_search(self._h, C.byref(fr), _ptr(k[3]), float(log_scale_factor), _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), C.byref(pt), float(th), _ptr(bi), _ptr(bd))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_fuse_search')
        return bi, bd

    class _Fr(C.Structure):
        _fields_ = [('K', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('minX', C.c_float), ('minY', C.c_float), ('maxX', C.c_float),
                    ('maxY', C.c_float), ('scaleFactors', C.c_void_p), ('nlevels', C.c_int)]

    class _Pt(C.Structure):
        _fields_ = [('M', C.c_int)] + [(n, C.c_void_p) for n in ('state', 'worldPos', 'normal', 'minDistance', 'maxDistance', 'descriptors')]

    def FuseSearchSim3(self, kf_kps, kf_desc, bounds, scale_factors, log_scale_factor, Tcw7, Ow, cam4, state, xyz, normal, min_d, max_d, mp_desc, th=3.0):
        """The search of ``ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)`` (src/ORBmatcher.cc:1340-1455): (bestIdx, bestDist)."""
        k = [_c(kf_kps, KP_DTYPE), _c(kf_desc, np.uint8), _c(scale_factors, np.float32)]
        m = [_c(state, np.uint8), _c(xyz, np.float32), _c(normal, np.float32), _c(min_d, np.float32), _c(max_d, np.float32), _c(mp_desc, np.uint8)]
        o = [_c(Tcw7, np.float32), _c(Ow, np.float32), _c(cam4, np.float32)]
        fr = self._Fr(len(k[0]), k[0].ctypes.data, k[1].ctypes.data, *[float(b) for b in bounds], k[2].ctypes.data, len(k[2]))
        pt = self._Pt(len(m[0]), *[a.ctypes.data for a in m])
        bi = np.zeros(len(m[0]), np.int32); bd = np.zeros(len(m[0]), np.int32)
        L = lib()
        L.orbm_fuse_search_sim3.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        rc = L.orbm_fuse_search_sim3(self._h, C.byref(fr), float(log_scale_factor), _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), C.byref(pt), float(th), _ptr(bi), _ptr(bd))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_fuse_search_sim3')
        return bi, bd

    def SearchBySim3(self, kf1, kf2, bounds, scale_factors, log_scale_factor, cam4, pre12, th=7.5):
        """``int ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th)`` (src/ORBmatcher.cc:1457-1674).  kf = dict(kps, desc, state, pcam (the map points
        in the OTHER keyframe's camera frame), min_d, max_d, mp_desc).  Returns (nFound, match12)."""
        keep = []

        def side(kf):
            a = [_c(kf['kps'], KP_DTYPE), _c(kf['desc'], np.uint8), _c(scale_factors, np.float32), _c(kf['state'], np.uint8), _c(kf['pcam'], np.float32),
                 _c(kf['min_d'], np.float32), _c(kf['max_d'], np.float32), _c(kf['mp_desc'], np.uint8)]
            keep.append(a)
            fr = self._Fr(len(a[0]), a[0].ctypes.data, a[1].ctypes.data, *[float(b) for b in bounds], a[2].ctypes.data, len(a[2]))
            pt = self._Pt(len(a[3]), a[3].ctypes.data, a[4].ctypes.data, None, a[5].ctypes.data, a[6].ctypes.data, a[7].ctypes.data)
            return fr, pt
        f1, p1 = side(kf1); f2, p2 = side(kf2)
        cam = _c(cam4, np.float32); pre = _c(pre12, np.int32)
        m12 = np.zeros(f1.K, np.int32); n = C.c_int(0)
        L = lib()
        L.orbm_search_by_sim3.argtypes = [C.c_void_p] * 5 + [C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.orbm_search_by_sim3(self._h, C.byref(f1), C.byref(p1), C.byref(f2), C.byref(p2), float(log_scale_factor), _ptr(cam), float(th), _ptr(pre), _ptr(m12), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_by_sim3')
        return n.value, m12

    def SearchByBoWKF(self, k1, d1, point1, fv1, k2, d2, point2, fv2):
        """``int ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)`` (src/ORBmatcher.cc:765-905): (nmatches, match12 [N1] = KF2 feature or -1)."""
        class _BF(C.Structure):
            _fields_ = [('N', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('nEntries', C.c_int), ('fvNode', C.c_void_p), ('fvFeature', C.c_void_p)]
        keep = [_c(k1, KP_DTYPE), _c(d1, np.uint8), _c(fv1[0], np.int32), _c(fv1[1], np.int32), _c(k2, KP_DTYPE), _c(d2, np.uint8),
                _c(fv2[0], np.int32), _c(fv2[1], np.int32), _c(point1, np.uint8), _c(point2, np.uint8)]
        a = _BF(len(keep[0]), keep[0].ctypes.data, keep[1].ctypes.data, len(keep[2]), keep[2].ctypes.data, keep[3].ctypes.data)
        b = _BF(len(keep[4]), keep[4].ctypes.data, keep[5].ctypes.data, len(keep[6]), keep[6].ctypes.data, keep[7].ctypes.data)
        m12 = np.full(len(keep[0]), -1, np.int32)
        n = C.c_int(0)
        L = lib()
        L.orbm_search_by_bow_kf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        rc = L.orbm_search_by_bow_kf(self._h, C.byref(a), _ptr(keep[8]), C.byref(b), _ptr(keep[9]), self.mfNNratio, int(self.mbCheckOrientation), _ptr(m12), C.byref(n))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_by_bow_kf')
        return n.value, m12

    def SearchForTriangulation(self, kf1, kf2_list, scale_factors, level_sigma2, ep, F12, coarse=False):
        """``int ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, false, bCoarse)`` (src/ORBmatcher.cc:907-1146) for one keyframe against
        several neighbours.  A keyframe is a dict(kps, desc, has_mp, fv=(nodes, features)); ep [n, 2], F12 [n, 9].  Returns (nmatches [n], matches12 [n, N1])."""
        class _TF(C.Structure):
            _fields_ = [('N', C.c_int), ('keypoints', C.c_void_p), ('descriptors', C.c_void_p), ('hasMapPoint', C.c_void_p), ('nEntries', C.c_int),
                        ('fvNode', C.c_void_p), ('fvFeature', C.c_void_p)]
        keep = []

        def mk(kf):
            a = [_c(kf['kps'], KP_DTYPE), _c(kf['desc'], np.uint8), _c(kf['has_mp'], np.uint8), _c(kf['fv'][0], np.int32), _c(kf['fv'][1], np.int32)]
            keep.append(a)
            return _TF(len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[3]), a[3].ctypes.data, a[4].ctypes.data)
        f1 = mk(kf1)
        n2 = len(kf2_list)
        f2 = (_TF * max(n2, 1))(*[mk(k) for k in kf2_list])
        sf, sg = _c(scale_factors, np.float32), _c(level_sigma2, np.float32)
        epa, Fa = _c(ep, np.float32).reshape(-1, 2), _c(F12, np.float32).reshape(-1, 9)
        m12 = np.full((n2, f1.N), -1, np.int32); nm = np.zeros(n2, np.int32)
        L = lib()
        L.orbm_search_for_triangulation.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_void_p]
        rc = L.orbm_search_for_triangulation(self._h, C.byref(f1), n2, f2, _ptr(sf), _ptr(sg), len(sf), _ptr(epa), _ptr(Fa), int(coarse), int(self.mbCheckOrientation),
                                             _ptr(m12), _ptr(nm))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_search_for_triangulation')
        return nm, m12

    def ComputeDistinctiveDescriptors(self, obs_list):
        """``MapPoint::ComputeDistinctiveDescriptors`` for a list of map points (each an [n, 32] u8 array of observed descriptors): index of the chosen row."""
        start = np.zeros(len(obs_list) + 1, np.int32)
        for i, o in enumerate(obs_list):
            start[i + 1] = start[i] + len(o)
        desc = np.concatenate([_c(o, np.uint8).reshape(-1, 32) for o in obs_list]) if start[-1] else np.zeros((0, 32), np.uint8)
        desc = np.ascontiguousarray(desc)
        best = np.zeros(len(obs_list), np.int32)
        L = lib()
        L.orbm_distinctive_descriptors.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.orbm_distinctive_descriptors(self._h, len(obs_list), _ptr(start), _ptr(desc), _ptr(best))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_distinctive_descriptors')
        return best

    def knnMatch2(self, query, train):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2) (src/Frame.cc:1144): (idx[Q,2], dist[Q,2])."""
        q = _c(query, np.uint8).reshape(-1, 32)
        t = _c(train, np.uint8).reshape(-1, 32)
        idx = np.full((len(q), 2), -1, np.int32)
        dist = np.full((len(q), 2), -1, np.int32)
        rc = lib().orbm_bf_knn2(self._h, _ptr(q), len(q), _ptr(t), len(t), _ptr(idx), _ptr(dist))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbm_bf_knn2')
        return idx, dist

    def last_launch_count(self):
        return lib().orbm_last_launch_count(self._h)


# =============================================================================================
# Optimizer (reference include/Optimizer.h:46-102)
# =============================================================================================
class _LbaProblem(C.Structure):
    _fields_ = [('nPoses', C.c_int), ('poses', C.c_void_p), ('poseFixed', C.c_void_p), ('cam', C.c_void_p),
                ('nPoints', C.c_int), ('points', C.c_void_p), ('nEdges', C.c_int), ('edgePoint', C.c_void_p),
                ('edgePose', C.c_void_p), ('obs', C.c_void_p), ('invSigma2', C.c_void_p), ('huberDelta', C.c_double),
                ('iterations', C.c_int), ('userLambdaInit', C.c_double), ('stopFlag', C.c_void_p)]


class _LbaResult(C.Structure):
    _fields_ = [('poses', C.c_void_p), ('points', C.c_void_p), ('edgeChi2', C.c_void_p), ('edgeDepthPositive', C.c_void_p),
                ('iterations', C.c_int), ('trials', C.c_int), ('lambda_', C.c_double), ('chi2', C.c_double),
                ('initialChi2', C.c_double), ('gpuLaunches', C.c_int)]


class Optimizer:
    """Mirror of the static ``ORB_SLAM3::Optimizer`` functions on the hot path.  The pointer-graph walk of
    ``LocalBundleAdjustment`` (src/Optimizer.cc:1125-1403) stays with the caller; this object runs the numeric core
    on flat arrays (``synth.lba_problem`` has the layout) and returns what the write-back / outlier test reads."""

    def __init__(self, max_poses=64, max_points=8192, max_edges=65536, max_batch=1, device=0):
        L = lib()
        L.lba_create_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.lba_destroy.argtypes = [C.c_void_p]
        L.lba_destroy.restype = None
        L.lba_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lba_solve_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.lba_upload_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lba_run_batch_device.argtypes = [C.c_void_p, C.c_void_p]
        L.lba_download_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.lba_last_cluster_size.argtypes = [C.c_void_p]
        self._h = C.c_void_p()
        rc = L.lba_create_batch(C.byref(self._h), max_poses, max_points, max_edges, max_batch, device)
        if rc != ORB_OK:
            self._h = None
            raise OrbError(rc, 'lba_create')

    def close(self):
        if getattr(self, '_h', None):
            lib().lba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    @staticmethod
    def _pack(prob, iterations, user_lambda_init, stop_flag):
        keep = dict(poses=_c(prob['poses'], np.float64), fixed=_c(prob['fixed'], np.uint8), cam=_c(prob['cam'], np.float32),
                    points=_c(prob['points'], np.float64), ep=_c(prob['edge_point'], np.int32), ek=_c(prob['edge_pose'], np.int32),
                    obs=_c(prob['obs'], np.float64), isg=_c(prob['inv_sigma2'], np.float32))
        nP, nL, nE = len(keep['poses']), len(keep['points']), len(keep['ep'])
        p = _LbaProblem(nP, _ptr(keep['poses']), _ptr(keep['fixed']), _ptr(keep['cam']), nL, _ptr(keep['points']), nE, _ptr(keep['ep']),
                        _ptr(keep['ek']), _ptr(keep['obs']), _ptr(keep['isg']), float(prob['huber_delta']), iterations, user_lambda_init,
                        _ptr(stop_flag) if stop_flag is not None else None)
        out = dict(poses=np.zeros((nP, 7)), points=np.zeros((nL, 3)), chi2=np.zeros(nE), depth_pos=np.zeros(nE, np.uint8))
        r = _LbaResult(_ptr(out['poses']), _ptr(out['points']), _ptr(out['chi2']), _ptr(out['depth_pos']))
        return keep, p, out, r

    @staticmethod
    def _finish(out, r):
        out.update(iters=r.iterations, trials=r.trials, lambda_=r.lambda_, final_chi2=r.chi2, initial_chi2=r.initialChi2, launches=r.gpuLaunches)
        return out

    def LocalBundleAdjustmentBatch(self, probs, iterations=10, user_lambda_init=0.0):
        """One LocalBundleAdjustment per stream / local map, all solved by one kernel launch."""
        packs = [self._pack(p, iterations, user_lambda_init, None) for p in probs]
        P = (_LbaProblem * len(probs))(*[k[1] for k in packs])
        R = (_LbaResult * len(probs))(*[k[3] for k in packs])
        rc = lib().lba_solve_batch(self._h, len(probs), P, R)
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_solve_batch')
        return [self._finish(k[2], R[i]) for i, k in enumerate(packs)]

    def upload(self, probs, iterations=10, user_lambda_init=0.0):
        self._packs = [self._pack(p, iterations, user_lambda_init, None) for p in probs]
        P = (_LbaProblem * len(probs))(*[k[1] for k in self._packs])
        rc = lib().lba_upload_batch(self._h, len(probs), P)
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_upload_batch')

    def run_device(self, stream=0):
        rc = lib().lba_run_batch_device(self._h, C.c_void_p(stream))
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_run_batch_device')

    def download(self):
        R = (_LbaResult * len(self._packs))(*[k[3] for k in self._packs])
        rc = lib().lba_download_batch(self._h, len(self._packs), R)
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_download_batch')
        return [self._finish(k[2], R[i]) for i, k in enumerate(self._packs)]

    def last_cluster_size(self):
        return lib().lba_last_cluster_size(self._h)

    def set_cluster_size(self, ctas):
        lib().lba_set_cluster_size.argtypes = [C.c_void_p, C.c_int]
        rc = lib().lba_set_cluster_size(self._h, ctas)
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_set_cluster_size')

    def LocalBundleAdjustment(self, prob, iterations=10, user_lambda_init=0.0, stop_flag=None):
        keep = dict(poses=_c(prob['poses'], np.float64), fixed=_c(prob['fixed'], np.uint8), cam=_c(prob['cam'], np.float32),
                    points=_c(prob['points'], np.float64), ep=_c(prob['edge_point'], np.int32), ek=_c(prob['edge_pose'], np.int32),
                    obs=_c(prob['obs'], np.float64), isg=_c(prob['inv_sigma2'], np.float32))
        nP, nL, nE = len(keep['poses']), len(keep['points']), len(keep['ep'])
        p = _LbaProblem(nP, _ptr(keep['poses']), _ptr(keep['fixed']), _ptr(keep['cam']), nL, _ptr(keep['points']), nE, _ptr(keep['ep']),
                        _ptr(keep['ek']), _ptr(keep['obs']), _ptr(keep['isg']), float(prob['huber_delta']), iterations, user_lambda_init,
                        _ptr(stop_flag) if stop_flag is not None else None)
        out = dict(poses=np.zeros((nP, 7)), points=np.zeros((nL, 3)), chi2=np.zeros(nE), depth_pos=np.zeros(nE, np.uint8))
        r = _LbaResult(_ptr(out['poses']), _ptr(out['points']), _ptr(out['chi2']), _ptr(out['depth_pos']))
        rc = lib().lba_solve(self._h, C.byref(p), C.byref(r))
        if rc != ORB_OK:
            raise OrbError(rc, 'lba_solve')
        out.update(iters=r.iterations, trials=r.trials, lambda_=r.lambda_, final_chi2=r.chi2, initial_chi2=r.initialChi2, launches=r.gpuLaunches)
        return out


def PoseOptimization(frames, device=0):
    """``Optimizer::PoseOptimization`` for a list of frames (dicts: pose [7], cam [4], Xw [N,3], obs [N,2], inv_sigma2 [N]).
    Returns a list of dicts(pose, outlier, inliers)."""
    count = len(frames)
    cap = max(1, max(len(f['obs']) for f in frames))
    N = np.array([len(f['obs']) for f in frames], np.int32)
    pose = np.stack([_c(f['pose'], np.float64) for f in frames])
    cam = np.stack([_c(f['cam'], np.float32) for f in frames])
    Xw = np.zeros((count, cap, 3)); obs = np.zeros((count, cap, 2)); isg = np.ones((count, cap), np.float32)
    for i, f in enumerate(frames):
        Xw[i, :N[i]] = f['Xw']; obs[i, :N[i]] = f['obs']; isg[i, :N[i]] = f['inv_sigma2']
    out_pose = np.zeros((count, 7)); outl = np.zeros((count, cap), np.uint8); ninl = np.zeros(count, np.int32)
    L = lib()
    L.pose_optimization_batch.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6 + [C.c_double] + [C.c_void_p] * 3 + [C.c_int]
    rc = L.pose_optimization_batch(count, cap, _ptr(N), _ptr(pose), _ptr(cam), _ptr(Xw), _ptr(obs), _ptr(isg),
                                   float(np.float32(np.sqrt(5.991))), _ptr(out_pose), _ptr(outl), _ptr(ninl), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'pose_optimization_batch')
    return [dict(pose=out_pose[i], outlier=outl[i, :N[i]].copy(), inliers=int(ninl[i])) for i in range(count)]


# =============================================================================================
# Inertial edges (reference src/G2oTypes.cc, src/ImuTypes.cc)
# =============================================================================================
IMU_PREINT_FLOATS = 292


class _ImuMonoEdges(C.Structure):
    _fields_ = [('nPoses', C.c_int), ('poses', C.c_void_p), ('extrinsics', C.c_void_p), ('cam', C.c_void_p), ('nPoints', C.c_int), ('points', C.c_void_p),
                ('nEdges', C.c_int), ('edgePoint', C.c_void_p), ('edgePose', C.c_void_p), ('obs', C.c_void_p), ('invSigma2', C.c_void_p), ('huberDelta', C.c_double)]


def imu_preintegrate(acc, gyr, dt, n_meas, bias6, noise4, device=0):
    """``IMU::Preintegrated`` of `count` intervals: acc / gyr [count, maxMeas, 3], dt [count, maxMeas], n_meas [count], bias6 [count, 6]."""
    acc = _c(acc, np.float32); gyr = _c(gyr, np.float32); dt = _c(dt, np.float32); n = _c(n_meas, np.int32); b = _c(bias6, np.float32); nz = _c(noise4, np.float32)
    count, mm = dt.shape
    out = np.zeros((count, IMU_PREINT_FLOATS), np.float32)
    L = lib()
    L.imu_preintegrate_batch.argtypes = [C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int]
    rc = L.imu_preintegrate_batch(count, _ptr(n), mm, _ptr(acc), _ptr(gyr), _ptr(dt), _ptr(b), _ptr(nz), _ptr(out), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'imu_preintegrate_batch')
    return out


def imu_information(preint, device=0):
    P = _c(preint, np.float32).reshape(-1, IMU_PREINT_FLOATS)
    n = len(P)
    info = np.zeros((n, 9, 9)); ig = np.zeros((n, 3, 3)); ia = np.zeros((n, 3, 3))
    L = lib()
    L.imu_information_batch.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int]
    rc = L.imu_information_batch(n, _ptr(P), _ptr(info), _ptr(ig), _ptr(ia), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'imu_information_batch')
    return info, ig, ia


def imu_inertial_edges(preint, states36, info9=None, huber_delta=0.0, jac=True, device=0):
    """``EdgeInertial::computeError`` / ``linearizeOplus`` for n edges.  Returns dict(err [n,9], J [n,9,24] | None, chi2, rho)."""
    P = _c(preint, np.float32).reshape(-1, IMU_PREINT_FLOATS); S = _c(states36, np.float64).reshape(-1, 36)
    n = len(S)
    err = np.zeros((n, 9)); J = np.zeros((n, 9, 24)) if jac else None
    info = _c(info9, np.float64) if info9 is not None else None
    chi2 = np.zeros(n); rho = np.zeros(n)
    L = lib()
    L.imu_inertial_edges.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.imu_inertial_edges(n, _ptr(P), _ptr(S), _ptr(info), float(huber_delta), _ptr(err), _ptr(J), _ptr(chi2), _ptr(rho), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'imu_inertial_edges')
    return dict(err=err, J=J, chi2=chi2 if info is not None else None, rho=rho if info is not None else None)


def imu_mono_edges(poses12, extrinsics24, cam, points, edge_point, edge_pose, obs, inv_sigma2, huber_delta=0.0, jac=True, device=0):
    """``EdgeMono`` over ``ImuCamPose`` vertices.  Returns dict(err [n,2], Jpoint [n,2,3], Jpose [n,2,6], chi2, rho, depth_pos)."""
    po = _c(poses12, np.float64).reshape(-1, 12); ex = _c(extrinsics24, np.float64).reshape(24); cm = _c(cam, np.float32).reshape(-1, 4)
    pt = _c(points, np.float64).reshape(-1, 3); ep = _c(edge_point, np.int32); ek = _c(edge_pose, np.int32); ob = _c(obs, np.float64).reshape(-1, 2)
    isg = _c(inv_sigma2, np.float32)
    n = len(ep)
    s = _ImuMonoEdges(len(po), po.ctypes.data, ex.ctypes.data, cm.ctypes.data, len(pt), pt.ctypes.data, n, ep.ctypes.data, ek.ctypes.data, ob.ctypes.data,
                      isg.ctypes.data, float(huber_delta))
    err = np.zeros((n, 2)); Jp = np.zeros((n, 2, 3)) if jac else None; Jx = np.zeros((n, 2, 6)) if jac else None
    chi2 = np.zeros(n); rho = np.zeros(n); dp = np.zeros(n, np.uint8)
    L = lib()
    L.imu_mono_edges.argtypes = [C.c_void_p] * 7 + [C.c_int]
    rc = L.imu_mono_edges(C.byref(s), _ptr(err), _ptr(Jp), _ptr(Jx), _ptr(chi2), _ptr(rho), _ptr(dp), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'imu_mono_edges')
    return dict(err=err, Jpoint=Jp, Jpose=Jx, chi2=chi2, rho=rho, depth_pos=dp)


def PoseInertialOptimizationLastKeyFrame(frames, extrinsics24, rec_init=False, device=0):
    """``int Optimizer::PoseInertialOptimizationLastKeyFrame(Frame*, bool bRecInit)`` (src/Optimizer.cc:4491-4873) for a list of frames, one CTA each.
    frame = dict(Xw [N,3], obs [N,2], inv_sigma2 [N], track_depth [N], cam [4], preint [IMU_PREINT_FLOATS], kf_state [21], state [21]).
    Returns a list of dict(state [21], outlier [N], H [15,15], ret)."""
    n = len(frames)
    cap = max(1, max(len(f['Xw']) for f in frames))
    N = np.array([len(f['Xw']) for f in frames], np.int32)
    Xw = np.zeros((n, cap, 3), np.float32); ob = np.zeros((n, cap, 2), np.float32); isg = np.zeros((n, cap), np.float32); td = np.zeros((n, cap), np.float32)
    for i, f in enumerate(frames):
        Xw[i, :N[i]] = f['Xw']; ob[i, :N[i]] = f['obs']; isg[i, :N[i]] = f['inv_sigma2']; td[i, :N[i]] = f['track_depth']
    cam = np.stack([_c(f['cam'], np.float32) for f in frames]); P = np.stack([_c(f['preint'], np.float32) for f in frames])
    kf = np.stack([_c(f['kf_state'], np.float64) for f in frames]); st = np.stack([_c(f['state'], np.float64) for f in frames]).copy()
    ex = _c(extrinsics24, np.float64)
    out = np.zeros((n, cap), np.uint8); H = np.zeros((n, 15, 15)); ret = np.zeros(n, np.int32)
    L = lib()
    L.pose_inertial_optimization_last_kf_batch.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 10 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.pose_inertial_optimization_last_kf_batch(n, cap, _ptr(N), _ptr(Xw), _ptr(ob), _ptr(isg), _ptr(td), _ptr(cam), _ptr(ex), _ptr(P), _ptr(kf), _ptr(st), int(rec_init),
                                                    _ptr(out), _ptr(H), _ptr(ret), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'pose_inertial_optimization_last_kf_batch')
    return [dict(state=st[i], outlier=out[i, :N[i]], H=H[i], ret=int(ret[i])) for i in range(n)]


def PoseInertialOptimizationLastFrame(frames, extrinsics24, rec_init=False, device=0):
    """``int Optimizer::PoseInertialOptimizationLastFrame(Frame*, bool bRecInit)`` (src/Optimizer.cc:4875-5289) for a list of frames.
    frame = dict(Xw, obs, inv_sigma2, track_depth, cam, preint_frame, preint_kf, prior_state [21], prior_H [15,15], prev_state [21], state [21]).
    Returns a list of dict(state, prev_state, outlier, H [15,15], ret)."""
    n = len(frames)
    cap = max(1, max(len(f['Xw']) for f in frames))
    N = np.array([len(f['Xw']) for f in frames], np.int32)
    Xw = np.zeros((n, cap, 3), np.float32); ob = np.zeros((n, cap, 2), np.float32); isg = np.zeros((n, cap), np.float32); td = np.zeros((n, cap), np.float32)
    for i, f in enumerate(frames):
        Xw[i, :N[i]] = f['Xw']; ob[i, :N[i]] = f['obs']; isg[i, :N[i]] = f['inv_sigma2']; td[i, :N[i]] = f['track_depth']
    st64 = lambda k: np.stack([_c(f[k], np.float64).reshape(-1) for f in frames]).copy()
    cam = np.stack([_c(f['cam'], np.float32) for f in frames]); Pf = np.stack([_c(f['preint_frame'], np.float32) for f in frames]); Pk = np.stack([_c(f['preint_kf'], np.float32) for f in frames])
    prior, pH, pv, st = st64('prior_state'), st64('prior_H'), st64('prev_state'), st64('state')
    ex = _c(extrinsics24, np.float64)
    out = np.zeros((n, cap), np.uint8); H = np.zeros((n, 15, 15)); ret = np.zeros(n, np.int32)
    L = lib()
    L.pose_inertial_optimization_last_frame_batch.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 13 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.pose_inertial_optimization_last_frame_batch(n, cap, _ptr(N), _ptr(Xw), _ptr(ob), _ptr(isg), _ptr(td), _ptr(cam), _ptr(ex), _ptr(Pf), _ptr(Pk), _ptr(prior), _ptr(pH),
                                                       _ptr(pv), _ptr(st), int(rec_init), _ptr(out), _ptr(H), _ptr(ret), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'pose_inertial_optimization_last_frame_batch')
    return [dict(state=st[i], prev_state=pv[i], outlier=out[i, :N[i]], H=H[i], ret=int(ret[i])) for i in range(n)]


# =============================================================================================
# DBoW2 vocabulary transform (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h)
# =============================================================================================
class _LocalInertialBAProblem(C.Structure):
    _fields_ = [('nKF', C.c_int32), ('nOpt', C.c_int32), ('kfState21', C.c_void_p), ('kfTcw12', C.c_void_p), ('cam4', C.c_void_p), ('extrinsics24', C.c_void_p),
                ('nInertial', C.c_int32), ('ieKf1', C.c_void_p), ('ieKf2', C.c_void_p), ('preint', C.c_void_p), ('ieRobust', C.c_void_p), ('ieInfoScale', C.c_void_p),
                ('nPoints', C.c_int32), ('points3', C.c_void_p), ('trackDepth', C.c_void_p),
                ('nEdges', C.c_int32), ('edgePoint', C.c_void_p), ('edgeKf', C.c_void_p), ('obs2', C.c_void_p), ('invSigma2', C.c_void_p),
                ('iterations', C.c_int32), ('bLarge', C.c_int32), ('lambdaInit', C.c_double)]


class _LocalInertialBAResult(C.Structure):
    _fields_ = [('kfState21', C.c_void_p), ('kfTcw12', C.c_void_p), ('points3', C.c_void_p), ('erase', C.c_void_p), ('edgeChi2', C.c_void_p), ('stats8', C.c_void_p), ('profile8', C.c_void_p)]


def _liba_marshal(probs):
    """ctypes arrays of LocalInertialBAProblem / LocalInertialBAResult for a list of problem dicts (synth.local_inertial_ba_problem layout + 'preint')."""
    n = len(probs)
    P = (_LocalInertialBAProblem * n)(); R = (_LocalInertialBAResult * n)()
    keep, outs = [], []
    for i, pr in enumerate(probs):
        a = dict(st=_c(pr['state'], np.float64), tc=_c(pr['tcw'], np.float64), cam=_c(pr['cam'], np.float32), ex=_c(pr['extr'], np.float64),
                 k1=_c(pr['ie_kf1'], np.int32), k2=_c(pr['ie_kf2'], np.int32), pre=_c(pr['preint'], np.float32), rob=_c(pr['ie_robust'], np.uint8),
                 sc=_c(pr['ie_info_scale'], np.float64), pts=_c(pr['points'], np.float64), td=_c(pr['track_depth'], np.float32), ep=_c(pr['e_pt'], np.int32),
                 ek=_c(pr['e_kf'], np.int32), ob=_c(pr['obs'], np.float64), isg=_c(pr['inv_sigma2'], np.float32))
        nE = len(a['ep'])
        o = dict(state=np.zeros_like(a['st']), tcw=np.zeros_like(a['tc']), points=np.zeros_like(a['pts']), erase=np.zeros(nE, np.uint8), chi2=np.zeros(nE), stats=np.zeros(8), prof=np.zeros(8))
        P[i] = _LocalInertialBAProblem(int(pr['n_kf']), int(pr['n_opt']), a['st'].ctypes.data, a['tc'].ctypes.data, a['cam'].ctypes.data, a['ex'].ctypes.data, len(a['k1']), a['k1'].ctypes.data, a['k2'].ctypes.data,
                                       a['pre'].ctypes.data, a['rob'].ctypes.data, a['sc'].ctypes.data, len(a['pts']), a['pts'].ctypes.data, a['td'].ctypes.data, nE, a['ep'].ctypes.data, a['ek'].ctypes.data,
                                       a['ob'].ctypes.data, a['isg'].ctypes.data, int(pr['iterations']), int(bool(pr['large'])), float(pr['lambda_init']))
        R[i] = _LocalInertialBAResult(o['state'].ctypes.data, o['tcw'].ctypes.data, o['points'].ctypes.data, o['erase'].ctypes.data, o['chi2'].ctypes.data, o['stats'].ctypes.data, o['prof'].ctypes.data)
        keep.append(a); outs.append(o)
    return P, R, keep, outs


def _liba_finish(outs, iters):
    return [dict(state=o['state'], tcw=o['tcw'], points=o['points'], erase=o['erase'], chi2=o['chi2'], iters=int(iters[i]), err=float(o['stats'][0]),
                 err_end=float(o['stats'][1]), failed=bool(o['stats'][2]), lam=float(o['stats'][3]), trials=int(o['stats'][4]), kernel_ms=float(o['stats'][6]) * 1e-6, phase_ms=o['prof'] * 1e-6)
            for i, o in enumerate(outs)]


def LocalInertialBA(probs, device=0):
    """``void Optimizer::LocalInertialBA(KeyFrame*, bool*, Map*, int&, int&, int&, int&, bool bLarge, bool bRecInit)`` (src/Optimizer.cc:2383-2958): the
    numeric core for a list of local maps, one persistent CTA each (``local_inertial_ba_batch``).  A problem is the dict of
    ``synth.local_inertial_ba_problem`` plus ``preint`` [nI, IMU_PREINT_FLOATS]: n_kf, n_opt, state [nKF,21], tcw [nKF,12], cam [nKF,4], extr [24], ie_kf1 / ie_kf2 /
    ie_robust / ie_info_scale [nI], points [nL,3], track_depth [nL], e_pt / e_kf / obs / inv_sigma2 [nE], iterations, lambda_init, large.
    Returns a list of dict(state, tcw, points, erase, chi2, iters, err, err_end, failed, lam, trials)."""
    P, R, keep, outs = _liba_marshal(probs)
    iters = np.zeros(len(probs), np.int32)
    L = lib()
    L.local_inertial_ba_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.local_inertial_ba_batch(len(probs), C.cast(P, C.c_void_p), C.cast(R, C.c_void_p), _ptr(iters), device)
    if rc != ORB_OK:
        raise OrbError(rc, 'local_inertial_ba_batch')
    return _liba_finish(outs, iters)


class _OrbVocabulary(C.Structure):
    _fields_ = [('nNodes', C.c_int), ('L', C.c_int), ('weighting', C.c_int), ('norm', C.c_int), ('childStart', C.c_void_p), ('children', C.c_void_p),
                ('descriptors', C.c_void_p), ('weight', C.c_void_p), ('wordId', C.c_void_p)]


class ORBVocabulary:
    """Mirror of ``ORB_SLAM3::ORBVocabulary`` (= ``DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>``) for ``transform``: the tree is
    given as flat arrays (child_start, children, desc, weight, word_id) and stays on the device."""

    def __init__(self, L, child_start, children, desc, weight, word_id, weighting=0, norm=1, device=0):
        self._keep = [_c(child_start, np.int32), _c(children, np.int32), _c(desc, np.uint8), _c(weight, np.float64), _c(word_id, np.int32)]
        s = _OrbVocabulary(len(self._keep[4]), int(L), int(weighting), int(norm), *[a.ctypes.data for a in self._keep])
        self._h = C.c_void_p()
        lib().orbv_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
        rc = lib().orbv_create(C.byref(self._h), C.byref(s), device)
        if rc != ORB_OK:
            self._h = None
            raise OrbError(rc, 'orbv_create')

    def close(self):
        if getattr(self, '_h', None):
            lib().orbv_destroy.argtypes = [C.c_void_p]
            lib().orbv_destroy.restype = None
            lib().orbv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, desc_list, levelsup=4):
        """``transform(features, BowVector, FeatureVector, levelsup)`` for a list of frames (each [n, 32] u8).  Returns per frame
        (word_id [w] i32, word_value [w] f64, fv_node [e] i32, fv_feature [e] i32) in the reference's std::map order."""
        B = len(desc_list)
        cap = max(1, max(len(d) for d in desc_list))
        desc = np.zeros((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
        for b, d in enumerate(desc_list):
            n[b] = len(d); desc[b, :len(d)] = d
        wid = np.zeros((B, cap), np.int32); wv = np.zeros((B, cap)); nw = np.zeros(B, np.int32)
        fn = np.zeros((B, cap), np.int32); ff = np.zeros((B, cap), np.int32); ne = np.zeros(B, np.int32)
        L = lib()
        L.orbv_transform_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        rc = L.orbv_transform_batch(self._h, B, _ptr(desc), _ptr(n), cap, int(levelsup), _ptr(wid), _ptr(wv), _ptr(nw), _ptr(fn), _ptr(ff), _ptr(ne))
        if rc != ORB_OK:
            raise OrbError(rc, 'orbv_transform_batch')
        return [(wid[b, :nw[b]].copy(), wv[b, :nw[b]].copy(), fn[b, :ne[b]].copy(), ff[b, :ne[b]].copy()) for b in range(B)]
