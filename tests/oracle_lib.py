"""ctypes loader for the CPU oracle (test infrastructure; never imported by the product)."""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, 'oracle', 'liborb_oracle.so')

KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'), ('octave', 'i4'), ('class_id', 'i4')])
assert KP_DTYPE.itemsize == 28

_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_ROOT, 'oracle', f) for f in os.listdir(os.path.join(_ROOT, 'oracle')) if f.endswith(('.cpp', '.h'))]
        if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
            subprocess.check_call(['make', '-s', '-C', os.path.join(_ROOT, 'oracle')])
        _lib = C.CDLL(_SO)
        _lib.orbo_create.restype = C.c_void_p
        _lib.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.orbo_destroy.argtypes = [C.c_void_p]
        _lib.orbo_fast_atan2.restype = C.c_float
        _lib.orbo_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """Mirror of ORBextractor (reference include/ORBextractor.h:43-109) over the oracle."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = C.c_void_p(self.L.orbo_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.orbo_destroy(self.h)
            self.h = None

    def __call__(self, img, lap=(0, 0)):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 4096
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self.L.orbo_extract(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], lap[0], lap[1], _p(kps), _p(desc), cap, C.byref(n))
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def tables(self):
        nl = self.nlevels
        s, i, g, ig = (np.zeros(nl, np.float32) for _ in range(4))
        f = np.zeros(nl, np.int32)
        u = np.zeros(16, np.int32)
        self.L.orbo_tables(self.h, _p(s), _p(i), _p(g), _p(ig), _p(f), _p(u))
        return dict(scale=s, inv_scale=i, sigma2=g, inv_sigma2=ig, features_per_level=f, umax=u)

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orbo_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.orbo_level_copy(self.h, l, _p(out))
        return out

    def candidates(self, l, cap=200000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.orbo_level_candidates(self.h, l, _p(out), cap)
        return out[:n].copy()

    def keypoints(self, l, cap=20000):
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.orbo_level_keypoints(self.h, l, _p(out), cap)
        return out[:n].copy()

    def distribute(self, cands, minX, maxX, minY, maxY, N):
        cands = np.ascontiguousarray(cands, KP_DTYPE)
        out = np.zeros(len(cands) + 8, KP_DTYPE)
        n = self.L.orbo_distribute(self.h, _p(cands), len(cands), minX, maxX, minY, maxY, N, _p(out), len(out))
        return out[:n].copy()


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orbo_resize_linear(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orbo_blur7(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), src.shape[1])
    return dst


def fast(roi, T):
    roi = np.ascontiguousarray(roi, np.uint8)
    cap = roi.size
    out = np.zeros((cap, 3), np.int32)
    n = lib().orbo_fast(_p(roi), roi.shape[1], roi.shape[0], roi.strides[0], T, _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().orbo_fast_atan2_n(_p(y), _p(x), _p(out), len(y))
    return out


def sincosf(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.zeros_like(a)
    c = np.zeros_like(a)
    lib().orbo_sincosf_n(_p(a), _p(s), _p(c), len(a))
    return s, c
