"""CPU tests: the oracle against the committed golden vectors (cv2-derived) and, when the cv2 wheel is
importable, against cv2 live on fresh random inputs.  These pin the OpenCV primitives the reference calls."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'primitives.npz'))


def test_resize_golden(gold):
    assert np.array_equal(O.resize_linear(gold['img'], 107, 80), gold['resize_img_107x80'])
    assert np.array_equal(O.resize_linear(gold['img'], 64, 48), gold['resize_img_64x48'])     # exact 2x -> INTER_AREA
    assert np.array_equal(O.resize_linear(gold['noise'], 69, 48), gold['resize_noise_69x48'])


def test_blur_golden(gold):
    assert np.array_equal(O.blur7(gold['img']), gold['blur_img'])
    assert np.array_equal(O.blur7(gold['noise']), gold['blur_noise'])


def test_fast_golden(gold):
    for name in ('img', 'noise'):
        for T in (20, 7):
            assert np.array_equal(O.fast(gold[name], T), gold['fast_%s_%d' % (name, T)]), (name, T)


def test_fastatan2_golden(gold):
    assert np.array_equal(O.fast_atan2(gold['atan_y'], gold['atan_x']), gold['atan_deg'])


def test_primitives_vs_cv2_live():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(5)
    from orb_slam3_modified_b200 import synth
    im = synth.frame(2)
    # full pyramid chain of the reference config
    prev = im
    inv = O.OracleExtractor().tables()['inv_scale']
    for l in range(1, 8):
        w, h = int(np.rint(np.float32(640) * inv[l])), int(np.rint(np.float32(480) * inv[l]))
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(O.resize_linear(prev, w, h), ref), l
        assert np.array_equal(O.blur7(ref), cv2.GaussianBlur(ref, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)), l
        prev = ref
    for it in range(25):
        w, h = int(rng.integers(7, 70)), int(rng.integers(4, 70))
        x, y = int(rng.integers(0, 640 - w)), int(rng.integers(0, 480 - h))
        roi = np.ascontiguousarray(im[y:y + h, x:x + w]) if it % 3 else rng.integers(0, 256, (h, w)).astype(np.uint8)
        for T in (20, 7):
            det = cv2.FastFeatureDetector_create(threshold=T, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in det.detect(roi)], np.int32).reshape(-1, 3)
            assert np.array_equal(O.fast(roi, T), ref), (w, h, T)


def test_sincosf_is_glibc():
    """The oracle calls the host libm; the CUDA path emulates glibc's FMA sincosf (tests/test_exact_math_cpu.py)."""
    a = np.linspace(0, 6.3, 1000).astype(np.float32)
    s, c = O.sincosf(a)
    assert np.allclose(s, np.sin(a.astype(np.float64)), atol=1e-7) and np.allclose(c, np.cos(a.astype(np.float64)), atol=1e-7)


@pytest.mark.parametrize('name', ['extract_640x480_t0', 'extract_640x480_t7', 'extract_1280x720_t5'])
def test_extract_golden(name):
    from orb_slam3_modified_b200 import synth
    rec = json.load(open(os.path.join(G, name + '.json')))
    im = synth.frame(rec['t'], rec['w'], rec['h'], rec['seed'])
    assert hashlib.sha256(im.tobytes()).hexdigest() == rec['image_sha256'], 'synthetic frame generator drifted'
    mono, kps, desc = O.OracleExtractor(rec['nf'], 1.2, 8, 20, 7)(im, tuple(rec['lap']))
    assert (mono, len(kps)) == (rec['mono'], rec['n'])
    assert hashlib.sha256(kps.tobytes()).hexdigest() == rec['kps_sha256']
    assert hashlib.sha256(desc.tobytes()).hexdigest() == rec['desc_sha256']


def test_extract_structure():
    """Properties the reference guarantees: per-level counts N_l..N_l+3, keypoints inside the 19-px margin,
    {0,1000} lapping on 640x480 -> everything filled from the back (monoIndex 0, octaves descending)."""
    from orb_slam3_modified_b200 import synth
    e = O.OracleExtractor()
    mono, kps, desc = e(synth.frame(4), (0, 1000))
    assert mono == 0
    assert list(kps['octave']) == sorted(kps['octave'], reverse=True)
    fpl = e.tables()['features_per_level']
    for l in range(8):
        n = int((kps['octave'] == l).sum())
        assert fpl[l] <= n <= fpl[l] + 3
        k = e.keypoints(l)
        lv = e.level(l)
        assert (k['x'] >= 19).all() and (k['x'] < lv.shape[1] - 19).all() and (k['y'] >= 19).all() and (k['y'] < lv.shape[0] - 19).all()
    assert desc.shape == (len(kps), 32)
