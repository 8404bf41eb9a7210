#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

* primitives.npz  -- outputs of the independent cv2 wheel (4.13.0 in the authoring container) for the five
  OpenCV primitives the reference calls: the oracle must reproduce them bit-for-bit.
* extract_*.json  -- SHA-256 digests + head/tail keypoints of the oracle's ORBextractor output on seeded
  synthetic frames (regression pin for the oracle itself and for the CUDA path at full size).
Run from the repo root: python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle_lib as O  # noqa: E402
from orb_slam3_modified_b200 import synth  # noqa: E402

out = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(2024)

img = synth.frame(0)[100:196, 200:328].copy()          # 96 x 128 textured crop
noise = rng.integers(0, 256, (57, 83)).astype(np.uint8)
g = {'img': img, 'noise': noise}
g['resize_img_107x80'] = cv2.resize(img, (107, 80), interpolation=cv2.INTER_LINEAR)
g['resize_img_64x48'] = cv2.resize(img, (64, 48), interpolation=cv2.INTER_LINEAR)      # exact 2x -> INTER_AREA path
g['resize_noise_69x48'] = cv2.resize(noise, (69, 48), interpolation=cv2.INTER_LINEAR)
g['blur_img'] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
g['blur_noise'] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
for name, im in (('img', img), ('noise', noise)):
    for T in (20, 7):
        det = cv2.FastFeatureDetector_create(threshold=T, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kps = det.detect(im)
        g['fast_%s_%d' % (name, T)] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
ay = rng.integers(-200000, 200000, 512).astype(np.float32)
ax = rng.integers(-200000, 200000, 512).astype(np.float32)
ay[:8] = 0
ax[4:12] = 0
g['atan_y'], g['atan_x'] = ay, ax
g['atan_deg'] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(ay, ax)], np.float32)
tq = rng.integers(0, 256, (40, 32)).astype(np.uint8)
tt = rng.integers(0, 256, (60, 32)).astype(np.uint8)
tt[7] = tt[3]
m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(tq, tt, k=2)
g['bf_q'], g['bf_t'] = tq, tt
g['bf_idx'] = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
g['bf_dist'] = np.array([[a.distance, b.distance] for a, b in m], np.int32)
np.savez_compressed(os.path.join(out, 'primitives.npz'), **g)

cases = [dict(name='extract_640x480_t0', t=0, w=640, h=480, seed=0, nf=1000, lap=(0, 1000)),
         dict(name='extract_640x480_t7', t=7, w=640, h=480, seed=1, nf=1000, lap=(0, 1000)),
         dict(name='extract_1280x720_t5', t=5, w=1280, h=720, seed=0, nf=1000, lap=(0, 1000))]
for c in cases:
    im = synth.frame(c['t'], c['w'], c['h'], c['seed'])
    mono, kps, desc = O.OracleExtractor(c['nf'], 1.2, 8, 20, 7)(im, c['lap'])
    rec = dict(c)
    rec.update(image_sha256=hashlib.sha256(im.tobytes()).hexdigest(), mono=int(mono), n=int(len(kps)),
               kps_sha256=hashlib.sha256(kps.tobytes()).hexdigest(), desc_sha256=hashlib.sha256(desc.tobytes()).hexdigest(),
               head=[[float(k['x']), float(k['y']), float(k['angle']), float(k['response']), int(k['octave'])] for k in kps[:8]],
               tail=[[float(k['x']), float(k['y']), float(k['angle']), float(k['response']), int(k['octave'])] for k in kps[-8:]],
               desc_head=desc[:2].tolist(), cv2_version=cv2.__version__)
    json.dump(rec, open(os.path.join(out, c['name'] + '.json'), 'w'), indent=1)
    print(c['name'], rec['n'], rec['mono'])
