"""Seeded matcher scenarios built from real extractor output of consecutive synthetic frames."""
import numpy as np

import oracle_lib as O
from orb_slam3_modified_b200 import synth

_CACHE = {}


def extract(t, seed=0, w=640, h=480, nf=1000):
    key = (t, seed, w, h, nf)
    if key not in _CACHE:
        _CACHE[key] = O.OracleExtractor(nf, 1.2, 8, 20, 7)(synth.frame(t, w, h, seed), (0, 1000))[1:]
    return _CACHE[key]


def last_frame_scene(t, seed=0, rng=None, pose_noise=0.004, frac_invalid=0.1, frac_noobs=0.1):
    """Current frame = frame t, last frame = frame t-1 whose keypoints became map points on the plane z=0."""
    rng = rng or np.random.default_rng(t * 31 + seed)
    kps, desc = extract(t, seed)
    lk, ld = extract(t - 1, seed)
    xyz = synth.backproject(np.stack([lk['x'], lk['y']], 1), t - 1, seed).astype(np.float32)
    M = len(lk)
    last = dict(valid=(rng.random(M) > frac_invalid).astype(np.uint8), xyz=xyz, octave=lk['octave'].astype(np.int32),
                angle=lk['angle'].astype(np.float32), hasObs=(rng.random(M) > frac_noobs).astype(np.uint8), descriptors=ld)
    T = synth.pose(t, seed)
    T[4:] += rng.normal(0, pose_noise, 3)
    q = T[:4] + rng.normal(0, pose_noise * 0.2, 4)
    T[:4] = q / np.linalg.norm(q)
    sf = O.OracleExtractor().tables()['scale']
    return dict(kps=kps, desc=desc, bounds=(0.0, 0.0, 640.0, 480.0), sf=sf, Tcw=T.astype(np.float32), cam=synth.camera(), last=last)


def local_map_scene(t, seed=0, rng=None, M=2500):
    """Local map = keypoints of several neighbouring frames projected into frame t (fields as Frame::isInFrustum writes them)."""
    rng = rng or np.random.default_rng(t * 17 + seed + 5)
    kps, desc = extract(t, seed)
    T = synth.pose(t, seed)
    w, x, y, z = T[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    cam = synth.camera()
    px, py, lv, dd, dep = [], [], [], [], []
    for dt in (-3, -2, -1, 1, 2):
        lk, ld = extract(t + dt, seed)
        P = synth.backproject(np.stack([lk['x'], lk['y']], 1), t + dt, seed)
        Pc = (R @ P.T).T + T[4:]
        u = cam[0] * Pc[:, 0] / Pc[:, 2] + cam[2] + rng.normal(0, 1.5, len(lk))
        v = cam[1] * Pc[:, 1] / Pc[:, 2] + cam[3] + rng.normal(0, 1.5, len(lk))
        px.append(u); py.append(v); lv.append(np.clip(lk['octave'] + rng.integers(-1, 2, len(lk)), 0, 7)); dd.append(ld); dep.append(Pc[:, 2])
    px, py, lv, dd, dep = (np.concatenate(a) for a in (px, py, lv, dd, dep))
    sel = rng.permutation(len(px))[:M]
    px, py, lv, dd, dep = px[sel], py[sel], lv[sel], dd[sel], dep[sel]
    inb = (px >= 0) & (px <= 640) & (py >= 0) & (py <= 480)
    n = len(px)
    pts = dict(inView=(inb & (rng.random(n) > 0.05)).astype(np.uint8), bad=(rng.random(n) < 0.03).astype(np.uint8), depth=dep.astype(np.float32),
               projX=px.astype(np.float32), projY=py.astype(np.float32), level=lv.astype(np.int32),
               viewCos=np.where(rng.random(n) < 0.5, 0.9995, 0.9).astype(np.float32), hasObs=(rng.random(n) > 0.1).astype(np.uint8), descriptors=dd)
    sf = O.OracleExtractor().tables()['scale']
    return dict(kps=kps, desc=desc, bounds=(0.0, 0.0, 640.0, 480.0), sf=sf, pts=pts)
