"""Seeded local-map scenes for Frame::isInFrustum: points in front of / behind / beside the camera, inside and outside the
scale-invariance range, with normals at all angles."""
import numpy as np


def scene(M=3000, seed=0, W=640, H=480):
    rng = np.random.default_rng(seed)
    # camera pose: small random rotation (Rodrigues) + translation
    w = rng.normal(0, 0.2, 3)
    th = np.linalg.norm(w) + 1e-12
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K).astype(np.float32)
    t = rng.normal(0, 0.5, 3).astype(np.float32)
    Ow = (-R.T.astype(np.float64) @ t.astype(np.float64)).astype(np.float32)
    cam = (500.0 + 10 * rng.random(), 505.0, W / 2 - 3.0, H / 2 + 2.0)
    # points: most in a frustum-ish box in camera coordinates, some behind / far outside
    pc = np.stack([rng.uniform(-6, 6, M), rng.uniform(-4.5, 4.5, M), rng.uniform(-2, 12, M)], 1)
    P = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)      # world = R^T (pc - t)
    PO = P.astype(np.float64) - Ow.astype(np.float64)
    dist = np.linalg.norm(PO, axis=1)
    n = PO / dist[:, None] + rng.normal(0, 0.6, (M, 3))
    n /= np.linalg.norm(n, axis=1)[:, None]
    flip = rng.random(M) < 0.15
    n[flip] *= -1
    dmax = (dist * rng.uniform(0.6, 3.0, M)).astype(np.float32)                       # mfMaxDistance
    dmin = (dmax / np.float32(1.2 ** 7) * rng.uniform(0.5, 1.5, M)).astype(np.float32)
    pts = dict(worldPos=P, normal=n.astype(np.float32), minDistInv=np.float32(0.8) * dmin, maxDistInv=np.float32(1.2) * dmax, maxDistance=dmax,
               minDistance=dmin)   # raw mfMinDistance: only oracle/_ref (the reference's own getters) reads it
    return dict(pts=pts, Rcw=R, tcw=t, Ow=Ow, cam=cam, bounds=(0.0, 0.0, float(W), float(H)),
                log_scale_factor=np.float32(np.log(np.float32(1.2))), n_levels=8, mbf=40.0)


def keypoints_near(px, py, level, mp_desc, rng, n_levels=8):
    """A frame whose keypoints sit within ~2 px of the given projections at the predicted (or the next finer) level, with
    descriptors a few bit flips away from the map points'."""
    import orb_slam3_modified_b200 as orb
    K = len(px)
    kps = np.zeros(K, orb.KP_DTYPE)
    kps['x'] = np.clip(px + rng.normal(0, 1.5, K), 1, 638).astype(np.float32)
    kps['y'] = np.clip(py + rng.normal(0, 1.5, K), 1, 478).astype(np.float32)
    kps['octave'] = np.clip(level - (rng.random(K) < 0.3), 0, n_levels - 1)
    kps['size'] = 31.0 * 1.2 ** kps['octave']
    kps['angle'] = rng.uniform(0, 360, K).astype(np.float32)
    kps['response'] = 50
    kps['class_id'] = -1
    desc = mp_desc.copy()
    for i in range(K):
        bits = rng.integers(0, 256, rng.integers(0, 25))
        for b in bits:
            desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    order = rng.permutation(K)
    sf = (1.2 ** np.arange(n_levels)).astype(np.float32)
    return kps[order], desc[order], sf
