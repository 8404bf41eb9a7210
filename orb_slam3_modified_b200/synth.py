"""Seeded synthetic inputs: textured frames for the extractor.

numpy only, deterministic for a given seed; used by tests/, bench.py and smoke().
"""
import numpy as np

_WORLD_CACHE = {}


def world_plane(seed=0, size=2048, nshapes=40000):
    """Grey-level plane of random rectangles/discs at log-uniform sizes, lightly blurred + noise."""
    key = (seed, size, nshapes)
    if key in _WORLD_CACHE:
        return _WORLD_CACHE[key]
    rng = np.random.default_rng(seed)
    img = np.full((size, size), 128.0, np.float32)
    sizes = np.exp(rng.uniform(np.log(2), np.log(100), nshapes)).astype(np.int32)
    xs = rng.integers(0, size, nshapes)
    ys = rng.integers(0, size, nshapes)
    greys = rng.integers(10, 246, nshapes)
    kinds = rng.integers(0, 2, nshapes)
    for s, x, y, g, k in zip(sizes, xs, ys, greys, kinds):
        x1, y1 = min(size, x + s), min(size, y + s)
        if k == 0:
            img[y:y1, x:x1] = g
        else:
            yy, xx = np.ogrid[y:y1, x:x1]
            r = s / 2.0
            m = (yy - (y + r)) ** 2 + (xx - (x + r)) ** 2 <= r * r
            img[y:y1, x:x1][m] = g
    # separable 3-tap blur (~sigma 0.8) then noise
    k = np.array([0.25, 0.5, 0.25], np.float32)
    img = k[0] * np.roll(img, 1, 0) + k[1] * img + k[2] * np.roll(img, -1, 0)
    img = k[0] * np.roll(img, 1, 1) + k[1] * img + k[2] * np.roll(img, -1, 1)
    img += rng.normal(0, 2.0, img.shape).astype(np.float32)
    out = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    _WORLD_CACHE[key] = out
    return out


def frame(t=0, width=640, height=480, seed=0):
    """Frame t of a smooth synthetic trajectory over the world plane (bilinear-sampled similarity view)."""
    world = world_plane(seed)
    size = world.shape[0]
    ang = 0.15 * np.sin(0.05 * t + seed)
    zoom = 1.0 + 0.2 * np.sin(0.031 * t + 1.0)
    cx = size / 2 + 300 * np.sin(0.02 * t + 0.3 * seed)
    cy = size / 2 + 300 * np.cos(0.017 * t)
    ys, xs = np.mgrid[0:height, 0:width].astype(np.float32)
    xs -= width / 2
    ys -= height / 2
    c, s = np.cos(ang) * zoom, np.sin(ang) * zoom
    wx = (c * xs - s * ys + cx).astype(np.float32)
    wy = (s * xs + c * ys + cy).astype(np.float32)
    x0 = np.clip(np.floor(wx).astype(np.int32), 0, size - 2)
    y0 = np.clip(np.floor(wy).astype(np.int32), 0, size - 2)
    fx = np.clip(wx - x0, 0, 1)
    fy = np.clip(wy - y0, 0, 1)
    w = world.astype(np.float32)
    v = (w[y0, x0] * (1 - fx) * (1 - fy) + w[y0, x0 + 1] * fx * (1 - fy)
         + w[y0 + 1, x0] * (1 - fx) * fy + w[y0 + 1, x0 + 1] * fx * fy)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def frames(n, width=640, height=480, seed=0, t0=0):
    return np.stack([frame(t0 + i, width, height, seed) for i in range(n)])


# ---------------------------------------------------------------------------------------------
# Camera geometry of frame(): the similarity view of the plane z = 0 is exactly a pinhole camera
# (fx = fy = F, principal point at the image centre, zero distortion) looking down +z from
# C = (mu*cx, mu*cy, -Zc), rotated by `ang` about the optical axis, with Zc = mu * zoom * F.
# ---------------------------------------------------------------------------------------------
F_PIX = 458.0
MU = 3.0 / F_PIX   # metres per world-plane pixel: ~3 m depth at zoom 1


def _view(t, seed, size=2048):
    ang = 0.15 * np.sin(0.05 * t + seed)
    zoom = 1.0 + 0.2 * np.sin(0.031 * t + 1.0)
    cx = size / 2 + 300 * np.sin(0.02 * t + 0.3 * seed)
    cy = size / 2 + 300 * np.cos(0.017 * t)
    return ang, zoom, cx, cy


def camera(width=640, height=480):
    """(fx, fy, cx, cy) of the synthetic pinhole camera."""
    return np.array([F_PIX, F_PIX, width / 2.0, height / 2.0], np.float32)


def pose(t, seed=0):
    """Tcw of frame t as (qw, qx, qy, qz, tx, ty, tz), float64."""
    ang, zoom, cx, cy = _view(t, seed)
    zc = MU * zoom * F_PIX
    C = np.array([MU * cx, MU * cy, -zc])
    ca, sa = np.cos(-ang), np.sin(-ang)
    Rcw = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1.0]])
    tcw = -Rcw @ C
    return np.array([np.cos(-ang / 2), 0, 0, np.sin(-ang / 2), tcw[0], tcw[1], tcw[2]])


def backproject(xy, t, seed=0, width=640, height=480):
    """World coordinates (on the plane z = 0) of pixel positions xy [n, 2] seen in frame t."""
    ang, zoom, cx, cy = _view(t, seed)
    xs = xy[:, 0].astype(np.float64) - width / 2
    ys = xy[:, 1].astype(np.float64) - height / 2
    c, s = np.cos(ang) * zoom, np.sin(ang) * zoom
    wx = c * xs - s * ys + cx
    wy = s * xs + c * ys + cy
    return np.stack([MU * wx, MU * wy, np.zeros_like(wx)], 1)
