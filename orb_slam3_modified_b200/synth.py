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
