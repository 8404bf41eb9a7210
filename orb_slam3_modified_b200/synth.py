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


# ---------------------------------------------------------------------------------------------
# Local bundle adjustment problems (BASELINE config 4): flat arrays in g2o's Hessian order.
# ---------------------------------------------------------------------------------------------
def _quat_from_rotvec(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.array([1.0, 0, 0, 0])
    ax = rv / th
    return np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * ax])


def _qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def _qrot(q, v):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return v @ R.T


def lba_problem(n_kf=20, n_pts=5000, obs_per_pt=8, n_fixed=1, seed=0, width=640, height=480, outlier_frac=0.05,
                pose_noise=(0.01, 0.5), point_noise=0.02):
    """Returns a dict of flat arrays: poses [n_kf,7] (qw,qx,qy,qz,t), fixed [n_kf], cam [n_kf,4] f32, points [n_pts,3],
    edge_point / edge_pose [n_e], obs [n_e,2], inv_sigma2 [n_e] f32 (+ ground truth).  Edges are grouped by point."""
    rng = np.random.default_rng(seed)
    cam = np.tile(np.array([F_PIX, F_PIX, width / 2.0, height / 2.0], np.float32), (n_kf, 1))
    # ground-truth keyframes: a slow sideways arc looking down +z
    gt = np.zeros((n_kf, 7))
    for i in range(n_kf):
        q = _quat_from_rotvec(np.array([0.03 * np.sin(0.4 * i), 0.02 * i - 0.2, 0.02 * np.cos(0.3 * i)]))
        c = np.array([0.25 * i - 0.125 * n_kf, 0.05 * np.sin(0.5 * i), 0.1 * np.cos(0.2 * i)])   # camera centre in world
        qc = q * np.array([1, -1, -1, -1])    # Rcw = Rwc^T
        gt[i, :4] = qc
        gt[i, 4:] = -_qrot(qc, c)
    inv_sigma2_tab = (1.0 / (np.float32(1.2) ** np.arange(8, dtype=np.float32)) ** 2).astype(np.float32)
    level_p = np.array([217, 181, 151, 126, 105, 87, 73, 60], np.float64)
    level_p /= level_p.sum()
    pts = np.zeros((n_pts, 3))
    e_pt, e_kf, obs, isg = [], [], [], []
    span = max(1, n_kf - obs_per_pt + 1)
    for j in range(n_pts):
        s = j % span
        kfs = list(range(s, min(n_kf, s + obs_per_pt)))
        mid = kfs[len(kfs) // 2]
        # a point in front of the middle keyframe of its window
        z = rng.uniform(4.0, 10.0)
        u, v = rng.uniform(60, width - 60), rng.uniform(60, height - 60)
        pc = np.array([(u - width / 2) * z / F_PIX, (v - height / 2) * z / F_PIX, z])
        qc, t = gt[mid, :4], gt[mid, 4:]
        pw = _qrot(qc * np.array([1, -1, -1, -1]), pc - t)
        pts[j] = pw
        for k in kfs:
            xc = _qrot(gt[k, :4], pw) + gt[k, 4:]
            if xc[2] <= 0.1:
                continue
            oct_ = rng.choice(8, p=level_p)
            sig = 1.2 ** oct_
            uv = np.array([F_PIX * xc[0] / xc[2] + width / 2, F_PIX * xc[1] / xc[2] + height / 2]) + rng.normal(0, sig, 2)
            if rng.random() < outlier_frac:
                uv += rng.choice([-1, 1], 2) * rng.uniform(10, 20, 2)
            e_pt.append(j); e_kf.append(k); obs.append(uv.astype(np.float32).astype(np.float64)); isg.append(inv_sigma2_tab[oct_])
    poses = gt.copy()
    for i in range(n_fixed, n_kf):
        dq = _quat_from_rotvec(rng.normal(0, np.deg2rad(pose_noise[1]), 3))
        poses[i, :4] = _qmul(dq, gt[i, :4])
        poses[i, 4:] = gt[i, 4:] + rng.normal(0, pose_noise[0], 3)
    # float -> double casts as in Optimizer.cc:1217-1218,1286
    poses = poses.astype(np.float32).astype(np.float64)
    points = (pts + rng.normal(0, point_noise, pts.shape)).astype(np.float32).astype(np.float64)
    fixed = np.zeros(n_kf, np.uint8)
    fixed[:n_fixed] = 1
    return dict(poses=poses, fixed=fixed, cam=cam, points=points, edge_point=np.array(e_pt, np.int32), edge_pose=np.array(e_kf, np.int32),
                obs=np.array(obs, np.float64).reshape(-1, 2), inv_sigma2=np.array(isg, np.float32), gt_poses=gt, gt_points=pts,
                huber_delta=float(np.float32(np.sqrt(5.991))))


def pose_opt_problem(n=400, seed=0, width=640, height=480, outlier_frac=0.15, pose_noise=(0.03, 1.5)):
    """A tracked frame for Optimizer::PoseOptimization: map points, their (noisy) observations, a perturbed initial pose."""
    rng = np.random.default_rng(seed)
    gt = np.concatenate([_quat_from_rotvec(rng.normal(0, 0.1, 3)), rng.normal(0, 0.3, 3)])
    z = rng.uniform(3, 12, n)
    u, v = rng.uniform(20, width - 20, n), rng.uniform(20, height - 20, n)
    pc = np.stack([(u - width / 2) * z / F_PIX, (v - height / 2) * z / F_PIX, z], 1)
    Xw = _qrot(gt[:4] * np.array([1, -1, -1, -1]), pc - gt[4:])
    octv = rng.integers(0, 8, n)
    obs = np.stack([u, v], 1) + rng.normal(0, 1, (n, 2)) * (1.2 ** octv)[:, None]
    bad = rng.random(n) < outlier_frac
    obs[bad] += rng.uniform(8, 30, (int(bad.sum()), 2)) * rng.choice([-1, 1], (int(bad.sum()), 2))
    pose = gt.copy()
    pose[:4] = _qmul(_quat_from_rotvec(rng.normal(0, np.deg2rad(pose_noise[1]), 3)), gt[:4])
    pose[4:] += rng.normal(0, pose_noise[0], 3)
    inv_sigma2 = (1.0 / (np.float32(1.2) ** octv.astype(np.float32)) ** 2).astype(np.float32)
    return dict(pose=pose.astype(np.float32).astype(np.float64), cam=np.array([F_PIX, F_PIX, width / 2, height / 2], np.float32),
                Xw=Xw.astype(np.float32).astype(np.float64), obs=obs.astype(np.float32).astype(np.float64), inv_sigma2=inv_sigma2, gt_pose=gt)


# ---------------------------------------------------------------------------------------------
# Inertial data (300 Hz accel + gyro from an analytic trajectory + gravity, EuRoC-like noise)
# ---------------------------------------------------------------------------------------------
IMU_NOISE = (1.7e-4 * np.sqrt(200.0), 2.0e-3 * np.sqrt(200.0), 1.9393e-5 / np.sqrt(200.0), 3.0e-3 / np.sqrt(200.0))   # ng, na, ngw, naw as Tracking builds them from EuRoC.yaml (NoiseGyro * sqrt(freq), ...)


def _rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def imu_trajectory(t):
    """Body pose in the world (Rwb, twb), velocity, body angular velocity and world acceleration of a smooth trajectory at time t [s]."""
    w = np.array([0.2 * np.sin(0.7 * t), 0.15 * np.cos(0.5 * t), 0.1 * np.sin(0.3 * t + 1.0)])          # rotation vector of Rwb(t)
    eps = 1e-5
    R = _rodrigues(w)
    w2 = np.array([0.2 * np.sin(0.7 * (t + eps)), 0.15 * np.cos(0.5 * (t + eps)), 0.1 * np.sin(0.3 * (t + eps) + 1.0)])
    dR = R.T @ _rodrigues(w2)
    omega = np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]) / (2 * eps)       # body rate
    p = np.array([0.5 * np.sin(0.4 * t), 0.3 * np.cos(0.3 * t), 0.2 * np.sin(0.2 * t)])
    v = np.array([0.2 * np.cos(0.4 * t), -0.09 * np.sin(0.3 * t), 0.04 * np.cos(0.2 * t)])
    a = np.array([-0.08 * np.sin(0.4 * t), -0.027 * np.cos(0.3 * t), -0.008 * np.sin(0.2 * t)])
    return R, p, v, omega, a


def imu_interval(t0, t1, rate=300.0, seed=0, bias=(0.02, -0.01, 0.03, 0.002, -0.001, 0.0015), noise=True):
    """IMU samples (acc [n,3], gyro [n,3], dt [n]) between two frame times: specific force in the body frame = Rbw (a - g_w), g_w = (0,0,-9.81),
    + bias (bax..baz, bwx..bwz) + white noise."""
    rng = np.random.default_rng(seed)
    n = max(1, int(round((t1 - t0) * rate)))
    dt = (t1 - t0) / n
    acc, gyr = [], []
    for i in range(n):
        R, p, v, om, a = imu_trajectory(t0 + (i + 0.5) * dt)
        f = R.T @ (a - np.array([0, 0, -9.81]))
        acc.append(f + np.array(bias[:3]) + (rng.normal(0, IMU_NOISE[1], 3) if noise else 0))
        gyr.append(om + np.array(bias[3:]) + (rng.normal(0, IMU_NOISE[0], 3) if noise else 0))
    return np.array(acc, np.float32), np.array(gyr, np.float32), np.full(n, dt, np.float32)


def inertial_edge_state(t0, t1, seed=0, perturb=1.0):
    """Vertex estimates of one EdgeInertial (keyframes at t0, t1): ground truth + a seeded perturbation."""
    rng = np.random.default_rng(seed + 77)
    R1, p1, v1, _, _ = imu_trajectory(t0)
    R2, p2, v2, _, _ = imu_trajectory(t1)
    s = dict(Rwb1=R1 @ _rodrigues(perturb * rng.normal(0, 0.01, 3)), twb1=p1 + perturb * rng.normal(0, 0.01, 3), v1=v1 + perturb * rng.normal(0, 0.02, 3),
             bg=np.array([0.002, -0.001, 0.0015]) + perturb * rng.normal(0, 1e-3, 3), ba=np.array([0.02, -0.01, 0.03]) + perturb * rng.normal(0, 1e-2, 3),
             Rwb2=R2 @ _rodrigues(perturb * rng.normal(0, 0.01, 3)), twb2=p2 + perturb * rng.normal(0, 0.01, 3), v2=v2 + perturb * rng.normal(0, 0.02, 3))
    return {k: np.ascontiguousarray(v, np.float64) for k, v in s.items()}


def stereo_pair(t=0, width=640, height=480, seed=0, baseline_px=22.0):
    """Left / right frames of a rectified stereo rig looking at the scene plane: the right view is the left one shifted by the disparity of
    the plane (all points share one depth in frame(), so the disparity is constant: mbf / Z = baseline_px) + independent sensor noise."""
    left = frame(t, width + 64, height, seed)
    rng = np.random.default_rng(1000 + seed * 31 + t)
    d = int(round(baseline_px))
    right = left[:, 32 + d:32 + d + width].astype(np.int16) + rng.integers(-2, 3, (height, width))
    return np.ascontiguousarray(left[:, 32:32 + width]), np.clip(right, 0, 255).astype(np.uint8)


IMU_EXTRINSICS = None


def imu_extrinsics():
    """Tcb / Tbc of the synthetic rig (camera a few centimetres off the body, slightly rotated): Rcb 9 | tcb 3 | Rbc 9 | tbc 3."""
    Rcb = _rodrigues(np.array([0.02, -0.015, 0.01])) @ np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    tcb = np.array([0.03, -0.02, 0.01])
    Rbc = Rcb.T
    tbc = -Rbc @ tcb
    return np.concatenate([Rcb.reshape(9), tcb, Rbc.reshape(9), tbc])


def pose_inertial_problem(t0=1.0, t1=1.2, n=300, seed=0, outlier_frac=0.1, width=640, height=480, perturb=1.0, noise_px=0.7):
    """One frame for Optimizer::PoseInertialOptimizationLastKeyFrame: the last keyframe at t0 (state known), the frame at t1 (state = ground truth + a
    perturbation), the IMU samples between them, n map points seen by the frame's camera with noisy observations and a share of gross outliers."""
    rng = np.random.default_rng(seed)
    ex = imu_extrinsics()
    Rcb, tcb = ex[:9].reshape(3, 3), ex[9:12]
    R1, p1, v1, _, _ = imu_trajectory(t0)
    R2, p2, v2, _, _ = imu_trajectory(t1)
    bias = np.array([0.02, -0.01, 0.03, 0.002, -0.001, 0.0015])                     # bax.. bwx..
    acc, gyr, dts = imu_interval(t0, t1, seed=seed, bias=tuple(bias))
    cam = camera(width, height)
    # points in front of the true camera
    Rcw = Rcb @ R2.T
    tcw = Rcb @ (-R2.T @ p2) + tcb
    z = rng.uniform(2.0, 25.0, n)
    u, v = rng.uniform(15, width - 15, n), rng.uniform(15, height - 15, n)
    Xc = np.stack([(u - cam[2]) * z / cam[0], (v - cam[3]) * z / cam[1], z], 1)
    Xw = (Xc - tcw) @ Rcw                                                            # Rcw^T (Xc - tcw)
    octave = rng.integers(0, 8, n)
    sig = 1.2 ** octave
    obs = np.stack([u, v], 1) + rng.normal(0, noise_px, (n, 2)) * sig[:, None]
    bad = rng.random(n) < outlier_frac
    obs[bad] += rng.uniform(8, 40, (int(bad.sum()), 2)) * rng.choice([-1, 1], (int(bad.sum()), 2))
    st = lambda R, p, v_, bg, ba: np.concatenate([R.reshape(9), p, v_, bg, ba]).astype(np.float64)
    kf = st(R1, p1, v1, bias[3:], bias[:3])
    fr = st(R2 @ _rodrigues(perturb * rng.normal(0, 0.01, 3)), p2 + perturb * rng.normal(0, 0.02, 3), v2 + perturb * rng.normal(0, 0.05, 3),
            bias[3:] + perturb * rng.normal(0, 2e-4, 3), bias[:3] + perturb * rng.normal(0, 2e-3, 3))
    return dict(Xw=Xw.astype(np.float32), obs=obs.astype(np.float32), inv_sigma2=(1.0 / 1.44 ** octave).astype(np.float32), track_depth=z.astype(np.float32), cam=cam,
                extr=ex, acc=acc, gyr=gyr, dt=dts, bias6=bias.astype(np.float32), kf_state=kf, state=fr, truth=st(R2, p2, v2, bias[3:], bias[:3]), gross=bad)


def pose_inertial_problem_last_frame(t0=1.0, t1=1.15, t2=1.2, n=300, seed=0, outlier_frac=0.1, perturb=1.0, noise_px=0.7, prior_sigma=(2e-3, 5e-3, 2e-2, 1e-4, 1e-3)):
    """One frame for Optimizer::PoseInertialOptimizationLastFrame: last keyframe at t0 (only its preintegration covariance is used), previous frame at
    t1 (free, held by a prior = its true state + noise with a diagonal information), the frame at t2."""
    pr = pose_inertial_problem(t1, t2, n=n, seed=seed, outlier_frac=outlier_frac, perturb=perturb, noise_px=noise_px)
    rng = np.random.default_rng(seed + 991)
    acc, gyr, dts = imu_interval(t0, t2, seed=seed + 1, bias=tuple(pr['bias6']))
    pr['acc_kf'], pr['gyr_kf'], pr['dt_kf'] = acc, gyr, dts
    truth_prev = pr['kf_state'].copy()                      # pose_inertial_problem's "keyframe" is the previous frame here
    sr, st, sv, sg, sa = prior_sigma
    prior = truth_prev.copy()
    prior[:9] = (prior[:9].reshape(3, 3) @ _rodrigues(rng.normal(0, sr, 3))).reshape(9)
    prior[9:12] += rng.normal(0, st, 3); prior[12:15] += rng.normal(0, sv, 3); prior[15:18] += rng.normal(0, sg, 3); prior[18:21] += rng.normal(0, sa, 3)
    H = np.diag(np.concatenate([np.full(3, 1 / sr ** 2), np.full(3, 1 / st ** 2), np.full(3, 1 / sv ** 2), np.full(3, 1 / sg ** 2), np.full(3, 1 / sa ** 2)]))
    A = rng.normal(0, 1, (15, 15)); Q, _ = np.linalg.qr(A)
    M = np.eye(15) + 0.05 * (Q - np.eye(15))               # a mild mixing so that the information is not diagonal
    pr['prior_state'] = prior; pr['prior_H'] = M.T @ H @ M
    pr['prev_state'] = prior.copy()                         # VertexPose(pFp) etc.: the previous frame's current estimate
    pr['truth_prev'] = truth_prev
    return pr


def local_inertial_ba_problem(n_opt=10, n_cov_fixed=3, n_pts=800, seed=0, kf_dt=0.25, t_end=6.0, width=640, height=480, outlier_frac=0.03, perturb=1.0,
                              noise_px=0.7, float_inputs=True, rec_init=False, large=False):
    """One Optimizer::LocalInertialBA graph (reference src/Optimizer.cc:2383-2958), flattened as the caller's graph walk leaves it:
    keyframes [0, n_opt) = the temporal window, NEWEST FIRST (vpOptimizableKFs order); keyframe n_opt = the window's predecessor (fixed, linked by
    the last EdgeInertial); n_cov_fixed older keyframes that only see points (fixed).  Inertial edge i links keyframe i+1 -> i; the one to the
    fixed keyframe carries the Huber kernel and information * 1e-2 (:2633-2643), all carry the kernel with rec_init.  Points are seen by every keyframe in
    whose image they fall.  States = ground truth + a seeded perturbation; with float_inputs they are rounded to float like KeyFrame storage and
    the camera pose Tcw is rounded separately (ImuCamPose(KeyFrame*) loads both, src/G2oTypes.cc:25-47)."""
    rng = np.random.default_rng(seed)
    ex = imu_extrinsics()
    Rcb, tcb = ex[:9].reshape(3, 3), ex[9:12]
    nkf = n_opt + 1 + n_cov_fixed
    times = [t_end - kf_dt * k for k in range(n_opt + 1)] + [t_end - kf_dt * (n_opt + 1) - 0.4 * (j + 1) for j in range(n_cov_fixed)]
    bias = np.array([0.02, -0.01, 0.03, 0.002, -0.001, 0.0015])                      # bax.. bwx..
    cam = camera(width, height)
    truth = np.zeros((nkf, 21)); state = np.zeros((nkf, 21)); tcw12 = np.zeros((nkf, 12)); tcw_true = np.zeros((nkf, 12))

    def cam_pose(R, p):
        Rcw = Rcb @ R.T
        return Rcw, Rcb @ (-R.T @ p) + tcb
    for k, t in enumerate(times):
        R, p, v, _, _ = imu_trajectory(t)
        truth[k] = np.concatenate([R.reshape(9), p, v, bias[3:], bias[:3]])
        Rc, tc = cam_pose(R, p)
        tcw_true[k] = np.concatenate([Rc.reshape(9), tc])
        if k < n_opt:
            Rn = R @ _rodrigues(perturb * rng.normal(0, 0.004, 3)); pn = p + perturb * rng.normal(0, 0.01, 3)
            state[k] = np.concatenate([Rn.reshape(9), pn, v + perturb * rng.normal(0, 0.03, 3), bias[3:] + perturb * rng.normal(0, 1e-4, 3),
                                       bias[:3] + perturb * rng.normal(0, 2e-3, 3)])
        else:
            Rn, pn = R, p
            state[k] = truth[k]
        Rc, tc = cam_pose(Rn, pn)
        tcw12[k] = np.concatenate([Rc.reshape(9), tc])
    if float_inputs:
        state = state.astype(np.float32).astype(np.float64)
        tcw12 = tcw12.astype(np.float32).astype(np.float64)
    # points in front of the newest camera, seen by every keyframe whose image contains them
    Rc0, tc0 = tcw_true[0, :9].reshape(3, 3), tcw_true[0, 9:]
    z = rng.uniform(2.0, 25.0, n_pts)
    u, v = rng.uniform(-40, width + 40, n_pts), rng.uniform(-30, height + 30, n_pts)
    Xc = np.stack([(u - cam[2]) * z / cam[0], (v - cam[3]) * z / cam[1], z], 1)
    Xw = (Xc - tc0) @ Rc0
    e_pt, e_kf, obs, isig = [], [], [], []
    for j in range(n_pts):
        for k in range(nkf):
            Xk = tcw_true[k, :9].reshape(3, 3) @ Xw[j] + tcw_true[k, 9:]
            if Xk[2] < 0.5:
                continue
            uu, vv = cam[0] * Xk[0] / Xk[2] + cam[2], cam[1] * Xk[1] / Xk[2] + cam[3]
            if not (5 < uu < width - 5 and 5 < vv < height - 5) or rng.random() < 0.15:
                continue
            octv = int(rng.integers(0, 8))
            o = np.array([uu, vv]) + rng.normal(0, noise_px, 2) * 1.2 ** octv
            if rng.random() < outlier_frac:
                o += rng.uniform(8, 40, 2) * rng.choice([-1, 1], 2)
            e_pt.append(j); e_kf.append(k); obs.append(o); isig.append(1.0 / 1.44 ** octv)
    e_pt = np.array(e_pt, np.int32); e_kf = np.array(e_kf, np.int32)
    # drop points with fewer than two observations, re-index
    cnt = np.bincount(e_pt, minlength=n_pts)
    keep = cnt >= 2
    remap = -np.ones(n_pts, np.int64); remap[keep] = np.arange(int(keep.sum()))
    sel = keep[e_pt]
    e_pt = remap[e_pt[sel]].astype(np.int32); e_kf = e_kf[sel]
    obs = np.array(obs)[sel].astype(np.float32).astype(np.float64); isig = np.array(isig, np.float32)[sel]
    pts_true = Xw[keep]
    pts = pts_true + perturb * rng.normal(0, 0.02, pts_true.shape)
    if float_inputs:
        pts = pts.astype(np.float32).astype(np.float64)
    imu = [imu_interval(times[i + 1], times[i], seed=seed * 100 + i, bias=tuple(bias)) for i in range(n_opt)]
    return dict(n_kf=nkf, n_opt=n_opt, state=np.ascontiguousarray(state), tcw=np.ascontiguousarray(tcw12), cam=np.tile(cam, (nkf, 1)).astype(np.float32), extr=ex,
                ie_kf1=np.arange(1, n_opt + 1, dtype=np.int32), ie_kf2=np.arange(0, n_opt, dtype=np.int32), imu=imu, bias6=bias.astype(np.float32),
                ie_robust=np.array([1 if (i == n_opt - 1 or rec_init) else 0 for i in range(n_opt)], np.uint8),
                ie_info_scale=np.array([1e-2 if i == n_opt - 1 else 1.0 for i in range(n_opt)]),
                points=np.ascontiguousarray(pts), track_depth=z[keep].astype(np.float32), e_pt=e_pt, e_kf=e_kf, obs=np.ascontiguousarray(obs), inv_sigma2=isig,
                iterations=4 if large else 10, lambda_init=1e-2 if large else 1.0, large=bool(large), truth=truth, points_true=pts_true, times=times)
