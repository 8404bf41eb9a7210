"""Latency lines for the BASELINE configurations bench.py's throughput metric does not cover (called by bench.py, rank 0, N=1):
  configs[0]  one 640x480 frame at a time, ORBextractor(1000, 1.2, 8) -- the way the reference is actually driven (30 fps, one frame
              per call, src/Tracking.cc:1572-1600 -> Frame::ExtractORB): operator() latency through the host C-ABI, the same on the
              device only, and operator() + SearchByProjection(last frame);
  configs[2]  1280x720 tracking of one frame against the local map (src/Tracking.cc:2859-2974, 3346-3416): operator(),
              Frame::isInFrustum over the local map points, SearchByProjection(local map), PoseOptimization.
Every number is a median over repetitions of host-API calls (inputs and outputs in host memory, synchronous) unless it says device."""
import time

import numpy as np


def _median_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def measure(orb, synth, device=0, reps=40):
    import torch
    out = {}
    # ---------------- configs[0]: single 640x480 frame ----------------
    W, H = 640, 480
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, W, H, 1, device)
    cap = ex.max_keypoints
    mt = orb.ORBmatcher(0.9, True, max_batch=1, max_keypoints=cap, max_mappoints=cap, device=device)
    img0, img1 = synth.frame(10, W, H, 1), synth.frame(11, W, H, 1)
    sf = ex.GetScaleFactors()
    cam = synth.camera(W, H)
    _, k0, d0 = ex(img0, (0, 1000))
    last = dict(valid=np.ones(len(k0), np.uint8), xyz=synth.backproject(np.stack([k0['x'], k0['y']], 1), 10, 1, W, H).astype(np.float32),
                octave=k0['octave'].astype(np.int32), angle=k0['angle'].astype(np.float32), hasObs=np.ones(len(k0), np.uint8), descriptors=d0)
    Tcw = synth.pose(11, 1).astype(np.float32)
    bounds = (0.0, 0.0, float(W), float(H))
    c1 = {'extract_host_ms': _median_ms(lambda: ex(img1, (0, 1000)), reps)}

    def ext_match():
        _, k, d = ex(img1, (0, 1000))
        F = orb.Frame(k, d, bounds, sf)
        return mt.SearchByProjection(F, last, 15.0, True, Tcw=Tcw, cam=cam)
    c1['extract_plus_match_host_ms'] = _median_ms(ext_match, reps)
    c1['matches'] = int(ext_match())
    dev = torch.device('cuda', device)
    d_img = torch.from_numpy(img1).to(dev).unsqueeze(0).contiguous()
    d_kps = torch.zeros((1, cap, 7), dtype=torch.float32, device=dev); d_desc = torch.zeros((1, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(1, dtype=torch.int32, device=dev); d_mono = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    ts = []
    for i in range(reps + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ex.extract_batch_device(d_img, d_kps, d_desc, d_n, d_mono, (0, 1000), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    c1['extract_device_ms'] = float(np.median(ts))
    c1['keypoints'] = int(d_n.item())
    c1['fps_single_stream_host'] = 1e3 / c1['extract_plus_match_host_ms']
    c1['note'] = 'batch of ONE frame: latency-bound (10 dependent kernel launches, 7 of them the pyramid chain); 30 fps needs 33 ms'
    out['configs[0] single 640x480 frame'] = c1
    ex.close(); mt.close()

    # ---------------- configs[2]: 1280x720 tracking against the local map ----------------
    W, H = 1280, 720
    t, seed = 20, 2
    ex = orb.ORBextractor(1000, 1.2, 8, 20, 7, W, H, 1, device)
    cap = ex.max_keypoints
    img = synth.frame(t, W, H, seed)
    cam = synth.camera(W, H)
    bounds = (0.0, 0.0, float(W), float(H))
    T = synth.pose(t, seed)
    Rcw = _rot(T[:4]); tcw = T[4:]; Ow = -Rcw.T @ tcw
    # local map: the keypoints of the neighbouring frames, back-projected onto the scene plane
    P, D, O = [], [], []
    for dt in (-2, -1, 1, 2):
        _, k, d = ex(synth.frame(t + dt, W, H, seed), (0, 1000))
        P.append(synth.backproject(np.stack([k['x'], k['y']], 1), t + dt, seed, W, H)); D.append(d); O.append(k['octave'])
    P, D, O = np.concatenate(P), np.concatenate(D), np.concatenate(O)
    M = len(P)
    PO = P - Ow
    dist = np.linalg.norm(PO, axis=1)
    dmax = (dist * 1.2 ** O).astype(np.float32)                  # mfMaxDistance = dist * scale^level of the reference observation
    dmin = (dmax / np.float32(1.2 ** 7)).astype(np.float32)
    pts = dict(worldPos=P.astype(np.float32), normal=(PO / dist[:, None]).astype(np.float32), minDistInv=np.float32(0.8) * dmin,
               maxDistInv=np.float32(1.2) * dmax, maxDistance=dmax)
    mt = orb.ORBmatcher(0.8, True, max_batch=1, max_keypoints=cap, max_mappoints=max(M, 1), device=device)
    sf = ex.GetScaleFactors()
    logsf = np.float32(np.log(np.float32(1.2)))
    c3 = {'local_map_points': int(M)}
    state = {}

    def s_extract():
        state['kd'] = ex(img, (0, 1000))[1:]

    def s_frustum():
        state['fr'] = mt.isInFrustum(pts, Rcw.astype(np.float32), tcw.astype(np.float32), Ow.astype(np.float32), cam, bounds, logsf, 8, 0.5, 40.0)

    def s_search():
        k, d = state['kd']
        fr = state['fr']
        F = orb.Frame(k, d, bounds, sf)
        mp = dict(inView=fr['inView'], bad=np.zeros(M, np.uint8), depth=fr['depth'], projX=fr['projX'], projY=fr['projY'], level=np.maximum(fr['level'], 0),
                  viewCos=fr['viewCos'], hasObs=np.ones(M, np.uint8), descriptors=D)
        state['n'] = mt.SearchByProjection(F, mp, 1.0)
        state['F'] = F

    def s_pose():
        F = state['F']
        k = state['kd'][0]
        sel = np.flatnonzero(F.match >= 0)
        isg = (1.0 / (np.float32(1.2) ** k['octave'][sel].astype(np.float32)) ** 2).astype(np.float32)
        fr = dict(pose=T, cam=cam, Xw=P[F.match[sel]], obs=np.stack([k['x'][sel], k['y'][sel]], 1).astype(np.float64), inv_sigma2=isg)
        state['po'] = orb.PoseOptimization([fr], device)[0]

    for name, fn in (('extract_host_ms', s_extract), ('is_in_frustum_host_ms', s_frustum), ('search_local_map_host_ms', s_search), ('pose_optimization_host_ms', s_pose)):
        c3[name] = _median_ms(fn, reps // 2)
    c3['frame_ms'] = sum(c3[k] for k in ('extract_host_ms', 'is_in_frustum_host_ms', 'search_local_map_host_ms', 'pose_optimization_host_ms'))
    c3['fps_single_stream_host'] = 1e3 / c3['frame_ms']
    c3['keypoints'], c3['in_view'], c3['matches'], c3['pose_inliers'] = int(len(state['kd'][0])), int(state['fr']['inView'].sum()), int(state['n']), int(state['po']['inliers'])
    c3['reference'] = 'Tracking::TrackLocalMap, src/Tracking.cc:2859-2974 + SearchLocalPoints :3346-3416'
    out['configs[2] 1280x720 tracking vs local map'] = c3
    # ---------------- mono-inertial tracking: the per-frame inertial pose optimisers ----------------
    # Tracking::TrackLocalMap calls PoseInertialOptimizationLastFrame / LastKeyFrame instead of PoseOptimization once the IMU is initialised
    # (src/Tracking.cc:2985-2994): one frame alone (latency) and the frames of 240 streams in one call, 700 matched map points each
    try:
        def problems(n, npts):
            prs = []
            for k in range(n):
                pr = synth.pose_inertial_problem_last_frame(seed=k % 8, n=npts, outlier_frac=0.1)
                pre = lambda a, g, d: orb.imu_preintegrate(a[None], g[None], d[None], [len(d)], pr['bias6'][None], synth.IMU_NOISE, device)[0]
                pr['preint_frame'] = pre(pr['acc'], pr['gyr'], pr['dt'])
                pr['preint_kf'] = pre(pr['acc_kf'], pr['gyr_kf'], pr['dt_kf'])
                pr['preint'] = pr['preint_frame']
                prs.append(pr)
            return prs
        one, many = problems(1, 700), problems(8, 700) * 30
        ex24 = one[0]['extr']
        r1 = orb.PoseInertialOptimizationLastFrame(one, ex24, device=device)[0]
        out['mono-inertial tracking: inertial pose optimisers (700 map points per frame)'] = {
            'last_frame_1_frame_host_ms': _median_ms(lambda: orb.PoseInertialOptimizationLastFrame(one, ex24, device=device), 20),
            'last_keyframe_1_frame_host_ms': _median_ms(lambda: orb.PoseInertialOptimizationLastKeyFrame(one, ex24, device=device), 20),
            'last_frame_240_frames_host_ms': _median_ms(lambda: orb.PoseInertialOptimizationLastFrame(many, ex24, device=device), 5, warm=1),
            'last_keyframe_240_frames_host_ms': _median_ms(lambda: orb.PoseInertialOptimizationLastKeyFrame(many, ex24, device=device), 5, warm=1),
            'inliers_of_700': int(r1['ret']),
            'reference': 'Optimizer::PoseInertialOptimizationLastFrame / LastKeyFrame, src/Optimizer.cc:4491-5289; includes the Python marshalling of the batch'}
    except Exception as exc:          # the latency lines are informative; never fail the bench line over them
        out['mono-inertial tracking: inertial pose optimisers (700 map points per frame)'] = {'error': repr(exc)[:200]}
    # ---------------- mono-inertial mapping: Optimizer::LocalInertialBA ----------------
    # LocalMapping::Run calls it instead of LocalBundleAdjustment once the IMU is initialised (src/LocalMapping.cc:129-151): 10 keyframes in the
    # temporal window + 7 fixed, ~2100 points, ~24.5k EdgeMono; one map alone (latency) and one map per SM in a single launch
    key = 'mono-inertial mapping: LocalInertialBA (10 + 7 keyframes, ~2100 points, ~24.5k edges)'
    try:
        def maps(n):
            prs = []
            for k in range(n):
                pr = synth.local_inertial_ba_problem(n_opt=10, n_cov_fixed=6, n_pts=2500, seed=12 + k)
                pr['preint'] = np.stack([orb.imu_preintegrate(a[None], g[None], d[None], [len(d)], pr['bias6'][None], synth.IMU_NOISE, device)[0] for a, g, d in pr['imu']])
                prs.append(pr)
            return prs
        one = maps(1)
        four = one + maps(4)[1:]
        r1 = orb.LocalInertialBA(one, device=device)[0]
        out[key] = {
            '1_map_host_ms': _median_ms(lambda: orb.LocalInertialBA(one, device=device), 10),
            '148_maps_host_ms': _median_ms(lambda: orb.LocalInertialBA(four * 37, device=device), 3, warm=1),
            '1_map_solver_ms': float(np.median([orb.LocalInertialBA(one, device=device)[0]['kernel_ms'] for _ in range(5)])),
            '148_maps_solver_ms_per_cta': float(np.median([r['kernel_ms'] for r in orb.LocalInertialBA(four * 37, device=device)])),
            'edges': int(len(one[0]['e_pt'])), 'points': int(len(one[0]['points'])), 'iterations': int(r1['iters']), 'lm_trials': int(r1['trials']),
            'erased_observations': int(r1['erase'].sum()),
            'reference': 'Optimizer::LocalInertialBA, src/Optimizer.cc:2383-2958; host = the API call incl. the Python marshalling and the packing of the graph, solver = the CTA\'s own globaltimer span'}
    except Exception as exc:
        out[key] = {'error': repr(exc)[:200]}
    return out
