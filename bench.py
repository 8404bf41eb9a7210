#!/usr/bin/env python3
"""bench.py -- frames/sec of the ORB-SLAM3 hot path (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic 640x480 mono frames per GPU
(BASELINE configs[1]: 640x480 mono stream, 1000 feats/frame, extended by the LBA share of configs[3]):
  * every stream contributes one frame: ORBextractor::operator() + SearchByProjection(current, last frame);
  * every KF_INTERVAL-th frame of a stream is a keyframe and triggers one LocalBundleAdjustment of the
    configs[3] size (20 keyframes x 5000 points x 40000 edges) on the mapping side -> B / KF_INTERVAL LBAs per step.
`value` is whole-job frames/s with all inputs already resident in HBM; `e2e` is the same metric through the
reference-facing C-ABI with HOST buffers (pinned), host<->device copies inside the timed region.
--impl reference times the CPU oracle (the reference itself cannot be built here: needs OpenCV C++/Eigen)
on all host cores for the same work mix.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'frames/sec (extract+match+LBA) 640x480 mono-inertial'
UNIT = 'frames/s'
W, H, NFEAT = 640, 480, 1000
KF_INTERVAL = 10     # one keyframe (-> one LBA) per 10 frames of a stream
LBA_CFG = dict(n_kf=20, n_pts=5000, obs_per_pt=8)   # BASELINE configs[3]
STAGES = ['extract', 'match(SearchByProjection last frame)', 'LBA(1 per %d frames)' % KF_INTERVAL]
WORKLOAD = 'configs[1]: 640x480 mono stream, 1000 feats/frame, extract+SearchByProjection, + configs[3]-sized LBA every %d frames' % KF_INTERVAL
# SURVEY.md 8d algorithmic bytes, 640x480, K = 1000 keypoints (S = 950,532 px over the 8 levels, A0 = 307,200, A7 = 23,986)
ALG = {'pyramid': 926546 + 643332, 'blur': 2 * 950532, 'fast_cells': 950532, 'quadtree_orient': 749 * 1000, 'assemble': 60 * 1000,
       'brief': 512 * 1000 + 32 * 1000}
ALG_BYTES_EXTRACT = 5742474
ALG_BYTES_MATCH = 552000
ALG_BYTES_LBA_PER_TRIAL = 13.6e6
DISTINCT = 32        # distinct synthetic streams; larger batches replicate them (separate buffers, same content)
TH_PROJ = 15.0       # SearchByProjection window for mono tracking (reference src/Tracking.cc:2884-2889)


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.m0 = [], None, 0
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap',
                 '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def mark(self):
        """Start of the region of interest: only samples taken from here on are reported."""
        self.m0 = len(self.rows)

    def samples(self):
        return len(self.rows) - self.m0

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = self.rows[self.m0:]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def stream_time(s, k):
    """Frame index of distinct stream s at parity k (consecutive frames t, t+1)."""
    return 5 * s + k


def make_frames(n, k, rank=0):
    from orb_slam3_modified_b200 import synth
    import numpy as np
    d = min(n, DISTINCT)
    base = np.stack([synth.frame(stream_time(s, k), W, H, seed=s % 4 + 4 * rank) for s in range(d)])
    return np.ascontiguousarray(np.concatenate([base] * ((n + d - 1) // d))[:n])


def stream_pose(s, k, rank=0, noise=0.003):
    """Motion-model prior of Tcw for stream s at parity k: the exact synthetic pose plus a seeded perturbation."""
    from orb_slam3_modified_b200 import synth
    import numpy as np
    d = s % DISTINCT
    rng = np.random.default_rng(1000 * rank + 2 * d + k)
    T = synth.pose(stream_time(d, k), seed=d % 4 + 4 * rank)
    T[4:] += rng.normal(0, noise, 3)
    return T.astype(np.float32)


def last_frame_slabs(kps_list, desc_list, k_last, cap, rank=0):
    """Map points of the 'last frame' (parity k_last) of every stream as fixed-capacity host slabs."""
    from orb_slam3_modified_b200 import synth
    import numpy as np
    B = len(kps_list)
    out = dict(nM=np.zeros(B, np.int32), valid=np.zeros((B, cap), np.uint8), xyz=np.zeros((B, cap, 3), np.float32),
               octave=np.zeros((B, cap), np.int32), angle=np.zeros((B, cap), np.float32), hasObs=np.zeros((B, cap), np.uint8),
               mpDesc=np.zeros((B, cap, 32), np.uint8))
    for b in range(B):
        k, d = kps_list[b], desc_list[b]
        m = len(k)
        sd = b % DISTINCT
        out['nM'][b] = m
        out['valid'][b, :m] = 1
        out['xyz'][b, :m] = synth.backproject(np.stack([k['x'], k['y']], 1), stream_time(sd, k_last), sd % 4 + 4 * rank, W, H)
        out['octave'][b, :m] = k['octave']
        out['angle'][b, :m] = k['angle']
        out['hasObs'][b, :m] = 1
        out['mpDesc'][b, :m] = d
    return out


def lba_problems(n, rank=0):
    from orb_slam3_modified_b200 import synth
    base = [synth.lba_problem(seed=100 * rank + i, **LBA_CFG) for i in range(min(n, 4))]
    return [base[i % len(base)] for i in range(n)]


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference path)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_mix(frames0, frames1, poses1, lba_prob, n_frames, threads):
    """Runs extract + SearchByProjection on n_frames frames and n_frames/KF_INTERVAL LBAs with the CPU oracle on `threads` threads.
    Returns (frames/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import numpy as np
    import oracle_lib as O
    from orb_slam3_modified_b200 import synth
    from concurrent.futures import ThreadPoolExecutor
    exs = [O.OracleExtractor(NFEAT, 1.2, 8, 20, 7) for _ in range(threads)]
    sf = exs[0].tables()['scale']
    cam = synth.camera(W, H)
    nsrc = len(frames0)
    last = [None] * nsrc
    for s in range(min(nsrc, n_frames)):     # untimed: the last frame's features / map points
        _, k, d = exs[0](frames0[s], (0, 1000))
        last[s] = dict(valid=np.ones(len(k), np.uint8), xyz=synth.backproject(np.stack([k['x'], k['y']], 1), stream_time(s, 0), s % 4, W, H).astype(np.float32),
                       octave=k['octave'].astype(np.int32), angle=k['angle'].astype(np.float32), hasObs=np.ones(len(k), np.uint8), descriptors=d)
    n_lba = n_frames // KF_INTERVAL

    def work(tid):
        e = exs[tid]
        for i in range(tid, n_frames, threads):
            s = i % nsrc
            _, k, d = e(frames1[s], (0, 1000))
            match = np.full(len(k), -1, np.int32)
            claimed = np.zeros(len(k), np.uint8)
            O.search_last_frame(k, d, (0.0, 0.0, float(W), float(H)), sf, poses1[s], cam, last[s], TH_PROJ, True, match, claimed)
        for i in range(tid, n_lba, threads):
            O.lba_solve(lba_prob)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def cpu_info():
    model = None
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    return {'model': model, 'nproc': os.cpu_count()}


_REF = {}


def _ref_init(nsrc):
    """Worker-process initialiser of the reference arm: oracle handles + the untimed 'last frame' of every source stream."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import numpy as np
    import oracle_lib as O
    from orb_slam3_modified_b200 import synth
    ex = O.OracleExtractor(NFEAT, 1.2, 8, 20, 7)
    f0, f1 = make_frames(nsrc, 0), make_frames(nsrc, 1)
    last = []
    for s in range(nsrc):
        _, k, d = ex(f0[s], (0, 1000))
        last.append(dict(valid=np.ones(len(k), np.uint8), xyz=synth.backproject(np.stack([k['x'], k['y']], 1), stream_time(s, 0), s % 4, W, H).astype(np.float32),
                         octave=k['octave'].astype(np.int32), angle=k['angle'].astype(np.float32), hasObs=np.ones(len(k), np.uint8), descriptors=d))
    _REF.update(O=O, ex=ex, f1=f1, last=last, sf=ex.tables()['scale'], cam=synth.camera(W, H), poses=[stream_pose(s, 1) for s in range(nsrc)],
                prob=lba_problems(1)[0], nsrc=nsrc)
    return True


def _ref_work(job):
    """One worker's share of a step: frames first .. n by stride (extract + SearchByProjection), then its share of the LBAs."""
    import numpy as np
    first, stride, n_frames, n_lba = job
    R = _REF
    O = R['O']
    for i in range(first, n_frames, stride):
        s = i % R['nsrc']
        _, k, d = R['ex'](R['f1'][s], (0, 1000))
        match = np.full(len(k), -1, np.int32)
        claimed = np.zeros(len(k), np.uint8)
        O.search_last_frame(k, d, (0.0, 0.0, float(W), float(H)), R['sf'], R['poses'][s], R['cam'], R['last'][s], TH_PROJ, True, match, claimed)
    for i in range(first, n_lba, stride):
        O.lba_solve(R['prob'])
    return True


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port; kind='port') on all host cores -- one worker PROCESS per core
    (threads would serialise on the interpreter lock in the numpy glue around the oracle calls)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    nsrc = 16
    n_frames = 4 * KF_INTERVAL * cores      # 40 frames + 4 LBAs per worker and step: about a second of work per core
    n_lba = n_frames // KF_INTERVAL
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_ref_init, initargs=(nsrc,)) as pool:
        jobs = [(w, cores, n_frames, n_lba) for w in range(cores)]
        warm = [(w, cores, KF_INTERVAL * cores, cores) for w in range(cores)]
        pool.map(_ref_work, warm, chunksize=1)                  # also makes sure every worker finished its initialiser
        for _ in range(max(args.warmup - 1, 0)):
            pool.map(_ref_work, warm, chunksize=1)
        t_all = 0.0
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_work, jobs, chunksize=1)
            t_all += time.perf_counter() - t0
    fps = n_frames * args.steps / t_all
    sample = '%d frames (extract+SearchByProjection) + %d LBAs (20 KF x 5000 pts x 40k edges) per step on %d worker processes' % (n_frames, n_lba, cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * t_all / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8 (extract/match), f64 (LBA)', 'data': 'synthetic', 'config': {'workload': WORKLOAD, 'stages': STAGES},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample, 'host': cpu_info()},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs the GPU is attached to (sysfs local_cpulist), so that pinned host memory is NUMA-local."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bus = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for tok in open('/sys/bus/pci/devices/%s/local_cpulist' % bus).read().strip().split(','):
            a, _, b = tok.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return '%s: %d cpus' % (bus, len(cpus))
    except (OSError, ValueError, AttributeError):
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='frames (streams) per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--dev-groups', type=int, default=4, help='stream groups (own handles + CUDA stream) in the device-resident measurement')
    ap.add_argument('--e2e-groups', type=int, default=2, help='stream groups (host threads with their own handles) in flight in the e2e measurement')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    import orb_slam3_modified_b200 as orb

    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU path in the product)')
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)     # pinned buffers (ours and the library's) land on the GPU's socket: ~53 vs ~20 GB/s over PCIe
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    B = args.batch
    NLBA = max(1, B // KF_INTERVAL)
    dev = torch.device('cuda', local)
    ex = orb.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, B, local)
    cap = ex.max_keypoints
    matcher = orb.ORBmatcher(0.9, True, max_batch=B, max_keypoints=cap, max_mappoints=cap, device=local)
    opt = orb.Optimizer(max_poses=LBA_CFG['n_kf'], max_points=LBA_CFG['n_pts'], max_edges=LBA_CFG['n_pts'] * LBA_CFG['obs_per_pt'],
                        max_batch=NLBA, device=local)
    sf = ex.GetScaleFactors()
    from orb_slam3_modified_b200 import synth
    cam = [float(c) for c in synth.camera(W, H)]

    # ---- inputs: two alternating sets (consecutive frames t / t+1 of every stream); step i extracts set i&1 and matches it
    # against the map points of set (i+1)&1.  2 x B x 307 KB (157 MB at B=256) > 126 MB L2, per-step working set ~1.3 GB.
    host_sets = [torch.from_numpy(make_frames(B, k, rank)).pin_memory() for k in range(2)]
    dev_sets = [h.to(dev) for h in host_sets]
    poses_h = [np.stack([stream_pose(s, k, rank) for s in range(B)]) for k in range(2)]
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_mono = torch.zeros(B, dtype=torch.int32, device=dev)
    d_match = torch.full((B, cap), -1, dtype=torch.int32, device=dev)
    d_claimed = torch.zeros((B, cap), dtype=torch.uint8, device=dev)
    d_nmatch = torch.zeros(B, dtype=torch.int32, device=dev)
    # untimed set-up: features of both sets -> last-frame map points (the map state the tracker would already hold)
    last_h, last_d = [], []
    for k in range(2):
        monos, kl, dl = ex.extract_batch(host_sets[k].numpy(), (0, 1000))
        slabs = last_frame_slabs(kl, dl, k, cap, rank)
        last_h.append(slabs)
        last_d.append({n: torch.from_numpy(v).to(dev) for n, v in slabs.items()})
    d_sf = torch.from_numpy(sf).to(dev)
    d_Tcw = [torch.from_numpy(p).to(dev) for p in poses_h]
    probs = lba_problems(NLBA, rank)
    opt.upload(probs)                      # flattened graphs resident in HBM for the `value` measurement
    from orb_slam3_modified_b200 import sharding
    gather = sharding.SlabGather(dist, world, d_kps, d_desc, d_n) if world > 1 else None
    stream = torch.cuda.current_stream()
    lba_stream = torch.cuda.Stream(device=dev)      # LocalMapping runs beside Tracking in the reference (src/System.cc:197)
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    # The B streams are served as DG groups, each with its own extractor / matcher handles on its own CUDA stream: the latency-bound
    # kernels of one group (quadtree, ordered commit) overlap the throughput-bound ones of the other (FAST, blur).
    DG = max(1, min(args.dev_groups, B))
    gb = [(g * B // DG, (g + 1) * B // DG) for g in range(DG)]
    if DG == 1:
        g_ex, g_mt, g_st = [ex], [matcher], [stream]
    else:
        g_ex = [orb.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, b1 - b0, local) for b0, b1 in gb]
        g_mt = [orb.ORBmatcher(0.9, True, max_batch=b1 - b0, max_keypoints=cap, max_mappoints=cap, device=local) for b0, b1 in gb]
        g_st = [torch.cuda.Stream(device=dev) for _ in gb]
    g_ev = [torch.cuda.Event() for _ in gb]

    def match_args(i, b0=0, b1=None):
        b1 = B if b1 is None else b1
        cur, lst = i & 1, (i + 1) & 1
        L = last_d[lst]
        return dict(batch=b1 - b0, kcap=cap, mcap=cap, nlevels=8, kps=d_kps[b0:b1], desc=d_desc[b0:b1], nK=d_n[b0:b1], scaleFactors=d_sf,
                    nM=L['nM'][b0:b1], valid=L['valid'][b0:b1], xyz=L['xyz'][b0:b1], octave=L['octave'][b0:b1], angle=L['angle'][b0:b1],
                    hasObs=L['hasObs'][b0:b1], mpDesc=L['mpDesc'][b0:b1], Tcw7=d_Tcw[cur][b0:b1], bounds=(0.0, 0.0, float(W), float(H)), cam=cam, reset=1)

    def step_device(i):
        # LocalMapping is asynchronous to Tracking in the reference (own thread, src/System.cc:197): the LBAs of step i are
        # enqueued on their own stream and only joined at the end of the timed region (all of them finish inside it).
        opt.run_device(lba_stream.cuda_stream)                       # B / KF_INTERVAL LBAs, one persistent kernel
        for g, (b0, b1) in enumerate(gb):
            st = g_st[g]
            g_ex[g].extract_batch_device(dev_sets[i & 1][b0:b1], d_kps[b0:b1], d_desc[b0:b1], d_n[b0:b1], d_mono[b0:b1], (0, 1000), st.cuda_stream)
            g_mt[g].search_last_frame_batch_device(match_args(i, b0, b1), TH_PROJ, d_match[b0:b1], d_claimed[b0:b1], d_nmatch[b0:b1], st.cuda_stream)
        if world > 1:   # shared-map exchange: one all-gather of the fixed-capacity keypoint/descriptor slabs
            join_groups()
            gather(d_kps, d_desc, d_n)
            fork_groups()        # the next step's kernels overwrite the slabs: after the gather

    def fork_groups():
        if DG > 1:
            ev_fork.record(stream)
            for st in g_st:
                st.wait_event(ev_fork)

    def join_groups():
        if DG > 1:
            for g, st in enumerate(g_st):
                g_ev[g].record(st)
                stream.wait_event(g_ev[g])

    def join_lba():
        ev_join.record(lba_stream)
        stream.wait_event(ev_join)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`) ----------------
    clocks = ClockSampler(local) if rank == 0 else None      # started before the warm-up: nvidia-smi needs ~0.2 s to deliver its first row
    for i in range(args.warmup):
        step_device(i)
    join_groups()
    join_lba()
    barrier()
    if clocks:
        clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_fork.record(stream)
    lba_stream.wait_event(ev_fork)
    e0.record(stream)
    fork_groups()
    for i in range(args.steps):
        step_device(i)
    join_groups()
    join_lba()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (sum(e.last_launch_count() for e in g_ex) + sum(m_.last_launch_count() for m_ in g_mt) + 1) * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if clocks and clocks.samples() == 0 and world == 1:
        # the timed region was shorter than the sampling period: sample the same load for a moment more (untimed), and say so
        t_end = time.perf_counter() + 0.4
        i = 0
        while time.perf_counter() < t_end and clocks.samples() < 2:
            step_device(i)
            join_groups()
            join_lba()
            torch.cuda.synchronize()
            i += 1
    clk = clocks.stop() if clocks else None
    if clk is not None:
        clk['sampled'] = 'timed region (+ an untimed continuation of the same load when it was shorter than the 50 ms sampling period)'
    value = world * B * args.steps / (ms * 1e-3)
    mean_kp = float(d_n.float().mean().item())
    mean_matches = float(d_nmatch.float().mean().item())

    # ---------------- per-stage split (events on the launching stream) + roofline of the dominant kernel ----------------
    stage = {}
    if rank == 0:
        ex.set_profiling(True)
        acc = {}
        reps = 3
        for i in range(reps):
            ex.extract_batch_device(dev_sets[i & 1], d_kps, d_desc, d_n, d_mono, (0, 1000), stream.cuda_stream)
            for k, v in ex.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        ex.set_profiling(False)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for i in range(reps):
            matcher.search_last_frame_batch_device(match_args(i), TH_PROJ, d_match, d_claimed, d_nmatch, stream.cuda_stream)
        a1.record(stream)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for i in range(reps):
            opt.run_device(stream.cuda_stream)
        b1.record(stream)
        torch.cuda.synchronize()
        stage = {k: v for k, v in acc.items()}
        stage['match(3 kernels)'] = a0.elapsed_time(a1) / reps
        stage['lba_cluster_kernel'] = b0.elapsed_time(b1) / reps
        lba_out = opt.download()
        mean_trials = float(np.mean([o['trials'] for o in lba_out]))
    torch.cuda.synchronize()

    # ---------------- end to end through the host C-ABI (`e2e`) ----------------
    # Host buffers in, host buffers out, every step.  The B streams are served as G groups, each by its own host thread with
    # its own extractor / matcher handles (the batch calls are synchronous, so two groups in flight let one group's PCIe
    # copies overlap the other's kernels), and the local maps' bundle adjustments by G mapping threads (the LocalMapping
    # thread of the reference, src/System.cc:197).  A step is complete when every group has its host results.
    e2e, e2e_steps, h2d, d2h = None, 0, 0, 0
    e2e_err = None
    if not args.no_e2e:
        try:
            from concurrent.futures import ThreadPoolExecutor
            G = max(1, min(args.e2e_groups, B, NLBA))
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
            kps_h = pin((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orb.KP_DTYPE).reshape(B, cap)
            desc_h = pin((B, cap, 32), torch.uint8)
            nK_h, mono_h, nmatch_h = pin((B,), torch.int32), pin((B,), torch.int32), pin((B,), torch.int32)
            match_h, claimed_h = pin((B, cap), torch.int32), pin((B, cap), torch.uint8)
            bounds = [(g * B // G, (g + 1) * B // G) for g in range(G)]
            lbounds = [(g * NLBA // G, (g + 1) * NLBA // G) for g in range(G)]
            if G == 1:
                exs, mts, opts = [ex], [matcher], [opt]
            else:
                exs = [orb.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, b1 - b0, local) for b0, b1 in bounds]
                mts = [orb.ORBmatcher(0.9, True, max_batch=b1 - b0, max_keypoints=cap, max_mappoints=cap, device=local) for b0, b1 in bounds]
                opts = [orb.Optimizer(max_poses=LBA_CFG['n_kf'], max_points=LBA_CFG['n_pts'], max_edges=LBA_CFG['n_pts'] * LBA_CFG['obs_per_pt'],
                                      max_batch=l1 - l0, device=local) for l0, l1 in lbounds]
            pool = ThreadPoolExecutor(2 * G)

            trace = [] if os.environ.get('BENCH_E2E_TRACE') else None

            def frames_job(g, i):
                cur, lst = i & 1, (i + 1) & 1
                b0, b1 = bounds[g]
                tq = time.perf_counter()
                exs[g].extract_batch_slabs(host_sets[cur].numpy()[b0:b1], kps_h[b0:b1], desc_h[b0:b1], nK_h[b0:b1], mono_h[b0:b1], (0, 1000))
                L = last_h[lst]
                d = dict(batch=b1 - b0, kcap=cap, mcap=cap, nlevels=8, kps=kps_h[b0:b1], desc=desc_h[b0:b1], nK=nK_h[b0:b1], scaleFactors=sf,
                         nM=L['nM'][b0:b1], valid=L['valid'][b0:b1], xyz=L['xyz'][b0:b1], octave=L['octave'][b0:b1], angle=L['angle'][b0:b1],
                         hasObs=L['hasObs'][b0:b1], mpDesc=L['mpDesc'][b0:b1], Tcw7=poses_h[cur][b0:b1], bounds=(0.0, 0.0, float(W), float(H)), cam=cam, reset=1)
                tm = time.perf_counter()
                mts[g].search_last_frame_batch(d, TH_PROJ, match_h[b0:b1], claimed_h[b0:b1], nmatch_h[b0:b1])
                if trace is not None:
                    trace.append(('frames', g, i, tq, tm, time.perf_counter()))
                return int(nK_h[b0:b1].sum())

            def lba_job(g):
                l0, l1 = lbounds[g]
                tq = time.perf_counter()
                r = opts[g].LocalBundleAdjustmentBatch(probs[l0:l1])        # host graphs in, optimised state out
                if trace is not None:
                    trace.append(('lba', g, -1, tq, tq, time.perf_counter()))
                return r

            pending = []          # bundle adjustments of the previous step: LocalMapping runs beside Tracking, one step behind

            def step_host(i):
                nonlocal pending
                lba = [pool.submit(lba_job, g) for g in range(G)]
                fr = [pool.submit(frames_job, g, i) for g in range(G)]
                nk = sum(j.result() for j in fr)
                outs = [j.result() for j in pending]
                pending = lba
                return nk, outs

            step_host(0)
            step_host(1)
            barrier()
            e2e_steps = max(2, min(args.steps, 8))
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                nk, outs = step_host(i)
            torch.cuda.synchronize()
            # every step of the timed region submitted one batch of bundle adjustments and collected one (the one submitted a
            # step earlier); the batches still in flight are collected outside the timed region
            dt = time.perf_counter() - t0
            outs = [j.result() for j in pending]
            if trace is not None and rank == 0:
                for kind, g, i, a, m, b in sorted(trace, key=lambda r: r[3])[-6 * G:]:
                    print('# %-6s g%d step %2d start %8.2f ms  first call %6.2f  total %6.2f' % (kind, g, i, (a - t0) * 1e3, (m - a) * 1e3, (b - a) * 1e3), file=sys.stderr)
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e = world * B * e2e_steps / float(t.item())
            p0 = probs[0]
            lba_h2d = NLBA * (p0['poses'].nbytes + p0['points'].nbytes + p0['obs'].nbytes + 3 * 4 * len(p0['edge_point']) + p0['cam'].nbytes)
            lba_d2h = NLBA * (p0['poses'].nbytes + p0['points'].nbytes + 9 * len(p0['edge_point']))
            h2d = B * W * H + B * cap * (28 + 32 + 5) + sum(v.nbytes for v in last_h[0].values()) + lba_h2d
            d2h = nk * 60 + 12 * B + B * cap * 5 + lba_d2h
            pool.shutdown()
        except Exception as exc:     # the device-resident result is still reported (single GPU); with several ranks a failure must stay fatal
            if world > 1:
                raise
            e2e, e2e_err = None, repr(exc)[:300]

    if rank == 0:
        peak, how = _peaks()
        traffic_tab = {}
        try:
            traffic_tab = json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))
        except (OSError, ValueError):
            pass

        def alg_of(k):
            if k == 'lba_cluster_kernel':
                return ALG_BYTES_LBA_PER_TRIAL * mean_trials * NLBA, 'LBA launch: %.1f LM trials x 13.6 MB x %d problems' % (mean_trials, NLBA)
            if k.startswith('match'):
                return ALG_BYTES_MATCH * B, '552,000 B/frame x %d frames' % B
            return ALG[k] * B, '%d B/frame x %d frames' % (ALG[k], B)

        def traffic_of(k):
            t = traffic_tab.get(k)
            # the table holds ncu DRAM bytes per launch of THIS workload (256 frames / 25 problems per launch); scale if --batch differs
            if not t:
                return None
            return t['bytes'] * (NLBA / 25.0 if k == 'lba_cluster_kernel' else B / 256.0)

        per_kernel = {}
        for k in stage:
            ab, _ = alg_of(k)
            per_kernel[k] = {'achieved': ab / (stage[k] * 1e-3) / 1e9, 'frac': ab / (stage[k] * 1e-3) / 1e9 / peak, 'traffic': traffic_of(k)}
        top = max(stage, key=stage.get)
        alg_bytes, per = alg_of(top)
        ach = alg_bytes / (stage[top] * 1e-3) / 1e9
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8 (extract/match), f64 (LBA)', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'stages': STAGES, 'frames_per_gpu_per_step': B, 'lba_per_gpu_per_step': NLBA,
                       'l2': 'inputs alternate between two %d-frame sets (2 x %.0f MB) > 126 MB L2; per-step working set > 1 GB' % (B, B * W * H / 1e6),
                       'mean_keypoints_per_frame': mean_kp, 'mean_matches_per_frame': mean_matches, 'lba_cluster_size': opt.last_cluster_size(),
                       'lba_mean_trials': mean_trials, 'host_numa_binding': numa, 'device_stream_groups': DG},
            'clocks': clk, 'gpu_launches': launches,
            'stage_ms_per_step': stage,
            'roofline': {'bound': 'hbm', 'kernel': top, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic_of(top),
                         'traffic_source': (traffic_tab.get(top) or {}).get('source'),
                         'peak_source': how, 'algorithmic_bytes_per_launch': alg_bytes, 'per': per, 'launch_ms': stage[top],
                         'note': 'FP64 LM solver bound by L1 look-ups of per-lane gathers and by cluster barriers, not by DRAM (DESIGN.md 5); frac is against the HBM copy peak as the contract asks',
                         'per_kernel': per_kernel,
                         'whole_step_frac': (ALG_BYTES_EXTRACT + ALG_BYTES_MATCH + ALG_BYTES_LBA_PER_TRIAL * mean_trials / KF_INTERVAL) * (value / world) / 1e9 / peak},
        }
        if e2e_err:
            out['e2e'], out['e2e_error'] = None, e2e_err
        if e2e is not None:
            out['e2e'] = {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': e2e_steps,
                          'stream_groups': max(1, min(args.e2e_groups, B, NLBA))}
        if world == 1 and not args.no_cpu_baseline:
            nsrc = 8
            f0, f1 = host_sets[0].numpy()[:nsrc], host_sets[1].numpy()[:nsrc]
            n_cpu = 64 * KF_INTERVAL                 # ~15-20 s of single-thread CPU work
            try:
                fps1, dt1 = cpu_oracle_mix(f0, f1, [poses_h[1][s] for s in range(nsrc)], probs[0], n_cpu, 1)
                out['cpu_baseline'] = {'value': fps1, 'unit': UNIT, 'cores': 1, 'kind': 'port', 'host': cpu_info(),
                                       'sample': '%d frames extract+SearchByProjection + %d LBA, oracle, 1 thread (%.1f s)' % (n_cpu, n_cpu // KF_INTERVAL, dt1)}
            except Exception as exc:
                out['cpu_baseline'] = None
                out['cpu_baseline_error'] = repr(exc)[:300]
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
