#!/usr/bin/env python3
"""bench.py -- frames/sec of the ORB-SLAM3 hot path (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic 640x480 frames per GPU
(BASELINE configs[1]: 640x480 mono stream, 1000 feats/frame).  `value` is whole-job frames/s with the
frames already resident in HBM; `e2e` is the same metric through the reference-facing C-ABI with
HOST buffers (pinned), H2D/D2H inside the timed region.  --impl reference times the CPU oracle
(the reference itself cannot be built here: needs OpenCV C++/Eigen) on all host cores.
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
# algorithmic bytes per extracted frame = 5S - A0 - A7 + 1321*K, 640x480, K=1000
ALG_BYTES_EXTRACT = 5742474


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
                 'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap',
                 '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


DISTINCT = 32   # distinct synthetic streams; larger batches replicate them (separate buffers, same content)
TH_PROJ = 15.0  # SearchByProjection window for mono tracking (reference src/Tracking.cc:2884-2889)


def stream_time(s, k):
    """Frame index of stream s at parity k (0: 'previous', 1: 'next')."""
    return 5 * s + k


def make_frames(n, k, rank=0):
    """Frame k (0/1) of n streams: stream s shows synth.frame(t = 5 s + k, seed = s % 4 + 4 rank)."""
    from orb_slam3_modified_b200 import synth
    import numpy as np
    d = min(n, DISTINCT)
    base = np.stack([synth.frame(stream_time(s, k), W, H, seed=s % 4 + 4 * rank) for s in range(d)])
    return np.ascontiguousarray(np.concatenate([base] * ((n + d - 1) // d))[:n])


def stream_pose(s, k, rank=0, noise=0.003):
    """Motion-model prior of Tcw for stream s at parity k: the exact synthetic pose plus a seeded perturbation."""
    from orb_slam3_modified_b200 import synth
    import numpy as np
    rng = np.random.default_rng(1000 * rank + 2 * s + k)
    T = synth.pose(stream_time(s % DISTINCT, k), seed=(s % DISTINCT) % 4 + 4 * rank)
    T[4:] += rng.normal(0, noise, 3)
    return T.astype(np.float32)


def cpu_oracle_throughput(frames, seconds_budget, threads):
    """frames/s of the CPU oracle (port of the reference path) on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    exs = [O.OracleExtractor(NFEAT, 1.2, 8, 20, 7) for _ in range(threads)]
    exs[0](frames[0], (0, 1000))
    t0 = time.perf_counter()
    exs[0](frames[0], (0, 1000))
    per = max(time.perf_counter() - t0, 1e-4)
    n = max(threads, min(len(frames) * 4, int(seconds_budget / per) * threads))
    n -= n % threads

    def work(k):
        e = exs[k]
        for i in range(k, n, threads):
            e(frames[i % len(frames)], (0, 1000))

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(work, range(threads)))
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    """Reference arm: the CPU implementation of the path on all host cores (oracle port; kind='port')."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = make_frames(32, seed=0)
    t_all, n_all = 0.0, 0
    for _ in range(args.warmup):
        cpu_oracle_throughput(frames, 0.5, cores)
    for _ in range(args.steps):
        fps, n, dt = cpu_oracle_throughput(frames, 3.0, cores)
        t_all += dt
        n_all += n
    fps = n_all / t_all
    sample = '%d frames of 640x480 per step on %d threads (extract only; matcher/LBA stages as in the GPU arm are added as they land)' % (n_all // max(args.steps, 1), cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * t_all / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u8', 'data': 'synthetic', 'config': {'workload': 'configs[1]: 640x480 mono stream, 1000 feats/frame', 'stages': STAGES},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


STAGES = ['extract']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='frames (streams) per GPU per step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    B = args.batch
    dev = torch.device('cuda', local)
    ex = orb.ORBextractor(NFEAT, 1.2, 8, 20, 7, W, H, B, local)
    cap = ex.max_keypoints

    # two alternating input sets so that consecutive steps never re-read the same frames from L2
    # (2 x B x 307 KB; with B=256 that is 157 MB > the 126 MB L2, and the per-step working set is ~1.3 GB)
    host_sets = [torch.from_numpy(make_frames(B, seed=rank * 2 + s)).pin_memory() for s in range(2)]
    dev_sets = [h.to(dev) for h in host_sets]
    d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_mono = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = None
    if world > 1:
        gathered = [torch.empty((world,) + t.shape, dtype=t.dtype, device=dev) for t in (d_kps, d_desc, d_n)]
    stream = torch.cuda.current_stream()

    def step_device(i):
        ex.extract_batch_device(dev_sets[i & 1], d_kps, d_desc, d_n, d_mono, (0, 1000), stream.cuda_stream)
        if world > 1:   # shared-map exchange: one all-gather of the fixed-capacity slabs
            for src, dst in zip((d_kps, d_desc, d_n), gathered):
                dist.all_gather_into_tensor(dst, src)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`) ----------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ex.last_launch_count() * args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clk = clocks.stop() if clocks else None
    value = world * B * args.steps / (ms * 1e-3)
    mean_kp = float(d_n.float().mean().item())

    # ---------------- end to end through the host C-ABI (`e2e`) ----------------
    def step_host(i):
        return ex.extract_batch(host_sets[i & 1].numpy(), (0, 1000))

    for i in range(max(1, min(args.warmup, 2))):
        step_host(i)
    barrier()
    e2e_steps = max(2, min(args.steps, 6))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        monos, kps, descs = step_host(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * B * e2e_steps / float(t.item())
    d2h = int(sum(len(k) for k in kps) * 60 + 12 * B)

    if rank == 0:
        peak, how = _peaks()
        ach = ALG_BYTES_EXTRACT * (value / world) / 1e9
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: 640x480 mono stream, 1000 feats/frame', 'stages': STAGES, 'frames_per_gpu_per_step': B,
                       'l2': 'inputs alternate between two %d-frame sets (2 x %.0f MB) > 126 MB L2' % (B, B * W * H / 1e6),
                       'mean_keypoints_per_frame': mean_kp},
            'clocks': clk, 'gpu_launches': launches,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': B * W * H, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps},
            'roofline': {'bound': 'hbm', 'kernel': 'whole extract step (per-kernel split in profiles/)', 'achieved': ach, 'peak': peak,
                         'unit': 'GB/s', 'frac': ach / peak, 'traffic': None, 'peak_source': how,
                         'algorithmic_bytes_per_frame': ALG_BYTES_EXTRACT},
        }
        if world == 1 and not args.no_cpu_baseline:
            frames = host_sets[0].numpy()[:32]
            fps1, n1, dt1 = cpu_oracle_throughput(frames, 8.0, 1)
            out['cpu_baseline'] = {'value': fps1, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                   'sample': '%d frames of 640x480, oracle ORBextractor, 1 thread (the reference mono path is single-threaded)' % n1}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
