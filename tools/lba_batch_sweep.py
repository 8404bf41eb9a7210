"""Throughput of the persistent LBA kernel for different (problems per launch, cluster size) choices (config-4 problems)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import orb_slam3_modified_b200 as orb
from orb_slam3_modified_b200 import synth
base = [synth.lba_problem(seed=i, n_kf=20, n_pts=5000, obs_per_pt=8) for i in range(4)]
st = torch.cuda.current_stream()
for nprob, cs in ((24, 0), (24, 4), (24, 6), (48, 3), (48, 0), (72, 2), (144, 1), (148, 1), (296, 1)):
    opt = orb.Optimizer(20, 5000, 40000, max_batch=nprob)
    opt.upload([base[i % 4] for i in range(nprob)])
    opt.set_cluster_size(cs)
    try:
        for _ in range(2):
            opt.run_device(st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            opt.run_device(st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print('problems %3d cluster %d (used %d): %.3f ms/launch  %.4f ms/problem' % (nprob, cs, opt.last_cluster_size(), ms, ms / nprob), flush=True)
    except Exception as e:
        print('problems %d cluster %d failed: %r' % (nprob, cs, e), flush=True)
    opt.close()
