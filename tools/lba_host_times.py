"""Host wall-clock of each call of the batch LBA pipeline (72 config-4 problems): upload (pack + H2D + structure kernel), run, download."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import orb_slam3_modified_b200 as orb
from orb_slam3_modified_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 72
base = [synth.lba_problem(seed=i, n_kf=20, n_pts=5000, obs_per_pt=8) for i in range(4)]
probs = [base[i % 4] for i in range(n)]
opt = orb.Optimizer(20, 5000, 40000, max_batch=n)
st = torch.cuda.Stream()
T = time.perf_counter
for it in range(4):
    torch.cuda.synchronize()
    t0 = T(); opt.upload(probs); t1 = T(); torch.cuda.synchronize(); t2 = T()
    opt.run_device(st.cuda_stream); t3 = T(); torch.cuda.synchronize(); t4 = T()
    out = opt.download(); t5 = T()
    print('upload call %.2f ms (+%.2f until device idle) | run call %.3f ms, kernel done after %.2f ms | download %.2f ms' %
          (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4)), flush=True)
