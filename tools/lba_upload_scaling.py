import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import bench as Bn
import orb_slam3_modified_b200 as orb
probs = Bn.lba_problems(32)
opt = orb.Optimizer(20,5000,40000,max_batch=32)
for n in (1, 2, 8, 16, 25, 32):
    opt.upload(probs[:n]); torch.cuda.synchronize()
    ts=[]
    for _ in range(3):
        t0=time.perf_counter(); opt.upload(probs[:n]); ts.append((time.perf_counter()-t0)*1e3)
    print('upload %2d problems: %.2f ms' % (n, min(ts)))
