import sys, ctypes as C; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, time
import orb_slam3_modified_b200 as orb
from orb_slam3_modified_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
opt = orb.Optimizer(20, 5000, 40000, max_batch=n)
p = synth.lba_problem()
import torch
opt.upload([p]*n)
if len(sys.argv) > 2: opt.set_cluster_size(int(sys.argv[2]))
for _ in range(2): opt.run_device()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record(); opt.run_device(torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
out = opt.download()
ns = np.zeros(10)
L = orb.lib(); L.lba_get_phase_ns.argtypes=[C.c_void_p,C.c_int,C.c_void_p]
L.lba_get_phase_ns(opt._h, 0, ns.ctypes.data_as(C.c_void_p))
names=['errors(it0)','build_points','build_poses','point_prep','schur_partial','ldlt','schur_combine','pose_trial','points_trial','-']
print('batch',n,'cluster',opt.last_cluster_size(),'kernel ms',e0.elapsed_time(e1),'iters',out[0]['iters'],'trials',out[0]['trials'])
for k,v in zip(names,ns): print('  %-14s %8.1f us total  %7.1f us/trial'%(k,v/1e3,v/1e3/max(out[0]['trials'],1)))
