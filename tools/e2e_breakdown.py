"""Where the end-to-end (host buffers in, host buffers out) step of bench.py spends its time: each C-ABI call timed alone."""
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import bench as Bn
import orb_slam3_modified_b200 as orb
from orb_slam3_modified_b200 import synth
B=256; W,H=640,480
ex = orb.ORBextractor(1000,1.2,8,20,7,W,H,B,0); cap=ex.max_keypoints
matcher = orb.ORBmatcher(0.9,True,max_batch=B,max_keypoints=cap,max_mappoints=cap)
NL=B//10
opt = orb.Optimizer(20,5000,40000,max_batch=NL)
host=[torch.from_numpy(Bn.make_frames(B,k)).pin_memory() for k in range(2)]
sf=ex.GetScaleFactors(); cam=[float(c) for c in synth.camera(W,H)]
poses=[np.stack([Bn.stream_pose(s,k) for s in range(B)]) for k in range(2)]
last=[]
for k in range(2):
    m,kl,dl=ex.extract_batch(host[k].numpy(),(0,1000)); last.append(Bn.last_frame_slabs(kl,dl,k,cap))
probs=Bn.lba_problems(NL)
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
kps_h = pin((B, cap, 7), torch.float32).view(np.uint8).reshape(B, cap, 28).view(orb.KP_DTYPE).reshape(B, cap)
desc_h = pin((B, cap, 32), torch.uint8)
nK_h, mono_h, nmatch_h = pin((B,), torch.int32), pin((B,), torch.int32), pin((B,), torch.int32)
match_h, claimed_h = pin((B, cap), torch.int32), pin((B, cap), torch.uint8)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    t0=T(); ex.extract_batch_slabs(host[it&1].numpy(), kps_h, desc_h, nK_h, mono_h, (0,1000)); t1=T()
    L=last[(it+1)&1]
    d=dict(batch=B,kcap=cap,mcap=cap,nlevels=8,kps=kps_h,desc=desc_h,nK=nK_h,scaleFactors=sf,nM=L['nM'],valid=L['valid'],xyz=L['xyz'],octave=L['octave'],angle=L['angle'],hasObs=L['hasObs'],mpDesc=L['mpDesc'],Tcw7=poses[it&1],bounds=(0.,0.,640.,480.),cam=cam,reset=1)
    matcher.search_last_frame_batch(d,15.0,match_h,claimed_h,nmatch_h); t2=T()
    ta=time.perf_counter(); opt.upload(probs); tb=T(); opt.run_device(); tc=T(); outs=opt.download(); td=T()
    print('extract_batch_slabs %.2f ms | match_batch %.2f | lba upload(pack+H2D) %.2f run %.2f download %.2f | sum %.2f'%((t1-t0)*1e3,(t2-t1)*1e3,(tb-ta)*1e3,(tc-tb)*1e3,(td-tc)*1e3,(td-t0)*1e3))
