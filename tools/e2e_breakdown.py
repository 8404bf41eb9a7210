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
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0=T(); monos,kl,dl=ex.extract_batch(host[it&1].numpy(),(0,1000)); t1=T()
    kps_h=np.zeros((B,cap),orb.KP_DTYPE); desc_h=np.zeros((B,cap,32),np.uint8); nK=np.zeros(B,np.int32)
    for b in range(B): nK[b]=len(kl[b]); kps_h[b,:nK[b]]=kl[b]; desc_h[b,:nK[b]]=dl[b]
    t2=T()
    L=last[(it+1)&1]
    d=dict(batch=B,kcap=cap,mcap=cap,nlevels=8,kps=kps_h,desc=desc_h,nK=nK,scaleFactors=sf,nM=L['nM'],valid=L['valid'],xyz=L['xyz'],octave=L['octave'],angle=L['angle'],hasObs=L['hasObs'],mpDesc=L['mpDesc'],Tcw7=poses[it&1],bounds=(0.,0.,640.,480.),cam=cam,reset=1)
    mh=np.full((B,cap),-1,np.int32); ch=np.zeros((B,cap),np.uint8); nm=np.zeros(B,np.int32)
    matcher.search_last_frame_batch(d,15.0,mh,ch,nm); t3=T()
    outs=opt.LocalBundleAdjustmentBatch(probs); t4=T()
    print('extract_batch %.1f ms | python slab copy %.1f | match_batch %.1f | lba_batch %.1f | total %.1f'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t4-t0)*1e3))
