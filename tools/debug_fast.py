import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, oracle_lib as O
import orb_slam3_modified_b200 as orb
from orb_slam3_modified_b200 import synth
im=synth.frame(0)
ex=orb.ORBextractor(1000,1.2,8,20,7,640,480,1)
oe=O.OracleExtractor()
ex(im,(0,1000)); oe(im,(0,1000))
for l in (0,7):
    c=ex.candidates(l); oc=oe.candidates(l)
    g=set(map(tuple,c.tolist())); o=set((int(a),int(b),int(r)) for a,b,r in zip(oc['x'],oc['y'],oc['response']))
    print('level',l,'gpu',len(g),'oracle',len(o),'common',len(g&o),'extra',len(g-o),'missing',len(o-g))
    gxy={(a,b):s for a,b,s in g}; oxy={(a,b):s for a,b,s in o}
    ex_=sorted(g-o)[:15]; print(' extra sample',ex_)
    print(' missing sample',sorted(o-g)[:15])
    # adjacency of extras to other gpu candidates
    adj=0
    for (a,b,s) in (g-o):
        for dx in (-1,0,1):
            for dy in (-1,0,1):
                if (dx or dy) and (a+dx,b+dy) in gxy: adj+=1
    print(' extras adjacent-to-candidate count',adj)
    same_xy_diff_score=sum(1 for (a,b,s) in (g-o) if (a,b) in oxy)
    print(' extras with same xy in oracle but different score',same_xy_diff_score)
    lv=oe.level(l)
    # full-res oracle score check for a few extras: use orbo_fast on a window
    for (a,b,s) in ex_[:5]:
        x=a+16; y=b+16
        roi=np.ascontiguousarray(lv[y-4:y+5, x-4:x+5])
        print('  extra',(a,b,s),'oracle window fast T=7:',O.fast(roi,7).tolist())
