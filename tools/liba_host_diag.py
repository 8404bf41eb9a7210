"""Where does the host-API time of local_inertial_ba_batch go?  20 consecutive 1-map calls (Python wall time, the library's own ORB_LIBA_TRACE split, the
solver's globaltimer span), with and without torch imported and with the cyclic GC off.  python tools/liba_host_diag.py [torch] [nogc]"""
import gc
import os
import sys
import time

import numpy as np

os.environ['ORB_LIBA_TRACE'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
if 'torch' in sys.argv:
    import torch
    torch.zeros(1, device='cuda')
if 'nogc' in sys.argv:
    gc.disable()
import orb_slam3_modified_b200 as orb   # noqa: E402
from orb_slam3_modified_b200 import synth   # noqa: E402
import oracle_lib as O   # noqa: E402

pr = synth.local_inertial_ba_problem(n_opt=10, n_cov_fixed=6, n_pts=2500, seed=12)
pr['preint'] = O.liba_preints(pr)
big = synth.local_inertial_ba_problem(n_opt=25, n_cov_fixed=4, n_pts=1500, seed=11, large=True)
big['preint'] = O.liba_preints(big)
for name, p in (('10+7', pr), ('bLarge', big), ('10+7 again', pr)):
    for k in range(8):
        t0 = time.perf_counter(); g = orb.LocalInertialBA([p])[0]; t = time.perf_counter() - t0
        print('%s call %d: python %.2f ms, solver %.2f ms' % (name, k, 1e3 * t, g['kernel_ms']), flush=True)
    if name == '10+7':
        orb.LocalInertialBA([p] * 148)
