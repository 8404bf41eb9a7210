"""LocalInertialBA on the GPU: solver time per map (the CTA's own globaltimer span, stats8[6]) and host-API time, for one map and for one map
per SM in a single launch; the oracle port on one host thread beside it.  Run on a B200: python tools/liba_time.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import orb_slam3_modified_b200 as orb   # noqa: E402
from orb_slam3_modified_b200 import synth   # noqa: E402
import oracle_lib as O   # noqa: E402

for name, kw in (('10+7 KF, ~2100 pts, ~24.5k edges', dict(n_opt=10, n_cov_fixed=6, n_pts=2500, seed=12)),
                 ('10+4 KF, ~500 pts, ~5k edges', dict(n_opt=10, n_cov_fixed=3, n_pts=600, seed=1)),
                 ('bLarge 25+5 KF, ~1300 pts, ~25k edges', dict(n_opt=25, n_cov_fixed=4, n_pts=1500, seed=11, large=True))):
    pr = synth.local_inertial_ba_problem(**kw)
    pr['preint'] = O.liba_preints(pr)
    t0 = time.perf_counter(); w = O.local_inertial_ba(pr, pr['preint']); t_cpu = time.perf_counter() - t0
    orb.LocalInertialBA([pr])
    ts, ks = [], []
    for _ in range(5):
        t0 = time.perf_counter(); g = orb.LocalInertialBA([pr])[0]; ts.append(time.perf_counter() - t0); ks.append(g['kernel_ms'])
    many = [pr] * 148
    orb.LocalInertialBA(many)
    t0 = time.perf_counter(); gm = orb.LocalInertialBA(many); t_many = time.perf_counter() - t0
    print('%s: %d edges, %d iterations / %d trials | 1 map: solver %.2f ms, host API %.2f ms | 148 maps in one launch: solver %.2f ms per CTA, host API %.1f ms '
          '(%.3f ms per map) | oracle, 1 thread: %.1f ms | 1-map phases [errors, build, Dinv/Y, Schur, LDLT, points, update, rest] ms: %s' % (name, len(pr['e_pt']), g['iters'], g['trials'], np.median(ks), 1e3 * np.median(ts),
                                                          np.median([x['kernel_ms'] for x in gm]), 1e3 * t_many, 1e3 * t_many / 148, 1e3 * t_cpu, np.round(g['phase_ms'], 2).tolist()))
