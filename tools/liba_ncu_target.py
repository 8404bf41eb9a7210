"""One LocalInertialBA launch for ncu / compute-sanitizer: python tools/liba_ncu_target.py [maps] [small]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import orb_slam3_modified_b200 as orb   # noqa: E402
from orb_slam3_modified_b200 import synth   # noqa: E402
import oracle_lib as O   # noqa: E402

maps = int(sys.argv[1]) if len(sys.argv) > 1 else 74
small = len(sys.argv) > 2
pr = synth.local_inertial_ba_problem(n_opt=4, n_cov_fixed=2, n_pts=80, seed=3) if small else synth.local_inertial_ba_problem(n_opt=10, n_cov_fixed=6, n_pts=2500, seed=12)
pr['preint'] = O.liba_preints(pr)
r = orb.LocalInertialBA([pr] * maps)
print(maps, 'maps', r[0]['iters'], 'iterations', r[0]['trials'], 'trials', '%.2f ms solver' % r[0]['kernel_ms'])
