"""GPU parity of Optimizer::LocalInertialBA (reference src/Optimizer.cc:2383-2958) through the C-ABI
(local_inertial_ba_batch) against the CPU oracle (oracle/local_inertial_ba_oracle.cpp, pinned by a dense Levenberg step from numerical
derivatives in tests/test_local_inertial_ba_cpu.py): same optimize() iteration count, LM trials, FAIL flag and erase flags, err / err_end / lambda to
1e-6 relative, reprojection residuals within the north star's 1e-4 px, keyframe states within 1e-8."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth
from test_local_inertial_ba_cpu import _same_solve

pytestmark = pytest.mark.gpu

CASES = [dict(n_opt=10, n_cov_fixed=3, n_pts=400, seed=1), dict(n_opt=14, n_cov_fixed=2, n_pts=300, seed=3, large=True), dict(n_opt=1, n_cov_fixed=2, n_pts=60, seed=5),
         dict(n_opt=4, n_cov_fixed=0, n_pts=120, seed=6, rec_init=True), dict(n_opt=6, n_cov_fixed=1, n_pts=200, seed=7, perturb=4.0),
         dict(n_opt=25, n_cov_fixed=4, n_pts=1500, seed=11, large=True), dict(n_opt=10, n_cov_fixed=6, n_pts=2500, seed=12)]


@pytest.fixture(scope='module')
def orb():
    import orb_slam3_modified_b200 as m
    m.lib()
    return m


def _problems(cases):
    prs = []
    for kw in cases:
        pr = synth.local_inertial_ba_problem(**kw)
        pr['preint'] = O.liba_preints(pr)
        prs.append(pr)
    return prs


def test_batch_of_different_local_maps_equals_the_oracle(orb):
    prs = _problems(CASES)
    got = orb.LocalInertialBA(prs)
    for pr, g in zip(prs, got):
        _same_solve(g, O.local_inertial_ba(pr, pr['preint']), pr)


def test_lambda_init_branch_rejected_steps_and_reproducibility(orb):
    a = _problems([dict(n_opt=5, n_cov_fixed=1, n_pts=150, seed=8)])[0]
    a['lambda_init'] = 0.0                                   # tau * max diagonal
    b = _problems([dict(n_opt=5, n_cov_fixed=1, n_pts=150, seed=9, perturb=12.0)])[0]
    b['lambda_init'] = 1e-12                                 # far-off start, (almost) undamped: rejected trials
    got = orb.LocalInertialBA([a, b, a, b])
    wa, wb = O.local_inertial_ba(a, a['preint']), O.local_inertial_ba(b, b['preint'])
    assert wb['trials'] > wb['iters']
    _same_solve(got[0], wa, a)
    _same_solve(got[1], wb, b, tol_state=1e-7)
    for k in ('state', 'tcw', 'points', 'erase', 'chi2'):    # ordered sums: the same map twice in one launch gives the same bytes
        assert got[0][k].tobytes() == got[2][k].tobytes() and got[1][k].tobytes() == got[3][k].tobytes()


def test_many_maps_in_one_launch(orb):
    prs = _problems([dict(n_opt=10, n_cov_fixed=2, n_pts=300, seed=100 + s) for s in range(6)])
    got = orb.LocalInertialBA(prs * 30)                      # 180 CTAs: more than one wave of the 148 SMs
    for i, pr in enumerate(prs):
        want = O.local_inertial_ba(pr, pr['preint'])
        _same_solve(got[i], want, pr)
        for rep in range(1, 30):
            assert got[i + 6 * rep]['state'].tobytes() == got[i]['state'].tobytes() and got[i + 6 * rep]['iters'] == got[i]['iters']


def test_argument_errors(orb):
    pr = _problems([dict(n_opt=3, n_cov_fixed=1, n_pts=40, seed=2)])[0]
    bad = dict(pr); bad['e_kf'] = pr['e_kf'].copy(); bad['e_kf'][0] = 99
    with pytest.raises(orb.OrbError) as e:
        orb.LocalInertialBA([bad])
    assert e.value.code == orb.ORB_ERR_ARG
    dup = dict(pr)
    dup['e_pt'] = np.concatenate([pr['e_pt'], pr['e_pt'][:1]]); dup['e_kf'] = np.concatenate([pr['e_kf'], pr['e_kf'][:1]])
    dup['obs'] = np.concatenate([pr['obs'], pr['obs'][:1]]); dup['inv_sigma2'] = np.concatenate([pr['inv_sigma2'], pr['inv_sigma2'][:1]])
    with pytest.raises(orb.OrbError) as e:
        orb.LocalInertialBA([dup])
    assert e.value.code == orb.ORB_ERR_ARG
