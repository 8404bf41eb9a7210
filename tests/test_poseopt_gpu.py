"""GPU parity of Optimizer::PoseOptimization: same inlier set, same return value, pose within 1e-9 of the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from orb_slam3_modified_b200 import synth

pytestmark = pytest.mark.gpu


def test_pose_optimization_batch():
    import orb_slam3_modified_b200 as m
    frames = [synth.pose_opt_problem(n=n, seed=s, outlier_frac=o) for s, (n, o) in enumerate([(400, 0.15), (900, 0.3), (60, 0.1), (9, 0.0), (300, 0.6), (1, 0.0), (2, 0.0), (3, 0.0)])]
    frames.append(synth.pose_opt_problem(n=350, seed=77, pose_noise=(0.5, 12.0)))    # bad prior: rejected LM trials
    outs = m.PoseOptimization(frames)
    for f, out in zip(frames, outs):
        ref = O.pose_optimization(f)
        assert out['inliers'] == ref['inliers'], (out['inliers'], ref['inliers'])
        assert np.array_equal(out['outlier'], ref['outlier'])
        assert np.allclose(out['pose'], ref['pose'], atol=1e-9)
    # fewer than 3 correspondences: the reference returns 0 and leaves the frame's pose alone (src/Optimizer.cc:996-997)
    for k in (5, 6):
        assert outs[k]['inliers'] == 0 and np.array_equal(outs[k]['pose'], np.asarray(frames[k]['pose'], np.float64)) and not outs[k]['outlier'].any()
    good = outs[0]
    assert np.abs(good['pose'] - frames[0]['gt_pose']).max() < 0.05 or np.abs(good['pose'] + frames[0]['gt_pose']).max() < 0.05
