"""CPU test: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        src = re.sub(r'/\*.*?\*/', '', open(h).read(), flags=re.S)
        names += re.findall(r'\b((?:orb|orbx|orbm|orbv|lba|pose|imu|local)_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import orb_slam3_modified_b200 as m
    L = m.lib()
    names = _declared()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.orb_compiled_sm() == 100 and L.orb_abi_version() >= 1


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every entry point must fail loudly (ORB_ERR_CUDA), never compute on the CPU."""
    import torch
    import pytest
    import orb_slam3_modified_b200 as m
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(m.OrbError) as ei:
        m.ORBextractor(1000, 1.2, 8, 20, 7)
    assert ei.value.code == m.ORB_ERR_CUDA
    # the stateless entry points as well: LocalInertialBA on a well-formed problem
    import oracle_lib as O
    from orb_slam3_modified_b200 import synth
    pr = synth.local_inertial_ba_problem(n_opt=2, n_cov_fixed=1, n_pts=20, seed=1)
    pr['preint'] = O.liba_preints(pr)
    with pytest.raises(m.OrbError) as ei:
        m.LocalInertialBA([pr])
    assert ei.value.code == m.ORB_ERR_CUDA


def test_keypoint_layout_is_cv_keypoint():
    import orb_slam3_modified_b200 as m
    assert m.KP_DTYPE.itemsize == 28
    assert [m.KP_DTYPE.fields[k][1] for k in ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')] == [0, 4, 8, 12, 16, 20, 24]
