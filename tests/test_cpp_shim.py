"""The C++ host mirror (include/orb_b200/orb_slam3.hpp) compiles with plain g++ against the C-ABI (CPU test) and
produces the oracle's result when run (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, 'tests', 'cpp_example')


def _build():
    lib = os.path.join(ROOT, 'orb_slam3_modified_b200')
    subprocess.check_call(['g++', '-std=c++14', '-O2', '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'cpp_example.cpp'),
                           '-o', EXE, '-L', lib, '-l:liborb_b200.so', '-Wl,-rpath,' + lib])


def test_cpp_shim_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_shim_runs_like_the_oracle(tmp_path):
    import oracle_lib as O
    from orb_slam3_modified_b200 import synth
    _build()
    img = synth.frame(2)
    p = tmp_path / 'img.raw'
    p.write_bytes(img.tobytes())
    out = subprocess.check_output([EXE, str(p), '480', '640']).decode().split()
    mono, kps, desc = O.OracleExtractor()(img, (0, 1000))
    s = 0
    for b in desc.tobytes():
        s = (s * 1315423911 + b) % (1 << 64)
    assert (int(out[0]), int(out[1]), int(out[2])) == (mono, len(kps), s)
