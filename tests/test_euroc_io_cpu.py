"""CPU: the host mirror of the reference's EuRoC readers and IMU front end (include/orb_b200/euroc_io.hpp) against the REFERENCE's own text:
LoadImages / LoadIMU (Examples/Monocular-Inertial/mono_inertial_euroc.cc:252-310, compiled verbatim into oracle/_ref/ref_euroc) on synthetic sequence files with comment
lines, blank lines and a missing final newline; Tracking::PreintegrateIMU's queue selection and integration steps (src/Tracking.cc:1646-1729) + IMU::Preintegrated
(oracle/_ref/libref_preint.so) against SelectImuFromQueue + FlattenForPreintegration feeding the oracle's preintegration; and a numpy restatement of both."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
from orb_slam3_modified_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', 'libeuroc_mirror.so')


def _mirror():
    src = os.path.join(ROOT, 'tests', 'euroc_wrap.cpp'); hdr = os.path.join(ROOT, 'include', 'orb_b200', 'euroc_io.hpp')
    if not os.path.exists(SO) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(SO):
        subprocess.check_call(['g++', '-O2', '-std=c++14', '-Wall', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'include'), '-o', SO, src])
    return C.CDLL(SO)


def _write_sequence(d, n_frames=25, rate=200.0, fps=20.0, t0=1403636579.763555584, final_newline=True):
    stamps_ns = [int(round((t0 + k / fps) * 1e9)) for k in range(n_frames)]
    lines = []
    for k, s in enumerate(stamps_ns):
        lines.append(str(s))
        if k % 7 == 3:
            lines.append('')                                  # blank lines are skipped by both readers
    (d / 'times.txt').write_text('\n'.join(lines) + ('\n' if final_newline else ''))
    rows = ['#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]']
    n_imu = int((n_frames / fps + 0.2) * rate)
    for i in range(n_imu):
        t = t0 - 0.05 + i / rate
        Rw, p, v, om, a = synth.imu_trajectory(t - t0)
        f = Rw.T @ (a - np.array([0, 0, -9.81]))
        rows.append('%d,%.12f,%.12f,%.12f,%.12f,%.12f,%.12f' % (int(round(t * 1e9)), om[0], om[1], om[2], f[0], f[1], f[2]))
        if i == 5:
            rows.append('')
    (d / 'imu.csv').write_text('\n'.join(rows) + ('\n' if final_newline else ''))
    return stamps_ns, n_imu


def _load(lib, prefix, d, cap=4096, stride=256):
    names = C.create_string_buffer(cap * stride); ts = np.zeros(cap)
    f = getattr(lib, prefix + '_euroc_load_images'); f.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = f(str(d / 'mav0/cam0/data').encode(), str(d / 'times.txt').encode(), cap, stride, names, ts.ctypes.data)
    imgs = [names.raw[i * stride:(i + 1) * stride].split(b'\0', 1)[0].decode() for i in range(n)]
    ti = np.zeros(cap); acc = np.zeros((cap, 3), np.float32); gyr = np.zeros((cap, 3), np.float32)
    g = getattr(lib, prefix + '_euroc_load_imu'); g.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    m = g(str(d / 'imu.csv').encode(), cap, ti.ctypes.data, acc.ctypes.data, gyr.ctypes.data)
    return imgs, ts[:n].copy(), ti[:m].copy(), acc[:m].copy(), gyr[:m].copy()


@pytest.mark.parametrize('final_newline', [True, False])
def test_loaders_equal_the_reference(tmp_path, final_newline):
    stamps_ns, n_imu = _write_sequence(tmp_path, final_newline=final_newline)
    got = _load(_mirror(), 'mirror', tmp_path)
    assert len(got[0]) == len(stamps_ns) and len(got[2]) == n_imu
    assert got[0][0] == str(tmp_path / 'mav0/cam0/data') + '/' + str(stamps_ns[0]) + '.png'
    assert np.array_equal(got[1], np.array(stamps_ns, np.float64) / 1e9)
    exe = os.path.join(ROOT, 'oracle', '_ref', 'ref_euroc')
    if not (R.available() and os.path.exists(exe)):
        pytest.skip('oracle/_ref is not built here')
    # the reference's readers run as their own process (oracle/_ref/ref_euroc prints every value as an exact hex float)
    out = subprocess.check_output([exe, str(tmp_path / 'mav0/cam0/data'), str(tmp_path / 'times.txt'), str(tmp_path / 'imu.csv')]).decode().split('\n')
    n = int(out[0].split()[1])
    names = [l.rsplit(' ', 1)[0] for l in out[1:1 + n]]
    tcam = np.array([float.fromhex(l.rsplit(' ', 1)[1]) for l in out[1:1 + n]])
    m = int(out[1 + n].split()[1])
    rows = np.array([[float.fromhex(x) for x in l.split()] for l in out[2 + n:2 + n + m]])
    assert got[0] == names and got[1].tobytes() == tcam.tobytes()
    assert got[2].tobytes() == rows[:, 0].tobytes() and np.array_equal(got[3], rows[:, 1:4].astype(np.float32)) and np.array_equal(got[4], rows[:, 4:7].astype(np.float32))


def test_imu_hand_over_queue_selection_and_integration_steps(tmp_path):
    """The example's per-frame hand-over (every sample with t <= tframe, mono_inertial_euroc.cc:170-183), Tracking::PreintegrateIMU's queue selection (src/Tracking.cc:1646-1672)
    and its mid-point / interpolated integration steps (:1688-1723) against a numpy restatement."""
    stamps_ns, n_imu = _write_sequence(tmp_path, n_frames=12)
    lib = _mirror()
    imgs, tc, ti, acc, gyr = _load(lib, 'mirror', tmp_path)
    nF, maxM = len(tc), 64
    A = np.zeros((nF, maxM, 3), np.float32); G = np.zeros((nF, maxM, 3), np.float32); DT = np.zeros((nF, maxM), np.float32); NM = np.zeros(nF, np.int32)
    # hand-over: ImuSince through the flatten entry of the wrapper gives the spans; the queue holds everything handed over so far
    b = np.zeros(nF, np.int32); e = np.zeros(nF, np.int32)
    A0 = np.zeros((nF - 1, maxM, 3), np.float32); G0 = np.zeros_like(A0); D0 = np.zeros((nF - 1, maxM), np.float32); N0 = np.zeros(nF - 1, np.int32)
    lib.mirror_euroc_flatten.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 1 + [C.c_int] + [C.c_void_p] * 6
    assert lib.mirror_euroc_flatten(str(tmp_path / 'imu.csv').encode(), nF, tc.ctypes.data, maxM, A0.ctypes.data, G0.ctypes.data, D0.ctypes.data, N0.ctypes.data, b.ctypes.data, e.ctypes.data)
    first = 0
    queued = np.zeros(nF, np.int32)
    for f in range(nF):
        if f > 0:
            lo = first
            while first < len(ti) and ti[first] <= tc[f]:
                first += 1
            assert (b[f], e[f]) == (lo, first)            # the reference's loop: while(vTimestampsImu[first_imu] <= vTimestampsCam[ni]) push, first_imu++
        queued[f] = first
    sF = np.zeros(nF, np.int32); sC = np.zeros(nF, np.int32)
    assert lib.mirror_tracking_flatten(len(ti), _p(np.ascontiguousarray(ti)), _p(acc), _p(gyr), nF, _p(np.ascontiguousarray(tc)), _p(queued), maxM, _p(sF), _p(sC), _p(A), _p(G), _p(DT), _p(NM))
    front, per = 0, 0.001
    for f in range(1, nF):
        sel = []
        while front < queued[f]:                              # Tracking.cc:1650-1672
            if ti[front] < tc[f - 1] - per:
                front += 1
            elif ti[front] < tc[f] - per:
                sel.append(front); front += 1
            else:
                sel.append(front)
                break
        assert sC[f] == len(sel) and (not sel or sF[f] == sel[0]) and sel == list(range(sel[0], sel[0] + len(sel)))
        n = len(sel) - 1                                      # const int n = mvImuFromLastFrame.size() - 1
        assert NM[f] == max(n, 0)
        for i in range(n):
            k = sel[i]
            a0, a1, w0, w1 = acc[k], acc[k + 1], gyr[k], gyr[k + 1]
            if i == 0 and i < n - 1:
                tab = np.float32(ti[k + 1] - ti[k]); tini = np.float32(ti[k] - tc[f - 1])
                wa = (a0 + a1 - (a1 - a0) * (tini / tab)) * np.float32(0.5); ww = (w0 + w1 - (w1 - w0) * (tini / tab)) * np.float32(0.5); ts = np.float32(ti[k + 1] - tc[f - 1])
            elif i < n - 1:
                wa = (a0 + a1) * np.float32(0.5); ww = (w0 + w1) * np.float32(0.5); ts = np.float32(ti[k + 1] - ti[k])
            elif i > 0:
                tab = np.float32(ti[k + 1] - ti[k]); tend = np.float32(ti[k + 1] - tc[f])
                wa = (a0 + a1 - (a1 - a0) * (tend / tab)) * np.float32(0.5); ww = (w0 + w1 - (w1 - w0) * (tend / tab)) * np.float32(0.5); ts = np.float32(tc[f] - ti[k])
            else:
                wa, ww, ts = a0, w0, np.float32(tc[f] - tc[f - 1])
            assert np.array_equal(A[f, i], wa.astype(np.float32)) and np.array_equal(G[f, i], ww.astype(np.float32)) and DT[f, i] == ts
    # preintegrating the steps of one frame pair reproduces the trajectory's relative rotation over the integrated span
    f = 6
    n = NM[f]
    P = O.imu_preintegrate(A[f, :n], G[f, :n], DT[f, :n], np.zeros(6, np.float32), synth.IMU_NOISE)
    t0 = stamps_ns[0] / 1e9
    ta, tb = tc[f - 1] - t0, tc[f - 1] - t0 + float(DT[f, :n].sum())
    Ra, Rb = synth.imu_trajectory(ta)[0], synth.imu_trajectory(tb)[0]
    assert np.abs(P[1:10].reshape(3, 3) - Ra.T @ Rb).max() < 2e-4 and abs((tb - ta) - (tc[f] - tc[f - 1])) < 1e-5      # the steps span exactly the frame interval


def test_preintegration_front_end_equals_the_reference_text(tmp_path):
    """Tracking::PreintegrateIMU end to end: the queue selection (src/Tracking.cc:1646-1678), the integration steps (:1680-1729) and Preintegrated::IntegrateNewMeasurement, all
    the reference's own text (oracle/_ref/libref_preint.so), against the mirror's SelectImuFromQueue + FlattenForPreintegration feeding the oracle's preintegration: the same
    samples are selected for every frame and the per-frame preintegration records agree to float rounding -- for aligned and unaligned camera / IMU clocks, and when the IMU
    data of a frame arrives late (the queue runs empty)."""
    so = os.path.join(ROOT, 'oracle', '_ref', 'libref_preint.so')
    if not os.path.exists(so):
        pytest.skip('oracle/_ref is not built here')
    O.lib()
    Lr = C.CDLL(so); Lm = _mirror()
    rng = np.random.default_rng(3)
    for case in range(4):
        rate, fps, nF = (200.0, 20.0, 14) if case % 2 == 0 else (317.0, 29.0, 16)
        t0 = 50.0
        tI = t0 - 0.03 + np.arange(int((nF / fps + 0.1) * rate)) / rate + (0.0 if case < 2 else 0.0007)
        acc = np.zeros((len(tI), 3), np.float32); gyr = np.zeros((len(tI), 3), np.float32)
        for i, t in enumerate(tI):
            Rw, p, v, om, a = synth.imu_trajectory(t - t0)
            acc[i] = Rw.T @ (a - np.array([0, 0, -9.81])) + rng.normal(0, 0.05, 3); gyr[i] = om + rng.normal(0, 0.002, 3)
        tF = t0 + np.arange(nF) / fps
        queued = np.array([int(np.searchsorted(tI, t, side='right')) for t in tF], np.int32)      # the example's hand-over: every sample with t <= tframe
        if case == 3:
            queued[5] = queued[4]                            # frame 5 arrives before its IMU data: the queue runs empty, the next frame sees a longer run
        bias = np.array([0.02, -0.01, 0.03, 0.002, -0.001, 0.0015], np.float32); noise = np.array(synth.IMU_NOISE, np.float32)
        sF = np.zeros(nF, np.int32); sC = np.zeros(nF, np.int32); Pref = np.zeros((nF, 292), np.float32)
        Lr.ref_tracking_preintegrate(len(tI), _p(tI), _p(acc), _p(gyr), nF, _p(tF), _p(queued), _p(bias), _p(noise), _p(sF), _p(sC), _p(Pref))
        maxM = 96
        mF = np.zeros(nF, np.int32); mC = np.zeros(nF, np.int32); A = np.zeros((nF, maxM, 3), np.float32); G = np.zeros((nF, maxM, 3), np.float32); DT = np.zeros((nF, maxM), np.float32)
        NM = np.zeros(nF, np.int32)
        assert Lm.mirror_tracking_flatten(len(tI), _p(tI), _p(acc), _p(gyr), nF, _p(tF), _p(queued), maxM, _p(mF), _p(mC), _p(A), _p(G), _p(DT), _p(NM))
        assert np.array_equal(sF, mF) and np.array_equal(sC, mC), (case, sF, mF, sC, mC)
        assert sC[1:].min() >= (0 if case == 3 else 2)
        for f in range(1, nF):
            n = NM[f]
            if sC[f] < 2:
                assert n == 0 and not Pref[f].any()
                continue
            P = O.imu_preintegrate(A[f, :n], G[f, :n], DT[f, :n], bias, synth.IMU_NOISE)
            for lo, hi in ((0, 1), (1, 10), (10, 13), (13, 16), (16, 25), (25, 34), (34, 43), (43, 52), (52, 61), (61, 67), (67, 292)):
                g, w = P[lo:hi].astype(np.float64), Pref[f, lo:hi].astype(np.float64)
                assert np.abs(g - w).max() <= 3e-6 * max(np.abs(w).max(), 1e-30), (case, f, lo, hi, np.abs(g - w).max())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)
