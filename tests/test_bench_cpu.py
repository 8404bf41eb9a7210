"""CPU checks of bench.py's host-side pieces: the clock sampler's bookkeeping, and the reference arm's JSON contract
(`--impl reference`: the oracle on the host cores, one worker process per core)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_clock_sampler_reports_only_the_marked_region():
    import bench

    class FakeProc:
        stdout = iter(['1965, 1965, Not Active, Not Active, Not Active, Not Active\n'] * 4)

        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass

    cs = bench.ClockSampler.__new__(bench.ClockSampler)
    cs.rows, cs.m0, cs.proc = [], 0, FakeProc()
    cs._read()
    assert cs.samples() == 4
    cs.mark()
    assert cs.samples() == 0
    cs.rows.append(['1830', '1965', 'Not Active', 'Not Active', 'Not Active', 'Active'])
    out = cs.stop()
    assert out == {'sm_mhz': 1830, 'sm_max_mhz': 1965, 'reasons': ['sw_power_cap'], 'samples': 1}


def test_reference_arm_json_contract():
    two = sorted(os.sched_getaffinity(0))[:2]
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                                  cwd=ROOT, preexec_fn=lambda: os.sched_setaffinity(0, two), timeout=600).decode().strip().splitlines()
    d = json.loads(out[-1])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True and d['value'] > 0
    have_ref = os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'libref_orb.so'))
    cb = d['cpu_baseline']
    assert cb['kind'] == ('reference' if have_ref else 'port') and cb['kind_per_stage']['lba'] == 'port'
    assert cb['cores'] == len(two) and cb['value'] == d['value'] and 0.4 <= cb["effective_cores_measured"] <= 2.6       # a timing ratio on a shared box: sanity bounds only
    assert cb['split']['extract_ms_per_frame'] > 1 and cb['split']['lba_ms_per_problem'] > 10
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['metric'].startswith('frames/sec') and d['config']['workload'].startswith('configs[1]')


def test_host_cores_respects_affinity():
    import bench
    n, how = bench.host_cores()
    assert 1 <= n <= len(os.sched_getaffinity(0)) and isinstance(how, str)


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'], cwd=ROOT, env=env, timeout=120)
    assert out.strip() == b''
