"""CPU: bench.py's own code path (timed loop, per-stage leg, overlapped end-to-end leg, sharded end-to-end leg, JSON
contract) executed at toy size on the host emulation build of the kernels (tests/bench_hostsim_runner.py).  The numbers
are meaningless; what is checked is that the harness runs and emits every key the contract asks for, at 1 and 2 devices."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, 'tests', 'bench_hostsim_runner.py')
KEYS = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
        'dtype', 'data', 'config', 'gpu_launches', 'clocks', 'e2e')


def _line(out):
    rows = [l for l in out.splitlines() if l.startswith('{"metric"')]
    assert len(rows) == 1, out[-2000:]
    return json.loads(rows[0])


def test_bench_single_device_on_the_emulator():
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'hostsim'))
    import build as hostsim_build
    hostsim_build.build()
    env = dict(os.environ, OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, RUNNER, '--agents', '2500', '--steps', '3', '--warmup', '3', '--no-cpu'], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    for k in KEYS + ('roofline',):
        assert k in d, k
    assert d['n_gpus'] == 1 and d['value'] > 0 and d['config']['workload'].startswith('BASELINE configs[4]')
    assert d['e2e']['value'] > 0 and d['e2e']['h2d_bytes_per_step'] == 2500 * 64 and d['e2e']['d2h_bytes_per_step'] == 2500 * 92   # (the velocity stays on the device)
    assert set(d['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
    assert d['e2e']['api'] == 'Swarm.step_from_host'      # (overlapped == serial, bit for bit: tests/test_gpu_step_from_host.py)


def test_bench_two_devices_on_the_emulator():
    env = dict(os.environ, OMP_NUM_THREADS='1')
    port = 29750 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), RUNNER, '--gpus', '2', '--agents', '2000', '--steps', '3', '--warmup', '3'],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = _line(r.stdout)
    for k in KEYS:
        assert k in d, k
    assert d['n_gpus'] == 2 and d['value'] > 0 and d['scaling'] == 'strong'
    assert d['e2e']['value'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > d['e2e']['h2d_bytes_per_step']
    assert 'e2e omitted' not in r.stderr
