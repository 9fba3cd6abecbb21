"""TEST INFRASTRUCTURE: run bench.py's own code path on the host emulation build of the kernels (tests/hostsim), at a toy
size, to check the HARNESS logic -- argument handling, the timed loop, the per-stage / end-to-end / sharded end-to-end
legs, the JSON line -- on a machine without a GPU.  Numbers printed this way mean nothing.

    python tests/bench_hostsim_runner.py --agents 3000 --steps 3 --warmup 3 --no-cpu
    python -m torch.distributed.run --nproc-per-node 2 ... tests/bench_hostsim_runner.py --gpus 2 --agents 4000 --steps 3

bench.py is executed from its source with two textual substitutions (device strings -> 'cpu', NCCL -> gloo); CUDA events,
streams and pinning are replaced by host stand-ins."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


class _Event:
    def __init__(self, *a, **k):
        self.t = None

    def record(self, *a, **k):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


def main():
    import pytest
    import torch
    import hostsim_util
    mp = pytest.MonkeyPatch()
    hostsim_util.activate(mp)
    hostsim_util.activate_streams(mp)
    mp.setattr(torch.cuda, 'Event', _Event)
    os.environ['SCAB_NO_GRAPH'] = '1'
    os.environ['SCAB_TRANSPORT'] = 'nccl'          # separate processes without shared device memory: staged send / recv (gloo)
    src = open(os.path.join(ROOT, 'bench.py')).read()
    for old, new in (("device = 'cuda:%d' % local", "device = 'cpu'"),
                     ("dist.init_process_group('nccl', device_id=torch.device(device))", "dist.init_process_group('gloo')")):
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    g = {'__name__': '__main__', '__file__': os.path.join(ROOT, 'bench.py')}
    exec(compile(src, os.path.join(ROOT, 'bench.py'), 'exec'), g)


if __name__ == '__main__':
    main()
