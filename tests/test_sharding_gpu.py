"""-m gpu (needs >= 2 GPUs, otherwise skipped): sharded run == 1-GPU run, bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('transport', ['p2p', 'nccl'])
@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_two_slabs_reproduce_one_gpu(policy, transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, SHARD_POLICY=policy, SHARD_N='20000', SHARD_STEPS='30', SHARD_TRANSPORT=transport)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', '29617', os.path.join(ROOT, 'tests', 'sharding_check.py')],
                       env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
