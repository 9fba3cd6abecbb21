"""CPU: the REAL multi-GPU path -- sca_b200.sharding.ShardedSwarm with the device-side halo pack / unpack and migration
kernels of csrc/scab_halo.cu -- on the host emulation build of the kernels (tests/hostsim), one process per "GPU", the
exchanges over gloo (tests/sharding_check.py with SHARD_HOSTSIM=1).  Same bar as tests/test_sharding_gpu.py holds the
hardware run to: the sharded run equals the single-device run of the same kernels bit for bit, migrations included.
(CUDA streams are no-ops and the step is launched eagerly here: the graph capture and NCCL are what only the GPU test sees.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('policy,world,n,steps,extra', [('sca', 2, 1600, 9, {}), ('orca_lp', 4, 3600, 9, {}),
                                                        ('rvo3d', 2, 1400, 12, {'SHARD_WAYPOINTS': '1', 'SHARD_MIG_CAP': '512'}),
                                                        ('sca', 2, 1000, 12, {'SHARD_DUBINS': '1', 'SHARD_MIG_CAP': '512'})],
                         ids=['sca-2', 'orca_lp-4', 'rvo3d-waypoints-2', 'sca-dubins-2'])
def test_sharded_swarm_on_emulated_devices_reproduces_the_single_device_run(policy, world, n, steps, extra):
    """`extra`: per-agent state that lives outside the 128-byte migration row -- the waypoint follower (agent.path,
    policy.now_goal) and the device-side Dubins tracker -- must travel with an agent that changes slabs
    (scab_migrate_extra_pack / _unpack); without it these two cases mismatch on ~30 agents (SHARD_NO_EXTRA=1)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'hostsim'))
    import build as hostsim_build
    hostsim_build.build()          # once, before the ranks race for it
    env = dict(os.environ, SHARD_HOSTSIM='1', SHARD_POLICY=policy, SHARD_N=str(n), SHARD_STEPS=str(steps), SHARD_WATCHDOG='600',
               OMP_NUM_THREADS='1', **extra)
    port = 29700 + (os.getpid() % 200) + world
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world), '--master-addr',
                        '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'sharding_check.py')],
                       env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:], r.stderr[-1500:])
    assert r.returncode == 0
    assert '-> OK' in r.stdout and 'migration-events 0 ' not in r.stdout


@pytest.mark.parametrize('policy,shards,n,steps,extra', [('sca', 3, 1500, 9, {}), ('orca_lp', 2, 1200, 9, {})], ids=['sca-3', 'orca_lp-2'])
def test_logical_shards_on_the_emulated_device(monkeypatch, policy, shards, n, steps, extra):
    """The one-process variant (sca_b200.sharding.LogicalShards, copies instead of send/recv) that the -m gpu suite runs on a
    single real GPU (tests/test_gpu_logical_shards.py), here on the host emulation build."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import hostsim_util
    hostsim_util.activate(monkeypatch)
    hostsim_util.activate_streams(monkeypatch)
    import logical_shards_check as lc
    info = lc.run(policy, shards, n, steps, 'cpu', mig_cap=256, **extra)
    print(info)
    assert all(v == 0 for v in info['mismatches'].values()) and not info['counter_mismatches'], info
    assert info['migrations'] > 0


def test_peer_to_peer_transport_falls_back_when_buffers_cannot_be_shared():
    """transport='p2p' between processes needs CUDA IPC; where that is unavailable (here: the emulated devices are separate
    host processes) every rank must agree to fall back to the staged transport -- and the run must still equal the single-device
    run bit for bit."""
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'hostsim'))
    import build as hostsim_build
    hostsim_build.build()
    env = dict(os.environ, SHARD_HOSTSIM='1', SHARD_TRANSPORT='p2p', SHARD_POLICY='rvo3d', SHARD_N='1400', SHARD_STEPS='6',
               SHARD_WATCHDOG='300', OMP_NUM_THREADS='1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
                        '127.0.0.1', '--master-port', str(29750 + os.getpid() % 100), os.path.join(ROOT, 'tests', 'sharding_check.py')],
                       env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-800:], r.stderr[-800:])
    assert r.returncode == 0 and '-> OK' in r.stdout
    assert 'falling back to transport=nccl' in r.stderr
