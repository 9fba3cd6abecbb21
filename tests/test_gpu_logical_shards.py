"""-m gpu, ONE GPU: the sharded path (k_halo_*, k_migrate_*, the ghost rows, the global-id tie-breaks) as S logical
shards in one process -- bit for bit equal to the single-device run, migrations included."""
import pytest

pytestmark = pytest.mark.gpu

CASES = [('sca', 2, 20000, 30, {}), ('sca', 4, 30000, 24, {}), ('orca_lp', 2, 20000, 30, {}), ('orca_lp', 4, 30000, 24, {}),
         ('srvo3d', 3, 12000, 16, {}),
         ('rvo3d', 2, 6000, 16, {'waypoints': True, 'mig_cap': 1024}), ('sca', 2, 4000, 16, {'dubins': True, 'mig_cap': 1024}),
         ('sca', 4, 6000, 12, {'dubins': True, 'mig_cap': 512}),
         # the staged path (pack -> copy -> unpack: what the NCCL transport does between processes)
         ('sca', 2, 12000, 16, {'transport': 'nccl'}), ('rvo3d', 2, 6000, 16, {'waypoints': True, 'mig_cap': 1024, 'transport': 'nccl'})]
IDS = ['sca-2', 'sca-4', 'orca_lp-2', 'orca_lp-4', 'srvo3d-3', 'rvo3d-waypoints-2', 'sca-dubins-2', 'sca-dubins-4', 'sca-2-staged',
       'rvo3d-waypoints-2-staged']


@pytest.mark.parametrize('policy,shards,n,steps,extra', CASES, ids=IDS)
def test_logical_shards_reproduce_the_single_device_run(request, monkeypatch, policy, shards, n, steps, extra):
    device = 'cuda:0'
    if request.config.getoption('--hostsim'):
        import hostsim_util
        hostsim_util.activate_streams(monkeypatch)
        device, n, steps = 'cpu', max(700 * shards, n // 16), min(steps, 10)      # (every slab wider than the halo)
        extra = dict(extra, mig_cap=min(extra.get('mig_cap', 4096), 256))
    import logical_shards_check as lc
    info = lc.run(policy, shards, n, steps, device, **extra)
    print(info)
    assert all(v == 0 for v in info['mismatches'].values()), info
    assert not info['counter_mismatches'], info
    assert info['migrations'] > 0, 'the case never migrated an agent: it does not test the migration kernels'
    if extra.get('waypoints'):
        assert info['popped'] > 0
    if extra.get('dubins'):
        assert info['replans'] >= n
