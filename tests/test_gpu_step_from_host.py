"""-m gpu: Swarm.step_from_host -- the end-to-end step for a caller whose state lives in host memory (what the mamp/envs
drop-in pays per step; bench.py's e2e leg) -- gives bit for bit what the device-resident Swarm.step gives, with and without
the overlapped copy schedule, and without moving the velocity (it is action[:3], mampenv.py:104)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_step_from_host_equals_the_resident_step(request, monkeypatch, policy):
    import torch
    from sca_b200 import capi, scenarios
    from sca_b200.swarm import Swarm
    if request.config.getoption('--hostsim'):
        import hostsim_util
        hostsim_util.activate_streams(monkeypatch)
    n = 20000 if not request.config.getoption('--hostsim') else 1500
    pos, goal, _ = scenarios.uniform_box_numpy(n, 4.0e6 / 560.0 ** 3, seed=9)

    def fresh():
        sw = Swarm(n)
        sw.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:6], initial_pos=pos, goal_pos=goal)
        sw.fit_grid()
        return sw
    ref = fresh()
    for _ in range(6):
        ref.step(policy, capi.NBR_KNN)
    torch.cuda.synchronize()
    want = ref.download()
    for overlap, chunks in ((False, 1), (True, 1), (True, 4), (True, 3)):
        sw = fresh()
        host = sw.host_buffers()
        for step in range(6):
            sw.step_from_host(host, policy, overlap=overlap, upload_vel=(step == 0), chunks=chunks, min_chunk_rows=256)
            # a host-resident caller: wipe the device copies of everything the next step uploads
            for k in ('pos', 'heading', 'total_dist'):
                sw.t[k].fill_(float('nan'))
            sw.t['status'].fill_(-1)
            sw.t['step_num'].fill_(-7)
        for k in ('pos', 'heading', 'status', 'total_dist', 'step_num', 'action'):
            assert np.array_equal(host[k].numpy(), want[k]), (overlap, chunks, k)
        assert np.array_equal(Swarm.host_velocity(host).numpy(), want['vel'])
        assert sw.counters()['n_active_total'] == ref.counters()['n_active_total']
