"""-m gpu: the mamp drop-in shims (GPUKDTree + GPU*Policy) driven the way MACAEnv._take_action drives
them (mampenv.py:27-40), teacher-forced from the golden fixtures."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu

CASES = [('circle16_srvo3d', 'GPUSRVO3DPolicy'), ('takeoff16_sca', 'GPUSCAPolicy'), ('dense60_orca_lp', 'GPUORCA3DPolicyOfficial'),
         ('dense40_orca_sampled', 'GPUORCA3DPolicy'), ('circle16_rvo3d', 'GPURVO3DPolicy')]


@pytest.mark.parametrize('name,cls', CASES)
def test_shim_matches_reference_actions(name, cls):
    import mini_mamp
    from sca_b200 import mamp_compat
    g = gu.load(name)
    pol_cls = getattr(mamp_compat, cls)
    n = g['pos'].shape[1]
    ext = gu.uses_external_vpref(str(g['policy']))
    T = min(g['pos'].shape[0], 40)
    agents = [mini_mamp.Agent(g['initial_pos'][i], g['goal_pos'][i], [0.0, 0.0, 0.0], float(g['radius'][i]),
                              float(g['pref_speed'][i]), pol_cls, i) for i in range(n)]
    obstacles = [mini_mamp.Obstacle(o[:3], float(o[3]), k) for k, o in enumerate(g['obstacles'])]
    step_holder = {'t': 0}
    provider = (lambda a: g['vpref_in'][step_holder['t']][a.id]) if ext else None
    tree = mamp_compat.GPUKDTree(agents, obstacles, vpref_provider=provider)
    if not ext:
        tree.vpref_provider = None
    bad = 0
    for t in range(T):
        step_holder['t'] = t
        s = gu.state_at(g, t)
        for i, a in enumerate(agents):       # teacher-force the reference's pre-step state into the objects
            a.pos_global_frame = s['pos'][i].copy()
            a.vel_global_frame = s['vel'][i].astype(np.float32)
            a.heading_global_frame = s['heading'][i].copy()
            a.is_at_goal = bool(s['status'][i] & 1)
            a.is_collision = bool(s['status'][i] & 2)
            a.is_out_of_max_time = bool(s['status'][i] & 4)
            a.total_dist = float(s['total_dist'][i])
            a.step_num = int(s['step_num'][i])
        tree.buildAgentTree()                                           # mampenv.py:28
        all_actions = np.zeros((n, 7), dtype=np.float32)               # mampenv.py:31
        for i, a in enumerate(agents):
            if a.is_at_goal or a.is_collision or a.is_out_of_max_time:  # mampenv.py:35
                continue
            all_actions[i, :] = a.policy.find_next_action({}, a, tree)  # mampenv.py:40
        ran = g['ran'][t].astype(bool)
        ref = g['action'][t].astype(np.float32)
        assert np.array_equal(all_actions[ran][:, :4], ref[ran][:, :4]), (name, t)
        assert np.abs(all_actions[ran][:, 4:] - ref[ran][:, 4:]).max(initial=0) < 5e-7
        for i in np.nonzero(ran & g['nbr_valid'][t].astype(bool))[0]:
            ids = [(o.id + n if o.is_obstacle else o.id) for o, _ in agents[i].neighbors]
            bad += ids != g['nbr_ids'][t][i][:g['nbr_cnt'][t][i]].tolist()
            assert agents[i].is_collision == bool(g['coll_query'][t][i])
    assert bad == 0


@pytest.mark.parametrize('name,cls', [('takeoff16_sca', 'GPUSCAPolicy'), ('circle16_rvo3d_dubins', 'GPURVO3dDubinsPolicy')])
def test_shim_default_provider_is_the_device_dubins_tracker(name, cls):
    """GPUSCAPolicy / GPURVO3dDubinsPolicy take V_des from the Dubins tracker ON THE DEVICE by default (the host-side reference
    tracker is 68 % of the reference's CPU time; *HostDubins keeps it).  Teacher-forced from the fixture; the tracker's own
    state free-runs on the device.  Bar = the Dubins bar of test_gpu_dubins.py: the device libm differs from glibc in the
    last ulp, which the reference's 1e-5 roundings can amplify to ~1e-4 of V_des."""
    import mini_mamp
    from sca_b200 import mamp_compat
    if name not in gu.golden_names():
        pytest.skip('fixture not present')
    g = gu.load(name)
    pol_cls = getattr(mamp_compat, cls)
    assert pol_cls.vpref_provider == 'device'
    n = g['pos'].shape[1]
    T = min(g['pos'].shape[0], 40)
    agents = [mini_mamp.Agent(g['initial_pos'][i], g['goal_pos'][i], [0.0, 0.0, 0.0], float(g['radius'][i]),
                              float(g['pref_speed'][i]), pol_cls, i) for i in range(n)]
    obstacles = [mini_mamp.Obstacle(o[:3], float(o[3]), k) for k, o in enumerate(g['obstacles'])]
    tree = mamp_compat.GPUKDTree(agents, obstacles)
    assert tree.vpref_provider is None and 'goal_heading' in tree.sw.t     # resolved to the device tracker
    tot = same = 0
    worst = 0.0
    for t in range(T):
        s = gu.state_at(g, t)
        for i, a in enumerate(agents):
            a.pos_global_frame = s['pos'][i].copy()
            a.vel_global_frame = s['vel'][i].astype(np.float32)
            a.heading_global_frame = s['heading'][i].copy()
            a.is_at_goal, a.is_collision, a.is_out_of_max_time = bool(s['status'][i] & 1), bool(s['status'][i] & 2), bool(s['status'][i] & 4)
            a.total_dist, a.step_num = float(s['total_dist'][i]), int(s['step_num'][i])
        tree.buildAgentTree()
        ran = g['ran'][t].astype(bool)
        act = np.array([a.policy.find_next_action({}, a, tree) for a in agents], np.float32)
        ref = g['action'][t].astype(np.float32)
        d = np.abs(act[ran][:, :4].astype('float64') - ref[ran][:, :4].astype('float64')).max(1)
        tot += int(ran.sum()); same += int((d == 0).sum()); worst = max(worst, float(d.max(initial=0)))
    print(name, 'agent-steps', tot, 'identical', same, 'worst', worst)
    assert same >= 0.9 * tot and worst <= 0.05
