"""-m gpu: KNN neighbour mode at larger N against the C oracle on seeded synthetic swarms
(the configurations of BASELINE configs 3-5 at sizes the oracle finishes in seconds)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(pos6, goal6, obstacles, policy, steps, lo=None, hi=None, full_search=True, radius=0.5, pref_speed=1.0):
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    import gpu_util as G
    n = len(pos6)
    import os
    orc.set_threads(max(8, min(32, os.cpu_count() or 8)))
    w = orc.World(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], obstacles=obstacles, initial_pos=pos6, goal_pos=goal6,
                  radius=radius, pref_speed=pref_speed)
    sw = Swarm(n, debug=True)
    sw.h.set_option(capi.OPT_FULL_SEARCH, 1 if full_search else 0)
    sw.set_state(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], initial_pos=pos6, goal_pos=goal6,
                 max_run_dist=w.max_run_dist, radius=radius, pref_speed=pref_speed)
    sw.set_obstacles(obstacles)
    lo = np.minimum(pos6[:, :3].min(0), goal6[:, :3].min(0)) - 15 if lo is None else lo
    hi = np.maximum(pos6[:, :3].max(0), goal6[:, :3].max(0)) + 15 if hi is None else hi
    sw.fit_grid(lo, hi)
    tot = dict(agents=0, nbr=0, tag=0, nsuit=0, vel=0, status=0, planes=0, flips=0)
    for t in range(steps):
        pre = dict(pos=w.pos.copy(), vel=w.vel.copy(), heading=w.heading.copy(), status=w.status.copy(),
                   total_dist=w.total_dist.copy(), step_num=w.step_num.copy())
        G.force_state(sw, pre)                       # teacher-forced from the oracle's state
        w.policy_step(policy, orc.NBR_KNN)
        sw.stage_grid(); sw.stage_neighbors(capi.NBR_KNN); sw.stage_policy(policy)
        out = G.policy_outputs(sw, pre['status'])
        ran = w.ran.astype(bool)
        assert np.array_equal(out['ran'].astype(bool), ran)
        has = ran & (w.nbr_cnt >= 0)
        tot['agents'] += int(ran.sum())
        tot['nbr'] += int(((out['nbr_ids'][has] != w.nbr_ids[has]).any(1) | (out['nbr_cnt'][has] != w.nbr_cnt[has])).sum())
        tot['tag'] += int((out['tag'][ran] != w.tag[ran]).sum())
        if policy != 'orca_lp':
            tot['nsuit'] += int(((out['n_suit'][ran] != w.n_suit[ran]) & (out['n_suit'][ran] >= 0)).sum())
        else:
            tot['planes'] += int(((out['plane_fail'][ran] != w.plane_fail[ran]) | (out['lp4_proj'][ran] != w.lp4_proj[ran])).sum())
        dv = np.abs(out['action'][:, :4].astype('float64') - w.action[:, :4].astype('float64')).max(1)
        tot['vel'] += int((dv[ran] > 2e-5).sum())
        tot['flips'] += int((dv[ran] > 0).sum())
        w.env_update(fast=True)
        sw.stage_status()
        post = G.post_state(sw)
        tot['status'] += int((post['status'] != w.status).sum())
    return tot


@pytest.mark.parametrize('full_search', [True, False], ids=['full', 'lazy'])
@pytest.mark.parametrize('policy', ['sca', 'srvo3d', 'rvo3d', 'orca_sampled', 'orca_lp'])
def test_uniform_swarm_config5_density(policy, full_search):
    from sca_b200 import scenarios
    pos, goal, obs = scenarios.uniform_box_numpy(6000, 3.0 * 1.0e6 / 560.0 ** 3, seed=11)
    rng = np.random.default_rng(5)
    L = pos[:, :3].max()
    obs = np.c_[rng.uniform(0, L, (40, 3)) + [0, 0, 20], rng.uniform(0.5, 2.0, 40)]
    tot = _compare(pos, goal, obs, policy, steps=5, full_search=full_search)
    print(policy, tot)
    assert tot['agents'] > 20000
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status', 'planes'):
        assert tot[k] == 0, (policy, tot)


@pytest.mark.parametrize('full_search', [True, False], ids=['full', 'lazy'])
@pytest.mark.parametrize('policy', ['sca', 'rvo3d', 'orca_lp'])
def test_mixed_radii_and_speeds_without_obstacles(policy, full_search):
    """Radii from 0.1 to 2.5 m and preferred speeds from 0.5 to 2 m/s in one swarm, no obstacles -- so the 16-NN runs on its
    scale path (reduced query ball, crowded pass) and the status stage on its list shortcut, whose bounds all depend on the
    largest radius and the fastest agent; colliders of large radius displace nearer small agents from the lists."""
    from sca_b200 import scenarios
    pos, goal, _ = scenarios.uniform_box_numpy(4000, 3.0 * 1.0e6 / 560.0 ** 3, seed=17)
    rng = np.random.default_rng(9)
    radius = np.exp(rng.uniform(np.log(0.1), np.log(2.5), len(pos)))
    pref = rng.uniform(0.5, 2.0, len(pos))
    tot = _compare(pos, goal, None, policy, steps=8, full_search=full_search, radius=radius, pref_speed=pref)
    print(policy, tot)
    assert tot['agents'] > 15000
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status', 'planes'):
        assert tot[k] == 0, (policy, tot)


def test_takeoff_landing_tiled_config3():
    from sca_b200 import scenarios
    pos, goal, obs = scenarios.takeoff_landing_tiled(pads=64)       # 1024 UAVs, 512 obstacles
    tot = _compare(pos, goal, obs, 'sca', steps=8)
    print(tot)
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status'):
        assert tot[k] == 0, tot


@pytest.mark.parametrize('policy', ['rvo3d', 'srvo3d'])
def test_dense_sphere_config4(policy):
    from sca_b200 import scenarios
    pos, goal, obs = scenarios.dense_sphere(4000)
    tot = _compare(pos, goal, obs, policy, steps=5)
    print(policy, tot)
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status'):
        assert tot[k] == 0, (policy, tot)


# ---------------------------------------------------------------- BASELINE configs 3 and 4 at their NAMED sizes
def test_takeoff_landing_config3_at_10000_uavs():
    """BASELINE configs[2]: run_sca.py:53-81 take-off / landing pads with the 8 sphere obstacles of :139-150, tiled 25 x 25
    (pitch 30 m) = 10 000 UAVs + 5 000 obstacles, SCA with the posture constraint (vertical start->goal: 2 x 128 + 1
    candidates, scaPolicy.py:188-190).  EVERY agent of the whole world against the oracle, teacher-forced, 6 steps."""
    from sca_b200 import scenarios
    pos, goal, obs = scenarios.takeoff_landing_tiled(pads=625)
    assert len(pos) == 10000 and len(obs) == 5000
    tot = _compare(pos, goal, obs, 'sca', steps=6)
    print(tot)
    assert tot['agents'] >= 59000
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status'):
        assert tot[k] == 0, tot


@pytest.mark.parametrize('policy', ['rvo3d', 'srvo3d'])
def test_dense_sphere_config4_at_100000_agents(policy):
    """BASELINE configs[3]: run_orca.py:36-54 Fibonacci sphere swap at 100 000 agents (R = 198 m, ~64 agents within
    neighborDist of each: every 16-slot list saturates).  The whole world against the oracle, 3 steps."""
    from sca_b200 import scenarios
    pos, goal, obs = scenarios.dense_sphere(100000)
    tot = _compare(pos, goal, obs, policy, steps=3, full_search=False)
    print(policy, tot)
    assert tot['agents'] == 300000
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status'):
        assert tot[k] == 0, (policy, tot)
