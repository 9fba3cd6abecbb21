"""-m gpu: BASELINE configs[4] at FULL size -- 10^6 UAVs on one B200 -- checked through size-independent properties.

(1) Locality (tests/locality.py): the device step of the whole swarm must equal, for every agent of a 60 m box,
    the oracle's step of the box grown by 11.5 m (a few thousand agents): ordered neighbour lists and distance keys,
    tags, plane counts, float32 actions, flags.  The box sits where the swarm is as dense as everywhere else, so
    16-slot saturation, ties and the lazy candidate tiers are all exercised at the bench's own working-set size.
(2) Whole-swarm invariants evaluated on the device: neighbour lists sorted by distance key, inside neighborDist,
    without the agent itself or duplicates; commanded speed <= maxSpeed; active-agent counter == agents that ran.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_FULL = 1000000
DENSITY = 1.0e6 / 560.0 ** 3
_cache = {}


def _scenario():
    if 'sc' not in _cache:
        from sca_b200 import scenarios
        pos6, goal6, _ = scenarios.uniform_box_torch(N_FULL, DENSITY, seed=7)
        _cache['sc'] = (pos6, goal6)
    return _cache['sc']


@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_million_agent_step_matches_oracle_on_a_sub_box(policy):
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    import gpu_util as G
    import locality as loc
    pos6, goal6 = _scenario()
    n = N_FULL
    L = (n / DENSITY) ** (1.0 / 3.0)
    sw = Swarm(n, debug=True)
    sw.set_state(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], initial_pos=pos6, goal_pos=goal6)
    sw.fit_grid(np.array([-15.0, -15.0, 5.0]), np.array([L + 15.0, L + 15.0, L + 35.0]))
    for _ in range(3):                      # leave the zero-velocity shortcut, build up real velocities
        sw.step(policy, capi.NBR_KNN)
    c0 = sw.counters()['n_active_total']
    pre = G.post_state(sw)                  # the state this step starts from (host copies)
    sw.stage_grid()
    sw.stage_neighbors(capi.NBR_KNN)
    sw.stage_policy(policy)
    out = G.policy_outputs(sw, pre['status'])
    sw.stage_status()
    post = G.post_state(sw)

    # (1) locality against the oracle
    outer_idx, inner = loc.pick_box(pre['pos'], [L / 2, L / 2, L / 2 + 20.0], half_inner=30.0)
    assert inner.sum() > 1000 and len(outer_idx) < 20000, (int(inner.sum()), len(outer_idx))
    orc.set_threads(8)
    w = loc.make_subworld(orc, pre, pos6, goal6, outer_idx)
    w.policy_step(policy, orc.NBR_KNN)
    w.env_update(fast=True)
    tot = loc.compare_inner(out, post, w, outer_idx, inner, n, policy)
    print(policy, tot)
    assert tot['agents'] > 1000 and tot['saturated'] > 100, tot
    for k in loc.EXACT:
        assert tot.get(k, 0) == 0, (policy, k, tot)

    # (2) whole-swarm invariants, on the device
    ran = torch.from_numpy(out['ran'].astype(bool)).to(sw.device)
    p = torch.from_numpy(pre['pos']).to(sw.device)
    n_live = loc.list_invariants(sw.t['nbr_cnt'][:n], sw.t['nbr_ids'][:n], sw.t['nbr_dsq'][:n], ran, p,
                                 float(sw.params.neighbor_dist))
    assert n_live > 10 * n           # ~24 agents within 10 m on average, 16 slots
    speed = sw.t['action'][:n, 3]
    # action[3] is a 5-decimal rounding; the ORCA variant rounds |v| + 1e-5 (orca3dPolicyOfficial.py:331-342)
    assert float(speed.max()) <= float(sw.params.max_speed) + 2.5e-5
    assert sw.counters()['n_active_total'] - c0 == int(out['ran'].sum())
    del sw
    torch.cuda.empty_cache()
