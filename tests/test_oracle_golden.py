"""CPU: pins the C oracle (oracle/sca_oracle.c) against the golden fixtures that the UNMODIFIED Python
reference produced (tests/golden/make_golden.py).  Teacher-forced: every fixture step's pre-state is
loaded, one oracle step is run, and every recorded quantity is compared."""
import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc


def run_fixture(name, g=None):
    g = gu.load(name) if g is None else g
    pol = str(g['policy'])
    T = g['pos'].shape[0]
    w = orc.World(g['pos'][0], g['goal'], heading=g['heading'][0], radius=g['radius'], pref_speed=g['pref_speed'],
                  obstacles=g['obstacles'], initial_pos=g['initial_pos'], goal_pos=g['goal_pos'])
    assert np.array_equal(w.max_run_dist, g['max_run_dist'])
    if 'waypoints' in g:          # agent.path of an external planner; the pop state evolves in the oracle
        w.set_waypoints(g['waypoints'], g['waypoint_count'])
    tot = {}
    for t in range(T):
        s = gu.state_at(g, t)
        if 'waypoints' in g:
            ng = g['now_goal'][t]
            same = np.isnan(ng) & np.isnan(w.now_goal) | (ng == w.now_goal)
            tot['path_bad'] = tot.get('path_bad', 0) + int((w.waypoint_left != g['path_left'][t]).sum()) + int((~same).sum())
        w.pos[:] = s['pos']
        w.vel[:] = s['vel'].astype(np.float32)
        w.heading[:] = s['heading']
        w.status[:] = s['status']
        w.total_dist[:] = s['total_dist']
        w.step_num[:] = s['step_num']
        w.set_vpref_ext(g['vpref_in'][t] if gu.uses_external_vpref(pol) else None)
        w.policy_step(pol, orc.NBR_REF_COMPAT)
        out = {k: getattr(w, k).copy() for k in ('ran', 'tag', 'coll_query', 'nbr_cnt', 'nbr_ids', 'nbr_dsq', 'n_suit',
                                                 'plane_fail', 'n_planes', 'lp4_proj', 'action')}
        r = gu.compare_step(g, t, out, who='%s step %d' % (name, t))
        r['perm_bad'] = int(not np.array_equal(w.agent_perm, g['agent_ids_perm'][t]))
        ran = g['ran'][t].astype(bool)
        if not gu.uses_external_vpref(pol):
            r['vpref_bad'] = int((w.vpref_in != g['vpref_in'][t])[ran].sum())
            r['vpref_store_bad'] = int((w.vpref_store != g['vpref_store'][t])[ran].sum())
        if pol == 'orca_lp':
            r['lp4_inner_fail'] = int((w.lp4_inner_fail != g['lp4_inner_fail'][t])[ran].sum())
        done = w.env_update()
        s2 = gu.state_at(g, t + 1)
        r['done_bad'] = int(done != bool(g['done'][t]))
        r['status_bad'] = int((w.status != s2['status']).sum())
        r['stepnum_bad'] = int((w.step_num != s2['step_num']).sum())
        r['vel_state_bad'] = int((w.vel != s2['vel'].astype(np.float32)).sum())
        # np.cos / np.sin (NumPy SIMD kernels) vs libm differ by an ulp now and then: 1e-12 on the kinematics
        r['pos_bad'] = int((np.abs(w.pos - s2['pos']) > 1e-12).sum())
        r['head_bad'] = int((np.abs(w.heading - s2['heading']) > 1e-12).sum())
        r['dist_bad'] = int((np.abs(w.total_dist - s2['total_dist']) > 1e-12).sum())
        for k, v in r.items():
            tot[k] = max(tot.get(k, 0), v) if k.endswith('maxerr') else tot.get(k, 0) + v
    return tot


@pytest.mark.parametrize('name', gu.golden_names())
def test_oracle_matches_reference(name):
    tot = run_fixture(name)
    assert tot['agents'] > 0
    for k, v in tot.items():
        if k in ('agents', 'vel_maxerr', 'n_suit_known', 'angle_ulp_flips'):
            continue
        assert v == 0, (name, k, tot)


def test_fixture_coverage():
    """The fixtures exercise every branch the scope table names."""
    tags = {}
    for name in gu.golden_names():
        g = gu.load(name)
        tags.setdefault(str(g['policy']), set()).update(np.unique(g['tag'][g['ran'].astype(bool)]).tolist())
    for pol in ('sca', 'srvo3d', 'rvo3d', 'rvo3d_dubins', 'orca_sampled'):
        assert gu.TAG_FIRST in tags[pol] and gu.TAG_SUITABLE in tags[pol], (pol, tags[pol])
    for pol in ('srvo3d', 'rvo3d', 'orca_sampled'):
        assert gu.TAG_FALLBACK in tags[pol], (pol, tags[pol])
    assert {gu.TAG_FIRST, gu.TAG_LP3_OK, gu.TAG_LP4} <= tags['orca_lp']
    g = gu.load('random1000_orca_lp')
    assert g['coll_query'].sum() > 0 and g['lp4_proj'].sum() > 0 and g['lp4_inner_fail'].sum() > 0
    assert (gu.load('takeoff16_sca')['nbr_ids'] >= 16).any()      # obstacle neighbours
    assert (gu.load('random1000_orca_lp')['nbr_cnt'] == 16).any()  # saturated lists (the eviction rule)
