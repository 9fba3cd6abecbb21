"""CPU: the oracle's Dubins v_pref provider (oracle/sca_dubins.c) against (a) the one known answer in the
reference (dubinsmaneuver3d.py:230-234, length 976.79) and (b) the V_des the reference's stateful tracker
(scaPolicy.py:264-338) produced at every step of the recorded SCA / RVO3D+Dubins rollouts."""
import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

DUBINS_FIXTURES = [n for n in gu.golden_names() if n.endswith('_sca') or n.endswith('_rvo3d_dubins')]


def test_known_answer_976_79():
    S = orc.dubins3d([-80, 10, 250, np.deg2rad(20), 0], [50, 70, 0, np.deg2rad(240), 0], 40.0, [np.deg2rad(-15), np.deg2rad(20)])
    assert round(S.length, 2) == 976.79
    assert abs(S.length - 976.7927058) < 1e-6
    assert S.n_points == 1001
    assert (list(S.h.mode), list(S.v.mode)) == ([2, 0, 2], [2, 1, 0])      # "RLR" + "RSL"
    p0, pe = orc.dubins_point(S, 0), orc.dubins_point(S, S.n_points - 1)
    assert np.allclose(p0, [-80, 10, 250]) and np.allclose(pe, [50, 70, 0], atol=1e-6)


def replay_provider(g, provider_step):
    """Walk a fixture teacher-forced and call provider_step(t, i, state...) for every active agent; yields the
    mismatch count against the recorded V_des."""
    T, n = g['pos'].shape[:2]
    goal_heading = g['goal_pos'][:, 3:6]
    dd = g['goal_pos'][:, :3] - g['initial_pos'][:, :3]
    zaxis = (np.abs(dd[:, 0]) <= 1e-5) & (np.abs(dd[:, 1]) <= 1e-5)
    last = [(False, 0.0)] * n       # agent.neighbors[0] as left by the last computeNeighbors
    bad, tot, worst = 0, 0, 0.0
    for t in range(T):
        s = gu.state_at(g, t)
        for i in np.nonzero(g['ran'][t])[0]:
            v = provider_step(i, last[i], s['pos'][i], s['vel'][i], s['heading'][i], g['goal'][i], goal_heading[i],
                              float(g['pref_speed'][i]), bool(zaxis[i]))
            err = np.abs(v - g['vpref_in'][t][i]).max()
            worst = max(worst, err)
            bad += err > 0
            tot += 1
        for i in np.nonzero(g['ran'][t] & g['nbr_valid'][t])[0]:
            c = int(g['nbr_cnt'][t][i])
            last[i] = (c > 0, float(g['nbr_dsq'][t][i][0]) if c > 0 else 0.0)
    return bad, tot, worst


@pytest.mark.parametrize('name', DUBINS_FIXTURES)
def test_vpref_matches_reference_tracker(name):
    g = gu.load(name)
    prov = orc.DubinsProvider(g['pos'].shape[1])

    def step(i, last, pos, vel, heading, goal, goal_heading, pref, zaxis):
        prov.set_last_neighbor(i, *last)
        return prov.vpref(i, pos, vel, heading, goal, goal_heading, pref, zaxis)
    bad, tot, worst = replay_provider(g, step)
    replans = sum(prov.S[i].replans for i in range(prov.n))
    print(name, 'agent-steps', tot, 'mismatches', bad, 'worst', worst, 'replans', replans)
    assert tot > 0 and bad == 0, (name, bad, tot, worst)
