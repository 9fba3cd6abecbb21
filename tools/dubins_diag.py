"""Diagnostic (not a test): where does the on-device Dubins tracker leave the oracle's?  Teacher-forced on a fixture."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import golden_util as gu
import gpu_util as G
from oracle import oracle as orc

for name in sys.argv[1:]:
    g = gu.load(name); pol = str(g['policy']); T, n = g['pos'].shape[:2]
    sw = G.swarm_from_golden(g); sw.enable_dubins(g['goal_pos'][:, 3:6])
    prov = orc.DubinsProvider(n)
    zaxis = (sw.t['status'][:n].cpu().numpy() & 16) != 0
    shown = 0
    for t in range(T):
        s = gu.state_at(g, t)
        G.force_state(sw, s)
        sw.stage_grid(); sw.stage_neighbors(1, pos_host=s['pos']); sw.stage_policy(pol)
        ran = g['ran'][t].astype(bool)
        v = sw.t['vpref_out'][:n].cpu().numpy()
        st = sw.dubins_states()
        for i in np.nonzero(ran)[0]:
            before = prov.S[i].replans
            vo = prov.vpref(i, s['pos'][i], s['vel'][i], s['heading'][i], g['goal'][i], g['goal_pos'][i, 3:6], float(g['pref_speed'][i]), bool(zaxis[i]))
            S = prov.S[i]
            d = np.abs((np.trunc(v[i] * 1e5) + 0.0) / 1e5 - vo).max()
            if d > 1.1e-5 and shown < 12:
                shown += 1
                print(name, 't', t, 'agent', i, 'd', d, 'replanned now: oracle', S.replans - before, 'gpu total', st['replans'][i], 'oracle total', S.replans)
                print('   pos', s['pos'][i].tolist(), 'heading', s['heading'][i].tolist(), 'vel', s['vel'][i].tolist(), 'goal', g['goal_pos'][i].tolist())
                print('   gpu   v', v[i].tolist(), 'len', st['length'][i], 'npts', st['n_points'][i], 'cur', st['cursor'][i], 'h', st['h']['mode'][i].tolist(), st['h']['r'][i], st['h']['t'][i], st['h']['p'][i], st['h']['q'][i],
                      'v', st['v']['mode'][i].tolist(), st['v']['r'][i], st['v']['t'][i], st['v']['p'][i], st['v']['q'][i], 'now_goal', st['now_goal'][i].tolist(), 'has_nbr', st['has_nbr'][i], st['nbr0_dsq'][i])
                print('   oracle v', vo.tolist(), 'len', S.length, 'npts', S.n_points, 'cur', S.cursor, 'h', list(S.h.mode), S.h.r, S.h.t, S.h.p, S.h.q,
                      'v', list(S.v.mode), S.v.r, S.v.t, S.v.p, S.v.q, 'now_goal', list(S.now_goal), 'has_nbr', S.has_nbr, S.nbr0_dsq)
        for i in np.nonzero(ran & g['nbr_valid'][t].astype(bool))[0]:
            c = int(g['nbr_cnt'][t][i])
            prov.set_last_neighbor(i, c > 0, float(g['nbr_dsq'][t][i][0]) if c > 0 else 0.0)
        sw.stage_status()
