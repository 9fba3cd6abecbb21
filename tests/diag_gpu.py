"""Diagnostic (not a test): per-component mismatch report of the CUDA path vs the golden fixtures."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import golden_util as gu
import gpu_util as G

names = sys.argv[1:] or gu.golden_names()
for name in names:
    g = gu.load(name); pol = str(g['policy']); T = g['pos'].shape[0]
    sw = G.swarm_from_golden(g)
    comp = np.zeros(7, int); tot = {}
    worst = []
    for t in range(T):
        s = gu.state_at(g, t)
        G.force_state(sw, s)
        sw.set_external_vpref(g['vpref_in'][t] if gu.uses_external_vpref(pol) else None)
        sw.stage_grid(); sw.stage_neighbors(1, pos_host=s['pos']); sw.stage_policy(pol)
        out = G.policy_outputs(sw, s['status'])
        r = gu.compare_step(g, t, out, dsq_rtol=4e-16)
        sw.stage_status()
        post = G.post_state(sw); s2 = gu.state_at(g, t + 1)
        r['status_bad'] = int((post['status'] != (s2['status'] & 15)).sum())
        for k, v in r.items():
            tot[k] = max(tot.get(k, 0), v) if k.endswith('maxerr') else tot.get(k, 0) + v
        ran = g['ran'][t].astype(bool)
        a = out['action'].astype(np.float32); b = g['action'][t].astype(np.float32)
        comp += ((a != b) & ran[:, None]).sum(0)
        bad = np.where((np.abs(a.astype('float64') - b.astype('float64')) > 1e-9).any(1) & ran)[0]
        for i in bad[:2]:
            if len(worst) < 4:
                worst.append((t, int(i), a[i].tolist(), b[i].tolist(), s['heading'][i].tolist(), int(g['tag'][t][i])))
    print(name, tot, 'per-component', comp.tolist())
    for w in worst:
        print('   ', w)
