"""GPU-side helpers for the parity tests: drive sca_b200.Swarm from a golden fixture / oracle world."""
import numpy as np

import golden_util as gu


def swarm_from_golden(g, debug=True, full_search=True):
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    n = g['pos'].shape[1]
    sw = Swarm(n, debug=debug)
    sw.h.set_option(capi.OPT_FULL_SEARCH, 1 if full_search else 0)
    s0 = gu.state_at(g, 0)
    sw.set_state(s0['pos'], g['goal'], heading=s0['heading'], vel=s0['vel'].astype(np.float32), radius=g['radius'],
                 pref_speed=g['pref_speed'], max_run_dist=g['max_run_dist'], initial_pos=g['initial_pos'],
                 goal_pos=g['goal_pos'])
    sw.set_obstacles(g['obstacles'])
    if 'waypoints' in g:       # agent.path; the pop state then evolves on the device
        sw.set_waypoints(g['waypoints'], g['waypoint_count'])
    lo = np.minimum(g['pos'].reshape(-1, 3).min(0), g['goal'].min(0)) - 5.0
    hi = np.maximum(g['pos'].reshape(-1, 3).max(0), g['goal'].max(0)) + 5.0
    sw.fit_grid(lo, hi)
    return sw


def force_state(sw, s):
    """Teacher-force a pre-step state (dict of numpy arrays) into the device arrays."""
    import torch
    host = {'pos': torch.from_numpy(np.ascontiguousarray(s['pos'], 'float64')),
            'vel': torch.from_numpy(np.ascontiguousarray(s['vel'], np.float32)),
            'heading': torch.from_numpy(np.ascontiguousarray(s['heading'], 'float64')),
            'status': torch.from_numpy(np.ascontiguousarray(s['status'], np.int32) & 15),
            'total_dist': torch.from_numpy(np.ascontiguousarray(s['total_dist'], 'float64')),
            'step_num': torch.from_numpy(np.ascontiguousarray(s['step_num'], np.int32))}
    sw.upload(host)


def policy_outputs(sw, pre_status):
    """Outputs after grid -> neighbours -> policy (before the status stage), fixture-shaped."""
    import torch
    torch.cuda.synchronize()
    out = {k: sw.t[k].cpu().numpy() for k in ('nbr_cnt', 'nbr_ids', 'nbr_dsq', 'tag', 'n_suit', 'plane_fail',
                                              'n_planes', 'lp4_proj', 'action')}
    ran = (np.asarray(pre_status) & 7) == 0
    out['ran'] = ran.astype(np.int32)
    st = sw.t['status'].cpu().numpy()
    out['coll_query'] = ((st & 2) != 0).astype(np.int32) * ran
    return out


def post_state(sw):
    import torch
    torch.cuda.synchronize()
    d = sw.download(('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num'))
    d['status'] = d['status'] & 15
    return d
