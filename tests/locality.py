"""Locality: a size-independent property of the collision-avoidance step, used to check FULL-SIZE runs.

Every stage of the step reads only agents within neighborDist = 10 m (agent.py:33; the collision check of
mampenv.py:61-80 reaches r_i + r_j plus one step of motion).  So if S is the set of agents inside a box and S+ the
agents inside the box grown by more than 10 m, stepping the world that holds only S+ gives, for the agents of S,
exactly what stepping the whole swarm gives: same ordered neighbour lists, tags, plane counts, actions, flags.
The oracle finishes a few thousand agents in milliseconds, so a 10^6-agent device step can be checked against it
on a sub-box (tests/test_gpu_full_size.py); tests/test_oracle_properties.py checks the property on the oracle itself.
"""
import numpy as np

STATE_KEYS = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')


def pick_box(pos, centre, half_inner, margin=11.5):
    """-> (outer_idx ascending, inner mask over outer_idx).  Chebyshev boxes around `centre`."""
    d = np.abs(np.asarray(pos)[:, :3] - np.asarray(centre)[None, :]).max(1)
    outer_idx = np.nonzero(d <= half_inner + margin)[0]
    return outer_idx, d[outer_idx] <= half_inner


def make_subworld(orc, pre, pos0_6, goal6, outer_idx):
    """Oracle world holding only the rows `outer_idx`, teacher-forced to the pre-step state `pre`."""
    w = orc.World(pre['pos'][outer_idx], goal6[outer_idx, :3], heading=pre['heading'][outer_idx],
                  initial_pos=pos0_6[outer_idx], goal_pos=goal6[outer_idx])
    w.vel[:] = np.asarray(pre['vel'], np.float32)[outer_idx]
    w.status[:] = np.asarray(pre['status'], np.int32)[outer_idx] & 15
    w.total_dist[:] = pre['total_dist'][outer_idx]
    w.step_num[:] = pre['step_num'][outer_idx]
    return w


def compare_inner(out, post, w, outer_idx, inner, n_full, policy):
    """`out` / `post`: policy outputs and post-step state of the FULL run (rows = global ids); `w`: the stepped
    sub-world.  Counts mismatches over the inner agents."""
    lut = np.full(n_full + 1, -1, np.int64)          # global id -> sub-world index (ascending, so id ties keep their order)
    lut[outer_idx] = np.arange(len(outer_idx))
    gi = outer_idx[inner]                            # global rows of the inner agents
    ran = w.ran.astype(bool)[inner]
    tot = dict(agents=int(ran.sum()), inner=int(inner.sum()), outer=len(outer_idx))
    assert np.array_equal(np.asarray(out['ran'])[gi].astype(bool), ran)
    has = ran & (w.nbr_cnt[inner] >= 0)              # computeNeighbors ran (not the zero-velocity shortcut)
    cnt_full = np.asarray(out['nbr_cnt'])[gi]
    ids_full = np.asarray(out['nbr_ids'])[gi].astype(np.int64)
    col = np.arange(ids_full.shape[1])[None, :]
    live = col < cnt_full[:, None]
    ids_sub = np.where(live, lut[np.where(live, ids_full, n_full)], -1)
    ids_ref = np.where(col < w.nbr_cnt[inner][:, None], w.nbr_ids[inner], -1)
    tot['nbr_outside'] = int((live & (ids_sub < 0))[has].sum())      # a neighbour the sub-world does not hold: box too small
    tot['nbr'] = int(((ids_sub != ids_ref).any(1) | (cnt_full != w.nbr_cnt[inner]))[has].sum())
    dsq_full = np.where(live, np.asarray(out['nbr_dsq'])[gi], 0.0)
    dsq_ref = np.where(col < w.nbr_cnt[inner][:, None], w.nbr_dsq[inner], 0.0)
    tot['dsq'] = int((dsq_full != dsq_ref).any(1)[has].sum())
    tot['saturated'] = int((cnt_full[has] == ids_full.shape[1]).sum())
    tot['tag'] = int((np.asarray(out['tag'])[gi][ran] != w.tag[inner][ran]).sum())
    if policy == 'orca_lp':
        tot['planes'] = int(((np.asarray(out['plane_fail'])[gi] != w.plane_fail[inner])
                             | (np.asarray(out['lp4_proj'])[gi] != w.lp4_proj[inner]))[ran].sum())
    else:
        ns = np.asarray(out['n_suit'])[gi]
        tot['nsuit'] = int(((ns != w.n_suit[inner]) & (ns >= 0))[ran].sum())     # the lazy search reports -1 (unknown)
    dv = np.abs(np.asarray(out['action'])[gi][:, :4].astype('float64') - w.action[inner][:, :4].astype('float64')).max(1)
    tot['vel'] = int((dv[ran] > 2e-5).sum())
    tot['flips'] = int((dv[ran] > 0).sum())
    if post is not None:
        tot['status'] = int(((np.asarray(post['status'])[gi] & 15) != (w.status[inner] & 15)).sum())
        same = dv == 0
        dp = np.abs(np.asarray(post['pos'])[gi] - w.pos[inner]).max(1)
        tot['pos'] = int((dp[same] > 1e-7).sum())
    return tot


EXACT = ('nbr_outside', 'nbr', 'dsq', 'tag', 'planes', 'nsuit', 'vel', 'status', 'pos')


def list_invariants(nbr_cnt, nbr_ids, nbr_dsq, ran, pos, neighbor_dist, sample_stride=97):
    """Whole-swarm invariants of the ordered neighbour lists, evaluated with torch on whatever device the tensors live
    on: lists sorted by distance key, inside neighborDist, without the agent itself or duplicates, and the key is the
    reference's round(|p_i - p_j|^2, 5) (insertAgentNeighbor, agent.py:79-99; util.py:100)."""
    import torch
    n, K = nbr_ids.shape
    dev = nbr_ids.device
    ids = nbr_ids.long()
    col = torch.arange(K, device=dev)[None, :]
    live = (col < nbr_cnt[:, None]) & ran[:, None]
    assert int(nbr_cnt[ran].max()) <= K
    assert bool((nbr_dsq[:, 1:] >= nbr_dsq[:, :-1])[live[:, 1:]].all()), 'neighbour lists are not sorted by distance key'
    assert bool((nbr_dsq[live] < neighbor_dist ** 2).all())
    me = torch.arange(n, device=dev)[:, None]
    assert not bool(((ids == me) & live).any()), 'an agent is its own neighbour'
    srt = torch.where(live, ids, -1 - col.expand_as(ids)).sort(1).values
    assert not bool((srt[:, 1:] == srt[:, :-1]).any()), 'duplicate neighbour'
    rows = torch.arange(0, n, sample_stride, device=dev)
    d2 = ((pos[rows][:, None, :] - pos[ids[rows].clamp(0, n - 1)]) ** 2).sum(2)
    key = torch.round(d2 * 1e5) / 1e5
    ok = (torch.abs(key - nbr_dsq[rows]) <= 1.5e-5) | ~live[rows]      # summation order may move the key by one unit
    assert bool(ok.all())
    return int(live.sum())
