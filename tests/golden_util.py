"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle/golden (GPU) parity tests."""
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, 'golden')

TAG_FIRST, TAG_SUITABLE, TAG_FALLBACK, TAG_LP3_OK, TAG_LP4 = 1, 2, 4, 8, 16
ST_AT_GOAL, ST_COLLISION, ST_TIMEOUT, ST_RUN_DONE = 1, 2, 4, 8
K = 16


def golden_names():
    # per-step fixtures only (map_*.npz holds the binvox reader's golden vector, rollout_*.npz the end state and metrics
    # of reference rollouts run to completion -- tests/test_rollout_metrics.py)
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))
                  if not os.path.basename(p).startswith(('map_', 'rollout_')))


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return {k: d[k] for k in d.files}


def uses_external_vpref(policy):
    # SCA / RVO3D+Dubins take v_pref from the (upstream, stateful) Dubins provider: SURVEY.md 8(f1)
    return policy in ('sca', 'rvo3d_dubins')


def state_at(g, t):
    """Pre-step state of fixture step t (t == T gives the state after the last kept step)."""
    T = g['pos'].shape[0]
    if t < T:
        return {k: g[k][t] for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')}
    return {k: g['after_' + k] for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')}


def compare_step(g, t, out, check_nbrs=True, vel_tol=2e-5, who='', dsq_rtol=0.0):
    """Compare one step's outputs `out` (dict of arrays, same keys as the fixture) with fixture step t.

    Returns a dict of mismatch counts; exactness requirements are asserted by the callers.
    """
    ran = g['ran'][t].astype(bool)
    res = {'agents': int(ran.sum())}
    assert np.array_equal(out['ran'].astype(bool), ran), who + ' active mask'
    res['tag'] = int((out['tag'][ran] != g['tag'][t][ran]).sum())
    res['coll_query'] = int((out['coll_query'][ran] != g['coll_query'][t][ran]).sum())
    if check_nbrs:
        nv = ran & g['nbr_valid'][t].astype(bool)
        res['nbr_cnt'] = int((out['nbr_cnt'][nv] != g['nbr_cnt'][t][nv]).sum())
        res['nbr_ids'] = int((out['nbr_ids'][nv] != g['nbr_ids'][t][nv]).any(axis=1).sum())
        # agent keys are decimal-rounded and must match exactly; obstacle keys (python float ** 2 -> libm pow)
        # are allowed dsq_rtol by the GPU tests
        dd = np.abs(out['nbr_dsq'][nv] - g['nbr_dsq'][t][nv])
        res['nbr_dsq'] = int((dd > dsq_rtol * np.abs(g['nbr_dsq'][t][nv])).any(axis=1).sum())
    if 'n_suit' in out and (g['tag'][t] & (TAG_SUITABLE | TAG_FALLBACK)).any():
        # -1 = not computed (lazy tiers decided before every candidate was examined, SCAB_OPT_FULL_SEARCH off)
        ns = out['n_suit'][ran]
        res['n_suit'] = int(((ns != g['n_suit'][t][ran]) & (ns >= 0)).sum())
        res['n_suit_known'] = int((ns >= 0).sum())
    if str(g['policy']) == 'orca_lp':
        res['plane_fail'] = int((out['plane_fail'][ran] != g['plane_fail'][t][ran]).sum())
        res['n_planes'] = int((out['n_planes'][ran] != g['n_planes'][t][ran]).sum())
        if 'lp4_proj' in out:
            res['lp4_proj'] = int((out['lp4_proj'][ran] != g['lp4_proj'][t][ran]).sum())
    a_ref = g['action'][t].astype(np.float32)
    a_out = out['action'].astype(np.float32)
    dv = np.abs(a_out[:, :4].astype('float64') - a_ref[:, :4].astype('float64')).max(axis=1)
    res['vel_bad'] = int((dv[ran] > vel_tol).sum())
    res['vel_maxerr'] = float(dv[ran].max()) if ran.any() else 0.0
    # velocity + speed (columns 0-3) are expected bit-identical as float32; the heading deltas (4-5) are
    # atan2(...) - heading, a cancelling difference where the device libm's last-ulp fp64 differences
    # stay visible: they get an absolute tolerance instead
    res['action_exact'] = int((a_out[ran][:, :4] != a_ref[ran][:, :4]).any(axis=1).sum())
    da = np.abs(a_out[:, 4:].astype('float64') - a_ref[:, 4:].astype('float64')).max(axis=1)
    # > 1e-9: counted (a last-ulp fp64 libm difference that flips the float32 rounding of a heading delta);
    # > 5e-7 (one float32 ulp at |angle| <= 2 pi): a real mismatch
    res['angle_ulp_flips'] = int((da[ran] > 1e-9).sum())
    res['angle_bad'] = int((da[ran] > 5e-7).sum())
    return res
