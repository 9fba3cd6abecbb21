"""CPU: unit/property tests of the oracle's building blocks against direct Python evaluation."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc


@settings(max_examples=2000, deadline=None)
@given(st.floats(min_value=-1e4, max_value=1e4, allow_nan=False))
def test_py_round5_matches_cpython(x):
    L = orc.lib()
    assert L.orc_py_round5(x) == round(x, 5) == L.orc_py_round5_slow(x)


def test_py_round5_ties():
    L = orc.lib()
    for k in range(-2000, 2000):
        x = (k + 0.5) / 1e5
        for y in (x, np.nextafter(x, 1e9), np.nextafter(x, -1e9)):
            assert L.orc_py_round5(float(y)) == round(float(y), 5)


def test_candidates_match_reference_formula():
    # scaPolicy.py:190-200
    for num_n in (256, 128):
        ref = []
        phi = (math.sqrt(5.0) - 1.0) / 2.0
        for rad in np.arange(0.5, 1.0 + 0.03, 1.0 - 0.5):
            for n in range(1, num_n + 1):
                z = (2 * n - 1) / num_n - 1
                x = math.sqrt(1 - z ** 2) * math.cos(2 * math.pi * n * phi)
                y = math.sqrt(1 - z ** 2) * math.sin(2 * math.pi * n * phi)
                ref.append([rad * x, rad * y, rad * z])
        got = orc.gen_candidates(num_n, 1.0)
        assert np.array_equal(got, np.array(ref))


def _world(n, box, seed, m=0):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, box, (n, 3)) + [0, 0, 20]
    goal = rng.uniform(0, box, (n, 3)) + [0, 0, 20]
    obs = None
    if m:
        obs = np.c_[rng.uniform(0, box, (m, 3)) + [0, 0, 20], rng.uniform(0.3, 1.5, m)]
    return orc.World(pos, goal, obstacles=obs)


def test_env_update_fast_equals_literal():
    for seed in range(4):
        a, b = _world(300, 12.0, seed, m=5), _world(300, 12.0, seed, m=5)
        for _ in range(6):
            a.policy_step('orca_lp', orc.NBR_KNN)
            b.policy_step('orca_lp', orc.NBR_KNN)
            da, db = a.env_update(fast=False), b.env_update(fast=True)
            assert da == db
            for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num'):
                assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert (a.status & 2).any()   # the dense box does produce collisions


def test_knn_is_true_16_nearest():
    w = _world(400, 25.0, 7, m=6)
    w.vel[:] = 0.1                      # leave the zero-velocity shortcut
    w.policy_step('rvo3d', orc.NBR_KNN)
    for i in range(0, 400, 17):
        if w.coll_query[i]:
            continue
        d = np.round(((w.pos - w.pos[i]) ** 2).sum(1), 5)
        d[i] = np.inf
        do = (np.array([round(math.sqrt(x), 5) for x in ((w.obs[:, :3] - w.pos[i]) ** 2).sum(1)]) - w.obs[:, 3]) ** 2
        allk = np.r_[do, d]
        uid = np.r_[400 + np.arange(6), np.arange(400)]
        order = np.lexsort((np.r_[np.arange(6), 6 + np.arange(400)], allk))
        order = [o for o in order if allk[o] < 100.0][:16]
        assert w.nbr_cnt[i] == len(order)
        assert np.array_equal(w.nbr_ids[i][:len(order)], uid[order])


def test_compat_equals_knn_when_lists_do_not_saturate():
    # with fewer than 16 in-range candidates the reference's list is the full sorted in-range set
    w1, w2 = _world(150, 40.0, 3), _world(150, 40.0, 3)
    w1.vel[:] = 0.1
    w2.vel[:] = 0.1
    w1.policy_step('rvo3d', orc.NBR_KNN)
    w2.policy_step('rvo3d', orc.NBR_REF_COMPAT)
    small = (w2.nbr_cnt < 16)
    assert small.sum() > 50
    assert np.array_equal(np.sort(w1.nbr_ids[small], axis=1), np.sort(w2.nbr_ids[small], axis=1))
    assert np.array_equal(w1.nbr_dsq[small], w2.nbr_dsq[small])


import pytest


@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_locality_sub_box_reproduces_the_full_world(policy):
    """The property the full-size GPU test relies on (tests/locality.py), checked on the oracle itself: stepping only
    the agents of a box grown by > neighborDist reproduces, for the agents inside the box, the step of the whole world."""
    import locality as loc
    from sca_b200 import scenarios
    n = 12000
    dens = 2.0 * 1.0e6 / 560.0 ** 3
    pos6, goal6, _ = scenarios.uniform_box_numpy(n, dens, seed=4)
    L = (n / dens) ** (1.0 / 3.0)
    orc.set_threads(8)
    full = orc.World(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], initial_pos=pos6, goal_pos=goal6)
    for _ in range(3):
        full.step(policy, orc.NBR_KNN, fast=True)
    pre = {k: getattr(full, k).copy() for k in loc.STATE_KEYS}
    outer_idx, inner = loc.pick_box(pre['pos'], [L / 2, L / 2, L / 2 + 20.0], half_inner=22.0)
    assert inner.sum() > 500 and len(outer_idx) < n // 2
    sub = loc.make_subworld(orc, pre, pos6, goal6, outer_idx)
    full.policy_step(policy, orc.NBR_KNN)
    sub.policy_step(policy, orc.NBR_KNN)
    out = {k: getattr(full, k).copy() for k in ('ran', 'nbr_cnt', 'nbr_ids', 'nbr_dsq', 'tag', 'n_suit', 'plane_fail',
                                                  'lp4_proj', 'action')}
    full.env_update(fast=True)
    sub.env_update(fast=True)
    post = {'status': full.status.copy(), 'pos': full.pos.copy()}
    tot = loc.compare_inner(out, post, sub, outer_idx, inner, n, policy)
    print(policy, tot)
    assert tot['agents'] > 500 and tot['saturated'] > 0
    for k in loc.EXACT:
        assert tot.get(k, 0) == 0, (policy, k, tot)
    assert tot['flips'] == 0
    # the whole-swarm list invariants the full-size GPU test evaluates on the device
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    n_live = loc.list_invariants(t(out['nbr_cnt']), t(out['nbr_ids']), t(out['nbr_dsq']), t(out['ran'].astype(bool)),
                                 t(pre['pos']), 10.0, sample_stride=7)
    assert n_live > 10 * n
