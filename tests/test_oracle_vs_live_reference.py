"""CPU, build container only: differential test of the C oracle against the UNMODIFIED Python reference, run LIVE on
fresh random scenarios (beyond the committed fixtures): random boxes of 10-26 agents at several densities and radii,
a few sphere obstacles, every policy.  The reference is imported from /root/reference through tests/golden/ref_harness.py
(same two shims as the fixture generator); the whole module is skipped where /root/reference does not exist (the GPU
box), and nothing here is on any product, bench or -m gpu path.

Every step of the reference rollout is replayed teacher-forced through the oracle (tests/test_oracle_golden.run_fixture)
and every recorded quantity must match exactly: kd-tree permutation, ordered neighbour lists and keys, branch tags,
suitable-candidate counts, plane counts, planeFail, LP4 projections, v_pref, action rows, flags.
"""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/mamp'), reason='the reference is not on this machine')

# (policy, agents, box half-width, radius, obstacles, steps, seed)
CASES = [
    ('sca', 12, 3.0, 0.5, 0, 7, 11),
    ('sca', 14, 4.0, 0.4, 3, 6, 12),
    ('rvo3d_dubins', 12, 3.5, 0.5, 2, 6, 13),
    ('srvo3d', 22, 3.0, 0.5, 3, 8, 14),
    ('srvo3d', 16, 5.0, 0.7, 0, 8, 15),
    ('rvo3d', 26, 3.0, 0.3, 4, 8, 16),
    ('orca_sampled', 20, 3.0, 0.5, 3, 8, 17),
    ('orca_lp', 26, 2.5, 0.5, 4, 10, 18),
    ('orca_lp', 24, 4.0, 0.8, 0, 10, 19),
]


def _scenario(n, r, n_obs, seed, radius):
    rng = random.Random(seed)
    u = lambda: rng.uniform(-r, r)
    pos = [np.array([u(), u(), u() + 30.0, rng.uniform(0, 2 * np.pi), 0.0, 0.0]) for _ in range(n)]
    goal = [np.array([u(), u(), u() + 30.0, rng.uniform(0, 2 * np.pi), 0.0, 0.0]) for _ in range(n)]
    obs = [[u(), u(), u() + 30.0, rng.uniform(0.3, 1.2)] for _ in range(n_obs)]
    return pos, goal, obs


@pytest.mark.parametrize('case', CASES, ids=lambda c: '%s-n%d-r%g-seed%d' % (c[0], c[1], c[3], c[6]))
def test_oracle_matches_the_live_reference(case):
    policy, n, r, radius, n_obs, steps, seed = case
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_golden as mg
    from test_oracle_golden import run_fixture
    g, done_steps, _ = mg.record(policy, _scenario(n, r, n_obs, seed, radius), keep=steps, radius=radius)
    tot = run_fixture('live %s seed %d' % (policy, seed), g=g)
    print(case, done_steps, tot)
    assert tot['agents'] > 0
    for k, v in tot.items():
        if k in ('agents', 'vel_maxerr', 'n_suit_known', 'angle_ulp_flips'):
            continue
        assert v == 0, (case, k, tot)
