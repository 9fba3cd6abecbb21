"""-m gpu: the result of a rollout does not depend on how the caller lays out the state rows -- ties and the
sequential collision-check order (mampenv.py:42-46) follow the agents' ids (scab_state_ptrs.gid), so bench.py and
ShardedSwarm are free to lay rows out cell by cell (scenarios.spatial_order) for coalesced gathers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_permuted_rows_give_identical_rollout(policy):
    import torch
    from sca_b200 import capi, scenarios
    from sca_b200.swarm import Swarm
    n, steps = 6000, 40
    pos, goal, _ = scenarios.uniform_box_numpy(n, density=2.0 * 1.0e6 / 560.0 ** 3, seed=3)   # dense: collisions do happen
    L = (n / (2.0 * 1.0e6 / 560.0 ** 3)) ** (1.0 / 3.0)
    lo, hi = np.array([-15.0, -15.0, 5.0]), np.array([L + 15.0, L + 15.0, L + 35.0])
    res = []
    for order in (np.arange(n), scenarios.spatial_order(pos), np.random.default_rng(1).permutation(n)):
        sw = Swarm(n, debug=False)
        sw.set_state(pos[order, :3], goal[order, :3], heading=pos[order, 3:])
        sw.set_global_ids(order.astype(np.int32), n)
        sw.fit_grid(lo, hi)
        for _ in range(steps):
            sw.step(policy, capi.NBR_KNN)
        torch.cuda.synchronize()
        d = sw.download()
        inv = np.empty(n, np.int64)
        inv[order] = np.arange(n)
        res.append({k: v[inv] for k, v in d.items()})
        res[-1]['status'] = res[-1]['status'] & 15
    print(policy, 'collisions', int(((res[0]['status'] & 2) != 0).sum()))
    for r in res[1:]:
        for k in res[0]:
            assert np.array_equal(res[0][k], r[k]), (policy, k)
