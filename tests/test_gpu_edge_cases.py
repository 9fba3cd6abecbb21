"""-m gpu: edge cases of the C ABI path -- empty and tiny swarms, everything finished, obstacles only,
the two 16-NN code paths against each other and against the oracle, and the error behaviour of the boundary."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POLICIES = ['sca', 'srvo3d', 'rvo3d', 'rvo3d_dubins', 'orca_sampled', 'orca_lp']


def _swarm(pos, goal, obstacles=None, debug=True, **kw):
    from sca_b200.swarm import Swarm
    sw = Swarm(len(pos), debug=debug)
    sw.set_state(pos, goal, **kw)
    sw.set_obstacles(obstacles)
    both = np.r_[pos, goal]
    sw.fit_grid(both.min(0) - 15, both.max(0) + 15)
    return sw


def test_empty_swarm_steps_to_done():
    import ctypes as C
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    sw = Swarm(0, capacity=16)
    assert sw.n == 0
    for pol in POLICIES:
        sw.step(pol, capi.NBR_KNN)
    c = sw.counters()
    assert c['n_active'] == 0 and c['n_not_done'] == 0 and c['n_collision'] == 0
    assert sw.is_done()


@pytest.mark.parametrize('policy', POLICIES)
def test_single_agent_flies_straight_like_the_oracle(policy):
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    pos, goal = np.array([[0.0, 0.0, 20.0]]), np.array([[6.0, 3.0, 24.0]])
    w = orc.World(pos, goal)
    sw = _swarm(pos, goal)
    done = False
    for t in range(200):
        d_o = w.step(policy, orc.NBR_KNN)
        sw.step(policy, capi.NBR_KNN)
        torch.cuda.synchronize()
        a = sw.action.cpu().numpy()[0].astype('float64')
        assert np.abs(a[:4] - w.action[0, :4]).max() <= 2e-5, (policy, t, a, w.action[0])
        assert int(sw.status.cpu().numpy()[0] & 15) == int(w.status[0]), (policy, t)
        assert sw.is_done() == d_o
        if d_o:
            done = True
            break
    assert done and (w.status[0] & 1), 'the lone agent reaches its goal'
    assert np.abs(sw.pos.cpu().numpy()[0] - w.pos[0]).max() <= 1e-4


def test_all_finished_is_a_no_op():
    import torch
    from sca_b200 import capi
    rng = np.random.default_rng(2)
    pos = rng.uniform(0, 30, (64, 3)) + [0, 0, 20]
    sw = _swarm(pos, pos + 0.1)          # everybody starts within near_goal of its goal
    sw.step('sca', capi.NBR_KNN)          # first step: flags AT_GOAL
    torch.cuda.synchronize()
    before = sw.download()
    assert ((before['status'] & 1) == 1).all() and sw.is_done()
    sw.step('sca', capi.NBR_KNN)
    torch.cuda.synchronize()
    after = sw.download()
    assert sw.counters()['n_active'] == 0
    for k in ('pos', 'heading', 'total_dist', 'step_num'):
        assert np.array_equal(before[k], after[k]), k
    assert (after['action'] == 0).all()


@pytest.mark.parametrize('policy', ['sca', 'orca_lp'])
def test_one_agent_among_obstacles_matches_oracle(policy):
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    rng = np.random.default_rng(4)
    pos, goal = np.array([[0.0, 0.0, 25.0]]), np.array([[30.0, 0.0, 25.0]])
    obs = np.c_[rng.uniform(4, 26, 30), rng.uniform(-6, 6, 30), rng.uniform(19, 31, 30), rng.uniform(0.4, 1.2, 30)]
    w = orc.World(pos, goal, obstacles=obs)
    sw = _swarm(pos, goal, obstacles=obs)
    for t in range(120):
        d_o = w.step(policy, orc.NBR_KNN)
        sw.step(policy, capi.NBR_KNN)
        torch.cuda.synchronize()
        if w.nbr_cnt[0] >= 0:
            c = int(w.nbr_cnt[0])
            assert int(sw.nbr_cnt.cpu().numpy()[0]) == c
            assert np.array_equal(sw.nbr_ids.cpu().numpy()[0, :c], w.nbr_ids[0, :c]), (policy, t)
        a = sw.action.cpu().numpy()[0].astype('float64')
        assert np.abs(a[:4] - w.action[0, :4]).max() <= 2e-5, (policy, t)
        assert int(sw.status.cpu().numpy()[0] & 15) == int(w.status[0]), (policy, t)
        if d_o:
            break


def test_knn_integer_key_path_equals_exact_path_and_oracle():
    """No obstacles -> the 32-bit integer-key fast path; one obstacle far away -> knn_exact for every agent.
    Both must produce the oracle's ordered lists (obstacle ids are reported as n + k and never appear here)."""
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi, scenarios
    pos6, goal6, _ = scenarios.uniform_box_numpy(5000, 4.0 * 1.0e6 / 560.0 ** 3, seed=21)
    pos, goal = pos6[:, :3], goal6[:, :3]
    vel = np.random.default_rng(1).normal(0, 0.4, (len(pos), 3)).astype(np.float32)
    w = orc.World(pos, goal, vel=vel)
    w.policy_step('srvo3d', orc.NBR_KNN)
    outs = []
    for obs in (None, np.array([[1.0e5, 1.0e5, 1.0e5, 1.0]])):
        sw = _swarm(pos, goal, obstacles=obs, vel=vel)
        sw.stage_grid()
        sw.stage_neighbors(capi.NBR_KNN)
        torch.cuda.synchronize()
        outs.append((sw.nbr_cnt.cpu().numpy(), sw.nbr_ids.cpu().numpy(), sw.nbr_dsq.cpu().numpy()))
    for cnt, ids, dsq in outs:
        assert np.array_equal(cnt, w.nbr_cnt)
        assert np.array_equal(ids, w.nbr_ids)
        assert np.array_equal(dsq, w.nbr_dsq)
    assert (w.nbr_cnt == 16).mean() > 0.5


def test_knn_reduced_ball_never_hides_a_collider():
    """The 16-NN query starts from last step's 16th-neighbour distance plus two steps of motion.  From the first collider
    on the reference keeps colliders only (agent.py:82-92), and a collider of large radius can sit OUTSIDE that ball while
    16 small agents sit inside it: the ball must never be smaller than r_i + r_max.  Stage-wise calls (no status stage in
    between, which would retire the colliding agents) reach the case; the lists are compared with the oracle's."""
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    rng = np.random.default_rng(5)
    small = []
    while len(small) < 24:                   # 24 small agents within 3 m of agent 0, >= 0.4 m apart
        p = rng.normal(0, 1.0, 3)
        p *= rng.uniform(0.5, 3.0) / np.linalg.norm(p)
        if all(np.linalg.norm(p - q) >= 0.4 for q in small + [np.zeros(3)]):
            small.append(p)
    c = np.array([60.0, 60.0, 60.0])
    pos = np.vstack([c, c + np.array(small), c + [0.0, 5.0 + 0.05 + 0.16, 0.0]])
    n = len(pos)
    goal = pos + [0.0, 40.0, 0.0]
    goal[-1] = pos[-1] - [0.0, 40.0, 0.0]    # the big agent flies at the cluster: 0.2 m closer per step
    radius = np.full(n, 0.05)
    radius[-1] = 5.0
    vel = (0.5 * (goal - pos) / np.linalg.norm(goal - pos, axis=1)[:, None]).astype(np.float32)   # (a zero velocity skips the query)
    sw = _swarm(pos, goal, radius=radius, vel=vel)
    seen = []
    for t in range(4):
        sw.stage_grid(); sw.stage_neighbors(capi.NBR_KNN)
        torch.cuda.synchronize()
        d = sw.download(('pos', 'vel'))
        w = orc.World(d['pos'][:n, :3], goal, radius=radius, vel=d['vel'][:n, :3])
        w.status[:] = sw.status.cpu().numpy()[:n] & 15       # (a query that saw a collider retires the agent, agent.py:83-85)
        w.policy_step('srvo3d', orc.NBR_KNN)
        cnt, ids = sw.nbr_cnt.cpu().numpy()[:n], sw.nbr_ids.cpu().numpy()[:n]
        assert np.array_equal(cnt, w.nbr_cnt), (t, cnt, w.nbr_cnt)
        for i in np.nonzero(cnt >= 0)[0]:
            assert np.array_equal(ids[i, :cnt[i]], w.nbr_ids[i, :cnt[i]]), (t, i)
        seen.append((int(cnt[0]), int(ids[0, 0])))
        sw.stage_policy('srvo3d')
    assert seen[0][0] == 16, 'agent 0 starts with a full list of small neighbours'
    assert (1, n - 1) in seen[1:3], 'then sees the big collider only, and is retired (%r)' % (seen,)


def test_knn_crowded_pass_never_hides_a_collider():
    """Same rule for the crowded-neighbourhood pass, which shrinks the ball from the hit count when more than 80 agents are
    in range: 150 small agents within 3 m and a collider of radius 5 m at 4.9 m."""
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    rng = np.random.default_rng(6)
    small = []
    while len(small) < 150:
        p = rng.normal(0, 1.0, 3)
        p *= 3.0 * rng.uniform(0.05, 1.0) ** (1 / 3) / np.linalg.norm(p)
        if all(np.linalg.norm(p - q) >= 0.25 for q in small + [np.zeros(3)]):
            small.append(p)
    c = np.array([60.0, 60.0, 60.0])
    pos = np.vstack([c, c + np.array(small), c + [0.0, 4.9, 0.0]])
    n = len(pos)
    goal = pos + [0.0, 40.0, 0.0]
    radius = np.full(n, 0.02)
    radius[-1] = 5.0
    vel = np.tile(np.float32([0.0, 0.5, 0.0]), (n, 1))
    w = orc.World(pos, goal, radius=radius, vel=vel)
    w.policy_step('srvo3d', orc.NBR_KNN)
    sw = _swarm(pos, goal, radius=radius, vel=vel)
    sw.stage_grid(); sw.stage_neighbors(capi.NBR_KNN)
    torch.cuda.synchronize()
    cnt, ids = sw.nbr_cnt.cpu().numpy()[:n], sw.nbr_ids.cpu().numpy()[:n]
    assert np.array_equal(cnt, w.nbr_cnt)
    for i in range(n):
        assert np.array_equal(ids[i, :cnt[i]], w.nbr_ids[i, :cnt[i]]), i
    assert cnt[0] == 1 and ids[0, 0] == n - 1 and (cnt == 16).sum() > 20


def test_boundary_errors_are_reported_not_swallowed():
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    sw = Swarm(8)
    sw.set_state(np.zeros((8, 3)) + np.arange(8)[:, None] * 3.0, np.ones((8, 3)) * 50)
    sw.fit_grid()
    with pytest.raises(capi.ScabError, match='bad policy'):
        sw.h.policy_step(99)
    with pytest.raises(capi.ScabError, match='bad mode'):
        sw.h.neighbors(7)
    with pytest.raises(capi.ScabError, match='REF_COMPAT needs scab_kdtree_build'):
        sw.h.neighbors(capi.NBR_REF_COMPAT)
    with pytest.raises(capi.ScabError, match='exceeds capacity'):
        sw._bind(9)
    with pytest.raises(capi.ScabError, match='cell_size < neighbor_dist'):
        sw.h.set_grid([0, 0, 0], 5.0, [4, 4, 4])
    with pytest.raises(capi.ScabError, match='never enabled'):
        sw.h.dubins_read(8)


def test_voxel_obstacle_field_at_scale_matches_oracle():
    """every occupied voxel of a binvox map as an r = 0.2 sphere (~1.7e4 obstacles in the library's
    static grid); agents fly past and over the field.  Ordered lists (obstacle ids included), tags and
    velocities against the oracle."""
    import io
    import os
    from sca_b200 import maps
    from test_gpu_knn_oracle import _compare
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'map_synthetic.npz'))
    obs = maps.obstacles_from_binvox(io.BytesIO(g['binvox'].tobytes()), center=(0.0, 0.0), above_every=1, floor_every=1)
    assert len(obs) > 15000
    c = obs[:, :3].mean(0)
    n = 96
    ang = 2 * np.pi * np.arange(n) / n
    pos = np.zeros((n, 6)); goal = np.zeros((n, 6))
    ring = 6.0 + 0.5 * (np.arange(n) % 3)
    pos[:, 0], pos[:, 1], pos[:, 2] = c[0] + ring * np.cos(ang), c[1] + ring * np.sin(ang), c[2] + 1.5 + 0.6 * (np.arange(n) % 4)
    goal[:, 0], goal[:, 1], goal[:, 2] = c[0] - ring * np.cos(ang), c[1] - ring * np.sin(ang), pos[:, 2]
    pos[:, 3] = ang + np.pi
    for policy in ('srvo3d', 'orca_lp'):
        tot = _compare(pos, goal, obs, policy, steps=12)
        print(policy, tot)
        for k in ('nbr', 'tag', 'nsuit', 'vel', 'status', 'planes'):
            assert tot[k] == 0, (policy, tot)


def test_fast_agents_collision_flags_match_the_oracle():
    """the collision stage's reach is derived from the motion bound (v_max dt), not from a fixed 1.25 m per
    step.  pref_speed = 14 m/s moves an agent 1.4 m per step; the reference's O(N^2) check (mampenv.py:61-80) has no
    reach limit, so the flags must still agree with the oracle's literal loop."""
    import torch
    from oracle import oracle as orc
    from sca_b200 import capi
    rng = np.random.default_rng(11)
    n = 400
    pos = rng.uniform(0, 28, (n, 3)) + [0, 0, 20]
    keep = np.ones(n, bool)
    for i in range(n):                      # >= 1.3 m apart at t = 0
        if keep[i]:
            d = np.linalg.norm(pos - pos[i], axis=1)
            keep &= ~((d < 1.3) & (np.arange(n) > i))
    pos = pos[keep]
    n = len(pos)
    goal = pos[rng.permutation(n)] + rng.normal(0, 2.0, (n, 3))
    w = orc.World(pos, goal, pref_speed=14.0)
    sw = _swarm(pos, goal, pref_speed=14.0)
    assert sw.v_bound == 14.0
    hits = 0
    for t in range(6):
        w.step('rvo3d', orc.NBR_KNN)         # fast=False: the literal sequential O(N^2) loop
        sw.step('rvo3d', capi.NBR_KNN)
        torch.cuda.synchronize()
        st = sw.status.cpu().numpy()[:n] & 15
        assert np.array_equal(st, w.status), (t, np.nonzero(st != w.status)[0][:10])
        hits = int((st & 2).astype(bool).sum())
    assert hits >= 4, 'the scenario must actually produce collisions (%d)' % hits
    sw.counters()                            # no sticky flag


def test_a_move_beyond_the_motion_bound_is_loud():
    import torch
    from sca_b200 import capi
    rng = np.random.default_rng(3)
    pos = rng.uniform(0, 40, (50, 3)) + [0, 0, 20]
    sw = _swarm(pos, pos[::-1].copy())
    sw.h.set_motion_bound(0.2)               # claims 0.02 m per step; the agents fly 0.1 m per step
    sw.step('rvo3d', capi.NBR_KNN)
    sw.step('rvo3d', capi.NBR_KNN)
    torch.cuda.synchronize()
    with pytest.raises(capi.ScabError, match='motion bound'):
        sw.counters()
    assert sw.counters(check=False)['overflow'] == capi.OVF_MOTION
    sw.h.reset_counters()
    sw.h.set_motion_bound(1.0)
    sw.step('rvo3d', capi.NBR_KNN)
    assert sw.counters()['overflow'] == 0


def test_a_radius_above_the_declared_bound_is_loud():
    """scab_set_max_radius is a promise the 16-NN query ball and the status shortcut rely on; a larger radius in the state
    raises the sticky overflow flag at the next cell-list build instead of silently dropping a collider."""
    from sca_b200 import capi
    rng = np.random.default_rng(3)
    pos = rng.uniform(0, 40, (200, 3)) + [0, 0, 20]
    goal = pos + [30.0, 0, 0]
    sw = _swarm(pos, goal, radius=np.where(np.arange(200) == 17, 1.5, 0.5))
    sw.step('rvo3d', capi.NBR_KNN)
    sw.counters()                            # Swarm.set_state declared max(radius) = 1.5: fine
    sw.h.set_max_radius(1.0)
    sw.step('rvo3d', capi.NBR_KNN)
    with pytest.raises(capi.ScabError, match='radius'):
        sw.counters()
    assert sw.counters(check=False)['overflow'] & capi.OVF_RADIUS


def test_status_stage_refuses_a_reach_beyond_the_cell_size():
    from sca_b200 import capi
    pos = np.array([[0.0, 0.0, 20.0], [8.0, 0.0, 20.0]])
    sw = _swarm(pos, pos[::-1].copy())
    sw.h.set_motion_bound(60.0)              # 6 m per step: partners may sit outside the 27-cell block of 10 m cells
    with pytest.raises(capi.ScabError, match='exceeds the cell size'):
        sw.step('rvo3d', capi.NBR_KNN)
    sw.fit_grid(pos.min(0) - 30, pos.max(0) + 30, cell=30.0)
    sw.step('rvo3d', capi.NBR_KNN)


def test_scab_create_validates_its_parameters():
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    for field, bad in (('cell_size', 5.0), ('time_step', 0.0), ('time_horizon', -1.0), ('neighbor_dist', float('nan'))):
        p = capi.default_params()
        setattr(p, field, bad)
        with pytest.raises(capi.ScabError, match='scab_create failed'):
            Swarm(4, params=p)


def test_halo_overflow_is_loud(request, monkeypatch):
    """A halo face buffer that is too small drops ghost records: the sticky device flag turns that into an error at the next
    counters() / is_done() instead of a quietly wrong rollout."""
    device = 'cuda:0'
    if request.config.getoption('--hostsim'):
        import hostsim_util
        hostsim_util.activate_streams(monkeypatch)
        device = 'cpu'
    from sca_b200 import capi, scenarios
    from sca_b200.sharding import LogicalShards
    n = 3000
    density = 4.0e6 / 560.0 ** 3
    pos, goal, _ = scenarios.uniform_box_numpy(n, density, seed=3)
    L = (n / density) ** (1 / 3)
    lo, hi = np.array([-20.0, -20.0, 0.0]), np.array([L + 20.0, L + 20.0, L + 40.0])
    ls = LogicalShards(pos, goal, 2, device, lo, hi, halo_cap=16)
    ls.step('rvo3d')
    with pytest.raises(capi.ScabError, match='halo face buffer overflowed'):
        ls.counters()
    ls.close()
