"""CPU, build container only: the drop-in recipe of INTEGRATION.md executed against the REAL reference package.

    agents = [Agent(..., policy=GPUSCAPolicy, ...)]; env = MACAEnv(); env.set_agents(agents, obstacles)
    env.kdTree = GPUKDTree(agents, obstacles)

An UNMODIFIED `mamp.envs.mampenv.MACAEnv` with the real `mamp.agents.agent.Agent` objects is driven through
`sca_b200.mamp_compat` (GPUKDTree + GPU*Policy) for a whole rollout and compared, step by step, with the same rollout under
the reference's own policy classes and kd-tree.  This container has no GPU, so the device-resident `Swarm` behind the
shim is replaced by `OracleSwarm` below -- the same method surface, the oracle as the stepper (test infrastructure; the
GPU-side counterpart with the real `Swarm` is tests/test_gpu_mamp_compat.py).  What is under test is the HOST logic of the
shim: the Agent attribute snapshot, the pinned staging buffers, what `find_next_action` returns and back-fills
(`agent.neighbors`, `v_pref`, `is_collision`, `path`, `policy.now_goal`), the upstream Dubins provider hand-off, and
that nothing in `mamp/envs` has to change.  Skipped where /root/reference does not exist (the GPU box).
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/mamp'), reason='the reference is not on this machine')


class OracleSwarm:
    """The subset of sca_b200.swarm.Swarm that mamp_compat.GPUKDTree uses, stepped by the oracle on the host."""

    def __init__(self, n, device='cpu', debug=True, **kw):
        self.n = n
        f8, f4, i4 = torch.float64, torch.float32, torch.int32
        self.t = {'pos': torch.zeros((n, 3), dtype=f8), 'vel': torch.zeros((n, 3), dtype=f4), 'heading': torch.zeros((n, 3), dtype=f8),
                  'status': torch.zeros(n, dtype=i4), 'total_dist': torch.zeros(n, dtype=f8), 'step_num': torch.zeros(n, dtype=i4),
                  'action': torch.zeros((n, 7), dtype=f4), 'nbr_cnt': torch.zeros(n, dtype=i4), 'nbr_ids': torch.zeros((n, 16), dtype=i4),
                  'nbr_dsq': torch.zeros((n, 16), dtype=f8), 'vpref_out': torch.zeros((n, 3), dtype=f8)}
        self._args, self._obs, self._wp, self.w, self._mode = None, None, None, None, 1

    def __getattr__(self, k):
        t = self.__dict__.get('t')
        if t is not None and k in t:
            return t[k]
        raise AttributeError(k)

    def set_state(self, pos, goal, **kw):
        self._args = (pos, goal, kw)

    def fit_grid(self, *a, **kw):
        return self

    def set_obstacles(self, xyzr):
        self._obs = np.asarray(xyzr, 'float64').reshape(-1, 4)
        self.w = None

    def set_waypoints(self, wp, counts):
        self._wp = (np.asarray(wp, 'float64'), np.asarray(counts, np.int32))
        self.t['waypoint_left'] = torch.zeros(self.n, dtype=torch.int32)
        self.t['now_goal'] = torch.full((self.n, 3), float('nan'), dtype=torch.float64)

    def set_external_vpref(self, v):
        self._vp = None if v is None else np.array(v.numpy() if torch.is_tensor(v) else v, 'float64')

    def _world(self):
        if self.w is None:
            from oracle import oracle as orc
            pos, goal, kw = self._args
            self.w = orc.World(pos, goal, heading=kw['heading'], radius=kw['radius'], pref_speed=kw['pref_speed'],
                               obstacles=self._obs, initial_pos=kw['initial_pos'], goal_pos=kw['goal_pos'])
            assert np.array_equal(self.w.max_run_dist, kw['max_run_dist'])
            if self._wp is not None:
                self.w.set_waypoints(*self._wp)
        return self.w

    def upload(self, h):
        w = self._world()
        for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num'):
            getattr(w, k)[:] = h[k].numpy()

    def stage_grid(self):
        pass

    def stage_neighbors(self, mode, pos_host=None):
        self._mode = mode

    def stage_policy(self, policy):
        w = self._world()
        w.set_vpref_ext(getattr(self, '_vp', None))
        w.policy_step(policy, self._mode)
        put = lambda k, a: self.t[k].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        put('action', w.action)
        put('nbr_cnt', w.nbr_cnt)
        put('nbr_ids', w.nbr_ids)
        put('nbr_dsq', w.nbr_dsq)
        put('vpref_out', w.vpref_store)
        put('status', w.status | (2 * (w.coll_query != 0)).astype(np.int32))      # agent.py:83-85, as the policy kernels do
        if self._wp is not None:
            put('waypoint_left', w.waypoint_left)
            put('now_goal', w.now_goal)


def _snapshot(agents):
    return dict(pos=np.array([a.pos_global_frame for a in agents], 'float64'),
                vel=np.array([np.asarray(a.vel_global_frame, 'float64') for a in agents]),
                heading=np.array([np.asarray(a.heading_global_frame, 'float64') for a in agents]),
                flags=np.array([[a.is_at_goal, a.is_collision, a.is_out_of_max_time] for a in agents]),
                total_dist=np.array([a.total_dist for a in agents]), step_num=np.array([a.step_num for a in agents]),
                v_pref=np.array([np.asarray(a.v_pref, 'float64') for a in agents]),
                path_left=np.array([len(a.path) for a in agents]))


def _nbr_ids(agents):
    n = len(agents)
    return [[(o.id + n if o.is_obstacle else o.id) for o, _ in a.neighbors] for a in agents]


# (policy, shim class, agents, box half-width, obstacles, steps, with waypoints)
CASES = [('srvo3d', 'GPUSRVO3DPolicy', 14, 4.0, 2, 24, False),
         ('rvo3d', 'GPURVO3DPolicy', 12, 4.0, 0, 30, True),
         ('orca_lp', 'GPUORCA3DPolicyOfficial', 20, 3.0, 3, 40, False),
         ('orca_sampled', 'GPUORCA3DPolicy', 12, 4.0, 2, 30, True),
         ('sca', 'GPUSCAPolicyHostDubins', 10, 5.0, 2, 30, False),
         ('rvo3d_dubins', 'GPURVO3dDubinsPolicyHostDubins', 8, 5.0, 0, 25, False)]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_unmodified_macaenv_runs_on_the_shims(case, monkeypatch):
    policy, shim_cls, n, r, n_obs, steps, with_paths = case
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import ref_harness as rh
    from test_oracle_vs_live_reference import _scenario
    Agent = rh.import_reference()
    from mamp.agents.obstacle import Obstacle
    from mamp.envs.mampenv import MACAEnv
    from sca_b200 import mamp_compat
    # no CUDA in this container: the Swarm is the oracle-backed stand-in, pinning / synchronising are no-ops
    monkeypatch.setattr(mamp_compat, 'Swarm', OracleSwarm)
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)

    pos, goal, obs = _scenario(n, r, n_obs, seed=100 + n, radius=0.5)

    def make(policy_cls, use_shim):
        agents = [Agent(start_pos=pos[i], goal_pos=goal[i], vel=[0.0, 0.0, 0.0], radius=0.5, pref_speed=1.0,
                        policy=policy_cls, id=i, dt=0.1) for i in range(n)]
        if with_paths:          # a one- or two-waypoint dog-leg per agent (popped from the end of the list)
            for i, a in enumerate(agents):
                s, g = np.array(pos[i][:3]), np.array(goal[i][:3])
                a.path = [(s + 0.7 * (g - s) + [0.0, 1.5, 0.0]).tolist(), (s + 0.3 * (g - s) + [0.0, 0.0, 1.5]).tolist()][i % 2:]
        obstacles = [Obstacle(pos=o[:3], shape_dict={'shape': 'sphere', 'feature': o[3]}, id=k) for k, o in enumerate(obs)]
        env = MACAEnv()
        with contextlib.redirect_stdout(io.StringIO()):
            env.set_agents(agents, obstacles=obstacles)
        if use_shim:
            env.kdTree = mamp_compat.GPUKDTree(agents, obstacles)          # <- the one added line of INTEGRATION.md
        return env, agents

    env_ref, ag_ref = make(rh.policy_class(policy), False)
    env_new, ag_new = make(getattr(mamp_compat, shim_cls), True)
    assert type(env_new) is MACAEnv and isinstance(env_new.kdTree, mamp_compat.GPUKDTree)
    worst = 0.0
    for t in range(steps):
        with contextlib.redirect_stdout(io.StringIO()):
            d_ref = env_ref.step({})
            d_new = env_new.step({})
        a, b = _snapshot(ag_ref), _snapshot(ag_new)
        assert d_ref == d_new, (policy, t)
        assert np.array_equal(a['flags'], b['flags']) and np.array_equal(a['step_num'], b['step_num']), (policy, t)
        assert np.array_equal(a['vel'], b['vel']), (policy, t)                  # float32 vx vy vz: bit for bit
        assert np.array_equal(a['path_left'], b['path_left']), (policy, t)
        # heading deltas are atan2 differences cast to float32: a last-ulp libm difference may flip that rounding
        worst = max(worst, float(np.abs(a['pos'] - b['pos']).max()), float(np.abs(a['heading'] - b['heading']).max()))
        assert worst < 1e-6, (policy, t, worst)
        active = ~a['flags'].any(1)
        ids_ref, ids_new = _nbr_ids(ag_ref), _nbr_ids(ag_new)
        for i in np.nonzero(active)[0]:
            assert ids_ref[i] == ids_new[i], (policy, t, i)
            assert np.allclose(a['v_pref'][i], b['v_pref'][i], rtol=0, atol=1e-12), (policy, t, i)
        if d_ref:
            break
    print(policy, 'steps', t + 1, 'worst |d pos / heading|', worst)
