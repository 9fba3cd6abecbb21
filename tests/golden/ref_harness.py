"""Reference harness: imports the UNMODIFIED wuuya1/SCA reference from /root/reference
(this container only -- the GPU box has no /root/reference) with the two shims of SURVEY.md 8(c)
and records, per step and per agent, everything the parity tests need.

Only used by tests/golden/make_golden.py (fixture generation) -- never at test/bench run time.

Shims (nothing in the reference is edited):
  1. stub `matplotlib` modules (mamp/policies/sca/dubinsmaneuver2d.py:13, dubinsmaneuver3d.py:11
     import pyplot for plot helpers only);
  2. `Agent.to_vector = no-op` (mamp/agents/agent.py:148 calls DataFrame.append, removed in pandas 2).
Instrumentation is done by wrapping module-level functions / methods (monkey-patching attributes),
so the reference's own code computes every recorded value.
"""
import contextlib
import io
import math
import random
import sys
import types

import numpy as np

REF = '/root/reference'

TAG_FIRST = 1      # zero-velocity shortcut (vA_post = 0.3 v_pref)
TAG_SUITABLE = 2   # sampled policies: at least one suitable candidate
TAG_FALLBACK = 4   # sampled policies: compute_without_suitV path
TAG_LP3_OK = 8     # ORCA-LP: linearProgram3 returned len(planes)
TAG_LP4 = 16       # ORCA-LP: linearProgram4 ran

ST_AT_GOAL, ST_COLLISION, ST_TIMEOUT, ST_RUN_DONE = 1, 2, 4, 8


def import_reference():
    for m in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'mpl_toolkits',
              'mpl_toolkits.mplot3d']:
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from mamp.agents.agent import Agent
    Agent.to_vector = lambda self: None
    return Agent


def policy_module(name):
    import importlib
    import_reference()
    return importlib.import_module({
        'sca': 'mamp.policies.sca.scaPolicy',
        'rvo3d_dubins': 'mamp.policies.sca.rvo3dDubinsPolicy',
        'rvo3d': 'mamp.policies.rvo3dPolicy',
        'srvo3d': 'mamp.policies.srvo3dPolicy',
        'orca_sampled': 'mamp.policies.orca3dPolicy',
        'orca_lp': 'mamp.policies.orca3dPolicyOfficial',
    }[name])


def policy_class(name):
    mod = policy_module(name)
    return getattr(mod, {
        'sca': 'SCAPolicy', 'rvo3d_dubins': 'RVO3dDubinsPolicy', 'rvo3d': 'RVO3DPolicy',
        'srvo3d': 'SRVO3DPolicy', 'orca_sampled': 'ORCA3DPolicy', 'orca_lp': 'ORCA3DPolicy'}[name])


# ----------------------------------------------------------------------------------------------
# Scenario generators: restated from run_example/ with explicit seeds (the reference is unseeded).
# ----------------------------------------------------------------------------------------------
def mod2pi(theta):
    return theta - 2.0 * math.pi * math.floor(theta / 2.0 / math.pi)


def scen_circle(n, rad):
    """run_example/run_sca.py:17-30 (goal carries the partner's heading)."""
    pos = []
    for j in range(n):
        pos.append([round(rad * np.cos(2 * j * np.pi / n), 2), round(rad * np.sin(2 * j * np.pi / n), 2),
                    10.0, round(mod2pi(2 * j * np.pi / n + np.pi), 5), 0, 0])
    goal = [pos[(j + int(n / 2)) % n] for j in range(n)]
    return pos, goal, []


def scen_sphere(n, rad=25.0, zc=30.0):
    """run_example/run_orca.py:36-54."""
    pos, goal = [], []
    phi = (math.sqrt(5.0) - 1.0) / 2.0
    for k in range(1, n + 1):
        z = (2 * k - 1) / n - 1
        x = math.sqrt(1 - z ** 2) * math.cos(2 * math.pi * k * phi)
        y = math.sqrt(1 - z ** 2) * math.sin(2 * math.pi * k * phi)
        p = np.array([rad * x, rad * y, rad * z, 0.0, 0.0, 0.0])
        pos.append(p)
        goal.append(-p)
    for i in range(n):
        pos[i][2] += zc
        goal[i][2] += zc
    return pos, goal, []


def scen_random(n, r=15.0, zc=30.0, seed=0):
    """run_example/run_orca.py:57-73 with random.seed(seed)."""
    random.seed(seed)
    pos = [np.array([random.uniform(-r, r), random.uniform(-r, r), random.uniform(-r, r), 0.0, 0.0, 0.0])
           for _ in range(n)]
    goal = [np.array([random.uniform(-r, r), random.uniform(-r, r), random.uniform(-r, r), 0.0, 0.0, 0.0])
            for _ in range(n)]
    for i in range(n):
        pos[i][2] += zc
        goal[i][2] += zc
    return pos, goal, []


def scen_takeoff_landing(n=16):
    """run_example/run_sca.py:53-81 + the 8 exp2 sphere obstacles of :139-150."""
    rad = 4.0
    landing = n - int(n / 2)
    takeoff = int(n / 2)
    pos = []
    for j in range(landing):
        pos.append([round(rad * np.cos(2 * j * np.pi / landing), 2), round(rad * np.sin(2 * j * np.pi / landing), 2),
                    10.0, round(np.pi / 2, 5), 0, 0])
    for j in range(landing, n):
        pos.append([round(rad * np.cos(2 * j * np.pi / takeoff), 2), round(rad * np.sin(2 * j * np.pi / takeoff), 2),
                    0.0, round(-np.pi / 2, 5), 0, 0])
    goal = [pos[j + landing] for j in range(landing)] + [pos[j - takeoff] for j in range(landing, n)]
    obs = [[round(rad * np.cos(2 * j * np.pi / 8), 2), round(rad * np.sin(2 * j * np.pi / 8), 2), 5.0, 1.0]
           for j in range(8)]
    return pos, goal, obs


# ----------------------------------------------------------------------------------------------
_ACTIVE = None


class _Proxy:
    """Attribute proxy to the active Recorder (module functions are wrapped once per process)."""

    def __getattr__(self, k):
        return getattr(_ACTIVE, k)


class Recorder:
    """Builds a reference MACAEnv for (policy, scenario) and records per-step data."""

    def __init__(self, policy, pos, goal, obstacles, radius=0.5, pref_speed=1.0, paths=None):
        Agent = import_reference()
        from mamp.agents.obstacle import Obstacle
        from mamp.envs.mampenv import MACAEnv
        self.policy = policy
        self.mod = policy_module(policy)
        cls = policy_class(policy)
        self.agents = [Agent(start_pos=pos[i], goal_pos=goal[i], vel=[0.0, 0.0, 0.0], radius=radius,
                             pref_speed=pref_speed, policy=cls, id=i, dt=0.1) for i in range(len(pos))]
        self.obstacles = [Obstacle(pos=o[:3], shape_dict={'shape': 'sphere', 'feature': o[3]}, id=i)
                          for i, o in enumerate(obstacles)]
        # waypoint lists of an external global planner (agent.path, agent.py:44): the policies pop() from the END
        self.paths = paths
        if paths is not None:
            for a, pth in zip(self.agents, paths):
                a.path = [list(map(float, w)) for w in pth]
        self.env = MACAEnv()
        with contextlib.redirect_stdout(io.StringIO()):
            self.env.set_agents(self.agents, obstacles=self.obstacles)
        self.n = len(self.agents)
        self.steps = []
        self._install()

    # -- instrumentation ---------------------------------------------------------------------
    def _install(self):
        global _ACTIVE
        _ACTIVE = self
        mod = self.mod
        rec = _Proxy()   # wrappers always talk to the most recently built Recorder
        self.cur = None  # dict for the agent currently inside find_next_action
        for ag in self.agents:
            self._wrap_policy(ag)
        if getattr(mod, '_scab_patched', False):
            return
        mod._scab_patched = True

        orig_cn = mod.computeNeighbors

        def computeNeighbors(agent, kdTree):
            rec.cur['nbr_valid'] = 1
            rec.cur['coll_before'] = int(agent.is_collision)
            return orig_cn(agent, kdTree)
        mod.computeNeighbors = computeNeighbors

        orig_vp = mod.compute_v_pref

        def compute_v_pref(*a):
            v = orig_vp(*a)
            rec.cur['vpref_in'] = np.array(v, dtype='float64')
            return v
        mod.compute_v_pref = compute_v_pref

        if hasattr(mod, 'intersect'):
            orig_suit = mod.compute_newV_is_suit

            def compute_newV_is_suit(*a):
                s = orig_suit(*a)
                rec.cur['n_suit'] += int(bool(s))
                return s
            mod.compute_newV_is_suit = compute_newV_is_suit
            orig_ws = mod.compute_without_suitV

            def compute_without_suitV(*a):
                rec.cur['tag'] |= TAG_FALLBACK
                return orig_ws(*a)
            mod.compute_without_suitV = compute_without_suitV
            orig_int = mod.intersect

            def intersect(*a):
                rec.cur['tag'] &= ~TAG_FIRST
                v = orig_int(*a)
                if not rec.cur['tag'] & TAG_FALLBACK:
                    rec.cur['tag'] |= TAG_SUITABLE
                return v
            mod.intersect = intersect
        if self.policy == 'orca_lp':
            cls = mod.ORCA3DPolicy
            orig_lp3, orig_lp4 = cls.linearProgram3, cls.linearProgram4

            def lp3(self_, planes, maxSpeed, vel_pref, dir_opt=False):
                r = orig_lp3(self_, planes, maxSpeed, vel_pref, dir_opt)
                if not dir_opt:
                    rec.cur['tag'] &= ~TAG_FIRST
                    rec.cur['plane_fail'] = r
                    rec.cur['n_planes'] = len(planes)
                    if r == len(planes):
                        rec.cur['tag'] |= TAG_LP3_OK
                else:
                    rec.cur['lp4_proj'] += 1
                    rec.cur['lp4_inner_fail'] += int(r < len(planes))
                return r

            def lp4(self_, planes, beginPlane, radius):
                rec.cur['tag'] |= TAG_LP4
                return orig_lp4(self_, planes, beginPlane, radius)
            cls.linearProgram3, cls.linearProgram4 = lp3, lp4

    def _wrap_policy(self, ag):
        rec = self
        orig = ag.policy.find_next_action

        def find_next_action(dict_comm, agent, kdTree):
            rec.cur = {'tag': TAG_FIRST, 'n_suit': 0, 'nbr_valid': 0, 'plane_fail': -1, 'n_planes': -1,
                       'lp4_proj': 0, 'lp4_inner_fail': 0, 'coll_before': 0,
                       'vpref_in': np.zeros(3)}
            act = orig(dict_comm, agent, kdTree)
            rec.per_agent[agent.id] = rec.cur
            rec.cur['action'] = np.asarray(act, dtype='float64')
            rec.cur['nbrs'] = [(('o', o.id) if o.is_obstacle else ('a', o.id), d) for o, d in agent.neighbors]
            rec.cur['coll_after'] = int(agent.is_collision)
            rec.cur['vpref_store'] = np.array(agent.v_pref, dtype='float64')
            return act
        ag.policy.find_next_action = find_next_action

    # -- state snapshots ---------------------------------------------------------------------
    def snapshot(self):
        ags = self.agents
        st = np.array([(ST_AT_GOAL * a.is_at_goal) | (ST_COLLISION * a.is_collision) |
                       (ST_TIMEOUT * a.is_out_of_max_time) | (ST_RUN_DONE * a.is_run_done) for a in ags],
                      dtype=np.int32)
        return dict(
            pos=np.array([a.pos_global_frame for a in ags], dtype='float64'),
            vel=np.array([np.asarray(a.vel_global_frame, dtype='float64') for a in ags]),
            heading=np.array([np.asarray(a.heading_global_frame, dtype='float64') for a in ags]),
            status=st,
            total_dist=np.array([a.total_dist for a in ags], dtype='float64'),
            step_num=np.array([a.step_num for a in ags], dtype=np.int32),
            vpref_prev=np.array([np.asarray(a.v_pref, dtype='float64') for a in ags]),
            path_left=np.array([len(a.path) for a in ags], dtype=np.int32),
            now_goal=np.array([np.full(3, np.nan) if getattr(a.policy, 'now_goal', None) is None
                               else np.asarray(a.policy.now_goal, dtype='float64') for a in ags]),
        )

    def step(self):
        n = self.n
        pre = self.snapshot()
        self.per_agent = {}
        with contextlib.redirect_stdout(io.StringIO()):
            done = self.env.step({})
        out = dict(pre)
        K = 16
        out['agent_ids_perm'] = np.array(self.env.kdTree.agentIDs, dtype=np.int32)
        out['ran'] = np.zeros(n, np.int32)
        out['nbr_valid'] = np.zeros(n, np.int32)
        out['nbr_cnt'] = np.zeros(n, np.int32)
        out['nbr_ids'] = np.full((n, K), -1, np.int32)   # agent id, or n + obstacle id
        out['nbr_dsq'] = np.zeros((n, K), 'float64')
        out['coll_query'] = np.zeros(n, np.int32)
        out['tag'] = np.zeros(n, np.int32)
        out['n_suit'] = np.zeros(n, np.int32)
        out['plane_fail'] = np.full(n, -1, np.int32)
        out['n_planes'] = np.full(n, -1, np.int32)
        out['lp4_proj'] = np.zeros(n, np.int32)
        out['lp4_inner_fail'] = np.zeros(n, np.int32)
        out['vpref_in'] = np.zeros((n, 3), 'float64')
        out['vpref_store'] = np.zeros((n, 3), 'float64')
        out['action'] = np.zeros((n, 7), 'float64')      # f64 list as returned; env stores float32(row)
        for i, c in self.per_agent.items():
            out['ran'][i] = 1
            out['nbr_valid'][i] = c['nbr_valid']
            out['coll_query'][i] = c['coll_after']
            out['tag'][i] = c['tag']
            out['n_suit'][i] = c['n_suit']
            out['plane_fail'][i] = c['plane_fail']
            out['n_planes'][i] = c['n_planes']
            out['lp4_proj'][i] = c['lp4_proj']
            out['lp4_inner_fail'][i] = c['lp4_inner_fail']
            out['vpref_in'][i] = c['vpref_in']
            out['vpref_store'][i] = c['vpref_store']
            out['action'][i] = c['action']
            if c['nbr_valid']:
                out['nbr_cnt'][i] = len(c['nbrs'])
                for k, ((kind, oid), d) in enumerate(c['nbrs']):
                    out['nbr_ids'][i, k] = oid if kind == 'a' else n + oid
                    out['nbr_dsq'][i, k] = d
        out['done'] = np.int32(bool(done))
        self.steps.append(out)
        return bool(done)

    def static(self):
        ags = self.agents
        return dict(
            goal=np.array([a.goal_global_frame for a in ags], dtype='float64'),
            radius=np.array([a.radius for a in ags], dtype='float64'),
            pref_speed=np.array([a.pref_speed for a in ags], dtype='float64'),
            initial_pos=np.array([a.initial_pos for a in ags], dtype='float64'),
            goal_pos=np.array([a.goal_pos for a in ags], dtype='float64'),
            max_run_dist=np.array([a.max_run_dist for a in ags], dtype='float64'),
            straight_len=np.array([a.straight_path_length for a in ags], dtype='float64'),
            desire_steps=np.array([a.desire_steps for a in ags], dtype=np.int32),
            obstacles=np.array([[*o.pos_global_frame, o.radius] for o in self.obstacles],
                               dtype='float64').reshape(-1, 4),
            **self._static_paths()
        )

    def _static_paths(self):
        if self.paths is None:
            return {}
        P = max(1, max(len(p) for p in self.paths))
        wp = np.full((self.n, P, 3), np.nan)
        for i, p in enumerate(self.paths):
            for k, w in enumerate(p):
                wp[i, k] = w
        return dict(waypoints=wp, waypoint_count=np.array([len(p) for p in self.paths], dtype=np.int32))

    def metrics(self):
        """run_example/run_sca.py:228-250."""
        ok = [a for a in self.agents if not a.is_collision and not a.is_out_of_max_time]
        ns = len(ok)
        all_dist = sum(a.total_dist for a in ok)
        all_straight = sum(a.straight_path_length for a in ok)
        all_des = sum(a.desire_steps for a in ok)
        all_steps = sum(a.step_num for a in ok)
        return np.array([ns / self.n,
                         ((all_steps - all_des) * 0.1) / ns if ns else np.nan,
                         (all_dist - all_straight) / ns if ns else np.nan,
                         all_dist / all_steps / 0.1 if all_steps else np.nan], dtype='float64')

    def pack(self, keep_every=1):
        d = {}
        keys = self.steps[0].keys()
        for k in keys:
            d[k] = np.stack([s[k] for s in self.steps])
        final = self.snapshot()
        for k, v in final.items():
            d['final_' + k] = v
        d.update(self.static())
        d['metrics'] = self.metrics()
        return d
