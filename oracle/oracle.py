"""ctypes binding of the parity oracle (oracle/sca_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under sca_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libsca_oracle.so')

POLICIES = {'sca': 0, 'srvo3d': 1, 'rvo3d': 2, 'rvo3d_dubins': 3, 'orca_sampled': 4, 'orca_lp': 5}
NBR_KNN, NBR_REF_COMPAT = 0, 1
K = 16


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ('sca_oracle.c', 'sca_dubins.c', 'Makefile')]
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(['make', '-C', HERE, '-B' if force else '-s'], stdout=subprocess.DEVNULL)
    return LIB


class KdNode(C.Structure):
    _fields_ = [('begin', C.c_int32), ('end', C.c_int32), ('left', C.c_int32), ('right', C.c_int32),
                ('mn', C.c_double * 3), ('mx', C.c_double * 3)]


_P = C.c_void_p


class OrcWorld(C.Structure):
    _fields_ = [('n', C.c_int32), ('m', C.c_int32)] + [(k, _P) for k in (
        'pos', 'vel', 'heading', 'goal', 'radius', 'pref_speed', 'status', 'total_dist', 'step_num',
        'max_run_dist', 'is_zaxis', 'obs', 'vpref_ext', 'agent_perm', 'obst_perm',
        'ran', 'nbr_cnt', 'nbr_ids', 'nbr_dsq', 'coll_query', 'tag', 'n_suit', 'plane_fail', 'n_planes',
        'lp4_proj', 'lp4_inner_fail', 'vpref_in', 'vpref_store', 'action64', 'action',
        'agent_nodes', 'obst_nodes', 'waypoints', 'waypoint_left', 'now_goal')] + [('waypoint_stride', C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_py_round5.restype = C.c_double
        _lib.orc_py_round5.argtypes = [C.c_double]
        _lib.orc_py_round5_slow.restype = C.c_double
        _lib.orc_py_round5_slow.argtypes = [C.c_double]
        _lib.orc_gen_candidates.argtypes = [C.c_int, C.c_double, _P]
        _lib.orc_set_threads.argtypes = [C.c_int]
        assert _lib.orc_sizeof_world() == C.sizeof(OrcWorld)
        assert _lib.orc_sizeof_kdnode() == C.sizeof(KdNode)
    return _lib


class World:
    """Host-side SoA world the oracle steps in place (mirrors the state the product keeps in HBM)."""

    def __init__(self, pos, goal, heading=None, radius=0.5, pref_speed=1.0, obstacles=None, vel=None,
                 initial_pos=None, goal_pos=None):
        n = len(pos)
        self.n = n
        f8 = lambda a, shape: np.ascontiguousarray(np.broadcast_to(np.asarray(a, 'float64'), shape)).copy()
        self.pos = f8(pos, (n, 3))
        self.goal = f8(goal, (n, 3))
        self.heading = f8(np.zeros((n, 3)) if heading is None else heading, (n, 3))
        self.vel = np.zeros((n, 3), np.float32) if vel is None else np.ascontiguousarray(vel, np.float32).copy()
        self.radius = f8(radius, (n,))
        self.pref_speed = f8(pref_speed, (n,))
        self.status = np.zeros(n, np.int32)
        self.total_dist = np.zeros(n, 'float64')
        self.step_num = np.zeros(n, np.int32)
        start = self.pos if initial_pos is None else np.asarray(initial_pos, 'float64')[:, :3]
        gl = self.goal if goal_pos is None else np.asarray(goal_pos, 'float64')[:, :3]
        # agent.py:74 max_run_dist = 3 * l3norm(start, goal)  (python round of sqrt)
        d = np.sqrt(((start - gl) ** 2).sum(1))
        self.max_run_dist = 3.0 * np.array([round(float(x), 5) for x in d])
        dd = gl - start
        self.is_zaxis = ((np.abs(dd[:, 0]) <= 1e-5) & (np.abs(dd[:, 1]) <= 1e-5)).astype(np.uint8)
        obs = np.zeros((0, 4)) if obstacles is None or len(obstacles) == 0 else np.asarray(obstacles, 'float64')
        self.obs = np.ascontiguousarray(obs.reshape(-1, 4))
        self.m = len(self.obs)
        self.vpref_ext = None
        self.agent_perm = np.arange(n, dtype=np.int32)
        self.obst_perm = np.arange(self.m, dtype=np.int32)
        self.agent_nodes = (KdNode * max(1, 2 * n - 1))()
        self.obst_nodes = (KdNode * max(1, 2 * self.m - 1))()
        i4 = lambda *s: np.zeros(s, np.int32)
        self.ran, self.nbr_cnt, self.nbr_ids = i4(n), i4(n), i4(n, K)
        self.nbr_dsq = np.zeros((n, K))
        self.coll_query, self.tag, self.n_suit = i4(n), i4(n), i4(n)
        self.plane_fail, self.n_planes, self.lp4_proj, self.lp4_inner_fail = i4(n), i4(n), i4(n), i4(n)
        self.vpref_in, self.vpref_store = np.zeros((n, 3)), np.zeros((n, 3))
        self.action64 = np.zeros((n, 7))
        self.action = np.zeros((n, 7), np.float32)
        self._obst_built = False
        self.waypoints = self.waypoint_left = self.now_goal = None
        self.waypoint_stride = 0

    def set_waypoints(self, waypoints, counts):
        """agent.path for every agent: waypoints [n, P, 3] in the reference's list order (pop() takes the last live
        entry), counts [n] = len(agent.path); resets policy.now_goal to None."""
        wp = np.ascontiguousarray(np.nan_to_num(np.asarray(waypoints, 'float64')))
        assert wp.shape[0] == self.n and wp.shape[2] == 3
        self.waypoints = wp
        self.waypoint_stride = int(wp.shape[1])
        self.waypoint_left = np.ascontiguousarray(counts, np.int32).copy()
        self.now_goal = np.full((self.n, 3), np.nan)

    def _struct(self):
        w = OrcWorld()
        w.n, w.m = self.n, self.m
        w.waypoint_stride = self.waypoint_stride
        for k, _ in OrcWorld._fields_[2:-1]:
            v = getattr(self, k)
            if v is None:
                setattr(w, k, None)
            elif isinstance(v, np.ndarray):
                assert v.flags['C_CONTIGUOUS']
                setattr(w, k, v.ctypes.data)
            else:
                setattr(w, k, C.cast(v, _P))
        return w

    def set_vpref_ext(self, v):
        self.vpref_ext = None if v is None else np.ascontiguousarray(v, 'float64').copy()

    def _ensure_obst(self, w):
        if not self._obst_built:
            lib().orc_kdtree_build_obstacles(C.byref(w))   # mampenv.py:20, once
            self._obst_built = True

    def kdtree_build(self):
        w = self._struct()
        lib().orc_kdtree_build_agents(C.byref(w))

    def policy_step(self, policy, nbr_mode=NBR_REF_COMPAT):
        w = self._struct()
        self._ensure_obst(w)
        lib().orc_policy_step(C.byref(w), POLICIES[policy], nbr_mode)

    def env_update(self, fast=False):
        w = self._struct()
        f = lib().orc_env_update_fast if fast else lib().orc_env_update
        return bool(f(C.byref(w)))

    def step(self, policy, nbr_mode=NBR_REF_COMPAT, fast=False):
        self.policy_step(policy, nbr_mode)
        return self.env_update(fast)

    def step_dubins(self, policy, prov, goal_heading, nbr_mode=NBR_REF_COMPAT):
        """One step of 'sca' / 'rvo3d_dubins' with V_des from the stateful Dubins tracker (scaPolicy.py:264-338)."""
        v = np.zeros((self.n, 3))
        act = np.nonzero((self.status & 7) == 0)[0]
        for i in act:
            v[i] = prov.vpref(i, self.pos[i], self.vel[i], self.heading[i], self.goal[i], goal_heading[i],
                              float(self.pref_speed[i]), bool(self.is_zaxis[i]))
        self.set_vpref_ext(v)
        self.policy_step(policy, nbr_mode)
        for i in act:
            if self.nbr_cnt[i] >= 0:      # computeNeighbors ran (not the zero-velocity shortcut)
                c = int(self.nbr_cnt[i])
                prov.set_last_neighbor(i, c > 0, float(self.nbr_dsq[i][0]) if c > 0 else 0.0)
        return self.env_update()


class Man2D(C.Structure):
    _fields_ = [('qi', C.c_double * 3), ('r', C.c_double), ('t', C.c_double), ('p', C.c_double), ('q', C.c_double),
                ('mode', C.c_int32 * 3), ('ok', C.c_int32), ('length', C.c_double)]


class DubinsState(C.Structure):
    _fields_ = [('in_use', C.c_int32), ('n_points', C.c_int32), ('cursor', C.c_int32), ('ss', C.c_double),
                ('length', C.c_double), ('qi0', C.c_double * 3), ('h', Man2D), ('v', Man2D), ('now_goal', C.c_double * 3),
                ('v_pref', C.c_double * 3), ('nbr0_dsq', C.c_double), ('has_nbr', C.c_int32), ('replans', C.c_int32)]


def dubins3d(qi, qf, rmin, pitchlims):
    """mamp/policies/sca/dubinsmaneuver3d.py:34 dubinsmaneuver3d -> DubinsState (length, n_points, manoeuvres)."""
    L = lib()
    assert L.orc_sizeof_dubins_state() == C.sizeof(DubinsState)
    S = DubinsState()
    a = lambda v: (C.c_double * len(v))(*[float(x) for x in v])
    L.orc_dubins3d(a(qi), a(qf), C.c_double(rmin), a(pitchlims), C.byref(S))
    return S


def dubins_point(S, i):
    out = (C.c_double * 3)()
    lib().orc_dubins_point(C.byref(S), int(i), out)
    return np.array(out[:])


class DubinsProvider:
    """Per-agent stateful Dubins v_pref tracker (scaPolicy.py:264-338), one DubinsState per agent."""

    def __init__(self, n):
        assert lib().orc_sizeof_dubins_state() == C.sizeof(DubinsState)
        self.S = (DubinsState * n)()
        self.n = n

    def set_last_neighbor(self, i, has, dsq):
        self.S[i].has_nbr = int(bool(has))
        self.S[i].nbr0_dsq = float(dsq)

    def vpref(self, i, pos, vel, heading, goal, goal_heading, pref_speed, is_zaxis):
        out = (C.c_double * 3)()
        d = lambda v: (C.c_double * 3)(*[float(x) for x in v])
        f = (C.c_float * 3)(*[float(np.float32(x)) for x in vel])
        lib().orc_dubins_vpref(C.byref(self.S[i]), d(pos), f, d(heading), d(goal), d(goal_heading), C.c_double(pref_speed),
                               int(is_zaxis), out)
        return np.array(out[:])


def set_threads(t):
    """Threads the oracle's policy loop may use (the reference itself is single-threaded)."""
    lib().orc_set_threads(int(t))


def gen_candidates(num_n=256, pref_speed=1.0):
    out = np.zeros((2 * num_n, 3))
    c = lib().orc_gen_candidates(num_n, pref_speed, out.ctypes.data)
    return out[:c]
