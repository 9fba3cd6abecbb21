"""Drop-in shims for the reference's duck-typed policy / kd-tree interface.

The reference's MACAEnv (mamp/envs/mampenv.py) drives one step as

    self.kdTree.buildAgentTree()                                             # :28
    for agent in active agents:
        all_actions[i] = agent.policy.find_next_action(dict_comm, agent, self.kdTree)   # :40
    for agent: update_velocitie(agent, all_actions[i]); check_agent_state(agent)        # :42-46

`GPUKDTree` replaces `KDTree` and each `GPU*Policy` replaces the policy class of the same name, so an
UNMODIFIED MACAEnv runs the batched CUDA step:

    agents = [Agent(..., policy=GPUSCAPolicy, ...)]
    env = MACAEnv(); env.set_agents(agents, obstacles)
    env.kdTree = GPUKDTree(agents, obstacles)          # the one line a maintainer adds

`buildAgentTree()` snapshots every Agent's attributes into pinned structure-of-arrays buffers, uploads
them, runs cell list -> neighbours (REF_COMPAT: bit-exact reference lists) -> policy kernels for ALL
agents at once and downloads the float32 action matrix; `find_next_action` then just returns row
`agent.id` and back-fills `agent.neighbors`, `agent.v_pref` and `agent.is_collision` exactly where the
reference policy would have left them.  The env's own Python kinematic update / collision check still
runs afterwards, unchanged (use sca_b200.env.BatchedEnv to keep the whole step on the device).
"""
import time

import numpy as np
import torch

from . import capi
from .swarm import Swarm


class GPUKDTree(object):
    """Same four methods as mamp/policies/kdTree.py:34 KDTree."""

    def __init__(self, agents, obstacles, policy=None, nbr_mode=capi.NBR_REF_COMPAT, device='cuda:0',
                 vpref_provider=None, backfill_neighbors=True):
        """`vpref_provider`: where SCA / RVO3D+Dubins take V_des from -- 'device' (default of GPUSCAPolicy / GPURVO3dDubinsPolicy:
        the Dubins tracker of scaPolicy.py:264-338 on the GPU), a callable agent -> V_des (e.g. reference_dubins_vpref, the
        reference's own host-side tracker: 68 % of the reference's CPU time stays on the host then), or None (straight line).
        `backfill_neighbors=False` skips rebuilding agent.neighbors as Python lists (the costliest part of the shim at scale)."""
        self.backfill_neighbors = bool(backfill_neighbors)
        self.agents = agents
        self.obstacles = obstacles or []
        n = len(agents)
        pol = policy or getattr(agents[0].policy, 'scab_policy', None)
        if pol is None:
            raise capi.ScabError('GPUKDTree: agents must carry a GPU*Policy (or pass policy=...)')
        self.policy = pol
        self.nbr_mode = nbr_mode
        self.vpref_provider = vpref_provider or getattr(agents[0].policy, 'vpref_provider', None)
        self.sw = Swarm(n, device=device, debug=True)
        pos = np.array([a.pos_global_frame for a in agents], 'float64')
        goal = np.array([a.goal_global_frame for a in agents], 'float64')
        self.sw.set_state(pos, goal, heading=np.array([a.heading_global_frame for a in agents], 'float64'),
                          radius=np.array([a.radius for a in agents], 'float64'),
                          pref_speed=np.array([a.pref_speed for a in agents], 'float64'),
                          max_run_dist=np.array([a.max_run_dist for a in agents], 'float64'),
                          initial_pos=np.array([a.initial_pos for a in agents], 'float64'),
                          goal_pos=np.array([a.goal_pos for a in agents], 'float64'))
        lo = np.minimum(pos.min(0), goal.min(0)) - 20.0
        hi = np.maximum(pos.max(0), goal.max(0)) + 20.0
        self.sw.fit_grid(lo, hi)
        # agent.path of an external global planner (agent.py:44): followed on the device (scab_state_ptrs.waypoints)
        self._has_paths = any(len(getattr(a, 'path', ()) or ()) for a in agents)
        if self._has_paths:
            P = max(len(a.path) for a in agents)
            wp = np.zeros((n, P, 3))
            for i, a in enumerate(agents):
                for k, w in enumerate(a.path):
                    wp[i, k] = np.asarray(w, 'float64')[:3]
            self.sw.set_waypoints(wp, [len(a.path) for a in agents])
        if self.vpref_provider == 'device':
            # the Dubins tracker itself on the GPU (scab_dubins_enable) instead of the host-side reference tracker
            self.sw.enable_dubins(np.array([a.goal_pos for a in agents], 'float64')[:, 3:6])
            self.vpref_provider = None
        self._host = {k: torch.zeros_like(self.sw.t[k], device='cpu').pin_memory()
                      for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')}
        self._vp = torch.zeros((n, 3), dtype=torch.float64).pin_memory()
        self.action = np.zeros((n, 7), np.float32)
        self.step_ms = 0.0
        for a in agents:
            a.policy._tree = self
        if self.obstacles:
            self.buildObstacleTree()

    # -- kdTree.py:158 ------------------------------------------------------------------------------
    def buildObstacleTree(self):
        xyzr = np.array([[*o.pos_global_frame, o.radius] for o in self.obstacles], 'float64').reshape(-1, 4)
        self.sw.set_obstacles(xyzr)

    # -- kdTree.py:56 -------------------------------------------------------------------------------
    def buildAgentTree(self):
        t0 = time.time()
        ags, h = self.agents, self._host
        pos, vel, hd = h['pos'].numpy(), h['vel'].numpy(), h['heading'].numpy()
        st, td, sn = h['status'].numpy(), h['total_dist'].numpy(), h['step_num'].numpy()
        # Agent attributes -> structure of arrays: one pass per attribute (the objects are the reference's; reading them
        # is the only per-agent Python left on this side of the boundary)
        pos[:] = [a.pos_global_frame for a in ags]
        vel[:] = [a.vel_global_frame for a in ags]
        hd[:] = [a.heading_global_frame for a in ags]
        flags = np.array([(a.is_at_goal, a.is_collision, a.is_out_of_max_time) for a in ags], dtype=bool).reshape(-1, 3)
        st[:] = flags @ np.array([capi.ST_AT_GOAL, capi.ST_COLLISION, capi.ST_TIMEOUT], np.int32)
        td[:] = [a.total_dist for a in ags]
        sn[:] = [a.step_num for a in ags]
        if self.vpref_provider is not None:
            # upstream of the boundary: the reference's own stateful Dubins tracker (scaPolicy.py:264-338)
            vp = self._vp.numpy()
            for i, a in enumerate(ags):
                if not (st[i] & 7):
                    vp[i] = self.vpref_provider(a)
            self.sw.set_external_vpref(self._vp)
        self.sw.upload(h)
        self.sw.stage_grid()
        self.sw.stage_neighbors(self.nbr_mode, pos_host=pos)
        self.sw.stage_policy(self.policy)
        torch.cuda.synchronize()
        self.action = self.sw.action.cpu().numpy()
        self._nbr_cnt = self.sw.nbr_cnt.cpu().numpy()
        self._nbr_ids = self.sw.nbr_ids.cpu().numpy()
        self._nbr_dsq = self.sw.nbr_dsq.cpu().numpy()
        self._status = self.sw.status.cpu().numpy()
        self._vpref = self.sw.vpref_out.cpu().numpy()
        if self._has_paths:
            self._path_left = self.sw.t['waypoint_left'].cpu().numpy()
            self._now_goal = self.sw.t['now_goal'].cpu().numpy()
        self.step_ms = (time.time() - t0) * 1e3

    # The per-agent queries already ran inside buildAgentTree; kept for interface compatibility.
    def computeAgentNeighbors(self, agent, rangeSq):
        return None

    def computeObstacleNeighbors(self, agent, rangeSq):
        return None

    def backfill(self, agent):
        i = agent.id
        n = len(self.agents)
        c = int(self._nbr_cnt[i])
        if c >= 0 and self.backfill_neighbors:     # computeNeighbors ran for this agent (scaPolicy.py:107-116)
            agent.neighbors = [((self.agents[j] if j < n else self.obstacles[j - n]), float(d))
                               for j, d in zip(self._nbr_ids[i, :c], self._nbr_dsq[i, :c])]
        if self._status[i] & capi.ST_COLLISION:
            agent.is_collision = True          # agent.py:83-85
        if self.vpref_provider is None:
            agent.v_pref = self._vpref[i].copy()
        if self._has_paths:          # what get_trajectory leaves behind: the popped list and policy.now_goal
            del agent.path[int(self._path_left[i]):]
            if not np.isnan(self._now_goal[i, 0]):
                agent.policy.now_goal = self._now_goal[i].copy()


class _GPUPolicy(object):
    """find_next_action(dict_comm, agent, kdTree) -> [vx, vy, vz, speed, d_yaw, d_pitch, d_roll]."""
    scab_policy = None
    vpref_provider = None

    def __init__(self):
        self.now_goal = None
        self.type = "internal"
        self._tree = None

    def find_next_action(self, dict_comm, agent, kdTree):
        tree = kdTree if isinstance(kdTree, GPUKDTree) else self._tree
        if tree is None:
            raise capi.ScabError('GPU policy used without a GPUKDTree: set env.kdTree = GPUKDTree(agents, obstacles)')
        self.now_goal = agent.goal_global_frame
        tree.backfill(agent)      # (overrides now_goal while the agent still follows waypoints)
        agent.total_time += tree.step_ms * 1e-3 / max(1, len(tree.agents))
        return tree.action[agent.id]


def reference_dubins_vpref(agent):
    """The reference's own Dubins v_pref tracker, imported from the user's mamp installation (host side, per agent)."""
    from mamp.policies.sca.scaPolicy import compute_v_pref
    return compute_v_pref(agent)


_reference_dubins_vpref = reference_dubins_vpref


class GPUSCAPolicy(_GPUPolicy):            # mamp/policies/sca/scaPolicy.py:18
    scab_policy = 'sca'
    vpref_provider = 'device'              # the Dubins tracker on the GPU (k_dubins_track / k_dubins_replan)


class GPURVO3dDubinsPolicy(_GPUPolicy):    # mamp/policies/sca/rvo3dDubinsPolicy.py:18
    scab_policy = 'rvo3d_dubins'
    vpref_provider = 'device'


class GPUSCAPolicyHostDubins(GPUSCAPolicy):                 # V_des from the reference's own host-side tracker instead
    vpref_provider = staticmethod(reference_dubins_vpref)


class GPURVO3dDubinsPolicyHostDubins(GPURVO3dDubinsPolicy):
    vpref_provider = staticmethod(reference_dubins_vpref)


class GPUSRVO3DPolicy(_GPUPolicy):         # mamp/policies/srvo3dPolicy.py:15
    scab_policy = 'srvo3d'


class GPURVO3DPolicy(_GPUPolicy):          # mamp/policies/rvo3dPolicy.py:15
    scab_policy = 'rvo3d'


class GPUORCA3DPolicy(_GPUPolicy):         # mamp/policies/orca3dPolicy.py:27 (what run_orca.py imports)
    scab_policy = 'orca_sampled'


class GPUORCA3DPolicyOfficial(_GPUPolicy):  # mamp/policies/orca3dPolicyOfficial.py:27
    scab_policy = 'orca_lp'
