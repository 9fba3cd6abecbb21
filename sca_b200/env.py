"""BatchedEnv: MACAEnv (mamp/envs/mampenv.py:8-105) with the whole step resident on the device.

    env = BatchedEnv('srvo3d'); env.set_agents(pos6, goal6, obstacles)
    while not env.step(): pass
    env.metrics()

`set_agents` takes the same [x, y, z, yaw, pitch, roll] rows Agent(start_pos=..., goal_pos=...) takes
(mamp/agents/agent.py:9-20); `step()` returns what MACAEnv.step returns (all agents done).
"""
import numpy as np

from . import capi
from .swarm import Swarm


class BatchedEnv(object):
    def __init__(self, policy, nbr_mode=capi.NBR_KNN, device='cuda:0', debug=False, dubins=None, log_every=0,
                 log_capacity=4096):
        """`dubins` (default: True for 'sca' and 'rvo3d_dubins', the two policies whose reference v_pref is the
        Dubins tracker of scaPolicy.py:264-338): run that tracker on the device; False = straight-line v_pref."""
        self.policy = policy
        self.dubins = (policy in ('sca', 'rvo3d_dubins')) if dubins is None else bool(dubins)
        self.nbr_mode = nbr_mode
        self.log_every, self.log_capacity, self.traj = int(log_every), int(log_capacity), None   # trajectory log (logs.py)
        self.device = device
        self.debug = debug
        self.sw = None
        self.steps = 0

    def set_agents(self, pos6, goal6, obstacles=None, radius=0.5, pref_speed=1.0, vel=None, margin=20.0, paths=None):
        """`paths`: optional per-agent waypoint lists (agent.path, popped from the end like the reference does)."""
        pos6, goal6 = np.asarray(pos6, 'float64'), np.asarray(goal6, 'float64')
        n = len(pos6)
        self.sw = Swarm(n, device=self.device, debug=self.debug)
        self.sw.set_state(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], vel=vel, radius=radius, pref_speed=pref_speed,
                          initial_pos=pos6, goal_pos=goal6)
        self.sw.set_obstacles(obstacles)
        lo = np.minimum(pos6[:, :3].min(0), goal6[:, :3].min(0)) - margin
        hi = np.maximum(pos6[:, :3].max(0), goal6[:, :3].max(0)) + margin
        self.sw.fit_grid(lo, hi)
        if self.dubins:
            self.sw.enable_dubins(goal6[:, 3:6])
        if paths is not None:
            P = max(1, max(len(p) for p in paths))
            wp = np.zeros((n, P, 3))
            for i, p in enumerate(paths):
                for k, w in enumerate(p):
                    wp[i, k] = w[:3]
            self.sw.set_waypoints(wp, [len(p) for p in paths])
        # agent.py:51-52 (metrics): straight_path_length = l3norm(start, goal) - 0.5, desire_steps
        d = np.round(np.sqrt(((pos6[:, :3] - goal6[:, :3]) ** 2).sum(1)), 5)
        self.straight_len = d - 0.5
        self.desire_steps = (self.straight_len / (np.broadcast_to(pref_speed, (n,)) * 0.1)).astype(np.int64)
        self.steps = 0
        self.obstacles = np.zeros((0, 4)) if obstacles is None else np.asarray(obstacles, 'float64').reshape(-1, 4)
        self.goal6, self.radius = goal6, radius
        if self.log_every > 0:
            from .logs import TrajectoryLog
            self.traj = TrajectoryLog(self.sw, self.log_capacity, self.log_every)
        return self

    def step(self, sync=True):
        """One MACAEnv.step.  With sync=False nothing is read back (returns None): call is_done() when needed."""
        self.sw.step(self.policy, self.nbr_mode)
        self.steps += 1
        if self.traj is not None:
            self.traj.record(self.steps)
        return self.sw.is_done() if sync else None

    def is_done(self):
        return self.sw.is_done()

    def run(self, max_steps=100000, check_every=16):
        while self.steps < max_steps:
            self.step(sync=False)
            if self.steps % check_every == 0 and self.is_done():
                break
        return self.steps

    def metrics(self):
        return self.sw.metrics(self.straight_len, self.desire_steps)

    def write_log(self, log_dir, compute_time=0.0):
        """env_cfg.json in the reference's schema (run_sca.py:198-259) + the trajectory log, if one was kept."""
        import os
        from . import logs
        d = self.sw.download(('status', 'total_dist', 'step_num'))
        info = logs.rollout_info(d['status'], d['total_dist'], d['step_num'], self.straight_len, self.desire_steps, self.radius,
                                 self.goal6, self.obstacles, compute_time)
        logs.write_env_cfg(os.path.join(log_dir, 'env_cfg.json'), info)
        if self.traj is not None:
            self.traj.save(log_dir)
        return info
