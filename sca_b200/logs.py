"""Rollout metrics and logs in the reference's own formats.

* `rollout_info`  -- the `info_dict_to_visualize` of run_example/run_sca.py:198-254 (written as env_cfg.json, read
  by the reference's visualization/), computed from the final state arrays instead of Agent objects.
* `TrajectoryLog` -- replaces the per-step `DataFrame.append` of Agent.to_vector (agent.py:126-148, called from
  mampenv.py:105 for EVERY agent after every kinematic update): a preallocated device buffer of
  [steps, n, 9] (pos, heading, vel) filled by device-to-device copies every `every` steps; `columns()` returns the
  reference's 13 ANIMATION_COLUMNS (agent.py:75-76).

Host-side bookkeeping only; PyTorch is used for the device buffer and the copies.
"""
import json
import os

import numpy as np

ANIMATION_COLUMNS = ['pos_x', 'pos_y', 'pos_z', 'alpha', 'beta', 'gamma', 'vel_x', 'vel_y', 'vel_z',
                     'gol_x', 'gol_y', 'gol_z', 'radius']           # agent.py:75-76
ST_COLLISION, ST_TIMEOUT = 2, 4


def rollout_info(status, total_dist, step_num, straight_len, desire_steps, radius, goal, obstacles=None,
                 compute_time=0.0, dt=0.1, group=None):
    """run_sca.py:198-254.  `compute_time` is the summed policy time of the successful agents in seconds
    (agent.total_time); a batched step has no per-agent time, so callers pass step time x successful / n."""
    status, total_dist, step_num = np.asarray(status), np.asarray(total_dist, 'float64'), np.asarray(step_num)
    n = len(status)
    ok = (status & (ST_COLLISION | ST_TIMEOUT)) == 0
    ns = int(ok.sum())
    all_dist = float(total_dist[ok].sum())
    all_straight = float(np.asarray(straight_len, 'float64')[ok].sum())
    all_des = int(np.asarray(desire_steps)[ok].sum())
    all_steps = int(step_num[ok].sum())
    radius = np.broadcast_to(np.asarray(radius, 'float64'), (n,))
    goal = np.asarray(goal, 'float64')
    info = {
        'all_agent_info': [{'id': i, 'gp': int(group[i]) if group is not None else 0, 'radius': float(radius[i]),
                            'goal_pos': goal[i, :3].tolist()} for i in range(n)],
        'all_obstacle': [{'position': [float(o[0]), float(o[1]), float(o[2])], 'shape': 'sphere', 'feature': float(o[3])}
                         for o in (np.zeros((0, 4)) if obstacles is None else np.asarray(obstacles, 'float64').reshape(-1, 4))],
        'all_compute_time': float(compute_time),
        'all_straight_distance': all_straight,
        'all_distance': all_dist,
        'successful_num': ns,
        'all_desire_step_num': all_des,
        'all_step_num': all_steps,
        'SuccessRate': ns / n if n else 0.0,
        'ExtraTime': ((all_steps - all_des) * dt) / ns if ns else float('nan'),
        'ExtraDistance': (all_dist - all_straight) / ns if ns else float('nan'),
        'AverageSpeed': all_dist / all_steps / dt if all_steps else float('nan'),
        'AverageCost': 1000 * float(compute_time) / all_steps if all_steps else float('nan'),
    }
    return info


def write_env_cfg(path, info):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as f:
        f.write(json.dumps(info, indent=4))


class TrajectoryLog(object):
    """Strided trajectory buffer on the device: row t holds (pos, heading, vel) of every agent after step
    (t + 1) * every."""

    def __init__(self, swarm, capacity, every=1):
        import torch
        self.sw, self.every, self.capacity = swarm, max(1, int(every)), int(capacity)
        self.buf = torch.zeros((self.capacity, swarm.n, 9), dtype=torch.float64, device=swarm.device)
        self.rows = 0
        self.dropped = 0

    def record(self, step_index):
        """Call after step number `step_index` (1-based); asynchronous on the current stream."""
        if step_index % self.every:
            return
        if self.rows >= self.capacity:
            self.dropped += 1
            return
        n, t = self.sw.n, self.sw.t
        row = self.buf[self.rows]
        row[:, 0:3].copy_(t['pos'][:n])
        row[:, 3:6].copy_(t['heading'][:n])
        row[:, 6:9].copy_(t['vel'][:n])
        self.rows += 1

    def columns(self):
        """[rows, n, 13] in the reference's ANIMATION_COLUMNS order."""
        n = self.sw.n
        a = self.buf[:self.rows].cpu().numpy()
        out = np.empty((self.rows, n, 13))
        out[:, :, 0:9] = a
        out[:, :, 9:12] = self.sw.t['goal'][:n].cpu().numpy()[None]
        out[:, :, 12] = self.sw.t['radius'][:n].cpu().numpy()[None]
        return out

    def save(self, log_dir):
        """trajs.npz (one array, all agents) and -- like the reference's one-sheet-per-agent trajs.xlsx -- one CSV per
        agent under trajs/ with the ANIMATION_COLUMNS header (openpyxl is not a dependency)."""
        os.makedirs(os.path.join(log_dir, 'trajs'), exist_ok=True)
        cols = self.columns()
        np.savez_compressed(os.path.join(log_dir, 'trajs.npz'), data=cols, columns=np.array(ANIMATION_COLUMNS), every=self.every)
        for i in range(cols.shape[1]):
            np.savetxt(os.path.join(log_dir, 'trajs', 'agent%d.csv' % i), cols[:, i], delimiter=',',
                       header=','.join(ANIMATION_COLUMNS), comments='')
        return cols.shape
