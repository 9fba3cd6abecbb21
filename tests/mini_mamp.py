"""A minimal stand-in for the reference's Agent / Obstacle objects (mamp/agents/agent.py:8-77,
obstacle.py:9-33): just the attributes the policy interface reads and writes.  The real `mamp` package
lives in /root/reference, which does not exist on the GPU box; the shims are duck-typed, so this is all
they need.  Test-only."""
import numpy as np


class Agent(object):
    def __init__(self, start_pos, goal_pos, vel, radius, pref_speed, policy, id, dt=0.1):
        self.policy = policy()
        self.initial_pos = np.array(start_pos, dtype='float64')
        self.goal_pos = np.array(goal_pos, dtype='float64')
        self.pos_global_frame = np.array(start_pos[:3], dtype='float64')
        self.goal_global_frame = np.array(goal_pos[:3], dtype='float64')
        self.heading_global_frame = np.array(start_pos[3:], dtype='float64')
        self.vel_global_frame = np.array(vel)
        self.radius = radius
        self.id = id
        self.pref_speed = pref_speed
        self.neighbors = []
        self.v_pref = np.zeros(3)
        self.total_time = 0.0
        self.total_dist = 0.0
        self.step_num = 0
        self.is_at_goal = False
        self.is_collision = False
        self.is_out_of_max_time = False
        d = self.initial_pos[:3] - self.goal_pos[:3]
        self.max_run_dist = 3.0 * round(float(np.sqrt((d ** 2).sum())), 5)
        self.is_obstacle = False


class Obstacle(object):
    def __init__(self, pos, radius, id):
        self.pos_global_frame = np.array(pos, dtype='float64')
        self.vel_global_frame = np.zeros(3)
        self.radius = radius
        self.id = id
        self.is_at_goal = True
        self.is_obstacle = True
