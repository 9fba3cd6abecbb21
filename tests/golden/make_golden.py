"""Generate the golden fixtures under tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only) through ref_harness.Recorder.

    python tests/golden/make_golden.py [case ...]

Each fixture holds, for the first `keep` steps of a reference rollout: the pre-step state of every
agent, the reference's outputs for that step (ordered neighbour lists, branch tags, plane counts,
planeFail, v_pref, the action row) and -- when `full` is set -- the rollout-level metrics of
run_example/run_sca.py:246-250 after running the reference to completion.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def dense(n, r, seed):
    pos, goal, _ = rh.scen_random(n, r=r, zc=30.0, seed=seed)
    return pos, goal, []


def circle_with_waypoints(n=16, rad=10.0):
    """the circle swap with a two-waypoint dog-leg per agent (agent.path as an external global planner
    would fill it; the policies pop() from the end of the list, so the first waypoint is stored last)."""
    pos, goal, obs = rh.scen_circle(n, rad)
    paths = []
    for j in range(n):
        s, g = np.array(pos[j][:3], 'float64'), np.array(goal[j][:3], 'float64')
        d = g - s
        lat = np.array([-d[1], d[0], 0.0]) / max(np.hypot(d[0], d[1]), 1e-9)
        wp1 = s + 0.35 * d + np.array([0.0, 0.0, 3.0 if j % 2 == 0 else -3.0])
        wp2 = s + 0.70 * d + 2.0 * lat
        paths.append([wp2.tolist(), wp1.tolist()] if j % 5 else [wp1.tolist()])     # some agents carry one waypoint only
    return pos, goal, obs, paths


CASES = {
    # name: (policy, scenario thunk, keep steps, run to completion?, max steps)
    'circle16_sca': ('sca', lambda: rh.scen_circle(16, 10.0), 130, False, 130),
    'circle16_rvo3d_dubins': ('rvo3d_dubins', lambda: rh.scen_circle(16, 10.0), 110, False, 110),
    'circle16_rvo3d': ('rvo3d', lambda: rh.scen_circle(16, 10.0), 120, True, 1500),
    'circle16_srvo3d': ('srvo3d', lambda: rh.scen_circle(16, 10.0), 120, True, 1500),
    'circle16_orca_sampled': ('orca_sampled', lambda: rh.scen_circle(16, 10.0), 120, False, 120),
    'circle16_orca_lp': ('orca_lp', lambda: rh.scen_circle(16, 10.0), 150, True, 1500),
    'circle100_sca': ('sca', lambda: rh.scen_circle(100, 18.0), 6, False, 6),
    'circle100_orca_lp': ('orca_lp', lambda: rh.scen_circle(100, 18.0), 40, True, 2000),
    'sphere100_rvo3d': ('rvo3d', lambda: rh.scen_sphere(100), 5, False, 5),
    'sphere100_srvo3d': ('srvo3d', lambda: rh.scen_sphere(100), 5, False, 5),
    'random1000_orca_lp': ('orca_lp', lambda: rh.scen_random(1000, seed=0), 6, False, 6),
    'random150_orca_sampled': ('orca_sampled', lambda: rh.scen_random(150, r=8.0, seed=1), 5, False, 5),
    'takeoff16_sca': ('sca', lambda: rh.scen_takeoff_landing(16), 80, False, 80),
    'takeoff16_srvo3d': ('srvo3d', lambda: rh.scen_takeoff_landing(16), 60, False, 60),
    'dense40_sca': ('sca', lambda: dense(40, 4.0, 2), 12, False, 12),
    'dense40_srvo3d': ('srvo3d', lambda: dense(40, 4.0, 3), 12, False, 12),
    'dense40_rvo3d': ('rvo3d', lambda: dense(40, 4.0, 4), 12, False, 12),
    'dense40_orca_sampled': ('orca_sampled', lambda: dense(40, 4.0, 5), 12, False, 12),
    'dense60_orca_lp': ('orca_lp', lambda: dense(60, 3.5, 6), 30, False, 30),
    'waypoints16_rvo3d': ('rvo3d', circle_with_waypoints, 160, False, 160),
    'waypoints16_srvo3d': ('srvo3d', circle_with_waypoints, 160, False, 160),
    'waypoints16_orca_sampled': ('orca_sampled', circle_with_waypoints, 160, False, 160),
    'waypoints16_orca_lp': ('orca_lp', circle_with_waypoints, 200, False, 200),
}


# Reference rollouts run to COMPLETION, of which only the first two steps, the end state and the metrics are kept
# (rollout_<name>.npz; excluded from the per-step fixture list): SCA with its Dubins tracker on the README scenarios.
ROLLOUTS = {
    'circle16_sca': ('sca', lambda: rh.scen_circle(16, 10.0), 3000),          # run_sca.py as shipped (R = 10, N = 16)
    'takeoff16_sca': ('sca', lambda: rh.scen_takeoff_landing(16), 3000),      # run_sca.py:53-81 + the 8 exp2 obstacles
    'circle100_sca': ('sca', lambda: rh.scen_circle(100, 18.0), 3000),        # BASELINE configs[0] (README, R = 18, N = 100)
}


def record(policy, sc, keep, full=False, max_steps=None, radius=0.5):
    """Run the reference on scenario `sc` = (pos, goal, obstacles[, paths]) and return the fixture dict."""
    max_steps = max_steps or keep
    pos, goal, obs = sc[:3]
    rec = rh.Recorder(policy, pos, goal, obs, radius=radius, paths=sc[3] if len(sc) > 3 else None)
    steps = 0
    done = False
    while steps < max_steps and not done:
        done = rec.step()
        steps += 1
        if steps == keep and not full:
            break
    all_steps = rec.steps
    rec.steps = all_steps[:keep]
    d = rec.pack()
    # pack() snapshots the state after the LAST executed step; keep the state right after `keep` too
    if len(all_steps) > keep:
        for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num'):
            d['after_' + k] = all_steps[keep][k]
    else:
        for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num'):
            d['after_' + k] = d['final_' + k]
    d['policy'] = np.array(policy)
    d['rollout_steps'] = np.int32(steps)
    d['rollout_done'] = np.int32(done)
    d['full'] = np.int32(full)
    return d, steps, done


def run_case(name):
    policy, scen, keep, full, max_steps = CASES[name]
    t0 = time.time()
    d, steps, done = record(policy, scen(), keep, full, max_steps)
    out = os.path.join(HERE, name + '.npz')
    np.savez_compressed(out, **d)
    print('%-26s steps=%d done=%d  %.1fs  %.0f kB  tags=%s' % (
        name, steps, done, time.time() - t0, os.path.getsize(out) / 1e3,
        np.unique(d['tag']).tolist()), flush=True)


def run_rollout(name):
    policy, scen, max_steps = ROLLOUTS[name]
    t0 = time.time()
    d, steps, done = record(policy, scen(), 2, True, max_steps)
    out = os.path.join(HERE, 'rollout_' + name + '.npz')
    np.savez_compressed(out, **d)
    print('rollout_%-18s steps=%d done=%d  %.1fs  metrics=%s' % (name, steps, done, time.time() - t0, d['metrics']), flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        if nm.startswith('rollout_'):
            run_rollout(nm[len('rollout_'):])
        else:
            run_case(nm)
