"""S logical shards on ONE device (sca_b200.sharding.LogicalShards: the real halo / migration kernels, copies instead of
NCCL) against the single-device Swarm of the same kernels: bit for bit, migrations included.  Used by the -m gpu test
(real sm_100a kernels on a 1-GPU box) and, on the host emulation build, by the CPU suite."""
import numpy as np

from sca_b200 import capi, scenarios
from sca_b200.sharding import LogicalShards
from sca_b200.swarm import Swarm


def run(policy, shards, n, steps, device, dubins=False, waypoints=False, migrate_every=4, mig_cap=4096, density_x=4.0, transport='local'):
    density = density_x * 1.0e6 / 560.0 ** 3          # denser than config 5: more interaction across the faces
    pos, goal, _ = scenarios.uniform_box_numpy(n, density, seed=3)
    L = (n / density) ** (1 / 3)
    lo, hi = np.array([-20.0, -20.0, 0.0]), np.array([L + 20.0, L + 20.0, L + 40.0])
    ls = LogicalShards(pos, goal, shards, device, lo, hi, migrate_every=migrate_every, mig_cap=mig_cap, transport=transport)
    wp = wp_cnt = None
    if waypoints:
        rng = np.random.default_rng(5)
        s0, g0 = pos[:, :3], goal[:, :3]
        d0 = (g0 - s0) / np.linalg.norm(g0 - s0, axis=1, keepdims=True)
        wp = np.stack([s0 + 0.8 * (g0 - s0) + rng.normal(0, 1.0, (n, 3)), s0 + 0.5 * (g0 - s0) + rng.normal(0, 1.0, (n, 3)),
                       s0 + 0.8 * d0 + rng.normal(0, 0.05, (n, 3))], 1)
        wp_cnt = (3 - (np.arange(n) % 2)).astype(np.int32)
        ls.set_waypoints(wp, wp_cnt)
    if dubins:
        ls.enable_dubins(goal[:, 3:6], policy)
    own0 = [set(sh._mine.tolist()) for sh in ls.shards]
    for _ in range(steps):
        ls.step(policy, capi.NBR_KNN)
    ls.check_overflow()
    keys = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num') + (('waypoint_left', 'now_goal') if waypoints else ())
    moved = 0
    for sh, o in zip(ls.shards, own0):
        g, _ = sh.gather_state(('status',))
        moved += len(set(g.tolist()) - o)
    got = ls.gather_state(keys)
    cnt = ls.counters()

    ref = Swarm(n, device=device)
    ref.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:6], initial_pos=pos, goal_pos=goal)
    ref.fit_grid(lo, hi)
    if waypoints:
        ref.set_waypoints(wp, wp_cnt)
    if dubins:
        ref.enable_dubins(goal[:, 3:6])
    for _ in range(steps):
        ref.step(policy, capi.NBR_KNN)
    import torch
    torch.cuda.synchronize()
    bad = {}
    for k in keys:
        a, b = got[k], ref.t[k][:n].cpu().numpy()
        if k == 'status':
            a, b = a & 15, b & 15
        if k == 'now_goal':
            a, b = np.nan_to_num(a, nan=-1e9), np.nan_to_num(b, nan=-1e9)
        bad[k] = int((a != b).reshape(n, -1).any(1).sum())
    refc = ref.counters()
    cbad = {k: (cnt[k], refc[k]) for k in ('n_active_total', 'n_active', 'n_not_done', 'n_collision') if cnt[k] != refc[k]}
    info = {'mismatches': bad, 'counter_mismatches': cbad, 'migrations': moved,
            'collisions': int((ref.t['status'][:n].cpu().numpy() & 2).astype(bool).sum())}
    if waypoints:
        info['popped'] = int((ref.t['waypoint_left'][:n].cpu().numpy() < wp_cnt).sum())
    if dubins:
        info['replans'] = int(ref.dubins_states()['replans'].sum())
    ls.close()
    return info
