"""Scenario generators: the reference's run_example/ builders with explicit seeds, plus the scaled
synthetic swarms of BASELINE.json (configs 3-5).  Each returns (pos[n,6], goal[n,6], obstacles[m,4]);
columns 3-5 are yaw, pitch, roll as in Agent(start_pos=[x, y, z, yaw, pitch, roll]) (agent.py:13-20)."""
import math
import random

import numpy as np


def _mod2pi(theta):   # mamp/util.py:112-113
    return theta - 2.0 * math.pi * math.floor(theta / 2.0 / math.pi)


def circle(n=100, rad=18.0):
    """run_example/run_sca.py:17-30 / run_orca.py:16-33 (README's 100-UAV comparison uses rad=18)."""
    pos = np.zeros((n, 6))
    for j in range(n):
        pos[j] = [round(rad * np.cos(2 * j * np.pi / n), 2), round(rad * np.sin(2 * j * np.pi / n), 2), 10.0,
                  round(_mod2pi(2 * j * np.pi / n + np.pi), 5), 0, 0]
    goal = pos[(np.arange(n) + n // 2) % n].copy()
    return pos, goal, np.zeros((0, 4))


def sphere(n=100, rad=25.0, zc=30.0):
    """run_example/run_orca.py:36-54: Fibonacci sphere, goal = antipode."""
    phi = (math.sqrt(5.0) - 1.0) / 2.0
    pos = np.zeros((n, 6))
    for k in range(1, n + 1):
        z = (2 * k - 1) / n - 1
        pos[k - 1, 0] = rad * math.sqrt(1 - z ** 2) * math.cos(2 * math.pi * k * phi)
        pos[k - 1, 1] = rad * math.sqrt(1 - z ** 2) * math.sin(2 * math.pi * k * phi)
        pos[k - 1, 2] = rad * z
    goal = -pos
    pos[:, 2] += zc
    goal[:, 2] += zc
    return pos, goal, np.zeros((0, 4))


def dense_sphere(n=100000, spacing=0.25):
    """BASELINE config 4: sphere swap scaled so that agents stay >= 1.1 m apart:
    R = 25 sqrt(n/100) * spacing."""
    rad = 25.0 * math.sqrt(n / 100.0) * spacing
    return sphere(n, rad=rad, zc=rad + 30.0)


def random_box(n=1000, r=15.0, zc=30.0, seed=0):
    """run_example/run_orca.py:57-73 with random.seed(seed) (the reference is unseeded)."""
    random.seed(seed)
    pos = np.zeros((n, 6))
    goal = np.zeros((n, 6))
    for i in range(n):
        pos[i, :3] = [random.uniform(-r, r), random.uniform(-r, r), random.uniform(-r, r)]
    for i in range(n):
        goal[i, :3] = [random.uniform(-r, r), random.uniform(-r, r), random.uniform(-r, r)]
    pos[:, 2] += zc
    goal[:, 2] += zc
    return pos, goal, np.zeros((0, 4))


def takeoff_landing(n=16):
    """run_example/run_sca.py:53-81 + the 8 exp2 sphere obstacles of :139-150."""
    rad = 4.0
    landing, takeoff = n - n // 2, n // 2
    pos = np.zeros((n, 6))
    for j in range(landing):
        pos[j] = [round(rad * np.cos(2 * j * np.pi / landing), 2), round(rad * np.sin(2 * j * np.pi / landing), 2),
                  10.0, round(np.pi / 2, 5), 0, 0]
    for j in range(landing, n):
        pos[j] = [round(rad * np.cos(2 * j * np.pi / takeoff), 2), round(rad * np.sin(2 * j * np.pi / takeoff), 2),
                  0.0, round(-np.pi / 2, 5), 0, 0]
    goal = np.concatenate([pos[landing:landing + landing], pos[:n - landing]])
    obs = np.array([[round(rad * np.cos(2 * j * np.pi / 8), 2), round(rad * np.sin(2 * j * np.pi / 8), 2), 5.0, 1.0]
                    for j in range(8)])
    return pos, goal, obs


def takeoff_landing_tiled(pads=625, pitch=30.0, n_pad=16):
    """BASELINE config 3: `pads` take-off/landing pads on a square lattice (pitch 30 m; D = 10 m so pads
    interact only weakly).  625 pads x 16 UAVs = 10^4 agents, 5000 obstacles."""
    side = int(math.ceil(math.sqrt(pads)))
    p0, g0, o0 = takeoff_landing(n_pad)
    P, G, O = [], [], []
    for k in range(pads):
        off = np.array([(k % side) * pitch, (k // side) * pitch, 0.0])
        p, g, o = p0.copy(), g0.copy(), o0.copy()
        p[:, :3] += off
        g[:, :3] += off
        o[:, :3] += off
        P.append(p); G.append(g); O.append(o)
    return np.concatenate(P), np.concatenate(G), np.concatenate(O)


def uniform_box_numpy(n, density=1.0e6 / 560.0 ** 3, seed=0, min_sep=1.1):
    """BASELINE config 5 on the host (small n): pos, goal ~ U(0, L)^3 with L^3 = n / density, yaw ~ U(0, 2pi),
    rejection to >= min_sep separation (2r + 0.1)."""
    rng = np.random.default_rng(seed)
    L = (n / density) ** (1.0 / 3.0)
    m = int(n * 1.08) + 64
    p = rng.uniform(0, L, (m, 3))
    # drop the later member of every pair closer than min_sep (hash grid)
    cell = np.floor(p / min_sep).astype(np.int64)
    key = {}
    keep = np.ones(m, bool)
    for i in range(m):
        c = tuple(cell[i])
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for j in key.get((c[0] + dx, c[1] + dy, c[2] + dz), ()):
                        if ((p[i] - p[j]) ** 2).sum() < min_sep ** 2:
                            ok = False
        if ok:
            key.setdefault(c, []).append(i)
        keep[i] = ok
    p = p[keep][:n]
    assert len(p) == n
    pos = np.zeros((n, 6))
    pos[:, :3] = p
    pos[:, 2] += 20.0
    pos[:, 3] = rng.uniform(0, 2 * np.pi, n)
    goal = np.zeros((n, 6))
    goal[:, :3] = rng.uniform(0, L, (n, 3))
    goal[:, 2] += 20.0
    return pos, goal, np.zeros((0, 4))


def uniform_box_torch(n, density=1.0e6 / 560.0 ** 3, seed=0, min_sep=1.1, device='cuda:0'):
    """BASELINE config 5 at scale (10^5-10^6): same distribution, generated on the device; the separation
    rejection uses the library's own 16-NN stage (drop the higher id of every pair closer than min_sep)."""
    import torch
    from .swarm import Swarm
    from . import capi
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = (n / density) ** (1.0 / 3.0)
    m = int(n * 1.06) + 1024
    p = torch.rand((m, 3), generator=g, device=device, dtype=torch.float64) * L
    p[:, 2] += 20.0
    sw = Swarm(m, device=device, debug=True)
    sw.set_state(p.cpu().numpy(), p.cpu().numpy(), vel=np.full((m, 3), 0.1, np.float32))
    sw.fit_grid()
    sw.stage_grid()
    sw.stage_neighbors(capi.NBR_KNN)
    torch.cuda.synchronize()
    ids, dsq, cnt = sw.nbr_ids, sw.nbr_dsq, sw.nbr_cnt
    me = torch.arange(m, device=device, dtype=torch.int32)[:, None]
    valid = torch.arange(capi.K, device=device)[None, :] < cnt[:, None]
    close = valid & (dsq < min_sep * min_sep) & (ids < me)
    keep = ~close.any(1)
    p = p[keep][:n]
    assert p.shape[0] == n, 'not enough points survived the separation filter'
    del sw
    pos = torch.zeros((n, 6), dtype=torch.float64, device=device)
    pos[:, :3] = p
    pos[:, 3] = torch.rand(n, generator=g, device=device, dtype=torch.float64) * (2 * math.pi)
    goal = torch.zeros((n, 6), dtype=torch.float64, device=device)
    goal[:, :3] = torch.rand((n, 3), generator=g, device=device, dtype=torch.float64) * L
    goal[:, 2] += 20.0
    return pos.cpu().numpy(), goal.cpu().numpy(), np.zeros((0, 4))


def spatial_order(pos, cell=10.0):
    """Permutation that lists the agents cell by cell (z-major, the order of the library's cell list).  Laying the
    state rows out in this order at t = 0 (the agents' own ids travel in `gid`, Swarm.set_global_ids) makes every
    gather / scatter between the caller's rows and the cell-sorted snapshot nearly coalesced; results do not depend
    on the row order because ties and the collision-check order follow the ids.  Agents move <= 0.1 m per step, so
    the layout stays coherent for thousands of steps."""
    p = np.asarray(pos, 'float64')[:, :3]
    c = np.floor((p - p.min(0)) / cell).astype(np.int64)
    dim = c.max(0) + 1
    return np.argsort((c[:, 2] * dim[1] + c[:, 1]) * dim[0] + c[:, 0], kind='stable')
