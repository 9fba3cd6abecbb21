"""CPU, world_size-2 and -4 gloo (at 4 ranks the inner slabs exchange on both faces at once): the multi-GPU step
PROTOCOL of sca_b200/sharding.py run end to end with the oracle as the per-rank stepper, against the single-process
oracle run of the same swarm.

ShardedSwarm's per-step sequence (`_step_body`) is replayed on the host with the module's own building blocks --
`SlabPlan` (equal-count slabs, halo width = neighborDist + 2 x drift), `pack_rows` / `unpack_rows` (full state rows),
`exchange_variable` (the send/recv pattern with the <= 2 slab neighbours) -- over a gloo process group:

    every `migrate_every` steps (and at step 1): agents outside their slab move to the neighbour rank
    halo rows of the agents within `width` of a face  ->  neighbour            (ghosts)
    local world = owned + ghosts, ordered by GLOBAL id (ties and the collision-check order follow it)  ->  policy
    the owners' actions of those same halo agents  ->  neighbour              (ghosts' new positions)
    kinematic update + collision / goal status; only owned rows are kept

The bar is the one tests/test_sharding_gpu.py holds the device path to on 2 GPUs: bit-for-bit equality with the
unsharded run, migrations and collisions included.  It pins down the plan itself (halo width, migration cadence,
what a ghost must carry) on every CPU run, where no second GPU is needed.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sca_b200 import sharding

KEYS = ('pos', 'vel', 'heading', 'goal', 'radius', 'pref_speed', 'status', 'total_dist', 'step_num', 'max_run_dist', 'gid')
CMP = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')


def _state_from_world(w, gid):
    t = {k: torch.from_numpy(np.array(getattr(w, k), copy=True)) for k in KEYS if k != 'gid'}
    t['gid'] = torch.from_numpy(np.asarray(gid, np.int32).copy())
    return t


def _take(t, idx):
    idx = torch.as_tensor(idx, dtype=torch.long)
    return {k: v[idx].clone() for k, v in t.items()}


def _cat(a, rows):
    """Append unpacked `rows` (float64 [k, ROW_COLS]) to the state dict `a`."""
    k = rows.shape[0]
    if k == 0:
        return a
    out = {key: torch.cat([v, torch.zeros((k,) + tuple(v.shape[1:]), dtype=v.dtype)]) for key, v in a.items()}
    sharding.unpack_rows(out, rows, len(a['gid']))
    return out


def _world_from_state(orc, t):
    """Oracle world over the rows of `t` (already ordered by gid), teacher-forced to their state."""
    n = len(t['gid'])
    w = orc.World(t['pos'].numpy(), t['goal'].numpy(), heading=t['heading'].numpy(), radius=t['radius'].numpy(),
                  pref_speed=t['pref_speed'].numpy())
    w.vel[:] = t['vel'].numpy()
    w.status[:] = t['status'].numpy()
    w.total_dist[:] = t['total_dist'].numpy()
    w.step_num[:] = t['step_num'].numpy()
    w.max_run_dist[:] = t['max_run_dist'].numpy()
    w.is_zaxis[:] = 0
    assert w.n == n
    return w


def _worker(rank, world, port, q, policy, pos6, goal6, steps, migrate_every, expect):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        orc.set_threads(2)
        n = len(pos6)
        plan = sharding.SlabPlan(pos6[:, 0], world, migrate_every=migrate_every)
        w0 = orc.World(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], initial_pos=pos6, goal_pos=goal6)
        own = _take(_state_from_world(w0, np.arange(n)), np.nonzero(plan.owner(pos6[:, 0]) == rank)[0])
        lo_f, hi_f = plan.faces(rank)
        stats = dict(migrated=0, halo_max=0, collisions=0)
        empty = torch.zeros((0, sharding.ROW_COLS), dtype=torch.float64)
        for step in range(steps):
            # -- migration (ShardedSwarm.step: steps > 0 and (steps % migrate_every == 0 or steps == 1))
            if step > 0 and (step % plan.migrate_every == 0 or step == 1):
                x = own['pos'][:, plan.axis]
                go_lo = torch.nonzero(x < lo_f)[:, 0] if rank > 0 else torch.zeros(0, dtype=torch.long)
                go_hi = torch.nonzero(x >= hi_f)[:, 0] if rank < world - 1 else torch.zeros(0, dtype=torch.long)
                from_lo, from_hi = sharding.exchange_variable(sharding.pack_rows(own, go_lo), sharding.pack_rows(own, go_hi), rank, world)
                keep = torch.ones(len(x), dtype=torch.bool)
                keep[go_lo] = False
                keep[go_hi] = False
                own = _cat(_cat(_take(own, torch.nonzero(keep)[:, 0]), from_lo), from_hi)
                stats['migrated'] += len(go_lo) + len(go_hi)
            # an owned agent never drifts further than plan.drift outside its slab between migrations
            x = own['pos'][:, plan.axis]
            assert float(x.min()) >= lo_f - plan.drift - 1e-9 and float(x.max()) < hi_f + plan.drift + 1e-9
            # -- pre-exchange: ghosts = the neighbour's agents within `width` of the shared face
            h_lo = torch.nonzero(x < lo_f + plan.width)[:, 0] if rank > 0 else torch.zeros(0, dtype=torch.long)
            h_hi = torch.nonzero(x >= hi_f - plan.width)[:, 0] if rank < world - 1 else torch.zeros(0, dtype=torch.long)
            g_lo, g_hi = sharding.exchange_variable(sharding.pack_rows(own, h_lo), sharding.pack_rows(own, h_hi), rank, world)
            stats['halo_max'] = max(stats['halo_max'], len(h_lo), len(h_hi))
            n_own = len(own['gid'])
            local = _cat(_cat(own, g_lo), g_hi)
            order = torch.argsort(local['gid'].long())
            inv = torch.empty_like(order)
            inv[order] = torch.arange(len(order))
            w = _world_from_state(orc, _take(local, order))
            # -- policy on the local world
            w.policy_step(policy, orc.NBR_KNN)
            # -- post-exchange: the owners' actions of the halo agents, in the order their rows were sent
            act = torch.from_numpy(w.action.astype(np.float64))
            a_lo, a_hi = sharding.exchange_variable(act[inv[h_lo]].contiguous(), act[inv[h_hi]].contiguous(), rank, world)
            ghosts = torch.cat([a_lo.reshape(-1, 7), a_hi.reshape(-1, 7)]) if (len(a_lo) + len(a_hi)) else torch.zeros((0, 7), dtype=torch.float64)
            assert ghosts.shape[0] == len(local['gid']) - n_own
            if ghosts.shape[0]:
                w.action[inv[n_own:].numpy()] = ghosts.numpy().astype(np.float32)
            # -- kinematic update + status; keep the owned rows
            w.env_update(fast=True)
            new = _state_from_world(w, local['gid'][order].numpy())
            own = _take(new, inv[:n_own])
        stats['collisions'] = int(((own['status'] & 2) != 0).sum())
        # -- compare the owned rows with the single-process run
        gid = own['gid'].long().numpy()
        bad = {k: int((own[k].numpy() != expect[k][gid]).reshape(len(gid), -1).any(1).sum()) for k in CMP}
        q.put((rank, len(gid), bad, stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('policy,steps,world', [('sca', 12, 2), ('orca_lp', 26, 4)], ids=['sca-2', 'orca_lp-4'])
def test_slab_protocol_reproduces_the_single_world(policy, steps, world):
    from oracle import oracle as orc
    from sca_b200 import scenarios
    n, migrate_every = 2400, 8
    # dense, and a few pairs start closer than 2r: collision mode of the neighbour lists and the collision flags are hit
    pos6, goal6, _ = scenarios.uniform_box_numpy(n, density=2.0 * 1.0e6 / 560.0 ** 3, seed=9, min_sep=0.97)
    orc.set_threads(8)
    w = orc.World(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], initial_pos=pos6, goal_pos=goal6)
    for _ in range(steps):
        w.step(policy, orc.NBR_KNN, fast=True)
    expect = {k: np.array(getattr(w, k), copy=True) for k in CMP}
    assert int(((expect['status'] & 2) != 0).sum()) > 0
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (7 if policy == 'sca' else 0) + 13 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, policy, pos6, goal6, steps, migrate_every, expect)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    print(policy, res)
    assert sum(r[1] for r in res) == n                      # every agent is owned by exactly one rank at the end
    for rank, cnt, bad, stats in res:
        assert all(v == 0 for v in bad.values()), (rank, bad)
        assert stats['halo_max'] > 0
    if steps > 16:
        assert sum(r[3]['migrated'] for r in res) > 0       # ownership did change hands during the run
