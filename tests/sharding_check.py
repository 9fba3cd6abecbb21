"""Run under torchrun on N >= 2 GPUs: an N-slab sharded run must reproduce the 1-GPU run of the same
kernels BIT FOR BIT (positions, velocities, headings, flags, counters), including across migrations.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/sharding_check.py

SHARD_HOSTSIM=1 runs the same check without GPUs: the kernels (halo pack / unpack, migration, the whole step) come from the
host emulation build (tests/hostsim), the exchanges go over gloo, streams are no-ops and the step is launched eagerly.
"""
import os
import sys


import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sca_b200 import capi, scenarios  # noqa: E402
from sca_b200.sharding import ShardedSwarm  # noqa: E402
from sca_b200.swarm import Swarm  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('SHARD_WATCHDOG', '75')), exit=True)
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    import datetime
    hostsim = os.environ.get('SHARD_HOSTSIM', '0') == '1'
    if hostsim:
        import pytest
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import hostsim_util
        mp = pytest.MonkeyPatch()
        hostsim_util.activate(mp)
        hostsim_util.activate_streams(mp)
        os.environ['SCAB_NO_GRAPH'] = '1'
        dev = 'cpu'
        dist.init_process_group('gloo', timeout=datetime.timedelta(seconds=600))
    else:
        torch.cuda.set_device(local)
        dev = 'cuda:%d' % local
        dist.init_process_group('nccl', device_id=torch.device(dev), timeout=datetime.timedelta(seconds=90))
    n = int(os.environ.get('SHARD_N', '30000'))
    steps = int(os.environ.get('SHARD_STEPS', '40'))
    policy = os.environ.get('SHARD_POLICY', 'sca')
    density = 4.0 * 1.0e6 / 560.0 ** 3            # denser than config 5: more interaction across the faces
    pos, goal, _ = scenarios.uniform_box_numpy(n, density, seed=3)
    L = (n / density) ** (1 / 3)
    lo, hi = np.array([-20.0, -20.0, 0.0]), np.array([L + 20.0, L + 20.0, L + 40.0])
    # transport: peer-to-peer direct writes over NVLink on real GPUs, staged send / recv (gloo) between the emulated devices
    transport = os.environ.get('SHARD_TRANSPORT', 'nccl' if hostsim else 'p2p')
    sh = ShardedSwarm(pos, goal, rank, world, dev, lo, hi, migrate_every=4, mig_cap=int(os.environ.get('SHARD_MIG_CAP', '4096')),
                      transport=transport)
    sh.warm_up(policy, capi.NBR_KNN)
    # optional per-agent state that lives outside the migration row and must travel with its agent
    use_dubins = os.environ.get('SHARD_DUBINS', '0') == '1'        # device-side Dubins tracker (policy sca / rvo3d_dubins)
    use_wp = os.environ.get('SHARD_WAYPOINTS', '0') == '1'         # two-waypoint dog-leg per agent (rvo3d / srvo3d / orca)
    wp = wp_cnt = None
    if use_wp:
        rng = np.random.default_rng(5)
        s0, g0 = pos[:, :3], goal[:, :3]
        d0 = (g0 - s0) / np.linalg.norm(g0 - s0, axis=1, keepdims=True)
        # list order = the reference's: pop() takes the LAST entry first -> [far, middle, near]; the near one (0.8 m ahead) is
        # reached and popped within the run, after which the middle one is the target (left = 1, now_goal = middle) for the
        # rest of it -- the state an agent must bring along when it changes slabs (an empty path means "fly to the goal")
        wp = np.stack([s0 + 0.8 * (g0 - s0) + rng.normal(0, 1.0, (n, 3)), s0 + 0.5 * (g0 - s0) + rng.normal(0, 1.0, (n, 3)),
                       s0 + 0.8 * d0 + rng.normal(0, 0.05, (n, 3))], 1)
        wp_cnt = (3 - (np.arange(n) % 2)).astype(np.int32)
        sh.set_waypoints(wp, wp_cnt)
    if use_dubins:
        sh.enable_dubins(goal[:, 3:6], policy)
    if os.environ.get('SHARD_NO_EXTRA', '0') == '1':               # negative control: the extra rows stay behind
        sh._extra = None
    if os.environ.get('SHARD_DEBUG'):
        def _dbg(w):
            sh.stream.synchronize()
            print('[rank %d step %d] reached %s (everything before it completed)' % (rank, sh.steps, w), flush=True)
        sh.debug_phase = _dbg
    dist.barrier()
    own0 = set(np.nonzero(sh.plan.owner(pos[:, 0]) == rank)[0].tolist())
    for _ in range(steps):
        sh.step(policy, capi.NBR_KNN)
    sh.check_overflow()
    keys = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num') + (('waypoint_left', 'now_goal') if use_wp else ())
    gid, st = sh.gather_state(keys)
    moved = len(set(gid.tolist()) - own0)          # agents this rank owns now but did not own at the start
    act = sh.counters()
    glob = sh.global_counters()                      # one all-reduce: the whole-swarm counters (done test)
    all_done = sh.is_done()                          # (collective: every rank calls it)
    gathered = [None] * world
    dist.gather_object((gid, st, act, moved), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        ref = Swarm(n, device=dev)
        ref.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:6], initial_pos=pos, goal_pos=goal)
        ref.fit_grid(lo, hi)
        if use_wp:
            ref.set_waypoints(wp, wp_cnt)
        if use_dubins:
            ref.enable_dubins(goal[:, 3:6])
        for _ in range(steps):
            ref.step(policy, capi.NBR_KNN)
        torch.cuda.synchronize()
        full = {k: ref.t[k][:n].cpu().numpy() for k in keys}
        seen = np.zeros(n, bool)
        bad = {}
        for (g, s, a, mv) in gathered:
            assert not seen[g].any(), 'an agent is owned by two ranks'
            seen[g] = True
            for k in full:
                a_, b_ = s[k], full[k][g]
                if k == 'status':
                    a_, b_ = a_ & 15, b_ & 15
                if k == 'now_goal':
                    a_, b_ = np.nan_to_num(a_, nan=-1e9), np.nan_to_num(b_, nan=-1e9)
                bad[k] = bad.get(k, 0) + int((a_ != b_).reshape(len(g), -1).any(1).sum())
        assert seen.all(), 'agents lost: %d' % int((~seen).sum())
        tot_active = sum(a['n_active_total'] for (_, _, a, _) in gathered)
        ref_active = ref.counters()['n_active_total']
        migrations = sum(mv for (_, _, _, mv) in gathered)
        refc = ref.counters()
        ok = all(v == 0 for v in bad.values()) and tot_active == ref_active
        for k in ('n_active_total', 'n_active', 'n_not_done', 'n_collision'):
            ok = ok and glob[k] == refc[k]
        ok = ok and all_done == ref.is_done()
        if use_wp:
            popped = int((ref.t['waypoint_left'][:n].cpu().numpy() < wp_cnt).sum())
            ok = ok and popped > 0
            print('waypoints popped by %d agents' % popped)
        if use_dubins:
            replans = int(ref.dubins_states()['replans'].sum())
            ok = ok and replans >= n
            print('dubins plans made: %d' % replans)
        print('sharding_check transport=%s world=%d n=%d steps=%d policy=%s mismatches=%s agent_steps %d vs %d  migration-events %d  '
              'collisions(ref)=%d  -> %s' % (transport, world, n, steps, policy, bad, tot_active, ref_active, migrations,
                                             int((full['status'] & 2).astype(bool).sum()), 'OK' if ok else 'FAIL'), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    rc = 0 if int(flag.item()) else 1
    sh.close()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(rc)      # skip the blocking process-group teardown


if __name__ == '__main__':
    main()
