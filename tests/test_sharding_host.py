"""CPU: host-side logic of the multi-GPU path (slab plan, row packing, neighbour exchange) with a
world_size-2 gloo process group -- the same code paths ShardedSwarm drives over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sca_b200 import sharding


def test_slab_plan_partitions_evenly():
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 560, 100000)
    for world in (1, 2, 4, 8):
        plan = sharding.SlabPlan(x, world)
        own = plan.owner(x)
        counts = np.bincount(own, minlength=world)
        assert counts.sum() == len(x) and counts.max() - counts.min() <= 2
        for r in range(world):
            lo, hi = plan.faces(r)
            assert ((x[own == r] >= lo) & (x[own == r] < hi)).all()
        # halo must cover the interaction radius plus the drift allowed between migrations
        assert plan.width >= 10.0 + 2 * plan.migrate_every * 0.1
        if world > 2:
            assert plan.min_slab_width() > plan.width


def _fake_state(n, seed):
    g = torch.Generator().manual_seed(seed)
    t = {'pos': torch.rand((n, 3), generator=g, dtype=torch.float64) * 100, 'vel': torch.rand((n, 3), generator=g),
         'heading': torch.rand((n, 3), generator=g, dtype=torch.float64), 'goal': torch.rand((n, 3), generator=g, dtype=torch.float64),
         'radius': torch.full((n,), 0.5, dtype=torch.float64), 'pref_speed': torch.ones(n, dtype=torch.float64),
         'status': torch.randint(0, 32, (n,), generator=g, dtype=torch.int32), 'total_dist': torch.rand(n, generator=g, dtype=torch.float64),
         'step_num': torch.randint(0, 1000, (n,), generator=g, dtype=torch.int32),
         'max_run_dist': torch.rand(n, generator=g, dtype=torch.float64), 'gid': torch.arange(n, dtype=torch.int32) + 1000 * seed}
    return t


def test_row_pack_roundtrip_is_exact():
    t = _fake_state(64, 1)
    idx = torch.tensor([3, 9, 10, 63])
    rows = sharding.pack_rows(t, idx)
    assert rows.shape == (4, sharding.ROW_COLS)
    t2 = {k: torch.zeros_like(v) for k, v in t.items()}
    sharding.unpack_rows(t2, rows, 5)
    for k in t:
        assert torch.equal(t2[k][5:9], t[k][idx]), k


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # fixed-size neighbour exchange (the per-step halo path)
        cap = 8
        send_lo = torch.full((cap, 8), float(10 * rank + 1), dtype=torch.float64)
        send_hi = torch.full((cap, 8), float(10 * rank + 2), dtype=torch.float64)
        recv = torch.full((2 * cap, 8), -1.0, dtype=torch.float64)
        sharding.exchange_with_neighbours(send_lo, send_hi, recv[:cap], recv[cap:], rank, world)
        exp_lo = float(10 * (rank - 1) + 2) if rank > 0 else -1.0          # left neighbour's "hi" buffer
        exp_hi = float(10 * (rank + 1) + 1) if rank < world - 1 else -1.0   # right neighbour's "lo" buffer
        ok1 = bool((recv[:cap] == exp_lo).all() and (recv[cap:] == exp_hi).all())
        # variable-size migration exchange
        t = _fake_state(32, rank + 1)
        rows_lo = sharding.pack_rows(t, torch.arange(0, 3 + rank))
        rows_hi = sharding.pack_rows(t, torch.arange(10, 15 + 2 * rank))
        from_lo, from_hi = sharding.exchange_variable(rows_lo, rows_hi, rank, world)
        peer = _fake_state(32, (1 - rank) + 1)
        if rank == 0:
            exp = sharding.pack_rows(peer, torch.arange(0, 3 + 1))          # rank 1 sent its "lo" rows to us
            ok2 = from_lo.shape[0] == 0 and torch.equal(from_hi, exp)
        else:
            exp = sharding.pack_rows(peer, torch.arange(10, 15))            # rank 0 sent its "hi" rows to us
            ok2 = from_hi.shape[0] == 0 and torch.equal(from_lo, exp)
        q.put((rank, ok1, bool(ok2)))
    finally:
        dist.destroy_process_group()


def test_neighbour_exchange_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True, True), (1, True, True)], res


def test_a_slab_narrower_than_the_halo_is_refused():
    """ghosts come from rank +- 1 only, so an interior slab must be at least as wide as the halo."""
    import pytest
    from sca_b200 import capi
    from sca_b200.sharding import ShardedSwarm, SlabPlan
    rng = np.random.default_rng(0)
    pos = np.zeros((4000, 6)); pos[:, :3] = rng.uniform(0, 60, (4000, 3))
    with pytest.raises(capi.ScabError, match='narrower than the halo width'):
        ShardedSwarm(pos, pos, 1, 8, 'cpu', np.zeros(3), np.full(3, 60.0))        # 7.5 m slabs, 11.6 m halo
    # faster agents and a larger neighbor_dist widen the halo
    a = SlabPlan(pos[:, 0], 2, neighbor_dist=10.0, max_speed=1.0, dt=0.1, migrate_every=8)
    b = SlabPlan(pos[:, 0], 2, neighbor_dist=15.0, max_speed=4.0, dt=0.1, migrate_every=8)
    assert abs(a.width - 11.6) < 1e-5 and abs(b.width - (15.0 + 2 * 8 * 0.4)) < 1e-5
