"""Multi-GPU spatial sharding: one process per GPU, slabs along x, halo exchange with the two slab neighbours.

The reference is single-process; this is the only parallelism the new build adds.
Per step and per rank (all asynchronous on the step's stream, fixed-size buffers, no host round trips):

    halo_pack  ->  exchange with the <= 2 slab neighbours  ->  halo_unpack          (ghost rows)
    cell list  ->  16-NN  ->  policy kernels (owned agents only)
    halo_pack_pos -> exchange -> halo_unpack_pos                                     (ghosts' new positions)
    status kernel (owned agents only)

The exchange is either fused into the pack kernels -- they write straight into the neighbour GPU's receive buffers over NVLink
(CUDA IPC), a flag kernel publishes the data, the unpack kernel waits for it: transport='p2p', no communication library in the
step -- or staged through torch.distributed send / recv (transport='nccl'; gloo in the CPU tests).  `LogicalShards` runs S slabs
in ONE process on one device with the same kernels (transport='local').
Every `migrate_every` steps agents that left their slab move to the neighbour rank (full state rows).
Ties and the sequential collision-check order follow GLOBAL agent ids, so an S-shard run reproduces the
1-GPU run of the same kernels bit for bit (tests/test_gpu_logical_shards.py, tests/test_sharding_gpu.py).

`SlabPlan` and the row packing are plain host logic (device agnostic) and are covered on CPU with
world_size-2 gloo tests (tests/test_sharding_host.py, tests/test_sharding_hostsim.py).
"""
import numpy as np
import torch

from . import capi

REC_DOUBLES = 8          # 64-byte halo record
KIND_HALO, KIND_POS, KIND_MIG, KIND_MIGX = 0, 1, 2, 3      # peer-to-peer exchange kinds (flag / sequence words of the control block)


def _view(ptr, shape, dtype, device):
    """A torch tensor over raw library-owned memory (scab_p2p_alloc): device memory through __cuda_array_interface__, host
    memory (the emulation build of the tests) through ctypes."""
    import ctypes as C
    n = int(np.prod(shape))
    if torch.device(device).type == 'cpu':
        ct = {torch.float64: C.c_double, torch.int64: C.c_int64, torch.int32: C.c_int32}[dtype]
        return torch.from_numpy(np.ctypeslib.as_array((ct * n).from_address(ptr))).view(*shape)
    ts = {torch.float64: '<f8', torch.int64: '<i8', torch.int32: '<i4'}[dtype]

    class _Raw:
        __cuda_array_interface__ = {'shape': tuple(int(x) for x in shape), 'typestr': ts, 'data': (int(ptr), False), 'version': 2}
    return torch.as_tensor(_Raw(), device=device)

ROW_COLS = 22            # full state row used for migration (19 used + 3 spare)


class SlabPlan:
    """Equal-count slabs along one axis + the halo geometry."""

    def __init__(self, coord, world, neighbor_dist=10.0, max_speed=1.0, dt=0.1, migrate_every=16, axis=0):
        coord = np.asarray(coord, 'float64')
        self.world = int(world)
        self.axis = axis
        qs = np.quantile(coord, np.linspace(0, 1, world + 1)[1:-1]) if world > 1 else np.zeros(0)
        self.edges = np.concatenate([[-np.inf], qs, [np.inf]])
        self.migrate_every = int(migrate_every)
        # an agent may sit up to `drift` outside its owner's slab before it migrates, and so may the ghosts
        self.drift = migrate_every * max_speed * dt
        self.width = neighbor_dist + 2.0 * self.drift + 1e-6

    def owner(self, coord):
        return np.clip(np.searchsorted(self.edges, np.asarray(coord, 'float64'), side='right') - 1, 0, self.world - 1)

    def faces(self, rank):
        return float(self.edges[rank]), float(self.edges[rank + 1])

    def min_slab_width(self):
        w = np.diff(self.edges[1:-1]) if self.world > 2 else np.array([np.inf])
        return float(w.min()) if len(w) else np.inf


def pack_rows(t, idx):
    """Full state rows (float64 [k, ROW_COLS]) of agents `idx` -- exact for f32 / int32 columns."""
    f = lambda k: t[k][idx].to(torch.float64)
    return torch.cat([f('pos'), f('vel'), f('heading'), f('goal'), f('radius')[:, None], f('pref_speed')[:, None],
                      f('status')[:, None], f('total_dist')[:, None], f('step_num')[:, None], f('max_run_dist')[:, None],
                      f('gid')[:, None], torch.zeros((len(idx), 3), dtype=torch.float64, device=t['pos'].device)], dim=1)


def unpack_rows(t, rows, row0):
    k = rows.shape[0]
    sl = slice(row0, row0 + k)
    t['pos'][sl] = rows[:, 0:3]
    t['vel'][sl] = rows[:, 3:6].to(torch.float32)
    t['heading'][sl] = rows[:, 6:9]
    t['goal'][sl] = rows[:, 9:12]
    t['radius'][sl] = rows[:, 12]
    t['pref_speed'][sl] = rows[:, 13]
    t['status'][sl] = rows[:, 14].to(torch.int32)
    t['total_dist'][sl] = rows[:, 15]
    t['step_num'][sl] = rows[:, 16].to(torch.int32)
    t['max_run_dist'][sl] = rows[:, 17]
    t['gid'][sl] = rows[:, 18].to(torch.int32)


def exchange_with_neighbours(send_lo, send_hi, recv_lo, recv_hi, rank, world, dist=None):
    """One grouped send/recv with the slab neighbours (rank-1 receives send_lo, rank+1 receives send_hi)."""
    if world == 1:
        return
    dist = dist or torch.distributed
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, send_lo, rank - 1), dist.P2POp(dist.irecv, recv_lo, rank - 1)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, send_hi, rank + 1), dist.P2POp(dist.irecv, recv_hi, rank + 1)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()


def exchange_variable(rows_lo, rows_hi, rank, world, dist=None):
    """Variable-length row exchange (migration): counts first, then payloads.  Returns (from_lo, from_hi)."""
    dist = dist or torch.distributed
    dev = rows_lo.device
    c_send = torch.tensor([rows_lo.shape[0], rows_hi.shape[0]], dtype=torch.int64, device=dev)
    c_lo, c_hi = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
    exchange_with_neighbours(c_send[0:1].clone(), c_send[1:2].clone(), c_lo, c_hi, rank, world, dist)
    n_lo, n_hi = int(c_lo.item()), int(c_hi.item())
    cols = rows_lo.shape[1]
    from_lo = torch.zeros((n_lo, cols), dtype=rows_lo.dtype, device=dev)
    from_hi = torch.zeros((n_hi, cols), dtype=rows_lo.dtype, device=dev)
    ops = []
    if rank > 0:
        if rows_lo.shape[0]:
            ops.append(dist.P2POp(dist.isend, rows_lo.contiguous(), rank - 1))
        if n_lo:
            ops.append(dist.P2POp(dist.irecv, from_lo, rank - 1))
    if rank < world - 1:
        if rows_hi.shape[0]:
            ops.append(dist.P2POp(dist.isend, rows_hi.contiguous(), rank + 1))
        if n_hi:
            ops.append(dist.P2POp(dist.irecv, from_hi, rank + 1))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return from_lo, from_hi


class ShardedSwarm:
    """The rank-local part of a swarm sharded over `world` GPUs.  Same step()/counters() surface as Swarm.

    Row layout (STATIC row count, so one captured CUDA graph replays every step):
        [0, own_cap)                 owned slots: n_own live agents, then parked ghosts
        [own_cap, own_cap + 2 cap)   halo ghosts received from the lower / upper slab neighbour
    Parked ghosts sit in a reserved, otherwise empty layer of the uniform grid: no kernel ever sees them as
    neighbours and every per-agent stage skips SCAB_ST_GHOST rows.
    """

    def __init__(self, pos6, goal6, rank, world, device, lo, hi, radius=0.5, pref_speed=1.0, migrate_every=8,
                 halo_cap=None, debug=False, use_graph=True, mig_cap=4096, params=None, transport='nccl', own_slack=0.2):
        """`transport`: how the halo / migration buffers reach the slab neighbours --
        'nccl'  pack -> torch.distributed send/recv -> unpack (any backend; gloo in the CPU tests);
        'p2p'   the pack kernels write straight into the neighbour's receive buffers over NVLink (CUDA IPC between the ranks of
                one node) and a flag kernel publishes them: no communication library in the step (scab_p2p.cu);
        'local' the same direct writes between logical shards of ONE process (LogicalShards wires the pointers)."""
        assert transport in ('nccl', 'p2p', 'local')
        self.transport = transport if world > 1 else 'nccl'
        from .swarm import Swarm
        pos6, goal6 = np.asarray(pos6, 'float64'), np.asarray(goal6, 'float64')
        n_global = len(pos6)
        self.rank, self.world, self.n_global = rank, world, n_global
        from . import capi as _capi
        prm = params if params is not None else _capi.default_params()
        pref_g = np.broadcast_to(np.asarray(pref_speed, 'float64'), (n_global,))
        v_bound = max(float(prm.max_speed), float(pref_g.max()) if n_global else 0.0, 0.5)    # == Swarm.v_bound over ALL shards
        self.plan = SlabPlan(pos6[:, 0], world, neighbor_dist=float(prm.neighbor_dist), max_speed=v_bound, dt=float(prm.time_step),
                             migrate_every=migrate_every)
        # ghosts come from the two adjacent slabs only: an interior slab narrower than the halo width would hide
        # neighbours (and collision partners) that live two slabs away
        if self.plan.min_slab_width() < self.plan.width:
            raise capi.ScabError('slab of width %.3f m is narrower than the halo width %.3f m: use fewer shards (or a smaller '
                                 'migrate_every)' % (self.plan.min_slab_width(), self.plan.width))
        own_all = self.plan.owner(pos6[:, 0])
        mine = np.nonzero(own_all == rank)[0]
        from .scenarios import spatial_order
        mine = mine[spatial_order(pos6[mine, :3])]     # rows laid out cell by cell: coalesced gathers / scatters
        self.n_own = len(mine)
        self._mine = mine
        # halo capacity per face: agents within `width` of a face, with head room
        # (the SAME value on every rank: the face buffers are exchanged as fixed-size messages)
        if halo_cap is None:
            near = 0
            for r in range(world):
                lo_f, hi_f = self.plan.faces(r)
                x = pos6[own_all == r, 0]
                near = max(near, int((x < lo_f + self.plan.width).sum()) if r > 0 else 0,
                           int((x >= hi_f - self.plan.width).sum()) if r < world - 1 else 0)
            halo_cap = int(near * 1.25) + 1024      # unused slots are parked ghost rows every per-row kernel still visits: keep it tight
                                                    # (an overflow is loud: the sticky OVF_HALO flag)
        self.cap = int(halo_cap) if world > 1 else 0
        self.own_cap = int(self.n_own * (1.0 + own_slack)) + 4096     # free rows for arrivals (an arrival without one is loud: OVF_MIGRATION_LOST)
        self.n_rows = self.own_cap + 2 * self.cap
        sw = Swarm(self.n_own, device=device, debug=debug, capacity=self.n_rows, params=params)
        rad_g = np.broadcast_to(np.asarray(radius, 'float64'), (n_global,))
        sw.set_state(pos6[mine, :3], goal6[mine, :3], heading=pos6[mine, 3:6], radius=rad_g[mine], pref_speed=pref_g[mine],
                     initial_pos=pos6[mine], goal_pos=goal6[mine])
        sw.h.set_max_radius(float(rad_g.max()))      # ghosts included: the bounds must hold across shards
        sw.v_bound = v_bound
        sw.h.set_motion_bound(v_bound)
        sw.set_global_ids(mine.astype(np.int32), n_global)
        # two extra z layers: an empty buffer layer and the layer where unused slots are parked
        lo, hi = np.asarray(lo, 'float64').copy(), np.asarray(hi, 'float64').copy()
        cell = float(max(sw.params.cell_size, sw.params.neighbor_dist))
        # the cell list only has to cover this slab plus its halo (agents beyond are clamped into the border
        # cells, which is still correct): the per-step scan over cells shrinks with the number of GPUs
        lo_f, hi_f = self.plan.faces(rank)
        ax = self.plan.axis
        if rank > 0:
            lo[ax] = max(lo[ax], lo_f - self.plan.width - 2 * cell)
        if rank < world - 1:
            hi[ax] = min(hi[ax], hi_f + self.plan.width + 2 * cell)
        dim = np.maximum(1, np.ceil((hi - lo) / cell).astype(np.int64))
        self._park_geom = (lo.copy(), cell, dim.copy())
        # physically remote (z = 1e9 clamps into the reserved top layer of the grid): an agent that climbs above the box is
        # clamped into that layer too and scans the parked rows, but can never be within neighbor_dist of one
        self.park = (float(lo[0] + 0.5 * cell), float(lo[1] + 0.5 * cell), 1.0e9)
        dim[2] += 2
        sw.h.set_grid(lo, cell, dim)
        self.sw = sw
        dev = sw.device
        f8 = dict(dtype=torch.float64, device=dev)
        self.send_lo, self.send_hi = torch.zeros((self.cap, REC_DOUBLES), **f8), torch.zeros((self.cap, REC_DOUBLES), **f8)
        self.psend_lo, self.psend_hi = torch.zeros((self.cap, 3), **f8), torch.zeros((self.cap, 3), **f8)
        # device-side migration: 128-byte full-state rows, 1 header row + mig_cap rows per face, and the free list
        self.mig_cap = int(mig_cap)
        mk = lambda: torch.zeros((self.mig_cap + 1, 16), **f8)
        self.msend_lo, self.msend_hi = mk(), mk()
        if self.transport != 'nccl' and not self._p2p_setup(dev):
            # peer memory is not available between these ranks (every rank agrees on that): the staged transport instead
            import warnings
            warnings.warn('sca_b200: peer-to-peer buffers could not be shared between the ranks; falling back to transport=nccl')
            self.transport = 'nccl'
        if self.transport == 'nccl':
            self.recv = torch.zeros((2 * self.cap, REC_DOUBLES), **f8)
            self.precv = torch.zeros((2 * self.cap, 3), **f8)
            self.mrecv_lo, self.mrecv_hi = mk(), mk()
        i4 = dict(dtype=torch.int32, device=dev)
        self.free_list = torch.zeros(self.own_cap, **i4)
        nfree = self.own_cap - self.n_own
        self.free_list[:nfree] = torch.arange(self.own_cap - 1, self.n_own - 1, -1, **i4)
        self.free_count = torch.tensor([nfree], **i4)
        self.mig_cnt = torch.zeros(4, **i4)
        self._extra = None       # per-agent state outside the migration row (Dubins tracker / waypoints): see _enable_extra
        self.steps = 0
        self._fill_parked()
        self._park_rows(self.n_own, self.n_rows)
        # a dedicated (non-default) stream: the legacy default stream would implicitly serialise the step
        # against NCCL's own stream
        self.stream = torch.cuda.Stream(device=dev)
        self.debug_phase = None
        import os as _os
        self.use_graph = bool(use_graph) and _os.environ.get('SCAB_NO_GRAPH', '0') != '1'
        self._graphs = {}
        sw._bind(self.n_rows)
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ peer-to-peer transport
    def _p2p_setup(self, dev):
        """One library-owned arena per rank that the neighbours can write: [control block | halo records | new positions |
        migration rows from below | from above]; the same offsets on every rank (cap and mig_cap are global)."""
        cap, mc = self.cap, self.mig_cap
        off, sizes = {}, (('ctl', 32 * 8), ('recv', 2 * cap * REC_DOUBLES * 8), ('precv', 2 * cap * 3 * 8),
                          ('mrecv_lo', (mc + 1) * 128), ('mrecv_hi', (mc + 1) * 128))
        total = 0
        for k, b in sizes:
            off[k] = total
            total += (b + 255) // 256 * 256
        self._off = off
        ipc = self.transport == 'p2p'
        try:
            self._arena, handle = self.sw.h.p2p_alloc(total, want_handle=ipc)
        except capi.ScabError:
            if not ipc:
                raise
            self._arena, handle = 0, None          # (still takes part in the collectives below)
        if ipc:
            peers = self._open_neighbours(handle)
            if peers is None:
                return False
        a = self._arena
        self.ctl = _view(a + off['ctl'], (32,), torch.int64, dev)
        self.recv = _view(a + off['recv'], (2 * cap, REC_DOUBLES), torch.float64, dev)
        self.precv = _view(a + off['precv'], (2 * cap, 3), torch.float64, dev)
        self.mrecv_lo = _view(a + off['mrecv_lo'], (mc + 1, 16), torch.float64, dev)
        self.mrecv_hi = _view(a + off['mrecv_hi'], (mc + 1, 16), torch.float64, dev)
        self._peer = {'lo': 0, 'hi': 0}          # base of the neighbours' arenas in THIS process's address space (0: none)
        self._xarena, self._xpeer = 0, {'lo': 0, 'hi': 0}
        if ipc:
            self._peer = {'lo': peers[0], 'hi': peers[1]}
        return True

    def _open_neighbours(self, handle):
        """Collective: every rank publishes the IPC handle of its arena and maps its slab neighbours'.  Returns None -- on
        EVERY rank -- if any rank could not allocate or map (no peer access between the devices, IPC disabled ...)."""
        handles = [None] * self.world
        torch.distributed.all_gather_object(handles, handle)
        ok, lo_b, hi_b = all(h is not None for h in handles), 0, 0
        if ok:
            try:
                lo_b = self.sw.h.p2p_open(handles[self.rank - 1]) if self.rank > 0 else 0
                hi_b = self.sw.h.p2p_open(handles[self.rank + 1]) if self.rank < self.world - 1 else 0
            except capi.ScabError:
                ok = False
        oks = [None] * self.world
        torch.distributed.all_gather_object(oks, bool(ok))
        return (lo_b, hi_b) if all(oks) else None

    def _remote(self, side, offset, fallback):
        """Address of byte `offset` of the `side` ('lo' / 'hi') neighbour's arena; a local scratch tensor on a side without one."""
        has = self.rank > 0 if side == 'lo' else self.rank < self.world - 1
        return self._peer[side] + offset if has else fallback.data_ptr()

    def _signal(self, kind):
        c = self.ctl.data_ptr()
        lo = self._peer['lo'] + self._off['ctl'] + 8 * (2 * kind + 1) if self.rank > 0 else 0             # I am its upper neighbour
        hi = self._peer['hi'] + self._off['ctl'] + 8 * (2 * kind) if self.rank < self.world - 1 else 0   # I am its lower neighbour
        self.sw.h.p2p_signal(c + 8 * (8 + kind), lo, hi, self._stream())

    def _wait(self, kind):
        """Folded into the unpack kernel that follows (no launch of its own): it waits until both neighbours' flags of this
        exchange kind reach this rank's own sequence number (set by _signal just before)."""
        c = self.ctl.data_ptr()
        self.sw.h.p2p_wait_next(c + 8 * (8 + kind), c + 8 * (2 * kind) if self.rank > 0 else 0,
                                c + 8 * (2 * kind + 1) if self.rank < self.world - 1 else 0)

    # ------------------------------------------------------------------
    def _park_rows(self, r0, r1):
        """Turn state rows [r0, r1) into parked ghosts, spread over the cells of the parking layer."""
        if r1 <= r0:
            return
        t = self.sw.t
        lo, cell, dim = self._park_geom
        dev = t['pos'].device
        i = torch.arange(r0, r1, device=dev)
        t['pos'][r0:r1, 0] = float(lo[0]) + ((i * 7919) % int(dim[0])).to(torch.float64) * cell + 0.5 * cell
        t['pos'][r0:r1, 1] = float(lo[1]) + ((i * 104729) % int(dim[1])).to(torch.float64) * cell + 0.5 * cell
        t['pos'][r0:r1, 2] = self.park[2]
        t['vel'][r0:r1] = 0
        t['radius'][r0:r1] = 0
        t['status'][r0:r1] = capi.ST_GHOST | capi.ST_AT_GOAL | capi.ST_RUN_DONE
        t['gid'][r0:r1] = 0x7fffffff

    def _fill_parked(self):
        """Edge ranks never receive from one side: those receive slots stay parked."""
        if self.cap == 0:
            return
        self.recv.zero_()
        self.recv[:, 0], self.recv[:, 1], self.recv[:, 2] = self.park
        # ints 8-10 = vx vy vz, 11 = status, 12 = gid (64-byte halo record)
        raw = self.recv.view(torch.int32).view(2 * self.cap, 2 * REC_DOUBLES)
        raw[:, 11] = capi.ST_GHOST | capi.ST_AT_GOAL | capi.ST_RUN_DONE
        raw[:, 12] = 0x7fffffff
        self.precv[:, 0], self.precv[:, 1], self.precv[:, 2] = self.park

    @property
    def t(self):
        return self.sw.t

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _mark(self, what):
        if self.debug_phase is not None:
            self.debug_phase(what)

    def _ghost_pre_pack(self):
        sw, cap = self.sw, self.cap
        lo_f, hi_f = self.plan.faces(self.rank)
        lo_face = lo_f if self.rank > 0 else -1e300            # no neighbour on that side: select nothing
        hi_face = hi_f if self.rank < self.world - 1 else 1e300
        if self.transport != 'nccl':      # straight into the neighbours' receive buffers
            # the lower neighbour's "from above" half, the upper neighbour's "from below" half
            out_lo = self._remote('lo', self._off['recv'] + cap * REC_DOUBLES * 8, self.send_lo)
            out_hi = self._remote('hi', self._off['recv'], self.send_hi)
            sw.h.halo_pack(self.own_cap, self.plan.axis, lo_face, hi_face, self.plan.width, self.park, out_lo, out_hi, cap, self._stream())
            self._signal(KIND_HALO)
            return KIND_HALO
        sw.h.halo_pack(self.own_cap, self.plan.axis, lo_face, hi_face, self.plan.width, self.park,
                       self.send_lo.data_ptr(), self.send_hi.data_ptr(), cap, self._stream())
        return (self.send_lo, self.send_hi, self.recv[:cap], self.recv[cap:])

    def _ghost_pre_unpack(self):
        self.sw.h.halo_unpack(self.recv.data_ptr(), 2 * self.cap, self.own_cap, self._stream())

    def _ghost_post_pack(self):
        sw, cap = self.sw, self.cap
        if self.transport != 'nccl':
            out_lo = self._remote('lo', self._off['precv'] + cap * 3 * 8, self.psend_lo)
            out_hi = self._remote('hi', self._off['precv'], self.psend_hi)
            sw.h.halo_pack_pos(out_lo, out_hi, cap, self.park, self._stream())
            self._signal(KIND_POS)
            return KIND_POS
        sw.h.halo_pack_pos(self.psend_lo.data_ptr(), self.psend_hi.data_ptr(), cap, self.park, self._stream())
        return (self.psend_lo, self.psend_hi, self.precv[:cap], self.precv[cap:])

    def _ghost_post_unpack(self):
        self.sw.h.halo_unpack_pos(self.precv.data_ptr(), 2 * self.cap, self.own_cap, self._stream())

    def _migrate_pack(self):
        """Agents that left the slab move to the neighbour rank -- on the device, fixed-size, no host sync."""
        sw = self.sw
        lo_f, hi_f = self.plan.faces(self.rank)
        lo_face = lo_f if self.rank > 0 else -1e300
        hi_face = hi_f if self.rank < self.world - 1 else 1e300
        lo, cell, dim = self._park_geom
        s = self._stream()
        x = self._extra
        if self.transport != 'nccl':
            out_lo = self._remote('lo', self._off['mrecv_hi'], self.msend_lo)
            out_hi = self._remote('hi', self._off['mrecv_lo'], self.msend_hi)
            sw.h.migrate_pack(self.own_cap, self.plan.axis, lo_face, hi_face, (lo[0], lo[1], cell, self.park[2]), int(dim[0]), int(dim[1]),
                              out_lo, out_hi, self.mig_cap, self.free_list.data_ptr(), self.free_count.data_ptr(), self.mig_cnt.data_ptr(), s)
            if x is not None:
                xlo = self._xpeer['lo'] + self._xoff['recv_hi'] if self.rank > 0 else x['send_lo'].data_ptr()
                xhi = self._xpeer['hi'] + self._xoff['recv_lo'] if self.rank < self.world - 1 else x['send_hi'].data_ptr()
                sw.h.migrate_extra_pack(xlo, xhi, self.mig_cap, self.mig_cnt.data_ptr(), s)
            self._signal(KIND_MIG)
            return [KIND_MIG]
        sw.h.migrate_pack(self.own_cap, self.plan.axis, lo_face, hi_face, (lo[0], lo[1], cell, self.park[2]), int(dim[0]), int(dim[1]),
                          self.msend_lo.data_ptr(), self.msend_hi.data_ptr(), self.mig_cap, self.free_list.data_ptr(),
                          self.free_count.data_ptr(), self.mig_cnt.data_ptr(), s)
        out = [(self.msend_lo, self.msend_hi, self.mrecv_lo, self.mrecv_hi)]
        if x is not None:        # the Dubins tracker / waypoint state of the same agents, same slots
            sw.h.migrate_extra_pack(x['send_lo'].data_ptr(), x['send_hi'].data_ptr(), self.mig_cap, self.mig_cnt.data_ptr(), s)
            out.append((x['send_lo'], x['send_hi'], x['recv_lo'], x['recv_hi']))
        return out

    def _migrate_unpack(self):
        sw, s, x = self.sw, self._stream(), self._extra
        sw.h.migrate_unpack(self.mrecv_lo.data_ptr(), self.mrecv_hi.data_ptr(), self.mig_cap, self.rank > 0, self.rank < self.world - 1,
                            self.free_list.data_ptr(), self.free_count.data_ptr(), self.mig_cnt.data_ptr(), s)
        if x is not None:
            sw.h.migrate_extra_unpack(x['recv_lo'].data_ptr(), x['recv_hi'].data_ptr(), self.mig_cap, s)

    def _step_phases(self, pol, nbr_mode, with_migration=False):
        """One step as a generator: static shapes, static pointers.  It yields at every exchange point the tuple
        (send_lo, send_hi, recv_lo, recv_hi) the caller has to move between the slab neighbours -- NCCL send/recv for
        one process per GPU (_step_body), plain device copies for several logical shards in one process (LogicalShards)."""
        sw = self.sw
        if self.world > 1:
            if with_migration:
                self._mark('migrate')
                for x in self._migrate_pack():
                    yield x
                    if isinstance(x, int):
                        self._wait(x)
                self._migrate_unpack()
            self._mark('pre-exchange')
            x = self._ghost_pre_pack()
            yield x
            if isinstance(x, int):
                self._wait(x)
            self._ghost_pre_unpack()
        s = self._stream()
        self._mark('grid')
        sw.h.build_grid(s)
        self._mark('neighbors')
        sw.h.neighbors(nbr_mode, s)
        self._mark('policy')
        sw.h.policy_step(pol, s)
        if self.world > 1:
            self._mark('post-exchange')
            x = self._ghost_post_pack()
            yield x
            if isinstance(x, int):
                self._wait(x)
            self._ghost_post_unpack()
        self._mark('status')
        sw.h.status_step(s)

    def _step_body(self, pol, nbr_mode, with_migration=False):
        """One step -> capturable in a CUDA graph (two variants).  Exchange points: torch.distributed send/recv ('nccl'), or
        nothing at all ('p2p': the data is already in the neighbours' buffers, the flag kernels order it)."""
        for x in self._step_phases(pol, nbr_mode, with_migration):
            if not isinstance(x, int):
                exchange_with_neighbours(x[0], x[1], x[2], x[3], self.rank, self.world)

    def _wants_migration(self):
        # step 1 also migrates, so that both graph variants are exercised eagerly during any warm-up
        return self.world > 1 and self.steps > 0 and (self.steps % self.plan.migrate_every == 0 or self.steps == 1)

    def step(self, policy, nbr_mode=capi.NBR_KNN):
        pol = capi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        with torch.cuda.stream(self.stream):
            mig = self._wants_migration()
            key = (pol, nbr_mode, mig)
            # the first steps run eagerly (NCCL creates its channels lazily on first use), then graphs replay
            if self.use_graph and self.debug_phase is None and self.steps >= 3:
                g = self._graphs.get(key)
                if g is None:
                    g = self._capture(pol, nbr_mode, mig)
                if g is not False:
                    g.replay()
                else:
                    self._step_body(pol, nbr_mode, mig)
            else:
                self._step_body(pol, nbr_mode, mig)
        self.steps += 1

    def _capture(self, pol, nbr_mode, mig=False):
        """Capture one step (kernels + both NCCL exchanges) into a CUDA graph; fall back to eager launches
        if the capture is refused."""
        try:
            g = torch.cuda.CUDAGraph()
            self.stream.synchronize()
            with torch.cuda.graph(g, stream=self.stream):
                self._step_body(pol, nbr_mode, mig)
            self._graphs[(pol, nbr_mode, mig)] = g
            return g
        except Exception as e:   # pragma: no cover - depends on the NCCL / driver combination
            import warnings
            warnings.warn('CUDA graph capture of the sharded step failed (%s); using eager launches' % (e,))
            self._graphs[(pol, nbr_mode, mig)] = False
            return False

    def warm_up(self, policy, nbr_mode=capi.NBR_KNN):
        """Load every kernel of the step (and every torch op migrate() uses) before any NCCL traffic: lazy
        module loading while a peer spins inside a collective is a classic deadlock.  State is restored."""
        sw = self.sw
        pol = capi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        keep = {k: v.clone() for k, v in sw.t.items()}
        with torch.cuda.stream(self.stream):
            s = self._stream()
            if self.world > 1:
                sw.h.halo_pack(self.own_cap, self.plan.axis, -1e300, 1e300, self.plan.width, self.park,
                               self.send_lo.data_ptr(), self.send_hi.data_ptr(), self.cap, s)
                sw.h.halo_unpack(self.recv.data_ptr(), 2 * self.cap, self.own_cap, s)
                sw.h.halo_pack_pos(self.psend_lo.data_ptr(), self.psend_hi.data_ptr(), self.cap, self.park, s)
                sw.h.halo_unpack_pos(self.precv.data_ptr(), 2 * self.cap, self.own_cap, s)
                lo, cell, dim = self._park_geom
                fc = self.free_count.clone()
                sw.h.migrate_pack(self.own_cap, self.plan.axis, -1e300, 1e300, (lo[0], lo[1], cell, self.park[2]), int(dim[0]), int(dim[1]),
                                  self.msend_lo.data_ptr(), self.msend_hi.data_ptr(), self.mig_cap, self.free_list.data_ptr(),
                                  self.free_count.data_ptr(), self.mig_cnt.data_ptr(), s)
                sw.h.migrate_unpack(self.mrecv_lo.data_ptr(), self.mrecv_hi.data_ptr(), self.mig_cap, 0, 0,
                                    self.free_list.data_ptr(), self.free_count.data_ptr(), self.mig_cnt.data_ptr(), s)
                self.free_count.copy_(fc)
            sw.h.build_grid(s)
            sw.h.neighbors(nbr_mode, s)
            sw.h.policy_step(pol, s)
            sw.h.status_step(s)
            self.stream.synchronize()
            for k, v in keep.items():
                sw.t[k].copy_(v)
            sw.h.reset_counters(s)
            t = sw.t
            self._fill_parked()
            self.stream.synchronize()

    # ------------------------------------------------------------------ state outside the 128-byte migration row
    def _enable_extra(self):
        """(Re)size the extra migration buffers: one row of scab_migrate_extra_row_doubles per migrated agent and face."""
        if self.world == 1:
            return
        if 2 * self.mig_cap > self.n_rows:
            raise capi.ScabError('sharded Dubins / waypoint state needs 2 * mig_cap <= rows per rank (%d > %d)' % (2 * self.mig_cap, self.n_rows))
        row = self.sw.h.migrate_extra_row_doubles()
        mk = lambda: torch.zeros((self.mig_cap, row), dtype=torch.float64, device=self.sw.device)
        self._graphs.clear()            # captured steps do not know about the extra exchange
        if self.transport == 'nccl':
            self._extra = {'send_lo': mk(), 'send_hi': mk(), 'recv_lo': mk(), 'recv_hi': mk(), 'row': row}
            return
        # peer-to-peer: a second neighbour-writable arena [rows from below | rows from above] (collective for 'p2p')
        half = (self.mig_cap * row * 8 + 255) // 256 * 256
        self._xoff = {'recv_lo': 0, 'recv_hi': half}
        self._xarena, handle = self.sw.h.p2p_alloc(2 * half, want_handle=self.transport == 'p2p')
        dev = self.sw.device
        self._extra = {'send_lo': mk(), 'send_hi': mk(), 'row': row,
                       'recv_lo': _view(self._xarena, (self.mig_cap, row), torch.float64, dev),
                       'recv_hi': _view(self._xarena + half, (self.mig_cap, row), torch.float64, dev)}
        if self.transport == 'p2p':
            peers = self._open_neighbours(handle)
            if peers is None:
                raise capi.ScabError('the extra migration buffers could not be shared between the ranks (the main ones could)')
            self._xpeer = {'lo': peers[0], 'hi': peers[1]}

    def enable_dubins(self, goal_heading, policy='sca'):
        """SCA / RVO3D+Dubins on several GPUs: V_des from the device-side Dubins tracker (Swarm.enable_dubins); the tracker
        state of an agent migrates with it.  `goal_heading` [n_global, 3] = goal_pos[3:6].  Call before the first step
        (row r still holds the agent it was given at construction).  Collective-free."""
        assert self.steps == 0, 'enable_dubins must be called before the first step'
        gh = np.zeros((self.n_rows, 3))
        gh[:len(self._mine)] = np.asarray(goal_heading, 'float64')[self._mine]
        self.sw.enable_dubins(gh)
        self.warm_up(policy)                     # load the tracker's kernels before any NCCL traffic ...
        with torch.cuda.stream(self.stream):
            self.sw.h.dubins_enable(True, self._stream())      # ... and start from fresh trackers again
        self.stream.synchronize()
        self._enable_extra()
        return self

    def set_waypoints(self, waypoints, counts=None):
        """agent.path per GLOBAL agent ([n_global, P, 3], counts [n_global]); the lists and policy.now_goal migrate with
        their agent.  Call before the first step."""
        assert self.steps == 0, 'set_waypoints must be called before the first step'
        wp = np.asarray(waypoints, 'float64')
        cnt = np.full(len(wp), wp.shape[1], np.int32) if counts is None else np.asarray(counts, np.int32)
        loc = np.zeros((self.n_rows,) + wp.shape[1:])
        loc[:len(self._mine)] = wp[self._mine]
        lc = np.zeros(self.n_rows, np.int32)
        lc[:len(self._mine)] = cnt[self._mine]
        n_live = self.sw.n
        self.sw.n = self.n_rows                   # Swarm.set_waypoints sizes its tensors from capacity and copies n rows
        self.sw.set_waypoints(loc, lc)
        self.sw._bind(self.n_rows)
        self._enable_extra()
        return self

    # ------------------------------------------------------------------ diagnostics
    def check_overflow(self):
        """Blocking: raise if a halo / migration buffer ever overflowed or an arrival found no free row."""
        with torch.cuda.stream(self.stream):
            c_lo, c_hi = self.sw.h.halo_counts(self._stream())
        m = self.mig_cnt.cpu().numpy()
        if max(c_lo, c_hi) > self.cap or m[2] or m[3]:
            raise capi.ScabError('sharding buffers overflowed: halo %d,%d (cap %d); migration deferred %d, lost %d'
                                 % (c_lo, c_hi, self.cap, m[2], m[3]))
        return {'halo': (c_lo, c_hi), 'migrated_last': (int(m[0]), int(m[1]))}

    def close(self):
        """Drop the captured graphs (they hold NCCL work; destroy_process_group would wait on them forever)."""
        self.stream.synchronize()
        self._graphs.clear()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ results
    def counters(self, check=True):
        """This rank's counters (blocking).  Raises ScabError if a halo / migration buffer ever overflowed, an arrival was
        lost or an agent outran the motion bound (sticky device flags: nothing is dropped silently)."""
        with torch.cuda.stream(self.stream):
            return self.sw.h.read_counters(self._stream(), check)

    def global_counters(self):
        """Whole-swarm counters: one all-reduce (SUM) of the per-rank device counters (the done test).  The
        overflow masks are OR-ed (as a per-bit sum), and every rank raises together if any rank's is set."""
        c = self.counters(check=False)
        keys = sorted(k for k, v in c.items() if isinstance(v, (int, np.integer)) and k != 'overflow')
        ovf = [(int(c['overflow']) >> b) & 1 for b in range(6)]
        if self.world > 1:
            t = torch.tensor([int(c[k]) for k in keys] + ovf, dtype=torch.int64, device=self.sw.device)
            torch.distributed.all_reduce(t)
            vals = t.tolist()
        else:
            vals = [int(c[k]) for k in keys] + ovf
        out = {k: int(v) for k, v in zip(keys, vals)}
        out['overflow'] = sum((1 << b) for b in range(6) if vals[len(keys) + b])
        if out['overflow']:
            raise capi.ScabError('sharded run overflow flags %d: %s' % (out['overflow'], '; '.join(
                t for b, t in capi.OVF_TEXT.items() if out['overflow'] & b)))
        return out

    def is_done(self):
        """MACAEnv.is_done (mampenv.py:51-59) over all slabs: every owned agent of every rank has finished."""
        return self.global_counters()['n_not_done'] == 0

    def gather_state(self, keys=('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')):
        """Owned rows keyed by global id (host)."""
        self.stream.synchronize()
        torch.cuda.synchronize()
        t = self.sw.t
        own = torch.nonzero((t['status'][:self.own_cap] & capi.ST_GHOST) == 0)[:, 0]
        self.n_own = int(own.numel())
        return t['gid'][own].cpu().numpy(), {k: t[k][own].cpu().numpy() for k in keys}


class LogicalShards:
    """S slabs of one swarm as S ShardedSwarm objects in ONE process on ONE device: the same halo / migration kernels,
    the same row layout and the same step sequence as the one-process-per-GPU run, with the NCCL send/recv of every
    exchange point replaced by a device copy between the neighbours' buffers (all on one stream).  SURVEY 4(iii): the
    sharded path can be held to "bit for bit equal to the single-device run" on a box with a single GPU."""

    def __init__(self, pos6, goal6, shards, device, lo, hi, **kw):
        self.world = int(shards)
        kw.setdefault('use_graph', False)
        kw.setdefault('transport', 'local')      # 'local': direct writes into the neighbour shard's buffers; 'nccl': staged copies
        self.shards = [ShardedSwarm(pos6, goal6, r, self.world, device, lo, hi, **kw) for r in range(self.world)]
        self.stream = self.shards[0].stream
        for sh in self.shards:
            sh.stream = self.stream
        self._wire()
        self.n_global = self.shards[0].n_global

    def _wire(self):
        """'local' transport: the neighbours' arenas are ordinary pointers of this process."""
        for r, sh in enumerate(self.shards):
            if sh.transport != 'local':
                continue
            lo, hi = (self.shards[r - 1] if r > 0 else None), (self.shards[r + 1] if r < self.world - 1 else None)
            sh._peer = {'lo': lo._arena if lo else 0, 'hi': hi._arena if hi else 0}
            sh._xpeer = {'lo': lo._xarena if lo else 0, 'hi': hi._xarena if hi else 0}

    def enable_dubins(self, goal_heading, policy='sca'):
        for sh in self.shards:
            sh.enable_dubins(goal_heading, policy)
        self._wire()
        return self

    def set_waypoints(self, waypoints, counts=None):
        for sh in self.shards:
            sh.set_waypoints(waypoints, counts)
        self._wire()
        return self

    def step(self, policy, nbr_mode=capi.NBR_KNN):
        pol = capi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        S = self.world
        with torch.cuda.stream(self.stream):
            mig = self.shards[0]._wants_migration()
            gens = [sh._step_phases(pol, nbr_mode, mig) for sh in self.shards]
            while True:
                reqs = [next(g, None) for g in gens]          # every shard runs up to its next exchange point
                if all(r is None for r in reqs):
                    break
                assert all(r is not None for r in reqs), 'logical shards fell out of lockstep'
                for r in range(S):
                    if isinstance(reqs[r], int):              # direct-write transport: the data is already there
                        continue
                    s_lo, s_hi = reqs[r][0], reqs[r][1]
                    if r > 0:
                        reqs[r - 1][3].copy_(s_lo)            # the lower neighbour's recv_hi
                    if r < S - 1:
                        reqs[r + 1][2].copy_(s_hi)            # the upper neighbour's recv_lo
        for sh in self.shards:
            sh.steps += 1

    def check_overflow(self):
        return [sh.check_overflow() for sh in self.shards]

    def counters(self):
        """Whole-swarm counters: the sum over the shards (what ShardedSwarm.global_counters all-reduces)."""
        tot = {}
        for sh in self.shards:
            for k, v in sh.counters().items():
                if isinstance(v, (int, np.integer)):
                    tot[k] = tot.get(k, 0) + int(v)
        return tot

    def is_done(self):
        return self.counters()['n_not_done'] == 0

    def gather_state(self, keys=('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')):
        """Owned rows of every shard keyed by global id (host); asserts that every agent has exactly one owner."""
        gid = np.zeros(0, np.int64)
        out = {k: [] for k in keys}
        for sh in self.shards:
            g, st = sh.gather_state(keys)
            gid = np.concatenate([gid, g.astype(np.int64)])
            for k in keys:
                out[k].append(st[k])
        assert len(np.unique(gid)) == len(gid) == self.n_global, 'an agent is owned by %s shards' % (
            'no' if len(gid) < self.n_global else 'two')
        order = np.argsort(gid)
        return {k: np.concatenate(v)[order] for k, v in out.items()}

    def close(self):
        for sh in self.shards:
            sh.close()
