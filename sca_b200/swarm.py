"""Device-resident swarm state (structure of arrays, owned through PyTorch) + the batched step.

Replaces the array-of-objects state of mamp/agents/agent.py:13-74 and the per-agent loop of
MACAEnv._take_action (mamp/envs/mampenv.py:27-49).  PyTorch is plumbing here (device memory, streams,
pinned host staging); all compute is libscab200's CUDA kernels, reached through the C ABI.
"""
import ctypes as C

import numpy as np
import torch

from . import capi

_F8, _F4, _I4 = torch.float64, torch.float32, torch.int32

# name -> (dtype, trailing shape, optional/debug?)
_FIELDS = {
    'pos': (_F8, (3,), False), 'vel': (_F4, (3,), False), 'heading': (_F8, (3,), False), 'goal': (_F8, (3,), False),
    'radius': (_F8, (), False), 'pref_speed': (_F8, (), False), 'status': (_I4, (), False),
    'total_dist': (_F8, (), False), 'step_num': (_I4, (), False), 'max_run_dist': (_F8, (), False),
    'vpref_ext': (_F8, (3,), True), 'vpref_out': (_F8, (3,), False), 'action': (_F4, (7,), False),
    'nbr_cnt': (_I4, (), True), 'nbr_ids': (_I4, (capi.K,), True), 'nbr_dsq': (_F8, (capi.K,), True),
    'tag': (_I4, (), True), 'n_suit': (_I4, (), True), 'plane_fail': (_I4, (), True), 'n_planes': (_I4, (), True),
    'lp4_proj': (_I4, (), True),
}
_DEBUG = ('nbr_cnt', 'nbr_ids', 'nbr_dsq', 'tag', 'n_suit', 'plane_fail', 'n_planes', 'lp4_proj')
_UPLOAD = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')


class Swarm:
    """N agents resident in HBM.  `debug=True` also allocates the observability outputs
    (ordered neighbour lists, branch tags, plane counts) the parity tests read back."""

    def __init__(self, n, device='cuda:0', debug=False, params=None, capacity=None):
        if not torch.cuda.is_available():
            raise capi.ScabError('sca_b200.Swarm needs a CUDA device; there is no CPU fallback')
        self.n = int(n)
        self.capacity = int(capacity or max(self.n, 1))
        self.device = torch.device(device)
        self.debug = bool(debug)
        self.params = params if params is not None else capi.default_params()
        self.t = {}
        for k, (dt, shp, opt) in _FIELDS.items():
            if k == 'vpref_ext' or (k in _DEBUG and not self.debug):
                continue
            self.t[k] = torch.zeros((self.capacity,) + shp, dtype=dt, device=self.device)
        self.h = capi.Handle(self.params, self.device.index or 0, self.capacity)
        self.m = 0
        self._pinned = {}
        self._bind()

    # ------------------------------------------------------------------ binding
    def _bind(self, n=None):
        """(Re)bind the SoA pointers; `n` <= capacity is the number of live rows (owned + halo ghosts)."""
        if n is not None:
            self.n = int(n)
        p = capi.StatePtrs()
        for k in capi.STATE_FIELDS:
            tns = self.t.get(k)
            setattr(p, k, tns.data_ptr() if tns is not None else None)
        p.n_global = int(getattr(self, 'n_global', 0))
        gh = self.t.get('goal_heading')
        p.goal_heading = gh.data_ptr() if gh is not None else None
        wp = self.t.get('waypoints')
        if wp is not None:
            p.waypoints, p.waypoint_left, p.now_goal = wp.data_ptr(), self.t['waypoint_left'].data_ptr(), self.t['now_goal'].data_ptr()
            p.waypoint_stride = int(wp.shape[1])
        self.h.bind_state(p, self.n)

    def set_global_ids(self, gid, n_global):
        """Sharded runs: global id of every row (ties and the collision-check order follow it)."""
        if 'gid' not in self.t:
            self.t['gid'] = torch.zeros(self.capacity, dtype=_I4, device=self.device)
        g = torch.as_tensor(gid, dtype=_I4)
        self.t['gid'][:g.numel()].copy_(g)
        self.n_global = int(n_global)
        self._bind()

    def __getattr__(self, k):
        t = self.__dict__.get('t')
        if t is not None and k in t:
            return t[k]
        raise AttributeError(k)

    def set_external_vpref(self, v):
        """Supply V_des from an upstream provider (SCA's Dubins tracker, scaPolicy.py:264-338); None ->
        straight-line v_pref computed on the device (rvo3dPolicy.py:182-196)."""
        if v is None:
            if 'vpref_ext' in self.t:
                del self.t['vpref_ext']
                self._bind()
            return
        v = torch.as_tensor(np.ascontiguousarray(v, dtype='float64')) if not torch.is_tensor(v) else v
        if 'vpref_ext' not in self.t:
            self.t['vpref_ext'] = torch.zeros((self.capacity, 3), dtype=_F8, device=self.device)
            self._bind()
        self.t['vpref_ext'][:v.shape[0]].copy_(v.to(self.device, _F8))

    def enable_dubins(self, goal_heading=None, on=True):
        """SCA / RVO3D+Dubins: take V_des from the on-device Dubins tracker (scaPolicy.py:264-338) instead of an
        external provider.  `goal_heading` [n,3] = goal_pos[3:6] (agent.py:21); resets every tracker."""
        if on:
            if 'goal_heading' not in self.t:
                self.t['goal_heading'] = torch.zeros((self.capacity, 3), dtype=_F8, device=self.device)
            self.t['goal_heading'].zero_()
            if goal_heading is not None:
                g = torch.as_tensor(np.ascontiguousarray(goal_heading, dtype='float64'))
                self.t['goal_heading'][:g.shape[0]].copy_(g)
            self.set_external_vpref(None)
            self._bind()
        self.h.dubins_enable(on, self._stream())
        return self

    def set_waypoints(self, waypoints, counts=None):
        """agent.path of an external global planner (agent.py:44) for RVO3D / S-RVO3D / ORCA: `waypoints` [n, P, 3] in the
        reference's list order (the policies pop() from the END, so the first waypoint to fly to is stored last),
        `counts` [n] = len(agent.path) (default P).  Resets policy.now_goal to None.  None removes the paths."""
        if waypoints is None:
            for k in ('waypoints', 'waypoint_left', 'now_goal'):
                self.t.pop(k, None)
            self._bind()
            return self
        wp = np.nan_to_num(np.ascontiguousarray(waypoints, dtype='float64'))
        n, P = wp.shape[0], max(1, wp.shape[1])
        cnt = np.full(n, wp.shape[1], np.int32) if counts is None else np.ascontiguousarray(counts, np.int32)
        self.t['waypoints'] = torch.zeros((self.capacity, P, 3), dtype=_F8, device=self.device)
        self.t['waypoints'][:n, :wp.shape[1]].copy_(torch.from_numpy(wp))
        self.t['waypoint_left'] = torch.zeros(self.capacity, dtype=_I4, device=self.device)
        self.t['waypoint_left'][:n].copy_(torch.from_numpy(cnt))
        self.t['now_goal'] = torch.full((self.capacity, 3), float('nan'), dtype=_F8, device=self.device)
        self._bind()
        return self

    def dubins_states(self):
        return self.h.dubins_read(self.n, self._stream())

    # ------------------------------------------------------------------ state in / out
    def set_state(self, pos, goal, heading=None, vel=None, radius=0.5, pref_speed=1.0, status=None,
                  total_dist=None, step_num=None, max_run_dist=None, initial_pos=None, goal_pos=None):
        n = self.n
        f8 = lambda a, shp: np.ascontiguousarray(np.broadcast_to(np.asarray(a, 'float64'), shp))
        pos, goal = f8(pos, (n, 3)), f8(goal, (n, 3))
        start = pos if initial_pos is None else np.asarray(initial_pos, 'float64')[:, :3]
        gl = goal if goal_pos is None else np.asarray(goal_pos, 'float64')[:, :3]
        if max_run_dist is None:   # agent.py:74
            max_run_dist = 3.0 * np.round(np.sqrt(((start - gl) ** 2).sum(1)), 5)
        dd = gl - start            # scaPolicy.py:188-189 is_zAxis
        zaxis = ((np.abs(dd[:, 0]) <= 1e-5) & (np.abs(dd[:, 1]) <= 1e-5)).astype(np.int32) * capi.ST_ZAXIS
        st = np.zeros(n, np.int32) if status is None else (np.asarray(status, np.int32) & 15)
        host = {
            'pos': pos, 'goal': goal, 'heading': f8(np.zeros((n, 3)) if heading is None else heading, (n, 3)),
            'vel': np.zeros((n, 3), np.float32) if vel is None else np.ascontiguousarray(vel, np.float32),
            'radius': f8(radius, (n,)), 'pref_speed': f8(pref_speed, (n,)), 'status': st | zaxis,
            'total_dist': f8(0.0 if total_dist is None else total_dist, (n,)),
            'step_num': np.zeros(n, np.int32) if step_num is None else np.ascontiguousarray(step_num, np.int32),
            'max_run_dist': f8(max_run_dist, (n,)),
        }
        self.h.set_max_radius(float(np.max(host['radius'])) if n else 0.0)
        # largest speed a policy can command: max_speed (the LP), pref_speed (outer candidate shell, |v_pref|), 0.5 (inner
        # shell).  The collision stage's reach is derived from it; a move beyond it raises the sticky OVF_MOTION flag.
        self.v_bound = max(float(self.params.max_speed), float(np.max(host['pref_speed'])) if n else 0.0, 0.5)
        self.h.set_motion_bound(self.v_bound)
        self._zaxis = torch.zeros(self.capacity, dtype=_I4, device=self.device)
        self._zaxis[:n].copy_(torch.as_tensor(zaxis))
        for k, a in host.items():
            self.t[k][:n].copy_(torch.from_numpy(np.array(a, copy=True)))
        return self

    def set_obstacles(self, xyzr):
        xyzr = np.zeros((0, 4)) if xyzr is None else np.asarray(xyzr, 'float64').reshape(-1, 4)
        self.m = len(xyzr)
        self.h.set_obstacles(xyzr)

    def fit_grid(self, lo=None, hi=None, margin=5.0, cell=None):
        """Size the uniform grid from a bounding box (default: current positions and goals)."""
        cell = float(cell or max(self.params.cell_size, self.params.neighbor_dist))
        if lo is None or hi is None:
            both = torch.cat([self.t['pos'][:self.n], self.t['goal'][:self.n]])
            lo = (both.min(0).values - margin).cpu().numpy()
            hi = (both.max(0).values + margin).cpu().numpy()
        lo, hi = np.asarray(lo, 'float64'), np.asarray(hi, 'float64')
        dim = np.maximum(1, np.ceil((hi - lo) / cell).astype(np.int64))
        while int(dim.prod()) > (1 << 24):     # keep the cell table modest; border cells absorb the rest
            cell *= 1.26
            dim = np.maximum(1, np.ceil((hi - lo) / cell).astype(np.int64))
        self.h.set_grid(lo, cell, dim)
        self.grid = (lo, cell, dim)
        return self

    def upload(self, host, stream=None):
        """Pinned-host -> device copy of the per-step inputs (pos, vel, heading, status, ...)."""
        nbytes = 0
        for k in _UPLOAD + (('vpref_ext',) if 'vpref_ext' in host else ()):
            if k not in host:
                continue
            src = host[k]
            if k == 'status':
                src = src  # caller supplies SCAB_ST_* bits; the static ZAXIS bit is OR-ed below
            self.t[k][:src.shape[0]].copy_(src, non_blocking=True)
            nbytes += src.numel() * src.element_size()
        if 'status' in host:
            self.t['status'].bitwise_or_(self._zaxis)
        return nbytes

    def download(self, keys=('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num', 'action')):
        return {k: self.t[k][:self.n].cpu().numpy() for k in keys if k in self.t}

    # ------------------------------------------------------------------ host-resident state: one step end to end
    HOST_STATE = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')

    def host_buffers(self):
        """Pinned host copies of the per-step state (+ the float32 action matrix the reference env consumes,
        mampenv.py:31), initialised from the device.  The dict step_from_host() works on."""
        n = self.n
        host = {k: self.t[k][:n].to('cpu', copy=True).pin_memory() for k in self.HOST_STATE}
        host['action'] = torch.zeros((n, 7), dtype=_F4).pin_memory()
        return host

    def step_from_host(self, host, policy, nbr_mode=capi.NBR_KNN, overlap=True, upload_vel=False, chunks=8, min_chunk_rows=65536):
        """One MACAEnv.step for a caller whose state lives in HOST memory (what a mamp/envs drop-in pays per step): upload the
        step's inputs from the pinned buffers of host_buffers(), run the four stages, download the new state and the action
        matrix into the same buffers, and return once they are there.  Bytes per agent: 64 up (pos, heading, status,
        total_dist, step_num), 92 down (the same + action).  The velocity is NOT moved in either direction unless
        upload_vel=True: update_velocitie sets it to action[:3] (mampenv.py:104) and the device did exactly that in the
        previous step, so its copy is current; host['vel'] is refreshed from the action rows on the host side lazily
        by host_velocity().  overlap=True hides the part of the upload the cell list does not read behind the cell list +
        16-NN (second stream), issues the policy stage as `chunks` row ranges (scab_policy_step_rows) so that the rows of a
        range come down while the next range computes (ranges of at least `min_chunk_rows` rows: below that the extra launches
        cost more than the overlap gives), and downloads what k_status writes behind it; the result is
        bit-identical to overlap=False (tests/test_gpu_step_from_host.py)."""
        pol = capi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        n = self.n
        t = self.t
        first = ('pos', 'status') + (('vel',) if upload_vel else ())
        late = ('heading', 'total_dist', 'step_num')
        early_down, late_down = ('pos', 'heading', 'total_dist', 'action'), ('status', 'step_num')
        main_s = torch.cuda.current_stream()
        if not overlap:
            for k in first + late:
                t[k][:n].copy_(host[k], non_blocking=True)
            t['status'].bitwise_or_(self._zaxis)
            if nbr_mode == capi.NBR_REF_COMPAT:
                self.h.kdtree_build(host['pos'].numpy())
            self.h.step(pol, nbr_mode, self._stream())
            for k in early_down + late_down:
                host[k].copy_(t[k][:n], non_blocking=True)
            main_s.synchronize()
            return host
        cs = self.__dict__.get('_copy_stream')
        if cs is None:
            cs = self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ev = [torch.cuda.Event() for _ in range(3)]
        ev_up, ev_in, ev_pol = self._ev[:3]
        for k in first:
            t[k][:n].copy_(host[k], non_blocking=True)
        t['status'].bitwise_or_(self._zaxis)
        ev_up.record(main_s)
        with torch.cuda.stream(cs):
            cs.wait_event(ev_up)              # (keeps the copy stream behind the previous step's downloads too)
            for k in late:
                t[k][:n].copy_(host[k], non_blocking=True)
            ev_in.record(cs)
        if nbr_mode == capi.NBR_REF_COMPAT:
            self.h.kdtree_build(host['pos'].numpy())
        self.h.build_grid(self._stream())
        self.h.neighbors(nbr_mode, self._stream())
        main_s.wait_event(ev_in)
        nch = max(1, min(int(chunks), n // max(1, int(min_chunk_rows)))) if 'goal_heading' not in t or 'vpref_ext' in t or pol not in (0, 3) else 1
        if nch <= 1:          # (also: the device Dubins tracker runs for every agent at once)
            self.h.policy_step(pol, self._stream())
            ev_pol.record(main_s)
            with torch.cuda.stream(cs):
                cs.wait_event(ev_pol)
                for k in early_down:
                    host[k].copy_(t[k][:n], non_blocking=True)
        else:
            while len(self._ev) < 3 + nch:
                self._ev.append(torch.cuda.Event())
            for c in range(nch):
                a, b = c * n // nch, (c + 1) * n // nch
                self.h.policy_step_rows(pol, a, b, self._stream())
                ev_c = self._ev[3 + c]
                ev_c.record(main_s)
                with torch.cuda.stream(cs):
                    cs.wait_event(ev_c)
                    for k in early_down:
                        host[k][a:b].copy_(t[k][a:b], non_blocking=True)
        self.h.status_step(self._stream())
        for k in late_down:
            host[k].copy_(t[k][:n], non_blocking=True)
        cs.synchronize()
        main_s.synchronize()                  # the caller needs the results on the host before its next step
        return host

    @staticmethod
    def host_velocity(host):
        """vel_global_frame after the step = action[:3] (mampenv.py:104), as a host view of the downloaded action rows."""
        return host['action'][:, :3]

    H2D_BYTES_PER_AGENT = 24 + 24 + 4 + 8 + 4
    D2H_BYTES_PER_AGENT = 24 + 24 + 4 + 8 + 4 + 28

    # ------------------------------------------------------------------ the step
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def step(self, policy, nbr_mode=capi.NBR_KNN):
        """One MACAEnv.step (mampenv.py:22-25) for every agent, asynchronously on the current stream."""
        pol = capi.POLICIES[policy] if isinstance(policy, str) else int(policy)
        if nbr_mode == capi.NBR_REF_COMPAT:
            self.h.kdtree_build(self.t['pos'].cpu().numpy())
        self.h.step(pol, nbr_mode, self._stream())

    def stage_grid(self):
        self.h.build_grid(self._stream())

    def stage_neighbors(self, nbr_mode=capi.NBR_KNN, pos_host=None):
        if nbr_mode == capi.NBR_REF_COMPAT:
            self.h.kdtree_build(self.t['pos'].cpu().numpy() if pos_host is None else pos_host)
        self.h.neighbors(nbr_mode, self._stream())

    def stage_policy(self, policy):
        self.h.policy_step(capi.POLICIES[policy] if isinstance(policy, str) else int(policy), self._stream())

    def stage_status(self):
        self.h.status_step(self._stream())

    def counters(self, check=True):
        """Step counters (blocking).  Raises ScabError if the sticky overflow mask is set (results may be wrong)."""
        return self.h.read_counters(self._stream(), check)

    def is_done(self):
        return self.counters()['n_not_done'] == 0

    def metrics(self, straight_len=None, desire_steps=None, dt=0.1):
        """Rollout metrics of run_example/run_sca.py:228-250 from the device counters."""
        st = self.t['status'].cpu().numpy()
        ok = ((st & (capi.ST_COLLISION | capi.ST_TIMEOUT)) == 0)
        td = self.t['total_dist'].cpu().numpy()
        sn = self.t['step_num'].cpu().numpy()
        ns = int(ok.sum())
        out = {'SuccessRate': ns / max(self.n, 1), 'successful_num': ns}
        if straight_len is not None and desire_steps is not None and ns:
            out['ExtraTime'] = float(((sn[ok].sum() - np.asarray(desire_steps)[ok].sum()) * dt) / ns)
            out['ExtraDistance'] = float((td[ok].sum() - np.asarray(straight_len)[ok].sum()) / ns)
        if ns and sn[ok].sum():
            out['AverageSpeed'] = float(td[ok].sum() / sn[ok].sum() / dt)
        return out
