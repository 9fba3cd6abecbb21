#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched collision-avoidance step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5] [--policy P] [--agents A]
                    [--skip-steps S] [--quick] [--no-cpu]

Default workload (config.workload): BASELINE.json configs[4], "SCA synthetic uniform-random swarm, 1,000,000 UAVs"
(pos/goal ~ U(0, 560 m)^3, >= 1.1 m initial separation, r = 0.5, pref_speed = 1, straight-line v_pref --
the Dubins provider is upstream of the path).  It fits one B200, so N = 1 runs all 10^6
agents on one GPU; N > 1 shards the SAME 10^6 agents into spatial slabs with a halo exchange (strong scaling).

`--config 1..4` run BASELINE configs[0..3] at their named sizes on one GPU (CONFIGS below; configs 1 and 2 with the
reference's own neighbour semantics, REF_COMPAT).  KNN mode -- the true 16 nearest -- is what the default workload uses; it is
NOT the reference's traversal-order neighbour list and the line says so.

A "step" is one full MACAEnv.step for every agent: cell list -> 16-NN -> SCA candidate search ->
kinematic update -> collision/goal status.  value = agent-steps/s with state resident in HBM;
e2e = the same through host buffers and the repo's public API for that case, Swarm.step_from_host (pinned H2D of the step's
inputs + D2H of its results inside the timed region), which is what a caller with host-resident state pays per step.
`--impl reference` times the CPU oracle port of the reference (the reference is ~40 agent-steps/s of
single-threaded Python and cannot run this workload) on a bounded sample with every host thread.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DENSITY = 1.0e6 / 560.0 ** 3          # agents per m^3 of BASELINE config 5
ALGO_BYTES = {'grid': 168, 'neighbors': 116, 'policy': 660, 'status': 144}   # SURVEY.md 8(d), per agent-step
# this library's kernels launched per step (profiles/r2_launches_{sca,lp}.csv): k_cell_count, k_scan_tile_sums, k_scan_tiles,
# k_scan_apply, k_reorder, k_neighbors_direct<pass 1>, <pass 2>, then k_vpref, k_select_fast<pass 1>, <pass 2>, k_select,
# k_integrate (sampled policies) or k_policy_lp<5>, k_policy_lp<4> (ORCA LP), then k_status; a sharded step adds the halo pack /
# signal / unpack kernels of its two exchanges (6; the migration kernels of every 8th step are not counted)
KERNELS_PER_STEP = {'orca_lp': 10}
KERNELS_PER_STEP_SAMPLED = 13
POLICY_KERNELS = {'orca_lp': 'k_policy_lp'}
METRIC = 'agent-steps/sec (SCA, 1M UAVs)'      # --policy other than sca: the same workload under that policy (named in config)

# BASELINE.json configs[0..4] as --config 1..5 (5 = the headline workload and the default).  `cpu`: the bounded sample the
# CPU legs run (the whole scenario where the oracle finishes it in seconds).
CONFIGS = {
    1: dict(workload='BASELINE configs[0]: run_sca.py circle scenario (R = 18 m), 100 fixed-wing UAVs, SCA with its Dubins v_pref '
                     'tracker on the device, REF_COMPAT neighbour lists (the reference\'s own kd-tree order)',
            policy='sca', nbr='compat', dubins=True, gen=lambda sc: sc.circle(100, 18.0), cpu=lambda sc: sc.circle(100, 18.0)),
    2: dict(workload='BASELINE configs[1]: run_orca.py random box (r = 15 m, seed 0), 1,000 agents, ORCA3D with linearProgram3 / '
                     'linearProgram4 (orca3dPolicyOfficial), REF_COMPAT neighbour lists',
            policy='orca_lp', nbr='compat', dubins=False, gen=lambda sc: sc.random_box(1000, 15.0, seed=0),
            cpu=lambda sc: sc.random_box(1000, 15.0, seed=0)),
    3: dict(workload='BASELINE configs[2]: SCA take-off / landing pads (run_sca.py:53-81 + the 8 obstacles of :139-150) tiled 25 x 25 at '
                     '30 m pitch: 10,000 UAVs + 5,000 obstacles, posture constraint, Dubins tracker on the device, KNN neighbour mode',
            policy='sca', nbr='knn', dubins=True, gen=lambda sc: sc.takeoff_landing_tiled(625), cpu=lambda sc: sc.takeoff_landing_tiled(64)),
    4: dict(workload='BASELINE configs[3]: dense Fibonacci-sphere swap (run_orca.py:36-54 scaled, R = 198 m), 100,000 agents, KNN neighbour mode',
            policy='rvo3d', nbr='knn', dubins=False, gen=lambda sc: sc.dense_sphere(100000), cpu=lambda sc: sc.dense_sphere(100000)),
}


def config_of(args):
    """-> (cfg dict or None for the default workload, policy, neighbour mode name)."""
    if args.config == 5:
        return None, args.policy or 'sca', 'knn'
    cfg = CONFIGS[args.config]
    pol = args.policy or cfg['policy']
    if args.config == 4 and pol not in ('rvo3d', 'srvo3d'):
        raise SystemExit('--config 4 is the RVO3D / S-RVO3D workload: --policy rvo3d|srvo3d')
    return cfg, pol, cfg['nbr']


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.p = None
        self.nvml = None
        self._stop = False
        try:        # NVML directly: a reading every 2 ms from the first moment on (a sharded timed region lasts ~20 ms, less than
            import pynvml                      # nvidia-smi needs to print its first line)
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(index)))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50',
                                       '-i', str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _poll(self):
        nv, h = self.nvml
        bits = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'), (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = ''
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                f = ['Active' if r & b else 'Not Active' for b, _ in bits]
                self.rows.append((time.time(), '%s, %s, , %s' % (sm, mx, ', '.join(f))))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.p is None and self.nvml is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        if self.nvml is not None:
            self._stop = True
            self.t.join(0.5)
        else:
            time.sleep(0.12)
            self.p.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        pad = (0.0, 0.0) if self.nvml is not None else (0.05, 0.1)        # (nvidia-smi's 50 ms period needs the slack)
        rows = [r for (t, r) in self.rows if t0 - pad[0] <= t <= t1 + pad[1]] or [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for nm, v in zip(names, f[3:7]):
                    if v == 'Active':
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_config_swarm(cfg, device):
    """BASELINE configs[0..3]: the reference's own scenario geometry on one GPU."""
    from sca_b200 import scenarios
    from sca_b200.swarm import Swarm
    pos, goal, obs = cfg['gen'](scenarios)
    n = len(pos)
    sw = Swarm(n, device=device, debug=False)
    sw.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:6], initial_pos=pos, goal_pos=goal)
    sw.set_obstacles(obs)
    lo = np.minimum(pos[:, :3].min(0), goal[:, :3].min(0)) - 20.0
    hi = np.maximum(pos[:, :3].max(0), goal[:, :3].max(0)) + 20.0
    sw.fit_grid(lo, hi)
    if cfg['dubins']:
        sw.enable_dubins(goal[:, 3:6])
    return sw


class _Agent(object):
    """The attributes of mamp/agents/agent.py:9-77 that the policy interface reads and writes (the reference package is not on
    the GPU box; the shims are duck-typed)."""

    def __init__(self, start, goal, policy, i):
        self.policy = policy()
        self.initial_pos, self.goal_pos = np.array(start, 'float64'), np.array(goal, 'float64')
        self.pos_global_frame, self.goal_global_frame = self.initial_pos[:3].copy(), self.goal_pos[:3].copy()
        self.heading_global_frame = self.initial_pos[3:].copy()
        self.vel_global_frame = np.zeros(3)
        self.radius, self.pref_speed, self.id = 0.5, 1.0, i
        self.neighbors, self.v_pref, self.path = [], np.zeros(3), []
        self.total_time = self.total_dist = 0.0
        self.step_num = 0
        self.is_at_goal = self.is_collision = self.is_out_of_max_time = self.is_obstacle = False
        self.max_run_dist = 3.0 * round(float(np.sqrt(((self.initial_pos[:3] - self.goal_pos[:3]) ** 2).sum())), 5)


def shim_rate(cfg, policy, steps):
    """agent-steps/s of the DROP-IN path (sca_b200.mamp_compat: GPUKDTree.buildAgentTree + find_next_action per agent, the two
    calls MACAEnv._take_action makes, mampenv.py:28,40) on Agent objects, followed by the env's own per-agent Python update
    (mampenv.py:83-105).  What a user of the unmodified reference gets after the one-line swap of INTEGRATION.md."""
    import math
    from sca_b200 import mamp_compat, scenarios
    pos, goal, obs = cfg['gen'](scenarios)
    cls = {'sca': mamp_compat.GPUSCAPolicy, 'orca_lp': mamp_compat.GPUORCA3DPolicyOfficial, 'rvo3d': mamp_compat.GPURVO3DPolicy,
           'srvo3d': mamp_compat.GPUSRVO3DPolicy}[policy]
    agents = [_Agent(pos[i], goal[i], cls, i) for i in range(len(pos))]

    class _Obs(object):
        def __init__(self, o, k):
            self.pos_global_frame, self.radius, self.id, self.is_obstacle, self.is_at_goal = np.array(o[:3]), float(o[3]), k, True, True
            self.vel_global_frame = np.zeros(3)
    tree = mamp_compat.GPUKDTree(agents, [_Obs(o, k) for k, o in enumerate(obs)])
    wrap = lambda a: (a + math.pi) % (2 * math.pi) - math.pi

    def one():
        tree.buildAgentTree()                                            # mampenv.py:28
        n_act = 0
        for a in agents:
            if a.is_at_goal or a.is_collision or a.is_out_of_max_time:
                continue
            act = a.policy.find_next_action({}, a, tree)                 # mampenv.py:40
            n_act += 1
            al, be = wrap(a.heading_global_frame[0] + act[4]), wrap(a.heading_global_frame[1] + act[5])     # mampenv.py:83-105
            d = act[3] * 0.1 * np.array([math.cos(be) * math.cos(al), math.cos(be) * math.sin(al), math.sin(be)])
            a.pos_global_frame += d
            a.total_dist += float(np.sqrt((d * d).sum()))
            a.heading_global_frame = [al, be, 0.0]
            a.vel_global_frame = np.array(act[:3])
            a.step_num += 1
        return n_act
    for _ in range(3):
        one()
    t0 = time.perf_counter()
    tot = sum(one() for _ in range(steps))
    dt = time.perf_counter() - t0
    return {'value': tot / dt, 'unit': 'agent-steps/s', 'ms_per_step': dt / steps * 1e3, 'steps': steps,
            'what': 'unmodified-MACAEnv drop-in: GPUKDTree.buildAgentTree + GPU*Policy.find_next_action on Agent objects + the env\'s '
                    'per-agent Python update; the reference\'s own policies run ~40 agent-steps/s on this scenario (BASELINE.md)'}


def make_swarm(n_total, rank, world, device, seed=0):
    from sca_b200 import scenarios
    from sca_b200.swarm import Swarm
    pos, goal, _ = scenarios.uniform_box_torch(n_total, DENSITY, seed=seed, device=device)
    L = (n_total / DENSITY) ** (1.0 / 3.0)
    lo, hi = np.array([-15.0, -15.0, 5.0]), np.array([L + 15.0, L + 15.0, L + 35.0])
    if world == 1:
        sw = Swarm(n_total, device=device, debug=False)
        if os.environ.get('SCAB_BENCH_NO_SORT', '0') != '1':
            # rows laid out cell by cell at t = 0 (ids travel in gid; results are independent of the row order)
            order = scenarios.spatial_order(pos[:, :3])
            pos, goal = pos[order], goal[order]
            sw.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:])
            sw.set_global_ids(order.astype(np.int32), n_total)
        else:
            sw.set_state(pos[:, :3], goal[:, :3], heading=pos[:, 3:])
        sw.fit_grid(lo, hi)
        return sw
    from sca_b200.sharding import ShardedSwarm
    # SCAB_TRANSPORT=nccl: pack -> ncclSend/ncclRecv -> unpack instead of the direct NVLink writes (A/B switch)
    return ShardedSwarm(pos, goal, rank, world, device, lo, hi, transport=os.environ.get('SCAB_TRANSPORT', 'p2p'))


def run_ours(args):
    import torch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = 'cuda:%d' % local
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(device))
    from sca_b200 import capi
    if os.environ.get('SCAB_BENCH_SIDE_STREAM', '0') == '1':      # debug: everything on a non-default (non-blocking) stream
        torch.cuda.set_stream(torch.cuda.Stream())
    cfg, policy, nbr_name = config_of(args)
    NBR = capi.NBR_REF_COMPAT if nbr_name == 'compat' else capi.NBR_KNN
    if cfg is not None and world > 1:
        raise SystemExit('--config 1-4 are single-GPU workloads (below the size where sharding pays off, SURVEY 8e)')
    if cfg is None:
        n_total = args.agents
        sw = make_swarm(n_total, rank, world, device)
    else:
        sw = make_config_swarm(cfg, device)
        n_total = sw.n
    for _ in range(args.skip_steps):          # a later part of the rollout (conflicts have developed), untimed
        sw.step(policy, NBR)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    if world > 1:
        sw.warm_up(policy, NBR)     # load every kernel before the first NCCL exchange
        torch.distributed.barrier()
    tstream = getattr(sw, 'stream', None) or torch.cuda.current_stream()   # the stream the kernels are launched on
    for _ in range(max(args.warmup, 3)):
        sw.step(policy, NBR)
    barrier()
    if os.environ.get('SCAB_BENCH_TRACE'):        # debug: per-step wall times (synchronised), not part of the measurement
        tr = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            sw.step(policy, NBR)
            tstream.synchronize()
            tr.append(round((time.perf_counter() - t0) * 1e3, 2))
        if rank == 0:
            sys.stderr.write('TRACE ms/step: %s\n' % tr)
    if os.environ.get('SCAB_BENCH_PHASES') and world > 1:   # debug: eager per-phase device times of the sharded step
        marks = []
        def mark(what):
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            marks.append((what, e))
        sw.debug_phase = mark
        acc = {}
        for _ in range(args.steps):
            del marks[:]
            sw.step(policy, NBR)
            with torch.cuda.stream(sw.stream):
                mark('end')
            sw.stream.synchronize()
            for (w, e), (_, e2) in zip(marks[:-1], marks[1:]):
                acc[w] = acc.get(w, 0.0) + e.elapsed_time(e2)
        sw.debug_phase = None
        sys.stderr.write('PHASES rank %d ms/step (eager): %s rows=%d own=%d cap=%d\n' % (
            rank, {k: round(v / args.steps, 4) for k, v in acc.items()}, sw.n_rows, sw.n_own, sw.cap))
        barrier()
    c0 = sw.counters()['n_active_total']
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record(tstream)
    for _ in range(args.steps):
        sw.step(policy, NBR)
    ev1.record(tstream)
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    agent_steps = sw.counters()['n_active_total'] - c0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if world > 1:
        t = torch.tensor([ms, float(agent_steps)], dtype=torch.float64, device=device)
        tmax = t.clone()
        torch.distributed.all_reduce(tmax, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        ms, agent_steps = float(tmax[0]), float(t[1])
    value = agent_steps / (ms * 1e-3)

    # ---- per-stage device times (1 GPU): which kernel dominates, and its roofline fraction
    stage_ms, roof = None, None
    if world == 1:
        names = ['grid', 'neighbors', 'policy', 'status']
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
        act0 = sw.counters()['n_active_total']
        for k in range(args.steps):
            evs[k][0].record(); sw.stage_grid()
            evs[k][1].record(); sw.stage_neighbors(NBR)
            evs[k][2].record(); sw.stage_policy(policy)
            evs[k][3].record(); sw.stage_status()
            evs[k][4].record()
        torch.cuda.synchronize()
        cdiag = sw.counters()
        act = (cdiag['n_active_total'] - act0) / args.steps
        stage_ms = {nm: float(np.mean([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)]))
                    for i, nm in enumerate(names)}
        # the reference's own cost metric: AverageCost = 1000 * sum(agent.total_time) / sum(step_num) (run_sca.py:250), i.e.
        # POLICY milliseconds per agent-step (neighbour query + velocity selection; kd-tree build and env update excluded)
        average_cost = (stage_ms['neighbors'] + stage_ms['policy']) / max(act, 1.0)
        dom = max(stage_ms, key=stage_ms.get)
        peak, how = measured_peak()
        units = n_total if dom in ('grid', 'status') else act     # grid/status touch every agent, the others the active ones
        achieved = ALGO_BYTES[dom] * units / (stage_ms[dom] * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'dram_traffic.json')) as f:
                traffic = json.load(f).get(dom if (dom != 'policy' or policy == 'sca') else 'policy_' + policy)   # ncu figures are per kernel
        except Exception:
            pass
        counters = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'kernel_counters.json')) as f:
                counters = json.load(f).get(dom if (dom != 'policy' or policy == 'sca') else 'policy_' + policy)
        except Exception:
            pass
        roof = {'bound': 'hbm', 'kernel': {'grid': 'k_cell_count+k_scan_*+k_reorder', 'neighbors': 'k_neighbors_compat' if nbr_name == 'compat' else 'k_neighbors_direct',
                                           'policy': POLICY_KERNELS.get(policy, 'k_vpref+k_select_fast<%s>+k_select<%s>+k_integrate' % (policy.upper(), policy.upper())),
                                           'status': 'k_status'}[dom],
                'achieved': round(achieved, 2), 'peak': peak, 'unit': 'GB/s', 'frac': round(achieved / peak, 5),
                'traffic': traffic, 'traffic_source': 'profiles/dram_traffic.json (ncu --set full of this build at 10^6 agents; not measured in this run)',
                'counters': counters, 'peak_source': how, 'algorithmic_bytes_per_agent_step': ALGO_BYTES[dom],
                'units_per_launch': int(units), 'stage_ms': {k: round(v, 4) for k, v in stage_ms.items()},
                'average_cost_ms_per_agent_step': average_cost,
                'last_step_diag': {k: cdiag[k] for k in ('n_active', 'n_recheck', 'n_tier2', 'n_slow', 'n_collision')},
                'note': 'the policy stage is ALU / instruction-issue bound (SURVEY.md 8d): HBM fraction is reported as asked, '
                        'stage_ms shows where the step goes'}

    # ---- end to end through host buffers (pinned H2D of the inputs, D2H of the results, every step)
    e2e = None
    if args.quick:
        pass
    elif world == 1:
        # host-resident state (pinned): every step goes through the repo's public API for that case, Swarm.step_from_host --
        # upload the step's inputs, the four stages, download the new state + the float32 action matrix the reference env
        # consumes -- and returns when the results are on the host.  The velocity stays on the device (it is action[:3],
        # mampenv.py:104).  SCAB_BENCH_E2E_OVERLAP=0 selects the serial copy schedule (same bytes, bit-identical results:
        # tests/test_gpu_step_from_host.py).
        host = sw.host_buffers()
        h2d, d2h = sw.H2D_BYTES_PER_AGENT * sw.n, sw.D2H_BYTES_PER_AGENT * sw.n
        steps_e = max(3, min(args.steps, 20))
        overlap = os.environ.get('SCAB_BENCH_E2E_OVERLAP', '1') == '1'
        e2e_chunks = int(os.environ.get('SCAB_E2E_CHUNKS', '8'))        # row ranges of the policy stage (downloads overlap the next range)
        sw.step_from_host(host, policy, NBR, overlap=overlap, upload_vel=True, chunks=e2e_chunks)
        a0 = sw.counters()['n_active_total']
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps_e):
            sw.step_from_host(host, policy, NBR, overlap=overlap, chunks=e2e_chunks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {'value': (sw.counters()['n_active_total'] - a0) / dt, 'unit': 'agent-steps/s',
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'steps': steps_e,
               'ms_per_step': round(dt / steps_e * 1e3, 3), 'api': 'Swarm.step_from_host', 'policy_row_ranges': e2e_chunks if overlap else 1,
               'copies': 'overlapped with the cell list / 16-NN / status kernels' if overlap else 'serial'}

    elif os.environ.get('SCAB_BENCH_E2E_SHARDED', '1') == '1':
        # sharded: every rank keeps the authoritative copy of its owned rows [0, own_cap) in pinned host memory; a step
        # uploads them, replays the sharded step (halo exchange, migration and all) and downloads the rows + the action
        # matrix.  Rows change owners on the device (migration), so the host copy is whatever came down last -- the
        # upload restores exactly that.  Max over ranks of the wall time, sums of the agent-steps and of the bytes.
        import torch.distributed as dist
        keys_state = ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')
        ok, host, host_action, rows, st = 1.0, None, None, 0, None
        try:
            rows, st = int(sw.own_cap), sw.stream
            host = {k: torch.empty((rows,) + tuple(sw.t[k].shape[1:]), dtype=sw.t[k].dtype).pin_memory() for k in keys_state}
            host_action = torch.empty((rows, 7), dtype=sw.t['action'].dtype).pin_memory()
            with torch.cuda.stream(st):
                for k in keys_state:
                    host[k].copy_(sw.t[k][:rows], non_blocking=True)
            st.synchronize()
        except Exception as e:       # never leave a peer alone inside a collective: agree on skipping instead
            ok = 0.0
            sys.stderr.write('rank %d: sharded e2e setup failed (%r); e2e omitted\n' % (rank, e))
        flag = torch.tensor([ok], dtype=torch.float64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag[0]) > 0:
            try:
                h2d = sum(v.numel() * v.element_size() for v in host.values())
                d2h = h2d + host_action.numel() * host_action.element_size()
                steps_e = max(3, min(args.steps, 20))

                def one_sharded():
                    with torch.cuda.stream(st):
                        for k in keys_state:
                            sw.t[k][:rows].copy_(host[k], non_blocking=True)
                    sw.step(policy, NBR)
                    with torch.cuda.stream(st):
                        for k in keys_state:
                            host[k].copy_(sw.t[k][:rows], non_blocking=True)
                        host_action.copy_(sw.t['action'][:rows], non_blocking=True)
                    st.synchronize()          # the caller needs the results on the host before the next step
                one_sharded()
                a0 = sw.counters()['n_active_total']
                barrier()
                t0 = time.perf_counter()
                for _ in range(steps_e):
                    one_sharded()
                barrier()
                dt = time.perf_counter() - t0
                tmax = torch.tensor([dt], dtype=torch.float64, device=device)
                tsum = torch.tensor([float(sw.counters()['n_active_total'] - a0), float(h2d), float(d2h)], dtype=torch.float64, device=device)
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
                e2e = {'value': float(tsum[0]) / float(tmax[0]), 'unit': 'agent-steps/s', 'h2d_bytes_per_step': int(tsum[1]),
                       'd2h_bytes_per_step': int(tsum[2]), 'steps': steps_e, 'ms_per_step': round(float(tmax[0]) / steps_e * 1e3, 3),
                       'copies': 'per rank: owned rows up, sharded step, owned rows + actions down'}
            except Exception as e:       # the same Python error hits every rank at the same point: all of them skip
                e2e = None
                sys.stderr.write('rank %d: sharded e2e failed (%r); e2e omitted\n' % (rank, e))

    cpu = None
    if world == 1 and not args.no_cpu and not args.quick:
        cpu = cpu_baseline(args.cpu_agents, 1, policy, budget_s=12.0, cfg=cfg)

    if rank == 0:
        if cfg is None:
            metric = METRIC
            conf = {'workload': 'BASELINE configs[4]: SCA synthetic uniform-random swarm, %d UAVs, density 1e6/(560 m)^3, '
                                'straight-line v_pref, KNN neighbour mode' % n_total,
                    'policy': policy, 'agents': n_total, 'neighbour_mode': 'KNN (true 16 nearest; NOT the reference\'s traversal-order '
                    'lists, SURVEY fact 3 -- REF_COMPAT reproduces those and is what --config 1 / 2 run)',
                    'l2': 'per-step working set (~%d MB) exceeds the 126 MB L2' % (n_total * 420 // 1000000),
                    'parallelism': 'spatial slabs x%d + halo exchange' % world if world > 1 else 'single GPU'}
        else:
            metric = 'agent-steps/sec (%s, %d agents)' % (policy.upper(), n_total)
            conf = {'workload': cfg['workload'], 'policy': policy, 'agents': n_total,
                    'neighbour_mode': 'REF_COMPAT (kd-tree rebuilt on the host every step, as mampenv.py:28 does)' if nbr_name == 'compat' else 'KNN',
                    'l2': 'working set fits the 126 MB L2: the step is launch / latency bound at this size' if n_total < 200000 else '',
                    'parallelism': 'single GPU'}
        if args.skip_steps:
            conf['rollout_window'] = 'steps %d..%d of the rollout' % (args.skip_steps + max(args.warmup, 3), args.skip_steps + max(args.warmup, 3) + args.steps)
        out = {'metric': metric, 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world, 'steps': args.steps,
               'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
               'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32 candidate tests + f64 decision scalars',
               'data': 'synthetic',
               'config': conf,
               # + 5 halo kernels per step when sharded: k_halo_fill, k_halo_pack, k_halo_unpack, k_halo_pack_pos, k_halo_unpack_pos
               'gpu_launches': (KERNELS_PER_STEP.get(policy, KERNELS_PER_STEP_SAMPLED) + (6 if world > 1 else 0)) * args.steps, 'clocks': clocks}
        if roof:
            out['roofline'] = roof
        if e2e:
            out['e2e'] = e2e
        if cpu:
            out['cpu_baseline'] = cpu
        if cfg is not None and args.config in (1, 2) and not args.quick:
            out['dropin_shim'] = shim_rate(cfg, policy, min(args.steps, 20))
        print(json.dumps(out))
    if world > 1:
        # captured graphs hold NCCL work: release them, then leave without the (blocking) group teardown
        sw.close()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def oracle_world(n_agents, policy, cfg):
    """The oracle's world for a CPU leg: the default workload at `n_agents` agents, or a --config scenario.
    -> (world, stepper(), sample description, neighbour mode)"""
    from oracle import oracle as orc
    from sca_b200 import scenarios
    if cfg is None:
        pos, goal, obs = scenarios.uniform_box_numpy(n_agents, DENSITY, seed=1)
        what, mode, dub = '%d agents at the workload density' % n_agents, orc.NBR_KNN, False
    else:
        pos, goal, obs = cfg['cpu'](scenarios)
        what, mode, dub = 'the scenario at %d agents' % len(pos), (orc.NBR_REF_COMPAT if cfg['nbr'] == 'compat' else orc.NBR_KNN), cfg['dubins']
    w = orc.World(pos[:, :3], goal[:, :3], heading=pos[:, 3:6], obstacles=obs, initial_pos=pos, goal_pos=goal)
    timing = {'policy': 0.0}
    if dub:
        prov = orc.DubinsProvider(len(pos))
        gh = np.ascontiguousarray(goal[:, 3:6])

        def one():
            t0 = time.perf_counter()
            w.step_dubins(policy, prov, gh, mode)
            timing['policy'] += time.perf_counter() - t0      # (includes the env update of step_dubins: small next to the tracker)
    else:
        def one():
            t0 = time.perf_counter()
            w.policy_step(policy, mode)
            timing['policy'] += time.perf_counter() - t0
            w.env_update(fast=True)
    return w, one, what, ('REF_COMPAT' if mode == orc.NBR_REF_COMPAT else 'KNN'), timing


def cpu_baseline(n_agents, threads, policy, budget_s=15.0, steps=None, cfg=None):
    """The CPU oracle port of the reference (oracle/sca_oracle.c) on a bounded sample of the same workload:
    the same density (or scenario) and policy.  Returns the cpu_baseline object of the bench contract."""
    from oracle import oracle as orc
    orc.set_threads(threads)
    w, one, what, mode, timing = oracle_world(n_agents, policy, cfg)
    for _ in range(2):                      # leave the zero-velocity shortcut (step 0)
        one()
    timing['policy'] = 0.0
    done_steps, agent_steps = 0, 0
    t0 = time.perf_counter()
    while True:
        one()
        agent_steps += int(w.ran.sum())
        done_steps += 1
        el = time.perf_counter() - t0
        if (steps and done_steps >= steps) or (not steps and (el > budget_s or done_steps >= 200)):
            break
    return {'value': agent_steps / el, 'unit': 'agent-steps/s', 'cores': threads, 'kind': 'port',
            'sample': '%s, %d steps, %s policy, %s neighbour mode, %.1f s' % (what, done_steps, policy, mode, el),
            'ms_per_step': el / done_steps * 1e3, 'steps': done_steps,
            'average_cost_ms_per_agent_step': 1e3 * timing['policy'] / max(agent_steps, 1)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = []
    total_agent_steps, total_t = 0, 0.0
    # each "step" of this arm is one oracle step over the bounded sample
    from oracle import oracle as orc
    orc.set_threads(threads)
    cfg, policy, _ = config_of(args)
    args.policy = policy
    w, one, what, mode, timing = oracle_world(args.cpu_agents, policy, cfg)
    n = w.n
    for _ in range(max(args.warmup, 2)):
        one()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one()
        dt = time.perf_counter() - t0
        per_step.append(dt)
        total_agent_steps += int(w.ran.sum())
        total_t += dt
    value = total_agent_steps / total_t
    sample = '%s per step (bounded sample), %s policy, %s neighbour mode' % (what, args.policy, mode)
    out = {'impl': 'reference', 'metric': METRIC if cfg is None else 'agent-steps/sec (%s, %d agents)' % (policy.upper(), len(cfg['gen'](__import__('sca_b200.scenarios', fromlist=['x']))[0])), 'value': value, 'unit': 'agent-steps/s',
           'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': args.steps, 'warmup': max(args.warmup, 2),
           'ms_per_step': total_t / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
           'dtype': 'f64', 'data': 'synthetic',
           'config': {'workload': ('BASELINE configs[4]: SCA synthetic uniform-random swarm (bounded sample: %d UAVs at the '
                                   '1M-UAV density), straight-line v_pref, KNN neighbour mode' % n) if cfg is None else cfg['workload'] + ' (CPU sample: %s)' % what,
                      'policy': args.policy, 'agents': n,
                      'note': 'the Python reference runs ~40 agent-steps/s single-threaded (BASELINE.md); this arm times its C '
                              'restatement (oracle/sca_oracle.c), which is pinned to the reference by the golden fixtures'},
           'cpu_baseline': {'value': value, 'unit': 'agent-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
           'e2e': {'value': value, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
           'gpu_launches': 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--agents', type=int, default=1000000)
    ap.add_argument('--policy', default=None, help='default: the policy of the chosen --config (sca for the headline workload)')
    ap.add_argument('--config', type=int, default=5, choices=[1, 2, 3, 4, 5], help='BASELINE.json configs[N-1]; 5 = the headline workload (default)')
    ap.add_argument('--skip-steps', type=int, default=0, help='untimed steps before the warm-up: sample a later part of the rollout')
    ap.add_argument('--cpu-agents', type=int, default=20000)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--quick', action='store_true', help='A/B runs: device-timed value + stage times only (no e2e, no CPU baseline)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
