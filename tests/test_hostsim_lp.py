"""CPU: the DEVICE SOURCE of the ORCA-LP policy stage (sca_b200/csrc/scab_policy_lp.cu: orca_plane, the fp32 plane
pre-filter, linearProgram1-4, cartesian2spherical, the kinematic update and the kernel bodies themselves) compiled for the
host by tests/hostsim (one thread at a time, -ffp-contract=off to mirror the device build's -fmad=false) and run against
(a) the reference's recorded outputs on the ORCA-LP fixtures and (b) the oracle on dense random swarms with obstacles.

All four launch configurations of `launch_policy_lp` are exercised: the register-only pass with the fp32 pre-filter + the
work-list pass (default), the same without the pre-filter (SCAB_LP_F64=1), the one-kernel version (SCAB_NO_FAST=1) and the
work-list pass with compact 32-byte planes (SCAB_LP_COMPACT=1); they must agree with each other bit for bit.  This is not
a substitute for the -m gpu parity tests (those run the compiled sm_100a code); it checks the kernel LOGIC on every CPU run and lets the LP kernels be edited without a GPU at hand.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SIM = os.path.join(HERE, 'hostsim')
LIB = os.path.join(SIM, 'libhostsim_lp.so')
CSRC = os.path.join(HERE, '..', 'sca_b200', 'csrc')


def _lib():
    deps = [os.path.join(SIM, f) for f in ('lp_host.cpp', 'cuda_host_shim.h')] + \
           [os.path.join(CSRC, f) for f in ('scab_policy_lp.cu', 'scab_common.cuh')] + [os.path.join(HERE, '..', 'include', 'scab200.h')]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-w', '-D_GNU_SOURCE',
                               '-I/usr/local/cuda/include', os.path.join(SIM, 'lp_host.cpp'), '-o', LIB])
    L = C.CDLL(LIB)
    L.hostsim_policy_lp.restype = C.c_int
    return L


def run_device_source(mode, w, pre, nbr_cnt, nbr_ids, coll_query):
    """One policy stage of the host-compiled device code on the pre-step state `pre` with the given neighbour lists.
    Returns the arrays the kernel writes."""
    from sca_b200 import capi
    L = _lib()
    n, m = w.n, w.m
    a = lambda x, dt: np.ascontiguousarray(x, dtype=dt).copy()
    st = {'pos': a(pre['pos'], 'f8'), 'vel': a(pre['vel'], 'f4'), 'heading': a(pre['heading'], 'f8'), 'goal': a(w.goal, 'f8'),
          'radius': a(w.radius, 'f8'), 'pref_speed': a(w.pref_speed, 'f8'), 'status': a(pre['status'] & 7, 'i4'),
          'total_dist': a(pre['total_dist'], 'f8'), 'step_num': a(pre['step_num'], 'i4'), 'max_run_dist': a(w.max_run_dist, 'f8'),
          'vpref_out': np.zeros((n, 3)), 'action': np.zeros((n, 7), 'f4'), 'tag': np.zeros(n, 'i4'), 'n_suit': np.zeros(n, 'i4'),
          'plane_fail': np.zeros(n, 'i4'), 'n_planes': np.zeros(n, 'i4'), 'lp4_proj': np.zeros(n, 'i4')}
    p = capi.StatePtrs()
    for k, v in st.items():
        setattr(p, k, v.ctypes.data)
    posr = np.ascontiguousarray(np.c_[pre['pos'], w.radius], 'f8')
    vel4 = np.zeros((n, 4), 'f4')
    vel4[:, :3] = pre['vel']
    vel4[:, 3] = (pre['status'] & 7).astype('i4').view('f4')        # status bits travel in vel.w (scab_grid.cu k_reorder)
    nbr = a(np.where(nbr_ids < 0, 0, nbr_ids), 'i4')
    cnt = a(np.maximum(nbr_cnt, 0), 'i4')
    qc = a(coll_query, 'i4')
    obs = np.ascontiguousarray(w.obs if m else np.zeros((1, 4)), 'f8')
    slow, slow_n, counters = np.zeros(max(n, 1), 'i4'), np.zeros(1, 'i4'), np.zeros(16, 'u8')
    rc = L.hostsim_policy_lp(C.c_int(mode), C.byref(p), C.c_int(n), C.c_int(m), C.c_void_p(posr.ctypes.data), C.c_void_p(vel4.ctypes.data),
                             C.c_void_p(nbr.ctypes.data), C.c_void_p(cnt.ctypes.data), C.c_void_p(qc.ctypes.data), C.c_void_p(obs.ctypes.data),
                             C.c_double(0.1), C.c_double(10.0), C.c_double(1.0), C.c_void_p(slow.ctypes.data),
                             C.c_void_p(slow_n.ctypes.data), C.c_void_p(counters.ctypes.data))
    assert rc == 0
    st['n_slow'] = int(slow_n[0])
    st['n_ran'] = int(counters[0])
    return st


def _pre(w):
    return {k: getattr(w, k).copy() for k in ('pos', 'vel', 'heading', 'status', 'total_dist', 'step_num')}


@pytest.mark.parametrize('name', ['random1000_orca_lp', 'dense60_orca_lp', 'circle16_orca_lp'])
def test_device_lp_source_matches_the_reference_fixtures(name):
    g = gu.load(name)
    T = min(g['pos'].shape[0], 12)
    w = orc.World(g['pos'][0], g['goal'], heading=g['heading'][0], radius=g['radius'], pref_speed=g['pref_speed'],
                  obstacles=g['obstacles'], initial_pos=g['initial_pos'], goal_pos=g['goal_pos'])
    agents = lp4 = deferred = 0
    for t in range(T):
        s = gu.state_at(g, t)
        for k in ('pos', 'heading', 'status', 'total_dist', 'step_num'):
            getattr(w, k)[:] = s[k]
        w.vel[:] = s['vel'].astype(np.float32)
        pre = _pre(w)
        w.policy_step('orca_lp', orc.NBR_REF_COMPAT)          # the reference's ordered neighbour lists (pinned elsewhere)
        ran = g['ran'][t].astype(bool)
        outs = [run_device_source(mode, w, pre, w.nbr_cnt, w.nbr_ids, w.coll_query) for mode in (1, 3, 0, 4, 5)]
        for o in outs:
            assert o['n_ran'] == int(ran.sum())
            for k in ('tag', 'plane_fail', 'n_planes', 'lp4_proj'):
                assert np.array_equal(o[k][ran], g[k][t][ran]), (name, t, k)
            ref = g['action'][t].astype(np.float32)
            assert np.array_equal(o['action'][ran][:, :4], ref[ran][:, :4]), (name, t)
            assert np.abs(o['action'][ran][:, 4:].astype('f8') - ref[ran][:, 4:].astype('f8')).max(initial=0) <= 5e-7
        for o in outs[1:]:                                      # the three launch configurations agree bit for bit
            for k in ('pos', 'vel', 'heading', 'total_dist', 'action', 'status', 'tag', 'plane_fail', 'lp4_proj', 'vpref_out'):
                assert np.array_equal(o[k], outs[0][k]), (name, t, k)
        assert outs[0]['n_slow'] == outs[1]['n_slow']           # the fp32 pre-filter defers exactly the same agents
        agents += int(ran.sum())
        lp4 += int((g['lp4_proj'][t][ran] > 0).sum())
        deferred += outs[0]['n_slow']
    print(name, 'agent-steps', agents, 'with LP4', lp4, 'deferred to the list pass', deferred)
    assert agents > 0 and deferred > 0


@pytest.mark.parametrize('seed', [3, 4])
def test_device_lp_source_matches_the_oracle_on_dense_swarms(seed):
    from sca_b200 import scenarios
    n = 3000
    pos6, goal6, _ = scenarios.uniform_box_numpy(n, density=3.0 * 1.0e6 / 560.0 ** 3, seed=seed, min_sep=0.98)
    rng = np.random.default_rng(seed)
    L = pos6[:, :3].max()
    obs = np.c_[rng.uniform(0, L, (30, 3)) + [0, 0, 20], rng.uniform(0.5, 2.0, 30)]
    orc.set_threads(8)
    w = orc.World(pos6[:, :3], goal6[:, :3], heading=pos6[:, 3:6], obstacles=obs, initial_pos=pos6, goal_pos=goal6)
    tot = dict(agents=0, deferred=0, lp4=0, collided=0)
    for t in range(6):
        pre = _pre(w)
        w.policy_step('orca_lp', orc.NBR_KNN)
        ran = w.ran.astype(bool)
        outs = [run_device_source(mode, w, pre, w.nbr_cnt, w.nbr_ids, w.coll_query) for mode in (1, 3, 0, 4, 5)]
        ref = {k: getattr(w, k).copy() for k in ('tag', 'plane_fail', 'n_planes', 'lp4_proj', 'action', 'vpref_store')}
        w.env_update(fast=True)
        for o in outs:
            for k in ('tag', 'plane_fail', 'n_planes', 'lp4_proj'):
                assert np.array_equal(o[k][ran], ref[k][ran]), (t, k)
            assert np.array_equal(o['action'][ran], ref['action'][ran]), t          # same libm on the host: every column exact
            assert np.array_equal(o['vpref_out'][ran], ref['vpref_store'][ran]), t
            # the kernel fuses the kinematic update (mampenv.py:83-105)
            assert np.array_equal(o['pos'], w.pos) and np.array_equal(o['heading'], w.heading), t
            assert np.array_equal(o['vel'], w.vel) and np.array_equal(o['total_dist'], w.total_dist), t
        assert outs[0]['n_slow'] == outs[1]['n_slow']
        tot['agents'] += int(ran.sum())
        tot['deferred'] += outs[0]['n_slow']
        tot['lp4'] += int((ref['lp4_proj'][ran] > 0).sum())
        tot['collided'] = int(((w.status & 2) != 0).sum())
    print(seed, tot)
    assert tot['deferred'] > 100 and tot['lp4'] > 0 and tot['collided'] > 0
