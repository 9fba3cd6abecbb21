"""CPU: the DEVICE SOURCE of every kernel (sca_b200/csrc/*.cu, unmodified) compiled for the host by tests/hostsim and driven
through the same C ABI, `sca_b200.capi` binding and `Swarm` as on the GPU -- every CUDA thread of a block is a host thread,
`*_sync` warp primitives and `__syncthreads` are rendezvous points, "device" memory is host memory.

This module runs a slice of the -m gpu parity tests on that build on every CPU run: the reference fixtures through
cell list -> neighbours -> policy kernels -> status for all six policies (the 8-lanes-per-agent `k_select_fast`, the
one-warp-per-agent `k_select`, both ORCA-LP passes, the Dubins tracker, waypoints, obstacles), and the KNN path against the
oracle on a dense random swarm.  The whole -m gpu suite runs the same way with `pytest tests -m gpu --hostsim` (11 min;
91 passed at the time of writing).  It validates kernel LOGIC and the ABI's host code without a GPU -- index arithmetic,
branches, work lists, warp-cooperative protocols -- not the compiled sm_100a code, its libm or its performance: the -m gpu
tests remain the parity tests proper.
"""
import numpy as np
import pytest

import golden_util as gu


@pytest.fixture()
def hostsim(monkeypatch):
    import hostsim_util
    return hostsim_util.activate(monkeypatch)


# (fixture, steps, full_search)
SLICE = [('circle16_sca', 14, False), ('circle16_sca', 6, True), ('takeoff16_sca', 10, False), ('dense40_sca', 8, False),
         ('circle16_rvo3d_dubins', 10, False), ('dense40_srvo3d', 8, False), ('dense40_rvo3d', 8, True),
         ('dense40_orca_sampled', 8, False), ('dense60_orca_lp', 12, True), ('waypoints16_orca_lp', 40, True),
         ('waypoints16_srvo3d', 40, False), ('sphere100_rvo3d', 3, False)]


@pytest.mark.parametrize('name,steps,full_search', SLICE, ids=['%s-%d-%s' % (a, b, 'full' if c else 'lazy') for a, b, c in SLICE])
def test_reference_fixture_on_the_host_built_kernels(hostsim, name, steps, full_search):
    import test_gpu_parity as T
    tot = T._run_fixture(name, max_steps=steps, full_search=full_search)
    print(name, tot)
    assert tot['agents'] > 0
    for k in T.EXACT:
        assert tot.get(k, 0) == 0, (name, k, tot)
    assert tot['vel_bad'] == 0 and tot['pos_bad'] == 0 and tot['head_bad'] == 0 and tot['dist_bad'] == 0, (name, tot)
    # same libm as the reference here: even the heading deltas are exact
    assert tot['action_exact'] == 0 and tot['angle_bad'] == 0 and tot['angle_ulp_flips'] == 0, (name, tot)


@pytest.mark.parametrize('policy', ['sca', 'orca_lp', 'orca_sampled'])
def test_knn_path_on_the_host_built_kernels_matches_the_oracle(hostsim, policy):
    import test_gpu_knn_oracle as K
    from sca_b200 import scenarios
    pos, goal, _ = scenarios.uniform_box_numpy(1500, 3.0 * 1.0e6 / 560.0 ** 3, seed=21, min_sep=0.98)
    rng = np.random.default_rng(6)
    L = pos[:, :3].max()
    obs = np.c_[rng.uniform(0, L, (20, 3)) + [0, 0, 20], rng.uniform(0.5, 2.0, 20)]
    tot = K._compare(pos, goal, obs, policy, steps=3, full_search=False)
    print(policy, tot)
    assert tot['agents'] > 4000
    for k in ('nbr', 'tag', 'nsuit', 'vel', 'status', 'planes', 'flips'):
        assert tot[k] == 0, (policy, tot)


def test_gpu_edge_cases_on_the_host_built_kernels(hostsim):
    import test_gpu_edge_cases as E
    E.test_empty_swarm_steps_to_done()
    for policy in ('sca', 'orca_lp'):
        E.test_single_agent_flies_straight_like_the_oracle(policy)
        E.test_one_agent_among_obstacles_matches_oracle(policy)
    E.test_all_finished_is_a_no_op()
    E.test_knn_integer_key_path_equals_exact_path_and_oracle()
    E.test_boundary_errors_are_reported_not_swallowed()


def test_device_dubins_tracker_on_the_host_built_kernels(hostsim):
    import test_gpu_dubins as D
    D.test_first_plan_matches_oracle_planner()
    D.test_vpref_teacher_forced('takeoff16_sca')
    D.test_free_running_rollout('circle16_rvo3d_dubins')


def test_mamp_shims_on_the_host_built_kernels(hostsim):
    import test_gpu_mamp_compat as M
    for name, cls in M.CASES[:3]:
        M.test_shim_matches_reference_actions(name, cls)


def test_all_fixtures_first_steps_on_the_host_built_kernels(hostsim):
    """Every committed fixture, first 8 steps, lazy search (the default configuration of the library)."""
    import test_gpu_parity as T
    for name in gu.golden_names():
        tot = T._run_fixture(name, max_steps=8, full_search=False)
        bad = {k: tot.get(k, 0) for k in T.EXACT + ('vel_bad', 'pos_bad', 'head_bad', 'dist_bad', 'action_exact', 'angle_ulp_flips') if tot.get(k, 0)}
        assert not bad, (name, bad)


def test_the_emulated_library_is_not_the_product(hostsim):
    """The emulation is reached only through this monkeypatch: the product binding points at libscab200.so."""
    import importlib
    from sca_b200 import capi
    assert capi.LIB_PATH.endswith('libscab200_hostsim.so')
    fresh = importlib.util.find_spec('sca_b200.capi')
    src = open(fresh.origin).read()
    assert 'hostsim' not in src and "libscab200.so" in src
