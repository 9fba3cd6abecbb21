"""Rollout metrics / logs in the reference's formats (sca_b200/logs.py).
CPU: rollout_info on the reference's recorded final state reproduces its four metrics and the env_cfg.json schema.
GPU: BatchedEnv's trajectory log follows the recorded trajectory."""
import json
import os

import numpy as np
import pytest

import golden_util as gu

SCHEMA = ['all_agent_info', 'all_obstacle', 'all_compute_time', 'all_straight_distance', 'all_distance', 'successful_num',
          'all_desire_step_num', 'all_step_num', 'SuccessRate', 'ExtraTime', 'ExtraDistance', 'AverageSpeed', 'AverageCost']


@pytest.mark.parametrize('name', ['circle16_rvo3d', 'circle16_orca_lp', 'dense40_sca', 'takeoff16_sca'])
def test_rollout_info_reproduces_reference_metrics(name, tmp_path):
    from sca_b200 import logs
    g = gu.load(name)
    info = logs.rollout_info(g['final_status'], g['final_total_dist'], g['final_step_num'], g['straight_len'], g['desire_steps'],
                             g['radius'], g['goal'], g['obstacles'], compute_time=1.5)
    assert list(info.keys()) == SCHEMA                       # run_sca.py:198-212, same order
    m = g['metrics']
    assert info['SuccessRate'] == m[0]
    assert np.allclose([info['ExtraTime'], info['ExtraDistance'], info['AverageSpeed']], m[1:], rtol=0, atol=1e-9)
    assert info['AverageCost'] == 1000 * 1.5 / info['all_step_num']
    assert len(info['all_agent_info']) == len(g['radius']) and len(info['all_obstacle']) == len(g['obstacles'])
    assert set(info['all_agent_info'][0]) == {'id', 'gp', 'radius', 'goal_pos'}
    if len(g['obstacles']):
        assert set(info['all_obstacle'][0]) == {'position', 'shape', 'feature'}
    p = str(tmp_path / 'log' / 'env_cfg.json')
    logs.write_env_cfg(p, info)
    assert json.load(open(p))['successful_num'] == info['successful_num']


@pytest.mark.gpu
def test_trajectory_log_follows_the_reference_rollout(tmp_path):
    from sca_b200 import capi, logs
    from sca_b200.env import BatchedEnv
    g = gu.load('circle16_rvo3d')
    env = BatchedEnv('rvo3d', nbr_mode=capi.NBR_REF_COMPAT, log_every=2, log_capacity=64)
    env.set_agents(g['initial_pos'], g['goal_pos'], g['obstacles'], radius=g['radius'], pref_speed=g['pref_speed'])
    T = 100
    for _ in range(T):
        env.step(sync=False)
    cols = env.traj.columns()
    assert cols.shape == (50, 16, 13)
    for r in range(50):                                   # row r = state after step 2 (r + 1) = the fixture's pre-state of that step
        s = gu.state_at(g, 2 * (r + 1))
        assert np.abs(cols[r, :, 0:3] - s['pos']).max() <= 1e-5
        assert np.abs(cols[r, :, 3:6] - s['heading']).max() <= 1e-5
        assert np.abs(cols[r, :, 6:9] - s['vel']).max() <= 2e-5
    assert np.array_equal(cols[0, :, 9:12], g['goal']) and np.array_equal(cols[0, :, 12], g['radius'])
    info = env.write_log(str(tmp_path / 'log'))
    assert os.path.exists(str(tmp_path / 'log' / 'env_cfg.json')) and os.path.exists(str(tmp_path / 'log' / 'trajs' / 'agent15.csv'))
    assert list(info.keys()) == SCHEMA
    head = open(str(tmp_path / 'log' / 'trajs' / 'agent0.csv')).readline().strip().split(',')
    assert head == logs.ANIMATION_COLUMNS
    for _ in range(40):                                   # capacity 64 rows: the log stops, the run does not
        env.step(sync=False)
    assert env.traj.rows == 64 and env.traj.dropped == 6
