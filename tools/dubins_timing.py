"""Timing of the SCA step with the on-device Dubins tracker (not a test): uniform box, KNN mode.
python tools/dubins_timing.py [n] [steps]"""
import sys
import time
sys.path.insert(0, '.')
import torch
from sca_b200 import capi, scenarios
from sca_b200.env import BatchedEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
pos, goal, obs = scenarios.uniform_box_torch(n)
for dub in (False, True):
    env = BatchedEnv('sca', nbr_mode=capi.NBR_KNN, dubins=dub)
    env.set_agents(pos, goal, obs)
    sw = env.sw
    per = []
    for it in range(steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(); sw.stage_grid(); ev[1].record(); sw.stage_neighbors(capi.NBR_KNN); ev[2].record()
        sw.stage_policy('sca'); ev[3].record(); sw.stage_status(); ev[4].record()
        torch.cuda.synchronize()
        per.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    line = 'dubins=%s n=%d' % (dub, n)
    for it in (0, 1, 2, 5, 10, steps - 1):
        line += ' | step %d: policy %.2f ms total %.2f ms' % (it, per[it][2], sum(per[it]))
    print(line)
    if dub:
        st = sw.dubins_states()
        print('replans per agent over %d steps: mean %.2f max %d' % (steps, st['replans'].mean(), st['replans'].max()))
    print(sw.counters(), env.metrics())
