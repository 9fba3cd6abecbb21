"""Where an end-to-end step goes: raw pinned copy rates and Swarm.step_from_host in its three schedules (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sca_b200 import capi
sw = bench.make_swarm(1000000, 0, 1, 'cuda:0')
for _ in range(5):
    sw.step('sca', capi.NBR_KNN)
torch.cuda.synchronize()
host = sw.host_buffers()
def t_copy(fn, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
up = lambda: [sw.t[k][:sw.n].copy_(host[k], non_blocking=True) for k in ('pos', 'heading', 'status', 'total_dist', 'step_num')]
dn = lambda: [host[k].copy_(sw.t[k][:sw.n], non_blocking=True) for k in ('pos', 'heading', 'status', 'total_dist', 'step_num', 'action')]
print('H2D 64 MB: %.3f ms   D2H 92 MB: %.3f ms' % (t_copy(up), t_copy(dn)))
both = lambda: (up(), dn())
print('both directions back to back on one stream: %.3f ms' % t_copy(both))
for name, kw in (('serial', dict(overlap=False)), ('overlap, 1 range', dict(overlap=True, chunks=1)), ('overlap, 4 ranges', dict(overlap=True, chunks=4)),
                 ('overlap, 8 ranges', dict(overlap=True, chunks=8))):
    sw.step_from_host(host, 'sca', upload_vel=True, **kw)
    print('%-20s %.3f ms/step' % (name, t_copy(lambda: sw.step_from_host(host, 'sca', **kw), 10)))
