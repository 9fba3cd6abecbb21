"""TEST INFRASTRUCTURE: run the library's C ABI on the host emulation build (tests/hostsim) instead of libscab200.so.

`activate(monkeypatch)` points sca_b200.capi at tests/hostsim/libscab200_hostsim.so -- the same csrc/*.cu sources compiled for
the host, every CUDA thread of a block a real host thread -- and lets `Swarm` keep its tensors in host memory.  Nothing in
the product knows about this: it is a pytest monkeypatch, undone at the end of the test.  It checks the kernels' LOGIC and
the ABI's host code on machines without a GPU; the -m gpu tests remain the parity tests proper."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def activate(monkeypatch):
    import torch
    sys.path.insert(0, os.path.join(HERE, 'hostsim'))
    import build as hostsim_build
    from sca_b200 import capi, swarm
    lib = hostsim_build.build()
    monkeypatch.setattr(capi, 'LIB_PATH', lib)
    monkeypatch.setattr(capi, '_lib', None)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
    monkeypatch.setattr(swarm.Swarm, '_stream', staticmethod(lambda: 0))
    orig_init = swarm.Swarm.__init__

    def init(self, n, device='cpu', debug=False, params=None, capacity=None):
        orig_init(self, n, device='cpu', debug=debug, params=params, capacity=capacity)
    monkeypatch.setattr(swarm.Swarm, '__init__', init)
    return lib


class _NoStream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_event(self, *a):
        pass


class _NoEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass


def activate_streams(monkeypatch):
    """On top of activate(): CUDA streams / devices as no-ops, for code that manages its own stream (ShardedSwarm)."""
    import contextlib
    import torch
    monkeypatch.setattr(torch.cuda, 'Stream', lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _NoStream())
    monkeypatch.setattr(torch.cuda, 'set_device', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'Event', _NoEvent)
