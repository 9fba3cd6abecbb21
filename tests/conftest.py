import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

# -m gpu tests that need real hardware even under --hostsim (size, several devices, CUDA streams / graphs)
HOSTSIM_SKIP = ('test_gpu_full_size', 'test_sharding_gpu', 'test_voxel_obstacle_field_at_scale', 'config4_at_100000')


def pytest_addoption(parser):
    parser.addoption('--hostsim', action='store_true', default=False,
                     help='run the -m gpu tests on the host emulation build of the kernels (tests/hostsim): a functional check of '
                          'the device source on a machine without a GPU, not a parity claim')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    if config.getoption('--hostsim'):
        import hostsim_util
        mp = pytest.MonkeyPatch()
        hostsim_util.activate(mp)
        config._hostsim_mp = mp


def pytest_unconfigure(config):
    mp = getattr(config, '_hostsim_mp', None)
    if mp is not None:
        mp.undo()


def pytest_collection_modifyitems(config, items):
    if config.getoption('--hostsim'):
        skip = pytest.mark.skip(reason='needs real hardware (not run on the host emulation)')
        for it in items:
            if 'gpu' in it.keywords and any(s in it.nodeid for s in HOSTSIM_SKIP):
                it.add_marker(skip)
        return
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)
