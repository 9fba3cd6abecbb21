"""CPU: the C-ABI library loads without a GPU and exports every symbol include/scab200.h declares;
the product fails loudly (no CPU fallback) when asked to compute without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'scab200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(scab_[a-z0-9_]+)\s*\(', src)))


def test_header_matches_binding_table():
    from sca_b200 import capi
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from sca_b200 import build, capi
    lib = build.build()
    L = ctypes.CDLL(lib)
    for s in header_symbols():
        assert hasattr(L, s), s
    L.scab_version.restype = ctypes.c_char_p
    assert b'sm_100a' in L.scab_version()
    p = capi.default_params()
    assert (p.neighbor_dist, p.time_step, p.time_horizon, p.max_speed, p.near_goal) == (10.0, 0.1, 10.0, 1.0, 0.5)


def test_struct_layouts_match_the_header():
    from sca_b200 import capi
    # scab_params: 5 doubles + 3 doubles + 1 double + 3 int32 + 1 int32
    assert ctypes.sizeof(capi.Params) == 9 * 8 + 16
    assert ctypes.sizeof(capi.StatePtrs) == 8 * (len(capi.STATE_FIELDS) + 6)   # + n_global, goal_heading, waypoints, waypoint_left, now_goal, waypoint_stride
    assert ctypes.sizeof(capi.Counters) == 64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from sca_b200 import capi
    from sca_b200.swarm import Swarm
    with pytest.raises(capi.ScabError):
        Swarm(8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'sca_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'libsca_oracle' not in txt, (f, 'uses the oracle')
