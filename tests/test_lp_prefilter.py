"""CPU: the guard band of the fp32 plane pre-filter in k_policy_lp<1> (tools/lp_prefilter_check.py replays the kernel's
fp32 sequence in NumPy float32 against float64).  The kernel trusts fp32 only for "this plane is satisfied by a margin";
the band must dwarf the worst fp32 error, and the constants here must be the ones compiled into the kernel."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def test_band_constants_match_the_kernel():
    import lp_prefilter_check as chk
    src = open(os.path.join(ROOT, 'sca_b200', 'csrc', 'scab_policy_lp.cu')).read()
    val = lambda name: float(re.search(r'#define %s ([0-9.e+-]+)f' % name, src).group(1))
    assert val('LP_F32_EPS') == chk.EPS and val('LP_F32_WL_MIN') == chk.WL_MIN and val('LP_F32_DIF_MIN') == chk.DIF_MIN


def test_fp32_margin_error_is_far_inside_the_band():
    import lp_prefilter_check as chk
    worst, wrong = chk.main(n=400000, quiet=True)
    assert wrong == 0                       # never "satisfied" in fp32 while violated in fp64
    assert worst * 50 < chk.EPS, worst
