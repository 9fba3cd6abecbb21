"""Error study behind the fp32 plane pre-filter of k_policy_lp<1> (sca_b200/csrc/scab_policy_lp.cu, orca_margin_f32).

The register-only LP pass only needs the SIGN of  m = n . (point - v0)  for each ORCA plane (orca3dPolicyOfficial.py:56-106
and :252), v0 = linearProgram3's starting point.  The kernel evaluates m in fp32 first and trusts a negative sign only
when m < -EPS and the construction is well conditioned (GUARDS below); everything else is re-evaluated with the
reference's fp64 expressions.  This script replays the kernel's fp32 sequence in NumPy float32 (IEEE, no FMA -- the
library is built with -fmad=false) against float64 on random and adversarial neighbour configurations and prints
the largest |m32 - m64| seen among the configurations that pass the guards.  EPS must stay well above it.

    python tools/lp_prefilter_check.py [samples]
"""
import sys

import numpy as np

F = np.float32
EPS = 2e-3            # kernel: LP_F32_EPS
WL_MIN = 1e-2         # kernel: LP_F32_WL_MIN   (|w| or |ww| below this: direction ill conditioned)
DIF_MIN = 5e-2        # kernel: LP_F32_DIF_MIN  (||p|^2 - R^2| below this: cone leg ill conditioned, collision switch near)
INV_T, INV_DT = 0.1, 10.0


def margin64(p, v, R, d):
    """fp64 margin per the reference; p = pB - pA, v = vA - vB, d = vA - v0."""
    distSq = (p * p).sum(1)
    RSq = R * R
    m = np.empty(len(p))
    # collision branch
    w = v - INV_DT * p
    wl = np.sqrt((w * w).sum(1))
    mc = (w / wl[:, None] * d).sum(1) + 0.5 * (R * INV_DT - wl)
    # no-collision
    w = v - INV_T * p
    wl2 = (w * w).sum(1)
    dp = (w * p).sum(1)
    capA = (dp < 0) & (dp * dp > RSq * wl2)
    wl = np.sqrt(wl2)
    mA = (w / wl[:, None] * d).sum(1) + 0.5 * (R * INV_T - wl)
    difSq = distSq - RSq
    with np.errstate(all='ignore'):
        cr = np.cross(p, v)
        wwSq = (cr * cr).sum(1) / difSq
        pl = np.sqrt(distSq)
        t = ((p * v).sum(1) / pl + np.sqrt(wwSq) * (R / pl)) / pl
        ww = v - t[:, None] * p
        wwl = np.sqrt((ww * ww).sum(1))
        mB = (ww / wwl[:, None] * d).sum(1) + 0.5 * (R * t - wwl)
    nocoll = distSq > RSq
    m = np.where(nocoll, np.where(capA, mA, mB), mc)
    return m


def margin32(p, v, R, d):
    """The kernel's fp32 sequence -> (m32, guards_ok)."""
    p, v, d, R = p.astype(F), v.astype(F), d.astype(F), R.astype(F)
    dot = lambda a, b: (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]
    distSq = dot(p, p)
    RSq = R * R
    with np.errstate(all='ignore'):
        w = v - F(INV_DT) * p
        wl = np.sqrt(dot(w, w))
        mc = dot(w / wl[:, None], d) + F(0.5) * (R * F(INV_DT) - wl)
        okc = wl >= F(WL_MIN)
        w = v - F(INV_T) * p
        wl2 = dot(w, w)
        dp = dot(w, p)
        capA = (dp < 0) & (dp * dp > RSq * wl2)
        wl = np.sqrt(wl2)
        mA = dot(w / wl[:, None], d) + F(0.5) * (R * F(INV_T) - wl)
        okA = wl >= F(WL_MIN)
        difSq = distSq - RSq
        cr = np.stack([p[:, 1] * v[:, 2] - p[:, 2] * v[:, 1], p[:, 2] * v[:, 0] - p[:, 0] * v[:, 2],
                       p[:, 0] * v[:, 1] - p[:, 1] * v[:, 0]], 1)
        crl = np.sqrt(dot(cr, cr))
        pl = np.sqrt(distSq)
        t = (dot(p, v) / pl + (crl / np.sqrt(difSq)) * (R / pl)) / pl
        ww = v - t[:, None] * p
        wwl = np.sqrt(dot(ww, ww))
        mB = dot(ww / wwl[:, None], d) + F(0.5) * (R * t - wwl)
        okB = wwl >= F(WL_MIN)
    nocoll = distSq > RSq
    m = np.where(nocoll, np.where(capA, mA, mB), mc)
    ok = np.where(nocoll, np.where(capA, okA, okB), okc) & np.isfinite(m) & (np.abs(distSq - RSq) >= F(DIF_MIN))
    return m.astype(np.float64), ok


def sample(rng, n, kind):
    unit = lambda a: a / np.linalg.norm(a, axis=1, keepdims=True)
    me = rng.uniform(0, 600, (n, 3))
    dirp = unit(rng.normal(size=(n, 3)))
    if kind == 'uniform':
        dist = rng.uniform(0.9, 10.0, n)
    elif kind == 'touching':
        dist = 1.1 + rng.normal(0, 0.05, n)
        dist[::4] = 1.1 + rng.normal(0, 1e-7, len(dist[::4]))      # on the collision / no-collision switch
    else:
        dist = rng.uniform(1.2, 10.0, n)
    p_true = dirp * dist[:, None]
    other = me + p_true
    p = other - me                                       # the kernel's fp64 subtraction of fp64 positions
    vA = (unit(rng.normal(size=(n, 3))) * rng.choice([0.3, 0.5, 1.0], n)[:, None] * rng.uniform(0.9, 1.0, n)[:, None]).astype(F)
    vB = (unit(rng.normal(size=(n, 3))) * rng.choice([0.0, 0.5, 1.0], n)[:, None]).astype(F)
    if kind == 'headon':                                 # relative velocity (anti)parallel to p: |p x v| ~ 0
        s = rng.choice([-1.0, 1.0], n)[:, None] * rng.uniform(0.1, 2.0, n)[:, None]
        vA = (dirp * s * 0.5 + rng.normal(0, 1e-4, (n, 3))).astype(F)
        vB = (-dirp * s * 0.5).astype(F)
    if kind == 'smallw':                                 # v ~ p / tau: |w| ~ 0 (cut-off centre)
        vB = np.zeros((n, 3), F)
        vA = (p * INV_T + rng.normal(0, 1e-3, (n, 3))).astype(F)
    if kind == 'capedge':                                # near the cap / cone boundary: (w.p)^2 ~ R^2 |w|^2
        # w at angle acos(R/|p|) from -p (+- noise)
        ang = np.arccos(np.clip(1.1 / dist, -1, 1)) + rng.normal(0, 1e-3, n)
        perp = unit(np.cross(dirp, rng.normal(size=(n, 3))))
        wdir = -dirp * np.cos(ang)[:, None] + perp * np.sin(ang)[:, None]
        w = wdir * rng.uniform(0.05, 1.0, n)[:, None]
        vA = (w + p * INV_T).astype(F)
        vB = np.zeros((n, 3), F)
    v = (vA - vB).astype(F)                              # fp32 subtraction, as the reference does
    v0 = unit(rng.normal(size=(n, 3))) * rng.uniform(0, 1, n)[:, None]
    if kind != 'uniform':
        v0 = vA.astype(np.float64) + rng.normal(0, 0.2, (n, 3))
    d = vA.astype(np.float64) - v0
    R = np.full(n, (0.5 + 0.05) + (0.5 + 0.05))
    return p, v, R, d


def main(n=None, quiet=False):
    n = n or (int(sys.argv[1]) if len(sys.argv) > 1 else 2000000)
    rng = np.random.default_rng(0)
    worst, wrong = 0.0, 0
    for kind in ('uniform', 'far', 'touching', 'headon', 'smallw', 'capedge'):
        p, v, R, d = sample(rng, n, kind)
        m64 = margin64(p, v.astype(np.float64), R, d)
        m32, ok = margin32(p, v, R, d)
        err = np.abs(m32 - m64)
        bad_safe = ok & (m32 < -EPS) & ~(m64 <= 0)        # the only claim the kernel trusts
        e = float(err[ok].max()) if ok.any() else 0.0
        worst = max(worst, e)
        wrong += int(bad_safe.sum())
        if not quiet:
            print('%-9s guards pass %5.1f %%  max|m32-m64| %.3e  p99.99 %.3e  trusted-safe %5.1f %%  in band %5.2f %%  wrong-safe %d'
                  % (kind, 100 * ok.mean(), e, float(np.quantile(err[ok], 0.9999)), 100 * (ok & (m32 < -EPS)).mean(),
                     100 * (ok & (np.abs(m32) <= EPS)).mean(), int(bad_safe.sum())))
    if not quiet:
        print('worst error under the guards: %.3e   EPS = %.1e   safety factor %.1f' % (worst, EPS, EPS / worst))
    return worst, wrong


if __name__ == '__main__':
    main()
