// scab_policy_lp.cu -- stage 3 for ORCA3D with the linear programs, one thread per agent, fp64:
//   planes        mamp/policies/orca3dPolicyOfficial.py:56-106
//   linearProgram1 :126-181, linearProgram2 :183-237, linearProgram3 :239-259, linearProgram4 :261-300
// fused with cartesian2spherical (:331-342, the `distance` variant) and the kinematic update
// (mamp/envs/mampenv.py:83-105).
//
// The reference's deviations from RVO2-3D are reproduced on purpose: the LP4 entry
// test at :264 is np.dot(normal, (point - v) > 0.0) -- a dot with a BOOLEAN vector -- and LP4 keeps no
// running distance.  Plane order == neighbour order, which is why the neighbour stage is order-exact.
// Planes live in shared memory (16 x 48 B per agent, transposed so that lanes hit distinct banks).
#include "scab_common.cuh"

#define LP_THREADS 64
#define LP_FAST_THREADS 128

struct Planes {
    // [component][plane][thread] : consecutive threads -> consecutive 8-byte words (conflict-free)
    double v[6][KNBR][LP_THREADS];
};

struct PlaneView {
    double *base;  // &P.v[0][0][tid] (shared, stride LP_THREADS) or a thread-local array (stride 1)
    int stride;
    __device__ __forceinline__ double &at(int comp, int k) const { return base[(comp * KNBR + k) * stride]; }
    __device__ __forceinline__ double px(int k) const { return at(0, k); }
    __device__ __forceinline__ double py(int k) const { return at(1, k); }
    __device__ __forceinline__ double pz(int k) const { return at(2, k); }
    __device__ __forceinline__ double nx(int k) const { return at(3, k); }
    __device__ __forceinline__ double ny(int k) const { return at(4, k); }
    __device__ __forceinline__ double nz(int k) const { return at(5, k); }
};

struct Vec3 { double x, y, z; };

// Compact planes (MODE 4, SCAB_LP_COMPACT=1): an ORCA plane is point = vA + u / 2 with u = s n (:104-105), so (n, s) -- 32 B
// instead of 48 B -- determines it, and the point is rebuilt on every read with the roundings of the original construction
// (s * n, * 0.5, + vA).  Two thirds of the shared memory per thread: 6-7 resident CTAs per SM instead of 4 in the list pass.
struct PlanesC {
    double v[4][KNBR][LP_THREADS];
};
struct PlaneViewC {
    double *base;   // &P.v[0][0][tid], stride LP_THREADS
    double vax, vay, vaz;
    __device__ __forceinline__ double &at(int comp, int k) const { return base[(comp * KNBR + k) * LP_THREADS]; }
};

__device__ __forceinline__ double dotv(const Vec3 &a, const Vec3 &b) { return dot3(a.x, a.y, a.z, b.x, b.y, b.z); }
__device__ __forceinline__ Vec3 crossv(const Vec3 &a, const Vec3 &b)
{   // np.cross: products and differences rounded separately
    Vec3 c;
    c.x = __dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y));
    c.y = __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z));
    c.z = __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    return c;
}
__device__ __forceinline__ Vec3 P_point(const PlaneView &P, int k) { return Vec3{P.px(k), P.py(k), P.pz(k)}; }
__device__ __forceinline__ Vec3 P_normal(const PlaneView &P, int k) { return Vec3{P.nx(k), P.ny(k), P.nz(k)}; }
__device__ __forceinline__ Vec3 P_normal(const PlaneViewC &P, int k) { return Vec3{P.at(0, k), P.at(1, k), P.at(2, k)}; }
__device__ __forceinline__ Vec3 P_point(const PlaneViewC &P, int k)
{
    const double sc = P.at(3, k);
    return Vec3{__dadd_rn(P.vax, __dmul_rn(0.5, __dmul_rn(sc, P.at(0, k)))), __dadd_rn(P.vay, __dmul_rn(0.5, __dmul_rn(sc, P.at(1, k)))),
                __dadd_rn(P.vaz, __dmul_rn(0.5, __dmul_rn(sc, P.at(2, k))))};
}
__device__ __forceinline__ Vec3 subv(const Vec3 &a, const Vec3 &b) { return Vec3{__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ Vec3 axpy(double s, const Vec3 &d, const Vec3 &p)
{   // p + s * d with the product rounded first (NumPy ufunc order)
    return Vec3{__dadd_rn(p.x, __dmul_rn(s, d.x)), __dadd_rn(p.y, __dmul_rn(s, d.y)), __dadd_rn(p.z, __dmul_rn(s, d.z))};
}

// orca3dPolicyOfficial.py:126-181
template <class PV>
__device__ bool lp1(const PV &P, int planeNo, const Vec3 &lpoint, const Vec3 &ldir, double maxSpeed,
                    const Vec3 &vel_pref, bool dir_opt, Vec3 &res)
{
    const double dp = dotv(lpoint, ldir);
    const double disc = __dsub_rn(__dadd_rn(__dmul_rn(dp, dp), __dmul_rn(maxSpeed, maxSpeed)), dotv(lpoint, lpoint));
    if (disc < 0.0) return false;
    const double sq = __dsqrt_rn(disc);
    double tLeft = __dsub_rn(-dp, sq), tRight = __dadd_rn(-dp, sq);
    for (int i = 0; i < planeNo; i++) {
        const Vec3 n = P_normal(P, i);
        const double num = dotv(subv(P_point(P, i), lpoint), n);
        const double den = dotv(ldir, n);
        if (__dmul_rn(den, den) <= RVO_EPS) {
            if (num > 0.0) return false;
            continue;
        }
        const double t = __ddiv_rn(num, den);
        if (den >= 0.0) tLeft = fmax(tLeft, t);
        else tRight = fmin(tRight, t);
        if (tLeft > tRight) return false;
    }
    double tt;
    if (dir_opt) {
        tt = (dotv(vel_pref, ldir) > 0.0) ? tRight : tLeft;
    } else {
        const double t = dotv(ldir, subv(vel_pref, lpoint));
        tt = (t < tLeft) ? tLeft : (t > tRight ? tRight : t);
    }
    res = axpy(tt, ldir, lpoint);
    return true;
}

// orca3dPolicyOfficial.py:183-237
template <class PV>
__device__ bool lp2(const PV &P, int planeNo, double maxSpeed, const Vec3 &vel_pref, bool dir_opt, Vec3 &res)
{
    const Vec3 pn = P_normal(P, planeNo), pp = P_point(P, planeNo);
    const double planeDist = dotv(pp, pn);
    const double planeDistSq = __dmul_rn(planeDist, planeDist);
    const double radiusSq = __dmul_rn(maxSpeed, maxSpeed);
    if (planeDistSq > radiusSq) return false;
    const double planeRadiusSq = __dsub_rn(radiusSq, planeDistSq);
    const Vec3 center{__dmul_rn(planeDist, pn.x), __dmul_rn(planeDist, pn.y), __dmul_rn(planeDist, pn.z)};
    if (dir_opt) {
        const double dn = dotv(vel_pref, pn);
        const Vec3 pov{__dsub_rn(vel_pref.x, __dmul_rn(dn, pn.x)), __dsub_rn(vel_pref.y, __dmul_rn(dn, pn.y)),
                       __dsub_rn(vel_pref.z, __dmul_rn(dn, pn.z))};
        const double povSq = dotv(pov, pov);
        if (povSq <= RVO_EPS) res = center;
        else res = axpy(__dsqrt_rn(__ddiv_rn(planeRadiusSq, povSq)), pov, center);
    } else {
        const double dpn = dotv(subv(pp, vel_pref), pn);
        res = axpy(dpn, pn, vel_pref);
        if (dotv(res, res) > radiusSq) {
            const Vec3 pr = subv(res, center);
            const double prSq = dotv(pr, pr);
            res = axpy(__dsqrt_rn(__ddiv_rn(planeRadiusSq, prSq)), pr, center);
        }
    }
    for (int i = 0; i < planeNo; i++) {
        const Vec3 ni = P_normal(P, i), pi = P_point(P, i);
        if (dotv(ni, subv(pi, res)) > 0.0) {
            const Vec3 cp = crossv(ni, pn);
            const double cpSq = dotv(cp, cp);
            if (cpSq <= RVO_EPS) return false;
            const double ncp = __dsqrt_rn(cpSq);
            const Vec3 ldir{__ddiv_rn(cp.x, ncp), __ddiv_rn(cp.y, ncp), __ddiv_rn(cp.z, ncp)};
            const Vec3 lnormal = crossv(ldir, pn);
            const double dp1 = dotv(subv(pi, pp), ni);
            const double dp2 = dotv(lnormal, ni);
            const Vec3 lpoint = axpy(__ddiv_rn(dp1, dp2), lnormal, pp);
            if (!lp1(P, i, lpoint, ldir, maxSpeed, vel_pref, dir_opt, res)) return false;
        }
    }
    return true;
}

// orca3dPolicyOfficial.py:239-259
template <class PV>
__device__ int lp3(const PV &P, int np, double maxSpeed, const Vec3 &vel_pref, bool dir_opt, Vec3 &res)
{
    if (dir_opt) {
        res = Vec3{__dmul_rn(vel_pref.x, maxSpeed), __dmul_rn(vel_pref.y, maxSpeed), __dmul_rn(vel_pref.z, maxSpeed)};
    } else if (dotv(vel_pref, vel_pref) > __dmul_rn(maxSpeed, maxSpeed)) {
        const double nv = __dsqrt_rn(dotv(vel_pref, vel_pref));
        res = Vec3{__dmul_rn(__ddiv_rn(vel_pref.x, nv), maxSpeed), __dmul_rn(__ddiv_rn(vel_pref.y, nv), maxSpeed),
                   __dmul_rn(__ddiv_rn(vel_pref.z, nv), maxSpeed)};
    } else {
        res = vel_pref;
    }
    for (int i = 0; i < np; i++) {
        if (dotv(P_normal(P, i), subv(P_point(P, i), res)) > 0.0) {
            const Vec3 temp = res;
            if (!lp2(P, i, maxSpeed, vel_pref, dir_opt, res)) {
                res = temp;
                return i;
            }
        }
    }
    return np;
}

// orca3dPolicyOfficial.py:261-300.  `Q` is a second plane set (the projected planes).
template <class PV>
__device__ void lp4(const PV &P, const PlaneView &Q, int np, int beginPlane, double radius, Vec3 &res, int &n_proj)
{
    for (int i = beginPlane; i < np; i++) {
        const Vec3 ni = P_normal(P, i), pi = P_point(P, i);
        // :264 np.dot(normal, (point - v) > 0.0): dot with a boolean vector, truthy iff != 0
        const double bsum = dot3(ni.x, ni.y, ni.z, (__dsub_rn(pi.x, res.x) > 0.0) ? 1.0 : 0.0,
                                 (__dsub_rn(pi.y, res.y) > 0.0) ? 1.0 : 0.0, (__dsub_rn(pi.z, res.z) > 0.0) ? 1.0 : 0.0);
        if (bsum != 0.0) {
            int nq = 0;
            for (int j = 0; j < i; j++) {
                const Vec3 nj = P_normal(P, j), pj = P_point(P, j);
                const Vec3 cp = crossv(nj, ni);
                Vec3 pt;
                if (dotv(cp, cp) <= RVO_EPS) {
                    if (dotv(ni, nj) > 0.0) continue;
                    pt = Vec3{__dmul_rn(0.5, __dadd_rn(pi.x, pj.x)), __dmul_rn(0.5, __dadd_rn(pi.y, pj.y)),
                              __dmul_rn(0.5, __dadd_rn(pi.z, pj.z))};
                } else {
                    const Vec3 ln = crossv(cp, ni);
                    const double dp1 = dotv(subv(pj, pi), nj);
                    const double dp2 = dotv(ln, nj);
                    pt = axpy(__ddiv_rn(dp1, dp2), ln, pi);
                }
                const Vec3 dn = subv(nj, ni);
                const double nn = __dsqrt_rn(dotv(dn, dn));
                Q.at(0, nq) = pt.x; Q.at(1, nq) = pt.y; Q.at(2, nq) = pt.z;
                Q.at(3, nq) = __ddiv_rn(dn.x, nn); Q.at(4, nq) = __ddiv_rn(dn.y, nn); Q.at(5, nq) = __ddiv_rn(dn.z, nn);
                nq++;
            }
            const Vec3 temp = res;
            n_proj++;
            if (lp3(Q, nq, radius, ni, true, res) < nq) res = temp;
        }
    }
}

// ORCA plane of neighbour slot `e` for the agent (me, mv): orca3dPolicyOfficial.py:58-105.
// point = vA + u/2 for every neighbour, static or not.
__device__ __forceinline__ void orca_plane(const StepArgs &A, int e, const double4 &me, const float4 &mv, Vec3 &point, Vec3 &n,
                                           double *sc_out = nullptr)
{
    double4 o;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (e < A.n) {
        o = A.s_posr[e];
        const float4 ov = A.s_vel[e];
        ox = ov.x; oy = ov.y; oz = ov.z;
    } else if (e < A.n + A.m) o = A.o_posr[e - A.n];
    else o = A.o_posr_by_id[e - A.n - A.m];
    const Vec3 rp{__dsub_rn(o.x, me.x), __dsub_rn(o.y, me.y), __dsub_rn(o.z, me.z)};
    const Vec3 rv{(double)__fsub_rn(mv.x, ox), (double)__fsub_rn(mv.y, oy), (double)__fsub_rn(mv.z, oz)};
    const double distSq = dotv(rp, rp);
    const double R = __dadd_rn(__dadd_rn(me.w, 0.05), __dadd_rn(o.w, 0.05));
    const double RSq = __dmul_rn(R, R);
    Vec3 u;
    double scv;      // u = scv * n
    if (distSq > RSq) {
        const double invT = A.inv_time_horizon;
        const Vec3 w{__dsub_rn(rv.x, __dmul_rn(invT, rp.x)), __dsub_rn(rv.y, __dmul_rn(invT, rp.y)),
                     __dsub_rn(rv.z, __dmul_rn(invT, rp.z))};
        const double wl2 = dotv(w, w);
        const double dp = dotv(w, rp);
        if (dp < 0.0 && __dmul_rn(dp, dp) > __dmul_rn(RSq, wl2)) {
            const double wl = __dsqrt_rn(wl2);
            n = Vec3{__ddiv_rn(w.x, wl), __ddiv_rn(w.y, wl), __ddiv_rn(w.z, wl)};
            const double sc = __dsub_rn(__dmul_rn(R, invT), wl);
            u = Vec3{__dmul_rn(sc, n.x), __dmul_rn(sc, n.y), __dmul_rn(sc, n.z)};
            scv = sc;
        } else {
            const double difSq = __dsub_rn(distSq, RSq);
            const double dpr = dotv(rp, rv);
            const Vec3 cr = crossv(rp, rv);
            const double wwSq = __ddiv_rn(dotv(cr, cr), difSq);
            const double pl = __dsqrt_rn(distSq);
            const double t = __ddiv_rn(__dadd_rn(__ddiv_rn(dpr, pl), __dmul_rn(__dsqrt_rn(wwSq), __ddiv_rn(R, pl))), pl);
            const Vec3 ww{__dsub_rn(rv.x, __dmul_rn(t, rp.x)), __dsub_rn(rv.y, __dmul_rn(t, rp.y)),
                          __dsub_rn(rv.z, __dmul_rn(t, rp.z))};
            const double wl = __dsqrt_rn(dotv(ww, ww));
            n = Vec3{__ddiv_rn(ww.x, wl), __ddiv_rn(ww.y, wl), __ddiv_rn(ww.z, wl)};
            const double sc = __dsub_rn(__dmul_rn(R, t), wl);
            u = Vec3{__dmul_rn(sc, n.x), __dmul_rn(sc, n.y), __dmul_rn(sc, n.z)};
            scv = sc;
        }
    } else {
        const double invDt = A.inv_time_step;
        const Vec3 w{__dsub_rn(rv.x, __dmul_rn(invDt, rp.x)), __dsub_rn(rv.y, __dmul_rn(invDt, rp.y)),
                     __dsub_rn(rv.z, __dmul_rn(invDt, rp.z))};
        const double wl = __dsqrt_rn(dotv(w, w));
        n = Vec3{__ddiv_rn(w.x, wl), __ddiv_rn(w.y, wl), __ddiv_rn(w.z, wl)};
        const double sc = __dsub_rn(__dmul_rn(R, invDt), wl);
        u = Vec3{__dmul_rn(sc, n.x), __dmul_rn(sc, n.y), __dmul_rn(sc, n.z)};
            scv = sc;
    }
    point = Vec3{__dadd_rn((double)mv.x, __dmul_rn(0.5, u.x)), __dadd_rn((double)mv.y, __dmul_rn(0.5, u.y)),
                 __dadd_rn((double)mv.z, __dmul_rn(0.5, u.z))};
    if (sc_out) *sc_out = scv;
}

// fp32 pre-filter of the register-only pass.  That pass only needs the SIGN of m = n . (point - v0) per plane (:252),
// and m = n . (vA - v0) + s / 2 with u = s n.  The sequence below is replayed in NumPy float32 against float64 by
// tools/lp_prefilter_check.py on random and adversarial configurations: under the guards the largest |m32 - m64| seen is
// 2.2e-5, so a margin below -LP_F32_EPS (~100x that) is trusted as "satisfied"; a margin above +LP_F32_EPS defers the agent
// to the list pass (always exact); anything else -- inside the band, an ill-conditioned construction, a NaN -- is
// re-evaluated with the reference's fp64 expressions.  No FMA contraction here (-fmad=false), IEEE div / sqrt.
#define LP_F32_EPS 2e-3f
#define LP_F32_WL_MIN 1e-2f
#define LP_F32_DIF_MIN 5e-2f
__device__ __forceinline__ float dot3f(float a0, float a1, float a2, float b0, float b1, float b2) { return (a0 * b0 + a1 * b1) + a2 * b2; }

// 0: certainly satisfied, 1: certainly violated, 2: ask fp64.  (dx, dy, dz) = (float)(vA - v0)
__device__ __forceinline__ int orca_margin_f32(const StepArgs &A, int e, const double4 &me, const float4 &mv, float dx, float dy, float dz)
{
    double4 o;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (e < A.n) {
        o = A.s_posr[e];
        const float4 ov = A.s_vel[e];
        ox = ov.x; oy = ov.y; oz = ov.z;
    } else if (e < A.n + A.m) o = A.o_posr[e - A.n];
    else o = A.o_posr_by_id[e - A.n - A.m];
    const float px = (float)__dsub_rn(o.x, me.x), py = (float)__dsub_rn(o.y, me.y), pz = (float)__dsub_rn(o.z, me.z);
    const float vx = mv.x - ox, vy = mv.y - oy, vz = mv.z - oz;
    const float R = (float)__dadd_rn(__dadd_rn(me.w, 0.05), __dadd_rn(o.w, 0.05));
    const float distSq = dot3f(px, py, pz, px, py, pz), RSq = R * R;
    float nx, ny, nz, wl, sc;
    bool ok;
    if (distSq > RSq) {
        const float invT = (float)A.inv_time_horizon;
        const float wx = vx - invT * px, wy = vy - invT * py, wz = vz - invT * pz;
        const float wl2 = dot3f(wx, wy, wz, wx, wy, wz), dp = dot3f(wx, wy, wz, px, py, pz);
        if (dp < 0.f && dp * dp > RSq * wl2) {
            wl = sqrtf(wl2);
            nx = wx / wl; ny = wy / wl; nz = wz / wl;
            sc = R * invT - wl;
            ok = wl >= LP_F32_WL_MIN;
        } else {
            const float difSq = distSq - RSq;
            const float cx = py * vz - pz * vy, cy = pz * vx - px * vz, cz = px * vy - py * vx;
            const float crl = sqrtf(dot3f(cx, cy, cz, cx, cy, cz));
            const float pl = sqrtf(distSq);
            const float t = (dot3f(px, py, pz, vx, vy, vz) / pl + (crl / sqrtf(difSq)) * (R / pl)) / pl;
            const float ax = vx - t * px, ay = vy - t * py, az = vz - t * pz;
            wl = sqrtf(dot3f(ax, ay, az, ax, ay, az));
            nx = ax / wl; ny = ay / wl; nz = az / wl;
            sc = R * t - wl;
            ok = wl >= LP_F32_WL_MIN;
        }
    } else {
        const float invDt = (float)A.inv_time_step;
        const float wx = vx - invDt * px, wy = vy - invDt * py, wz = vz - invDt * pz;
        wl = sqrtf(dot3f(wx, wy, wz, wx, wy, wz));
        nx = wx / wl; ny = wy / wl; nz = wz / wl;
        sc = R * invDt - wl;
        ok = wl >= LP_F32_WL_MIN;
    }
    // the collision / no-collision switch at |p| = R is a discontinuity of the construction: stay clear of it in every branch
    ok = ok && fabsf(distSq - RSq) >= LP_F32_DIF_MIN;
    const float m = dot3f(nx, ny, nz, dx, dy, dz) + 0.5f * sc;
    if (ok && m < -LP_F32_EPS) return 0;      // a NaN fails both comparisons
    if (ok && m > LP_F32_EPS) return 1;
    return 2;
}

// linearProgram3's starting point (:240-250, dir_opt False): v_pref clamped to the maxSpeed sphere
__device__ __forceinline__ Vec3 lp3_start(const Vec3 &vel_pref, double maxSpeed)
{
    if (dotv(vel_pref, vel_pref) > __dmul_rn(maxSpeed, maxSpeed)) {
        const double nv = __dsqrt_rn(dotv(vel_pref, vel_pref));
        return Vec3{__dmul_rn(__ddiv_rn(vel_pref.x, nv), maxSpeed), __dmul_rn(__ddiv_rn(vel_pref.y, nv), maxSpeed),
                    __dmul_rn(__ddiv_rn(vel_pref.z, nv), maxSpeed)};
    }
    return vel_pref;
}

// MODE 0: every agent, planes in shared memory (SCAB_NO_FAST=1, the A/B switch).
// MODE 1: fast path, no plane storage.  In linearProgram3 (:251-258) a plane is only ever re-read after SOME plane was
//         violated by the running result; until then the result is the starting point.  So each plane is built in
//         registers and tested against the starting point; if none is violated that IS the answer (planeFail == nc).
//         The first violation appends the agent to A.slow_list and leaves every output untouched.
//         Each plane is first judged in fp32 (orca_margin_f32); only the undecided ones are built in fp64.
// MODE 2: the agents on A.slow_list, planes in shared memory, full LP3 / LP4.
// MODE 3: MODE 1 without the fp32 pre-filter (SCAB_LP_F64=1, the A/B switch for it).
// MODE 5: MODE 1 compiled for 8 resident CTAs per SM (64 registers, a few spills; SCAB_LP_PASS1_CTAS8=1) -- to be timed.
// MODE 4: MODE 2 with compact planes (SCAB_LP_COMPACT=1; PlaneViewC) -- not the default until it has been timed on the GPU.
template <int MODE>
__global__ void __launch_bounds__((MODE & 1) ? LP_FAST_THREADS : LP_THREADS, (MODE == 5) ? 8 : ((MODE & 1) ? 6 : 1)) k_policy_lp(StepArgs A)
{
    extern __shared__ double lp_smem[];
    const int tid = threadIdx.x;
    unsigned long long ran = 0;
    constexpr bool LIST = (MODE == 2 || MODE == 4);
    const int n_work = LIST ? min(*A.slow_count, A.n) : policy_items(A);
    if (LIST && blockIdx.x == 0 && tid == 0) atomicAdd(&A.counters[6], (unsigned long long)n_work);
    for (int wi = blockIdx.x * blockDim.x + tid; wi < n_work; wi += gridDim.x * blockDim.x) {
        const int s = LIST ? A.slow_list[wi] : policy_item(A, wi);
        const int id = A.s_id[s];
        const double4 me = A.s_posr[s];
        const float4 mv = A.s_vel[s];
        const int flags = __float_as_int(mv.w);
        Action7 act;
        double npx, npy, npz;
        if (flags & SCAB_ST_GHOST) continue;   // halo copy: owned (and stepped) by another GPU
        if (flags & FL_INACTIVE) {
#pragma unroll
            for (int k = 0; k < 7; k++) act.a[k] = 0.f;
            kinematic_update(A, id, act, me.x, me.y, me.z, &npx, &npy, &npz);
            if (A.st.tag) A.st.tag[id] = 0;
            if (A.st.plane_fail) A.st.plane_fail[id] = -1;
            if (A.st.n_planes) A.st.n_planes[id] = -1;
            if (A.st.lp4_proj) A.st.lp4_proj[id] = 0;
            continue;
        }
        double v_store[3], v_des[3];
        double target[3];
        // the fast pass may still defer this agent: it only peeks at the waypoint state and commits the pop when it finishes
        waypoint_target(A.st, id, me.x, me.y, me.z, true, target, !(MODE & 1));
        v_pref_straight(&A.st.goal[3 * id], target, me.x, me.y, me.z, A.st.pref_speed[id], true, v_store, v_des);
        const double nv32 = __dsqrt_rn(sumsq_f32(mv.x, mv.y, mv.z));
        const bool first = py_round5(__dadd_rn(nv32, 1e-5)) <= 1e-5;   // :52 distance(vel, 0) <= 1e-5
        int tag = 0, plane_fail = -1, n_planes = -1, n_proj = 0;
        Vec3 res;
        if (first) {
            tag = SCAB_TAG_FIRST;
            res = Vec3{__dmul_rn(0.3, v_des[0]), __dmul_rn(0.3, v_des[1]), __dmul_rn(0.3, v_des[2])};
        } else {
            const int nc = A.s_nbr_cnt[s];
            const Vec3 vp{v_des[0], v_des[1], v_des[2]};
            // (requesting all 16 neighbour records towards L1 up front measured SLOWER on B200: 0.536 -> 0.578 ms for the stage --
            //  most agents leave the plane loop early, so the extra requests are wasted traffic)
            if (MODE & 1) {
                res = lp3_start(vp, A.max_speed);
                bool violated = false;
                unsigned ask = (MODE == 3) ? ((nc >= 32) ? 0xffffffffu : ((1u << nc) - 1u)) : 0u;   // planes to build in fp64
                if (MODE == 1 || MODE == 5) {
                    const float dx = (float)__dsub_rn((double)mv.x, res.x), dy = (float)__dsub_rn((double)mv.y, res.y),
                                dz = (float)__dsub_rn((double)mv.z, res.z);
                    for (int k = 0; k < nc && !violated; k++) {
                        const int c = orca_margin_f32(A, A.s_nbr[(size_t)s * KNBR + k], me, mv, dx, dy, dz);
                        violated = c == 1;
                        if (c == 2) ask |= 1u << k;
                    }
                }
                while (ask && !violated) {
                    const int k = __ffs(ask) - 1;
                    ask &= ask - 1u;
                    Vec3 pt, n;
                    orca_plane(A, A.s_nbr[(size_t)s * KNBR + k], me, mv, pt, n);
                    violated = dotv(n, subv(pt, res)) > 0.0;            // :252
                }
                if (violated) {
                    A.slow_list[atomicAdd(A.slow_count, 1)] = s;
                    continue;
                }
                plane_fail = nc;
                tag |= SCAB_TAG_LP3_OK;
            } else if (MODE == 4) {
                PlanesC *PS = reinterpret_cast<PlanesC *>(lp_smem);
                double qloc[6 * KNBR];   // projected planes of linearProgram4: rare path, thread-local
                PlaneViewC P{&PS[0].v[0][0][tid], (double)mv.x, (double)mv.y, (double)mv.z};
                PlaneView Q{qloc, 1};
                for (int k = 0; k < nc; k++) {
                    Vec3 pt, n;
                    double sc;
                    orca_plane(A, A.s_nbr[(size_t)s * KNBR + k], me, mv, pt, n, &sc);
                    P.at(0, k) = n.x; P.at(1, k) = n.y; P.at(2, k) = n.z; P.at(3, k) = sc;
                }
                plane_fail = lp3(P, nc, A.max_speed, vp, false, res);   // :108
                if (plane_fail == nc) tag |= SCAB_TAG_LP3_OK;
                if (plane_fail < nc) {                                   // :110-111
                    tag |= SCAB_TAG_LP4;
                    lp4(P, Q, nc, plane_fail, A.max_speed, res, n_proj);
                }
            } else {
                Planes *PS = reinterpret_cast<Planes *>(lp_smem);
                double qloc[6 * KNBR];   // projected planes of linearProgram4: rare path, thread-local
                PlaneView P{&PS[0].v[0][0][tid], LP_THREADS}, Q{qloc, 1};
                for (int k = 0; k < nc; k++) {
                    Vec3 pt, n;
                    orca_plane(A, A.s_nbr[(size_t)s * KNBR + k], me, mv, pt, n);
                    P.at(0, k) = pt.x; P.at(1, k) = pt.y; P.at(2, k) = pt.z;
                    P.at(3, k) = n.x; P.at(4, k) = n.y; P.at(5, k) = n.z;
                }
                plane_fail = lp3(P, nc, A.max_speed, vp, false, res);   // :108
                if (plane_fail == nc) tag |= SCAB_TAG_LP3_OK;
                if (plane_fail < nc) {                                   // :110-111
                    tag |= SCAB_TAG_LP4;
                    lp4(P, Q, nc, plane_fail, A.max_speed, res, n_proj);
                }
            }
            n_planes = nc;
        }
        ran++;
        if ((MODE & 1) && A.st.waypoints != nullptr) waypoint_target(A.st, id, me.x, me.y, me.z, true, target, true);
        act = cartesian2spherical(res.x, res.y, res.z, A.st.heading[3 * id], A.st.heading[3 * id + 1], true);
        kinematic_update(A, id, act, me.x, me.y, me.z, &npx, &npy, &npz);
        if (A.s_qcoll[s]) A.st.status[id] = flags | SCAB_ST_COLLISION;   // agent.py:83-85
        if (A.st.vpref_out) { A.st.vpref_out[3 * id] = v_store[0]; A.st.vpref_out[3 * id + 1] = v_store[1]; A.st.vpref_out[3 * id + 2] = v_store[2]; }
        if (A.st.tag) A.st.tag[id] = tag;
        if (A.st.plane_fail) A.st.plane_fail[id] = plane_fail;
        if (A.st.n_planes) A.st.n_planes[id] = n_planes;
        if (A.st.lp4_proj) A.st.lp4_proj[id] = n_proj;
        if (A.st.n_suit) A.st.n_suit[id] = 0;
    }
    // one counter update per warp (every thread of the warp leaves the loop before this point)
    __syncwarp();
    for (int o = 16; o; o >>= 1) ran += __shfl_xor_sync(0xffffffffu, ran, o);
    if ((threadIdx.x & 31) == 0 && ran) {
        atomicAdd(&A.counters[0], ran);
        atomicAdd(&A.counters[4], ran);
    }
}

#ifndef SCAB_HOSTSIM   // tests/hostsim compiles the device code above for the host (one thread at a time); no launcher there
cudaError_t launch_policy_lp(const StepArgs &A, cudaStream_t s)
{
    if (A.n <= 0 || A.row_count == 0) return cudaSuccess;
    const int items = A.row_count < 0 ? A.n : A.row_count;
    LaunchCfg &cfg = *A.cfg;      // this handle's device and switches (slots 12, 13: the two list-pass kernels)
    int &per_sm = cfg.occ[12], &per_sm_c = cfg.occ[13];
    const int no_fast = cfg.no_fast, sms = cfg.sms, f64_only = cfg.lp_f64, compact = cfg.lp_compact, ctas8 = cfg.lp_ctas8;
    const size_t smem = sizeof(Planes), smem_c = sizeof(PlanesC);   // 6*16*64*8 = 49152 B, 4*16*64*8 = 32768 B
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_policy_lp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_policy_lp<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_policy_lp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_policy_lp<2>, LP_THREADS, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_policy_lp<4>, LP_THREADS, smem_c);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (per_sm_c < 1) per_sm_c = 1;
    }
    const int nb = (items + LP_THREADS - 1) / LP_THREADS;
    if (no_fast) {
        k_policy_lp<0><<<nb, LP_THREADS, smem, s>>>(A);
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(A.slow_count, 0, sizeof(int32_t), s);
    if (e != cudaSuccess) return e;
    const int nbf = (items + LP_FAST_THREADS - 1) / LP_FAST_THREADS;
    if (f64_only) k_policy_lp<3><<<nbf, LP_FAST_THREADS, 0, s>>>(A);
    else if (ctas8) k_policy_lp<5><<<nbf, LP_FAST_THREADS, 0, s>>>(A);
    else k_policy_lp<1><<<nbf, LP_FAST_THREADS, 0, s>>>(A);
    // the agents with a violated plane: one resident wave, strided over the list
    if (compact) {
        const int grid_c = nb < sms * per_sm_c ? nb : sms * per_sm_c;
        k_policy_lp<4><<<grid_c, LP_THREADS, smem_c, s>>>(A);
        return cudaGetLastError();
    }
    const int grid = nb < sms * per_sm ? nb : sms * per_sm;
    k_policy_lp<2><<<grid, LP_THREADS, smem, s>>>(A);
    return cudaGetLastError();
}
#endif
