// scab_policy_sampled.cu -- stage 3 for the five sampled-velocity policies:
//   SCA            mamp/policies/sca/scaPolicy.py:26-73, 119-240
//   S-RVO3D        mamp/policies/srvo3dPolicy.py:23-69, 100-231
//   RVO3D          mamp/policies/rvo3dPolicy.py:23-69, 102-196
//   RVO3D+Dubins   mamp/policies/sca/rvo3dDubinsPolicy.py:26-71, 119-194
//   ORCA sampled   mamp/policies/orca3dPolicy.py:38-120, 328-439
// as a chain of three kernels on the caller's stream:
//   k_vpref      thread/agent  v_pref (rvo3dPolicy.py:182-196) + the zero-velocity gate (scaPolicy.py:34)
//   k_select     WARP/agent    constraints + the 513-candidate search (the hot kernel)
//   k_integrate  thread/agent  cartesian2spherical (util.py:44-55) + kinematic update (mampenv.py:83-105)
//
// Candidate set (scaPolicy.py:190-211): two Fibonacci-sphere shells {0.5, pref_speed} x N unit directions
// (N = 256, or 128 for SCA when start->goal is vertical) + v_pref last.  k_select per warp:
//   A  posture test (util.py:6-20).  cos(angle(vA, rad*u)) does not depend on the shell radius, so one
//      fp32 dot per DIRECTION decides both shells; directions inside a 1e-6 guard band are re-evaluated
//      with the reference's fp64 expression.  Survivors (~15 % of the sphere) are ballot-compacted.
//   B  cone / half-space tests in fp32 with an error-bounded guard band; hits inside the band are
//      re-evaluated in fp64 with the reference's asin/acos form (util.py:30-41).
//   C  selection without sorting: the stable sort + prefix walk of shunted_strategy
//      (scaPolicy.py:119-145) is three lexicographic warp reductions over keys computed once in fp64
//      with the reference's decimal rounding (the keys are integers in units of 1e-5).
//   D  the "no suitable velocity" path (scaPolicy.py:148-165, 224-238) entirely in fp64.
#include "scab_common.cuh"
#include <limits.h>
#include <type_traits>

#define WARPS_PER_CTA 4
#define MAX_CAND 513
#define FULL 0xffffffffu

struct WarpSmem {
    // fp32 cones / planes, packed for 128-bit broadcast loads
    float4 cp[KNBR];   // cone: pB - pA (xyz), p.h (w)            | plane: normal (xyz), normal . point (w)
    float4 ch[KNBR];   // cone: h = 0.5 (vA + vB) (xyz), h.h (w)  | plane: unused
    float2 cg[KNBR];   // cone: max(|p|^2 - R^2, 0), guard band   | plane: -, guard band
    int cidx[KNBR];    // compacted slot -> neighbour index k (the fp64 copies below stay indexed by k)
    // fp64 exact copies for the guard-band recheck and the fallback path
    double dpx[KNBR], dpy[KNBR], dpz[KNBR];   // cone: pB (absolute)  | plane: point
    double dhx[KNBR], dhy[KNBR], dhz[KNBR];   // cone: p0 = pA (+ h)   | plane: normal
    double dR[KNBR];
    // ORCA-sampled fallback needs relpos / the v_dif offset too
    double rpx[KNBR], rpy[KNBR], rpz[KNBR];
    double ofx[KNBR], ofy[KNBR], ofz[KNBR];
    unsigned long long key[MAX_CAND + 7];   // per live slot: sort key (suitable: units of 1e-5; fallback: fp64 bits)
    unsigned int dvu[MAX_CAND + 7];         // per live slot: l3norm(v, vA) in units of 1e-5 (shunting)
    unsigned short list[MAX_CAND + 7];      // candidate index per slot; bit 15 = live (suitable)
};

struct Agent {
    double px, py, pz, r;
    float vx, vy, vz;
    double vdx, vdy, vdz;       // v_des (V_des, truncated)
    double pref;
    double nvA;                 // (double) np.linalg.norm(float32 vA)
    int N;                      // directions per shell
    const double *dir;          // [N,3]
    const float4 *dirf;
    double shell1;              // second shell radius (np.arange semantics)
};

__device__ __forceinline__ void cand64(const Agent &a, int c, double &x, double &y, double &z)
{
    if (c >= 2 * a.N) { x = a.vdx; y = a.vdy; z = a.vdz; return; }
    int n = (c >= a.N) ? c - a.N : c;
    double rad = (c >= a.N) ? a.shell1 : 0.5;
    x = __dmul_rn(rad, a.dir[3 * n]); y = __dmul_rn(rad, a.dir[3 * n + 1]); z = __dmul_rn(rad, a.dir[3 * n + 2]);
}

// util.py:6-20 satisfied_constraint; acos(c) <= pi/4 evaluated as c >= cos(pi/4)
__device__ __noinline__ bool posture_ok(const Agent &a, double time_step, double x, double y, double z)
{
    double next_z = __dadd_rn(a.pz, __dmul_rn(time_step, z));
    double den = __dmul_rn(a.nvA, __dsqrt_rn(dot3(x, y, z, x, y, z)));
    double ct = __ddiv_rn(dot3((double)a.vx, (double)a.vy, (double)a.vz, x, y, z), den);
    return (ct >= 0.70710678118654757) && (next_z >= 0.0);
}

// util.py:30-41 is_intersect in fp64 (pAB, R, v_dif)
__device__ __noinline__ bool is_intersect64(double ax, double ay, double az, double R, double vx, double vy, double vz)
{
    double dist = __dsqrt_rn(dot3(ax, ay, az, ax, ay, az));
    if (dist <= R) dist = R;
    double tb = asin(__ddiv_rn(R, dist));
    double nv = __dsqrt_rn(dot3(vx, vy, vz, vx, vy, vz));
    double tc = acos(__ddiv_rn(dot3(ax, ay, az, vx, vy, vz), __dmul_rn(dist, nv)));
    return !(tb <= tc);
}

// Guard-band recheck: the same predicate in its algebraic fp64 form, p.v > |v| sqrt(d^2 - R^2) with
// d = max(|p|, R); only when THAT is within 1e-12 of the boundary is the asin/acos form evaluated.
__device__ __forceinline__ bool is_intersect64_fast(double ax, double ay, double az, double R, double vx, double vy, double vz)
{
    const double pp = dot3(ax, ay, az, ax, ay, az), vv = dot3(vx, vy, vz, vx, vy, vz), dt = dot3(ax, ay, az, vx, vy, vz);
    const double c2 = fmax(pp - R * R, 0.0);
    const double df = dt * dt - vv * c2;
    if (fabs(df) <= 1e-12 * (pp * vv) || !(vv > 0.0)) return is_intersect64(ax, ay, az, R, vx, vy, vz);
    return dt > 0.0 && df > 0.0;
}



// lexicographic warp arg-min over (a, b); lanes without a value pass a = ~0ull
__device__ __forceinline__ void warp_lexmin(unsigned long long &a, unsigned &b)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long a2 = __shfl_xor_sync(FULL, a, o);
        unsigned b2 = __shfl_xor_sync(FULL, b, o);
        if (a2 < a || (a2 == a && b2 < b)) { a = a2; b = b2; }
    }
}
__device__ __forceinline__ bool lex_less(unsigned long long a, unsigned b, unsigned long long a2, unsigned b2)
{
    return a < a2 || (a == a2 && b < b2);
}
// Same arg-min, but one REDUX when every participating key fits 22 bits (suitable-candidate keys are
// distances in units of 1e-5, i.e. < 2^22 for speeds below ~20 m/s) and c < 1024.
__device__ __forceinline__ void warp_min_key_c(unsigned long long &a, unsigned &b)
{
    const bool none = (a == ~0ull);
    const bool small = none || (a < (1ull << 22) && b < 1024u);
    if (__all_sync(FULL, small)) {
        const unsigned packed = none ? 0xffffffffu : (((unsigned)a << 10) | b);
        const unsigned m = __reduce_min_sync(FULL, packed);
        if (m == 0xffffffffu) { a = ~0ull; b = 0xffffffffu; }
        else { a = (unsigned long long)(m >> 10); b = m & 1023u; }
    } else {
        warp_lexmin(a, b);
    }
}

// ------------------------------------------------------------------ k_vpref (thread per agent)
template <bool USE_DISTANCE, bool ALLOW_EXT>
__global__ void __launch_bounds__(256) k_vpref(StepArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= policy_items(A)) return;
    const int s = policy_item(A, t);
    const float4 mv = A.s_vel[s];
    if (__float_as_int(mv.w) & FL_INACTIVE) return;
    const int id = A.s_id[s];
    const double4 me = A.s_posr[s];
    double v_store[3], v_des[3];
    if (ALLOW_EXT && A.st.vpref_ext != nullptr) {
        for (int d = 0; d < 3; d++) v_des[d] = v_store[d] = A.st.vpref_ext[3 * id + d];
    } else {
        double target[3] = {A.st.goal[3 * id], A.st.goal[3 * id + 1], A.st.goal[3 * id + 2]};
        if (!ALLOW_EXT) waypoint_target(A.st, id, me.x, me.y, me.z, USE_DISTANCE, target);   // SCA / RVO3D+Dubins never follow agent.path
        v_pref_straight(&A.st.goal[3 * id], target, me.x, me.y, me.z, A.st.pref_speed[id], USE_DISTANCE, v_store, v_des);
    }
    // first-step gate: l3norm(vA,0) <= 1e-5 (scaPolicy.py:34) / distance(vA,0) <= 1e-5 (orca3dPolicy.py:52)
    const double nv32 = __dsqrt_rn(sumsq_f32(mv.x, mv.y, mv.z));
    const bool first = USE_DISTANCE ? (py_round5(__dadd_rn(nv32, 1e-5)) <= 1e-5) : (py_round5(nv32) <= 1e-5);
    // w carries pref_speed (> 0); its sign is the zero-velocity-gate flag
    const double pref_w = A.st.pref_speed[id];
    A.s_vdes[s] = make_double4(v_des[0], v_des[1], v_des[2], first ? -pref_w : pref_w);
    if (A.st.vpref_out) {
        A.st.vpref_out[3 * id] = v_store[0]; A.st.vpref_out[3 * id + 1] = v_store[1]; A.st.vpref_out[3 * id + 2] = v_store[2];
    }
}

// ------------------------------------------------------------------ k_select (warp per agent)
#include "scab_select_fast.cuh"

// USE_LIST: process only the agents the fast kernel appended to A.slow_list
template <int POLICY, bool USE_LIST>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 5) k_select(StepArgs A)
{
    constexpr bool IS_ORCA = (POLICY == SCAB_POLICY_ORCA_SAMPLED);
    constexpr bool SHUNT = (POLICY == SCAB_POLICY_SCA || POLICY == SCAB_POLICY_SRVO3D);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpSmem *smem = reinterpret_cast<WarpSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpSmem &W = smem[wid];
    const int warps_total = gridDim.x * WARPS_PER_CTA;
    unsigned long long n_recheck = 0;

    const int n_work = USE_LIST ? min(*A.slow_count2, A.n) : policy_items(A);       // (the list the second fast pass leaves)
    if (USE_LIST && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.counters[6], (unsigned long long)n_work);
    for (int wi = blockIdx.x * WARPS_PER_CTA + wid; wi < n_work; wi += warps_total) {
        const int s = USE_LIST ? A.slow_list2[wi] : policy_item(A, wi);
        const float4 mv = A.s_vel[s];
        const int flags = __float_as_int(mv.w);
        if (flags & FL_INACTIVE) continue;
        const double4 vd = A.s_vdes[s];
        if (!USE_LIST && A.row_count < 0) {   // hide the two dependent global round trips of the NEXT agents behind this agent's arithmetic:
            // header lines two agents ahead -> L2; neighbour records one agent ahead -> L2 (their index list
            // was prefetched one iteration ago)
            const int s2 = s + 2 * warps_total, s1 = s + warps_total;
            if (s2 < A.n) {
                if (lane == 0) prefetch_l2(&A.s_vel[s2]);
                if (lane == 1) prefetch_l2(&A.s_vdes[s2]);
                if (lane == 2) prefetch_l2(&A.s_posr[s2]);
                if (lane == 3) prefetch_l2(&A.s_nbr_cnt[s2]);
                if (lane == 4) prefetch_l2(&A.s_nbr[(size_t)s2 * KNBR]);
            }
            if (s1 < A.n && lane < KNBR) {
                const int e1 = A.s_nbr[(size_t)s1 * KNBR + lane];
                if (e1 >= 0 && e1 < A.n) { prefetch_l2(&A.s_posr[e1]); prefetch_l2(&A.s_vel[e1]); }
            }
        }
        if (vd.w < 0.0) {   // zero-velocity shortcut: vA_post = 0.3 v_pref (scaPolicy.py:38)
            if (lane == 0)
                A.s_vpost[s] = make_double4(__dmul_rn(0.3, vd.x), __dmul_rn(0.3, vd.y), __dmul_rn(0.3, vd.z),
                                            __longlong_as_double((long long)SCAB_TAG_FIRST));
            continue;
        }
        const double4 me = A.s_posr[s];
        Agent a;
        a.px = me.x; a.py = me.y; a.pz = me.z; a.r = me.w;
        a.vx = mv.x; a.vy = mv.y; a.vz = mv.z;
        a.pref = vd.w;
        a.nvA = (double)norm3f_np(a.vx, a.vy, a.vz);
        const bool zaxis = (POLICY == SCAB_POLICY_SCA) && (flags & SCAB_ST_ZAXIS);
        a.N = zaxis ? 128 : 256;
        a.dir = zaxis ? A.dir128 : A.dir256;
        a.dirf = zaxis ? A.dir128f : A.dir256f;
        {   // np.arange(0.5, pref + 0.03, pref - 0.5)[1]
            double stepv = __dsub_rn(a.pref, 0.5);
            double second = __dadd_rn(0.5, stepv);
            a.shell1 = __dadd_rn(0.5, __dsub_rn(second, 0.5));
        }
        a.vdx = vd.x; a.vdy = vd.y; a.vdz = vd.z;
        const int C = 2 * a.N + 1;
        int tag = 0, n_suit = 0;
        const int nc = A.s_nbr_cnt[s];
        // largest candidate speed (both shells and v_pref), slightly inflated: scales the fp32 guard bands
        const float vmaxf = 1.0001f * (float)fmax(fmax(0.5, fabs(a.shell1)),
                                                  sqrt(a.vdx * a.vdx + a.vdy * a.vdy + a.vdz * a.vdz));

        // ---------------------------------------------------------------- constraints (lane k)
        // Culling: every candidate that passes the posture test lies in the solid cone {|v| <= vmax,
        // angle(v, vA) <= pi/4}, which fits in the ball B(m, rho), m = vmax/sqrt2 * vA/|vA|, rho = vmax/sqrt2.
        // A constraint whose forbidden region stays >= 1e-2 clear of that ball cannot reject (or come near
        // the guard band of) any tested candidate, so phase B skips it.  (Phase D uses all of them.)
        float4 cP = make_float4(0.f, 0.f, 0.f, 0.f), cH = cP;
        float2 cG = make_float2(0.f, 0.f);
        bool keep = false;
        const float fnvA = (float)a.nvA;
        const float ball_r = 0.70711f * vmaxf;
        const float bmx = ball_r * a.vx / fnvA, bmy = ball_r * a.vy / fnvA, bmz = ball_r * a.vz / fnvA;
        if (lane < nc) {
            const int e = A.s_nbr[(size_t)s * KNBR + lane];
            double4 o;
            float ox = 0.f, oy = 0.f, oz = 0.f;
            bool at_goal = true, is_obst = true;
            if (e < A.n) {
                o = A.s_posr[e];
                float4 ov = A.s_vel[e];
                ox = ov.x; oy = ov.y; oz = ov.z;
                at_goal = (__float_as_int(ov.w) & SCAB_ST_AT_GOAL) != 0;
                is_obst = false;
            } else if (e < A.n + A.m) {
                o = A.o_posr[e - A.n];
            } else {
                o = A.o_posr_by_id[e - A.n - A.m];
            }
            const double R = IS_ORCA ? __dadd_rn(__dadd_rn(a.r, 0.05), __dadd_rn(o.w, 0.05))
                                     : __dadd_rn(__dadd_rn(o.w, 0.05), __dadd_rn(a.r, 0.05));
            const double rx = __dsub_rn(o.x, a.px), ry = __dsub_rn(o.y, a.py), rz = __dsub_rn(o.z, a.pz);
            W.dR[lane] = R;
            if (!IS_ORCA) {
                // scaPolicy.py:47-60: apex p0 = pA (static / at goal) else pA + 0.5 (vB + vA) in float32
                float hx = 0.f, hy = 0.f, hz = 0.f;
                if (!at_goal) {
                    hx = __fmul_rn(0.5f, __fadd_rn(ox, a.vx));
                    hy = __fmul_rn(0.5f, __fadd_rn(oy, a.vy));
                    hz = __fmul_rn(0.5f, __fadd_rn(oz, a.vz));
                }
                W.dpx[lane] = o.x; W.dpy[lane] = o.y; W.dpz[lane] = o.z;
                W.dhx[lane] = at_goal ? a.px : __dadd_rn(a.px, (double)hx);
                W.dhy[lane] = at_goal ? a.py : __dadd_rn(a.py, (double)hy);
                W.dhz[lane] = at_goal ? a.pz : __dadd_rn(a.pz, (double)hz);
                const double pp = dot3(rx, ry, rz, rx, ry, rz);
                const float fpx = (float)rx, fpy = (float)ry, fpz = (float)rz;
                const float hh = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
                cP = make_float4(fpx, fpy, fpz, fmaf(fpz, hz, fmaf(fpy, hy, fpx * hx)));
                cH = make_float4(hx, hy, hz, hh);
                // guard band: |lhs - rhs| error <= ~6e-6 * |p|^2 (|v| + |h|)^2 (DESIGN.md "Numerics")
                const float vh_max = vmaxf + sqrtf(hh);
                const float c2f = (float)fmax(pp - R * R, 0.0);
                cG = make_float2(c2f, 2e-5f * (float)pp * vh_max * vh_max);
                {   // distance from the ball centre (relative to the apex h) to the cone {w: p.w > |w| c}
                    const float d = sqrtf((float)pp), Rf = (float)R;
                    const float qx = bmx - hx, qy = bmy - hy, qz = bmz - hz;
                    const float qa = (qx * fpx + qy * fpy + qz * fpz) / d;
                    const float qr = sqrtf(fmaxf(qx * qx + qy * qy + qz * qz - qa * qa, 0.f));
                    const float sn = fminf(Rf / d, 1.f), cs = sqrtf(c2f) / d;
                    const float dist = qr * cs - qa * sn;
                    keep = !(dist > ball_r + 1e-2f) || !((float)pp > 1.02f * Rf * Rf);
                }
            } else {
                // orca3dPolicy.py:57-107 (== orca3dPolicyOfficial.py:56-106)
                const double vrx = (double)__fsub_rn(a.vx, ox), vry = (double)__fsub_rn(a.vy, oy),
                             vrz = (double)__fsub_rn(a.vz, oz);
                const double distSq = dot3(rx, ry, rz, rx, ry, rz);
                const double RSq = __dmul_rn(R, R);
                double nx, ny, nz, ux, uy, uz;
                if (distSq > RSq) {
                    const double invT = A.inv_time_horizon;
                    double wx = __dsub_rn(vrx, __dmul_rn(invT, rx)), wy = __dsub_rn(vry, __dmul_rn(invT, ry)),
                           wz = __dsub_rn(vrz, __dmul_rn(invT, rz));
                    double wl2 = dot3(wx, wy, wz, wx, wy, wz);
                    double dp = dot3(wx, wy, wz, rx, ry, rz);
                    if (dp < 0.0 && __dmul_rn(dp, dp) > __dmul_rn(RSq, wl2)) {
                        double wl = __dsqrt_rn(wl2);
                        nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                        double sc = __dsub_rn(__dmul_rn(R, invT), wl);
                        ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                    } else {
                        double difSq = __dsub_rn(distSq, RSq);
                        double dpr = dot3(rx, ry, rz, vrx, vry, vrz);
                        double cx = __dsub_rn(__dmul_rn(ry, vrz), __dmul_rn(rz, vry));
                        double cy = __dsub_rn(__dmul_rn(rz, vrx), __dmul_rn(rx, vrz));
                        double cz = __dsub_rn(__dmul_rn(rx, vry), __dmul_rn(ry, vrx));
                        double wwSq = __ddiv_rn(dot3(cx, cy, cz, cx, cy, cz), difSq);
                        double pl = __dsqrt_rn(distSq);
                        double t = __ddiv_rn(__dadd_rn(__ddiv_rn(dpr, pl), __dmul_rn(__dsqrt_rn(wwSq), __ddiv_rn(R, pl))), pl);
                        double wx = __dsub_rn(vrx, __dmul_rn(t, rx)), wy = __dsub_rn(vry, __dmul_rn(t, ry)),
                               wz = __dsub_rn(vrz, __dmul_rn(t, rz));
                        double wl = __dsqrt_rn(dot3(wx, wy, wz, wx, wy, wz));
                        nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                        double sc = __dsub_rn(__dmul_rn(R, t), wl);
                        ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                    }
                } else {
                    const double invDt = A.inv_time_step;
                    double wx = __dsub_rn(vrx, __dmul_rn(invDt, rx)), wy = __dsub_rn(vry, __dmul_rn(invDt, ry)),
                           wz = __dsub_rn(vrz, __dmul_rn(invDt, rz));
                    double wl = __dsqrt_rn(dot3(wx, wy, wz, wx, wy, wz));
                    nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                    double sc = __dsub_rn(__dmul_rn(R, invDt), wl);
                    ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                }
                const double qx = __dadd_rn((double)a.vx, __dmul_rn(0.5, ux)), qy = __dadd_rn((double)a.vy, __dmul_rn(0.5, uy)),
                             qz = __dadd_rn((double)a.vz, __dmul_rn(0.5, uz));
                W.dpx[lane] = qx; W.dpy[lane] = qy; W.dpz[lane] = qz;
                W.dhx[lane] = nx; W.dhy[lane] = ny; W.dhz[lane] = nz;
                cP = make_float4((float)nx, (float)ny, (float)nz, (float)(nx * qx + ny * qy + nz * qz));
                cG = make_float2(0.f, 1e-5f * (vmaxf + (float)sqrt(qx * qx + qy * qy + qz * qz)));
                // half-space n.v >= n.point: the smallest n.v over the ball is n.m - rho
                keep = !((cP.x * bmx + cP.y * bmy + cP.z * bmz) - ball_r - cP.w > 1e-2f);
                W.rpx[lane] = rx; W.rpy[lane] = ry; W.rpz[lane] = rz;
                // orca3dPolicy.py:387: v_dif = v - 0.5 (vA + vB) if |vB| > 1e-5 else v
                const double nvB = is_obst ? 0.0 : (double)norm3f_np(ox, oy, oz);
                const bool mov = nvB > 1e-5;
                W.ofx[lane] = mov ? (double)__fmul_rn(0.5f, __fadd_rn(a.vx, ox)) : 0.0;
                W.ofy[lane] = mov ? (double)__fmul_rn(0.5f, __fadd_rn(a.vy, oy)) : 0.0;
                W.ofz[lane] = mov ? (double)__fmul_rn(0.5f, __fadd_rn(a.vz, oz)) : 0.0;
            }
        }

        const unsigned keep_mask = __ballot_sync(FULL, keep);
        const int nce = __popc(keep_mask);
        if (keep) {
            const int slot = __popc(keep_mask & ((1u << lane) - 1));
            W.cp[slot] = cP; W.ch[slot] = cH; W.cg[slot] = cG; W.cidx[slot] = lane;
        }

        // ---------------------------------------------------------------- A: posture + compaction
        // cos(angle(vA, rad*u)) = (vA.u)/(|vA||u|) for both shells: one fp32 dot per direction.  The
        // altitude test pA.z + dt*v.z >= 0 can only fail within dt*max|v| of the ground.
        int ncand = 0;
        {
            const float fnv = (float)a.nvA;
            const float thr = 0.70710678f * fnv;
            const float band = 2e-6f * fnv;
            const double vmax = fmax(fmax(0.5, fabs(a.shell1)), fabs(a.vdz));
            const bool near_ground = !(a.pz > A.time_step * vmax + 1e-9);
            for (int base = 0; base < a.N; base += 32) {
                const int n = base + lane;   // N is a multiple of 32
                const float4 u = a.dirf[n];
                const float dt = fmaf(a.vz, u.z, fmaf(a.vy, u.y, a.vx * u.x));
                bool p0 = dt >= thr, p1 = p0;
                if (fabsf(dt - thr) <= band || near_ground) {
                    double x, y, z;
                    cand64(a, n, x, y, z);
                    p0 = posture_ok(a, A.time_step, x, y, z);
                    cand64(a, a.N + n, x, y, z);
                    p1 = posture_ok(a, A.time_step, x, y, z);
                }
                unsigned msk = __ballot_sync(FULL, p0);
                if (p0) W.list[ncand + __popc(msk & ((1u << lane) - 1))] = (unsigned short)n;
                ncand += __popc(msk);
                msk = __ballot_sync(FULL, p1);
                if (p1) W.list[ncand + __popc(msk & ((1u << lane) - 1))] = (unsigned short)(a.N + n);
                ncand += __popc(msk);
            }
            const bool pv = posture_ok(a, A.time_step, a.vdx, a.vdy, a.vdz);   // v_pref itself, last (scaPolicy.py:206)
            if (pv) {
                if (lane == 0) W.list[ncand] = (unsigned short)(2 * a.N);
                ncand++;
            }
        }
        __syncwarp();

        // ---------------------------------------------------------------- B: constraint tests
        for (int j = lane; j < ncand; j += 32) {
            const int c = W.list[j];
            float vx, vy, vz;
            if (c >= 2 * a.N) { vx = (float)a.vdx; vy = (float)a.vdy; vz = (float)a.vdz; }
            else {
                const int n = (c >= a.N) ? c - a.N : c;
                const float rad = (c >= a.N) ? (float)a.shell1 : 0.5f;
                const float4 u = a.dirf[n];
                vx = rad * u.x; vy = rad * u.y; vz = rad * u.z;
            }
            const float vv0 = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
            bool suit = true;
            for (int kk = 0; kk < nce && suit; kk++) {
                bool inside, unsure;
                const float4 P = W.cp[kk];
                const float2 G = W.cg[kk];
                if (!IS_ORCA) {
                    const float4 H = W.ch[kk];
                    const float dt = fmaf(P.z, vz, fmaf(P.y, vy, P.x * vx)) - P.w;
                    const float vh = fmaf(H.z, vz, fmaf(H.y, vy, H.x * vx));
                    const float vv = fmaf(-2.f, vh, vv0) + H.w;
                    const float df = fmaf(dt, dt, -vv * G.x);   // lhs - rhs
                    inside = (dt > 0.f) && (df > 0.f);
                    unsure = fabsf(df) <= G.y;
                } else {
                    const float val = fmaf(P.z, vz, fmaf(P.y, vy, P.x * vx)) - P.w;
                    inside = !(val >= 0.f);
                    unsure = fabsf(val) <= G.y;
                }
                if (unsure) {
                    const int k = W.cidx[kk];
                    n_recheck++;
                    double x, y, z;
                    cand64(a, c, x, y, z);
                    if (!IS_ORCA) {
                        // scaPolicy.py:180: v_dif = new_v + pA - p_0
                        double dx = __dsub_rn(__dadd_rn(x, a.px), W.dhx[k]), dy = __dsub_rn(__dadd_rn(y, a.py), W.dhy[k]),
                               dz = __dsub_rn(__dadd_rn(z, a.pz), W.dhz[k]);
                        inside = is_intersect64_fast(__dsub_rn(W.dpx[k], a.px), __dsub_rn(W.dpy[k], a.py),
                                                     __dsub_rn(W.dpz[k], a.pz), W.dR[k], dx, dy, dz);
                    } else {
                        // orca3dPolicy.py:328-333 is_inORCA
                        double val = dot3(__dsub_rn(x, W.dpx[k]), __dsub_rn(y, W.dpy[k]), __dsub_rn(z, W.dpz[k]),
                                          W.dhx[k], W.dhy[k], W.dhz[k]);
                        inside = !(val >= 0.0);
                    }
                }
                if (inside) suit = false;
            }
            W.list[j] = (unsigned short)(c | (suit ? 0x8000 : 0));
            n_suit += suit ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_suit += __shfl_xor_sync(FULL, n_suit, o);
        __syncwarp();

        // ---------------------------------------------------------------- C / D: selection
        // Sorted order of the reference's list.sort == (key, generation index) lexicographic.
        const bool fallback = (n_suit == 0);
        tag = fallback ? SCAB_TAG_FALLBACK : SCAB_TAG_SUITABLE;
        const int total = fallback ? C : ncand;
        const long long thr_units = (POLICY == SCAB_POLICY_SCA) ? (fallback ? 5000 : 3000) : 10000;
        const double thr = (POLICY == SCAB_POLICY_SCA) ? (fallback ? 5e-2 : 3e-2) : 1e-1;
        auto slot_cand = [&](int j) -> int { return fallback ? j : (int)(W.list[j] & 0x7fff); };
        auto slot_live = [&](int j) -> bool { return fallback ? true : (W.list[j] & 0x8000) != 0; };

        // keys, once per live slot
        unsigned long long ka = ~0ull;
        unsigned kb = 0xffffffffu;
        // `heavy`: run the time-to-collision loop of compute_without_suitV (posture-passing candidates only)
        auto eval_key = [&](int j, int c, bool heavy) {
            double x, y, z;
            cand64(a, c, x, y, z);
            const double dist = __dsqrt_rn(sumsq3(__dsub_rn(x, a.vdx), __dsub_rn(y, a.vdy), __dsub_rn(z, a.vdz)));
            unsigned long long key;
            if (!fallback) {
                key = (unsigned long long)py_round5_units(dist);     // l3norm(v, v_pref) in units of 1e-5
            } else {
                // compute_without_suitV (scaPolicy.py:148-165 / orca3dPolicy.py:381-397)
                double tcmin = 0.0;
                bool have = false;
                if (heavy) {
                    for (int k = 0; k < nc; k++) {
                        double ax, ay, az, dx, dy, dz;
                        if (!IS_ORCA) {
                            ax = __dsub_rn(W.dpx[k], a.px); ay = __dsub_rn(W.dpy[k], a.py); az = __dsub_rn(W.dpz[k], a.pz);
                            dx = __dsub_rn(__dadd_rn(x, a.px), W.dhx[k]); dy = __dsub_rn(__dadd_rn(y, a.py), W.dhy[k]);
                            dz = __dsub_rn(__dadd_rn(z, a.pz), W.dhz[k]);
                        } else {
                            ax = W.rpx[k]; ay = W.rpy[k]; az = W.rpz[k];
                            dx = __dsub_rn(x, W.ofx[k]); dy = __dsub_rn(y, W.ofy[k]); dz = __dsub_rn(z, W.ofz[k]);
                        }
                        // same decision as is_intersect64: the algebraic form defers to asin/acos near the boundary
                        if (is_intersect64_fast(ax, ay, az, W.dR[k], dx, dy, dz)) {
                            const double dv = dot3(dx, dy, dz, ax, ay, az);
                            const double v2 = dot3(dx, dy, dz, dx, dy, dz);
                            const double discr = __dsub_rn(__dmul_rn(dv, dv),
                                                           __dmul_rn(v2, __dsub_rn(dot3(ax, ay, az, ax, ay, az), __dmul_rn(W.dR[k], W.dR[k]))));
                            double tc = __ddiv_rn(__dsub_rn(dv, __dsqrt_rn(discr)), v2);
                            if (tc < 0) tc = 0.0;
                            if (!have || tc < tcmin) { tcmin = tc; have = true; }
                        }
                    }
                }
                const double tcv = __dadd_rn(have ? tcmin : 0.0, 1e-5);
                const double kd = __dadd_rn(__ddiv_rn(0.2, tcv), py_round5(dist));
                key = (unsigned long long)__double_as_longlong(kd);   // kd > 0: the bit pattern is monotone
            }
            W.key[j] = key;
            if (SHUNT) {   // l3norm(v, vA) in units (scaPolicy.py:128)
                const double dva = __dsqrt_rn(sumsq3(__dsub_rn(x, (double)a.vx), __dsub_rn(y, (double)a.vy), __dsub_rn(z, (double)a.vz)));
                W.dvu[j] = (unsigned int)py_round5_units(dva);
            }
            if (lex_less(key, (unsigned)c, ka, kb)) { ka = key; kb = (unsigned)c; }
        };
        if (!fallback) {
            for (int j = lane; j < total; j += 32)
                if (slot_live(j)) eval_key(j, slot_cand(j), false);
        } else {
            // Every one of the C candidates gets a key, but only those that pass the posture test (the compacted
            // list of phase A, ~15 %) walk the neighbours; doing them first, lanes together, keeps this path
            // -- one warp, a single agent's latency -- short.
            for (int j = lane; j < C; j += 32) W.key[j] = ~0ull;
            __syncwarp();
            for (int j = lane; j < ncand; j += 32) {
                const int c = (int)(W.list[j] & 0x7fff);
                eval_key(c, c, true);
            }
            __syncwarp();
            for (int j = lane; j < C; j += 32)
                if (W.key[j] == ~0ull) eval_key(j, j, false);
        }
        // 1. head of the sorted list
        warp_min_key_c(ka, kb);
        int best = (int)kb;
        const int live_total = fallback ? C : n_suit;
        __syncwarp();
        if (SHUNT && live_total > 1) {
            // scaPolicy.py:119-145 shunted_strategy
            double x0, y0, z0;
            cand64(a, best, x0, y0, z0);
            const long long d0u = (long long)py_round5_units(
                __dsqrt_rn(sumsq3(__dsub_rn(x0, (double)a.vx), __dsub_rn(y0, (double)a.vy), __dsub_rn(z0, (double)a.vz))));
            // 2. first element of the sorted list that breaks the prefix: abs(d0 - di) < thr fails
            unsigned long long fa = ~0ull;
            unsigned fb = 0xffffffffu;
            for (int j = lane; j < total; j += 32) {
                if (!slot_live(j)) continue;
                const long long diu = (long long)W.dvu[j];
                const long long du = d0u > diu ? d0u - diu : diu - d0u;
                bool within;
                if (du < thr_units - 1) within = true;
                else if (du > thr_units + 1) within = false;
                else within = fabs(__dsub_rn(div1e5((double)d0u), div1e5((double)diu))) < thr;
                if (!within) {
                    const int c = slot_cand(j);
                    if (lex_less(W.key[j], (unsigned)c, fa, fb)) { fa = W.key[j]; fb = (unsigned)c; }
                }
            }
            warp_min_key_c(fa, fb);
            // 3. min-phi / max-phi over the prefix; ties -> first in sorted order
            long long pmin = LLONG_MAX;
            unsigned long long pmin_k = ~0ull;
            unsigned pmin_c = 0xffffffffu;
            long long pmax = LLONG_MIN;
            unsigned long long pmax_k = ~0ull;
            unsigned pmax_c = 0xffffffffu;
            for (int j = lane; j < total; j += 32) {
                if (!slot_live(j)) continue;
                const int c = slot_cand(j);
                const unsigned long long k = W.key[j];
                if (!lex_less(k, (unsigned)c, fa, fb)) continue;
                double x, y, z;
                cand64(a, c, x, y, z);
                // util.py:145-150 get_phi
                double phi = (y >= 0) ? atan2(y, x) : __dadd_rn(2 * SCAB_PI, atan2(y, x));
                const long long pi_units = (long long)trunc(__dmul_rn(phi, 1e5));
                if (pi_units < pmin || (pi_units == pmin && lex_less(k, (unsigned)c, pmin_k, pmin_c))) {
                    pmin = pi_units; pmin_k = k; pmin_c = (unsigned)c;
                }
                if (pi_units > pmax || (pi_units == pmax && lex_less(k, (unsigned)c, pmax_k, pmax_c))) {
                    pmax = pi_units; pmax_k = k; pmax_c = (unsigned)c;
                }
            }
            {   // phi in units of 1e-5 is in [-314160, 628319]: bias into unsigned and use REDUX
                const unsigned umin = (pmin == LLONG_MAX) ? 0xffffffffu : (unsigned)(pmin + 1000000);
                const unsigned umax = (pmax == LLONG_MIN) ? 0u : (unsigned)(pmax + 1000000);
                const unsigned gmin = __reduce_min_sync(FULL, umin), gmax = __reduce_max_sync(FULL, umax);
                if (umin != gmin) { pmin_k = ~0ull; pmin_c = 0xffffffffu; }
                if (umax != gmax) { pmax_k = ~0ull; pmax_c = 0xffffffffu; }
                warp_min_key_c(pmin_k, pmin_c);
                warp_min_key_c(pmax_k, pmax_c);
                pmin = (long long)gmin - 1000000;
                pmax = (long long)gmax - 1000000;
            }
            const double phi_min = div1e5((double)pmin), phi_max = div1e5((double)pmax);
            best = (fabs(__dsub_rn(phi_max, phi_min)) <= SCAB_PI) ? (int)pmin_c : (int)pmax_c;
        }
        if (lane == 0) {
            double bx, by, bz;
            cand64(a, best, bx, by, bz);
            // scaPolicy.py:239 truncation; w packs tag | n_suit << 8
            A.s_vpost[s] = make_double4(trunc5(bx), trunc5(by), trunc5(bz),
                                        __longlong_as_double((long long)tag | ((long long)n_suit << 8)));
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_recheck += __shfl_xor_sync(FULL, n_recheck, o);
    if (lane == 0 && n_recheck) atomicAdd(&A.counters[3], n_recheck);
}

// ------------------------------------------------------------------ k_integrate (thread per agent)
template <bool IS_ORCA>
__global__ void __launch_bounds__(256) k_integrate(StepArgs A)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool ran = false;
    if (t < policy_items(A)) {
        const int s = policy_item(A, t);
        const int id = A.s_id[s];
        const double4 me = A.s_posr[s];
        const int flags = __float_as_int(A.s_vel[s].w);
        Action7 act;
        double npx, npy, npz;
        if (flags & SCAB_ST_GHOST) {
            // halo copy: owned (and stepped) by another GPU
        } else if (flags & FL_INACTIVE) {
            // mampenv.py:35-36: skipped agents keep the zero action row
#pragma unroll
            for (int k = 0; k < 7; k++) act.a[k] = 0.f;
            kinematic_update(A, id, act, me.x, me.y, me.z, &npx, &npy, &npz);
            if (A.st.tag) A.st.tag[id] = 0;
            if (A.st.n_suit) A.st.n_suit[id] = 0;
        } else {
            ran = true;
            const double4 vp = A.s_vpost[s];
            const long long bits = __double_as_longlong(vp.w);
            const int tag = (int)(bits & 0xff), n_suit = (int)(bits >> 8);
            act = cartesian2spherical(vp.x, vp.y, vp.z, A.st.heading[3 * id], A.st.heading[3 * id + 1], false);
            kinematic_update(A, id, act, me.x, me.y, me.z, &npx, &npy, &npz);
            // RVO family: computeNeighbors is not called on the zero-velocity shortcut (scaPolicy.py:34-45);
            // the ORCA policies call it before the gate (orca3dPolicy.py:50-52)
            const bool nbrs_ran = IS_ORCA || !(tag & SCAB_TAG_FIRST);
            if (nbrs_ran && A.s_qcoll[s]) A.st.status[id] = flags | SCAB_ST_COLLISION;   // agent.py:83-85
            if (!nbrs_ran && A.st.nbr_cnt) A.st.nbr_cnt[id] = -1;
            if (A.st.tag) A.st.tag[id] = tag;
            if (A.st.n_suit) A.st.n_suit[id] = n_suit;
        }
    }
    const unsigned msk = __ballot_sync(FULL, ran);
    if ((threadIdx.x & 31) == 0 && msk) {
        atomicAdd(&A.counters[0], (unsigned long long)__popc(msk));
        atomicAdd(&A.counters[4], (unsigned long long)__popc(msk));
    }
}

template <int POLICY>
static cudaError_t launch_select(const StepArgs &A, int sms, cudaStream_t s)
{
    constexpr bool IS_ORCA = (POLICY == SCAB_POLICY_ORCA_SAMPLED);
    const int items = A.row_count < 0 ? A.n : A.row_count;
    const int nb256 = (items + 255) / 256;
    // resident CTAs per SM of the two persistent kernels on THIS handle's device (slots 2 * POLICY, 2 * POLICY + 1)
    LaunchCfg &cfg = *A.cfg;
    int &per_sm = cfg.occ[2 * POLICY], &per_sm_fast = cfg.occ[2 * POLICY + 1], &per_sm_list = cfg.occ[14 + POLICY];
    const int no_fast = cfg.no_fast;
    const size_t smem = sizeof(WarpSmem) * WARPS_PER_CTA, smem_list = sizeof(GroupSmemT<POLICY == SCAB_POLICY_ORCA_SAMPLED>) * FG_WARPS * FG_PER_WARP,
                 smem_fast = smem_list + 384 * sizeof(float4);
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_select<POLICY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select<POLICY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select_fast<POLICY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_select_fast<POLICY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_list);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_list, k_select_fast<POLICY, true>, FG_WARPS * 32, smem_list);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_select<POLICY, true>, WARPS_PER_CTA * 32, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fast, k_select_fast<POLICY, false>, FG_WARPS * 32, smem_fast);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (per_sm_fast < 1) per_sm_fast = 1;
        if (per_sm_list < 1) per_sm_list = 1;
    }
    // persistent kernels, exactly one resident wave each: grid = SMs x resident CTAs per SM
    const int need = (items + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int grid = need < sms * per_sm ? need : sms * per_sm;
    if (no_fast) {
        k_select<POLICY, false><<<grid, WARPS_PER_CTA * 32, smem, s>>>(A);
        k_integrate<IS_ORCA><<<nb256, 256, 0, s>>>(A);
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(A.slow_count, 0, 2 * sizeof(int32_t), s);       // both work lists (slow_count2 == slow_count + 1)
    if (e != cudaSuccess) return e;
    const int need_f = (items + FG_WARPS * FG_PER_WARP - 1) / (FG_WARPS * FG_PER_WARP);
    const int grid_f = need_f < sms * per_sm_fast ? need_f : sms * per_sm_fast;
    k_select_fast<POLICY, false><<<grid_f, FG_WARPS * 32, smem_fast, s>>>(A);
    // pass 2: the agents tier 1 did not pin (~1 %), eight lanes each again but examining EVERY candidate; what that cannot
    // handle either (no suitable velocity at all, > 63 surviving directions) goes on to the general kernel, one warp per
    // agent (usually nobody: the launch is then one warp's look at an empty list)
    const int grid_l = need_f < sms * per_sm_list ? need_f : sms * per_sm_list;
    k_select_fast<POLICY, true><<<grid_l, FG_WARPS * 32, smem_list, s>>>(A);
    k_select<POLICY, true><<<grid, WARPS_PER_CTA * 32, smem, s>>>(A);
    k_integrate<IS_ORCA><<<nb256, 256, 0, s>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_policy_sampled(const StepArgs &A, int policy, cudaStream_t s)
{
    if (A.n <= 0 || A.row_count == 0) return cudaSuccess;
    const int sms = A.cfg->sms;
    const int nb = ((A.row_count < 0 ? A.n : A.row_count) + 255) / 256;
    if (A.row_count >= 0 && A.dub_states && !A.st.vpref_ext && (policy == SCAB_POLICY_SCA || policy == SCAB_POLICY_RVO3D_DUBINS))
        return cudaErrorNotSupported;       // the Dubins tracker kernels run for every agent
    const int grid = sms;   // launch_select multiplies by the resident CTAs per SM
    cudaError_t e;
    switch (policy) {
    case SCAB_POLICY_SCA:
        if (A.dub_states && !A.st.vpref_ext) { e = launch_vpref_dubins(A, s); if (e != cudaSuccess) return e; }
        else k_vpref<false, true><<<nb, 256, 0, s>>>(A);
        e = launch_select<SCAB_POLICY_SCA>(A, grid, s);
        break;
    case SCAB_POLICY_SRVO3D:
        k_vpref<false, false><<<nb, 256, 0, s>>>(A);
        e = launch_select<SCAB_POLICY_SRVO3D>(A, grid, s);
        break;
    case SCAB_POLICY_RVO3D:
        k_vpref<false, false><<<nb, 256, 0, s>>>(A);
        e = launch_select<SCAB_POLICY_RVO3D>(A, grid, s);
        break;
    case SCAB_POLICY_RVO3D_DUBINS:
        if (A.dub_states && !A.st.vpref_ext) { e = launch_vpref_dubins(A, s); if (e != cudaSuccess) return e; }
        else k_vpref<false, true><<<nb, 256, 0, s>>>(A);
        e = launch_select<SCAB_POLICY_RVO3D_DUBINS>(A, grid, s);
        break;
    case SCAB_POLICY_ORCA_SAMPLED:
        k_vpref<true, false><<<nb, 256, 0, s>>>(A);
        e = launch_select<SCAB_POLICY_ORCA_SAMPLED>(A, grid, s);
        break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}
