// scab_dubins.cu -- SCA's Dubins v_pref provider on the device (the first "next" row).
// Replaces, for SCAB_POLICY_SCA / SCAB_POLICY_RVO3D_DUBINS when no external v_pref is bound:
//   compute_v_pref      mamp/policies/sca/scaPolicy.py:264-338  (stateful tracker, one state per agent)
//   compute_dubins      :92-104, update_dubins :243-250, dubins_path_node_pop :253-261
//   dubinsmaneuver3d    mamp/policies/sca/dubinsmaneuver3d.py:34-162 (radius line search of Vana et al.)
//   2-D Dubins words    mamp/policies/sca/dubinsmaneuver2d.py:33-218, get_coordinates :260-297
//   is_parallel         mamp/util.py:125-137
// The reference materialises the sampled path as a Python list and pops points off it; here the path is its
// two 2-D manoeuvres plus a cursor, and sample i is evaluated on demand with the same arithmetic.
// Two kernels, one thread per agent, fp64: k_dubins_track follows the current path (one or two samples) and lists
// the agents that need a new plan; k_dubins_replan plans those (~100 radius trials x 12 Dubins words each).
#include "scab_common.cuh"

struct Man2D {
    double qi[3], r, t, p, q, length;
    int mode[3];   // 0 = L, 1 = S, 2 = R
    int ok;
};

struct DubState {
    int in_use, n_points, cursor, has_nbr;
    double ss, length, qi0[3];
    Man2D h, v;
    double now_goal[3], v_pref[3], nbr0_dsq;
    int replans, pad;          // diagnostics
};

size_t scab_dubins_state_bytes() { return sizeof(DubState); }

#define ML 0
#define MS 1
#define MR 2

__device__ __forceinline__ double mod2pi_d(double th) { return th - 2.0 * SCAB_PI * floor(th / 2.0 / SCAB_PI); }

// dubinsmaneuver2d.py:33-145
struct WordTrig { double sa, ca, sb, cb, c_ab; };   // the five trig values every word shares, evaluated once per 2-D plan

__device__ __forceinline__ bool plan_word(int w, double alpha, double beta, double d, const WordTrig &T, double &t, double &p, double &q, int *mode)
{
    const double sa = T.sa, ca = T.ca, sb = T.sb, cb = T.cb, c_ab = T.c_ab;
    switch (w) {
    case 0: {   // LSL
        mode[0] = ML; mode[1] = MS; mode[2] = ML;
        const double tmp0 = d + sa - sb;
        const double p2 = 2 + (d * d) - (2 * c_ab) + (2 * d * (sa - sb));
        if (p2 < 0) return false;
        const double tmp1 = atan2((cb - ca), tmp0);
        t = mod2pi_d(-alpha + tmp1); p = sqrt(p2); q = mod2pi_d(beta - tmp1);
        return true;
    }
    case 1: {   // RSR
        mode[0] = MR; mode[1] = MS; mode[2] = MR;
        const double tmp0 = d - sa + sb;
        const double p2 = 2 + (d * d) - (2 * c_ab) + (2 * d * (sb - sa));
        if (p2 < 0) return false;
        const double tmp1 = atan2((ca - cb), tmp0);
        t = mod2pi_d(alpha - tmp1); p = sqrt(p2); q = mod2pi_d(-beta + tmp1);
        return true;
    }
    case 2: {   // LSR
        mode[0] = ML; mode[1] = MS; mode[2] = MR;
        const double p2 = -2 + (d * d) + (2 * c_ab) + (2 * d * (sa + sb));
        if (p2 < 0) return false;
        p = sqrt(p2);
        const double tmp2 = atan2((-ca - cb), (d + sa + sb)) - atan2(-2.0, p);
        t = mod2pi_d(-alpha + tmp2); q = mod2pi_d(-mod2pi_d(beta) + tmp2);
        return true;
    }
    case 3: {   // RSL
        mode[0] = MR; mode[1] = MS; mode[2] = ML;
        const double p2 = (d * d) - 2 + (2 * c_ab) - (2 * d * (sa + sb));
        if (p2 < 0) return false;
        p = sqrt(p2);
        const double tmp2 = atan2((ca + cb), (d - sa - sb)) - atan2(2.0, p);
        t = mod2pi_d(alpha - tmp2); q = mod2pi_d(beta - tmp2);
        return true;
    }
    case 4: {   // RLR
        mode[0] = MR; mode[1] = ML; mode[2] = MR;
        const double tmp = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0;
        if (fabs(tmp) > 1.0) return false;
        p = mod2pi_d(2 * SCAB_PI - acos(tmp));
        t = mod2pi_d(alpha - atan2(ca - cb, d - sa + sb) + mod2pi_d(p / 2.0));
        q = mod2pi_d(alpha - beta - t + mod2pi_d(p));
        return true;
    }
    default: {  // LRL
        mode[0] = ML; mode[1] = MR; mode[2] = ML;
        const double tmp = (6. - d * d + 2 * c_ab + 2 * d * (-sa + sb)) / 8.;
        if (fabs(tmp) > 1) return false;
        p = mod2pi_d(2 * SCAB_PI - acos(tmp));
        t = mod2pi_d(-alpha - atan2(ca - cb, d + sa - sb) + p / 2.);
        q = mod2pi_d(mod2pi_d(beta) - alpha - t + mod2pi_d(p));
        return true;
    }
    }
}

// dubinsmaneuver2d.py:148-218
__device__ __noinline__ void plan2d(const double *start, const double *end, double c, Man2D &m)
{
    m.qi[0] = start[0]; m.qi[1] = start[1]; m.qi[2] = start[2];
    m.r = c;
    m.ok = 0;
    m.t = m.p = m.q = 0.0;
    m.mode[0] = m.mode[1] = m.mode[2] = MS;
    const double ex = end[0] - start[0], ey = end[1] - start[1];
    const double D = sqrt(ex * ex + ey * ey);
    const double d = D / c;
    const double theta = mod2pi_d(atan2(ey, ex));
    const double alpha = mod2pi_d(start[2] - theta), beta = mod2pi_d(end[2] - theta);
    double bcost = INFINITY;
    WordTrig T;
    sincos(alpha, &T.sa, &T.ca);
    sincos(beta, &T.sb, &T.cb);
    T.c_ab = cos(alpha - beta);
#pragma unroll
    for (int w = 0; w < 6; w++) {
        double t, p, q;
        int mode[3];
        if (!plan_word(w, alpha, beta, d, T, t, p, q, mode)) continue;
        const double cost = c * (fabs(t) + fabs(p) + fabs(q));
        if (bcost > cost) {
            m.t = t; m.p = p; m.q = q;
            m.mode[0] = mode[0]; m.mode[1] = mode[1]; m.mode[2] = mode[2];
            bcost = cost;
            m.ok = 1;
        }
    }
    m.length = bcost;
}

// dubinsmaneuver3d.py:135-162
__device__ bool try_to_construct(const double *qi, const double *qf, double Rmin, double pl_lo, double pl_hi, double hr, Man2D &h, Man2D &v)
{
    const double qi2[3] = {qi[0], qi[1], qi[3]}, qf2[3] = {qf[0], qf[1], qf[3]};
    plan2d(qi2, qf2, hr, h);
    const double qi3[3] = {0.0, qi[2], qi[4]}, qf3[3] = {h.length, qf[2], qf[4]};
    const double vc = sqrt(1.0 / (Rmin * Rmin) - 1.0 / (hr * hr));
    if (vc < 1e-5) return false;
    plan2d(qi3, qf3, 1.0 / vc, v);
    if (v.mode[0] == MR && v.mode[1] == ML && v.mode[2] == MR) return false;
    if (v.mode[0] == MR) {
        if (qi[4] - v.t < pl_lo) return false;
    } else {
        if (qi[4] + v.t > pl_hi) return false;
    }
    return true;
}

// dubinsmaneuver2d.py:283-297
__device__ __forceinline__ void seg(double offset, const double *qi, int mode, double *q)
{
    if (mode == ML) {
        q[0] = qi[0] + sin(qi[2] + offset) - sin(qi[2]);
        q[1] = qi[1] - cos(qi[2] + offset) + cos(qi[2]);
        q[2] = qi[2] + offset;
    } else if (mode == MR) {
        q[0] = qi[0] - sin(qi[2] - offset) + sin(qi[2]);
        q[1] = qi[1] + cos(qi[2] - offset) - cos(qi[2]);
        q[2] = qi[2] - offset;
    } else {
        q[0] = qi[0] + cos(qi[2]) * offset;
        q[1] = qi[1] + sin(qi[2]) * offset;
        q[2] = qi[2];
    }
}
// dubinsmaneuver2d.py:260-280
__device__ void get_coordinates(const Man2D &m, double offset, double *q)
{
    const double noffset = offset / m.r;
    const double qi[3] = {0., 0., m.qi[2]};
    double q1[3], q2[3];
    seg(m.t, qi, m.mode[0], q1);
    seg(m.p, q1, m.mode[1], q2);
    if (noffset < m.t) seg(noffset, qi, m.mode[0], q);
    else if (noffset < (m.t + m.p)) seg(noffset - m.t, q1, m.mode[1], q);
    else seg(noffset - m.t - m.p, q2, m.mode[2], q);
    q[0] = q[0] * m.r + qi[0];
    q[1] = q[1] * m.r + qi[1];
    q[2] = mod2pi_d(q[2]);
}

// maneuver3d.path[i][:3] (dubinsmaneuver3d.py:116-132 compute_sampling)
__device__ void dubins_point(const DubState &S, int i, double *out)
{
    const double offset = i * S.ss;
    double qSZ[3], qXY[3];
    get_coordinates(S.v, offset, qSZ);
    get_coordinates(S.h, qSZ[0], qXY);
    out[0] = qXY[0] + S.qi0[0];
    out[1] = qXY[1] + S.qi0[1];
    out[2] = qSZ[1] + S.qi0[2];
}

// dubinsmaneuver3d.py:34-113
__device__ __noinline__ void dubins3d(const double *qi, const double *qf, double Rmin, double pl_lo, double pl_hi, DubState &S)
{
    Man2D fbh, fbv, fch, fcv;
    double b = 1.0;
    bool fb = try_to_construct(qi, qf, Rmin, pl_lo, pl_hi, Rmin * b, fbh, fbv);
    for (int guard = 0; !fb && guard < 64; guard++) {
        b *= 2.0;
        fb = try_to_construct(qi, qf, Rmin, pl_lo, pl_hi, Rmin * b, fbh, fbv);
    }
    double step = 0.1;
    while (fabs(step) > 1e-10) {
        double c = b + step;
        if (c < 1.0) c = 1.0;
        const bool fc = try_to_construct(qi, qf, Rmin, pl_lo, pl_hi, Rmin * c, fch, fcv);
        if (fc && fcv.length < fbv.length) {
            b = c; fbh = fch; fbv = fcv;
            step *= 2.;
            continue;
        }
        step *= -0.1;
    }
    S.h = fbh; S.v = fbv;
    S.length = fbv.length;
    S.qi0[0] = qi[0]; S.qi0[1] = qi[1]; S.qi0[2] = qi[2];
    S.ss = (S.length > 100) ? S.length / 1000 : 0.1;
    S.n_points = max((int)ceil((S.length + S.ss) / S.ss), 0);   // len(np.arange(0, length + ss, ss))
    S.cursor = 0;
}

__device__ bool is_parallel_d(float vx, float vy, float vz, const double *v2)
{   // util.py:125-137 with vec1 = float32 velocity
    const float n1 = norm3f_np(vx, vy, vz);
    const double n2 = sqrt(dot3(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]));
    if (n1 <= 1e-5 || n2 <= 1e-5) return true;
    const double a0 = (double)__fdiv_rn(vx, n1), a1 = (double)__fdiv_rn(vy, n1), a2 = (double)__fdiv_rn(vz, n1);
    return np_round5(1.0 - fabs(dot3(a0, a1, a2, v2[0] / n2, v2[1] / n2, v2[2] / n2))) < 3e-3;
}

__device__ void replan(DubState &S, const double *pos, const double *heading, const double *goal, const double *goal_heading)
{   // scaPolicy.py:92-104, 275-279 (turning_radius 1.5, pitchlims +-pi/4: agent.py:24,27)
    const double qi[5] = {pos[0], pos[1], pos[2], heading[0], heading[1]};
    const double qf[5] = {goal[0], goal[1], goal[2], goal_heading[0], goal_heading[1]};
    dubins3d(qi, qf, 1.5, -SCAB_PI / 4, SCAB_PI / 4, S);
    S.replans++;
    for (int k = 0; k < 4; k++) if (S.cursor < S.n_points) S.cursor++;    // dubins_path_node_pop
    dubins_point(S, S.cursor, S.now_goal);                                 // agent.dubins_path.pop()
    S.cursor++;
}

__device__ void update_dubins(DubState &S, const double *pos, const double *goal)
{   // scaPolicy.py:243-250
    const double dis = l3norm_d(pos[0], pos[1], pos[2], S.now_goal[0], S.now_goal[1], S.now_goal[2]);
    if (dis < S.ss * 2) {
        if (S.cursor < S.n_points) { dubins_point(S, S.cursor, S.now_goal); S.cursor++; }
        else { S.now_goal[0] = goal[0]; S.now_goal[1] = goal[1]; S.now_goal[2] = goal[2]; }
    }
}

// Inputs of one agent's compute_v_pref, gathered once
struct AgentIn {
    int id, flags;
    float4 mv;
    double pos[3], goal[3], heading[3], gh[3], pref, dis_goal;
};

__device__ __forceinline__ void load_agent(const StepArgs &A, int s, AgentIn &I)
{
    I.mv = A.s_vel[s];
    I.flags = __float_as_int(I.mv.w);
    I.id = A.s_id[s];
    const int id = I.id;
    const double4 me = A.s_posr[s];
    I.pos[0] = me.x; I.pos[1] = me.y; I.pos[2] = me.z;
    for (int d = 0; d < 3; d++) {
        I.goal[d] = A.st.goal[3 * id + d];
        I.heading[d] = A.st.heading[3 * id + d];
        I.gh[d] = A.st.goal_heading ? A.st.goal_heading[3 * id + d] : 0.0;
    }
    I.pref = A.st.pref_speed[id];
    I.dis_goal = l3norm_d(I.pos[0], I.pos[1], I.pos[2], I.goal[0], I.goal[1], I.goal[2]);
}

// scaPolicy.py:329-338: normalise, the at-goal zero, the 1e-5 truncation; then remember this step's
// agent.neighbors[0][1] (read by the NEXT compute_v_pref, scaPolicy.py:299)
__device__ __forceinline__ void finish_vpref(const StepArgs &A, int s, const AgentIn &I, DubState &S, const double *dif)
{
    const int id = I.id;
    const double nrm = py_round5(__dsqrt_rn(sumsq3(dif[0], dif[1], dif[2])));
    for (int d = 0; d < 3; d++) S.v_pref[d] = __ddiv_rn(__dmul_rn(dif[d], I.pref), nrm);
    if (I.dis_goal < 0.2) S.v_pref[0] = S.v_pref[1] = S.v_pref[2] = 0.0;     // reached(goal, pos, 0.2)
    const double nv32 = __dsqrt_rn(sumsq_f32(I.mv.x, I.mv.y, I.mv.z));
    const bool first = py_round5(nv32) <= 1e-5;                              // scaPolicy.py:34
    A.s_vdes[s] = make_double4(trunc5(S.v_pref[0]), trunc5(S.v_pref[1]), trunc5(S.v_pref[2]), first ? -I.pref : I.pref);
    if (A.st.vpref_out) { A.st.vpref_out[3 * id] = S.v_pref[0]; A.st.vpref_out[3 * id + 1] = S.v_pref[1]; A.st.vpref_out[3 * id + 2] = S.v_pref[2]; }
    if (!first) {   // the zero-velocity shortcut does not query neighbours (scaPolicy.py:34-45)
        const double d0 = A.dub_d0[id];
        S.has_nbr = d0 >= 0.0;
        S.nbr0_dsq = d0 >= 0.0 ? d0 : 0.0;
    }
}

// Pass 1 (thread per agent, cell-sorted order): follow the current path (scaPolicy.py:264-327).  Agents that need
// a new plan are appended to A.slow_list -- a re-plan is ~100x the work of tracking, so it runs in its own
// kernel with full warps instead of a few divergent lanes.
__global__ void __launch_bounds__(128) k_dubins_track(const StepArgs A, DubState *states)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool need_plan = false;
    if (s < A.n && !(__float_as_int(A.s_vel[s].w) & FL_INACTIVE)) {
        AgentIn I;
        load_agent(A, s, I);
        DubState &S = states[I.id];
        if (!S.in_use) {
            need_plan = true;
        } else {
            update_dubins(S, I.pos, I.goal);
            const double dis = l3norm_d(I.pos[0], I.pos[1], I.pos[2], S.now_goal[0], S.now_goal[1], S.now_goal[2]);
            const double max_size = py_round5(6 * S.ss);
            const double gx = I.goal[0] - I.pos[0], gy = I.goal[1] - I.pos[1], gz = I.goal[2] - I.pos[2];
            const double ct = dot3((double)I.mv.x, (double)I.mv.y, (double)I.mv.z, gx, gy, gz) /
                              ((double)norm3f_np(I.mv.x, I.mv.y, I.mv.z) * sqrt(dot3(gx, gy, gz, gx, gy, gz)));
            const double cm = (1.0 < ct) ? 1.0 : ct;             // python min(ct, 1.0) keeps a NaN
            const double theta = py_round5(acos(cm));
            const double deg100 = py_round5(100.0 * (SCAB_PI / 180.0));
            const double min_dist_ob = S.has_nbr ? py_round5(sqrt(S.nbr0_dsq)) : 10.0;
            const bool condition_dist = (I.flags & SCAB_ST_ZAXIS) ? (min_dist_ob >= 2.0 * 1.5) : false;
            if (((is_parallel_d(I.mv.x, I.mv.y, I.mv.z, S.v_pref) || I.dis_goal <= 3.0 * 1.5) && dis < max_size) || (theta >= deg100) ||
                condition_dist) {
                update_dubins(S, I.pos, I.goal);
                double dif[3];
                if (S.cursor < S.n_points) for (int d = 0; d < 3; d++) dif[d] = S.now_goal[d] - I.pos[d];
                else for (int d = 0; d < 3; d++) dif[d] = I.goal[d] - I.pos[d];
                finish_vpref(A, s, I, S, dif);
            } else {
                need_plan = true;
            }
        }
    }
    // warp-aggregated append
    const unsigned msk = __ballot_sync(0xffffffffu, need_plan);
    if (msk) {
        const int lane = threadIdx.x & 31, leader = __ffs(msk) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(A.slow_count, __popc(msk));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (need_plan) A.slow_list[base + __popc(msk & ((1u << lane) - 1))] = s;
    }
}

// Pass 2 (thread per listed agent): compute_dubins + the first pops (scaPolicy.py:92-104,275-279,322-327)
__global__ void __launch_bounds__(64) k_dubins_replan(const StepArgs A, DubState *states)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *A.slow_count) return;
    const int s = A.slow_list[k];
    AgentIn I;
    load_agent(A, s, I);
    DubState &S = states[I.id];
    S.in_use = 1;
    replan(S, I.pos, I.heading, I.goal, I.gh);
    double dif[3];
    for (int d = 0; d < 3; d++) dif[d] = S.now_goal[d] - I.pos[d];
    finish_vpref(A, s, I, S, dif);
}

cudaError_t launch_vpref_dubins(const StepArgs &A, cudaStream_t s)
{
    if (A.n <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(A.slow_count, 0, sizeof(int32_t), s);
    if (e != cudaSuccess) return e;
    k_dubins_track<<<(A.n + 127) / 128, 128, 0, s>>>(A, (DubState *)A.dub_states);
    k_dubins_replan<<<(A.n + 63) / 64, 64, 0, s>>>(A, (DubState *)A.dub_states);   // surplus blocks exit on the device-side count
    return cudaGetLastError();
}
