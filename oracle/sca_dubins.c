/*
 * sca_dubins.c -- TEST INFRASTRUCTURE ONLY (part of the parity oracle, see sca_oracle.c).
 *
 * Plain-C restatement of SCA's Dubins v_pref provider:
 *   mamp/policies/sca/scaPolicy.py:92-104 (compute_dubins), :243-261 (update_dubins, dubins_path_node_pop),
 *   :264-338 (compute_v_pref);  mamp/policies/sca/dubinsmaneuver3d.py:34-162;
 *   mamp/policies/sca/dubinsmaneuver2d.py:33-218, 260-297;  mamp/util.py:125-137 (is_parallel).
 * The reference materialises the sampled path as a Python list and pops points off it; here the path is
 * kept as its two 2-D manoeuvres plus a cursor, and point i is evaluated on demand (same arithmetic).
 *
 * Pinned by: the known answer of dubinsmaneuver3d.py:230-234 (length 976.79) and the per-step V_des recorded
 * from the reference's SCA rollouts in tests/golden/*_sca.npz (tests/test_oracle_dubins.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define PI 3.141592653589793

typedef struct {
    double qi[3];      /* start (x, y, yaw) of the 2-D manoeuvre */
    double r;          /* r_min */
    double t, p, q;    /* segment lengths in units of r */
    int32_t mode[3];   /* 0 = L, 1 = S, 2 = R */
    int32_t ok;
    double length;
} Man2D;

typedef struct {
    int32_t in_use;            /* agent.is_use_dubins */
    int32_t n_points, cursor;  /* sampled points; next index path.pop() would return */
    double ss;                 /* agent.dubins_sampling_size */
    double length;
    double qi0[3];             /* maneuver3d.qi[0:3] */
    Man2D h, v;
    double now_goal[3];        /* agent.dubins_now_goal */
    double v_pref[3];          /* agent.v_pref (untruncated, previous step) */
    double nbr0_dsq;           /* agent.neighbors[0][1] of the last computeNeighbors */
    int32_t has_nbr;
    int32_t replans;           /* diagnostics */
} DubinsState;

int orc_sizeof_dubins_state(void) { return (int)sizeof(DubinsState); }

static double mod2pi(double th) { return th - 2.0 * PI * floor(th / 2.0 / PI); }   /* util.py:112-113 */
static double pyround5(double x)
{
    if (!isfinite(x)) return x;
    double p = x * 1e5, e = fma(x, 1e5, -p), r = rint(p);
    if (fabs(p - r) == 0.5 && e != 0.0) r = (e > 0.0) ? ceil(p) : floor(p);
    return r / 1e5;
}
static double nprnd5(double x) { return rint(x * 1e5) / 1e5; }
static double l3n(const double *a, const double *b)
{
    volatile double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    volatile double s = d0 * d0, q = d1 * d1;
    s = s + q; q = d2 * d2; s = s + q;
    return pyround5(sqrt(s));
}
static double dot3d(const double *a, const double *b)
{
    double d = a[0] * b[0];
    d = fma(a[1], b[1], d);
    return fma(a[2], b[2], d);
}

enum { ML = 0, MS = 1, MR = 2 };

/* dubinsmaneuver2d.py:33-145.  Returns 0 when the word does not exist. */
static int plan_word(int w, double alpha, double beta, double d, double *t, double *p, double *q, int32_t *mode)
{
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), c_ab = cos(alpha - beta);
    switch (w) {
    case 0: {   /* LSL */
        mode[0] = ML; mode[1] = MS; mode[2] = ML;
        double tmp0 = d + sa - sb;
        double p2 = 2 + (d * d) - (2 * c_ab) + (2 * d * (sa - sb));
        if (p2 < 0) return 0;
        double tmp1 = atan2((cb - ca), tmp0);
        *t = mod2pi(-alpha + tmp1); *p = sqrt(p2); *q = mod2pi(beta - tmp1);
        return 1;
    }
    case 1: {   /* RSR */
        mode[0] = MR; mode[1] = MS; mode[2] = MR;
        double tmp0 = d - sa + sb;
        double p2 = 2 + (d * d) - (2 * c_ab) + (2 * d * (sb - sa));
        if (p2 < 0) return 0;
        double tmp1 = atan2((ca - cb), tmp0);
        *t = mod2pi(alpha - tmp1); *p = sqrt(p2); *q = mod2pi(-beta + tmp1);
        return 1;
    }
    case 2: {   /* LSR */
        mode[0] = ML; mode[1] = MS; mode[2] = MR;
        double p2 = -2 + (d * d) + (2 * c_ab) + (2 * d * (sa + sb));
        if (p2 < 0) return 0;
        *p = sqrt(p2);
        double tmp2 = atan2((-ca - cb), (d + sa + sb)) - atan2(-2.0, *p);
        *t = mod2pi(-alpha + tmp2); *q = mod2pi(-mod2pi(beta) + tmp2);
        return 1;
    }
    case 3: {   /* RSL */
        mode[0] = MR; mode[1] = MS; mode[2] = ML;
        double p2 = (d * d) - 2 + (2 * c_ab) - (2 * d * (sa + sb));
        if (p2 < 0) return 0;
        *p = sqrt(p2);
        double tmp2 = atan2((ca + cb), (d - sa - sb)) - atan2(2.0, *p);
        *t = mod2pi(alpha - tmp2); *q = mod2pi(beta - tmp2);
        return 1;
    }
    case 4: {   /* RLR */
        mode[0] = MR; mode[1] = ML; mode[2] = MR;
        double tmp = (6.0 - d * d + 2.0 * c_ab + 2.0 * d * (sa - sb)) / 8.0;
        if (fabs(tmp) > 1.0) return 0;
        *p = mod2pi(2 * PI - acos(tmp));
        *t = mod2pi(alpha - atan2(ca - cb, d - sa + sb) + mod2pi(*p / 2.0));
        *q = mod2pi(alpha - beta - *t + mod2pi(*p));
        return 1;
    }
    default: {  /* LRL */
        mode[0] = ML; mode[1] = MR; mode[2] = ML;
        double tmp = (6. - d * d + 2 * c_ab + 2 * d * (-sa + sb)) / 8.;
        if (fabs(tmp) > 1) return 0;
        *p = mod2pi(2 * PI - acos(tmp));
        *t = mod2pi(-alpha - atan2(ca - cb, d + sa - sb) + *p / 2.);
        *q = mod2pi(mod2pi(beta) - alpha - *t + mod2pi(*p));
        return 1;
    }
    }
}

/* dubinsmaneuver2d.py:148-218 (the sampled course of generate_course is never used by the 3-D planner) */
static Man2D plan2d(const double *start, const double *end, double c)
{
    Man2D m;
    memset(&m, 0, sizeof m);
    m.qi[0] = start[0]; m.qi[1] = start[1]; m.qi[2] = start[2];
    m.r = c;
    double ex = end[0] - start[0], ey = end[1] - start[1];
    volatile double two = 2.0;
    double D = sqrt(pow(ex, two) + pow(ey, two));
    double d = D / c;
    double theta = mod2pi(atan2(ey, ex));
    double alpha = mod2pi(start[2] - theta), beta = mod2pi(end[2] - theta);
    double bcost = INFINITY;
    for (int w = 0; w < 6; w++) {
        double t, p, q;
        int32_t mode[3];
        if (!plan_word(w, alpha, beta, d, &t, &p, &q, mode)) continue;
        double cost = c * (fabs(t) + fabs(p) + fabs(q));
        if (bcost > cost) {
            m.t = t; m.p = p; m.q = q;
            m.mode[0] = mode[0]; m.mode[1] = mode[1]; m.mode[2] = mode[2];
            bcost = cost;
            m.ok = 1;
        }
    }
    m.length = bcost;
    return m;
}

/* dubinsmaneuver3d.py:135-162 */
static int try_to_construct(const double *qi, const double *qf, double Rmin, const double *pitchlims, double hr, Man2D *h, Man2D *v)
{
    double qi2[3] = { qi[0], qi[1], qi[3] }, qf2[3] = { qf[0], qf[1], qf[3] };
    *h = plan2d(qi2, qf2, hr);
    double qi3[3] = { 0.0, qi[2], qi[4] }, qf3[3] = { h->length, qf[2], qf[4] };
    volatile double two = 2.0;
    double vc = sqrt(1.0 / pow(Rmin, two) - 1.0 / pow(hr, two));
    if (vc < 1e-5) return 0;
    double vr = 1.0 / vc;
    *v = plan2d(qi3, qf3, vr);
    if (v->mode[0] == MR && v->mode[1] == ML && v->mode[2] == MR) return 0;   /* "RLR" (the reference tests it twice) */
    if (v->mode[0] == MR) {
        if (qi[4] - v->t < pitchlims[0]) return 0;
    } else {
        if (qi[4] + v->t > pitchlims[1]) return 0;
    }
    return 1;
}

/* dubinsmaneuver2d.py:283-297 */
static void seg(double offset, const double *qi, int mode, double *q)
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
/* dubinsmaneuver2d.py:260-280 */
static void get_coordinates(const Man2D *m, double offset, double *q)
{
    double noffset = offset / m->r;
    double qi[3] = { 0., 0., m->qi[2] };
    double l1 = m->t, l2 = m->p, q1[3], q2[3];
    seg(l1, qi, m->mode[0], q1);
    seg(l2, q1, m->mode[1], q2);
    if (noffset < l1) seg(noffset, qi, m->mode[0], q);
    else if (noffset < (l1 + l2)) seg(noffset - l1, q1, m->mode[1], q);
    else seg(noffset - l1 - l2, q2, m->mode[2], q);
    q[0] = q[0] * m->r + qi[0];
    q[1] = q[1] * m->r + qi[1];
    q[2] = mod2pi(q[2]);
}

/* dubinsmaneuver3d.py:34-113 + compute_sampling :116-132 (the sample count only) */
void orc_dubins3d(const double *qi, const double *qf, double Rmin, const double *pitchlims, DubinsState *S)
{
    Man2D fah, fav, fbh, fbv, fch, fcv;
    double b = 1.0;
    int fa = try_to_construct(qi, qf, Rmin, pitchlims, Rmin * 1.0, &fah, &fav);
    int fb = try_to_construct(qi, qf, Rmin, pitchlims, Rmin * b, &fbh, &fbv);
    (void)fa;
    int guard = 0;
    while (!fb && guard++ < 64) {
        b *= 2.0;
        fb = try_to_construct(qi, qf, Rmin, pitchlims, Rmin * b, &fbh, &fbv);
    }
    double step = 0.1;
    while (fabs(step) > 1e-10) {
        double c = b + step;
        if (c < 1.0) c = 1.0;
        int fc = try_to_construct(qi, qf, Rmin, pitchlims, Rmin * c, &fch, &fcv);
        if (fc) {
            if (fcv.length < fbv.length) {
                b = c; fbh = fch; fbv = fcv;
                step *= 2.;
                continue;
            }
        }
        step *= -0.1;
    }
    S->h = fbh; S->v = fbv;
    S->length = fbv.length;
    S->qi0[0] = qi[0]; S->qi0[1] = qi[1]; S->qi0[2] = qi[2];
    double ss = 0.1;
    if (S->length > 100) ss = S->length / 1000;
    S->ss = ss;
    /* len(np.arange(0, length + ss, ss)) */
    volatile double stop = S->length + ss;
    S->n_points = (int32_t)ceil(stop / ss);
    if (S->n_points < 0) S->n_points = 0;
    S->cursor = 0;
}

/* maneuver3d.path[i][:3] (compute_sampling) */
void orc_dubins_point(const DubinsState *S, int i, double *out)
{
    double offset = i * S->ss;     /* np.arange: start + i*step with start = 0 */
    double qSZ[3], qXY[3];
    get_coordinates(&S->v, offset, qSZ);
    get_coordinates(&S->h, qSZ[0], qXY);
    out[0] = qXY[0] + S->qi0[0];
    out[1] = qXY[1] + S->qi0[1];
    out[2] = qSZ[1] + S->qi0[2];
}

/* mamp/util.py:125-137 with vec1 = float32 velocity, vec2 = float64 */
static int is_parallel(const float *v1, const double *v2)
{
    double dsum = 0.0;
    for (int i = 0; i < 3; i++) { volatile float pr = v1[i] * v1[i]; dsum += (double)pr; }
    float n1 = sqrtf((float)dsum);
    double n2 = sqrt(dot3d(v2, v2));
    if (n1 <= 1e-5 || n2 <= 1e-5) return 1;
    double a[3], bb[3];
    for (int i = 0; i < 3; i++) { volatile float q = v1[i] / n1; a[i] = (double)q; bb[i] = v2[i] / n2; }
    return nprnd5(1.0 - fabs(dot3d(a, bb))) < 3e-3;
}

static void replan(DubinsState *S, const double *pos, const double *heading, const double *goal, const double *goal_heading,
                   double turning_radius)
{   /* scaPolicy.py:92-104, 275-279 */
    double qi[6] = { pos[0], pos[1], pos[2], heading[0], heading[1], heading[2] };
    double qf[6] = { goal[0], goal[1], goal[2], goal_heading[0], goal_heading[1], goal_heading[2] };
    double pl[2] = { -PI / 4, PI / 4 };
    orc_dubins3d(qi, qf, turning_radius, pl, S);
    S->replans++;
    for (int k = 0; k < 4; k++) if (S->cursor < S->n_points) S->cursor++;          /* dubins_path_node_pop */
    orc_dubins_point(S, S->cursor, S->now_goal);                                     /* agent.dubins_path.pop() */
    S->cursor++;
}

static void update_dubins(DubinsState *S, const double *pos, const double *goal)
{   /* scaPolicy.py:243-250 */
    double dis = l3n(pos, S->now_goal);
    if (dis < S->ss * 2) {
        if (S->cursor < S->n_points) { orc_dubins_point(S, S->cursor, S->now_goal); S->cursor++; }
        else { S->now_goal[0] = goal[0]; S->now_goal[1] = goal[1]; S->now_goal[2] = goal[2]; }
    }
}

/* scaPolicy.py:264-338 compute_v_pref for one agent.  Returns V_des in v_des[3]. */
void orc_dubins_vpref(DubinsState *S, const double *pos, const float *vel, const double *heading, const double *goal,
                      const double *goal_heading, double pref_speed, int is_zaxis, double *v_des)
{
    const double turning_radius = 1.5, neighbor_dist = 10.0;   /* agent.py:24,33 */
    double dis_goal = l3n(pos, goal);
    double k = 3.0 * turning_radius;
    double dif[3];
    if (!S->in_use) {
        S->in_use = 1;
        replan(S, pos, heading, goal, goal_heading, turning_radius);
        for (int d = 0; d < 3; d++) dif[d] = S->now_goal[d] - pos[d];
    } else {
        update_dubins(S, pos, goal);
        double dis = l3n(pos, S->now_goal);
        double max_size = pyround5(6 * S->ss);
        double pApG[3] = { goal[0] - pos[0], goal[1] - pos[1], goal[2] - pos[2] };
        double vA[3] = { vel[0], vel[1], vel[2] };
        double dsum = 0.0;
        for (int i = 0; i < 3; i++) { volatile float pr = vel[i] * vel[i]; dsum += (double)pr; }
        float nvA = sqrtf((float)dsum);
        double ct = dot3d(vA, pApG) / ((double)nvA * sqrt(dot3d(pApG, pApG)));
        double cm = (1.0 < ct) ? 1.0 : ct;        /* python min(ct, 1.0): returns ct when it is NaN */
        double theta = nprnd5(acos(cm));          /* round(python float, 5); NaN stays NaN */
        theta = isnan(theta) ? theta : pyround5(acos(cm));
        double deg100 = pyround5(100.0 * (PI / 180.0));
        double min_dist_ob = S->has_nbr ? pyround5(sqrt(S->nbr0_dsq)) : (double)lrint(neighbor_dist);
        int condition_dist = is_zaxis ? (min_dist_ob >= 2.0 * turning_radius) : 0;
        if (((is_parallel(vel, S->v_pref) || dis_goal <= k) && dis < max_size) || (theta >= deg100) || condition_dist) {
            update_dubins(S, pos, goal);
            if (S->cursor < S->n_points) for (int d = 0; d < 3; d++) dif[d] = S->now_goal[d] - pos[d];
            else for (int d = 0; d < 3; d++) dif[d] = goal[d] - pos[d];
        } else {
            replan(S, pos, heading, goal, goal_heading, turning_radius);
            for (int d = 0; d < 3; d++) dif[d] = S->now_goal[d] - pos[d];
        }
    }
    double zero[3] = { 0, 0, 0 };
    double nrm = l3n(dif, zero);
    for (int d = 0; d < 3; d++) { volatile double q = dif[d] * pref_speed; S->v_pref[d] = q / nrm; }
    if (l3n(goal, pos) < 0.2) S->v_pref[0] = S->v_pref[1] = S->v_pref[2] = 0.0;   /* reached(goal, pos, 0.2) */
    for (int d = 0; d < 3; d++) v_des[d] = (trunc(S->v_pref[d] * 1e5) + 0.0) / 1e5;
}
