/*
 * sca_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A plain-C, double-precision CPU restatement of the per-timestep collision-avoidance step of
 * wuuya1/SCA (reference: /root/reference/mamp, pure Python/NumPy).  Nothing under sca_b200/ (the
 * product) may include, link, import or execute this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / CPU baseline.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against fixtures under
 * tests/golden/ that were produced by running the unmodified Python reference in the build
 * container (tests/golden/make_golden.py + ref_harness.py).
 *
 * Every function cites the reference file:line it restates.  Arithmetic follows the reference's
 * NumPy/Python evaluation order and its decimal-rounding conventions:
 *   - round(np.float64, 5)  -> rint(x*1e5)/1e5             (NumPy scalar __round__)
 *   - round(python float,5) -> correctly rounded decimal  (CPython float.__round__)
 *   - int(x*1e5)/1e5        -> trunc
 *   - np.dot on 3-vectors   -> x0*y0, then two FMAs (OpenBLAS ddot tail loop, measured in the
 *                              build container); float32 dot: f32 products summed in double.
 *   - velocities are float32 after the first env update (mampenv.py:31,104); t=0 velocities are 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define K_MAX 16              /* agent.py:32 maxNeighbors */
#define NEIGHBOR_DIST 10.0    /* agent.py:33 */
#define TIME_STEP 0.1         /* config.py:2 DT */
#define TIME_HORIZON 10.0     /* agent.py:35 */
#define MAX_SPEED 1.0         /* agent.py:36 */
#define NEAR_GOAL 0.5         /* config.py:3 */
#define RVO_EPS 1e-5          /* config.py:4 */
#define MAX_LEAF 10           /* kdTree.py:53 */
#define PI 3.141592653589793

enum { POL_SCA = 0, POL_SRVO3D = 1, POL_RVO3D = 2, POL_RVO3D_DUBINS = 3, POL_ORCA_SAMPLED = 4, POL_ORCA_LP = 5 };
enum { NBR_KNN = 0, NBR_REF_COMPAT = 1 };
enum { TAG_FIRST = 1, TAG_SUITABLE = 2, TAG_FALLBACK = 4, TAG_LP3_OK = 8, TAG_LP4 = 16 };
enum { ST_AT_GOAL = 1, ST_COLLISION = 2, ST_TIMEOUT = 4, ST_RUN_DONE = 8 };

typedef struct {
    int32_t begin, end, left, right;
    double mn[3], mx[3];
} KdNode;

typedef struct {
    int32_t n, m;
    double *pos;         /* [n,3] */
    float *vel;          /* [n,3] */
    double *heading;     /* [n,3] */
    double *goal;        /* [n,3] */
    double *radius;      /* [n] */
    double *pref_speed;  /* [n] */
    int32_t *status;     /* [n] */
    double *total_dist;  /* [n] */
    int32_t *step_num;   /* [n] */
    double *max_run_dist;/* [n] */
    uint8_t *is_zaxis;   /* [n] scaPolicy.py:188-189 */
    double *obs;         /* [m,4] x y z r */
    double *vpref_ext;   /* [n,3] or NULL: externally supplied V_des (Dubins provider upstream) */
    int32_t *agent_perm; /* [n] persistent KDTree.agentIDs (kdTree.py:43) */
    int32_t *obst_perm;  /* [m] */
    /* outputs */
    int32_t *ran, *nbr_cnt, *nbr_ids;
    double *nbr_dsq;
    int32_t *coll_query, *tag, *n_suit, *plane_fail, *n_planes, *lp4_proj, *lp4_inner_fail;
    double *vpref_in, *vpref_store, *action64;
    float *action;
    /* scratch owned by the caller */
    KdNode *agent_nodes; /* [2n-1] */
    KdNode *obst_nodes;  /* [2m-1] */
    /* waypoint following (agent.path / policy.now_goal, e.g. rvo3dPolicy.py:71-86); NULL: no paths */
    double *waypoints;        /* [n, waypoint_stride, 3] in the reference's list order: pop() takes the LAST live entry */
    int32_t *waypoint_left;   /* [n] len(agent.path)                                                    in/out */
    double *now_goal;         /* [n,3] policy.now_goal, NaN = None                                      in/out */
    int32_t waypoint_stride;
} OrcWorld;

/* ------------------------------------------------------------------ rounding conventions */
static double np_round5(double x) { return rint(x * 1e5) / 1e5; }

/* CPython float.__round__(x, 5): round-half-even on the exact binary value. */
static double py_round5(double x)
{
    if (!isfinite(x)) return x;
    double p = x * 1e5;
    double e = fma(x, 1e5, -p);
    double r = rint(p);
    if (fabs(p - r) == 0.5 && e != 0.0) r = (e > 0.0) ? ceil(p) : floor(p);
    return r / 1e5;
}
double orc_py_round5(double x) { return py_round5(x); }
double orc_py_round5_slow(double x)
{
    char buf[400];
    snprintf(buf, sizeof buf, "%.5f", x);
    return strtod(buf, NULL);
}

/* int(x*1e5)/1e5: Python ints have no negative zero, so -0.0 becomes +0.0 */
static double trunc5(double x) { return (trunc(x * 1e5) + 0.0) / 1e5; }

/* ------------------------------------------------------------------ numpy-like vector helpers */
static double dot3(const double *a, const double *b)
{
    double d = a[0] * b[0];
    d = fma(a[1], b[1], d);
    d = fma(a[2], b[2], d);
    return d;
}
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
static float dot3f(const float *a, const float *b)
{
    double d = 0.0;
    for (int i = 0; i < 3; i++) {
        volatile float p = a[i] * b[i];
        d += (double)p;
    }
    return (float)d;
}
static float norm3f(const float *a) { return sqrtf(dot3f(a, a)); }
static void cross3(const double *a, const double *b, double *c)
{ /* np.cross: products and difference rounded separately (ufuncs) */
    volatile double t0 = a[1] * b[2], t1 = a[2] * b[1];
    volatile double t2 = a[2] * b[0], t3 = a[0] * b[2];
    volatile double t4 = a[0] * b[1], t5 = a[1] * b[0];
    c[0] = t0 - t1;
    c[1] = t2 - t3;
    c[2] = t4 - t5;
}
static double sqr(double x) { return x * x; }

/* util.py:103-105 l3norm on float64 data: Python round of math.sqrt */
static double l3norm(const double *a, const double *b)
{
    volatile double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    volatile double s = d0 * d0;
    volatile double q = d1 * d1;
    s = s + q;
    q = d2 * d2;
    s = s + q;
    return py_round5(sqrt(s));
}
/* util.py:99-100 l3normsq: NumPy-scalar round */
static double l3normsq(const double *a, const double *b)
{
    volatile double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    volatile double s = d0 * d0;
    volatile double q = d1 * d1;
    s = s + q;
    q = d2 * d2;
    s = s + q;
    return np_round5(s);
}
/* l3norm(vA, [0,0,0]) / distance(vA,[0,0,0]) with vA float32: the squares and sums stay float32 */
static double sumsq_f32(const float *v)
{
    volatile float s = v[0] * v[0];
    volatile float q = v[1] * v[1];
    s = s + q;
    q = v[2] * v[2];
    s = s + q;
    return (double)s;
}
/* util.py:140-142 distance(): +1e-5 before rounding */
static double distance_fn(const double *a, const double *b)
{
    volatile double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    volatile double s = d0 * d0;
    volatile double q = d1 * d1;
    s = s + q;
    q = d2 * d2;
    s = s + q;
    return py_round5(sqrt(s) + 1e-5);
}
/* util.py:108-109 pi_2_pi with NumPy's floored remainder */
static double pi_2_pi(double angle)
{
    double a = angle + PI, b = 2 * PI;
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0) != (m < 0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m - PI;
}
/* util.py:145-150 get_phi */
static double get_phi(const double *v)
{
    double phi = (v[1] >= 0) ? atan2(v[1], v[0]) : 2 * PI + atan2(v[1], v[0]);
    return (trunc(phi * 1e5) + 0.0) / 1e5;
}

/* ------------------------------------------------------------------ kd-tree (kdTree.py:56-122) */
static void kd_build(const double *pos, int32_t *ids, KdNode *nodes, int begin, int end, int node)
{
    KdNode *nd = &nodes[node];
    nd->begin = begin;
    nd->end = end;
    nd->left = nd->right = 0;
    for (int c = 0; c < 3; c++) nd->mn[c] = nd->mx[c] = pos[3 * ids[begin] + c];
    for (int i = begin + 1; i < end; i++)
        for (int c = 0; c < 3; c++) {
            double v = pos[3 * ids[i] + c];
            if (v > nd->mx[c]) nd->mx[c] = v;
            if (v < nd->mn[c]) nd->mn[c] = v;
        }
    if (end - begin > MAX_LEAF) {
        double d0 = nd->mx[0] - nd->mn[0], d1 = nd->mx[1] - nd->mn[1], d2 = nd->mx[2] - nd->mn[2];
        int coord = (d0 > d1 && d0 > d2) ? 0 : (d1 > d2 ? 1 : 2);
        double split = 0.5 * (nd->mx[coord] + nd->mn[coord]);
        int left = begin, right = end;
        while (left < right) {
            while (left < right && pos[3 * ids[left] + coord] < split) left++;
            while (right > left && pos[3 * ids[right - 1] + coord] >= split) right--;
            if (left < right) {
                int32_t t = ids[left];
                ids[left] = ids[right - 1];
                ids[right - 1] = t;
                left++;
                right--;
            }
        }
        int leftSize = left - begin;
        if (leftSize == 0) {
            leftSize++;
            left++;
            right++;
        }
        nd->left = node + 1;
        nd->right = node + 2 * leftSize;
        kd_build(pos, ids, nodes, begin, left, nd->left);
        kd_build(pos, ids, nodes, left, end, nd->right);
    }
}

void orc_kdtree_build_agents(OrcWorld *w)
{
    if (w->n > 0) kd_build(w->pos, w->agent_perm, w->agent_nodes, 0, w->n, 0);
}
void orc_kdtree_build_obstacles(OrcWorld *w)
{ /* obstacle positions are the first 3 of each [x y z r] row: build on a packed copy */
    if (w->m <= 0) return;
    double *p = (double *)malloc(sizeof(double) * 3 * w->m);
    for (int i = 0; i < w->m; i++)
        for (int c = 0; c < 3; c++) p[3 * i + c] = w->obs[4 * i + c];
    kd_build(p, w->obst_perm, w->obst_nodes, 0, w->m, 0);
    free(p);
}

/* ------------------------------------------------------------------ neighbour list (agent.py:79-124) */
typedef struct {
    int cnt;
    int32_t id[K_MAX + 1];   /* agent id, or n + obstacle id */
    double dsq[K_MAX + 1];
    int is_collision;
} NbrList;

static void nbr_append_sorted(NbrList *L, int32_t id, double dsq)
{ /* pop() when full, append, stable sort by distSq: newcomer goes after equal keys */
    if (L->cnt == K_MAX) L->cnt--;
    int k = L->cnt;
    while (k > 0 && L->dsq[k - 1] > dsq) {
        L->dsq[k] = L->dsq[k - 1];
        L->id[k] = L->id[k - 1];
        k--;
    }
    L->dsq[k] = dsq;
    L->id[k] = id;
    L->cnt++;
}

/* agent.py:79-99.  rangeSq is passed BY VALUE in the reference, so it never shrinks. */
static void insert_agent_neighbor(const OrcWorld *w, int self, int other, double rangeSq, NbrList *L)
{
    if (self == other) return;
    double distSq = l3normsq(&w->pos[3 * self], &w->pos[3 * other]);
    if (distSq < sqr(w->radius[self] + w->radius[other]) && distSq < rangeSq) {
        if (!L->is_collision) {
            L->is_collision = 1;
            L->cnt = 0;
        }
        nbr_append_sorted(L, other, distSq);
    } else if (!L->is_collision && distSq < rangeSq) {
        nbr_append_sorted(L, other, distSq);
    }
}
/* agent.py:101-124 */
static void insert_obstacle_neighbor(const OrcWorld *w, int self, int ob, double rangeSq, NbrList *L)
{
    const double *po = &w->obs[4 * ob];
    double distSq1 = l3normsq(&w->pos[3 * self], po);
    double t = l3norm(&w->pos[3 * self], po) - po[3];
    volatile double two = 2.0;     /* keep gcc from folding pow(t, 2) into t*t */
    double distSq = pow(t, two);   /* python float ** 2 -> libm pow (not always == t*t to the last ulp) */
    if (distSq1 < sqr(w->radius[self] + po[3]) && distSq < rangeSq) {
        if (!L->is_collision) {
            L->is_collision = 1;
            L->cnt = 0;
        }
        nbr_append_sorted(L, w->n + ob, distSq);
    } else if (!L->is_collision && distSq < rangeSq) {
        nbr_append_sorted(L, w->n + ob, distSq);
    }
}

static double box_dist_sq(const KdNode *nd, const double *p)
{ /* kdTree.py:132-138: six sqr(max(0, .)) terms summed left to right */
    double s = 0.0;
    for (int c = 0; c < 3; c++) {
        s += sqr(fmax(0.0, nd->mn[c] - p[c]));
        s += sqr(fmax(0.0, p[c] - nd->mx[c]));
    }
    return s;
}

/* kdTree.py:127-156 */
static void query_agent_tree(const OrcWorld *w, int self, double rangeSq, int node, NbrList *L)
{
    const KdNode *nd = &w->agent_nodes[node];
    if (nd->end - nd->begin <= MAX_LEAF) {
        for (int i = nd->begin; i < nd->end; i++) insert_agent_neighbor(w, self, w->agent_perm[i], rangeSq, L);
    } else {
        double dl = box_dist_sq(&w->agent_nodes[nd->left], &w->pos[3 * self]);
        double dr = box_dist_sq(&w->agent_nodes[nd->right], &w->pos[3 * self]);
        if (dl < dr) {
            if (dl < rangeSq) {
                query_agent_tree(w, self, rangeSq, nd->left, L);
                if (dr < rangeSq) query_agent_tree(w, self, rangeSq, nd->right, L);
            }
        } else {
            if (dr < rangeSq) {
                query_agent_tree(w, self, rangeSq, nd->right, L);
                if (dl < rangeSq) query_agent_tree(w, self, rangeSq, nd->left, L);
            }
        }
    }
}
/* kdTree.py:232-262 */
static void query_obstacle_tree(const OrcWorld *w, int self, double rangeSq, int node, NbrList *L)
{
    if (w->m <= 0) return;
    const KdNode *nd = &w->obst_nodes[node];
    if (nd->end - nd->begin <= MAX_LEAF) {
        for (int i = nd->begin; i < nd->end; i++) insert_obstacle_neighbor(w, self, w->obst_perm[i], rangeSq, L);
    } else {
        double dl = box_dist_sq(&w->obst_nodes[nd->left], &w->pos[3 * self]);
        double dr = box_dist_sq(&w->obst_nodes[nd->right], &w->pos[3 * self]);
        if (dl < dr) {
            if (dl < rangeSq) {
                query_obstacle_tree(w, self, rangeSq, nd->left, L);
                if (dr < rangeSq) query_obstacle_tree(w, self, rangeSq, nd->right, L);
            }
        } else {
            if (dr < rangeSq) {
                query_obstacle_tree(w, self, rangeSq, nd->right, L);
                if (dl < rangeSq) query_obstacle_tree(w, self, rangeSq, nd->left, L);
            }
        }
    }
}

/* True-KNN variant (documented deviation): what the reference would compute if
 * rangeSq shrank (RVO2-3D semantics): the 16 smallest by (distSq, obstacles-before-agents, id), with the
 * same collision-mode rule.  Candidates come from a uniform grid (cell = neighborDist) rebuilt by
 * orc_policy_step; the result does not depend on the visiting order. */
typedef struct {
    double o[3], inv;
    int nx, ny, nz;
    int32_t *start, *items;   /* [ncell+1], [n] */
} KnnGrid;
static KnnGrid g_grid = { {0, 0, 0}, 0, 0, 0, 0, NULL, NULL };

static int grid_cell(const KnnGrid *g, const double *p, int *cx, int *cy, int *cz)
{
    *cx = (int)floor((p[0] - g->o[0]) * g->inv);
    *cy = (int)floor((p[1] - g->o[1]) * g->inv);
    *cz = (int)floor((p[2] - g->o[2]) * g->inv);
    if (*cx < 0) *cx = 0; if (*cx >= g->nx) *cx = g->nx - 1;
    if (*cy < 0) *cy = 0; if (*cy >= g->ny) *cy = g->ny - 1;
    if (*cz < 0) *cz = 0; if (*cz >= g->nz) *cz = g->nz - 1;
    return (*cz * g->ny + *cy) * g->nx + *cx;
}
static void grid_build(const OrcWorld *w)
{
    KnnGrid *g = &g_grid;
    free(g->start); free(g->items);
    double mn[3] = { 1e300, 1e300, 1e300 }, mx[3] = { -1e300, -1e300, -1e300 };
    for (int i = 0; i < w->n; i++)
        for (int c = 0; c < 3; c++) {
            if (w->pos[3 * i + c] < mn[c]) mn[c] = w->pos[3 * i + c];
            if (w->pos[3 * i + c] > mx[c]) mx[c] = w->pos[3 * i + c];
        }
    double cs = NEIGHBOR_DIST;
    for (;;) {
        g->nx = (int)floor((mx[0] - mn[0]) / cs) + 1; g->ny = (int)floor((mx[1] - mn[1]) / cs) + 1; g->nz = (int)floor((mx[2] - mn[2]) / cs) + 1;
        if ((double)g->nx * g->ny * g->nz <= 4.0e6 + 8.0 * w->n) break;
        cs *= 1.5;
    }
    g->inv = 1.0 / cs;
    for (int c = 0; c < 3; c++) g->o[c] = mn[c];
    int ncell = g->nx * g->ny * g->nz;
    g->start = (int32_t *)calloc((size_t)ncell + 1, sizeof(int32_t));
    g->items = (int32_t *)malloc(sizeof(int32_t) * (size_t)(w->n > 0 ? w->n : 1));
    int cx, cy, cz;
    for (int i = 0; i < w->n; i++) g->start[grid_cell(g, &w->pos[3 * i], &cx, &cy, &cz) + 1]++;
    for (int c = 0; c < ncell; c++) g->start[c + 1] += g->start[c];
    int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncell > 0 ? ncell : 1));
    memcpy(fill, g->start, sizeof(int32_t) * (size_t)ncell);
    for (int i = 0; i < w->n; i++) g->items[fill[grid_cell(g, &w->pos[3 * i], &cx, &cy, &cz)]++] = i;
    free(fill);
}

static void knn_insert(NbrList *L, int32_t uid, double dsq)
{   /* uid: obstacles n + o, agents id -- ordered (dsq, obstacle-first, id) */
#define UKEY(x, n_) ((x) >= (n_) ? (long long)(x) - (n_) : (long long)(x) + (1LL << 40))
    extern int g_knn_n;
    long long ku = UKEY(uid, g_knn_n);
    if (L->cnt == K_MAX) {
        double dl = L->dsq[K_MAX - 1];
        long long kl = UKEY(L->id[K_MAX - 1], g_knn_n);
        if (!(dsq < dl || (dsq == dl && ku < kl))) return;
        L->cnt--;
    }
    int k = L->cnt;
    while (k > 0 && (L->dsq[k - 1] > dsq || (L->dsq[k - 1] == dsq && UKEY(L->id[k - 1], g_knn_n) > ku))) {
        L->dsq[k] = L->dsq[k - 1];
        L->id[k] = L->id[k - 1];
        k--;
    }
    L->dsq[k] = dsq;
    L->id[k] = uid;
    L->cnt++;
#undef UKEY
}
int g_knn_n = 0;

static void knn_neighbors(const OrcWorld *w, int self, NbrList *L)
{
    const double rangeSq = NEIGHBOR_DIST * NEIGHBOR_DIST;
    const KnnGrid *g = &g_grid;
    int cx, cy, cz;
    grid_cell(g, &w->pos[3 * self], &cx, &cy, &cz);
    for (int pass = 0; pass < 2; pass++) {
        /* pass 0: is there any collider?  pass 1: collect */
        int coll = L->is_collision;
        if (pass == 1) L->cnt = 0;
        for (int o = 0; o < w->m; o++) {
            const double *po = &w->obs[4 * o];
            double d1 = l3normsq(&w->pos[3 * self], po);
            double t = l3norm(&w->pos[3 * self], po) - po[3];
            double d = t * t;
            int is_c = d1 < sqr(w->radius[self] + po[3]) && d < rangeSq;
            if (pass == 0) { if (is_c) L->is_collision = 1; }
            else if (coll ? is_c : (d < rangeSq)) knn_insert(L, w->n + o, d);
        }
        for (int z = (cz > 0 ? cz - 1 : 0); z <= cz + 1 && z < g->nz; z++)
            for (int y = (cy > 0 ? cy - 1 : 0); y <= cy + 1 && y < g->ny; y++)
                for (int x = (cx > 0 ? cx - 1 : 0); x <= cx + 1 && x < g->nx; x++) {
                    int c = (z * g->ny + y) * g->nx + x;
                    for (int q = g->start[c]; q < g->start[c + 1]; q++) {
                        int j = g->items[q];
                        if (j == self) continue;
                        double d = l3normsq(&w->pos[3 * self], &w->pos[3 * j]);
                        int is_c = d < sqr(w->radius[self] + w->radius[j]) && d < rangeSq;
                        if (pass == 0) { if (is_c) L->is_collision = 1; }
                        else if (coll ? is_c : (d < rangeSq)) knn_insert(L, j, d);
                    }
                }
    }
}

/* scaPolicy.py:107-116 computeNeighbors (identical in all six policies) */
static void compute_neighbors(const OrcWorld *w, int self, int mode, NbrList *L)
{
    L->cnt = 0;
    L->is_collision = 0;
    if (mode == NBR_REF_COMPAT) {
        double rangeSq = NEIGHBOR_DIST * NEIGHBOR_DIST;
        query_obstacle_tree(w, self, rangeSq, 0, L);
        query_agent_tree(w, self, rangeSq, 0, L);
    } else {
        knn_neighbors(w, self, L);
    }
}

/* ------------------------------------------------------------------ neighbour object view */
typedef struct {
    double p[3];
    float v[3];
    double r;
    int at_goal;   /* obj.is_at_goal (obstacle.py:26: always True for obstacles) */
    int is_obst;
} Obj;

static void get_obj(const OrcWorld *w, int32_t id, Obj *o)
{
    if (id >= w->n) {
        const double *po = &w->obs[4 * (id - w->n)];
        o->p[0] = po[0]; o->p[1] = po[1]; o->p[2] = po[2];
        o->v[0] = o->v[1] = o->v[2] = 0.f;
        o->r = po[3];
        o->at_goal = 1;
        o->is_obst = 1;
    } else {
        for (int c = 0; c < 3; c++) { o->p[c] = w->pos[3 * id + c]; o->v[c] = w->vel[3 * id + c]; }
        o->r = w->radius[id];
        o->at_goal = (w->status[id] & ST_AT_GOAL) != 0;
        o->is_obst = 0;
    }
}

/* ------------------------------------------------------------------ util.py:30-41 is_intersect */
static int is_intersect_rel(const double *pAB, double R, const double *v_dif)
{
    double dist = norm3(pAB);
    if (dist <= R) dist = R;
    double tb = asin(R / dist);
    double tc = acos(dot3(pAB, v_dif) / (dist * norm3(v_dif)));
    return !(tb <= tc);
}

/* util.py:6-20 satisfied_constraint */
static int satisfied_constraint(const OrcWorld *w, int a, const double *vc)
{
    double vA[3] = { w->vel[3 * a], w->vel[3 * a + 1], w->vel[3 * a + 2] };
    volatile double step = TIME_STEP * vc[2];
    double next_z = w->pos[3 * a + 2] + step;
    double costheta = dot3(vA, vc) / ((double)norm3f(&w->vel[3 * a]) * norm3(vc));
    if (costheta > 1.0) costheta = 1.0;
    else if (costheta < -1.0) costheta = -1.0;
    double theta = acos(costheta);
    return (theta <= PI / 4 && next_z >= 0.0);
}

/* ------------------------------------------------------------------ candidates (scaPolicy.py:190-211) */
#define MAX_CAND 1025
static int gen_candidates(int num_N, double pref_speed, double (*cand)[3])
{
    const double param_phi = (sqrt(5.0) - 1.0) / 2.0;
    double start = 0.5, stepv = pref_speed - 0.5, stop = pref_speed + 0.03;
    int len = (int)ceil((stop - start) / stepv);           /* np.arange length */
    if (len < 0) len = 0;
    if (len > 2) len = 2;                                   /* unsupported: pref_speed <= 0.53 */
    volatile double second = start + stepv;
    double delta = second - start;
    int c = 0;
    for (int i = 0; i < len; i++) {
        double rad = (i == 0) ? start : start + i * delta;
        for (int n = 1; n <= num_N; n++) {
            double z = (double)(2 * n - 1) / num_N - 1;
            double sq = sqrt(1 - z * z);
            double x = sq * cos(2 * PI * n * param_phi);
            double y = sq * sin(2 * PI * n * param_phi);
            cand[c][0] = rad * x; cand[c][1] = rad * y; cand[c][2] = rad * z;
            c++;
        }
    }
    return c;
}
/* exported for tests and for cross-checking the product's constant table */
int orc_gen_candidates(int num_N, double pref_speed, double *out)
{
    return gen_candidates(num_N, pref_speed, (double(*)[3])out);
}

/* ------------------------------------------------------------------ cones / planes */
typedef struct {
    double p0[3], pA[3], pB[3], R;   /* RVO_BA = [transl_vB_vA, pA, pB, R] (scaPolicy.py:59) */
} Cone;
typedef struct {
    double point[3], normal[3];      /* orca3dPolicyOfficial.py:21-24 */
    double R, relpos[3];
    float vB[3];
    int vB_is_f32;
} Plane;

/* stable sort of idx[0..n) by key (python list.sort(key=...)) -- merge sort */
static void stable_sort(int *idx, const double *key, int n, int *tmp)
{
    if (n < 2) return;
    int h = n / 2;
    stable_sort(idx, key, h, tmp);
    stable_sort(idx + h, key, n - h, tmp);
    int i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (key[idx[j]] < key[idx[i]]) ? idx[j++] : idx[i++];
    while (i < h) tmp[k++] = idx[i++];
    while (j < n) tmp[k++] = idx[j++];
    memcpy(idx, tmp, sizeof(int) * n);
}

/* scaPolicy.py:119-145 / srvo3dPolicy.py:100-126 */
static int shunted_strategy(const OrcWorld *w, int a, double (*cand)[3], const int *order, int n, double thr)
{
    double vA[3] = { w->vel[3 * a], w->vel[3 * a + 1], w->vel[3 * a + 2] };
    int cnt = 1;
    double d0 = l3norm(cand[order[0]], vA);
    int i = 1;
    for (;;) {
        if (fabs(d0 - l3norm(cand[order[i]], vA)) < thr) {
            cnt++;
            i++;
            if (i == n) break;
        } else break;
    }
    int imin = 0, imax = 0;
    double pmin = get_phi(cand[order[0]]), pmax = pmin;
    for (int k = 1; k < cnt; k++) {
        double ph = get_phi(cand[order[k]]);
        if (ph < pmin) { pmin = ph; imin = k; }
        if (ph > pmax) { pmax = ph; imax = k; }
    }
    if (fabs(pmax - pmin) <= PI) {
        /* min([vmin, vmax], key=phi) -> first minimal */
        return (pmax < pmin) ? order[imax] : order[imin];
    } else {
        /* max([vmin, vmax], key=phi) -> first maximal */
        return (pmax > pmin) ? order[imax] : order[imin];
    }
}

typedef struct { int tag, n_suit; } SelInfo;

/* intersect(): scaPolicy.py:187-240, rvo3dPolicy.py:141-179, srvo3dPolicy.py:185-231,
 * rvo3dDubinsPolicy.py:156-194, orca3dPolicy.py:400-439 */
static void sampled_select(const OrcWorld *w, int a, int policy, const double *v_pref,
                           const Cone *cones, const Plane *planes, int nc, double *vA_post, SelInfo *info)
{
    static __thread double cand[MAX_CAND][3];
    static __thread double key[MAX_CAND];
    static __thread int suit_idx[MAX_CAND], unsuit_idx[MAX_CAND], tmp[MAX_CAND];
    static __thread unsigned char posture[MAX_CAND];
    int num_N = (policy == POL_SCA && w->is_zaxis[a]) ? 128 : 256;
    int C = gen_candidates(num_N, w->pref_speed[a], cand);
    cand[C][0] = v_pref[0]; cand[C][1] = v_pref[1]; cand[C][2] = v_pref[2];
    C++;
    int ns = 0, nu = 0;
    const double *pA = &w->pos[3 * a];
    for (int c = 0; c < C; c++) {
        /* compute_newV_is_suit (scaPolicy.py:168-184 / orca3dPolicy.py:366-379) */
        posture[c] = (unsigned char)satisfied_constraint(w, a, cand[c]);
        int suit = 1;
        if (nc == 0) {
            if (!posture[c]) suit = 0;
        }
        for (int k = 0; k < nc && suit; k++) {
            if (policy == POL_ORCA_SAMPLED) {
                double rel[3] = { cand[c][0] - planes[k].point[0], cand[c][1] - planes[k].point[1],
                                  cand[c][2] - planes[k].point[2] };
                int in_orca = dot3(rel, planes[k].normal) >= 0.0;
                if (!in_orca || !posture[c]) suit = 0;
            } else {
                double v_dif[3], pAB[3];
                for (int d = 0; d < 3; d++) {
                    volatile double t = cand[c][d] + cones[k].pA[d];
                    v_dif[d] = t - cones[k].p0[d];
                    pAB[d] = cones[k].pB[d] - cones[k].pA[d];
                }
                if (is_intersect_rel(pAB, cones[k].R, v_dif) || !posture[c]) suit = 0;
            }
        }
        if (suit) suit_idx[ns++] = c; else unsuit_idx[nu++] = c;
    }
    info->n_suit = ns;
    int shunt = (policy == POL_SCA || policy == POL_SRVO3D);
    int best;
    if (ns > 0) {
        info->tag |= TAG_SUITABLE;
        for (int k = 0; k < ns; k++) key[suit_idx[k]] = l3norm(cand[suit_idx[k]], v_pref);
        stable_sort(suit_idx, key, ns, tmp);
        if (shunt && ns > 1)
            best = shunted_strategy(w, a, cand, suit_idx, ns, policy == POL_SCA ? 3e-2 : 1e-1);
        else
            best = suit_idx[0];
    } else {
        info->tag |= TAG_FALLBACK;
        for (int u = 0; u < nu; u++) {
            int c = unsuit_idx[u];
            /* compute_without_suitV (scaPolicy.py:148-165 / orca3dPolicy.py:381-397) */
            double tcmin = 0.0;
            int have = 0;
            for (int k = 0; k < nc; k++) {
                double v_dif[3], pAB[3], R;
                if (policy == POL_ORCA_SAMPLED) {
                    const Plane *pl = &planes[k];
                    R = pl->R;
                    double vBd[3] = { pl->vB[0], pl->vB[1], pl->vB[2] };
                    for (int d = 0; d < 3; d++) pAB[d] = pl->relpos[d];
                    double nvB = pl->vB_is_f32 ? (double)norm3f(pl->vB) : norm3(vBd);
                    if (nvB > 1e-5) {
                        for (int d = 0; d < 3; d++) {
                            volatile float s = w->vel[3 * a + d] + pl->vB[d];
                            volatile float hs = 0.5f * s;
                            v_dif[d] = cand[c][d] - (double)hs;
                        }
                    } else {
                        for (int d = 0; d < 3; d++) v_dif[d] = cand[c][d];
                    }
                } else {
                    R = cones[k].R;
                    for (int d = 0; d < 3; d++) {
                        volatile double t = cand[c][d] + cones[k].pA[d];
                        v_dif[d] = t - cones[k].p0[d];
                        pAB[d] = cones[k].pB[d] - cones[k].pA[d];
                    }
                }
                if (is_intersect_rel(pAB, R, v_dif) && posture[c]) {
                    double dv = dot3(v_dif, pAB);
                    double discr = sqr(dv) - dot3(v_dif, v_dif) * (dot3(pAB, pAB) - sqr(R));
                    double tc = (dv - sqrt(discr)) / dot3(v_dif, v_dif);
                    if (tc < 0) tc = 0.0;
                    if (!have || tc < tcmin) { tcmin = tc; have = 1; }
                }
            }
            if (!have) tcmin = 0.0;
            double tcv = tcmin + 1e-5;
            key[c] = (0.2 / tcv) + l3norm(cand[c], v_pref);
        }
        if (shunt) {
            stable_sort(unsuit_idx, key, nu, tmp);
            if (nu > 1)
                best = shunted_strategy(w, a, cand, unsuit_idx, nu, policy == POL_SCA ? 5e-2 : 1e-1);
            else
                best = unsuit_idx[0];
        } else {
            best = unsuit_idx[0];
            for (int u = 1; u < nu; u++)
                if (key[unsuit_idx[u]] < key[best]) best = unsuit_idx[u];
        }
    }
    (void)pA;
    for (int d = 0; d < 3; d++) vA_post[d] = trunc5(cand[best][d]);
}

/* ------------------------------------------------------------------ ORCA planes
 * orca3dPolicyOfficial.py:56-106 (identical orca3dPolicy.py:57-107) */
static void build_plane(const OrcWorld *w, int a, const Obj *o, Plane *pl)
{
    double relPos[3], relVel[3];
    const float *vA = &w->vel[3 * a];
    for (int d = 0; d < 3; d++) {
        relPos[d] = o->p[d] - w->pos[3 * a + d];
        volatile float rv = vA[d] - o->v[d];     /* float32 - float32 (obstacle: exact either way) */
        relVel[d] = (double)rv;
    }
    double distSq = dot3(relPos, relPos);
    volatile double agent_rad = w->radius[a] + 0.05;
    volatile double obj_rad = o->r + 0.05;
    double R = agent_rad + obj_rad;
    double RSq = R * R;
    double u[3];
    if (distSq > RSq) {
        const double invT = 1.0 / TIME_HORIZON;
        double wv[3];
        for (int d = 0; d < 3; d++) { volatile double t = invT * relPos[d]; wv[d] = relVel[d] - t; }
        double wLenSq = dot3(wv, wv);
        double dp = dot3(wv, relPos);
        if (dp < 0.0 && sqr(dp) > RSq * wLenSq) {
            double wLen = sqrt(wLenSq);
            double s = (R * invT - wLen);
            for (int d = 0; d < 3; d++) { pl->normal[d] = wv[d] / wLen; u[d] = s * pl->normal[d]; }
        } else {
            double difSq = distSq - RSq;
            double dpr = dot3(relPos, relVel);
            double cr[3];
            cross3(relPos, relVel, cr);
            double wwSq = dot3(cr, cr) / difSq;
            double pLen = norm3(relPos);
            double pAp1 = dpr / pLen;
            double p1ot = sqrt(wwSq) * (R / pLen);
            double pAot = pAp1 + p1ot;
            double t = pAot / pLen;
            double ww[3];
            for (int d = 0; d < 3; d++) { volatile double q = t * relPos[d]; ww[d] = relVel[d] - q; }
            double wwLen = norm3(ww);
            double s = (R * t - wwLen);
            for (int d = 0; d < 3; d++) { pl->normal[d] = ww[d] / wwLen; u[d] = s * pl->normal[d]; }
        }
    } else {
        const double invDt = 1.0 / TIME_STEP;
        double wv[3];
        for (int d = 0; d < 3; d++) { volatile double t = invDt * relPos[d]; wv[d] = relVel[d] - t; }
        double wLen = norm3(wv);
        double s = (R * invDt - wLen);
        for (int d = 0; d < 3; d++) { pl->normal[d] = wv[d] / wLen; u[d] = s * pl->normal[d]; }
    }
    for (int d = 0; d < 3; d++) {
        volatile double hu = 0.5 * u[d];
        pl->point[d] = (double)vA[d] + hu;
        pl->relpos[d] = relPos[d];
        pl->vB[d] = o->v[d];
    }
    pl->R = R;
    pl->vB_is_f32 = !o->is_obst;
}

/* ------------------------------------------------------------------ LP1..LP4
 * orca3dPolicyOfficial.py:126-300 */
typedef struct { double v[3]; int lp4_proj, lp4_inner_fail; } LPState;

static int lp1(LPState *S, const Plane *planes, int planeNo, const double *lpoint, const double *ldir,
               double maxSpeed, const double *vel_pref, int dir_opt)
{
    double dp = dot3(lpoint, ldir);
    double disc = sqr(dp) + sqr(maxSpeed) - dot3(lpoint, lpoint);
    if (disc < 0.0) return 0;
    double sq = sqrt(disc);
    double tLeft = -dp - sq, tRight = -dp + sq;
    for (int i = 0; i < planeNo; i++) {
        double diff[3] = { planes[i].point[0] - lpoint[0], planes[i].point[1] - lpoint[1], planes[i].point[2] - lpoint[2] };
        double num = dot3(diff, planes[i].normal);
        double den = dot3(ldir, planes[i].normal);
        if (sqr(den) <= RVO_EPS) {
            if (num > 0.0) return 0;
            continue;
        }
        double t = num / den;
        if (den >= 0.0) tLeft = fmax(tLeft, t);
        else tRight = fmin(tRight, t);
        if (tLeft > tRight) return 0;
    }
    double tt;
    if (dir_opt) {
        tt = (dot3(vel_pref, ldir) > 0.0) ? tRight : tLeft;
    } else {
        double diff[3] = { vel_pref[0] - lpoint[0], vel_pref[1] - lpoint[1], vel_pref[2] - lpoint[2] };
        double t = dot3(ldir, diff);
        tt = (t < tLeft) ? tLeft : (t > tRight ? tRight : t);
    }
    for (int d = 0; d < 3; d++) { volatile double q = tt * ldir[d]; S->v[d] = lpoint[d] + q; }
    return 1;
}

static int lp2(LPState *S, const Plane *planes, int planeNo, double maxSpeed, const double *vel_pref, int dir_opt)
{
    const Plane *P = &planes[planeNo];
    double planeDist = dot3(P->point, P->normal);
    double planeDistSq = sqr(planeDist);
    double radiusSq = sqr(maxSpeed);
    if (planeDistSq > radiusSq) return 0;
    double planeRadiusSq = radiusSq - planeDistSq;
    double center[3];
    for (int d = 0; d < 3; d++) center[d] = planeDist * P->normal[d];
    if (dir_opt) {
        double dn = dot3(vel_pref, P->normal);
        double pov[3];
        for (int d = 0; d < 3; d++) { volatile double q = dn * P->normal[d]; pov[d] = vel_pref[d] - q; }
        double povSq = dot3(pov, pov);
        if (povSq <= RVO_EPS) {
            for (int d = 0; d < 3; d++) S->v[d] = center[d];
        } else {
            double s = sqrt(planeRadiusSq / povSq);
            for (int d = 0; d < 3; d++) { volatile double q = s * pov[d]; S->v[d] = center[d] + q; }
        }
    } else {
        double diff[3] = { P->point[0] - vel_pref[0], P->point[1] - vel_pref[1], P->point[2] - vel_pref[2] };
        double dpn = dot3(diff, P->normal);
        for (int d = 0; d < 3; d++) { volatile double q = dpn * P->normal[d]; S->v[d] = vel_pref[d] + q; }
        if (dot3(S->v, S->v) > radiusSq) {
            double pr[3] = { S->v[0] - center[0], S->v[1] - center[1], S->v[2] - center[2] };
            double prSq = dot3(pr, pr);
            double s = sqrt(planeRadiusSq / prSq);
            for (int d = 0; d < 3; d++) { volatile double q = s * pr[d]; S->v[d] = center[d] + q; }
        }
    }
    for (int i = 0; i < planeNo; i++) {
        double diff[3] = { planes[i].point[0] - S->v[0], planes[i].point[1] - S->v[1], planes[i].point[2] - S->v[2] };
        if (dot3(planes[i].normal, diff) > 0.0) {
            double cp[3];
            cross3(planes[i].normal, P->normal, cp);
            if (dot3(cp, cp) <= RVO_EPS) return 0;
            double ldir[3], lnormal[3], lpoint[3];
            double ncp = norm3(cp);
            for (int d = 0; d < 3; d++) ldir[d] = cp[d] / ncp;
            cross3(ldir, P->normal, lnormal);
            double d12[3] = { planes[i].point[0] - P->point[0], planes[i].point[1] - P->point[1], planes[i].point[2] - P->point[2] };
            double dp1 = dot3(d12, planes[i].normal);
            double dp2 = dot3(lnormal, planes[i].normal);
            double f = dp1 / dp2;
            for (int d = 0; d < 3; d++) { volatile double q = f * lnormal[d]; lpoint[d] = P->point[d] + q; }
            if (!lp1(S, planes, i, lpoint, ldir, maxSpeed, vel_pref, dir_opt)) return 0;
        }
    }
    return 1;
}

static int lp3(LPState *S, const Plane *planes, int np, double maxSpeed, const double *vel_pref, int dir_opt)
{
    if (dir_opt) {
        for (int d = 0; d < 3; d++) S->v[d] = vel_pref[d] * maxSpeed;
    } else if (dot3(vel_pref, vel_pref) > sqr(maxSpeed)) {
        double nv = norm3(vel_pref);
        for (int d = 0; d < 3; d++) { volatile double q = vel_pref[d] / nv; S->v[d] = q * maxSpeed; }
    } else {
        for (int d = 0; d < 3; d++) S->v[d] = vel_pref[d];
    }
    for (int i = 0; i < np; i++) {
        double diff[3] = { planes[i].point[0] - S->v[0], planes[i].point[1] - S->v[1], planes[i].point[2] - S->v[2] };
        if (dot3(planes[i].normal, diff) > 0.0) {
            double temp[3] = { S->v[0], S->v[1], S->v[2] };
            if (!lp2(S, planes, i, maxSpeed, vel_pref, dir_opt)) {
                S->v[0] = temp[0]; S->v[1] = temp[1]; S->v[2] = temp[2];
                return i;
            }
        }
    }
    return np;
}

static void lp4(LPState *S, const Plane *planes, int np, int beginPlane, double radius)
{
    Plane proj[K_MAX];
    for (int i = beginPlane; i < np; i++) {
        /* orca3dPolicyOfficial.py:264: np.dot(normal, (point - v) > 0.0) -- dot with a BOOLEAN vector:
         * the sum of the normal components whose (point - v) component is positive; truthy iff != 0 */
        double bsum = 0.0;
        {
            double b[3];
            for (int d = 0; d < 3; d++) b[d] = (planes[i].point[d] - S->v[d] > 0.0) ? 1.0 : 0.0;
            bsum = dot3(planes[i].normal, b);
        }
        if (bsum != 0.0) {
            int npj = 0;
            for (int j = 0; j < i; j++) {
                Plane *pp = &proj[npj];
                double cp[3];
                cross3(planes[j].normal, planes[i].normal, cp);
                if (dot3(cp, cp) <= RVO_EPS) {
                    if (dot3(planes[i].normal, planes[j].normal) > 0.0) continue;
                    for (int d = 0; d < 3; d++) { volatile double s = planes[i].point[d] + planes[j].point[d]; pp->point[d] = 0.5 * s; }
                } else {
                    double ln[3];
                    cross3(cp, planes[i].normal, ln);
                    double dj[3] = { planes[j].point[0] - planes[i].point[0], planes[j].point[1] - planes[i].point[1], planes[j].point[2] - planes[i].point[2] };
                    double dp1 = dot3(dj, planes[j].normal);
                    double dp2 = dot3(ln, planes[j].normal);
                    double f = dp1 / dp2;
                    for (int d = 0; d < 3; d++) { volatile double q = f * ln[d]; pp->point[d] = planes[i].point[d] + q; }
                }
                double dn[3] = { planes[j].normal[0] - planes[i].normal[0], planes[j].normal[1] - planes[i].normal[1], planes[j].normal[2] - planes[i].normal[2] };
                double nn = norm3(dn);
                for (int d = 0; d < 3; d++) pp->normal[d] = dn[d] / nn;
                npj++;
            }
            double temp[3] = { S->v[0], S->v[1], S->v[2] };
            S->lp4_proj++;
            if (lp3(S, proj, npj, radius, planes[i].normal, 1) < npj) {
                S->lp4_inner_fail++;
                S->v[0] = temp[0]; S->v[1] = temp[1]; S->v[2] = temp[2];
            }
        }
    }
}

/* ------------------------------------------------------------------ v_pref, action */
/* rvo3dPolicy.py:182-196 (use_distance=0) / orca3dPolicy*.py:345-359 (use_distance=1) */
/* get_trajectory: rvo3dPolicy.py:71-86, srvo3dPolicy.py:71-86 (l3norm) / orca3dPolicy.py:298-312,
 * orca3dPolicyOfficial.py:302-316 (distance).  Writes the target compute_v_pref aims at. */
static void get_trajectory(OrcWorld *w, int a, int use_distance, double *target)
{
    const double *goal = &w->goal[3 * a], *pos = &w->pos[3 * a];
    if (w->waypoints == NULL || w->waypoint_left[a] <= 0) {
        for (int d = 0; d < 3; d++) target[d] = goal[d];
        if (w->now_goal) for (int d = 0; d < 3; d++) w->now_goal[3 * a + d] = goal[d];
        return;
    }
    double *ng = &w->now_goal[3 * a];
    const double *wp = &w->waypoints[(size_t)a * w->waypoint_stride * 3];
#define POP_WAYPOINT() do { int k_ = --w->waypoint_left[a]; ng[0] = wp[3 * k_]; ng[1] = wp[3 * k_ + 1]; ng[2] = wp[3 * k_ + 2]; } while (0)
    if (isnan(ng[0])) POP_WAYPOINT();                     /* first */
    double dis = use_distance ? distance_fn(pos, ng) : l3norm(pos, ng);
    double d_ng_goal = use_distance ? distance_fn(ng, goal) : l3norm(ng, goal);
    double d_pos_goal = use_distance ? distance_fn(pos, goal) : l3norm(pos, goal);
    if (dis <= 1.0 * w->radius[a]) {                      /* update_now_goal_dist = 1.0 */
        if (w->waypoint_left[a] > 0) POP_WAYPOINT();
    } else if (d_ng_goal >= d_pos_goal) {
        if (w->waypoint_left[a] > 0) POP_WAYPOINT();
    }
#undef POP_WAYPOINT
    for (int d = 0; d < 3; d++) target[d] = ng[d];
}

static void compute_v_pref_straight(OrcWorld *w, int a, int use_distance, int follow_path, double *v_store, double *v_des)
{
    double dif[3], zero[3] = { 0, 0, 0 }, target[3];
    if (follow_path) get_trajectory(w, a, use_distance, target);
    else for (int d = 0; d < 3; d++) target[d] = w->goal[3 * a + d];
    for (int d = 0; d < 3; d++) dif[d] = target[d] - w->pos[3 * a + d];
    double nrm = use_distance ? distance_fn(dif, zero) : l3norm(dif, zero);
    nrm = (trunc(nrm * 1e5) + 0.0) / 1e5;
    for (int d = 0; d < 3; d++) { volatile double q = dif[d] * w->pref_speed[a]; v_store[d] = q / nrm; }
    /* reached(goal, pos, 0.2): util.py:23-27 */
    if (l3norm(&w->goal[3 * a], &w->pos[3 * a]) < 0.2) v_store[0] = v_store[1] = v_store[2] = 0.0;
    for (int d = 0; d < 3; d++) v_des[d] = trunc5(v_store[d]);
}

/* util.py:44-55 (use_distance=0) / orca3dPolicyOfficial.py:331-342 (use_distance=1) */
static void cartesian2spherical(const OrcWorld *w, int a, const double *v, int use_distance, double *act)
{
    double zero[3] = { 0, 0, 0 };
    double speed = use_distance ? distance_fn(v, zero) : l3norm(v, zero);
    double alpha = 0, beta = 0;
    if (!(speed < 0.001)) {
        alpha = atan2(v[1], v[0]) - w->heading[3 * a];
        volatile double x2 = v[0] * v[0];   /* pow(x,2) */
        volatile double y2 = v[1] * v[1];
        beta = atan2(v[2], sqrt(x2 + y2)) - w->heading[3 * a + 1];
    }
    act[0] = v[0]; act[1] = v[1]; act[2] = v[2]; act[3] = speed; act[4] = alpha; act[5] = beta; act[6] = 0.0;
}

/* ------------------------------------------------------------------ one policy call
 * find_next_action: scaPolicy.py:26-73, rvo3dPolicy.py:23-69, srvo3dPolicy.py:23-69,
 * rvo3dDubinsPolicy.py:26-71, orca3dPolicy.py:38-120, orca3dPolicyOfficial.py:37-124 */
static void policy_one(OrcWorld *w, int a, int policy, int nbr_mode)
{
    int is_orca = (policy == POL_ORCA_SAMPLED || policy == POL_ORCA_LP);
    int ext_vpref = (policy == POL_SCA || policy == POL_RVO3D_DUBINS) && w->vpref_ext != NULL;
    double v_store[3], v_des[3];
    if (ext_vpref) {
        for (int d = 0; d < 3; d++) v_des[d] = v_store[d] = w->vpref_ext[3 * a + d];
    } else {
        /* SCA / RVO3D+Dubins define get_trajectory but never call it (their v_pref is the Dubins tracker) */
        compute_v_pref_straight(w, a, is_orca, policy != POL_SCA && policy != POL_RVO3D_DUBINS, v_store, v_des);
    }
    for (int d = 0; d < 3; d++) { w->vpref_in[3 * a + d] = v_des[d]; w->vpref_store[3 * a + d] = v_store[d]; }
    w->ran[a] = 1;
    SelInfo info = { 0, 0 };
    NbrList L;
    L.cnt = 0; L.is_collision = 0;
    int have_nbrs = 0;
    double vA_post[3];
    const float *vA = &w->vel[3 * a];
    int first;
    if (is_orca) {
        compute_neighbors(w, a, nbr_mode, &L);          /* before the gate (orca3dPolicyOfficial.py:50) */
        have_nbrs = 1;
        first = py_round5(sqrt(sumsq_f32(vA)) + 1e-5) <= 1e-5;
    } else {
        first = py_round5(sqrt(sumsq_f32(vA))) <= 1e-5;
    }
    w->plane_fail[a] = -1; w->n_planes[a] = -1; w->lp4_proj[a] = 0; w->lp4_inner_fail[a] = 0;
    if (first) {
        info.tag = TAG_FIRST;
        for (int d = 0; d < 3; d++) vA_post[d] = 0.3 * v_des[d];
    } else {
        if (!is_orca) { compute_neighbors(w, a, nbr_mode, &L); have_nbrs = 1; }
        Cone cones[K_MAX];
        Plane planes[K_MAX];
        int nc = L.cnt;
        for (int k = 0; k < nc; k++) {
            Obj o;
            get_obj(w, L.id[k], &o);
            if (is_orca) {
                build_plane(w, a, &o, &planes[k]);
            } else {
                Cone *c = &cones[k];
                volatile double agent_rad = w->radius[a] + 0.05;
                volatile double obj_rad = o.r + 0.05;
                for (int d = 0; d < 3; d++) {
                    c->pA[d] = w->pos[3 * a + d];
                    c->pB[d] = o.p[d];
                    if (o.at_goal) c->p0[d] = c->pA[d];
                    else {
                        volatile float s = o.v[d] + vA[d];
                        volatile float h = 0.5f * s;
                        c->p0[d] = c->pA[d] + (double)h;
                    }
                }
                c->R = obj_rad + agent_rad;
            }
        }
        if (policy == POL_ORCA_LP) {
            LPState S; S.v[0] = S.v[1] = S.v[2] = 0.0; S.lp4_proj = 0; S.lp4_inner_fail = 0;
            int pf = lp3(&S, planes, nc, MAX_SPEED, v_des, 0);
            w->plane_fail[a] = pf; w->n_planes[a] = nc;
            if (pf == nc) info.tag |= TAG_LP3_OK;
            if (pf < nc) { info.tag |= TAG_LP4; lp4(&S, planes, nc, pf, MAX_SPEED); }
            w->lp4_proj[a] = S.lp4_proj; w->lp4_inner_fail[a] = S.lp4_inner_fail;
            for (int d = 0; d < 3; d++) vA_post[d] = S.v[d];
        } else {
            sampled_select(w, a, policy, v_des, cones, planes, nc, vA_post, &info);
        }
    }
    double act[7];
    cartesian2spherical(w, a, vA_post, policy == POL_ORCA_LP, act);
    for (int k = 0; k < 7; k++) { w->action64[7 * a + k] = act[k]; w->action[7 * a + k] = (float)act[k]; }
    w->tag[a] = info.tag;
    w->n_suit[a] = info.n_suit;
    w->coll_query[a] = L.is_collision;
    if (have_nbrs) {
        w->nbr_cnt[a] = L.cnt;
        for (int k = 0; k < K_MAX; k++) {
            w->nbr_ids[K_MAX * a + k] = k < L.cnt ? L.id[k] : -1;
            w->nbr_dsq[K_MAX * a + k] = k < L.cnt ? L.dsq[k] : 0.0;
        }
    } else {
        w->nbr_cnt[a] = -1;   /* computeNeighbors not called this step: the reference list is stale */
    }
}

/* The reference is single-threaded; agents are independent inside the policy loop, so the CPU baseline
 * may use every host thread (orc_set_threads).  Results do not depend on the thread count. */
#include <pthread.h>
static int g_threads = 1;
void orc_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }
typedef struct { OrcWorld *w; int policy, nbr_mode; volatile int *next; } PJob;
static void *policy_worker(void *arg)
{
    PJob *J = (PJob *)arg;
    for (;;) {
        int a0 = __sync_fetch_and_add(J->next, 64);
        if (a0 >= J->w->n) break;
        int a1 = a0 + 64 < J->w->n ? a0 + 64 : J->w->n;
        for (int a = a0; a < a1; a++) {
            if (J->w->status[a] & (ST_AT_GOAL | ST_COLLISION | ST_TIMEOUT)) continue;
            policy_one(J->w, a, J->policy, J->nbr_mode);
        }
    }
    return NULL;
}
static void parallel_policy(OrcWorld *w, int policy, int nbr_mode)
{
    volatile int next = 0;
    PJob J = { w, policy, nbr_mode, &next };
    if (g_threads <= 1) { policy_worker(&J); return; }
    pthread_t th[256];
    int nt = g_threads;
    for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, policy_worker, &J);
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

/* mampenv.py:27-40: kd-tree rebuild + policy loop over active agents */
void orc_policy_step(OrcWorld *w, int policy, int nbr_mode)
{
    if (nbr_mode == NBR_REF_COMPAT) orc_kdtree_build_agents(w);
    for (int i = 0; i < w->n * 7; i++) { w->action[i] = 0.f; w->action64[i] = 0.0; }
    for (int a = 0; a < w->n; a++) { w->ran[a] = 0; w->coll_query[a] = 0; w->tag[a] = 0; w->n_suit[a] = 0; w->nbr_cnt[a] = -1; }
    if (nbr_mode == NBR_KNN) { g_knn_n = w->n; grid_build(w); }
    parallel_policy(w, policy, nbr_mode);
    /* insertAgentNeighbor sets agent.is_collision on the querying agent itself (agent.py:83-84) */
    for (int a = 0; a < w->n; a++)
        if (w->ran[a] && w->coll_query[a]) w->status[a] |= ST_COLLISION;
}

/* mampenv.py:83-105 update_velocitie */
static void update_one(OrcWorld *w, int a)
{
    const float *act = &w->action[7 * a];
    double speed = (double)act[3];
    double al = pi_2_pi(w->heading[3 * a] + (double)act[4]);
    double be = pi_2_pi(w->heading[3 * a + 1] + (double)act[5]);
    double ga = pi_2_pi(w->heading[3 * a + 2] + (double)act[6]);
    volatile double t = speed * cos(be);
    volatile double tx = t * cos(al);
    double dx = tx * TIME_STEP;
    volatile double ty = t * sin(al);
    double dy = ty * TIME_STEP;
    volatile double tz = speed * sin(be);
    double dz = tz * TIME_STEP;
    volatile double s = dx * dx;
    volatile double q = dy * dy;
    s = s + q;
    q = dz * dz;
    s = s + q;
    w->total_dist[a] += sqrt(s);
    w->pos[3 * a] += dx; w->pos[3 * a + 1] += dy; w->pos[3 * a + 2] += dz;
    w->heading[3 * a] = al; w->heading[3 * a + 1] = be; w->heading[3 * a + 2] = ga;
    w->vel[3 * a] = act[0]; w->vel[3 * a + 1] = act[1]; w->vel[3 * a + 2] = act[2];
}

/* mampenv.py:42-46 + :61-80 (literal, sequential, O(N^2)) then is_done :51-59 */
int orc_env_update(OrcWorld *w)
{
    int n = w->n;
    for (int i = 0; i < n; i++) {
        update_one(w, i);
        if (!(w->status[i] & ST_AT_GOAL)) w->step_num[i] += 1;
        for (int o = 0; o < w->m; o++) {
            double d = l3norm(&w->pos[3 * i], &w->obs[4 * o]);
            if (d <= w->radius[i] + w->obs[4 * o + 3]) w->status[i] |= ST_COLLISION;
        }
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            double dx = w->pos[3 * i] - w->pos[3 * j];
            if (dx > 4.0 || dx < -4.0) continue;    /* radii are O(1): pure speed-up, same result */
            double d = l3norm(&w->pos[3 * i], &w->pos[3 * j]);
            if (d <= w->radius[i] + w->radius[j]) {
                if (!(w->status[j] & ST_AT_GOAL)) w->status[j] |= ST_COLLISION;
                if (!(w->status[i] & ST_AT_GOAL)) w->status[i] |= ST_COLLISION;
            }
        }
        if (w->total_dist[i] > w->max_run_dist[i]) w->status[i] |= ST_TIMEOUT;
    }
    int all = 1;
    for (int i = 0; i < n; i++) {
        if (l3norm(&w->pos[3 * i], &w->goal[3 * i]) <= NEAR_GOAL) w->status[i] |= ST_AT_GOAL;
        if (w->status[i] & (ST_AT_GOAL | ST_COLLISION | ST_TIMEOUT)) w->status[i] |= ST_RUN_DONE;
        if (!(w->status[i] & ST_RUN_DONE)) all = 0;
    }
    return all;
}

/* Same result as orc_env_update, O(N) with a hash grid: pair (lo<hi) collides iff
 * d(new_lo, old_hi) <= R or d(new_lo, new_hi) <= R (derivation in DESIGN.md).  Used for large-N
 * baselines; tests check it against the literal version. */
int orc_env_update_fast(OrcWorld *w)
{
    int n = w->n;
    double *oldp = (double *)malloc(sizeof(double) * 3 * n);
    memcpy(oldp, w->pos, sizeof(double) * 3 * n);
    int32_t *was_goal = (int32_t *)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) {
        was_goal[i] = w->status[i] & ST_AT_GOAL;
        update_one(w, i);
        if (!was_goal[i]) w->step_num[i] += 1;
    }
    /* grid over old positions, cell = 4 m (max radius sum assumed < 3.8 m) */
    double mn[3] = { 1e300, 1e300, 1e300 };
    for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) if (oldp[3 * i + c] < mn[c]) mn[c] = oldp[3 * i + c];
    const double cs = 4.0;
    uint32_t H = 1; while (H < (uint32_t)(2 * n + 16)) H <<= 1;
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * H), *next = (int32_t *)malloc(sizeof(int32_t) * n);
    for (uint32_t k = 0; k < H; k++) head[k] = -1;
#define CELLH(cx, cy, cz) ((((uint32_t)(cx) * 73856093u) ^ ((uint32_t)(cy) * 19349663u) ^ ((uint32_t)(cz) * 83492791u)) & (H - 1))
    for (int i = n - 1; i >= 0; i--) {
        int cx = (int)floor((oldp[3 * i] - mn[0]) / cs), cy = (int)floor((oldp[3 * i + 1] - mn[1]) / cs), cz = (int)floor((oldp[3 * i + 2] - mn[2]) / cs);
        uint32_t h = CELLH(cx, cy, cz);
        next[i] = head[h]; head[h] = i;
    }
    uint8_t *coll = (uint8_t *)calloc(n, 1);
    for (int i = 0; i < n; i++) {
        for (int o = 0; o < w->m; o++) {
            double d = l3norm(&w->pos[3 * i], &w->obs[4 * o]);
            if (d <= w->radius[i] + w->obs[4 * o + 3]) w->status[i] |= ST_COLLISION;
        }
        int cx = (int)floor((oldp[3 * i] - mn[0]) / cs), cy = (int)floor((oldp[3 * i + 1] - mn[1]) / cs), cz = (int)floor((oldp[3 * i + 2] - mn[2]) / cs);
        for (int ax = -1; ax <= 1; ax++) for (int ay = -1; ay <= 1; ay++) for (int az = -1; az <= 1; az++) {
            uint32_t h = CELLH(cx + ax, cy + ay, cz + az);
            /* distinct cells may share a bucket: a bucket may be visited more than once (harmless) */
            for (int j = head[h]; j >= 0; j = next[j]) {
                if (j <= i) continue;
                double R = w->radius[i] + w->radius[j];
                if (l3norm(&w->pos[3 * i], &oldp[3 * j]) <= R || l3norm(&w->pos[3 * i], &w->pos[3 * j]) <= R) {
                    coll[i] = 1; coll[j] = 1;
                }
            }
        }
    }
#undef CELLH
    int all = 1;
    for (int i = 0; i < n; i++) {
        if (coll[i] && !was_goal[i]) w->status[i] |= ST_COLLISION;
        if (w->total_dist[i] > w->max_run_dist[i]) w->status[i] |= ST_TIMEOUT;
    }
    for (int i = 0; i < n; i++) {
        if (l3norm(&w->pos[3 * i], &w->goal[3 * i]) <= NEAR_GOAL) w->status[i] |= ST_AT_GOAL;
        if (w->status[i] & (ST_AT_GOAL | ST_COLLISION | ST_TIMEOUT)) w->status[i] |= ST_RUN_DONE;
        if (!(w->status[i] & ST_RUN_DONE)) all = 0;
    }
    free(oldp); free(was_goal); free(head); free(next); free(coll);
    return all;
}

/* MACAEnv.step (mampenv.py:22-25) */
int orc_step(OrcWorld *w, int policy, int nbr_mode, int fast_update)
{
    orc_policy_step(w, policy, nbr_mode);
    return fast_update ? orc_env_update_fast(w) : orc_env_update(w);
}

int orc_sizeof_world(void) { return (int)sizeof(OrcWorld); }
int orc_sizeof_kdnode(void) { return (int)sizeof(KdNode); }
