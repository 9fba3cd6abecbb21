// scab_api.cu -- the C ABI of libscab200.so (include/scab200.h): handle, library-owned scratch,
// static obstacle structures, the host-side replay of the reference's kd-tree build (REF_COMPAT), and
// the stage launchers.  No torch types, no exceptions across the boundary.
#include "scab_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define TIER_DELTA_DEFAULT 0.28f  // round 2, with the second pass at + 0.15 (policy stage, 1M SCA): 0.24 -> 1.244 ms, 0.26 -> 1.233, 0.28 -> 1.239 (best whole
                                  // step), 0.30 -> 1.244, 0.31 -> 1.288, 0.33 -> 1.41 (one more round of cone tests for most agents).  Round 1: 0.12 -> 5.0 ms, 0.18 -> 4.1, 0.25 -> 3.0, 0.4 -> 3.0, 0.55 -> 3.6, 0.75 -> 4.9 (policy stage, 1M SCA)

struct scab_handle {
    int device = 0;
    int64_t capacity = 0;
    scab_params p{};
    StepArgs A{};
    LaunchCfg cfg{};
    bool bound = false;
    std::string err;
    // library-owned device scratch
    void *d_block = nullptr;
    int32_t *d_cell_count = nullptr, *d_cell_start = nullptr;
    int64_t cell_cap = 0;
    double *d_dir256 = nullptr, *d_dir128 = nullptr;
    float4 *d_dir256f = nullptr, *d_dir128f = nullptr;
    uint32_t *d_plut256 = nullptr, *d_plut128 = nullptr;
    unsigned long long *d_counters = nullptr;
    // obstacles
    int32_t m = 0;
    double4 *d_o_posr = nullptr, *d_o_posr_by_id = nullptr;
    int32_t *d_o_id = nullptr, *d_o_cell_start = nullptr, *d_kd_obst_perm = nullptr;
    KdNodeDev *d_kd_obst = nullptr;
    // reference kd-tree over agents (REF_COMPAT)
    std::vector<int32_t> kd_perm;          // persistent agentIDs (kdTree.py:43-45)
    std::vector<KdNodeDev> kd_nodes;
    KdNodeDev *d_kd_agent = nullptr;
    int32_t *d_kd_agent_perm = nullptr;
    bool kd_valid = false;
    // halo bookkeeping (which owned rows were packed for each face, and how many)
    int32_t *d_sent_lo = nullptr, *d_sent_hi = nullptr, *d_halo_cnt = nullptr;
    // migration bookkeeping: rows vacated by the last scab_migrate_pack / rows given to the arrivals of the last unpack
    int32_t *d_mig_sent_row = nullptr, *d_mig_recv_row = nullptr;
    int32_t mig_cap = 0;
    // peer-to-peer exchange buffers (scab_p2p_alloc) and peer mappings (scab_p2p_open)
    std::vector<void *> p2p_own, p2p_peer;
    P2PWait pending_wait{nullptr, nullptr, nullptr};      // scab_p2p_wait_next: consumed by the next unpack launch
    // Dubins v_pref provider (allocated by the first scab_dubins_enable)
    void *d_dub_states = nullptr;
    double *d_dub_d0 = nullptr;
};

static int fail(scab_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(h, SCAB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

extern "C" const char *scab_version(void) { return "scab200 0.1 (sm_100a)"; }

extern "C" void scab_default_params(scab_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->neighbor_dist = 10.0;   // agent.py:33
    p->time_step = 0.1;        // config.py:2
    p->time_horizon = 10.0;    // agent.py:35
    p->max_speed = 1.0;        // agent.py:36
    p->near_goal = 0.5;        // config.py:3
    p->grid_origin[0] = p->grid_origin[1] = p->grid_origin[2] = 0.0;
    p->cell_size = 10.0;
    p->grid_dim[0] = p->grid_dim[1] = p->grid_dim[2] = 1;
}

extern "C" const char *scab_last_error(const scab_handle *h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int scab_set_max_radius(scab_handle *h, double r_max)
{
    if (!h || !(r_max >= 0.0)) return fail(h, SCAB_ERR_INVALID, "scab_set_max_radius: bad argument");
    h->A.max_radius = r_max;
    return SCAB_OK;
}

extern "C" int scab_set_motion_bound(scab_handle *h, double v_max)
{
    if (!h || !(v_max > 0.0) || !std::isfinite(v_max)) return fail(h, SCAB_ERR_INVALID, "scab_set_motion_bound: bad argument");
    // 0.1 % slack: v_pref is dif * pref / trunc5(|dif|) (rvo3dPolicy.py:186-190), i.e. up to 5e-5 faster than pref near the goal
    h->A.move_bound = v_max * h->p.time_step * 1.001 + 1e-9;
    return SCAB_OK;
}

extern "C" int scab_set_option(scab_handle *h, int option, int value)
{
    if (!h) return SCAB_ERR_INVALID;
    if (option == SCAB_OPT_FULL_SEARCH) { h->A.full_search = value ? 1 : 0; return SCAB_OK; }
    return fail(h, SCAB_ERR_INVALID, "scab_set_option: unknown option");
}

// Fibonacci-sphere unit directions (scaPolicy.py:196-199), evaluated with the host libm exactly as the
// reference does (math.sqrt / math.cos / math.sin), then uploaded.
static void make_dirs(int N, std::vector<double> &d64, std::vector<float4> &d32)
{
    const double pi = 3.141592653589793;
    const double param_phi = (sqrt(5.0) - 1.0) / 2.0;
    d64.resize(3 * (size_t)N);
    d32.resize(N);
    for (int n = 1; n <= N; n++) {
        volatile double zq = (double)(2 * n - 1) / N;
        double z = zq - 1;
        volatile double z2 = z * z;
        double sq = sqrt(1 - z2);
        volatile double a1 = 2 * pi;
        volatile double a2 = a1 * n;
        double ang = a2 * param_phi;
        volatile double x = sq * cos(ang);
        volatile double y = sq * sin(ang);
        d64[3 * (n - 1)] = x; d64[3 * (n - 1) + 1] = y; d64[3 * (n - 1) + 2] = z;
        d32[n - 1] = make_float4((float)x, (float)y, (float)z, 0.f);
    }
}

// Posture lookup table of k_select_fast (scab_select_fast.cuh posture_bin is the forward map).  For every bin of the
// LUT_B x LUT_B octahedral map: c = direction of the bin centre, rho = largest angle between c and any direction of the bin
// (sampled on a 9 x 9 lattice over the bin grown by 5 %, + 2 mrad for the lattice and for fp32 rounding in posture_bin).
// Direction u passes util.py:6-20's angle test for EVERY vA of the bin if angle(u, c) <= pi/4 - rho - 1 mrad ("sure"), for
// none if angle(u, c) >= pi/4 + rho + 1 mrad, and is tested by the kernel otherwise ("maybe").
static void make_posture_lut(int N, const std::vector<double> &d64, std::vector<uint32_t> &lut)
{
    const int B = SCAB_POSTURE_LUT_B;
    lut.assign((size_t)B * B * 16, 0u);
    auto decode = [](double u, double v, double out[3]) {
        double z = 1.0 - std::fabs(u) - std::fabs(v), x = u, y = v;
        if (z < 0.0) {
            x = (1.0 - std::fabs(v)) * (u >= 0.0 ? 1.0 : -1.0);
            y = (1.0 - std::fabs(u)) * (v >= 0.0 ? 1.0 : -1.0);
        }
        const double l = std::sqrt(x * x + y * y + z * z);
        out[0] = x / l; out[1] = y / l; out[2] = z / l;
    };
    const double quarter = 0.78539816339744831;
    for (int iy = 0; iy < B; iy++)
        for (int ix = 0; ix < B; ix++) {
            const double u0 = -1.0 + 2.0 * ix / B, v0 = -1.0 + 2.0 * iy / B, w = 2.0 / B;
            double c[3];
            decode(u0 + 0.5 * w, v0 + 0.5 * w, c);
            double rho = 0.0;
            for (int a = 0; a <= 8; a++)
                for (int b = 0; b <= 8; b++) {
                    const double uu = std::min(1.0, std::max(-1.0, u0 + w * (-0.05 + 1.1 * a / 8.0)));
                    const double vv = std::min(1.0, std::max(-1.0, v0 + w * (-0.05 + 1.1 * b / 8.0)));
                    double q[3];
                    decode(uu, vv, q);
                    const double dt = std::min(1.0, std::max(-1.0, q[0] * c[0] + q[1] * c[1] + q[2] * c[2]));
                    rho = std::max(rho, std::acos(dt));
                }
            rho += 2e-3;
            const double cos_sure = std::cos(std::max(0.0, quarter - rho - 1e-3)), cos_never = std::cos(quarter + rho + 1e-3);
            const bool no_sure = quarter - rho - 1e-3 <= 0.0;
            uint32_t *row = &lut[((size_t)iy * B + ix) * 16];
            for (int n = 0; n < N; n++) {
                const double dt = d64[3 * n] * c[0] + d64[3 * n + 1] * c[1] + d64[3 * n + 2] * c[2];
                const int lane = n & 7, bit = n >> 3;
                if (!no_sure && dt >= cos_sure) row[lane] |= 1u << bit;
                else if (dt > cos_never) row[8 + lane] |= 1u << bit;
            }
        }
}

static void set_geom(GridGeom &g, const double origin[3], double cs, const int32_t dim[3])
{
    g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
    g.inv_cs = 1.0 / cs;
    g.nx = std::max(1, dim[0]); g.ny = std::max(1, dim[1]); g.nz = std::max(1, dim[2]);
    g.ncell = g.nx * g.ny * g.nz;
}

static int ensure_cells(scab_handle *h, int64_t ncell)
{
    if (ncell <= h->cell_cap) return SCAB_OK;
    if (h->d_cell_count) cudaFree(h->d_cell_count);
    if (h->d_cell_start) cudaFree(h->d_cell_start);
    h->d_cell_count = h->d_cell_start = nullptr;
    h->cell_cap = 0;
    CK(cudaMalloc(&h->d_cell_count, sizeof(int32_t) * (size_t)ncell));
    CK(cudaMalloc(&h->d_cell_start, sizeof(int32_t) * (size_t)(ncell + 1)));
    h->cell_cap = ncell;
    h->A.cell_count = h->d_cell_count;
    h->A.cell_start = h->d_cell_start;
    return SCAB_OK;
}

extern "C" int scab_set_grid(scab_handle *h, const double origin[3], double cell_size, const int32_t dim[3])
{
    if (!h || !origin || !dim) return fail(h, SCAB_ERR_INVALID, "scab_set_grid: null argument");
    if (!(cell_size >= h->p.neighbor_dist)) return fail(h, SCAB_ERR_INVALID, "scab_set_grid: cell_size < neighbor_dist");
    int64_t nc = (int64_t)std::max(1, dim[0]) * std::max(1, dim[1]) * std::max(1, dim[2]);
    if (nc > (int64_t)1 << 28) return fail(h, SCAB_ERR_INVALID, "scab_set_grid: too many cells");
    CK(cudaSetDevice(h->device));
    set_geom(h->A.g, origin, cell_size, dim);
    for (int d = 0; d < 3; d++) { h->p.grid_origin[d] = origin[d]; h->p.grid_dim[d] = dim[d]; }
    h->p.cell_size = cell_size;
    return ensure_cells(h, nc);
}

extern "C" int scab_create(scab_handle **out, const scab_params *p, int device, int64_t capacity)
{
    if (!out || !p || capacity <= 0) return SCAB_ERR_INVALID;
    *out = nullptr;
    // the 27-cell scans need cell_size >= neighbor_dist; the plane / integration arithmetic divides by these
    const double chk[] = {p->neighbor_dist, p->time_step, p->time_horizon, p->max_speed, p->near_goal, p->cell_size,
                          p->grid_origin[0], p->grid_origin[1], p->grid_origin[2]};
    for (double v : chk) if (!std::isfinite(v)) return SCAB_ERR_INVALID;
    if (!(p->neighbor_dist > 0.0) || !(p->time_step > 0.0) || !(p->time_horizon > 0.0) || !(p->max_speed > 0.0) ||
        !(p->cell_size >= p->neighbor_dist) || p->grid_dim[0] < 1 || p->grid_dim[1] < 1 || p->grid_dim[2] < 1)
        return SCAB_ERR_INVALID;
    scab_handle *h = new (std::nothrow) scab_handle();
    if (!h) return SCAB_ERR_NOMEM;
    h->device = device;
    h->capacity = capacity;
    h->p = *p;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return SCAB_ERR_CUDA; }
    StepArgs &A = h->A;
    A.nbr_dist_sq = p->neighbor_dist * p->neighbor_dist;   // agent.neighborDist ** 2 (scaPolicy.py:112)
    A.time_step = p->time_step;
    A.inv_time_horizon = 1.0 / p->time_horizon;
    A.inv_time_step = 1.0 / p->time_step;
    A.max_speed = p->max_speed;
    A.near_goal = p->near_goal;
    A.tier_delta = TIER_DELTA_DEFAULT;
    A.row_begin = 0; A.row_count = -1;
    // until scab_set_motion_bound: no agent moves more than 1.25 m per step (12.5 m/s at the reference's dt)
    A.move_bound = 1.25;
    {   // launch configuration of THIS handle: its device's SM count, the A/B switches as they are set right now
        LaunchCfg &c = h->cfg;
        if (cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c.sms <= 0) c.sms = 148;
        auto on = [](const char *name) { const char *v = getenv(name); return (v && v[0] == '1') ? 1 : 0; };
        c.no_fast = on("SCAB_NO_FAST");        // every agent through the general kernels (sampled policies and ORCA-LP)
        c.lp_f64 = on("SCAB_LP_F64");          // ORCA-LP register-only pass without the fp32 pre-filter
        auto off = [](const char *name) { const char *v = getenv(name); return (v && v[0] == '0') ? 0 : 1; };
        // both measured faster on B200 (policy stage 0.574 -> 0.534 ms at 10^6 agents, profiles/r2_summary.md): default on
        c.lp_compact = off("SCAB_LP_COMPACT");     // ORCA-LP list pass with 32-byte planes (7 resident CTAs instead of 4); =0: 48-byte planes
        c.lp_ctas8 = off("SCAB_LP_PASS1_CTAS8");   // ORCA-LP register-only pass at 64 registers (8 resident CTAs); =0: 80 registers
        c.knn_exact = on("SCAB_KNN_EXACT");    // 16-NN with fp64 total-order keys for every agent
        A.cfg = &h->cfg;
    }
    if (const char *td = getenv("SCAB_TIER_DELTA")) { float v = (float)atof(td); if (v > 0.f) A.tier_delta = v; }   // tuning knob
    // one block for all per-agent scratch
    const size_t cap = (size_t)capacity;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_posr = carve(cap * sizeof(double4)), o_gid = carve(cap * 4),
                 o_vdes = carve(cap * sizeof(double4)), o_vpost = carve(cap * sizeof(double4)),
                 o_vel = carve(cap * sizeof(float4)), o_posf = carve(cap * sizeof(float4)), o_id = carve(cap * 4), o_rank = carve(cap * 4),
                 o_cid = carve(cap * 4), o_cloc = carve(cap * 4), o_nbr = carve(cap * 4 * KNBR),
                 o_ncnt = carve(cap * 4), o_qc = carve(cap * 4), o_r16 = carve(cap * 4), o_kdp = carve(cap * 4),
                 o_slo = carve(cap * 4), o_shi = carve(cap * 4), o_hcnt = carve(64),
                 o_slow = carve(cap * 4), o_slowc = carve(64), o_slow2 = carve(cap * 4), o_srow = carve(cap * 4), o_rrow = carve(cap * 4),
                 o_kdn = carve((2 * cap) * sizeof(KdNodeDev));
    e = cudaMalloc(&h->d_block, off);
    if (e != cudaSuccess) { delete h; return SCAB_ERR_NOMEM; }
    char *b = (char *)h->d_block;
    A.s_posr = (double4 *)(b + o_posr); A.s_gid = (int32_t *)(b + o_gid); A.s_vel = (float4 *)(b + o_vel);
    A.s_posf = (float4 *)(b + o_posf);
    A.s_vdes = (double4 *)(b + o_vdes); A.s_vpost = (double4 *)(b + o_vpost);
    A.s_id = (int32_t *)(b + o_id); A.rank = (int32_t *)(b + o_rank);
    A.cell_id = (int32_t *)(b + o_cid); A.cell_local = (int32_t *)(b + o_cloc);
    A.s_nbr = (int32_t *)(b + o_nbr); A.s_nbr_cnt = (int32_t *)(b + o_ncnt); A.s_qcoll = (int32_t *)(b + o_qc);
    h->d_kd_agent_perm = (int32_t *)(b + o_kdp); h->d_kd_agent = (KdNodeDev *)(b + o_kdn);
    A.slow_list = (int32_t *)(b + o_slow); A.slow_count = (int32_t *)(b + o_slowc);
    A.r16_prev = (float *)(b + o_r16);
    cudaMemset(A.r16_prev, 0x7f, cap * 4);        // 3.4e38: no previous query
    A.slow_list2 = (int32_t *)(b + o_slow2); A.slow_count2 = A.slow_count + 1;      // adjacent: one memset resets both
    h->d_mig_sent_row = (int32_t *)(b + o_srow); h->d_mig_recv_row = (int32_t *)(b + o_rrow);
    h->d_sent_lo = (int32_t *)(b + o_slo); h->d_sent_hi = (int32_t *)(b + o_shi); h->d_halo_cnt = (int32_t *)(b + o_hcnt);
    A.kd_agent = h->d_kd_agent; A.kd_agent_perm = h->d_kd_agent_perm;
    // candidate tables
    std::vector<double> d64; std::vector<float4> d32;
    bool ok = true;
    std::vector<uint32_t> lut;
    make_dirs(256, d64, d32);
    make_posture_lut(256, d64, lut);
    ok &= cudaMalloc(&h->d_dir256, d64.size() * 8) == cudaSuccess && cudaMalloc(&h->d_dir256f, d32.size() * 16) == cudaSuccess &&
          cudaMalloc(&h->d_plut256, lut.size() * 4) == cudaSuccess;
    if (ok) { cudaMemcpy(h->d_dir256, d64.data(), d64.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(h->d_dir256f, d32.data(), d32.size() * 16, cudaMemcpyHostToDevice);
              cudaMemcpy(h->d_plut256, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice); }
    make_dirs(128, d64, d32);
    make_posture_lut(128, d64, lut);
    ok &= cudaMalloc(&h->d_dir128, d64.size() * 8) == cudaSuccess && cudaMalloc(&h->d_dir128f, d32.size() * 16) == cudaSuccess &&
          cudaMalloc(&h->d_plut128, lut.size() * 4) == cudaSuccess;
    if (ok) { cudaMemcpy(h->d_dir128, d64.data(), d64.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(h->d_dir128f, d32.data(), d32.size() * 16, cudaMemcpyHostToDevice);
              cudaMemcpy(h->d_plut128, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice); }
    ok &= cudaMalloc(&h->d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess;
    if (!ok) { scab_destroy(h); return SCAB_ERR_NOMEM; }
    cudaMemset(h->d_counters, 0, 8 * sizeof(unsigned long long));
    A.dir256 = h->d_dir256; A.dir128 = h->d_dir128; A.dir256f = h->d_dir256f; A.dir128f = h->d_dir128f;
    A.plut256 = h->d_plut256; A.plut128 = h->d_plut128;
    A.counters = h->d_counters;
    set_geom(A.g, p->grid_origin, p->cell_size, p->grid_dim);
    int32_t one[3] = {1, 1, 1};
    double zero[3] = {0, 0, 0};
    set_geom(A.og, zero, p->cell_size, one);
    int rc = ensure_cells(h, A.g.ncell);
    if (rc != SCAB_OK) { scab_destroy(h); return rc; }
    *out = h;
    return SCAB_OK;
}

extern "C" int scab_destroy(scab_handle *h)
{
    if (!h) return SCAB_ERR_INVALID;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_block, h->d_cell_count, h->d_cell_start, h->d_dir256, h->d_dir128, h->d_dir256f, h->d_dir128f, h->d_plut256, h->d_plut128,
                    h->d_counters, h->d_o_posr, h->d_o_posr_by_id, h->d_o_id, h->d_o_cell_start, h->d_kd_obst_perm, h->d_kd_obst,
                    h->d_dub_states, h->d_dub_d0};
    for (void *q : ptrs) if (q) cudaFree(q);
    for (void *q : h->p2p_peer) cudaIpcCloseMemHandle(q);
    for (void *q : h->p2p_own) cudaFree(q);
    delete h;
    return SCAB_OK;
}

extern "C" int scab_bind_state(scab_handle *h, const scab_state_ptrs *s, int64_t n)
{
    if (!h || !s) return fail(h, SCAB_ERR_INVALID, "scab_bind_state: null argument");
    if (n < 0 || n > h->capacity) return fail(h, SCAB_ERR_CAPACITY, "scab_bind_state: n exceeds capacity");
    if (!s->pos || !s->vel || !s->heading || !s->goal || !s->radius || !s->pref_speed || !s->status || !s->total_dist ||
        !s->step_num || !s->max_run_dist || !s->action)
        return fail(h, SCAB_ERR_INVALID, "scab_bind_state: a required state pointer is NULL");
    h->A.st = *s;
    h->A.n = (int)n;
    cudaMemset(h->A.r16_prev, 0x7f, (size_t)h->capacity * 4);      // the rows may hold other agents now
    h->A.obst_base = (int)(s->n_global > 0 ? s->n_global : n);
    h->bound = true;
    h->kd_valid = false;
    if ((int64_t)h->kd_perm.size() != n) {   // KDTree.__init__: agentIDs = [0..n) (kdTree.py:43-45)
        h->kd_perm.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) h->kd_perm[(size_t)i] = (int32_t)i;
    }
    return SCAB_OK;
}

// ------------------------------------------------------------------ reference kd-tree (host replay)
// kdTree.py:60-122 buildAgentTreeRecursive / :162-227 buildObstacleTreeRecursive
static void kd_build(const double *pos, int stride, int32_t *ids, KdNodeDev *nodes, int begin, int end, int node)
{
    KdNodeDev &nd = nodes[node];
    nd.begin = begin; nd.end = end; nd.left = nd.right = 0;
    for (int c = 0; c < 3; c++) nd.mn[c] = nd.mx[c] = pos[(size_t)stride * ids[begin] + c];
    for (int i = begin + 1; i < end; i++)
        for (int c = 0; c < 3; c++) {
            const double v = pos[(size_t)stride * ids[i] + c];
            nd.mx[c] = std::max(nd.mx[c], v);
            nd.mn[c] = std::min(nd.mn[c], v);
        }
    if (end - begin > 10) {   // max_leaf_size, kdTree.py:53
        const double d0 = nd.mx[0] - nd.mn[0], d1 = nd.mx[1] - nd.mn[1], d2 = nd.mx[2] - nd.mn[2];
        const int coord = (d0 > d1 && d0 > d2) ? 0 : (d1 > d2 ? 1 : 2);
        const double split = 0.5 * (nd.mx[coord] + nd.mn[coord]);
        int left = begin, right = end;
        while (left < right) {
            while (left < right && pos[(size_t)stride * ids[left] + coord] < split) left++;
            while (right > left && pos[(size_t)stride * ids[right - 1] + coord] >= split) right--;
            if (left < right) {
                std::swap(ids[left], ids[right - 1]);
                left++;
                right--;
            }
        }
        int leftSize = left - begin;
        if (leftSize == 0) { leftSize++; left++; right++; }
        nd.left = node + 1;
        nd.right = node + 2 * leftSize;
        kd_build(pos, stride, ids, nodes, begin, left, nd.left);
        kd_build(pos, stride, ids, nodes, left, end, nd.right);
    }
}

extern "C" int scab_kdtree_build(scab_handle *h, const double *pos_host, int64_t n)
{
    if (!h || !pos_host) return fail(h, SCAB_ERR_INVALID, "scab_kdtree_build: null argument");
    if (!h->bound || n != h->A.n) return fail(h, SCAB_ERR_INVALID, "scab_kdtree_build: n does not match the bound state");
    if (n == 0) return SCAB_OK;
    CK(cudaSetDevice(h->device));
    h->kd_nodes.assign((size_t)(2 * n), KdNodeDev{});
    kd_build(pos_host, 3, h->kd_perm.data(), h->kd_nodes.data(), 0, (int)n, 0);
    CK(cudaMemcpy(h->d_kd_agent, h->kd_nodes.data(), sizeof(KdNodeDev) * (size_t)(2 * n - 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_kd_agent_perm, h->kd_perm.data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    h->kd_valid = true;
    return SCAB_OK;
}

extern "C" int scab_kdtree_get_perm(scab_handle *h, int32_t *out_host, int64_t n)
{
    if (!h || !out_host || n != (int64_t)h->kd_perm.size()) return fail(h, SCAB_ERR_INVALID, "scab_kdtree_get_perm: bad argument");
    memcpy(out_host, h->kd_perm.data(), sizeof(int32_t) * (size_t)n);
    return SCAB_OK;
}

extern "C" int scab_set_obstacles(scab_handle *h, const double *xyzr, int32_t m)
{
    if (!h || m < 0 || (m > 0 && !xyzr)) return fail(h, SCAB_ERR_INVALID, "scab_set_obstacles: bad argument");
    CK(cudaSetDevice(h->device));
    void *old[] = {h->d_o_posr, h->d_o_posr_by_id, h->d_o_id, h->d_o_cell_start, h->d_kd_obst_perm, h->d_kd_obst};
    for (void *q : old) if (q) cudaFree(q);
    h->d_o_posr = h->d_o_posr_by_id = nullptr; h->d_o_id = h->d_o_cell_start = h->d_kd_obst_perm = nullptr; h->d_kd_obst = nullptr;
    h->m = m;
    StepArgs &A = h->A;
    A.m = m;
    A.o_posr = A.o_posr_by_id = nullptr; A.o_id = A.o_cell_start = A.kd_obst_perm = nullptr; A.kd_obst = nullptr;
    if (m == 0) return SCAB_OK;
    // static obstacle cell list: cell >= neighbor_dist + r_max so that (|pA-pO| - r_O)^2 < D^2 stays
    // inside the 27-cell neighbourhood (agent.py:106)
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300}, rmax = 0.0;
    for (int i = 0; i < m; i++) {
        for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], xyzr[4 * i + c]); mx[c] = std::max(mx[c], xyzr[4 * i + c]); }
        rmax = std::max(rmax, xyzr[4 * i + 3]);
    }
    const double cs = h->p.neighbor_dist + rmax + 1e-6;
    int32_t dim[3];
    for (int c = 0; c < 3; c++) dim[c] = std::max(1, std::min(1024, (int)std::floor((mx[c] - mn[c]) / cs) + 1));
    set_geom(A.og, mn, cs, dim);
    std::vector<int32_t> cell(m), start((size_t)A.og.ncell + 1, 0), order(m);
    for (int i = 0; i < m; i++) {
        int cx = std::min(std::max((int)std::floor((xyzr[4 * i] - A.og.ox) * A.og.inv_cs), 0), A.og.nx - 1);
        int cy = std::min(std::max((int)std::floor((xyzr[4 * i + 1] - A.og.oy) * A.og.inv_cs), 0), A.og.ny - 1);
        int cz = std::min(std::max((int)std::floor((xyzr[4 * i + 2] - A.og.oz) * A.og.inv_cs), 0), A.og.nz - 1);
        cell[i] = (cz * A.og.ny + cy) * A.og.nx + cx;
        start[(size_t)cell[i] + 1]++;
    }
    for (int c = 0; c < A.og.ncell; c++) start[(size_t)c + 1] += start[c];
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < m; i++) order[fill[cell[i]]++] = i;   // stable: id order inside a cell
    std::vector<double4> sorted(m), byid(m);
    std::vector<int32_t> oid(m);
    for (int i = 0; i < m; i++) byid[i] = make_double4(xyzr[4 * i], xyzr[4 * i + 1], xyzr[4 * i + 2], xyzr[4 * i + 3]);
    for (int k = 0; k < m; k++) { sorted[k] = byid[order[k]]; oid[k] = order[k]; }
    // reference obstacle kd-tree (built once: mampenv.py:20)
    std::vector<int32_t> perm(m);
    for (int i = 0; i < m; i++) perm[i] = i;
    std::vector<KdNodeDev> nodes((size_t)2 * m);
    kd_build(xyzr, 4, perm.data(), nodes.data(), 0, m, 0);
    CK(cudaMalloc(&h->d_o_posr, sizeof(double4) * m));
    CK(cudaMalloc(&h->d_o_posr_by_id, sizeof(double4) * m));
    CK(cudaMalloc(&h->d_o_id, 4 * (size_t)m));
    CK(cudaMalloc(&h->d_o_cell_start, 4 * start.size()));
    CK(cudaMalloc(&h->d_kd_obst_perm, 4 * (size_t)m));
    CK(cudaMalloc(&h->d_kd_obst, sizeof(KdNodeDev) * nodes.size()));
    CK(cudaMemcpy(h->d_o_posr, sorted.data(), sizeof(double4) * m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_o_posr_by_id, byid.data(), sizeof(double4) * m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_o_id, oid.data(), 4 * (size_t)m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_o_cell_start, start.data(), 4 * start.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_kd_obst_perm, perm.data(), 4 * (size_t)m, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_kd_obst, nodes.data(), sizeof(KdNodeDev) * nodes.size(), cudaMemcpyHostToDevice));
    A.o_posr = h->d_o_posr; A.o_posr_by_id = h->d_o_posr_by_id; A.o_id = h->d_o_id; A.o_cell_start = h->d_o_cell_start;
    A.kd_obst_perm = h->d_kd_obst_perm; A.kd_obst = h->d_kd_obst;
    return SCAB_OK;
}

// ------------------------------------------------------------------ stages
static int check_ready(scab_handle *h, const char *who)
{
    if (!h) return SCAB_ERR_INVALID;
    if (!h->bound) return fail(h, SCAB_ERR_INVALID, std::string(who) + ": no state bound");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(h, SCAB_ERR_CUDA, cudaGetErrorString(e));
    return SCAB_OK;
}

extern "C" int scab_build_grid(scab_handle *h, void *stream)
{
    int rc = check_ready(h, "scab_build_grid");
    if (rc) return rc;
    CK(launch_grid_build(h->A, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_neighbors(scab_handle *h, int mode, void *stream)
{
    int rc = check_ready(h, "scab_neighbors");
    if (rc) return rc;
    if (mode != SCAB_NBR_KNN && mode != SCAB_NBR_REF_COMPAT) return fail(h, SCAB_ERR_INVALID, "scab_neighbors: bad mode");
    if (mode == SCAB_NBR_REF_COMPAT && !h->kd_valid && h->A.n > 0)
        return fail(h, SCAB_ERR_INVALID, "scab_neighbors: REF_COMPAT needs scab_kdtree_build first");
    CK(launch_neighbors(h->A, mode, (cudaStream_t)stream));
    h->A.last_nbr_mode = mode;
    return SCAB_OK;
}

extern "C" int scab_policy_step(scab_handle *h, int policy, void *stream)
{
    int rc = check_ready(h, "scab_policy_step");
    if (rc) return rc;
    if (policy < SCAB_POLICY_SCA || policy > SCAB_POLICY_ORCA_LP) return fail(h, SCAB_ERR_INVALID, "scab_policy_step: bad policy");
    if (policy == SCAB_POLICY_ORCA_LP) CK(launch_policy_lp(h->A, (cudaStream_t)stream));
    else CK(launch_policy_sampled(h->A, policy, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_policy_step_rows(scab_handle *h, int policy, int64_t row_begin, int64_t row_end, void *stream)
{
    int rc = check_ready(h, "scab_policy_step_rows");
    if (rc) return rc;
    if (policy < SCAB_POLICY_SCA || policy > SCAB_POLICY_ORCA_LP) return fail(h, SCAB_ERR_INVALID, "scab_policy_step_rows: bad policy");
    if (row_begin < 0 || row_end < row_begin || row_end > h->A.n) return fail(h, SCAB_ERR_INVALID, "scab_policy_step_rows: bad row range");
    StepArgs A = h->A;
    A.row_begin = (int)row_begin; A.row_count = (int)(row_end - row_begin);
    cudaError_t e = (policy == SCAB_POLICY_ORCA_LP) ? launch_policy_lp(A, (cudaStream_t)stream) : launch_policy_sampled(A, policy, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported) return fail(h, SCAB_ERR_INVALID, "scab_policy_step_rows: not available with the device Dubins tracker");
    CK(e);
    return SCAB_OK;
}

extern "C" int scab_status_step(scab_handle *h, void *stream)
{
    int rc = check_ready(h, "scab_status_step");
    if (rc) return rc;
    // every possible collision partner must lie in the 27-cell block around the agent's cell (mampenv.py:61-80 scans all agents)
    if (2.0 * h->A.max_radius + 2.0 * h->A.move_bound + 1e-3 > 1.0 / h->A.g.inv_cs)
        return fail(h, SCAB_ERR_INVALID, "scab_status_step: 2 r_max + 2 v_max dt exceeds the cell size (scab_set_grid with larger cells)");
    CK(launch_status(h->A, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_step(scab_handle *h, int policy, int nbr_mode, void *stream)
{
    int rc = scab_build_grid(h, stream);
    if (rc) return rc;
    rc = scab_neighbors(h, nbr_mode, stream);
    if (rc) return rc;
    rc = scab_policy_step(h, policy, stream);
    if (rc) return rc;
    return scab_status_step(h, stream);
}

// ------------------------------------------------------------------ Dubins v_pref provider
extern "C" int64_t scab_dubins_state_size(void) { return (int64_t)scab_dubins_state_bytes(); }

extern "C" int scab_dubins_enable(scab_handle *h, int on, void *stream)
{
    if (!h) return SCAB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (!on) { h->A.dub_states = nullptr; h->A.dub_d0 = nullptr; return SCAB_OK; }
    const size_t bytes = scab_dubins_state_bytes() * (size_t)h->capacity;
    if (!h->d_dub_states) {
        if (cudaMalloc(&h->d_dub_states, bytes) != cudaSuccess || cudaMalloc(&h->d_dub_d0, sizeof(double) * (size_t)h->capacity) != cudaSuccess) {
            if (h->d_dub_states) cudaFree(h->d_dub_states);
            h->d_dub_states = nullptr; h->d_dub_d0 = nullptr;
            (void)cudaGetLastError();
            return fail(h, SCAB_ERR_NOMEM, "scab_dubins_enable: out of device memory");
        }
    }
    // all-zero state == agent.is_use_dubins False, empty path, no neighbours seen yet
    CK(cudaMemsetAsync(h->d_dub_states, 0, bytes, (cudaStream_t)stream));
    CK(cudaMemsetAsync(h->d_dub_d0, 0, sizeof(double) * (size_t)h->capacity, (cudaStream_t)stream));
    h->A.dub_states = h->d_dub_states;
    h->A.dub_d0 = h->d_dub_d0;
    return SCAB_OK;
}

extern "C" int scab_dubins_read(scab_handle *h, void *out_host, int64_t n, void *stream)
{
    if (!h || !out_host || n < 0 || n > h->capacity) return fail(h, SCAB_ERR_INVALID, "scab_dubins_read: bad argument");
    if (!h->d_dub_states) return fail(h, SCAB_ERR_INVALID, "scab_dubins_read: the provider was never enabled");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out_host, h->d_dub_states, scab_dubins_state_bytes() * (size_t)n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_read_counters(scab_handle *h, scab_counters *out, void *stream)
{
    if (!h || !out) return fail(h, SCAB_ERR_INVALID, "scab_read_counters: null argument");
    CK(cudaSetDevice(h->device));
    unsigned long long v[8];
    CK(cudaMemcpyAsync(v, h->d_counters, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    memset(out, 0, sizeof(*out));
    out->n_active = (int64_t)v[0]; out->n_not_done = (int64_t)v[1]; out->n_collision = (int64_t)v[2]; out->n_recheck = (int64_t)v[3];
    out->n_active_total = (int64_t)v[4];
    out->n_tier2 = (int64_t)v[5]; out->n_slow = (int64_t)v[6];
    out->overflow = (int64_t)v[7];
    return SCAB_OK;
}

extern "C" int scab_reset_counters(scab_handle *h, void *stream)
{
    if (!h) return SCAB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), (cudaStream_t)stream));
    return SCAB_OK;
}

// ------------------------------------------------------------------ multi-GPU halo
extern "C" int scab_halo_pack(scab_handle *h, int64_t n_own, int axis, double lo_face, double hi_face, double width,
                              const double park[3], void *out_lo, void *out_hi, int32_t cap, void *stream)
{
    int rc = check_ready(h, "scab_halo_pack");
    if (rc) return rc;
    if (n_own < 0 || n_own > h->A.n || axis < 0 || axis > 2 || cap < 0 || cap > h->capacity || !park || (cap > 0 && (!out_lo || !out_hi)))
        return fail(h, SCAB_ERR_INVALID, "scab_halo_pack: bad argument");
    CK(launch_halo_pack(h->A, (int)n_own, axis, lo_face, hi_face, width, park, out_lo, out_hi, cap, h->d_sent_lo, h->d_sent_hi,
                        h->d_halo_cnt, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_halo_unpack(scab_handle *h, const void *recs, int32_t count, int64_t row0, void *stream)
{
    int rc = check_ready(h, "scab_halo_unpack");
    if (rc) return rc;
    if (count < 0 || row0 < 0 || row0 + count > h->capacity || (count > 0 && !recs)) return fail(h, SCAB_ERR_INVALID, "scab_halo_unpack: bad argument");
    CK(launch_halo_unpack(h->A, recs, count, (int)row0, (cudaStream_t)stream, h->pending_wait));
    h->pending_wait = P2PWait{nullptr, nullptr, nullptr};
    return SCAB_OK;
}

extern "C" int scab_halo_pack_pos(scab_handle *h, double *out_lo, double *out_hi, int32_t cap, const double park[3], void *stream)
{
    int rc = check_ready(h, "scab_halo_pack_pos");
    if (rc) return rc;
    if (cap < 0 || cap > h->capacity || !park || (cap > 0 && (!out_lo || !out_hi))) return fail(h, SCAB_ERR_INVALID, "scab_halo_pack_pos: bad argument");
    CK(launch_halo_pack_pos(h->A, h->d_sent_lo, h->d_sent_hi, h->d_halo_cnt, cap, out_lo, out_hi, park, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_halo_unpack_pos(scab_handle *h, const double *pos, int32_t count, int64_t row0, void *stream)
{
    int rc = check_ready(h, "scab_halo_unpack_pos");
    if (rc) return rc;
    if (count < 0 || row0 < 0 || row0 + count > h->capacity || (count > 0 && !pos)) return fail(h, SCAB_ERR_INVALID, "scab_halo_unpack_pos: bad argument");
    CK(launch_halo_unpack_pos(h->A, pos, count, (int)row0, (cudaStream_t)stream, h->pending_wait));
    h->pending_wait = P2PWait{nullptr, nullptr, nullptr};
    return SCAB_OK;
}

extern "C" int scab_halo_counts(scab_handle *h, int32_t out_host[2], void *stream)
{
    if (!h || !out_host) return fail(h, SCAB_ERR_INVALID, "scab_halo_counts: null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(out_host, h->d_halo_cnt, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_migrate_pack(scab_handle *h, int64_t own_cap, int axis, double lo_face, double hi_face, const double park_geom[4],
                                 int32_t nx, int32_t ny, void *out_lo, void *out_hi, int32_t cap, int32_t *free_list, int32_t *free_count,
                                 int32_t *cnt, void *stream)
{
    int rc = check_ready(h, "scab_migrate_pack");
    if (rc) return rc;
    if (own_cap < 0 || own_cap > h->A.n || axis < 0 || axis > 2 || cap <= 0 || !park_geom || nx <= 0 || ny <= 0 || !out_lo || !out_hi ||
        !free_list || !free_count || !cnt)
        return fail(h, SCAB_ERR_INVALID, "scab_migrate_pack: bad argument");
    // the vacated rows are remembered (when the scratch holds 2 * cap of them) for scab_migrate_extra_pack
    h->mig_cap = (2 * (int64_t)cap <= h->capacity) ? cap : 0;
    CK(launch_migrate_pack(h->A, (int)own_cap, axis, lo_face, hi_face, park_geom, nx, ny, out_lo, out_hi, cap, free_list, free_count, cnt,
                           h->mig_cap ? h->d_mig_sent_row : nullptr, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_migrate_unpack(scab_handle *h, const void *in_lo, const void *in_hi, int32_t cap, int use_lo, int use_hi,
                                   int32_t *free_list, int32_t *free_count, int32_t *cnt, void *stream)
{
    int rc = check_ready(h, "scab_migrate_unpack");
    if (rc) return rc;
    if (cap <= 0 || !in_lo || !in_hi || !free_list || !free_count || !cnt) return fail(h, SCAB_ERR_INVALID, "scab_migrate_unpack: bad argument");
    CK(launch_migrate_unpack(h->A, in_lo, in_hi, cap, use_lo, use_hi, free_list, free_count, cnt,
                             (2 * (int64_t)cap <= h->capacity) ? h->d_mig_recv_row : nullptr, (cudaStream_t)stream, h->pending_wait));
    h->pending_wait = P2PWait{nullptr, nullptr, nullptr};
    return SCAB_OK;
}

extern "C" int64_t scab_migrate_extra_row_doubles(scab_handle *h)
{
    if (!h) return SCAB_ERR_INVALID;
    return (int64_t)migrate_extra_row_doubles(h->A);
}

extern "C" int scab_migrate_extra_pack(scab_handle *h, double *out_lo, double *out_hi, int32_t cap, const int32_t *cnt, void *stream)
{
    int rc = check_ready(h, "scab_migrate_extra_pack");
    if (rc) return rc;
    if (cap <= 0 || !out_lo || !out_hi || !cnt || 2 * (int64_t)cap > h->capacity || h->mig_cap != cap)
        return fail(h, SCAB_ERR_INVALID, "scab_migrate_extra_pack: call scab_migrate_pack with the same cap first (2 * cap <= capacity)");
    CK(launch_migrate_extra_pack(h->A, h->d_mig_sent_row, cnt, cap, out_lo, out_hi, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_migrate_extra_unpack(scab_handle *h, const double *in_lo, const double *in_hi, int32_t cap, void *stream)
{
    int rc = check_ready(h, "scab_migrate_extra_unpack");
    if (rc) return rc;
    if (cap <= 0 || !in_lo || !in_hi || 2 * (int64_t)cap > h->capacity)
        return fail(h, SCAB_ERR_INVALID, "scab_migrate_extra_unpack: bad argument");
    CK(launch_migrate_extra_unpack(h->A, h->d_mig_recv_row, cap, in_lo, in_hi, (cudaStream_t)stream));
    return SCAB_OK;
}


// ------------------------------------------------------------------ peer-to-peer exchange (scab_p2p.cu)
static_assert(sizeof(cudaIpcMemHandle_t) == SCAB_IPC_HANDLE_BYTES, "scab200.h: SCAB_IPC_HANDLE_BYTES");

extern "C" int scab_p2p_alloc(scab_handle *h, int64_t bytes, void **dev_ptr, unsigned char *handle_out)
{
    if (!h || bytes <= 0 || !dev_ptr) return fail(h, SCAB_ERR_INVALID, "scab_p2p_alloc: bad argument");
    CK(cudaSetDevice(h->device));
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { (void)cudaGetLastError(); return fail(h, SCAB_ERR_NOMEM, "scab_p2p_alloc: out of device memory"); }
    CK(cudaMemset(p, 0, (size_t)bytes));
    h->p2p_own.push_back(p);
    *dev_ptr = p;
    if (handle_out) {     // NULL: the buffer is only shared inside this process (logical shards on one device)
        cudaIpcMemHandle_t ih;
        CK(cudaIpcGetMemHandle(&ih, p));
        memcpy(handle_out, &ih, sizeof(ih));
    }
    return SCAB_OK;
}

extern "C" int scab_p2p_open(scab_handle *h, const unsigned char *handle, void **dev_ptr)
{
    if (!h || !handle || !dev_ptr) return fail(h, SCAB_ERR_INVALID, "scab_p2p_open: bad argument");
    CK(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ih;
    memcpy(&ih, handle, sizeof(ih));
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    h->p2p_peer.push_back(p);
    *dev_ptr = p;
    return SCAB_OK;
}

extern "C" int scab_p2p_signal(scab_handle *h, void *local_seq, void *peer_flag_lo, void *peer_flag_hi, void *stream)
{
    if (!h || !local_seq) return fail(h, SCAB_ERR_INVALID, "scab_p2p_signal: bad argument");
    CK(cudaSetDevice(h->device));
    CK(launch_p2p_signal((unsigned long long *)local_seq, (unsigned long long *)peer_flag_lo, (unsigned long long *)peer_flag_hi, (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_p2p_wait(scab_handle *h, void *local_expect, const void *flag_from_lo, const void *flag_from_hi, void *stream)
{
    if (!h || !local_expect) return fail(h, SCAB_ERR_INVALID, "scab_p2p_wait: bad argument");
    CK(cudaSetDevice(h->device));
    CK(launch_p2p_wait(h->A, (unsigned long long *)local_expect, (const unsigned long long *)flag_from_lo, (const unsigned long long *)flag_from_hi,
                       (cudaStream_t)stream));
    return SCAB_OK;
}

extern "C" int scab_p2p_wait_next(scab_handle *h, const void *local_seq, const void *flag_from_lo, const void *flag_from_hi)
{
    if (!h || !local_seq) return fail(h, SCAB_ERR_INVALID, "scab_p2p_wait_next: bad argument");
    h->pending_wait = P2PWait{(const unsigned long long *)local_seq, (const unsigned long long *)flag_from_lo, (const unsigned long long *)flag_from_hi};
    return SCAB_OK;
}
