// scab_common.cuh -- shared declarations of libscab200 (sm_100a only).
//
// Numerics contract (DESIGN.md "Numerics"): every decision scalar of the reference is evaluated in
// fp64 with the reference's own decimal-rounding conventions (mamp/util.py:100-106,140-150); the
// library is compiled with -fmad=false so the only fused multiply-adds are the explicit ones that
// reproduce NumPy's 3-element dot (x0*y0 followed by two FMAs).  Bulk candidate tests run in fp32
// with an error-bounded guard band; hits inside the band are re-evaluated in fp64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/scab200.h"

#define KNBR SCAB_MAX_NEIGHBORS
#define SCAB_POSTURE_LUT_B 64
#define SCAB_PI 3.141592653589793
#define RVO_EPS 1e-5  // mamp/configs/config.py:4

// internal flag bits carried in the snapshot's vel.w (low bits == SCAB_ST_*)
// agents the policy stage skips: finished ones (mampenv.py:35) and halo ghosts owned by another GPU
#define FL_INACTIVE (SCAB_ST_AT_GOAL | SCAB_ST_COLLISION | SCAB_ST_TIMEOUT | SCAB_ST_GHOST)

struct GridGeom {
    double ox, oy, oz, inv_cs;
    int nx, ny, nz, ncell;
};

// kd-tree node of the reference (kdTree.py:14-21), REF_COMPAT mode only
struct KdNodeDev {
    int32_t begin, end, left, right;
    double mn[3], mx[3];
};

// Host only: launch configuration of ONE handle (its device's SM count, resident CTAs per SM of the persistent kernels on
// that device, the A/B switches read from the environment by scab_create).  Kernels never dereference it.
#define SCAB_OCC_SLOTS 24
struct LaunchCfg {
    int sms;
    int no_fast, lp_f64, lp_compact, lp_ctas8, knn_exact;
    int occ[SCAB_OCC_SLOTS];   // per kernel (slot constants next to each launcher); 0 = not queried on this device yet
};

// sticky bits of counters[7] (scab_counters.overflow): a result may be WRONG once one is set
#define OVF_HALO 1ull        // k_halo_pack dropped records: a face buffer was too small
#define OVF_MIG_DEFER 2ull   // k_migrate_pack could not send an agent (buffer full): it stays outside its slab
#define OVF_MIG_LOST 4ull    // k_migrate_unpack found no free row for an arrival: the agent is gone
#define OVF_MOTION 8ull      // an agent moved farther in one step than the bound the collision stage relies on
#define OVF_P2P_TIMEOUT 16ull // a peer-to-peer halo wait gave up after ~2 s (the neighbour rank never published its data)
#define OVF_RADIUS 32ull      // an agent's radius exceeds scab_set_max_radius (the 16-NN ball and the status shortcut rely on it)

// Everything a kernel needs; passed by value.
struct StepArgs {
    scab_state_ptrs st;
    int n;           // owned agents
    // scab_policy_step_rows: the policy stage for the state rows [row_begin, row_begin + row_count) only (row_count < 0: every
    // agent, in cell-sorted order).  Work item t of a per-agent kernel is then the agent of row row_begin + t.
    int row_begin, row_count;
    int m;           // obstacles
    GridGeom g;      // agent grid
    GridGeom og;     // obstacle grid (static)
    // constants
    double nbr_dist_sq, time_step, inv_time_horizon, inv_time_step, max_speed, near_goal;
    // snapshot, sorted by cell
    double4 *s_posr;    // x y z r
    float4 *s_vel;      // vx vy vz, w = status bits
    float4 *s_posf;     // fp32 copy of the position relative to the grid origin (x y z, w unused): the 16-NN tile scan's pre-filter
    int32_t *s_id;      // sorted idx -> row (local agent index)
    int32_t *s_gid;     // sorted idx -> global agent id (== row unless sharded)
    int obst_base;      // obstacle k is reported as obst_base + k in nbr_ids (n, or n_global when sharded)
    int32_t *rank;      // agent id -> sorted idx
    double4 *s_vdes;    // V_des (x y z) + zero-velocity-gate flag (w)        [k_vpref -> k_select]
    double4 *s_vpost;   // vA_post (x y z) + packed tag | n_suit << 8 (w)     [k_select -> k_integrate]
    int32_t *cell_id, *cell_local, *cell_count, *cell_start;  // [n], [n], [ncell], [ncell+1]
    int32_t *s_nbr;     // [n,16] sorted idx, or n + sorted obstacle idx
    int32_t *s_nbr_cnt; // [n]
    int32_t *s_qcoll;   // [n] collision flag raised by the neighbour query (agent.py:83-85)
    float *r16_prev;    // [capacity], by ROW: distance of the 16th neighbour found by the previous query (huge: unknown) -- the
                        // next query only looks that far plus two steps of motion, and falls back if that was not enough
    int32_t *slow_list; // [n] agents the 8-lane fast selection kernel hands to the general kernel
    int32_t *slow_count;
    int32_t *slow_list2, *slow_count2;   // agents the second (every-candidate) fast pass hands to the general kernel
    double max_radius;  // largest agent radius (0: unknown -> the status stage always scans the cell list)
    double move_bound;  // no agent moves farther than this in one step (scab_set_motion_bound; kinematic_update checks it)
    LaunchCfg *cfg;     // host pointer (the owning handle's launch configuration)
    int last_nbr_mode;  // mode of the neighbour lists currently in s_nbr
    float tier_delta;   // lazy tiers: tier 1 = candidates within this distance of the one closest to v_pref
    int full_search;    // SCAB_OPT_FULL_SEARCH: examine every candidate (exact n_suit) instead of the lazy tiers
    // obstacles (sorted by obstacle cell) + reference kd-trees
    const double4 *o_posr;
    const int32_t *o_id, *o_cell_start;
    const KdNodeDev *kd_agent, *kd_obst;
    const int32_t *kd_agent_perm, *kd_obst_perm;
    const double4 *o_posr_by_id;  // unsorted (id order), for the kd-tree replay
    // candidate direction tables (host libm, uploaded once): unit Fibonacci-sphere directions
    const double *dir256, *dir128;   // [N,3] f64
    const float4 *dir256f, *dir128f; // [N] f32 (x y z 0)
    // posture lookup tables (scab_api.cu make_posture_lut): per bin of an octahedral map of vA's direction, 8 words of
    // directions that pass angle(vA, u) <= pi/4 for EVERY vA of the bin and 8 words of directions that may; lane l's word
    // holds direction 8 i + l in bit i
    const uint32_t *plut256, *plut128;   // [SCAB_POSTURE_LUT_B^2][16]
    unsigned long long *counters;    // scab_counters
    // Dubins v_pref provider (scab_dubins.cu); both NULL while it is disabled
    void *dub_states;                // [capacity] DubState, indexed by row
    double *dub_d0;                  // [capacity] this step's agent.neighbors[0][1] (-1: empty list), indexed by row
};

// ------------------------------------------------------------------ cache hints
#ifdef __CUDACC__
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else     // host builds of the device source (tests/hostsim)
static inline void prefetch_l2(const void *) {}
static inline void prefetch_l1(const void *) {}
#endif

// ------------------------------------------------------------------ rounding conventions
// x / 1e5, correctly rounded, as one multiply and two FMAs (Markstein's division by a constant with y = RN(1/d)) instead
// of the ~25-instruction IEEE division: every decimal rounding of the reference ends in this division.  Bit-identical to
// x / 1e5 for every integer-valued |x| <= 2^33 (exhaustive check) and on 4e8 random doubles (tools/div1e5_check.c); zero
// stays zero.  Not for infinities / NaN (callers that can see them test first).
__device__ __forceinline__ double div1e5(double x)
{
    const double q = __dmul_rn(x, 1e-5);
    const double r = __fma_rn(-q, 1e5, x);
    return __fma_rn(r, 1e-5, q);
}
__device__ __forceinline__ double np_round5(double x)
{  // round(np.float64, 5): NumPy scalar __round__ -> rint(x*1e5)/1e5 (util.py:100 l3normsq)
    return div1e5(rint(__dmul_rn(x, 1e5)));
}
__device__ __forceinline__ double py_round5_units(double x)
{  // CPython float.__round__(x,5) numerator: round-half-even on the EXACT value of x*1e5
    double p = __dmul_rn(x, 1e5);
    double e = __fma_rn(x, 1e5, -p);
    double r = rint(p);
    if (fabs(p - r) == 0.5 && e != 0.0) r = (e > 0.0) ? ceil(p) : floor(p);
    return r;
}
__device__ __forceinline__ double py_round5(double x)
{
    if (!isfinite(x)) return x;
    return div1e5(py_round5_units(x));
}
__device__ __forceinline__ double trunc5(double x)
{  // int(x*1e5)/1e5 ; python ints have no -0
    return div1e5(__dadd_rn(trunc(__dmul_rn(x, 1e5)), 0.0));
}
__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2)
{  // np.dot on 3-vectors (OpenBLAS ddot tail): product then two FMAs
    return __fma_rn(a2, b2, __fma_rn(a1, b1, __dmul_rn(a0, b0)));
}
__device__ __forceinline__ double sumsq3(double d0, double d1, double d2)
{  // python: d0**2 + d1**2 + d2**2, each op rounded
    return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}
__device__ __forceinline__ double l3norm_d(double ax, double ay, double az, double bx, double by, double bz)
{  // util.py:103-105
    return py_round5(__dsqrt_rn(sumsq3(__dsub_rn(ax, bx), __dsub_rn(ay, by), __dsub_rn(az, bz))));
}
__device__ __forceinline__ double l3normsq_d(double ax, double ay, double az, double bx, double by, double bz)
{  // util.py:99-100
    return np_round5(sumsq3(__dsub_rn(ax, bx), __dsub_rn(ay, by), __dsub_rn(az, bz)));
}
__device__ __forceinline__ double sumsq_f32(float x, float y, float z)
{  // l3norm(vA,[0,0,0]) with float32 vA: squares and sums stay float32
    return (double)__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float norm3f_np(float x, float y, float z)
{  // np.linalg.norm(float32[3]): f32 products accumulated in double (OpenBLAS sdot), f32 sqrt
    double d = (double)__fmul_rn(x, x);
    d = __dadd_rn(d, (double)__fmul_rn(y, y));
    d = __dadd_rn(d, (double)__fmul_rn(z, z));
    return __fsqrt_rn((float)d);
}
__device__ __forceinline__ double pi_2_pi(double angle)
{  // util.py:108-109 with NumPy's floored remainder
    double a = __dadd_rn(angle, SCAB_PI);
    const double b = 2 * SCAB_PI;
    double m = fmod(a, b);
    if (m != 0.0) {
        if (m < 0) m = __dadd_rn(m, b);
    } else {
        m = 0.0;
    }
    return __dsub_rn(m, SCAB_PI);
}

__device__ __forceinline__ int cell_of(const GridGeom &g, double x, double y, double z, int &cx, int &cy, int &cz)
{
    cx = (int)floor((x - g.ox) * g.inv_cs);
    cy = (int)floor((y - g.oy) * g.inv_cs);
    cz = (int)floor((z - g.oz) * g.inv_cs);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    cz = min(max(cz, 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}

// ------------------------------------------------------------------ shared device routines
struct Action7 { float a[7]; };

// util.py:44-55 (use_distance=false) / orca3dPolicyOfficial.py:331-342 (use_distance=true)
__device__ __forceinline__ Action7 cartesian2spherical(double vx, double vy, double vz, double yaw, double pitch,
                                                       bool use_distance)
{
    double nrm = __dsqrt_rn(sumsq3(vx, vy, vz));
    double speed = use_distance ? py_round5(__dadd_rn(nrm, 1e-5)) : py_round5(nrm);
    double alpha = 0.0, beta = 0.0;
    if (!(speed < 0.001)) {
        alpha = __dsub_rn(atan2(vy, vx), yaw);
        beta = __dsub_rn(atan2(vz, __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)))), pitch);
    }
    Action7 r;
    r.a[0] = (float)vx; r.a[1] = (float)vy; r.a[2] = (float)vz; r.a[3] = (float)speed;
    r.a[4] = (float)alpha; r.a[5] = (float)beta; r.a[6] = 0.f;
    return r;
}

__device__ __forceinline__ double distance_d(double ax, double ay, double az, double bx, double by, double bz)
{  // util.py:140-142 distance(): round(norm + 1e-5, 5)
    return py_round5(__dadd_rn(__dsqrt_rn(sumsq3(__dsub_rn(ax, bx), __dsub_rn(ay, by), __dsub_rn(az, bz))), 1e-5));
}

// get_trajectory (rvo3dPolicy.py:71-86, srvo3dPolicy.py:71-86 with l3norm; orca3dPolicy.py:298-312,
// orca3dPolicyOfficial.py:302-316 with distance): the point compute_v_pref aims at.  Pops agent.path from its end.
__device__ __forceinline__ void waypoint_target(const scab_state_ptrs &st, int id, double px, double py, double pz, bool use_distance,
                                                double *target, bool commit = true)
{   // commit == false: compute the target only, leave agent.path / now_goal as they are (a later call commits)
    const double *g = &st.goal[3 * id];
    target[0] = g[0]; target[1] = g[1]; target[2] = g[2];
    if (st.waypoints == nullptr) return;
    double *ng = &st.now_goal[3 * id];
    int left = st.waypoint_left[id];
    if (left <= 0) { if (commit) { ng[0] = g[0]; ng[1] = g[1]; ng[2] = g[2]; } return; }     // path exhausted: now_goal = goal
    const double *wp = st.waypoints + (size_t)id * (size_t)st.waypoint_stride * 3;
    double nx = ng[0], ny = ng[1], nz = ng[2];
    if (isnan(nx)) { left--; nx = wp[3 * left]; ny = wp[3 * left + 1]; nz = wp[3 * left + 2]; }   // first
    const double dis = use_distance ? distance_d(px, py, pz, nx, ny, nz) : l3norm_d(px, py, pz, nx, ny, nz);
    const double d_ng_goal = use_distance ? distance_d(nx, ny, nz, g[0], g[1], g[2]) : l3norm_d(nx, ny, nz, g[0], g[1], g[2]);
    const double d_pos_goal = use_distance ? distance_d(px, py, pz, g[0], g[1], g[2]) : l3norm_d(px, py, pz, g[0], g[1], g[2]);
    if ((dis <= __dmul_rn(1.0, st.radius[id]) || d_ng_goal >= d_pos_goal) && left > 0) {   // update_now_goal_dist = 1.0
        left--; nx = wp[3 * left]; ny = wp[3 * left + 1]; nz = wp[3 * left + 2];
    }
    if (commit) {
        st.waypoint_left[id] = left;
        ng[0] = nx; ng[1] = ny; ng[2] = nz;
    }
    target[0] = nx; target[1] = ny; target[2] = nz;
}

// rvo3dPolicy.py:182-196 (use_distance=false) / orca3dPolicy*.py:345-359 (use_distance=true): aims at `t` (the goal or
// the current waypoint); the at-goal zero tests the GLOBAL goal `g`
__device__ __forceinline__ void v_pref_straight(const double *g, const double *t, double px, double py, double pz, double pref,
                                                bool use_distance, double *v_store, double *v_des)
{
    double dx = __dsub_rn(t[0], px), dy = __dsub_rn(t[1], py), dz = __dsub_rn(t[2], pz);
    double nr = __dsqrt_rn(sumsq3(dx, dy, dz));
    double nrm = use_distance ? py_round5(__dadd_rn(nr, 1e-5)) : py_round5(nr);
    nrm = trunc5(nrm);
    v_store[0] = __ddiv_rn(__dmul_rn(dx, pref), nrm);
    v_store[1] = __ddiv_rn(__dmul_rn(dy, pref), nrm);
    v_store[2] = __ddiv_rn(__dmul_rn(dz, pref), nrm);
    const double ng = __dsqrt_rn(sumsq3(__dsub_rn(g[0], px), __dsub_rn(g[1], py), __dsub_rn(g[2], pz)));
    if (py_round5(ng) < 0.2) { v_store[0] = 0.0; v_store[1] = 0.0; v_store[2] = 0.0; }  // reached(), util.py:23
    v_des[0] = trunc5(v_store[0]); v_des[1] = trunc5(v_store[1]); v_des[2] = trunc5(v_store[2]);
}

// mampenv.py:83-105 update_velocitie; writes the agent's state in place and returns the new position
__device__ __forceinline__ void kinematic_update(const StepArgs &A, int id, const Action7 &act, double px, double py,
                                                 double pz, double *npx, double *npy, double *npz)
{
    double speed = (double)act.a[3];
    double al = pi_2_pi(__dadd_rn(A.st.heading[3 * id + 0], (double)act.a[4]));
    double be = pi_2_pi(__dadd_rn(A.st.heading[3 * id + 1], (double)act.a[5]));
    double ga = pi_2_pi(__dadd_rn(A.st.heading[3 * id + 2], (double)act.a[6]));
    double sb, cb, sa, ca;
    sincos(be, &sb, &cb);
    sincos(al, &sa, &ca);
    double t = __dmul_rn(speed, cb);
    double dx = __dmul_rn(__dmul_rn(t, ca), A.time_step);
    double dy = __dmul_rn(__dmul_rn(t, sa), A.time_step);
    double dz = __dmul_rn(__dmul_rn(speed, sb), A.time_step);
    double len = __dsqrt_rn(sumsq3(dx, dy, dz));
    if (len > A.move_bound) atomicOr(&A.counters[7], OVF_MOTION);   // the collision stage's reach assumes the bound: be loud
    A.st.total_dist[id] = __dadd_rn(A.st.total_dist[id], len);
    *npx = __dadd_rn(px, dx); *npy = __dadd_rn(py, dy); *npz = __dadd_rn(pz, dz);
    A.st.pos[3 * id + 0] = *npx; A.st.pos[3 * id + 1] = *npy; A.st.pos[3 * id + 2] = *npz;
    A.st.heading[3 * id + 0] = al; A.st.heading[3 * id + 1] = be; A.st.heading[3 * id + 2] = ga;
    A.st.vel[3 * id + 0] = act.a[0]; A.st.vel[3 * id + 1] = act.a[1]; A.st.vel[3 * id + 2] = act.a[2];
#pragma unroll
    for (int k = 0; k < 7; k++) A.st.action[7 * id + k] = act.a[k];
}

// the spin of a folded peer-to-peer wait (see P2PWait); call with every thread of the block, before any early return
__device__ __forceinline__ void p2p_block_wait(const unsigned long long *seq, const unsigned long long *flag_lo,
                                               const unsigned long long *flag_hi, unsigned long long *counters)
{
    if (seq == nullptr) return;          // (uniform across the block)
    if (threadIdx.x == 0) {
        const unsigned long long e = *seq;
        const long long t0 = clock64();
        for (;;) {
            const bool lo_ok = !flag_lo || *(volatile const unsigned long long *)flag_lo >= e;
            const bool hi_ok = !flag_hi || *(volatile const unsigned long long *)flag_hi >= e;
            if (lo_ok && hi_ok) break;
            if (clock64() - t0 > 8000000000ll) { atomicOr(&counters[7], OVF_P2P_TIMEOUT); break; }
            __nanosleep(40);
        }
        __threadfence_system();          // acquire: the rest of the block reads what the peers wrote
    }
    __syncthreads();
}

// number of per-agent work items of a policy kernel and the sorted index of item t (see StepArgs::row_count)
__device__ __forceinline__ int policy_items(const StepArgs &A) { return A.row_count < 0 ? A.n : A.row_count; }
__device__ __forceinline__ int policy_item(const StepArgs &A, int t) { return A.row_count < 0 ? t : A.rank[A.row_begin + t]; }

// kernels' host-side launchers (one translation unit per stage)
cudaError_t launch_grid_build(const StepArgs &A, cudaStream_t s);
cudaError_t launch_neighbors(const StepArgs &A, int mode, cudaStream_t s);
cudaError_t launch_policy_sampled(const StepArgs &A, int policy, cudaStream_t s);
cudaError_t launch_policy_lp(const StepArgs &A, cudaStream_t s);
cudaError_t launch_status(const StepArgs &A, cudaStream_t s);
cudaError_t launch_vpref_dubins(const StepArgs &A, cudaStream_t s);
size_t scab_dubins_state_bytes();
cudaError_t launch_halo_pack(const StepArgs &A, int n_own, int axis, double lo_face, double hi_face, double width,
                             const double park[3], void *out_lo, void *out_hi, int cap, int32_t *sent_lo, int32_t *sent_hi,
                             int32_t *cnt, cudaStream_t s);
// a peer-to-peer wait folded into the next unpack kernel (scab_p2p_wait_next): every block's first thread spins until both
// neighbours' flags reach this rank's own sequence number of that exchange, then the block proceeds
struct P2PWait { const unsigned long long *seq, *flag_lo, *flag_hi; };
cudaError_t launch_halo_unpack(const StepArgs &A, const void *recs, int count, int row0, cudaStream_t s, P2PWait w = P2PWait{nullptr, nullptr, nullptr});
cudaError_t launch_halo_pack_pos(const StepArgs &A, const int32_t *sent_lo, const int32_t *sent_hi, const int32_t *cnt, int cap,
                                 double *out_lo, double *out_hi, const double park[3], cudaStream_t s);
cudaError_t launch_halo_unpack_pos(const StepArgs &A, const double *p, int count, int row0, cudaStream_t s, P2PWait w = P2PWait{nullptr, nullptr, nullptr});
cudaError_t launch_migrate_pack(const StepArgs &A, int own_cap, int axis, double lo_face, double hi_face, const double geom[4], int nx, int ny,
                                void *out_lo, void *out_hi, int cap, int32_t *free_list, int32_t *free_count, int32_t *cnt, int32_t *sent_row,
                                cudaStream_t s);
cudaError_t launch_migrate_unpack(const StepArgs &A, const void *in_lo, const void *in_hi, int cap, int use_lo, int use_hi,
                                  int32_t *free_list, int32_t *free_count, int32_t *cnt, int32_t *recv_row, cudaStream_t s,
                                  P2PWait w = P2PWait{nullptr, nullptr, nullptr});
int migrate_extra_row_doubles(const StepArgs &A);
cudaError_t launch_p2p_signal(unsigned long long *local_seq, unsigned long long *peer_flag_lo, unsigned long long *peer_flag_hi, cudaStream_t s);
cudaError_t launch_p2p_wait(const StepArgs &A, unsigned long long *local_expect, const unsigned long long *flag_from_lo,
                            const unsigned long long *flag_from_hi, cudaStream_t s);
cudaError_t launch_migrate_extra_pack(const StepArgs &A, const int32_t *sent_row, const int32_t *cnt, int cap, double *out_lo, double *out_hi,
                                      cudaStream_t s);
cudaError_t launch_migrate_extra_unpack(const StepArgs &A, const int32_t *recv_row, int cap, const double *in_lo, const double *in_hi,
                                        cudaStream_t s);
