// scab_halo.cu -- multi-GPU spatial sharding: device-side pack / unpack of the halo records that are
// exchanged with the two slab neighbours over NCCL (the reference is single-process: no counterpart).
//
// The exchange is FIXED SIZE and fully asynchronous (no host round trip for counts): each face buffer
// holds `cap` 64-byte records; unused slots carry a "parked" ghost placed in a reserved, otherwise empty
// layer of the uniform grid so that it can never be anybody's neighbour.  Two exchanges per step:
//   pre   [x y z r | vx vy vz status | gid]  of every owned agent within `width` of a slab face
//         -> become read-only ghost rows (SCAB_ST_GHOST) on the neighbour GPU;
//   post  the same agents' positions after the kinematic update (the sequential collision-check order of
//         mampenv.py:42-46 needs both the old and the new position of a partner).
#include "scab_common.cuh"

struct HaloRec {       // 64 B
    double x, y, z, r;
    float vx, vy, vz;
    int32_t status;
    int32_t gid;
    int32_t pad[3];
};

__global__ void __launch_bounds__(256) k_halo_fill(HaloRec *lo, HaloRec *hi, int cap, double px, double py, double pz, int32_t *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { cnt[0] = 0; cnt[1] = 0; }      // the per-face counters of the pack kernel that follows
    if (i >= cap) return;
    HaloRec r;
    r.x = px; r.y = py; r.z = pz; r.r = 0.0;
    r.vx = r.vy = r.vz = 0.f;
    r.status = SCAB_ST_GHOST | SCAB_ST_AT_GOAL | SCAB_ST_RUN_DONE;
    r.gid = 0x7fffffff;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    lo[i] = r;
    hi[i] = r;
}

__global__ void __launch_bounds__(256) k_halo_pack(StepArgs A, int n_own, int axis, double lo_face, double hi_face, double width,
                                                   HaloRec *out_lo, HaloRec *out_hi, int cap, int32_t *sent_lo, int32_t *sent_hi,
                                                   int32_t *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    if (A.st.status[i] & SCAB_ST_GHOST) return;   // unused owned slots are kept as parked ghosts (static row count)
    const double c = A.st.pos[3 * i + axis];
    const bool to_lo = c < lo_face + width, to_hi = c >= hi_face - width;
    if (!to_lo && !to_hi) return;
    HaloRec r;
    r.x = A.st.pos[3 * i]; r.y = A.st.pos[3 * i + 1]; r.z = A.st.pos[3 * i + 2]; r.r = A.st.radius[i];
    r.vx = A.st.vel[3 * i]; r.vy = A.st.vel[3 * i + 1]; r.vz = A.st.vel[3 * i + 2];
    r.status = (A.st.status[i] & 15) | SCAB_ST_GHOST;
    r.gid = A.st.gid ? A.st.gid[i] : i;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    if (to_lo) {
        const int slot = atomicAdd(&cnt[0], 1);
        if (slot < cap) { out_lo[slot] = r; sent_lo[slot] = i; }
        else atomicOr(&A.counters[7], OVF_HALO);
    }
    if (to_hi) {
        const int slot = atomicAdd(&cnt[1], 1);
        if (slot < cap) { out_hi[slot] = r; sent_hi[slot] = i; }
        else atomicOr(&A.counters[7], OVF_HALO);
    }
}

__global__ void __launch_bounds__(256) k_halo_unpack(StepArgs A, const HaloRec *recs, int count, int row0, P2PWait w)
{
    p2p_block_wait(w.seq, w.flag_lo, w.flag_hi, A.counters);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const HaloRec r = recs[k];
    const int i = row0 + k;
    A.st.pos[3 * i] = r.x; A.st.pos[3 * i + 1] = r.y; A.st.pos[3 * i + 2] = r.z;
    ((double *)A.st.radius)[i] = r.r;
    A.st.vel[3 * i] = r.vx; A.st.vel[3 * i + 1] = r.vy; A.st.vel[3 * i + 2] = r.vz;
    A.st.status[i] = r.status;
    if (A.st.gid) ((int32_t *)A.st.gid)[i] = r.gid;
}

// post exchange: positions of the agents sent in the pre exchange, same slots
__global__ void __launch_bounds__(256) k_halo_pack_pos(StepArgs A, const int32_t *sent_lo, const int32_t *sent_hi, const int32_t *cnt,
                                                       int cap, double *out_lo, double *out_hi, double px, double py, double pz)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cap) return;
    const int side = k / cap, slot = k - side * cap;
    const int c = min(cnt[side], cap);
    double *o = (side ? out_hi : out_lo) + 3 * (size_t)slot;
    if (slot < c) {
        const int i = (side ? sent_hi : sent_lo)[slot];
        o[0] = A.st.pos[3 * i]; o[1] = A.st.pos[3 * i + 1]; o[2] = A.st.pos[3 * i + 2];
    } else {
        o[0] = px; o[1] = py; o[2] = pz;
    }
}

__global__ void __launch_bounds__(256) k_halo_unpack_pos(StepArgs A, const double *p, int count, int row0, P2PWait w)
{
    p2p_block_wait(w.seq, w.flag_lo, w.flag_hi, A.counters);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = row0 + k;
    A.st.pos[3 * i] = p[3 * k]; A.st.pos[3 * i + 1] = p[3 * k + 1]; A.st.pos[3 * i + 2] = p[3 * k + 2];
}

cudaError_t launch_halo_pack(const StepArgs &A, int n_own, int axis, double lo_face, double hi_face, double width,
                             const double park[3], void *out_lo, void *out_hi, int cap, int32_t *sent_lo, int32_t *sent_hi,
                             int32_t *cnt, cudaStream_t s)
{
    if (cap > 0) k_halo_fill<<<(cap + 255) / 256, 256, 0, s>>>((HaloRec *)out_lo, (HaloRec *)out_hi, cap, park[0], park[1], park[2], cnt);
    else {
        cudaError_t e = cudaMemsetAsync(cnt, 0, 2 * sizeof(int32_t), s);
        if (e != cudaSuccess) return e;
    }
    if (n_own > 0)
        k_halo_pack<<<(n_own + 255) / 256, 256, 0, s>>>(A, n_own, axis, lo_face, hi_face, width, (HaloRec *)out_lo, (HaloRec *)out_hi,
                                                       cap, sent_lo, sent_hi, cnt);
    return cudaGetLastError();
}

cudaError_t launch_halo_unpack(const StepArgs &A, const void *recs, int count, int row0, cudaStream_t s, P2PWait w)
{
    if (count > 0) k_halo_unpack<<<(count + 255) / 256, 256, 0, s>>>(A, (const HaloRec *)recs, count, row0, w);
    return cudaGetLastError();
}

cudaError_t launch_halo_pack_pos(const StepArgs &A, const int32_t *sent_lo, const int32_t *sent_hi, const int32_t *cnt, int cap,
                                 double *out_lo, double *out_hi, const double park[3], cudaStream_t s)
{
    if (cap > 0)
        k_halo_pack_pos<<<(2 * cap + 255) / 256, 256, 0, s>>>(A, sent_lo, sent_hi, cnt, cap, out_lo, out_hi, park[0], park[1], park[2]);
    return cudaGetLastError();
}

cudaError_t launch_halo_unpack_pos(const StepArgs &A, const double *p, int count, int row0, cudaStream_t s, P2PWait w)
{
    if (count > 0) k_halo_unpack_pos<<<(count + 255) / 256, 256, 0, s>>>(A, p, count, row0, w);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ device-side migration
// Agents that left the slab move to the neighbour GPU as 128-byte full-state rows through fixed-size
// buffers (row 0 of each buffer is a header carrying the count), entirely on the device: the vacated
// state row becomes a parked ghost and its index is pushed on a free list; arrivals pop free rows.
struct MigRow {        // 128 B
    double pos[3];
    double heading[3];
    double goal[3];
    double radius, pref_speed, total_dist, max_run_dist;
    float vel[3];
    int32_t status, step_num, gid;
};
static_assert(sizeof(MigRow) == 128, "MigRow must be 128 bytes");

struct ParkGeom { double lox, loy, cell, z; int nx, ny; };

__device__ __forceinline__ void park_row(const StepArgs &A, int i, const ParkGeom &P)
{
    A.st.pos[3 * i] = P.lox + (((long long)i * 7919) % P.nx + 0.5) * P.cell;
    A.st.pos[3 * i + 1] = P.loy + (((long long)i * 104729) % P.ny + 0.5) * P.cell;
    A.st.pos[3 * i + 2] = P.z;
    A.st.vel[3 * i] = A.st.vel[3 * i + 1] = A.st.vel[3 * i + 2] = 0.f;
    ((double *)A.st.radius)[i] = 0.0;
    A.st.status[i] = SCAB_ST_GHOST | SCAB_ST_AT_GOAL | SCAB_ST_RUN_DONE;
    if (A.st.gid) ((int32_t *)A.st.gid)[i] = 0x7fffffff;
}

__global__ void __launch_bounds__(256) k_migrate_pack(StepArgs A, int own_cap, int axis, double lo_face, double hi_face, ParkGeom P,
                                                      MigRow *out_lo, MigRow *out_hi, int cap, int32_t *free_list, int32_t *free_count,
                                                      int32_t *cnt, int32_t *sent_row)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= own_cap) return;
    const int st = A.st.status[i];
    if (st & SCAB_ST_GHOST) return;
    const double c = A.st.pos[3 * i + axis];
    const int side = (c < lo_face) ? 0 : (c >= hi_face ? 1 : -1);
    if (side < 0) return;
    const int slot = atomicAdd(&cnt[side], 1);
    if (slot >= cap) { atomicAdd(&cnt[2], 1); atomicOr(&A.counters[7], OVF_MIG_DEFER); return; }   // overflow: the agent stays (counted, sticky flag)
    MigRow r;
    for (int d = 0; d < 3; d++) {
        r.pos[d] = A.st.pos[3 * i + d]; r.heading[d] = A.st.heading[3 * i + d]; r.goal[d] = A.st.goal[3 * i + d];
        r.vel[d] = A.st.vel[3 * i + d];
    }
    r.radius = A.st.radius[i]; r.pref_speed = A.st.pref_speed[i]; r.total_dist = A.st.total_dist[i];
    r.max_run_dist = A.st.max_run_dist[i];
    r.status = st; r.step_num = A.st.step_num[i]; r.gid = A.st.gid ? A.st.gid[i] : i;
    (side ? out_hi : out_lo)[1 + slot] = r;
    if (sent_row) sent_row[side * cap + slot] = i;          // for k_migrate_extra_pack
    park_row(A, i, P);
    free_list[atomicAdd(free_count, 1)] = i;
}

__global__ void k_migrate_header(MigRow *out_lo, MigRow *out_hi, int cap, const int32_t *cnt)
{
    if (threadIdx.x == 0) { out_lo[0].gid = min(cnt[0], cap); out_hi[0].gid = min(cnt[1], cap); }
}

__global__ void __launch_bounds__(256) k_migrate_unpack(StepArgs A, const MigRow *in_lo, const MigRow *in_hi, int cap, int use_lo, int use_hi,
                                                        int32_t *free_list, int32_t *free_count, int32_t *cnt, int32_t *recv_row, P2PWait w)
{
    p2p_block_wait(w.seq, w.flag_lo, w.flag_hi, A.counters);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cap) return;
    if (recv_row) recv_row[k] = -1;                          // row given to arrival k (for k_migrate_extra_unpack); -1: none
    const int side = k / cap, slot = k - side * cap;
    if (side == 0 ? !use_lo : !use_hi) return;
    const MigRow *in = side ? in_hi : in_lo;
    const int count = min(max(in[0].gid, 0), cap);
    if (slot >= count) return;
    const int fi = atomicSub(free_count, 1) - 1;
    if (fi < 0) { atomicAdd(free_count, 1); atomicAdd(&cnt[3], 1); atomicOr(&A.counters[7], OVF_MIG_LOST); return; }   // no free row: lost (sticky flag; scab_read_counters reports it)
    const int i = free_list[fi];
    if (recv_row) recv_row[k] = i;
    const MigRow r = in[1 + slot];
    for (int d = 0; d < 3; d++) {
        A.st.pos[3 * i + d] = r.pos[d]; A.st.heading[3 * i + d] = r.heading[d]; ((double *)A.st.goal)[3 * i + d] = r.goal[d];
        A.st.vel[3 * i + d] = r.vel[d];
    }
    ((double *)A.st.radius)[i] = r.radius; ((double *)A.st.pref_speed)[i] = r.pref_speed; A.st.total_dist[i] = r.total_dist;
    ((double *)A.st.max_run_dist)[i] = r.max_run_dist;
    A.st.status[i] = r.status; A.st.step_num[i] = r.step_num;
    if (A.st.gid) ((int32_t *)A.st.gid)[i] = r.gid;
}

cudaError_t launch_migrate_pack(const StepArgs &A, int own_cap, int axis, double lo_face, double hi_face, const double geom[4], int nx, int ny,
                                void *out_lo, void *out_hi, int cap, int32_t *free_list, int32_t *free_count, int32_t *cnt, int32_t *sent_row,
                                cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(cnt, 0, 2 * sizeof(int32_t), s);    // cnt[2], cnt[3] (overflow / lost) accumulate
    if (e != cudaSuccess) return e;
    ParkGeom P{geom[0], geom[1], geom[2], geom[3], nx, ny};
    if (own_cap > 0)
        k_migrate_pack<<<(own_cap + 255) / 256, 256, 0, s>>>(A, own_cap, axis, lo_face, hi_face, P, (MigRow *)out_lo, (MigRow *)out_hi, cap,
                                                            free_list, free_count, cnt, sent_row);
    k_migrate_header<<<1, 32, 0, s>>>((MigRow *)out_lo, (MigRow *)out_hi, cap, cnt);
    return cudaGetLastError();
}

cudaError_t launch_migrate_unpack(const StepArgs &A, const void *in_lo, const void *in_hi, int cap, int use_lo, int use_hi,
                                  int32_t *free_list, int32_t *free_count, int32_t *cnt, int32_t *recv_row, cudaStream_t s, P2PWait w)
{
    if (cap > 0)
        k_migrate_unpack<<<(2 * cap + 255) / 256, 256, 0, s>>>(A, (const MigRow *)in_lo, (const MigRow *)in_hi, cap, use_lo, use_hi, free_list,
                                                              free_count, cnt, recv_row, w);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ per-agent state outside the 128-byte row
// The Dubins tracker (library-owned DubState + this step's nearest-neighbour key, the caller's goal_heading)
// and the waypoint follower (agent.path, its length, policy.now_goal) are row-indexed too, so they travel with
// a migrating agent: one extra row of `row_doubles` doubles per migrated agent, same slot as its MigRow.
//   [0, dub)  DubState   [dub] dub_d0   [+1..+3] goal_heading   [+4..+6] now_goal   [+7] waypoint_left   [+8 ...) waypoints
__global__ void __launch_bounds__(128) k_migrate_extra_pack(StepArgs A, const int32_t *sent_row, const int32_t *cnt, int cap, double *out_lo,
                                                            double *out_hi, int row_doubles, int dub_doubles)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cap) return;
    const int side = k / cap, slot = k - side * cap;
    if (slot >= min(cnt[side], cap)) return;
    const int i = sent_row[k];
    double *o = (side ? out_hi : out_lo) + (size_t)slot * row_doubles;
    int w = 0;
    if (A.dub_states) {
        const double *src = (const double *)((const char *)A.dub_states + (size_t)i * dub_doubles * sizeof(double));
        for (int q = 0; q < dub_doubles; q++) o[q] = src[q];
    } else for (int q = 0; q < dub_doubles; q++) o[q] = 0.0;
    w = dub_doubles;
    o[w++] = A.dub_d0 ? A.dub_d0[i] : 0.0;
    for (int d = 0; d < 3; d++) o[w++] = A.st.goal_heading ? A.st.goal_heading[3 * i + d] : 0.0;
    for (int d = 0; d < 3; d++) o[w++] = A.st.now_goal ? A.st.now_goal[3 * i + d] : 0.0;
    o[w++] = A.st.waypoint_left ? (double)A.st.waypoint_left[i] : 0.0;
    if (A.st.waypoints) {
        const int nq = 3 * (int)A.st.waypoint_stride;
        const double *src = A.st.waypoints + (size_t)i * nq;
        for (int q = 0; q < nq; q++) o[w + q] = src[q];
    }
}

__global__ void __launch_bounds__(128) k_migrate_extra_unpack(StepArgs A, const int32_t *recv_row, int cap, const double *in_lo,
                                                              const double *in_hi, int row_doubles, int dub_doubles)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cap) return;
    const int i = recv_row[k];
    if (i < 0) return;
    const int side = k / cap, slot = k - side * cap;
    const double *o = (side ? in_hi : in_lo) + (size_t)slot * row_doubles;
    int w = dub_doubles;
    if (A.dub_states) {
        double *dst = (double *)((char *)A.dub_states + (size_t)i * dub_doubles * sizeof(double));
        for (int q = 0; q < dub_doubles; q++) dst[q] = o[q];
    }
    if (A.dub_d0) A.dub_d0[i] = o[w];
    w++;
    for (int d = 0; d < 3; d++, w++) if (A.st.goal_heading) ((double *)A.st.goal_heading)[3 * i + d] = o[w];
    for (int d = 0; d < 3; d++, w++) if (A.st.now_goal) A.st.now_goal[3 * i + d] = o[w];
    if (A.st.waypoint_left) A.st.waypoint_left[i] = (int32_t)o[w];
    w++;
    if (A.st.waypoints) {
        const int nq = 3 * (int)A.st.waypoint_stride;
        double *dst = (double *)A.st.waypoints + (size_t)i * nq;
        for (int q = 0; q < nq; q++) dst[q] = o[w + q];
    }
}

int migrate_extra_row_doubles(const StepArgs &A)
{
    return (int)(scab_dubins_state_bytes() / sizeof(double)) + 8 + (A.st.waypoints ? 3 * (int)A.st.waypoint_stride : 0);
}

cudaError_t launch_migrate_extra_pack(const StepArgs &A, const int32_t *sent_row, const int32_t *cnt, int cap, double *out_lo, double *out_hi,
                                      cudaStream_t s)
{
    const int dub = (int)(scab_dubins_state_bytes() / sizeof(double));
    if (cap > 0) k_migrate_extra_pack<<<(2 * cap + 127) / 128, 128, 0, s>>>(A, sent_row, cnt, cap, out_lo, out_hi, migrate_extra_row_doubles(A), dub);
    return cudaGetLastError();
}

cudaError_t launch_migrate_extra_unpack(const StepArgs &A, const int32_t *recv_row, int cap, const double *in_lo, const double *in_hi,
                                        cudaStream_t s)
{
    const int dub = (int)(scab_dubins_state_bytes() / sizeof(double));
    if (cap > 0) k_migrate_extra_unpack<<<(2 * cap + 127) / 128, 128, 0, s>>>(A, recv_row, cap, in_lo, in_hi, migrate_extra_row_doubles(A), dub);
    return cudaGetLastError();
}
