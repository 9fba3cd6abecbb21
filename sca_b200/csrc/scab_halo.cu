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

__global__ void __launch_bounds__(256) k_halo_fill(HaloRec *lo, HaloRec *hi, int cap, double px, double py, double pz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    }
    if (to_hi) {
        const int slot = atomicAdd(&cnt[1], 1);
        if (slot < cap) { out_hi[slot] = r; sent_hi[slot] = i; }
    }
}

__global__ void __launch_bounds__(256) k_halo_unpack(StepArgs A, const HaloRec *recs, int count, int row0)
{
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

__global__ void __launch_bounds__(256) k_halo_unpack_pos(StepArgs A, const double *p, int count, int row0)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = row0 + k;
    A.st.pos[3 * i] = p[3 * k]; A.st.pos[3 * i + 1] = p[3 * k + 1]; A.st.pos[3 * i + 2] = p[3 * k + 2];
}

cudaError_t launch_halo_pack(const StepArgs &A, int n_own, int axis, double lo_face, double hi_face, double width,
                             const double park[3], void *out_lo, void *out_hi, int cap, int32_t *sent_lo, int32_t *sent_hi,
                             int32_t *cnt, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(cnt, 0, 2 * sizeof(int32_t), s);
    if (e != cudaSuccess) return e;
    if (cap > 0) k_halo_fill<<<(cap + 255) / 256, 256, 0, s>>>((HaloRec *)out_lo, (HaloRec *)out_hi, cap, park[0], park[1], park[2]);
    if (n_own > 0)
        k_halo_pack<<<(n_own + 255) / 256, 256, 0, s>>>(A, n_own, axis, lo_face, hi_face, width, (HaloRec *)out_lo, (HaloRec *)out_hi,
                                                       cap, sent_lo, sent_hi, cnt);
    return cudaGetLastError();
}

cudaError_t launch_halo_unpack(const StepArgs &A, const void *recs, int count, int row0, cudaStream_t s)
{
    if (count > 0) k_halo_unpack<<<(count + 255) / 256, 256, 0, s>>>(A, (const HaloRec *)recs, count, row0);
    return cudaGetLastError();
}

cudaError_t launch_halo_pack_pos(const StepArgs &A, const int32_t *sent_lo, const int32_t *sent_hi, const int32_t *cnt, int cap,
                                 double *out_lo, double *out_hi, const double park[3], cudaStream_t s)
{
    if (cap > 0)
        k_halo_pack_pos<<<(2 * cap + 255) / 256, 256, 0, s>>>(A, sent_lo, sent_hi, cnt, cap, out_lo, out_hi, park[0], park[1], park[2]);
    return cudaGetLastError();
}

cudaError_t launch_halo_unpack_pos(const StepArgs &A, const double *p, int count, int row0, cudaStream_t s)
{
    if (count > 0) k_halo_unpack_pos<<<(count + 255) / 256, 256, 0, s>>>(A, p, count, row0);
    return cudaGetLastError();
}
