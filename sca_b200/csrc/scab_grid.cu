// scab_grid.cu -- stage 1: uniform-grid cell list over all agents, rebuilt every step.
// Replaces KDTree.buildAgentTree() (mamp/policies/kdTree.py:56-122, called from mampenv.py:28).
//
// count (atomic histogram) -> exclusive scan over cells -> reorder into a cell-sorted snapshot
//   s_posr[ x y z r ] (32 B)  s_vel[ vx vy vz status ] (16 B)  s_id  rank
// The snapshot is what every later stage gathers from, so the policy stage can update the caller's
// state arrays in place without racing with neighbour reads (the reference reads a frozen snapshot of
// the other agents too: positions only change in mampenv.py:42-46).
#include "scab_common.cuh"

__global__ void __launch_bounds__(256) k_cell_count(StepArgs A)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int cx, cy, cz;
    int c = cell_of(A.g, A.st.pos[3 * i], A.st.pos[3 * i + 1], A.st.pos[3 * i + 2], cx, cy, cz);
    A.cell_id[i] = c;
    A.cell_local[i] = atomicAdd(&A.cell_count[c], 1);
}

// Single-CTA exclusive scan: each thread owns a contiguous chunk. Cells <= a few 1e5.
__global__ void __launch_bounds__(1024) k_cell_scan(StepArgs A)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int ncell = A.g.ncell;
    if (t == 0) carry_s = 0;
    __syncthreads();
    // tiles of T*4 cells: coalesced int4-free simple loop (cells count is small vs agents)
    for (int base = 0; base < ncell; base += T * 4) {
        int idx = base + t * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (idx + k < ncell) ? A.cell_count[idx + k] : 0;
        int sum = v[0] + v[1] + v[2] + v[3];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = (lane < (T >> 5)) ? warp_sums[lane] : 0;
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += y;
            }
            warp_sums[lane] = wi - w;  // exclusive
        }
        __syncthreads();
        int excl = carry_s + warp_sums[wid] + incl - sum;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (idx + k < ncell) A.cell_start[idx + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (t == T - 1) carry_s = excl;
        __syncthreads();
    }
    if (t == 0) A.cell_start[ncell] = carry_s;
}

__global__ void __launch_bounds__(256) k_reorder(StepArgs A)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int dst = A.cell_start[A.cell_id[i]] + A.cell_local[i];
    double4 p;
    p.x = A.st.pos[3 * i]; p.y = A.st.pos[3 * i + 1]; p.z = A.st.pos[3 * i + 2];
    p.w = A.st.radius[i];
    float4 v;
    v.x = A.st.vel[3 * i]; v.y = A.st.vel[3 * i + 1]; v.z = A.st.vel[3 * i + 2];
    v.w = __int_as_float(A.st.status[i]);
    A.s_posr[dst] = p;
    A.s_vel[dst] = v;
    A.s_id[dst] = i;
    A.s_gid[dst] = A.st.gid ? A.st.gid[i] : i;
    A.rank[i] = dst;
}

cudaError_t launch_grid_build(const StepArgs &A, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(A.cell_count, 0, sizeof(int32_t) * A.g.ncell, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(A.counters, 0, sizeof(unsigned long long) * 4, s);     // per-step counters [0..3]
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(A.counters + 5, 0, sizeof(unsigned long long) * 2, s); // per-step diagnostics [5..6]
    if (e != cudaSuccess) return e;
    if (A.n > 0) {
        int nb = (A.n + 255) / 256;
        k_cell_count<<<nb, 256, 0, s>>>(A);
        k_cell_scan<<<1, 1024, 0, s>>>(A);
        k_reorder<<<nb, 256, 0, s>>>(A);
    }
    return cudaGetLastError();
}
