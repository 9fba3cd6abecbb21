// scab_grid.cu -- stage 1: uniform-grid cell list over all agents, rebuilt every step.
// Replaces KDTree.buildAgentTree() (mamp/policies/kdTree.py:56-122, called from mampenv.py:28).
//
// count (atomic histogram) -> exclusive scan over cells (3 small launches) -> reorder into a cell-sorted snapshot
//   s_posr[ x y z r ] (32 B)  s_vel[ vx vy vz status ] (16 B)  s_id  rank
// The snapshot is what every later stage gathers from, so the policy stage can update the caller's
// state arrays in place without racing with neighbour reads (the reference reads a frozen snapshot of
// the other agents too: positions only change in mampenv.py:42-46).
#include "scab_common.cuh"

__global__ void __launch_bounds__(256) k_cell_count(StepArgs A)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {     // per-step counters [0..3] and diagnostics [5..6] (nothing accumulates into them before this stage ends)
        A.counters[0] = A.counters[1] = A.counters[2] = A.counters[3] = 0ull;
        A.counters[5] = A.counters[6] = 0ull;
    }
    if (i >= A.n) return;
    int cx, cy, cz;
    int c = cell_of(A.g, A.st.pos[3 * i], A.st.pos[3 * i + 1], A.st.pos[3 * i + 2], cx, cy, cz);
    A.cell_id[i] = c;
    A.cell_local[i] = atomicAdd(&A.cell_count[c], 1);
}

// Exclusive scan over the cell counts in three small launches (tiles of 4096 cells, one CTA each):
//   k_scan_tile_sums  tile total -> parked in cell_start[tile * SCAN_TILE] (a slot the last pass overwrites anyway)
//   k_scan_tiles      one CTA: exclusive scan of the tile totals in place, grand total -> cell_start[ncell]
//   k_scan_apply      tile-local exclusive scan + the tile's base
// (a single-CTA scan over the 2*10^5 cells of the 10^6-agent workload took 89 us, half of the grid stage)
#define SCAN_TILE 4096

__device__ __forceinline__ int block_exclusive_scan_1024(int sum, int *warp_sums, int &block_total)
{   // 1024 threads; returns the exclusive prefix of `sum` over the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int w = warp_sums[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warp_sums[lane] = wi - w;  // exclusive
        if (lane == 31) warp_sums[32] = wi;
    }
    __syncthreads();
    block_total = warp_sums[32];
    return warp_sums[wid] + incl - sum;
}

__global__ void __launch_bounds__(1024) k_scan_tile_sums(StepArgs A)
{
    __shared__ int warp_sums[33];
    const int ncell = A.g.ncell, idx = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) sum += (idx + k < ncell) ? A.cell_count[idx + k] : 0;
    int total;
    block_exclusive_scan_1024(sum, warp_sums, total);
    if (threadIdx.x == 0) A.cell_start[blockIdx.x * SCAN_TILE] = total;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(StepArgs A, int ntiles)
{
    __shared__ int warp_sums[33];
    int carry = 0;
    for (int base = 0; base < ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int v = (t < ntiles) ? A.cell_start[t * SCAN_TILE] : 0;
        int total;
        const int excl = block_exclusive_scan_1024(v, warp_sums, total);
        if (t < ntiles) A.cell_start[t * SCAN_TILE] = carry + excl;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) A.cell_start[A.g.ncell] = carry;
}

__global__ void __launch_bounds__(1024) k_scan_apply(StepArgs A)
{
    __shared__ int warp_sums[33];
    __shared__ int base_s;
    const int ncell = A.g.ncell, idx = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    if (threadIdx.x == 0) base_s = A.cell_start[blockIdx.x * SCAN_TILE];
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = (idx + k < ncell) ? A.cell_count[idx + k] : 0;
    int total;
    int excl = block_exclusive_scan_1024(v[0] + v[1] + v[2] + v[3], warp_sums, total) + base_s;   // (base_s is read after the scan's barriers)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (idx + k < ncell) A.cell_start[idx + k] = excl;
        excl += v[k];
    }
}

__global__ void __launch_bounds__(256) k_reorder(StepArgs A)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int dst = A.cell_start[A.cell_id[i]] + A.cell_local[i];
    double4 p;
    p.x = A.st.pos[3 * i]; p.y = A.st.pos[3 * i + 1]; p.z = A.st.pos[3 * i + 2];
    p.w = A.st.radius[i];
    if (A.max_radius > 0.0 && p.w > A.max_radius) atomicOr(&A.counters[7], OVF_RADIUS);   // a declared bound must hold: be loud
    float4 v;
    v.x = A.st.vel[3 * i]; v.y = A.st.vel[3 * i + 1]; v.z = A.st.vel[3 * i + 2];
    v.w = __int_as_float(A.st.status[i]);
    A.s_posr[dst] = p;
    A.s_posf[dst] = make_float4((float)(p.x - A.g.ox), (float)(p.y - A.g.oy), (float)(p.z - A.g.oz), 0.f);
    A.s_vel[dst] = v;
    A.s_id[dst] = i;
    A.s_gid[dst] = A.st.gid ? A.st.gid[i] : i;
    A.rank[i] = dst;
}

cudaError_t launch_grid_build(const StepArgs &A, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(A.cell_count, 0, sizeof(int32_t) * A.g.ncell, s);
    if (e != cudaSuccess) return e;
    if (A.n <= 0) {     // (otherwise k_cell_count zeroes them)
        e = cudaMemsetAsync(A.counters, 0, sizeof(unsigned long long) * 4, s);     // per-step counters [0..3]
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(A.counters + 5, 0, sizeof(unsigned long long) * 2, s); // per-step diagnostics [5..6]
        if (e != cudaSuccess) return e;
    }
    if (A.n > 0) {
        int nb = (A.n + 255) / 256;
        k_cell_count<<<nb, 256, 0, s>>>(A);
        const int ntiles = (A.g.ncell + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_tile_sums<<<ntiles, 1024, 0, s>>>(A);
        k_scan_tiles<<<1, 1024, 0, s>>>(A, ntiles);
        k_scan_apply<<<ntiles, 1024, 0, s>>>(A);
        k_reorder<<<nb, 256, 0, s>>>(A);
    }
    return cudaGetLastError();
}
