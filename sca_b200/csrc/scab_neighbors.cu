// scab_neighbors.cu -- stage 2: the neighbour query.
// Replaces computeNeighbors (mamp/policies/sca/scaPolicy.py:107-116), i.e.
//   KDTree.computeObstacleNeighbors / computeAgentNeighbors (kdTree.py:124-156, 229-262)
//   -> Agent.insertObstacleNeighbor / insertAgentNeighbor   (mamp/agents/agent.py:79-124).
//
// Two semantics (include/scab200.h scab_nbr_mode):
//   KNN         cell-list scan, exact 16 smallest by (distSq, obstacles-before-agents, id);
//   REF_COMPAT  replays the reference's kd-tree traversal and its insertion rule literally (the
//               by-value rangeSq never shrinks and a full list evicts its last entry unconditionally),
//               so the ORDERED lists are bit-identical to the reference's.
// Distances are fp64 with the reference's decimal rounding (util.py:100,104), so ties and the strict
// `< 100.0` / `< (rA+rB)^2` cuts are reproduced exactly.
#include "scab_common.cuh"
#include <float.h>
#include <stdlib.h>

struct Key {
    double d;
    int u;  // obstacles: obstacle id ; agents: m + agent id  (obstacles sort first on ties)
};
__device__ __forceinline__ bool key_less(double d, int u, double d2, int u2)
{
    return d < d2 || (d == d2 && u < u2);
}

// Register-resident sorted list of the 16 best keys; slots beyond cnt hold +inf.
struct TopK {
    double d[KNBR];
    int u[KNBR];
    int s[KNBR];  // payload: sorted index (agents) or n + sorted obstacle index
    int cnt;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int k = 0; k < KNBR; k++) { d[k] = DBL_MAX; u[k] = INT_MAX; s[k] = -1; }
        cnt = 0;
    }
    __device__ __forceinline__ void insert(double nd, int nu, int ns)
    {
        if (!key_less(nd, nu, d[KNBR - 1], u[KNBR - 1])) return;
#pragma unroll
        for (int k = KNBR - 1; k >= 1; k--) {
            bool lt_prev = key_less(nd, nu, d[k - 1], u[k - 1]);
            bool lt_here = key_less(nd, nu, d[k], u[k]);
            if (lt_prev) { d[k] = d[k - 1]; u[k] = u[k - 1]; s[k] = s[k - 1]; }
            else if (lt_here) { d[k] = nd; u[k] = nu; s[k] = ns; }
        }
        if (key_less(nd, nu, d[0], u[0])) { d[0] = nd; u[0] = nu; s[0] = ns; }
        cnt = min(cnt + 1, KNBR);
    }
};

__device__ __forceinline__ void write_lists(const StepArgs &A, int s, int id, int cnt, const double *d, const int *u,
                                            const int *sx, int qcoll)
{
    A.s_nbr_cnt[s] = cnt;
    A.s_qcoll[s] = qcoll;
    if (A.dub_d0) A.dub_d0[id] = cnt > 0 ? d[0] : -1.0;
    int4 *dst = reinterpret_cast<int4 *>(A.s_nbr + (size_t)s * KNBR);
#pragma unroll
    for (int q = 0; q < KNBR / 4; q++) dst[q] = make_int4(sx[4 * q], sx[4 * q + 1], sx[4 * q + 2], sx[4 * q + 3]);
    if (A.st.nbr_cnt) A.st.nbr_cnt[id] = cnt;
    if (A.st.nbr_ids) {
#pragma unroll
        for (int k = 0; k < KNBR; k++)
            A.st.nbr_ids[(size_t)id * KNBR + k] = (k < cnt) ? (u[k] < A.m ? A.obst_base + u[k] : u[k] - A.m) : -1;
    }
    if (A.st.nbr_dsq) {
#pragma unroll
        for (int k = 0; k < KNBR; k++) A.st.nbr_dsq[(size_t)id * KNBR + k] = (k < cnt) ? d[k] : 0.0;
    }
}

// The sorted list of the general path.  It lives in local memory and is updated by a plain insertion loop: this path runs
// for every agent only when obstacles share the 16 slots (the reference's own scenarios, N <= 10^4) and otherwise for the
// rare agents with tied distance keys, so compact code matters more than speed (the register-resident, fully unrolled
// version made every neighbour kernel 200 KB of SASS).  The tie-break id (obstacle id, or m + global agent id) is looked up
// only when two distances compare equal.
struct TopK2 {
    double d[KNBR];
    int s[KNBR];  // payload: sorted index (agents) or n + sorted obstacle index
    int cnt;
    __device__ __forceinline__ void clear() { cnt = 0; }
    static __device__ __forceinline__ int uid(const StepArgs &A, int sx)
    {
        return sx >= A.n ? A.o_id[sx - A.n] : A.m + A.s_gid[sx];
    }
    static __device__ __forceinline__ bool less(const StepArgs &A, double d1, int s1, double d2, int s2)
    {
        if (d1 != d2) return d1 < d2;
        return uid(A, s1) < uid(A, s2);
    }
    __device__ __noinline__ void insert(const StepArgs &A, double nd, int ns)
    {
        if (cnt == KNBR && !less(A, nd, ns, d[KNBR - 1], s[KNBR - 1])) return;
        int k = min(cnt, KNBR - 1);            // a full list drops its last entry
#pragma unroll 1
        while (k > 0 && less(A, nd, ns, d[k - 1], s[k - 1])) { d[k] = d[k - 1]; s[k] = s[k - 1]; k--; }
        d[k] = nd; s[k] = ns;
        cnt = min(cnt + 1, KNBR);
    }
};

__device__ __noinline__ void write_lists2(const StepArgs &A, int s, int id, const TopK2 &L, int qcoll)
{
    A.s_nbr_cnt[s] = L.cnt;
    A.s_qcoll[s] = qcoll;
    if (A.dub_d0) A.dub_d0[id] = L.cnt > 0 ? L.d[0] : -1.0;
#pragma unroll 1
    for (int k = 0; k < KNBR; k++) {
        const bool live = k < L.cnt;
        const int sx = live ? L.s[k] : -1;
        A.s_nbr[(size_t)s * KNBR + k] = sx;
        if (A.st.nbr_ids) A.st.nbr_ids[(size_t)id * KNBR + k] = !live ? -1 : (sx >= A.n ? A.obst_base + A.o_id[sx - A.n] : A.s_gid[sx]);
        if (A.st.nbr_dsq) A.st.nbr_dsq[(size_t)id * KNBR + k] = live ? L.d[k] : 0.0;
    }
    if (A.st.nbr_cnt) A.st.nbr_cnt[id] = L.cnt;
}

// ------------------------------------------------------------------ KNN over the cell list
#ifndef KNN_CAP
#define KNN_CAP 48      // candidate slots per agent of the tile variant (in range + pre-filter guard)
#endif
// The exact general path (obstacles, ties on the decimal-rounded distance, more than KNN_CAP raw candidates):
// fp64 keys + the (distSq, obstacles-before-agents, id) comparator.
__device__ __noinline__ void knn_exact(const StepArgs &A, int s, int id, const double4 me)
{
    TopK2 L;
    L.clear();
    bool coll = false;
    const double rangeSq = A.nbr_dist_sq;

    // obstacles first (scaPolicy.py:114), from the static obstacle cell list
    if (A.m > 0) {
        int cx, cy, cz;
        cell_of(A.og, me.x, me.y, me.z, cx, cy, cz);
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            const int z = cz + r / 3 - 1, y = cy + r % 3 - 1;
            if (z < 0 || z >= A.og.nz || y < 0 || y >= A.og.ny) continue;
            const int row = (z * A.og.ny + y) * A.og.nx;
            const int beg = A.o_cell_start[row + max(cx - 1, 0)];
            const int end = A.o_cell_start[row + min(cx + 1, A.og.nx - 1) + 1];
#pragma unroll 1
            for (int j = beg; j < end; j++) {
                const double4 o = A.o_posr[j];
                // agent.py:104-107
                double ss = sumsq3(__dsub_rn(me.x, o.x), __dsub_rn(me.y, o.y), __dsub_rn(me.z, o.z));
                double dsq1 = np_round5(ss);
                double t = __dsub_rn(py_round5(__dsqrt_rn(ss)), o.w);
                double dsq = __dmul_rn(t, t);
                double rr = __dadd_rn(me.w, o.w);
                bool is_c = dsq1 < __dmul_rn(rr, rr) && dsq < rangeSq;
                if (is_c && !coll) { coll = true; L.clear(); }
                if (coll ? is_c : (dsq < rangeSq)) L.insert(A, dsq, A.n + j);
            }
        }
    }
    // agents (scaPolicy.py:116): the 27 cells, a cheap raw-distance reject, then the exact decimal rounding
    {
        int cx, cy, cz;
        cell_of(A.g, me.x, me.y, me.z, cx, cy, cz);
        const double lim = rangeSq + 1e-4;   // the 5-decimal rounding cannot bring anything beyond this back in range
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            const int z = cz + r / 3 - 1, y = cy + r % 3 - 1;
            if (z < 0 || z >= A.g.nz || y < 0 || y >= A.g.ny) continue;
            const int row = (z * A.g.ny + y) * A.g.nx;
            const int beg = A.cell_start[row + max(cx - 1, 0)];
            const int end = A.cell_start[row + min(cx + 1, A.g.nx - 1) + 1];
#pragma unroll 1
            for (int j = beg; j < end; j++) {
                const double4 o = A.s_posr[j];
                const double ex = me.x - o.x, ey = me.y - o.y, ez = me.z - o.z;
                if (ex * ex + ey * ey + ez * ez > lim || j == s) continue;
                const double dsq = l3normsq_d(me.x, me.y, me.z, o.x, o.y, o.z);   // agent.py:81
                if (!(dsq < rangeSq)) continue;
                const double rr = __dadd_rn(me.w, o.w);
                const bool is_c = dsq < __dmul_rn(rr, rr);
                if (is_c && !coll) { coll = true; L.clear(); }
                if (coll && !is_c) continue;
                L.insert(A, dsq, j);
            }
        }
    }
    write_lists2(A, s, id, L, coll ? 1 : 0);
    A.r16_prev[id] = 3.0e38f;
}


// ------------------------------------------------------------------ the scale path
// Common case (no obstacles in the merged list, distinct distance keys), two phases per agent (one thread each):
//   phase 1  fp32 squared distance of every record of the (trimmed) 27 cells against rangeSq + a guard band that covers the
//            fp32 error; survivors go to a shared-memory column per thread, so that
//   phase 2  all lanes walk their candidates together: the reference's fp64 distance with its decimal rounding
//            (util.py:99-100) -- an exact integer number of 1e-5 units (< 2^24 for neighbor_dist < 12.9 m), so
//            (units << 8 | candidate slot) is a 32-bit sort key and inserting into the sorted top-16 is a branch-free
//            min/max network on 16 registers:  new[k] = min(old[k], max(old[k-1], x)).
// A tie on the distance (the reference breaks it by agent id) is detected afterwards -- two equal neighbours in the list,
// or the best rejected key equal to the last kept one -- and sends the agent to knn_exact, as does a full column.
// `rangeSq`: the squared query radius the candidates were collected for -- neighbor_dist^2, or less when a dense
// neighbourhood overflowed the candidate column and the scan was repeated with a smaller ball (the 16 nearest are then
// certainly inside it, provided it still holds 16 agents: checked here).
// Returns 0 when the lists were written, 1 when the agent needs the exact path (tie, full column), 2 when the reduced ball
// held fewer than 16 agents (the caller retries with the full radius).
template <class GetJ>
__device__ __forceinline__ int knn_phase2(const StepArgs &A, int s, int id, const double4 me, int nbuf, bool overflow, double rangeSq, GetJ getj)
{
    unsigned L[KNBR];
#pragma unroll
    for (int k = 0; k < KNBR; k++) L[k] = 0xffffffffu;
    unsigned rej_min = 0xffffffffu;
    int cnt = 0;
    bool coll = false;
#pragma unroll 1
    for (int i0 = 0; i0 < nbuf; i0 += 4) {
        double4 o4[4];             // four gathers in flight per thread (L1 / L2 hits)
        int j4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { j4[q] = getj(min(i0 + q, nbuf - 1)); o4[q] = A.s_posr[j4[q]]; }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + q;
            if (i >= nbuf || j4[q] == s) continue;
            const double4 o = o4[q];
            // l3normsq (util.py:99-100): round(sum, 5) = rint(sum * 1e5) / 1e5
            const double units = rint(__dmul_rn(sumsq3(__dsub_rn(me.x, o.x), __dsub_rn(me.y, o.y), __dsub_rn(me.z, o.z)), 1e5));
            const double dsq = div1e5(units);
            if (!(dsq < rangeSq)) continue;
            const double rr = __dadd_rn(me.w, o.w);
            const bool is_c = dsq < __dmul_rn(rr, rr);      // agent.py:83-85: from the first collider on, colliders only
            if (is_c && !coll) {
                coll = true;
#pragma unroll
                for (int k = 0; k < KNBR; k++) L[k] = 0xffffffffu;
                rej_min = 0xffffffffu;
                cnt = 0;
            }
            if (coll && !is_c) continue;
            const unsigned x = ((unsigned)(int)units << 8) | (unsigned)i;
            rej_min = min(rej_min, max(L[KNBR - 1], x));
#pragma unroll
            for (int k = KNBR - 1; k >= 1; k--) L[k] = min(L[k], max(L[k - 1], x));
            L[0] = min(L[0], x);
            cnt++;
        }
    }
    const int n_in = min(cnt, KNBR);
    if (!overflow && rangeSq < A.nbr_dist_sq && !coll && cnt < KNBR) return 2;
    bool tie = overflow || (cnt > KNBR && (rej_min >> 8) == (L[KNBR - 1] >> 8));
#pragma unroll
    for (int k = 1; k < KNBR; k++) tie |= (k < n_in) && ((L[k] >> 8) == (L[k - 1] >> 8));
    if (tie) return 1;
    // how far the next query has to look: the 16th neighbour's distance (a shorter or colliders-only list says nothing)
    A.r16_prev[id] = (n_in == KNBR && !coll) ? (float)sqrt(div1e5((double)(L[KNBR - 1] >> 8))) * 1.000001f : 3.0e38f;
    int sx[KNBR];
#pragma unroll
    for (int k = 0; k < KNBR; k++) sx[k] = (k < n_in) ? getj((int)(L[k] & 0xffu)) : -1;
    A.s_nbr_cnt[s] = n_in;
    A.s_qcoll[s] = coll ? 1 : 0;
    if (A.dub_d0) A.dub_d0[id] = n_in > 0 ? div1e5((double)(L[0] >> 8)) : -1.0;
    int4 *dst = reinterpret_cast<int4 *>(A.s_nbr + (size_t)s * KNBR);
#pragma unroll
    for (int q = 0; q < KNBR / 4; q++) dst[q] = make_int4(sx[4 * q], sx[4 * q + 1], sx[4 * q + 2], sx[4 * q + 3]);
    if (A.st.nbr_cnt) A.st.nbr_cnt[id] = n_in;
    if (A.st.nbr_ids) {
#pragma unroll
        for (int k = 0; k < KNBR; k++) A.st.nbr_ids[(size_t)id * KNBR + k] = (k < n_in) ? A.s_gid[sx[k]] : -1;
    }
    if (A.st.nbr_dsq) {
#pragma unroll
        for (int k = 0; k < KNBR; k++) A.st.nbr_dsq[(size_t)id * KNBR + k] = (k < n_in) ? div1e5((double)(L[k] >> 8)) : 0.0;
    }
    return 0;
}

// The trimmed x-range of one of the nine x-rows around cell (cx, cy, cz) for a query ball of squared radius `limt`:
// (fx, fy, fz) is the offset inside the own cell, so the gap to a neighbouring slab is f or cs - f.  Returns false when
// the ball does not reach the row.  (Agents clamped into a border cell from outside the grid scan the full block; cell
// contents are never closer than their cell.)
struct RowTrim {
    double fx, fy, fz, cs, limt;
    bool trim;
    __device__ __forceinline__ void init(const GridGeom &g, const double4 &me, int cx, int cy, int cz, double lim)
    {
        cs = 1.0 / g.inv_cs;
        fx = (me.x - g.ox) - cx * cs; fy = (me.y - g.oy) - cy * cs; fz = (me.z - g.oz) - cz * cs;
        trim = fx >= 0.0 && fx <= cs && fy >= 0.0 && fy <= cs && fz >= 0.0 && fz <= cs;
        limt = lim + 1e-6;
    }
    __device__ __forceinline__ bool row(int r, int cx, int &x0, int &x1) const
    {
        const int dz = r / 3 - 1, dy = r % 3 - 1;
        const double gz = dz == 0 ? 0.0 : (dz < 0 ? fz : cs - fz);
        const double gy = dy == 0 ? 0.0 : (dy < 0 ? fy : cs - fy);
        const double rem = limt - gz * gz - gy * gy;
        if (trim && rem < 0.0) return false;
        x0 = (trim && fx * fx > rem) ? cx : cx - 1;
        x1 = (trim && (cs - fx) * (cs - fx) > rem) ? cx : cx + 1;
        return true;
    }
};

// fp32 error of the squared distance of two points within ~10 m of each other at (grid-relative) coordinates of magnitude S:
// <= 2 d (e_x + e_y + e_z) with e ~ 2^-23 (|c| + 10), i.e. < 2.5e-6 S + 2e-4
__device__ __forceinline__ float prefilter_limit(double lim, float mx, float my, float mz)
{
    return (float)lim + 1e-3f + 4e-6f * (fabsf(mx) + fabsf(my) + fabsf(mz));
}

#ifndef KNN_VARIANT
#define KNN_VARIANT 1      // 0: tile-staged scan (k_neighbors_tile)   1: records straight from L1 / L2 (k_neighbors_direct)
#endif
#ifndef KNN_UNROLL
#define KNN_UNROLL 4       // records loaded per inner iteration of the direct scan (must divide 32)
#endif
#ifndef KNN_DIRECT_CTAS
#define KNN_DIRECT_CTAS 6    // resident CTAs per SM the direct kernel is compiled for (80 registers, no spills); measured on B200 at 10^6
#endif                       // agents: 5 -> 0.398 ms, 6 -> 0.372, 7 (72 registers, spills) -> 0.387, 8 -> 0.424; KNN_UNROLL 8: +0.004

// ---- variant 0: one CTA per WORK ITEM = NT_W consecutive cells of one x-row of the grid.  Its agents are a contiguous range
// of the cell-sorted snapshot, and everything they can see lies in nine contiguous ranges (the 3 x 3 neighbouring x-rows,
// cells cx0-1 .. cx0+NT_W): those are staged ONCE into shared memory as fp32 float4 records (coalesced 128-bit loads of
// s_posf, the copy k_reorder writes) and every agent of the item scans its own 27 cells out of the tile.  Candidates are
// recorded as 16-bit (row << 11 | offset in the row's range).  An item whose ranges do not fit the tile (very dense cells)
// sends its agents through knn_exact.
#ifndef NT_W
#define NT_W 16
#define NT_THREADS 128
#define NT_TILE 1408
#endif
#ifndef NT_MIN_CTAS
#define NT_MIN_CTAS 5
#endif

__global__ void __launch_bounds__(NT_THREADS, NT_MIN_CTAS) k_neighbors_tile(const __grid_constant__ StepArgs A, int segs)
{
    __shared__ float4 tile[NT_TILE];
    __shared__ unsigned short buf[KNN_CAP][NT_THREADS];
    __shared__ int cs[9][NT_W + 3];
    __shared__ int rbase[9];
    __shared__ int use_tile;
    static_assert(NT_TILE <= 2048, "candidate codes hold an 11-bit offset");
    const int tid = threadIdx.x;
    const int row = blockIdx.x / segs, cx0 = (blockIdx.x - row * segs) * NT_W;
    const int nx = A.g.nx, ny = A.g.ny;
    const int y0 = row % ny, z0 = row / ny;
    const int a_beg = A.cell_start[row * nx + cx0], a_end = A.cell_start[row * nx + min(cx0 + NT_W, nx)];
    if (a_beg == a_end) return;
    {   // nothing to do for an item whose agents are all finished / ghosts (parked rows of a sharded run)
        int act = 0;
        for (int s = a_beg + tid; s < a_end; s += NT_THREADS) {
            const bool inact = (__float_as_int(A.s_vel[s].w) & FL_INACTIVE) != 0;
            if (inact) {
                const int id = A.s_id[s];
                A.s_nbr_cnt[s] = -1;
                A.s_qcoll[s] = 0;
                if (A.st.nbr_cnt) A.st.nbr_cnt[id] = -1;
            } else act = 1;
        }
        if (!__syncthreads_or(act)) return;
    }
    // cell bounds of the nine x-rows: entry k <-> cell cx0 - 1 + k (clamped); rows outside the grid are empty
    for (int t = tid; t < 9 * (NT_W + 3); t += NT_THREADS) {
        const int r = t / (NT_W + 3), k = t - r * (NT_W + 3);
        const int z = z0 + r / 3 - 1, y = y0 + r % 3 - 1;
        int v = 0;
        if (z >= 0 && z < A.g.nz && y >= 0 && y < ny) v = A.cell_start[(z * ny + y) * nx + min(max(cx0 - 1 + k, 0), nx)];
        cs[r][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int r = 0; r < 9; r++) { rbase[r] = tot; tot += cs[r][NT_W + 2] - cs[r][0]; }
        use_tile = tot <= NT_TILE;
    }
    __syncthreads();
    const bool staged = use_tile != 0;
    if (staged) {
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            const int src = cs[r][0], len = cs[r][NT_W + 2] - src, dst = rbase[r];
            for (int i = tid; i < len; i += NT_THREADS) tile[dst + i] = A.s_posf[src + i];
        }
    }
    __syncthreads();

#pragma unroll 1
    for (int s = a_beg + tid; s < a_end; s += NT_THREADS) {
        const int flags = __float_as_int(A.s_vel[s].w);
        if (flags & FL_INACTIVE) continue;
        const int id = A.s_id[s];
        const double4 me = A.s_posr[s];
        if (!staged) { knn_exact(A, s, id, me); continue; }
        // ---- phase 1: fp32 pre-filter over the (trimmed) 27 cells, out of the tile
        int nbuf = 0;
        bool overflow = false;
        int cx, cy, cz;
        cell_of(A.g, me.x, me.y, me.z, cx, cy, cz);
        const double lim = A.nbr_dist_sq + 1e-4;   // the 5-decimal rounding cannot bring anything beyond this back in range
        RowTrim T;
        T.init(A.g, me, cx, cy, cz, lim);
        const float mx = (float)(me.x - A.g.ox), my = (float)(me.y - A.g.oy), mz = (float)(me.z - A.g.oz);
        const float limf = prefilter_limit(lim, mx, my, mz);
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            int x0, x1;
            if (!T.row(r, cx, x0, x1)) continue;
            const int r0 = cs[r][0];
            const int beg = cs[r][max(x0, 0) - (cx0 - 1)] - r0, end = cs[r][min(x1, nx - 1) + 1 - (cx0 - 1)] - r0;   // offsets in the row's range
            const float4 *rec = tile + rbase[r];
            const unsigned code = (unsigned)r << 11;
#pragma unroll 1
            for (int j0 = beg; j0 < end; j0 += 4) {
                float4 o[4];
#pragma unroll
                for (int q = 0; q < 4; q++) o[q] = rec[min(j0 + q, end - 1)];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float ex = mx - o[q].x, ey = my - o[q].y, ez = mz - o[q].z;
                    const float ss = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
                    if (ss <= limf && j0 + q < end) {           // (the agent itself included: phase 2 drops it)
                        if (nbuf < KNN_CAP) buf[nbuf++][tid] = (unsigned short)(code | (unsigned)(j0 + q));
                        else overflow = true;
                    }
                }
            }
        }
        if (knn_phase2(A, s, id, me, nbuf, overflow, A.nbr_dist_sq, [&](int i) { const unsigned c = buf[i][tid]; return cs[c >> 11][0] + (int)(c & 2047u); }))
            knn_exact(A, s, id, me);
    }
}

// ---- variant 1: thread per agent, records straight from global memory (s_posf: 16-byte fp32 records, L1 / L2 hits).
// DENSE = false: every agent, one scan; an agent with more than KNN_CAP candidates in range is appended to a work list.
// DENSE = true: the agents of that list, scanned again with half the squared radius until the column holds them (the 16
// nearest of a crowded neighbourhood are certainly inside a smaller ball).  Keeping the retry out of the first kernel keeps
// a warp from re-scanning for one crowded lane (BASELINE config 4 has ~64 agents in range of each other; the uniform
// workload gets there locally late in a rollout).
// Candidates are recorded as 16-bit codes (row << 12 | offset in the row's range): 80 slots per agent in 20 KB per CTA.
#ifndef KND_CAP
#define KND_CAP 80      // measured at 6 resident CTAs (10^6 uniform / late rollout / config 4): 64 -> 0.363 / 0.409 / 0.653 ms,
#endif                  // 80 -> 0.366 / 0.413 / 0.221, 96 -> 0.371 / 0.419 / 0.226, 112 -> 0.386 / 0.440 / 0.304
template <bool DENSE>
__global__ void __launch_bounds__(128, KNN_DIRECT_CTAS) k_neighbors_direct(const __grid_constant__ StepArgs A)
{
    __shared__ unsigned short buf[KND_CAP][128];
    __shared__ int rowbase[9][128];          // first sorted index of each row's range, per thread (decodes the candidate codes)
    const int t = blockIdx.x * blockDim.x + threadIdx.x, tid = threadIdx.x;
    const int n_work = DENSE ? min(*A.slow_count, A.n) : A.n;
    if (t >= n_work) return;
    const int s = DENSE ? A.slow_list[t] : t;
    const int id = A.s_id[s];
    const int flags = __float_as_int(A.s_vel[s].w);
    if (flags & FL_INACTIVE) {
        A.s_nbr_cnt[s] = -1;
        A.s_qcoll[s] = 0;
        if (A.st.nbr_cnt) A.st.nbr_cnt[id] = -1;
        return;
    }
    const double4 me = A.s_posr[s];
    int nbuf = 0;
    bool overflow = false;
    int cx, cy, cz;
    cell_of(A.g, me.x, me.y, me.z, cx, cy, cz);
    const float mx = (float)(me.x - A.g.ox), my = (float)(me.y - A.g.oy), mz = (float)(me.z - A.g.oz);
    // Temporal coherence: neither this agent nor its neighbours move more than move_bound per step, so the 16th neighbour is
    // at most 2 move_bound farther away than it was at the previous query.  (If the state was changed from outside in
    // between, the ball may hold fewer than 16 agents: phase 2 notices and the agent is redone with the full radius.)
    // A collider displaces every non-collider from the list (agent.py:82-92) however far it is, so the ball also has to hold
    // every possible collider: r_i + the largest radius (scab_set_max_radius; unknown -> no reduced ball).
    double range_used = A.nbr_dist_sq;
    if (!DENSE && A.max_radius > 0.0) {
        const double r = fmax((double)A.r16_prev[id] + 2.0 * A.move_bound, me.w + A.max_radius) + 1e-3;
        if (r * r < range_used) range_used = r * r;
    }
    RowTrim T;          // (the smaller ball also reaches fewer cells)
    T.init(A.g, me, cx, cy, cz, range_used + 1e-4);
    float limf = prefilter_limit(range_used + 1e-4, mx, my, mz);
    // the (up to) nine x-rows of the block: all their cell_start bounds are fetched before any row is scanned
    int rbeg[9], rend[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int z = cz + r / 3 - 1, y = cy + r % 3 - 1;
        int x0, x1;
        const bool ok = z >= 0 && z < A.g.nz && y >= 0 && y < A.g.ny && T.row(r, cx, x0, x1);
        const int rowi = (z * A.g.ny + y) * A.g.nx;
        rbeg[r] = ok ? A.cell_start[rowi + max(x0, 0)] : 0;
        rend[r] = ok ? A.cell_start[rowi + min(x1, A.g.nx - 1) + 1] : 0;
        rowbase[r][tid] = rbeg[r];
    }
    // Per row, the hits of up to 32 consecutive records are collected as a bit mask (one instruction per record instead of
    // a predicated store sequence) and then drained into the thread's shared-memory column.  A neighbourhood too dense for
    // the column (> KNN_CAP agents in range: BASELINE config 4 has ~64) is scanned again with half the squared radius.
#pragma unroll 1
    for (int attempt = 0; attempt < (DENSE ? 5 : 1); attempt++) {
    nbuf = 0;
    overflow = false;
    int nhit = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int end = rend[r];
        if (end - rbeg[r] > 4096) overflow = true;               // (a row the 12-bit offsets cannot address: exact path)
#pragma unroll 1
        for (int c0 = rbeg[r]; c0 < end; c0 += 32) {
            const int n = min(32, end - c0);
            unsigned m = 0u;
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += KNN_UNROLL) {
                float4 o[KNN_UNROLL];          // independent loads in flight per thread
#pragma unroll
                for (int q = 0; q < KNN_UNROLL; q++) o[q] = __ldg(&A.s_posf[min(c0 + j0 + q, end - 1)]);
                unsigned mq = 0u;
#pragma unroll
                for (int q = 0; q < KNN_UNROLL; q++) {
                    const float ex = mx - o[q].x, ey = my - o[q].y, ez = mz - o[q].z;
                    const float ss = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
                    mq |= (ss <= limf ? 1u : 0u) << q;           // (the agent itself included: phase 2 drops it)
                }
                m |= mq << j0;
            }
            if (n < 32) m &= (1u << n) - 1u;                      // the clamped tail loads repeat the last record
            if (DENSE) nhit += __popc(m);
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                if (nbuf < KND_CAP) buf[nbuf++][tid] = (unsigned short)((r << 12) | ((c0 + b - rbeg[r]) & 4095));
                else overflow = true;
            }
        }
    }
    if (!DENSE || !overflow) break;
    // shrink the ball so that it holds ~36 of the nhit agents seen (count ~ r^3; a flatter crowd just takes another round),
    // but never below the reach of a collider (see above); a ball that cannot shrink leaves the agent to knn_exact
    const double floor_r = me.w + A.max_radius + 1e-3;
    if (!(A.max_radius > 0.0) || range_used <= floor_r * floor_r) break;
    const float sc = cbrtf(36.f / (float)nhit);
    range_used = fmax(range_used * (double)fminf(sc * sc, 0.8f), floor_r * floor_r);
    limf = prefilter_limit(range_used + 1e-4, mx, my, mz);
    }
    if (!DENSE && overflow) { A.slow_list[atomicAdd(A.slow_count, 1)] = s; return; }      // (the policy stage resets the list for its own use)
    const int rc = knn_phase2(A, s, id, me, nbuf, overflow, range_used, [&](int i) { const unsigned c = buf[i][tid]; return rowbase[c >> 12][tid] + (int)(c & 4095u); });
    if (rc == 2 && !DENSE) { A.slow_list[atomicAdd(A.slow_count, 1)] = s; return; }        // the reduced ball was too small: second pass, full radius
    if (rc) knn_exact(A, s, id, me);
}

// the general path for every agent (obstacles in the merged list, SCAB_KNN_EXACT=1, neighbor_dist >= 12.9 m): thread per agent
__global__ void __launch_bounds__(128) k_neighbors_exact(const __grid_constant__ StepArgs A)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    const int id = A.s_id[s];
    const int flags = __float_as_int(A.s_vel[s].w);
    if (flags & FL_INACTIVE) {
        A.s_nbr_cnt[s] = -1;
        A.s_qcoll[s] = 0;
        if (A.st.nbr_cnt) A.st.nbr_cnt[id] = -1;
        return;
    }
    knn_exact(A, s, id, A.s_posr[s]);
}

// ------------------------------------------------------------------ REF_COMPAT: literal replay
struct RefList {
    double d[KNBR];
    int u[KNBR];
    int s[KNBR];
    int cnt;
    bool coll;
    // agent.py:87-90 / 94-97: pop() when full, append, stable sort (newcomer after equal keys)
    __device__ __forceinline__ void append_sorted(double nd, int nu, int ns)
    {
        if (cnt == KNBR) cnt--;
        int k = cnt;
        while (k > 0 && d[k - 1] > nd) {
            d[k] = d[k - 1]; u[k] = u[k - 1]; s[k] = s[k - 1];
            k--;
        }
        d[k] = nd; u[k] = nu; s[k] = ns;
        cnt++;
    }
};

__device__ __forceinline__ double box_dist_sq(const KdNodeDev &nd, double x, double y, double z)
{  // kdTree.py:132-138, summed left to right
    double s = 0.0, t;
    t = fmax(0.0, __dsub_rn(nd.mn[0], x)); s = __dadd_rn(s, __dmul_rn(t, t));
    t = fmax(0.0, __dsub_rn(x, nd.mx[0])); s = __dadd_rn(s, __dmul_rn(t, t));
    t = fmax(0.0, __dsub_rn(nd.mn[1], y)); s = __dadd_rn(s, __dmul_rn(t, t));
    t = fmax(0.0, __dsub_rn(y, nd.mx[1])); s = __dadd_rn(s, __dmul_rn(t, t));
    t = fmax(0.0, __dsub_rn(nd.mn[2], z)); s = __dadd_rn(s, __dmul_rn(t, t));
    t = fmax(0.0, __dsub_rn(z, nd.mx[2])); s = __dadd_rn(s, __dmul_rn(t, t));
    return s;
}

#define KD_STACK 64
#define KD_MAX_LEAF 10  // kdTree.py:53

__global__ void __launch_bounds__(128) k_neighbors_compat(StepArgs A)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n) return;
    const int id = A.s_id[s];
    const int flags = __float_as_int(A.s_vel[s].w);
    if (flags & FL_INACTIVE) {
        A.s_nbr_cnt[s] = -1;
        A.s_qcoll[s] = 0;
        if (A.st.nbr_cnt) A.st.nbr_cnt[id] = -1;
        return;
    }
    const double4 me = A.s_posr[s];
    const double rangeSq = A.nbr_dist_sq;
    RefList L;
    L.cnt = 0;
    L.coll = false;
    int stack[KD_STACK];
    for (int pass = 0; pass < 2; pass++) {
        const bool obst = (pass == 0);
        if (obst && A.m <= 0) continue;     // kdTree.py:233 empty obstacle tree
        const KdNodeDev *nodes = obst ? A.kd_obst : A.kd_agent;
        const int32_t *perm = obst ? A.kd_obst_perm : A.kd_agent_perm;
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const int node = stack[--sp];
            const KdNodeDev nd = nodes[node];
            if (nd.end - nd.begin <= KD_MAX_LEAF) {
                for (int i = nd.begin; i < nd.end; i++) {
                    const int other = perm[i];
                    if (obst) {  // agent.py:101-124
                        const double4 o = A.o_posr_by_id[other];
                        double ss = sumsq3(__dsub_rn(me.x, o.x), __dsub_rn(me.y, o.y), __dsub_rn(me.z, o.z));
                        double dsq1 = np_round5(ss);
                        double t = __dsub_rn(py_round5(__dsqrt_rn(ss)), o.w);
                        double dsq = __dmul_rn(t, t);
                        double rr = __dadd_rn(me.w, o.w);
                        if (dsq1 < __dmul_rn(rr, rr) && dsq < rangeSq) {
                            if (!L.coll) { L.coll = true; L.cnt = 0; }
                            L.append_sorted(dsq, other, -1 - other);
                        } else if (!L.coll && dsq < rangeSq) {
                            L.append_sorted(dsq, other, -1 - other);
                        }
                    } else {     // agent.py:79-99
                        if (other == id) continue;
                        const int so = A.rank[other];
                        const double4 o = A.s_posr[so];
                        double dsq = l3normsq_d(me.x, me.y, me.z, o.x, o.y, o.z);
                        double rr = __dadd_rn(me.w, o.w);
                        if (dsq < __dmul_rn(rr, rr) && dsq < rangeSq) {
                            if (!L.coll) { L.coll = true; L.cnt = 0; }
                            L.append_sorted(dsq, A.m + other, so);
                        } else if (!L.coll && dsq < rangeSq) {
                            L.append_sorted(dsq, A.m + other, so);
                        }
                    }
                }
            } else {
                double dl = box_dist_sq(nodes[nd.left], me.x, me.y, me.z);
                double dr = box_dist_sq(nodes[nd.right], me.x, me.y, me.z);
                // kdTree.py:147-156: nearer child first; the farther one only if the nearer qualified
                if (dl < dr) {
                    if (dl < rangeSq) {
                        if (dr < rangeSq && sp < KD_STACK) stack[sp++] = nd.right;
                        if (sp < KD_STACK) stack[sp++] = nd.left;
                    }
                } else {
                    if (dr < rangeSq) {
                        if (dl < rangeSq && sp < KD_STACK) stack[sp++] = nd.left;
                        if (sp < KD_STACK) stack[sp++] = nd.right;
                    }
                }
            }
        }
    }
    // obstacles were tagged with -1-id payloads (unsorted ids): translate to n + sorted-obstacle index
    // is unnecessary here -- the policy stage only needs (pos, r) so it accepts n + obstacle ID in
    // compat mode through o_posr_by_id.  Encode as n + m + id to tell the two apart.
    int sx[KNBR];
#pragma unroll
    for (int k = 0; k < KNBR; k++) sx[k] = (k < L.cnt) ? (L.s[k] < 0 ? A.n + A.m + (-1 - L.s[k]) : L.s[k]) : -1;
    write_lists(A, s, id, L.cnt, L.d, L.u, sx, L.coll ? 1 : 0);
}

cudaError_t launch_neighbors(const StepArgs &A, int mode, cudaStream_t s)
{
    if (A.n <= 0) return cudaSuccess;
    int nb = (A.n + 127) / 128;
    if (mode == SCAB_NBR_REF_COMPAT) k_neighbors_compat<<<nb, 128, 0, s>>>(A);
    else {
        // the integer-key fast path needs: no obstacles in the merged list, and distance keys below 2^24 units of 1e-5
        const bool fast_ok = A.m == 0 && A.nbr_dist_sq * 1e5 < 16777215.0 && !A.cfg->knn_exact;   // SCAB_KNN_EXACT=1: A/B switch
        if (fast_ok && KNN_VARIANT == 0) {
            const int segs = (A.g.nx + NT_W - 1) / NT_W;
            k_neighbors_tile<<<A.g.ny * A.g.nz * segs, NT_THREADS, 0, s>>>(A, segs);
        } else if (fast_ok) {
            cudaError_t e = cudaMemsetAsync(A.slow_count, 0, sizeof(int32_t), s);
            if (e != cudaSuccess) return e;
            k_neighbors_direct<false><<<nb, 128, 0, s>>>(A);
            k_neighbors_direct<true><<<nb, 128, 0, s>>>(A);       // the crowded agents (usually none: the blocks exit at once)
        }
        else k_neighbors_exact<<<nb, 128, 0, s>>>(A);
    }
    return cudaGetLastError();
}
