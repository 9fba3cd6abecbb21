// scab_status.cu -- stage 4: collision / timeout / goal flags and the done reduction.
// Replaces the second loop of MACAEnv._take_action (mamp/envs/mampenv.py:42-46 -> check_agent_state
// :61-80) and MACAEnv.is_done (:51-59).
//
// The reference's check is O(N^2) and SEQUENTIAL: agent i is tested right after its own move against
// every other agent, of which j < i have already moved and j > i have not.  For a pair lo < hi this is
//     collide(lo, hi)  <=>  l3norm(new_lo, old_hi) <= R   or   l3norm(new_lo, new_hi) <= R
// (the first test happens when lo is processed, the second when hi is), with R = r_lo + r_hi; either
// hit flags both agents unless the flagged one is already at its goal.  That pair rule is evaluated here
// in parallel over the cell list built from the OLD positions: a step moves an agent by <= max_speed*dt
// << cell size, so every possible partner is inside the 27-cell neighbourhood.
#include "scab_common.cuh"

__global__ void __launch_bounds__(128) k_status(StepArgs A)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool not_done = false, collided = false;
    const int snap_flags = (s < A.n) ? __float_as_int(A.s_vel[s].w) : SCAB_ST_GHOST;
    if (!(snap_flags & SCAB_ST_GHOST)) {   // halo ghosts are partners only: their owner GPU flags them
        const int id = A.s_id[s];
        const int gid = A.s_gid[s];
        const double4 oldp = A.s_posr[s];
        // positions after the kinematic update: the policy stage wrote them in place (ghost rows are
        // refreshed by the halo exchange between the two stages)
        const double4 newp = make_double4(A.st.pos[3 * id], A.st.pos[3 * id + 1], A.st.pos[3 * id + 2], 0.0);
        const bool was_goal = (snap_flags & SCAB_ST_AT_GOAL) != 0;
        int st = A.st.status[id];   // includes the collision bit a neighbour query may have raised this step
        if (!was_goal) A.st.step_num[id] += 1;   // mampenv.py:44-45

        bool coll_obst = false, coll_agent = false;
        // obstacles (mampenv.py:63-66): flag regardless of at_goal
        if (A.m > 0) {
            int cx, cy, cz;
            cell_of(A.og, newp.x, newp.y, newp.z, cx, cy, cz);
            for (int dz = -1; dz <= 1; dz++) {
                int z = cz + dz;
                if (z < 0 || z >= A.og.nz) continue;
                for (int dy = -1; dy <= 1; dy++) {
                    int y = cy + dy;
                    if (y < 0 || y >= A.og.ny) continue;
                    int row = (z * A.og.ny + y) * A.og.nx;
                    int beg = A.o_cell_start[row + max(cx - 1, 0)];
                    int end = A.o_cell_start[row + min(cx + 1, A.og.nx - 1) + 1];
                    for (int j = beg; j < end; j++) {
                        const double4 o = A.o_posr[j];
                        if (l3norm_d(newp.x, newp.y, newp.z, o.x, o.y, o.z) <= __dadd_rn(oldp.w, o.w)) coll_obst = true;
                    }
                }
            }
        }
        // agents (mampenv.py:68-75) through the pair rule
        const double reach = 2.0 * A.move_bound + 1e-3;   // both partners move, plus slack for the 5-decimal rounding
        auto test_pair = [&](int j) {
            const double4 oo = A.s_posr[j];
            // cheap reject on the old positions: each agent moves <= move_bound per step
            const double ex = oo.x - oldp.x, ey = oo.y - oldp.y, ez = oo.z - oldp.z;
            const double R = __dadd_rn(oldp.w, oo.w);
            const double lim = R + reach;
            if (ex * ex + ey * ey + ez * ez > lim * lim) return;
            const int jrow = A.s_id[j];
            const double4 on = make_double4(A.st.pos[3 * jrow], A.st.pos[3 * jrow + 1], A.st.pos[3 * jrow + 2], 0.0);
            const int jid = A.s_gid[j];
            bool hit;
            if (gid < jid)
                hit = l3norm_d(newp.x, newp.y, newp.z, oo.x, oo.y, oo.z) <= R ||
                      l3norm_d(newp.x, newp.y, newp.z, on.x, on.y, on.z) <= R;
            else
                hit = l3norm_d(on.x, on.y, on.z, oldp.x, oldp.y, oldp.z) <= R ||
                      l3norm_d(on.x, on.y, on.z, newp.x, newp.y, newp.z) <= R;
            coll_agent |= hit;
        };
        // Shortcut: this step's neighbour list already holds EVERY agent within reach (r_i + r_max + `reach`) when it is
        // not saturated and reach stays inside neighbor_dist, or (true-KNN lists) when its farthest entry is an agent
        // beyond that reach.
        bool from_list = false;
        int ncnt = -1;
        const double my_reach = oldp.w + A.max_radius + reach;
        if (A.max_radius > 0.0) {
            ncnt = A.s_nbr_cnt[s];
            if (ncnt >= 0 && !A.s_qcoll[s]) {
                if (ncnt < KNBR) from_list = (my_reach + 0.01) * (my_reach + 0.01) < A.nbr_dist_sq;
                else if (A.last_nbr_mode == SCAB_NBR_KNN) {
                    const int e = A.s_nbr[(size_t)s * KNBR + KNBR - 1];
                    if (e < A.n) {
                        const double4 oo = A.s_posr[e];
                        const double ex = oo.x - oldp.x, ey = oo.y - oldp.y, ez = oo.z - oldp.z;
                        const double lim = my_reach + 0.1;
                        from_list = ex * ex + ey * ey + ez * ez > lim * lim;
                    }
                }
            }
        }
        if (from_list) {
            // the list is sorted by distance (both modes insert in order): without obstacle entries mixed in, the
            // first agent beyond reach ends the walk -- typically after one or two entries
            const double brk = my_reach + 0.01;
            const bool sorted_agents = (A.m == 0);
            for (int k = 0; k < ncnt; k++) {
                const int e = A.s_nbr[(size_t)s * KNBR + k];
                if (e >= A.n) continue;
                if (sorted_agents) {
                    const double4 oo = A.s_posr[e];
                    const double ex = oo.x - oldp.x, ey = oo.y - oldp.y, ez = oo.z - oldp.z;
                    if (ex * ex + ey * ey + ez * ez > brk * brk) break;
                }
                test_pair(e);
            }
        } else {
            int cx, cy, cz;
            cell_of(A.g, oldp.x, oldp.y, oldp.z, cx, cy, cz);
            for (int dz = -1; dz <= 1; dz++) {
                int z = cz + dz;
                if (z < 0 || z >= A.g.nz) continue;
                for (int dy = -1; dy <= 1; dy++) {
                    int y = cy + dy;
                    if (y < 0 || y >= A.g.ny) continue;
                    int row = (z * A.g.ny + y) * A.g.nx;
                    int beg = A.cell_start[row + max(cx - 1, 0)];
                    int end = A.cell_start[row + min(cx + 1, A.g.nx - 1) + 1];
                    for (int j = beg; j < end; j++) {
                        if (j == s) continue;
                        test_pair(j);
                    }
                }
            }
        }
        if (coll_obst) st |= SCAB_ST_COLLISION;
        if (coll_agent && !was_goal) st |= SCAB_ST_COLLISION;
        if (A.st.total_dist[id] > A.st.max_run_dist[id]) st |= SCAB_ST_TIMEOUT;   // :77-80
        // is_done (:51-59)
        const double *g = &A.st.goal[3 * id];
        if (l3norm_d(newp.x, newp.y, newp.z, g[0], g[1], g[2]) <= A.near_goal) st |= SCAB_ST_AT_GOAL;
        if (st & (SCAB_ST_AT_GOAL | SCAB_ST_COLLISION | SCAB_ST_TIMEOUT)) st |= SCAB_ST_RUN_DONE;
        A.st.status[id] = st;
        not_done = !(st & SCAB_ST_RUN_DONE);
        collided = (st & SCAB_ST_COLLISION) != 0;
    }
    const unsigned m1 = __ballot_sync(0xffffffffu, not_done), m2 = __ballot_sync(0xffffffffu, collided);
    if ((threadIdx.x & 31) == 0) {
        if (m1) atomicAdd(&A.counters[1], (unsigned long long)__popc(m1));
        if (m2) atomicAdd(&A.counters[2], (unsigned long long)__popc(m2));
    }
}

cudaError_t launch_status(const StepArgs &A, cudaStream_t s)
{
    if (A.n <= 0) return cudaSuccess;
    int nb = (A.n + 127) / 128;
    k_status<<<nb, 128, 0, s>>>(A);
    return cudaGetLastError();
}
