// scab_select_fast.cuh -- the common-case candidate search: EIGHT lanes per agent, four agents per warp.
// (Included by scab_policy_sampled.cu after the shared helpers.)
//
// k_select (one warp per agent, scab_policy_sampled.cu) leaves most lanes idle outside the cone loop: per
// agent only ~75 of the 513 candidates survive the posture test and ~45 stay suitable, while the scalar
// prologue/epilogue and the reductions are paid once per WARP.  Here a group of 8 lanes owns an agent, so a
// warp instruction advances four agents.  Groups run independently (every warp primitive uses the group's
// 8-lane mask).  Anything outside the common case -- no suitable candidate (scaPolicy.py:224-238), more
// surviving directions than DIR_CAP, keys beyond 22 bits -- is appended to a work list and handled by the
// general kernel k_select afterwards, so both kernels together cover every agent exactly once.
//
// Per group:  constraints (2 per lane)
//             -> A: posture test per DIRECTION (both shells share cos(angle)) from a lookup table over the direction of vA:
//                   per bin of an octahedral map, the directions that pass for EVERY vA of the bin ("sure") and the ones
//                   that may ("maybe"; only these ~10 of 256 are tested)
//             -> tier 1: the candidates within tier_delta of the one closest to v_pref (squared fp32 distances)
//             -> B: cone tests per direction, the two shells share the dot products p.u and h.u
//             -> C: keys (fp64, decimal-rounded, packed key<<10|c) + REDUX selection / shunting.
// If tier 1 does not pin the reference's answer (1 % of the agents) the agent goes to the work list as well: the general
// kernel examines every candidate.
#pragma once

#define FG_LANES 8
#define FG_PER_WARP 4
#define FG_WARPS 4        // warps per CTA
#define DIR_CAP 64        // surviving directions per agent (a pi/4 cap holds ~38 of 256)
#define LUT_B SCAB_POSTURE_LUT_B
#define LIVE_CAP 136      // 2 * DIR_CAP + v_pref, rounded up
#define VPREF_ENTRY 0x3ff
#ifndef LIST_EXTRA_DELTA
#define LIST_EXTRA_DELTA 0.15f    // second pass: tier width + 0.15 (measured on B200, policy stage at 10^6 agents: +0.05 1.259 ms,
                                  // +0.1 1.245, +0.15 1.244, +0.25 1.247, +0.4 1.262, +0.55 1.269 -- a narrower second tier leaves more
                                  // agents to the general kernel, whose first wave of 2960 costs the same 20-36 us however full it is)
#endif
#define LIST_EXTRA_DELTA_WIDE 0.4f // ... unless the list is long (a dense scene: BASELINE config 4 lists 93 % of its agents and the narrow tier
#define LIST_NARROW_MAX 24000     // would pass 90 % of them on): beyond ~8 waves' worth of leftovers the second pass examines tier + 0.4
#define LIST_MIN_AGENTS 3000      // second pass: shorter lists go straight to the general kernel
#ifndef FG_LIST_CTAS
#define FG_LIST_CTAS 5            // second pass: resident CTAs per SM it is compiled for.  6 (80 registers, a few spills) paid off while the
#endif                            // pass was one long wave at tier + 0.4; with the narrow tier 5 (96 registers, no spills) is 0.6 % faster
#ifndef FG_MIN_CTAS
#define FG_MIN_CTAS 5   // resident CTAs per SM the kernel is compiled for (96 registers): 1.39 -> 1.27 ms for the policy stage on B200; 6 spills
#endif
// tier 1 = candidates within StepArgs::tier_delta (default 0.28, SCAB_TIER_DELTA) of the one closest to v_pref

// fp64 copies of the constraints for the guard-band recheck.  Only the ORCA planes keep them in shared memory;
// a cone is cheap to rebuild from the neighbour record (recheck_cone64), and without these 896 bytes a fifth CTA
// fits on the SM.
struct Exact64 {
    double dpx[KNBR], dpy[KNBR], dpz[KNBR];
    double dhx[KNBR], dhy[KNBR], dhz[KNBR];
};
struct NoExact64 {};

template <bool WITH64>
struct GroupSmemT : public std::conditional<WITH64, Exact64, NoExact64>::type {
    float4 cp[KNBR];
    float4 ch[KNBR];
    float2 cg[KNBR];
    int cidx[KNBR];
    unsigned short act[LIVE_CAP];    // tier-1 candidates: n | shell << 12 ; n == VPREF_ENTRY = v_pref itself
    unsigned int kc[LIVE_CAP];       // live candidates: c, later key << 10 | c
    unsigned int dvu[LIVE_CAP];      // l3norm(v, vA) in units of 1e-5
    int pad[4];                      // the four groups of a warp start in different banks
};

// the guard-band recheck of one cone (scaPolicy.py:47-60,180 + util.py:30-41) rebuilt from the neighbour record
__device__ __noinline__ bool recheck_cone64(const StepArgs &A, int s, int k, const Agent &a, double x, double y, double z)
{
    const int e = A.s_nbr[(size_t)s * KNBR + k];
    double4 o;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    bool at_goal = true;
    if (e < A.n) {
        o = A.s_posr[e];
        const float4 ov = A.s_vel[e];
        ox = ov.x; oy = ov.y; oz = ov.z;
        at_goal = (__float_as_int(ov.w) & SCAB_ST_AT_GOAL) != 0;
    } else if (e < A.n + A.m) {
        o = A.o_posr[e - A.n];
    } else {
        o = A.o_posr_by_id[e - A.n - A.m];
    }
    const double R = __dadd_rn(__dadd_rn(o.w, 0.05), __dadd_rn(a.r, 0.05));
    double p0x = a.px, p0y = a.py, p0z = a.pz;
    if (!at_goal) {
        p0x = __dadd_rn(a.px, (double)__fmul_rn(0.5f, __fadd_rn(ox, a.vx)));
        p0y = __dadd_rn(a.py, (double)__fmul_rn(0.5f, __fadd_rn(oy, a.vy)));
        p0z = __dadd_rn(a.pz, (double)__fmul_rn(0.5f, __fadd_rn(oz, a.vz)));
    }
    const double dx = __dsub_rn(__dadd_rn(x, a.px), p0x), dy = __dsub_rn(__dadd_rn(y, a.py), p0y), dz = __dsub_rn(__dadd_rn(z, a.pz), p0z);
    return is_intersect64_fast(__dsub_rn(o.x, a.px), __dsub_rn(o.y, a.py), __dsub_rn(o.z, a.pz), R, dx, dy, dz);
}

template <bool LIST>
__device__ __forceinline__ void push_slow(const StepArgs &A, int s)
{   // first pass -> the list of the second pass; second pass -> the list of the general kernel
    const int w = atomicAdd(LIST ? A.slow_count2 : A.slow_count, 1);
    (LIST ? A.slow_list2 : A.slow_list)[w] = s;
}

// Bin of the posture lookup table for a (non-zero) vector: octahedral map of its direction, LUT_B x LUT_B bins.  The host
// builds the table with the inverse map (scab_api.cu make_posture_lut) and a margin that covers fp32 rounding here.
__device__ __forceinline__ int posture_bin(float x, float y, float z)
{
    const float inv = 1.f / (fabsf(x) + fabsf(y) + fabsf(z));
    float ox = x * inv, oy = y * inv;
    if (z < 0.f) {
        const float tx = (1.f - fabsf(oy)) * (ox >= 0.f ? 1.f : -1.f);
        oy = (1.f - fabsf(ox)) * (oy >= 0.f ? 1.f : -1.f);
        ox = tx;
    }
    const int ix = min(max((int)((ox + 1.f) * (0.5f * LUT_B)), 0), LUT_B - 1);
    const int iy = min(max((int)((oy + 1.f) * (0.5f * LUT_B)), 0), LUT_B - 1);
    return iy * LUT_B + ix;
}

// LIST = false: every agent, tier 1 only.  LIST = true: the agents the first pass could not pin, with a wider tier
// (tier_delta + LIST_EXTRA_DELTA); a short list skips this pass -- the general kernel finishes up to one wave of agents in
// one warp's latency (~35 us), which a second 8-lane pass cannot beat.
template <int POLICY, bool LIST>
__global__ void __launch_bounds__(FG_WARPS * 32, (POLICY == SCAB_POLICY_ORCA_SAMPLED) ? 4 : (LIST ? FG_LIST_CTAS : FG_MIN_CTAS)) k_select_fast(const __grid_constant__ StepArgs A)
{
    const bool full_search = A.full_search != 0;
    constexpr bool IS_ORCA = (POLICY == SCAB_POLICY_ORCA_SAMPLED);
    constexpr bool SHUNT = (POLICY == SCAB_POLICY_SCA || POLICY == SCAB_POLICY_SRVO3D);
    typedef GroupSmemT<IS_ORCA> GroupSmem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GroupSmem *gs = reinterpret_cast<GroupSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane >> 3, l8 = lane & 7;
    const unsigned gmask = 0xffu << (8 * g);
    const unsigned lt8 = (1u << l8) - 1;
    GroupSmem &W = gs[wid * FG_PER_WARP + g];
    const int groups_total = gridDim.x * FG_WARPS * FG_PER_WARP;
    unsigned long long n_recheck = 0, n_tier2 = 0;
    // fp32 direction tables, once per CTA (6 KB): conflict-free float4 reads in phases A and B
    const float4 *s_dir256 = A.dir256f, *s_dir128 = A.dir128f;       // (second pass: straight from global memory / L1)
    if (!LIST) {
        float4 *t256 = reinterpret_cast<float4 *>(smem_raw + sizeof(GroupSmem) * FG_WARPS * FG_PER_WARP);
        for (int i = threadIdx.x; i < 256; i += blockDim.x) t256[i] = A.dir256f[i];
        for (int i = threadIdx.x; i < 128; i += blockDim.x) t256[256 + i] = A.dir128f[i];
        s_dir256 = t256; s_dir128 = t256 + 256;
        __syncthreads();
    }

    const int n_work = LIST ? min(*A.slow_count, A.n) : policy_items(A);
    const float tier_delta = A.tier_delta + (LIST ? (n_work <= LIST_NARROW_MAX ? LIST_EXTRA_DELTA : LIST_EXTRA_DELTA_WIDE) : 0.f);
    if (LIST && n_work < LIST_MIN_AGENTS) {      // hand the whole list on (uniform branch: every thread sees the same count)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_work; i += gridDim.x * blockDim.x) A.slow_list2[i] = A.slow_list[i];
        if (blockIdx.x == 0 && threadIdx.x == 0) *A.slow_count2 = n_work;
        return;
    }
    for (int wi = (blockIdx.x * FG_WARPS + wid) * FG_PER_WARP + g; wi < n_work; wi += groups_total) {
        const int s = LIST ? A.slow_list[wi] : policy_item(A, wi);
        const float4 mv = A.s_vel[s];
        const int flags = __float_as_int(mv.w);
        if (!LIST && A.row_count < 0) {   // L2 prefetch of this group's next agent (header lines) and of its neighbour records
            const int s1 = s + groups_total;
            if (s1 < A.n) {
                if (l8 == 0) prefetch_l2(&A.s_vel[s1]);
                if (l8 == 1) prefetch_l2(&A.s_vdes[s1]);
                if (l8 == 2) prefetch_l2(&A.s_posr[s1]);
                if (l8 == 3) prefetch_l2(&A.s_nbr[(size_t)s1 * KNBR]);
            }
        }
        if (flags & FL_INACTIVE) continue;
        const double4 vd = A.s_vdes[s];
        if (vd.w < 0.0) {   // zero-velocity shortcut: vA_post = 0.3 v_pref (scaPolicy.py:38)
            if (l8 == 0)
                A.s_vpost[s] = make_double4(__dmul_rn(0.3, vd.x), __dmul_rn(0.3, vd.y), __dmul_rn(0.3, vd.z),
                                            __longlong_as_double((long long)SCAB_TAG_FIRST));
            continue;
        }
        const double4 me = A.s_posr[s];
        Agent a;
        a.px = me.x; a.py = me.y; a.pz = me.z; a.r = me.w;
        a.vx = mv.x; a.vy = mv.y; a.vz = mv.z;
        a.pref = vd.w;
        a.nvA = (double)norm3f_np(a.vx, a.vy, a.vz);
        const bool zaxis = (POLICY == SCAB_POLICY_SCA) && (flags & SCAB_ST_ZAXIS);
        a.N = zaxis ? 128 : 256;
        a.dir = zaxis ? A.dir128 : A.dir256;
        a.dirf = zaxis ? A.dir128f : A.dir256f;
        {
            double stepv = __dsub_rn(a.pref, 0.5);
            double second = __dadd_rn(0.5, stepv);
            a.shell1 = __dadd_rn(0.5, __dsub_rn(second, 0.5));
        }
        a.vdx = vd.x; a.vdy = vd.y; a.vdz = vd.z;
        const int nc = A.s_nbr_cnt[s];
        const float r1f = (float)a.shell1;
        const float vmaxf = 1.0001f * (float)fmax(fmax(0.5, fabs(a.shell1)), sqrt(a.vdx * a.vdx + a.vdy * a.vdy + a.vdz * a.vdz));
        const float fnvA = (float)a.nvA;
        const float ball_r = 0.70711f * vmaxf;
        const float inv_nvA = 1.f / fnvA;
        const float bmx = ball_r * a.vx * inv_nvA, bmy = ball_r * a.vy * inv_nvA, bmz = ball_r * a.vz * inv_nvA;
        const float vdx = (float)a.vdx, vdy = (float)a.vdy, vdz = (float)a.vdz;
        // my two neighbours (k = l8, l8 + 8): indices now, records towards L1 while the posture / tier phases run
        const int e0 = (l8 < nc) ? A.s_nbr[(size_t)s * KNBR + l8] : -1;
        const int e1 = (l8 + FG_LANES < nc) ? A.s_nbr[(size_t)s * KNBR + FG_LANES + l8] : -1;
        if (e0 >= 0 && e0 < A.n) { prefetch_l1(&A.s_posr[e0]); prefetch_l1(&A.s_vel[e0]); }
        if (e1 >= 0 && e1 < A.n) { prefetch_l1(&A.s_posr[e1]); prefetch_l1(&A.s_vel[e1]); }
        // the posture table row of this agent: requested before the constraint gathers, consumed after them
        const float4 *tab = zaxis ? s_dir128 : s_dir256;
        unsigned sure = 0u, maybe = zaxis ? 0xffffu : 0xffffffffu;      // degenerate vA: test every direction
        if (fnvA > 1e-30f) {
            const unsigned *row = (zaxis ? A.plut128 : A.plut256) + (size_t)posture_bin(a.vx, a.vy, a.vz) * 16;
            sure = row[l8]; maybe = row[8 + l8];
        }

        // ------------------------------------------------------------ A: posture per direction
        // Lane l8 owns directions n = 8 i + l8 (bit i).  `sure` directions pass for every vA of the table bin; `maybe`
        // directions get the fp32 test, and the reference's fp64 expression inside its 2e-6 guard band.
        unsigned m0, m1;
        {
            const float thr = 0.70710678f * fnvA;
            const float band = 2e-6f * fnvA;
            m0 = sure; m1 = sure;
#pragma unroll 1
            for (unsigned b = maybe; b; b &= b - 1) {
                const int i = __ffs(b) - 1;
                const float4 u = tab[FG_LANES * i + l8];
                const float dt = fmaf(a.vz, u.z, fmaf(a.vy, u.y, a.vx * u.x));
                bool p0 = dt >= thr, p1 = p0;
                if (fabsf(dt - thr) <= band) {
                    double x, y, z;
                    cand64(a, FG_LANES * i + l8, x, y, z);
                    p0 = posture_ok(a, A.time_step, x, y, z);
                    cand64(a, a.N + FG_LANES * i + l8, x, y, z);
                    p1 = posture_ok(a, A.time_step, x, y, z);
                }
                m0 |= (p0 ? 1u : 0u) << i;
                m1 |= (p1 ? 1u : 0u) << i;
            }
            // altitude: pA.z + dt v.z >= 0 (util.py:19) can only fail within dt * max|v| of the ground
            const double vmax = fmax(fmax(0.5, fabs(a.shell1)), fabs(a.vdz));
            if (!(a.pz > A.time_step * vmax + 1e-9)) {
#pragma unroll 1
                for (unsigned b = m0 | m1; b; b &= b - 1) {
                    const int i = __ffs(b) - 1;
                    const double uz = a.dir[3 * (FG_LANES * i + l8) + 2];
                    if (!(__dadd_rn(a.pz, __dmul_rn(A.time_step, __dmul_rn(0.5, uz))) >= 0.0)) m0 &= ~(1u << i);
                    if (!(__dadd_rn(a.pz, __dmul_rn(A.time_step, __dmul_rn(a.shell1, uz))) >= 0.0)) m1 &= ~(1u << i);
                }
            }
        }
        const bool pv = posture_ok(a, A.time_step, a.vdx, a.vdy, a.vdz);    // v_pref itself, the last candidate (scaPolicy.py:206)
        const int ndir = __reduce_add_sync(gmask, __popc(m0 | m1)) + (pv ? 1 : 0);
        if (ndir >= DIR_CAP) {           // not the common case: leave it to the general kernel
            if (l8 == 0) push_slow<LIST>(A, s);
            continue;
        }

        // ------------------------------------------------------------ tier 1
        // The answer only depends on the head of the key-sorted suitable list and (with shunting) on the short prefix
        // before the first element that breaks it, so only the candidates within tier_delta of the one closest to v_pref
        // (APPROXIMATE fp32 distances) are examined.  If they pin the result -- head (and prefix-breaking element) found
        // with exact keys below the tier boundary minus a margin covering the fp32 key error -- the other ~90 % never
        // matter; otherwise the agent goes to the general kernel.  full_search (scab_set_option) examines everything
        // and yields the exact n_suit of the reference.
        // |r u - v_pref|^2 = r^2 - 2 r u.v_pref + |v_pref|^2 for a unit direction u (|u|^2 = 1 to 2e-7: these distances
        // only order the candidates, to 1e-3)
        const float vdvd = fmaf(vdz, vdz, fmaf(vdy, vdy, vdx * vdx));
        const float c0 = 0.25f + vdvd, c1 = fmaf(r1f, r1f, vdvd), r2 = 2.f * r1f;
        float T1 = 3.0e38f;
        if (!full_search) {
            float kmin = 0.f;                        // v_pref, when it passes the posture test, is at distance 0 from itself
            if (!pv) {
                kmin = 3.0e38f;
#pragma unroll 1
                for (unsigned b = m0 | m1; b; b &= b - 1) {
                    const int i = __ffs(b) - 1;
                    const float4 u = tab[FG_LANES * i + l8];
                    const float uv = fmaf(u.z, vdz, fmaf(u.y, vdy, u.x * vdx));
                    if ((m0 >> i) & 1u) kmin = fminf(kmin, fmaxf(c0 - uv, 0.f));               // squared distances:
                    if ((m1 >> i) & 1u) kmin = fminf(kmin, fmaxf(fmaf(-r2, uv, c1), 0.f));    // one sqrt after the min
                }
                kmin = sqrtf(__uint_as_float(__reduce_min_sync(gmask, __float_as_uint(kmin))));   // non-negative floats order as uints
            }
            T1 = kmin + tier_delta;
        }
        const float T1sq = T1 * T1;      // (+inf under full_search)
        // exact keys strictly below this are certainly smaller than any key outside tier 1 (fp32 key error <= 1e-3)
        const unsigned tier_units = full_search ? 0xffffffffu : (unsigned)fmaxf((T1 - 1.1e-3f) * 1.0e5f, 0.f);
        int nact = 0;
        {
            unsigned a0 = 0, a1 = 0;             // per lane: direction 8 i + l8 -> bit i of the shell-0 / shell-1 masks
#pragma unroll 1
            for (unsigned b = m0 | m1; b; b &= b - 1) {
                const int i = __ffs(b) - 1;
                const float4 u = tab[FG_LANES * i + l8];
                const float uv = fmaf(u.z, vdz, fmaf(u.y, vdy, u.x * vdx));
                if (c0 - uv <= T1sq) a0 |= m0 & (1u << i);                 // |v - v_pref|^2, fp32
                if (fmaf(-r2, uv, c1) <= T1sq) a1 |= m1 & (1u << i);
            }
            // one entry per CANDIDATE (tier 1 around a v_pref of length pref_speed is almost all outer shell: testing the
            // two shells of a direction together would do half of its arithmetic for nothing)
            const int cnt = __popc(a0) + __popc(a1);
            int pre = cnt;
#pragma unroll
            for (int o = 1; o < FG_LANES; o <<= 1) {
                const int t = __shfl_up_sync(gmask, pre, o, FG_LANES);
                if (l8 >= o) pre += t;
            }
            nact = __shfl_sync(gmask, pre, FG_LANES - 1, FG_LANES);
            int slot = pre - cnt;
            while (a0) {
                const int i = __ffs(a0) - 1;
                a0 &= a0 - 1;
                W.act[slot++] = (unsigned short)(FG_LANES * i + l8);
            }
            while (a1) {
                const int i = __ffs(a1) - 1;
                a1 &= a1 - 1;
                W.act[slot++] = (unsigned short)((FG_LANES * i + l8) | (1u << 12));
            }
            if (pv) {
                if (l8 == 0) W.act[nact] = (unsigned short)VPREF_ENTRY;
                nact++;
            }
        }
        __syncwarp(gmask);

        // ------------------------------------------------------------ constraints: lane l8 builds k = l8, l8 + 8
        // A constraint is kept only if its forbidden region comes within 1e-2 of BOTH balls that enclose every candidate
        // this kernel will test: the posture ball (all posture-feasible velocities) and the tier ball B(v_pref, T1).
        int nce = 0;
        const float rho_t = T1 + 2e-3f;          // fp32 tier distances are good to 1e-3 (+inf under full_search)
        for (int kb = 0; kb < nc; kb += FG_LANES) {
            const int k = kb + l8;
            float4 cP = make_float4(0.f, 0.f, 0.f, 0.f), cH = cP;
            float2 cG = make_float2(0.f, 0.f);
            bool keep = false;
            if (k < nc) {
                const int e = kb ? e1 : e0;
                double4 o;
                float ox = 0.f, oy = 0.f, oz = 0.f;
                bool at_goal = true;
                if (e < A.n) {
                    o = A.s_posr[e];
                    const float4 ov = A.s_vel[e];
                    ox = ov.x; oy = ov.y; oz = ov.z;
                    at_goal = (__float_as_int(ov.w) & SCAB_ST_AT_GOAL) != 0;
                } else if (e < A.n + A.m) {
                    o = A.o_posr[e - A.n];
                } else {
                    o = A.o_posr_by_id[e - A.n - A.m];
                }
                const double R = IS_ORCA ? __dadd_rn(__dadd_rn(a.r, 0.05), __dadd_rn(o.w, 0.05))
                                         : __dadd_rn(__dadd_rn(o.w, 0.05), __dadd_rn(a.r, 0.05));
                const double rx = __dsub_rn(o.x, a.px), ry = __dsub_rn(o.y, a.py), rz = __dsub_rn(o.z, a.pz);
                if (!IS_ORCA) {   // scaPolicy.py:47-60
                    float hx = 0.f, hy = 0.f, hz = 0.f;
                    if (!at_goal) {
                        hx = __fmul_rn(0.5f, __fadd_rn(ox, a.vx));
                        hy = __fmul_rn(0.5f, __fadd_rn(oy, a.vy));
                        hz = __fmul_rn(0.5f, __fadd_rn(oz, a.vz));
                    }
                    const double pp = dot3(rx, ry, rz, rx, ry, rz);
                    const float fpx = (float)rx, fpy = (float)ry, fpz = (float)rz;
                    const float hh = fmaf(hz, hz, fmaf(hy, hy, hx * hx));
                    cP = make_float4(fpx, fpy, fpz, fmaf(fpz, hz, fmaf(fpy, hy, fpx * hx)));
                    cH = make_float4(hx, hy, hz, hh);
                    const float vh_max = vmaxf + sqrtf(hh);
                    const float c2f = (float)fmax(pp - R * R, 0.0);
                    cG = make_float2(c2f, 2e-5f * (float)pp * vh_max * vh_max);
                    // distance from a ball centre (relative to the apex h) to the cone {w: p.w > |w| c} -- approximate
                    // reciprocal square roots are fine here: the margin is 1e-2
                    const float ppf = (float)pp, Rf = (float)R;
                    const float invd = rsqrtf(ppf);
                    const float sn = fminf(Rf * invd, 1.f), cs = (c2f > 0.f ? c2f * rsqrtf(c2f) : 0.f) * invd;
                    auto clear_of = [&](float cx, float cy, float cz, float rad) {
                        const float qx = cx - hx, qy = cy - hy, qz = cz - hz;
                        const float qa = (qx * fpx + qy * fpy + qz * fpz) * invd;
                        const float q2 = fmaxf(qx * qx + qy * qy + qz * qz - qa * qa, 0.f);
                        const float qr = q2 > 0.f ? q2 * rsqrtf(q2) : 0.f;
                        return qr * cs - qa * sn > rad + 1e-2f;
                    };
                    keep = !(ppf > 1.02f * Rf * Rf) || !(clear_of(bmx, bmy, bmz, ball_r) || clear_of(vdx, vdy, vdz, rho_t));
                } else {          // orca3dPolicy.py:57-107
                    const double vrx = (double)__fsub_rn(a.vx, ox), vry = (double)__fsub_rn(a.vy, oy), vrz = (double)__fsub_rn(a.vz, oz);
                    const double distSq = dot3(rx, ry, rz, rx, ry, rz);
                    const double RSq = __dmul_rn(R, R);
                    double nx, ny, nz, ux, uy, uz;
                    if (distSq > RSq) {
                        const double invT = A.inv_time_horizon;
                        double wx = __dsub_rn(vrx, __dmul_rn(invT, rx)), wy = __dsub_rn(vry, __dmul_rn(invT, ry)), wz = __dsub_rn(vrz, __dmul_rn(invT, rz));
                        double wl2 = dot3(wx, wy, wz, wx, wy, wz);
                        double dp = dot3(wx, wy, wz, rx, ry, rz);
                        if (dp < 0.0 && __dmul_rn(dp, dp) > __dmul_rn(RSq, wl2)) {
                            double wl = __dsqrt_rn(wl2);
                            nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                            double sc = __dsub_rn(__dmul_rn(R, invT), wl);
                            ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                        } else {
                            double difSq = __dsub_rn(distSq, RSq);
                            double dpr = dot3(rx, ry, rz, vrx, vry, vrz);
                            double cx = __dsub_rn(__dmul_rn(ry, vrz), __dmul_rn(rz, vry));
                            double cy = __dsub_rn(__dmul_rn(rz, vrx), __dmul_rn(rx, vrz));
                            double cz = __dsub_rn(__dmul_rn(rx, vry), __dmul_rn(ry, vrx));
                            double wwSq = __ddiv_rn(dot3(cx, cy, cz, cx, cy, cz), difSq);
                            double pl = __dsqrt_rn(distSq);
                            double t = __ddiv_rn(__dadd_rn(__ddiv_rn(dpr, pl), __dmul_rn(__dsqrt_rn(wwSq), __ddiv_rn(R, pl))), pl);
                            double wx = __dsub_rn(vrx, __dmul_rn(t, rx)), wy = __dsub_rn(vry, __dmul_rn(t, ry)), wz = __dsub_rn(vrz, __dmul_rn(t, rz));
                            double wl = __dsqrt_rn(dot3(wx, wy, wz, wx, wy, wz));
                            nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                            double sc = __dsub_rn(__dmul_rn(R, t), wl);
                            ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                        }
                    } else {
                        const double invDt = A.inv_time_step;
                        double wx = __dsub_rn(vrx, __dmul_rn(invDt, rx)), wy = __dsub_rn(vry, __dmul_rn(invDt, ry)), wz = __dsub_rn(vrz, __dmul_rn(invDt, rz));
                        double wl = __dsqrt_rn(dot3(wx, wy, wz, wx, wy, wz));
                        nx = __ddiv_rn(wx, wl); ny = __ddiv_rn(wy, wl); nz = __ddiv_rn(wz, wl);
                        double sc = __dsub_rn(__dmul_rn(R, invDt), wl);
                        ux = __dmul_rn(sc, nx); uy = __dmul_rn(sc, ny); uz = __dmul_rn(sc, nz);
                    }
                    const double qx = __dadd_rn((double)a.vx, __dmul_rn(0.5, ux)), qy = __dadd_rn((double)a.vy, __dmul_rn(0.5, uy)),
                                 qz = __dadd_rn((double)a.vz, __dmul_rn(0.5, uz));
                    if constexpr (IS_ORCA) {
                        W.dpx[k] = qx; W.dpy[k] = qy; W.dpz[k] = qz;
                        W.dhx[k] = nx; W.dhy[k] = ny; W.dhz[k] = nz;
                    }
                    cP = make_float4((float)nx, (float)ny, (float)nz, (float)(nx * qx + ny * qy + nz * qz));
                    cG = make_float2(0.f, 1e-5f * (vmaxf + (float)sqrt(qx * qx + qy * qy + qz * qz)));
                    // half-space n.v >= n.point: the smallest n.v over a ball is n.centre - radius
                    keep = !((cP.x * bmx + cP.y * bmy + cP.z * bmz) - ball_r - cP.w > 1e-2f ||
                             (cP.x * vdx + cP.y * vdy + cP.z * vdz) - rho_t - cP.w > 1e-2f);
                }
            }
            const unsigned km = (__ballot_sync(gmask, keep) >> (8 * g)) & 0xffu;
            if (keep) {
                const int slot = nce + __popc(km & lt8);
                W.cp[slot] = cP; W.ch[slot] = cH; W.cg[slot] = cG; W.cidx[slot] = k;
            }
            nce += __popc(km);
        }
        __syncwarp(gmask);       // the packed constraints (and the tier-1 entries above) are read by every lane of the group

        // ---- B: cone tests, one candidate per lane and round
        int nlive = 0, best = -1;
        unsigned kfail = 0xffffffffu;
        bool overflow = false;
        for (int jb = 0; jb < nact; jb += FG_LANES) {
            const int ja = jb + l8;
            bool ok = false;
            int c = 0;
            if (ja < nact) {
                const unsigned ae = W.act[ja];
                const int n = ae & 0x3ff;
                const bool isv = (n == VPREF_ENTRY), outer = (ae >> 12) != 0;
                const float4 u = isv ? make_float4(vdx, vdy, vdz, 0.f) : tab[n];
                const float uu = fmaf(u.z, u.z, fmaf(u.y, u.y, u.x * u.x));
                const float ra = isv ? 1.f : (outer ? r1f : 0.5f);
                c = isv ? 2 * a.N : (outer ? a.N + n : n);
                ok = true;
                for (int kk = 0; kk < nce && ok; kk++) {
                    const float4 P = W.cp[kk];
                    const float2 G = W.cg[kk];
                    const float pa = fmaf(P.z, u.z, fmaf(P.y, u.y, P.x * u.x));
                    float df;
                    bool in;
                    if (!IS_ORCA) {
                        const float4 H = W.ch[kk];
                        const float hb = fmaf(H.z, u.z, fmaf(H.y, u.y, H.x * u.x));
                        const float dt = fmaf(ra, pa, -P.w);
                        const float vv = fmaf(-2.f * ra, hb, ra * ra * uu) + H.w;
                        df = fmaf(dt, dt, -vv * G.x);
                        in = (dt > 0.f) && (df > 0.f);
                    } else {
                        df = fmaf(ra, pa, -P.w);          // n.v - n.q: violated iff negative
                        in = !(df >= 0.f);
                    }
                    if (fabsf(df) <= G.y) {              // inside the fp32 guard band: the reference's fp64 expression
                        const int k = W.cidx[kk];
                        n_recheck++;
                        double x, y, z;
                        cand64(a, c, x, y, z);
                        if constexpr (!IS_ORCA) {
                            in = recheck_cone64(A, s, k, a, x, y, z);
                        } else {
                            double val = dot3(__dsub_rn(x, W.dpx[k]), __dsub_rn(y, W.dpy[k]), __dsub_rn(z, W.dpz[k]),
                                              W.dhx[k], W.dhy[k], W.dhz[k]);
                            in = !(val >= 0.0);
                        }
                    }
                    if (in) ok = false;
                }
            }
            // compact the surviving candidates of this round (generation index c) into the live list
            const unsigned b0 = (__ballot_sync(gmask, ok) >> (8 * g)) & 0xffu;
            if (ok) W.kc[nlive + __popc(b0 & lt8)] = c;
            nlive += __popc(b0);
        }
        __syncwarp(gmask);

        // ---- C: exact keys of the live candidates
        for (int jb = 0; jb < nlive; jb += FG_LANES) {
            const int j = jb + l8;
            if (j < nlive) {
                const int c = (int)W.kc[j];
                double x, y, z;
                cand64(a, c, x, y, z);
                const double dist = __dsqrt_rn(sumsq3(__dsub_rn(x, a.vdx), __dsub_rn(y, a.vdy), __dsub_rn(z, a.vdz)));
                const double ku = py_round5_units(dist);          // l3norm(v, v_pref) in units of 1e-5
                if (!(ku < 4194304.0)) overflow = true;
                W.kc[j] = ((unsigned)ku << 10) | (unsigned)c;
                if (SHUNT) {
                    const double dva = __dsqrt_rn(sumsq3(__dsub_rn(x, (double)a.vx), __dsub_rn(y, (double)a.vy), __dsub_rn(z, (double)a.vz)));
                    W.dvu[j] = (unsigned int)py_round5_units(dva);
                }
            }
        }
        // keys beyond 22 bits, or nothing suitable in tier 1 (possibly "no suitable velocity" at all, scaPolicy.py:224-238)
        if (__any_sync(gmask, overflow) || nlive == 0) {
            if (l8 == 0) { push_slow<LIST>(A, s); if (nlive == 0 && !LIST) n_tier2++; }
            continue;
        }
        __syncwarp(gmask);

        // ---- head of the sorted list, and (shunting) the first element that breaks the prefix
        bool decided;
        {
            unsigned local_min = 0xffffffffu;
#pragma unroll 1
            for (int j = l8; j < nlive; j += FG_LANES) local_min = min(local_min, W.kc[j]);
            const unsigned head = __reduce_min_sync(gmask, local_min);
            best = (int)(head & 1023u);
            if (SHUNT) {
                const long long thr_units = (POLICY == SCAB_POLICY_SCA) ? 3000 : 10000;
                const double thr = (POLICY == SCAB_POLICY_SCA) ? 3e-2 : 1e-1;
                double x0, y0, z0;
                cand64(a, best, x0, y0, z0);
                const long long d0u = (long long)py_round5_units(
                    __dsqrt_rn(sumsq3(__dsub_rn(x0, (double)a.vx), __dsub_rn(y0, (double)a.vy), __dsub_rn(z0, (double)a.vz))));
                unsigned fmin_ = 0xffffffffu;
                for (int j = l8; j < nlive; j += FG_LANES) {
                    const long long diu = (long long)W.dvu[j];
                    const long long du = d0u > diu ? d0u - diu : diu - d0u;
                    bool within;
                    if (du < thr_units - 1) within = true;
                    else if (du > thr_units + 1) within = false;
                    else within = fabs(__dsub_rn(div1e5((double)d0u), div1e5((double)diu))) < thr;
                    if (!within) fmin_ = min(fmin_, W.kc[j]);
                }
                kfail = __reduce_min_sync(gmask, fmin_);
                decided = full_search || (kfail != 0xffffffffu && (kfail >> 10) < tier_units);
            } else {
                decided = full_search || ((head >> 10) < tier_units);
            }
        }
        if (!decided) {                   // tier 1 does not pin the answer: every candidate, in the general kernel
            if (l8 == 0) { push_slow<LIST>(A, s); if (!LIST) n_tier2++; }
            continue;
        }
        int nprefix = 0;
        if (SHUNT && nlive > 1) {
#pragma unroll 1
            for (int j = l8; j < nlive; j += FG_LANES) nprefix += (W.kc[j] < kfail) ? 1 : 0;
            nprefix = __reduce_add_sync(gmask, nprefix);
        }
        if (SHUNT && nprefix > 1) {      // a one-element prefix is its own min-phi / max-phi: best stays the head
            // scaPolicy.py:136-145: min-phi / max-phi over the prefix {(key, c) < kfail}; ties -> first in sorted order
            unsigned umin = 0xffffffffu, umax = 0u, kmin_ = 0xffffffffu, kmax_ = 0xffffffffu;
            for (int j = l8; j < nlive; j += FG_LANES) {
                const unsigned kcj = W.kc[j];
                if (kcj < kfail) {
                    double x, y, z;
                    cand64(a, (int)(kcj & 1023u), x, y, z);
                    double phi = (y >= 0) ? atan2(y, x) : __dadd_rn(2 * SCAB_PI, atan2(y, x));   // util.py:145-150
                    const unsigned pu = (unsigned)((long long)trunc(__dmul_rn(phi, 1e5)) + 1000000);
                    if (pu < umin || (pu == umin && kcj < kmin_)) { umin = pu; kmin_ = kcj; }
                    if (pu > umax || (pu == umax && kcj < kmax_) || kmax_ == 0xffffffffu) { umax = pu; kmax_ = kcj; }
                }
            }
            const unsigned gmin = __reduce_min_sync(gmask, umin), gmax = __reduce_max_sync(gmask, umax);
            const unsigned cmin = __reduce_min_sync(gmask, umin == gmin ? kmin_ : 0xffffffffu);
            const unsigned cmax = __reduce_min_sync(gmask, (umax == gmax && kmax_ != 0xffffffffu) ? kmax_ : 0xffffffffu);
            const double phi_min = div1e5((double)((long long)gmin - 1000000)), phi_max = div1e5((double)((long long)gmax - 1000000));
            best = (fabs(__dsub_rn(phi_max, phi_min)) <= SCAB_PI) ? (int)(cmin & 1023u) : (int)(cmax & 1023u);
        }
        if (l8 == 0) {
            double bx, by, bz;
            cand64(a, best, bx, by, bz);
            // n_suit is only known when every candidate was examined (full_search); otherwise report -1
            const long long ns = full_search ? (long long)nlive : -1ll;
            A.s_vpost[s] = make_double4(trunc5(bx), trunc5(by), trunc5(bz),
                                        __longlong_as_double((long long)SCAB_TAG_SUITABLE | (ns << 8)));
        }
        __syncwarp(gmask);
    }
    for (int o = 4; o > 0; o >>= 1) n_recheck += __shfl_xor_sync(gmask, n_recheck, o);
    if (l8 == 0 && n_recheck) atomicAdd(&A.counters[3], n_recheck);
    if (l8 == 0 && n_tier2) atomicAdd(&A.counters[5], n_tier2);
}
