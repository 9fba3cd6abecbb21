/*
 * scab200.h -- C ABI of libscab200.so: the B200-native batched collision-avoidance step.
 *
 * The reference (wuuya1/SCA, pure Python) has no FFI; its boundary is the duck-typed
 * policy/kd-tree interface MACAEnv drives once per step:
 *     mamp/envs/mampenv.py:28   self.kdTree.buildAgentTree()
 *     mamp/envs/mampenv.py:40   agent.policy.find_next_action(dict_comm, agent, self.kdTree)
 *     mamp/envs/mampenv.py:43   update_velocitie(agent, all_actions[i, :])
 *     mamp/envs/mampenv.py:46   self.check_agent_state(agent)
 *     mamp/envs/mampenv.py:24   self.is_done(NEAR_GOAL_THRESHOLD)
 * Each entry point below names the reference code it replaces.  sca_b200/mamp_compat.py is the
 * Python-side binding that plugs these into an unmodified MACAEnv.
 *
 * Conventions: every function returns 0 on success or a negative scab_status; the message of the last
 * failure is available from scab_last_error().  No C++ types, no exceptions cross the boundary.  All
 * kernels are launched asynchronously on the caller's stream (pass the cudaStream_t as void*; NULL =
 * legacy default stream).  One handle per device; a handle is not thread-safe; distinct handles are
 * independent.  State arrays are CALLER-OWNED device memory (PyTorch tensors in this repo): the library
 * never allocates, frees or keeps them alive -- the caller must keep them alive until scab_destroy or
 * the next scab_bind_state.
 */
#ifndef SCAB200_H
#define SCAB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCAB_MAX_NEIGHBORS 16 /* mamp/agents/agent.py:32 maxNeighbors */

typedef struct scab_handle scab_handle;

typedef enum {
    SCAB_OK = 0,
    SCAB_ERR_INVALID = -1,   /* bad argument / state not bound */
    SCAB_ERR_CUDA = -2,      /* CUDA runtime error (message in scab_last_error) */
    SCAB_ERR_CAPACITY = -3,  /* n exceeds the capacity given to scab_create */
    SCAB_ERR_NOMEM = -4
} scab_status;

/* Policy selector: one per reference policy class. */
typedef enum {
    SCAB_POLICY_SCA = 0,          /* mamp/policies/sca/scaPolicy.py:18 SCAPolicy */
    SCAB_POLICY_SRVO3D = 1,       /* mamp/policies/srvo3dPolicy.py:15 SRVO3DPolicy */
    SCAB_POLICY_RVO3D = 2,        /* mamp/policies/rvo3dPolicy.py:15 RVO3DPolicy */
    SCAB_POLICY_RVO3D_DUBINS = 3, /* mamp/policies/sca/rvo3dDubinsPolicy.py:18 RVO3dDubinsPolicy */
    SCAB_POLICY_ORCA_SAMPLED = 4, /* mamp/policies/orca3dPolicy.py:27 ORCA3DPolicy (run_orca.py imports this) */
    SCAB_POLICY_ORCA_LP = 5       /* mamp/policies/orca3dPolicyOfficial.py:27 ORCA3DPolicy (linearProgram3/4) */
} scab_policy;

/* Neighbour semantics. */
typedef enum {
    SCAB_NBR_KNN = 0,        /* true 16 nearest by (distSq, obstacles first, id): what RVO2-3D computes */
    SCAB_NBR_REF_COMPAT = 1  /* bit-exact replay of the reference: kdTree.py:127-156 traversal order +
                                agent.py:79-124 insertion with the by-value rangeSq (never shrinks) */
} scab_nbr_mode;

/* Status bits (mamp/agents/agent.py:70-73 flags). */
#define SCAB_ST_AT_GOAL 1
#define SCAB_ST_COLLISION 2
#define SCAB_ST_TIMEOUT 4
#define SCAB_ST_RUN_DONE 8
#define SCAB_ST_ZAXIS 16      /* static: start->goal is vertical (scaPolicy.py:188-190 -> 2x128 candidates) */
#define SCAB_ST_GHOST 32      /* read-only copy of an agent owned by another GPU (halo): visible as a neighbour /
                                 collision partner, never stepped, never modified (no reference counterpart) */

/* Branch tags written to state.tag (which path produced the velocity). */
#define SCAB_TAG_FIRST 1      /* zero-velocity shortcut vA_post = 0.3 v_pref (scaPolicy.py:34-40) */
#define SCAB_TAG_SUITABLE 2   /* sampled policies: a suitable candidate existed (scaPolicy.py:218) */
#define SCAB_TAG_FALLBACK 4   /* sampled policies: compute_without_suitV path (scaPolicy.py:224-238) */
#define SCAB_TAG_LP3_OK 8     /* ORCA-LP: linearProgram3 satisfied every plane (orca3dPolicyOfficial.py:108) */
#define SCAB_TAG_LP4 16       /* ORCA-LP: linearProgram4 fallback ran (orca3dPolicyOfficial.py:110-111) */

/* Constants of mamp/agents/agent.py:23-42 and mamp/configs/config.py:1-4 plus the cell-list geometry. */
typedef struct {
    double neighbor_dist;  /* 10.0  agent.py:33 */
    double time_step;      /* 0.1   config.py:2 */
    double time_horizon;   /* 10.0  agent.py:35 */
    double max_speed;      /* 1.0   agent.py:36 */
    double near_goal;      /* 0.5   config.py:3 */
    double grid_origin[3]; /* lower corner of the uniform grid (agents outside are clamped into the border cells) */
    double cell_size;      /* >= neighbor_dist */
    int32_t grid_dim[3];   /* cells per axis */
    int32_t reserved;
} scab_params;

/* Device pointers to the caller-owned structure-of-arrays state (replaces the per-object attributes
 * of mamp/agents/agent.py:13-36,56-74).  n rows each.  Optional pointers may be NULL. */
typedef struct {
    double *pos;           /* [n,3] pos_global_frame (f64)                          in/out */
    float *vel;            /* [n,3] vel_global_frame (f32, as after mampenv.py:104) in/out */
    double *heading;       /* [n,3] heading_global_frame (yaw, pitch, roll)         in/out */
    const double *goal;    /* [n,3] goal_global_frame                               in */
    const double *radius;  /* [n]                                                   in */
    const double *pref_speed;   /* [n]                                              in */
    int32_t *status;       /* [n] SCAB_ST_* bits                                    in/out */
    double *total_dist;    /* [n]                                                   in/out */
    int32_t *step_num;     /* [n]                                                   in/out */
    const double *max_run_dist; /* [n] 3*|start-goal| (agent.py:74)                 in */
    const double *vpref_ext;    /* [n,3] optional: externally supplied V_des (e.g. the Dubins provider of
                                   scaPolicy.py:264-338); NULL -> straight-line v_pref (rvo3dPolicy.py:182-196) */
    double *vpref_out;     /* [n,3] optional: agent.v_pref as stored by compute_v_pref  out */
    float *action;         /* [n,7] vx vy vz speed d_yaw d_pitch d_roll (mampenv.py:31) out */
    /* observability (optional) */
    int32_t *nbr_cnt;      /* [n]    len(agent.neighbors), -1 if computeNeighbors did not run  out */
    int32_t *nbr_ids;      /* [n,16] agent id, or n + obstacle id; -1 padding                  out */
    double *nbr_dsq;       /* [n,16] the sort key (agent.py:81,106)                            out */
    int32_t *tag;          /* [n] SCAB_TAG_* bits                                              out */
    int32_t *n_suit;       /* [n] sampled policies: number of suitable candidates              out */
    int32_t *plane_fail;   /* [n] ORCA-LP: linearProgram3's return value                       out */
    int32_t *n_planes;     /* [n] ORCA-LP: len(orcaPlanes)                                     out */
    int32_t *lp4_proj;     /* [n] ORCA-LP: number of inner dir-opt LP3 solves in LP4           out */
    /* multi-GPU sharding (optional) */
    const int32_t *gid;    /* [n] global agent id of each row (NULL: row index). Ties and the sequential
                              collision-check order (mampenv.py:42-46) follow the GLOBAL id              in */
    int64_t n_global;      /* total agents across all shards (0: n); obstacle ids are reported as n_global + k */
    /* Dubins v_pref provider (optional; see scab_dubins_enable) */
    const double *goal_heading; /* [n,3] goal_heading_frame = goal_pos[3:6] (agent.py:21); NULL: zeros           in */
    /* Waypoint following (optional): agent.path of an external global planner + policy.now_goal, i.e. get_trajectory
     * of rvo3dPolicy.py:71-86, srvo3dPolicy.py:71-86, orca3dPolicy.py:298-312, orca3dPolicyOfficial.py:302-316
     * (SCA and RVO3D+Dubins define it but never call it).  NULL waypoints: v_pref aims at the goal. */
    const double *waypoints;    /* [n, waypoint_stride, 3] in the reference's list order: pop() takes the LAST live entry in */
    int32_t *waypoint_left;     /* [n] len(agent.path)                                                      in/out */
    double *now_goal;           /* [n,3] policy.now_goal; NaN in [i,0] = None (not started)                 in/out */
    int64_t waypoint_stride;    /* waypoint slots per agent */
} scab_state_ptrs;

/* Step counters, read back with scab_read_counters (device->host copy of 8 x int64). */
typedef struct {
    int64_t n_active;      /* agents that ran the policy this step (not at_goal/collision/timeout) */
    int64_t n_not_done;    /* agents without RUN_DONE after the step; 0 <=> MACAEnv.is_done() is True */
    int64_t n_collision;   /* agents with the collision flag after the step */
    int64_t n_recheck;     /* fp32 guard-band hits re-evaluated in fp64 (diagnostic) */
    int64_t n_active_total;/* cumulative n_active since scab_create / scab_reset_counters (agent-steps) */
    int64_t n_tier2;       /* diagnostics: agents whose lazy tier 1 did not pin the result (tier 2 examined) */
    int64_t n_slow;        /* diagnostics: agents handed to the general selection kernel */
    int64_t overflow;      /* STICKY bit mask (cleared only by scab_reset_counters); non-zero = results may be wrong:
                              1 a halo face buffer was too small (records dropped), 2 a migration buffer was full (an agent
                              stayed outside its slab), 4 an arriving agent found no free row (lost), 8 an agent moved farther
                              in one step than scab_set_motion_bound allows (the collision stage's reach assumes the bound),
                              16 a peer-to-peer halo wait timed out, 32 an agent's radius exceeds scab_set_max_radius */
} scab_counters;
#define SCAB_OVF_HALO 1
#define SCAB_OVF_MIGRATION_DEFERRED 2
#define SCAB_OVF_MIGRATION_LOST 4
#define SCAB_OVF_MOTION 8
#define SCAB_OVF_RADIUS 32

const char *scab_version(void);
void scab_default_params(scab_params *p);

/* Handle life cycle.  `capacity` bounds n (library-owned scratch is sized from it). */
int scab_create(scab_handle **out, const scab_params *p, int device, int64_t capacity);
int scab_destroy(scab_handle *h);
const char *scab_last_error(const scab_handle *h);

/* Options.  SCAB_OPT_FULL_SEARCH (default 0): the sampled policies normally examine the candidates in two
 * tiers of increasing distance to v_pref and stop as soon as the reference's result is pinned (the selected
 * velocity and branch tag are identical; state.n_suit is then reported as -1).  1 = test all 513 candidates
 * like the reference does, which also yields its exact count of suitable candidates in state.n_suit. */
#define SCAB_OPT_FULL_SEARCH 1
int scab_set_option(scab_handle *h, int option, int value);

/* Largest agent radius in the swarm (0 = unknown).  When known, the status stage reuses the step's neighbour
 * lists instead of re-scanning the cell list wherever those lists provably hold every possible collision partner,
 * and the 16-NN query may look inside a ball smaller than neighbor_dist (never smaller than r_i + r_max: a collider
 * displaces every other neighbour from the reference's list, agent.py:82-92).  Unknown: full-radius queries only.
 * A radius above a declared bound raises SCAB_OVF_RADIUS in scab_counters.overflow at the next cell-list build. */
int scab_set_max_radius(scab_handle *h, double r_max);

/* Largest speed any agent can be commanded (max of max_speed, every pref_speed, 0.5 = the inner candidate shell, and the
 * norm of any externally supplied V_des).  The collision stage (mampenv.py:61-80 is O(N^2) and has no such limit) only
 * looks for partners within r_i + r_j + 2 v_max dt of an agent; the kinematic update checks every move against the bound
 * and raises SCAB_OVF_MOTION in scab_counters.overflow if it is ever exceeded.  Default: 1.25 m per step. */
int scab_set_motion_bound(scab_handle *h, double v_max);

/* Replace the grid geometry (e.g. after the swarm's bounding box changed). */
int scab_set_grid(scab_handle *h, const double origin[3], double cell_size, const int32_t dim[3]);

/* Bind the SoA state (replaces Agent attribute storage, agent.py:13-74). */
int scab_bind_state(scab_handle *h, const scab_state_ptrs *s, int64_t n);

/* Static obstacles: HOST array [m,4] = x y z r (mamp/agents/obstacle.py:10-33).  Builds the static
 * obstacle cell list and the reference's obstacle kd-tree (kdTree.py:158-227; mampenv.py:20). */
int scab_set_obstacles(scab_handle *h, const double *xyzr_host, int32_t m);

/* REF_COMPAT only: rebuild the reference's agent kd-tree from HOST positions [n,3] (kdTree.py:56-122,
 * including its persistent agentIDs permutation) and upload it.  Replaces KDTree.buildAgentTree(). */
int scab_kdtree_build(scab_handle *h, const double *pos_host, int64_t n);
/* Copy the current agentIDs permutation to `out_host` [n] (test hook). */
int scab_kdtree_get_perm(scab_handle *h, int32_t *out_host, int64_t n);

/* Stage 1: uniform-grid cell list over all agents (replaces the per-step kd-tree rebuild, mampenv.py:28). */
int scab_build_grid(scab_handle *h, void *stream);
/* Stage 2: neighbour query for every active agent (replaces computeNeighbors, scaPolicy.py:107-116 ->
 * kdTree.py:124-156,229-262 -> agent.py:79-124). */
int scab_neighbors(scab_handle *h, int mode, void *stream);
/* Stage 3: v_pref, constraint construction, velocity selection, action, kinematic update
 * (replaces Policy.find_next_action + update_velocitie, mampenv.py:83-105). */
int scab_policy_step(scab_handle *h, int policy, void *stream);
/* Stage 3 for the state rows [row_begin, row_end) only.  The stage may be issued as several row ranges that together cover
 * every row exactly once (any order): a caller whose state lives in host memory starts downloading the results of a range
 * while the next one computes (sca_b200.Swarm.step_from_host).  The union of the ranges gives exactly what scab_policy_step
 * gives.  Not available while the device Dubins tracker is enabled (its kernels run for every agent): SCAB_ERR_INVALID. */
int scab_policy_step_rows(scab_handle *h, int policy, int64_t row_begin, int64_t row_end, void *stream);
/* Stage 4: collision / timeout / goal flags and the done reduction (mampenv.py:51-80). */
int scab_status_step(scab_handle *h, void *stream);
/* Stages 1-4 back to back: one MACAEnv.step (mampenv.py:22-25). */
int scab_step(scab_handle *h, int policy, int nbr_mode, void *stream);

/* SCA's Dubins v_pref provider on the device (replaces compute_v_pref / compute_dubins / update_dubins,
 * mamp/policies/sca/scaPolicy.py:92-104,243-338, with dubinsmaneuver3d.py:34-162 and dubinsmaneuver2d.py:33-297).
 * on = 1 allocates (once) and resets one tracker state per agent row (agent.is_use_dubins = False); from then on
 * SCAB_POLICY_SCA and SCAB_POLICY_RVO3D_DUBINS take V_des from the tracker whenever state.vpref_ext is NULL.
 * on = 0 returns to the straight-line v_pref.  Single-GPU only (tracker states do not migrate between slabs). */
int scab_dubins_enable(scab_handle *h, int on, void *stream);
/* Test hook: bytes per tracker state, and a blocking copy of the first n states to host memory. */
int64_t scab_dubins_state_size(void);
int scab_dubins_read(scab_handle *h, void *out_host, int64_t n, void *stream);

/* Blocking read of the counters written by the last scab_status_step on `stream`. */
int scab_read_counters(scab_handle *h, scab_counters *out_host, void *stream);
/* Zero the cumulative counters (asynchronously on `stream`). */
int scab_reset_counters(scab_handle *h, void *stream);

/* ---- Multi-GPU spatial sharding (no reference counterpart: the reference is single-process). ----
 * Rows [0, n_own) of the bound state are owned; rows [n_own, n) are halo ghosts (SCAB_ST_GHOST).
 * scab_halo_pack: write the 64-byte record [x y z r | vx vy vz status | gid pad] of every owned agent
 * whose coordinate along `axis` is within `width` of the slab faces lo_face / hi_face into the device
 * buffers out_lo / out_hi (`cap` records each; unused slots hold a parked ghost at `park`, a position
 * inside a reserved empty grid layer).  Fixed-size and asynchronous: nothing is read back. */
int scab_halo_pack(scab_handle *h, int64_t n_own, int axis, double lo_face, double hi_face, double width,
                   const double park[3], void *out_lo, void *out_hi, int32_t cap, void *stream);
/* Scatter `count` received records into state rows [row0, row0 + count). */
int scab_halo_unpack(scab_handle *h, const void *recs, int32_t count, int64_t row0, void *stream);
/* After the policy stage: the new positions ([cap,3] f64 per face) of the agents packed by the last scab_halo_pack. */
int scab_halo_pack_pos(scab_handle *h, double *out_lo, double *out_hi, int32_t cap, const double park[3], void *stream);
int scab_halo_unpack_pos(scab_handle *h, const double *pos, int32_t count, int64_t row0, void *stream);
/* Device-side migration between slabs (asynchronous, fixed size, no host round trip).
 * scab_migrate_pack: every owned agent (rows [0, own_cap), ghosts skipped) whose coordinate left
 * [lo_face, hi_face) is written as a 128-byte full-state row into out_lo / out_hi (1 header row carrying the
 * count + `cap` rows each), its state row becomes a parked ghost and is pushed on the caller-owned device free
 * list (`free_list` int32[own_cap], `free_count` int32).  park_geom = {x0, y0, cell, z} and nx, ny spread the
 * parked rows over the reserved grid layer.  scab_migrate_unpack pops one free row per received agent.
 * `cnt` (device int32[4]): [0],[1] packed per face, [2] agents that did not fit (stay put), [3] arrivals lost. */
int scab_migrate_pack(scab_handle *h, int64_t own_cap, int axis, double lo_face, double hi_face, const double park_geom[4],
                      int32_t nx, int32_t ny, void *out_lo, void *out_hi, int32_t cap, int32_t *free_list, int32_t *free_count,
                      int32_t *cnt, void *stream);
int scab_migrate_unpack(scab_handle *h, const void *in_lo, const void *in_hi, int32_t cap, int use_lo, int use_hi,
                        int32_t *free_list, int32_t *free_count, int32_t *cnt, void *stream);
/* Per-agent state that lives outside the 128-byte row travels with a migrating agent in one extra row of
 * scab_migrate_extra_row_doubles(h) doubles, same slot as its state row: the Dubins tracker (the library's tracker state,
 * scab_state_ptrs.goal_heading; scaPolicy.py:264-338 keeps it on the Agent object) and the waypoint follower
 * (waypoints, waypoint_left, now_goal; agent.path / policy.now_goal in the reference).  Call order per migration:
 * scab_migrate_pack, scab_migrate_extra_pack (same cap, the same `cnt`), exchange both buffer pairs ([cap, row] f64,
 * no header), scab_migrate_unpack, scab_migrate_extra_unpack.  Needs 2 * cap <= capacity. */
int64_t scab_migrate_extra_row_doubles(scab_handle *h);
int scab_migrate_extra_pack(scab_handle *h, double *out_lo, double *out_hi, int32_t cap, const int32_t *cnt, void *stream);
int scab_migrate_extra_unpack(scab_handle *h, const double *in_lo, const double *in_hi, int32_t cap, void *stream);
/* ---- The same exchanges WITHOUT a communication library: peer-to-peer over NVLink.  The out_lo / out_hi arguments of
 * scab_halo_pack, scab_halo_pack_pos, scab_migrate_pack and scab_migrate_extra_pack may be PEER memory: the slab neighbour's
 * receive buffers.  scab_p2p_alloc allocates a zeroed device buffer that another process of the node can map
 * (handle_out: SCAB_IPC_HANDLE_BYTES bytes to hand to the peer; NULL = no handle wanted), scab_p2p_open maps a peer's
 * buffer from its handle; both are released by scab_destroy.  After its pack kernels a rank calls scab_p2p_signal:
 * seq = ++*local_seq, then (system-scope fence) *peer_flag_lo = *peer_flag_hi = seq (NULL = no neighbour on that side);
 * before its unpack kernels it calls scab_p2p_wait: e = ++*local_expect, spin until both flags >= e.  All four are
 * uint64 words in device memory (the flags inside a scab_p2p_alloc'ed buffer, written by the peers).  Asynchronous on `stream`,
 * capturable in a CUDA graph.  A wait that is not satisfied within ~2 s gives up and sets SCAB_OVF_P2P_TIMEOUT. */
#define SCAB_IPC_HANDLE_BYTES 64
#define SCAB_OVF_P2P_TIMEOUT 16
int scab_p2p_alloc(scab_handle *h, int64_t bytes, void **dev_ptr, unsigned char *handle_out);
int scab_p2p_open(scab_handle *h, const unsigned char *handle, void **dev_ptr);
int scab_p2p_signal(scab_handle *h, void *local_seq, void *peer_flag_lo, void *peer_flag_hi, void *stream);
int scab_p2p_wait(scab_handle *h, void *local_expect, const void *flag_from_lo, const void *flag_from_hi, void *stream);
/* The same wait WITHOUT a kernel of its own: folded into the next scab_halo_unpack / scab_halo_unpack_pos / scab_migrate_unpack
 * launch of this handle (every block's first thread spins, then the block proceeds).  The expected value is read from
 * *local_seq at run time: this rank's own sequence number of that exchange -- both sides signal once per exchange, so the
 * peers' numbers match it.  Call scab_p2p_signal for the exchange first. */
int scab_p2p_wait_next(scab_handle *h, const void *local_seq, const void *flag_from_lo, const void *flag_from_hi);

/* Blocking: how many agents the last scab_halo_pack selected per face (> cap means the buffers overflowed). */
int scab_halo_counts(scab_handle *h, int32_t out_host[2], void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCAB200_H */
