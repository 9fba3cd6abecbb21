"""ctypes binding of libscab200.so (include/scab200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at load time, and every
entry point raises ScabError on a non-zero status.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SCAB_LIB_PATH') or os.path.join(HERE, 'libscab200.so')   # override: A/B builds

POLICIES = {'sca': 0, 'srvo3d': 1, 'rvo3d': 2, 'rvo3d_dubins': 3, 'orca_sampled': 4, 'orca_lp': 5}
NBR_KNN, NBR_REF_COMPAT = 0, 1
K = 16

ST_AT_GOAL, ST_COLLISION, ST_TIMEOUT, ST_RUN_DONE, ST_ZAXIS, ST_GHOST = 1, 2, 4, 8, 16, 32
TAG_FIRST, TAG_SUITABLE, TAG_FALLBACK, TAG_LP3_OK, TAG_LP4 = 1, 2, 4, 8, 16

# every symbol include/scab200.h declares (tests/test_capi_symbols.py checks the header against this)
OPT_FULL_SEARCH = 1

OVF_HALO, OVF_MIGRATION_DEFERRED, OVF_MIGRATION_LOST, OVF_MOTION, OVF_P2P_TIMEOUT, OVF_RADIUS = 1, 2, 4, 8, 16, 32
IPC_HANDLE_BYTES = 64
OVF_TEXT = {1: 'a halo face buffer overflowed (ghost records dropped)', 2: 'a migration buffer was full (an agent stayed outside its slab)',
            4: 'an arriving agent found no free row (agent lost)',
            8: 'an agent moved farther in one step than the motion bound the collision stage relies on (scab_set_motion_bound)',
            16: 'a peer-to-peer halo wait timed out (the neighbour rank never published its data)',
            32: 'an agent has a larger radius than scab_set_max_radius declared (the 16-NN query ball and the status shortcut rely on it)'}

SYMBOLS = ['scab_version', 'scab_default_params', 'scab_set_option', 'scab_set_max_radius', 'scab_set_motion_bound', 'scab_create', 'scab_destroy', 'scab_last_error', 'scab_set_grid',
           'scab_bind_state', 'scab_set_obstacles', 'scab_kdtree_build', 'scab_kdtree_get_perm', 'scab_build_grid',
           'scab_neighbors', 'scab_policy_step', 'scab_policy_step_rows', 'scab_status_step', 'scab_step', 'scab_read_counters',
           'scab_reset_counters', 'scab_halo_pack', 'scab_halo_unpack', 'scab_halo_pack_pos', 'scab_halo_unpack_pos',
           'scab_halo_counts', 'scab_migrate_pack', 'scab_migrate_unpack', 'scab_migrate_extra_row_doubles', 'scab_migrate_extra_pack',
           'scab_migrate_extra_unpack', 'scab_p2p_alloc', 'scab_p2p_open', 'scab_p2p_signal', 'scab_p2p_wait', 'scab_p2p_wait_next', 'scab_dubins_enable', 'scab_dubins_state_size', 'scab_dubins_read']


class ScabError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [('neighbor_dist', C.c_double), ('time_step', C.c_double), ('time_horizon', C.c_double),
                ('max_speed', C.c_double), ('near_goal', C.c_double), ('grid_origin', C.c_double * 3),
                ('cell_size', C.c_double), ('grid_dim', C.c_int32 * 3), ('reserved', C.c_int32)]


STATE_FIELDS = ['pos', 'vel', 'heading', 'goal', 'radius', 'pref_speed', 'status', 'total_dist', 'step_num',
                'max_run_dist', 'vpref_ext', 'vpref_out', 'action', 'nbr_cnt', 'nbr_ids', 'nbr_dsq', 'tag', 'n_suit',
                'plane_fail', 'n_planes', 'lp4_proj', 'gid']


class StatePtrs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in STATE_FIELDS] + [('n_global', C.c_int64), ('goal_heading', C.c_void_p), ('waypoints', C.c_void_p),
                                                               ('waypoint_left', C.c_void_p), ('now_goal', C.c_void_p), ('waypoint_stride', C.c_int64)]


# mirror of DubState / Man2D in csrc/scab_dubins.cu (test hook only)
_MAN2D = [('qi', '<f8', 3), ('r', '<f8'), ('t', '<f8'), ('p', '<f8'), ('q', '<f8'), ('length', '<f8'), ('mode', '<i4', 3), ('ok', '<i4')]
DUBINS_STATE_DTYPE = [('in_use', '<i4'), ('n_points', '<i4'), ('cursor', '<i4'), ('has_nbr', '<i4'), ('ss', '<f8'), ('length', '<f8'),
                      ('qi0', '<f8', 3), ('h', _MAN2D), ('v', _MAN2D), ('now_goal', '<f8', 3), ('v_pref', '<f8', 3), ('nbr0_dsq', '<f8'),
                      ('replans', '<i4'), ('pad', '<i4')]


class Counters(C.Structure):
    _fields_ = [('n_active', C.c_int64), ('n_not_done', C.c_int64), ('n_collision', C.c_int64),
                ('n_recheck', C.c_int64), ('n_active_total', C.c_int64), ('n_tier2', C.c_int64), ('n_slow', C.c_int64),
                ('overflow', C.c_int64)]


_lib = None


def lib():
    """Load libscab200.so (built in-tree by sca_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScabError('libscab200.so is missing (%s): run `python -m sca_b200.build`; there is no CPU fallback'
                        % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.scab_version.restype = C.c_char_p
    L.scab_last_error.restype = C.c_char_p
    L.scab_last_error.argtypes = [vp]
    L.scab_default_params.restype = None
    L.scab_default_params.argtypes = [C.POINTER(Params)]
    L.scab_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int, i64]
    L.scab_destroy.argtypes = [vp]
    L.scab_set_grid.argtypes = [vp, C.POINTER(C.c_double), C.c_double, C.POINTER(i32)]
    L.scab_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.scab_set_max_radius.argtypes = [vp, C.c_double]
    L.scab_set_motion_bound.argtypes = [vp, C.c_double]
    L.scab_bind_state.argtypes = [vp, C.POINTER(StatePtrs), i64]
    L.scab_set_obstacles.argtypes = [vp, vp, i32]
    L.scab_kdtree_build.argtypes = [vp, vp, i64]
    L.scab_kdtree_get_perm.argtypes = [vp, vp, i64]
    L.scab_build_grid.argtypes = [vp, vp]
    L.scab_neighbors.argtypes = [vp, C.c_int, vp]
    L.scab_policy_step.argtypes = [vp, C.c_int, vp]
    L.scab_status_step.argtypes = [vp, vp]
    L.scab_policy_step_rows.argtypes = [vp, C.c_int, i64, i64, vp]
    L.scab_step.argtypes = [vp, C.c_int, C.c_int, vp]
    L.scab_read_counters.argtypes = [vp, C.POINTER(Counters), vp]
    L.scab_reset_counters.argtypes = [vp, vp]
    L.scab_dubins_enable.argtypes = [vp, C.c_int, vp]
    L.scab_dubins_state_size.restype = i64
    L.scab_dubins_state_size.argtypes = []
    L.scab_dubins_read.argtypes = [vp, vp, i64, vp]
    d3 = C.POINTER(C.c_double)
    L.scab_halo_pack.argtypes = [vp, i64, C.c_int, C.c_double, C.c_double, C.c_double, d3, vp, vp, i32, vp]
    L.scab_halo_unpack.argtypes = [vp, vp, i32, i64, vp]
    L.scab_halo_pack_pos.argtypes = [vp, vp, vp, i32, d3, vp]
    L.scab_halo_unpack_pos.argtypes = [vp, vp, i32, i64, vp]
    L.scab_halo_counts.argtypes = [vp, C.POINTER(i32), vp]
    L.scab_migrate_pack.argtypes = [vp, i64, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), i32, i32, vp, vp, i32, vp, vp, vp, vp]
    L.scab_migrate_unpack.argtypes = [vp, vp, vp, i32, C.c_int, C.c_int, vp, vp, vp, vp]
    L.scab_migrate_extra_row_doubles.restype = i64
    L.scab_migrate_extra_row_doubles.argtypes = [vp]
    L.scab_migrate_extra_pack.argtypes = [vp, vp, vp, i32, vp, vp]
    L.scab_migrate_extra_unpack.argtypes = [vp, vp, vp, i32, vp]
    L.scab_p2p_alloc.argtypes = [vp, i64, C.POINTER(vp), vp]
    L.scab_p2p_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.scab_p2p_signal.argtypes = [vp, vp, vp, vp, vp]
    L.scab_p2p_wait.argtypes = [vp, vp, vp, vp, vp]
    L.scab_p2p_wait_next.argtypes = [vp, vp, vp, vp]
    _lib = L
    return L


def default_params():
    p = Params()
    lib().scab_default_params(C.byref(p))
    return p


class Handle:
    """RAII wrapper of scab_handle."""

    def __init__(self, params, device, capacity):
        self._h = C.c_void_p()
        self._L = lib()
        rc = self._L.scab_create(C.byref(self._h), C.byref(params), int(device), int(capacity))
        if rc != 0:
            raise ScabError('scab_create failed with status %d' % rc)

    def _ck(self, rc):
        if rc != 0:
            msg = self._L.scab_last_error(self._h)
            raise ScabError('libscab200 status %d: %s' % (rc, msg.decode() if msg else '?'))

    def close(self):
        if self._h:
            self._L.scab_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_max_radius(self, r_max):
        self._ck(self._L.scab_set_max_radius(self._h, float(r_max)))

    def set_motion_bound(self, v_max):
        self._ck(self._L.scab_set_motion_bound(self._h, float(v_max)))

    def set_option(self, option, value):
        self._ck(self._L.scab_set_option(self._h, int(option), int(value)))

    def set_grid(self, origin, cell_size, dim):
        o = (C.c_double * 3)(*[float(x) for x in origin])
        d = (C.c_int32 * 3)(*[int(x) for x in dim])
        self._ck(self._L.scab_set_grid(self._h, o, float(cell_size), d))

    def bind_state(self, ptrs, n):
        self._ck(self._L.scab_bind_state(self._h, C.byref(ptrs), int(n)))

    def set_obstacles(self, xyzr_host):
        import numpy as np
        a = np.ascontiguousarray(xyzr_host, dtype='float64').reshape(-1, 4)
        self._ck(self._L.scab_set_obstacles(self._h, a.ctypes.data if len(a) else None, len(a)))

    def kdtree_build(self, pos_host):
        import numpy as np
        a = np.ascontiguousarray(pos_host, dtype='float64')
        self._ck(self._L.scab_kdtree_build(self._h, a.ctypes.data, a.shape[0]))

    def kdtree_perm(self, n):
        import numpy as np
        out = np.zeros(n, np.int32)
        self._ck(self._L.scab_kdtree_get_perm(self._h, out.ctypes.data, n))
        return out

    def build_grid(self, stream=0):
        self._ck(self._L.scab_build_grid(self._h, stream))

    def neighbors(self, mode, stream=0):
        self._ck(self._L.scab_neighbors(self._h, mode, stream))

    def policy_step(self, policy, stream=0):
        self._ck(self._L.scab_policy_step(self._h, policy, stream))

    def policy_step_rows(self, policy, row_begin, row_end, stream=0):
        self._ck(self._L.scab_policy_step_rows(self._h, policy, int(row_begin), int(row_end), stream))

    def status_step(self, stream=0):
        self._ck(self._L.scab_status_step(self._h, stream))

    def step(self, policy, nbr_mode, stream=0):
        self._ck(self._L.scab_step(self._h, policy, nbr_mode, stream))

    def read_counters(self, stream=0, check=True):
        """The step counters.  A non-zero sticky `overflow` mask means results may be wrong: raises unless check=False."""
        c = Counters()
        self._ck(self._L.scab_read_counters(self._h, C.byref(c), stream))
        if check and c.overflow:
            raise ScabError('libscab200 overflow flags %d: %s' % (c.overflow, '; '.join(t for b, t in OVF_TEXT.items() if c.overflow & b)))
        return {'overflow': c.overflow, 'n_active': c.n_active, 'n_not_done': c.n_not_done, 'n_collision': c.n_collision,
                'n_recheck': c.n_recheck, 'n_active_total': c.n_active_total, 'n_tier2': c.n_tier2, 'n_slow': c.n_slow}

    def reset_counters(self, stream=0):
        self._ck(self._L.scab_reset_counters(self._h, stream))

    # ---- Dubins v_pref provider (scaPolicy.py:264-338 on the device)
    def dubins_enable(self, on=True, stream=0):
        self._ck(self._L.scab_dubins_enable(self._h, 1 if on else 0, stream))

    def dubins_read(self, n, stream=0):
        """Tracker states of rows [0, n) as a structured numpy array (test hook)."""
        import numpy as np
        out = np.zeros(int(n), dtype=DUBINS_STATE_DTYPE)
        assert out.dtype.itemsize == self._L.scab_dubins_state_size(), (out.dtype.itemsize, self._L.scab_dubins_state_size())
        if n:
            self._ck(self._L.scab_dubins_read(self._h, out.ctypes.data, int(n), stream))
        return out

    # ---- multi-GPU halo (device pointers are plain ints: tensor.data_ptr())
    def halo_pack(self, n_own, axis, lo_face, hi_face, width, park, out_lo, out_hi, cap, stream=0):
        p = (C.c_double * 3)(*[float(x) for x in park])
        self._ck(self._L.scab_halo_pack(self._h, int(n_own), int(axis), float(lo_face), float(hi_face), float(width), p,
                                        out_lo, out_hi, int(cap), stream))

    def halo_unpack(self, recs, count, row0, stream=0):
        self._ck(self._L.scab_halo_unpack(self._h, recs, int(count), int(row0), stream))

    def halo_pack_pos(self, out_lo, out_hi, cap, park, stream=0):
        p = (C.c_double * 3)(*[float(x) for x in park])
        self._ck(self._L.scab_halo_pack_pos(self._h, out_lo, out_hi, int(cap), p, stream))

    def halo_unpack_pos(self, pos, count, row0, stream=0):
        self._ck(self._L.scab_halo_unpack_pos(self._h, pos, int(count), int(row0), stream))

    def migrate_pack(self, own_cap, axis, lo_face, hi_face, park_geom, nx, ny, out_lo, out_hi, cap, free_list, free_count, cnt, stream=0):
        g = (C.c_double * 4)(*[float(x) for x in park_geom])
        self._ck(self._L.scab_migrate_pack(self._h, int(own_cap), int(axis), float(lo_face), float(hi_face), g, int(nx), int(ny),
                                           out_lo, out_hi, int(cap), free_list, free_count, cnt, stream))

    def migrate_unpack(self, in_lo, in_hi, cap, use_lo, use_hi, free_list, free_count, cnt, stream=0):
        self._ck(self._L.scab_migrate_unpack(self._h, in_lo, in_hi, int(cap), int(use_lo), int(use_hi), free_list, free_count, cnt, stream))

    def migrate_extra_row_doubles(self):
        return int(self._L.scab_migrate_extra_row_doubles(self._h))

    def migrate_extra_pack(self, out_lo, out_hi, cap, cnt, stream=0):
        self._ck(self._L.scab_migrate_extra_pack(self._h, out_lo, out_hi, int(cap), cnt, stream))

    def migrate_extra_unpack(self, in_lo, in_hi, cap, stream=0):
        self._ck(self._L.scab_migrate_extra_unpack(self._h, in_lo, in_hi, int(cap), stream))

    # ---- peer-to-peer exchange (NVLink, CUDA IPC)
    def p2p_alloc(self, nbytes, want_handle=True):
        """-> (device pointer, IPC handle bytes or None).  Zeroed; released by scab_destroy."""
        ptr = C.c_void_p()
        buf = (C.c_ubyte * IPC_HANDLE_BYTES)() if want_handle else None
        self._ck(self._L.scab_p2p_alloc(self._h, int(nbytes), C.byref(ptr), C.cast(buf, C.c_void_p) if want_handle else None))
        return int(ptr.value), (bytes(buf) if want_handle else None)

    def p2p_open(self, handle):
        ptr = C.c_void_p()
        buf = (C.c_ubyte * IPC_HANDLE_BYTES).from_buffer_copy(handle)
        self._ck(self._L.scab_p2p_open(self._h, C.cast(buf, C.c_void_p), C.byref(ptr)))
        return int(ptr.value)

    def p2p_signal(self, local_seq, peer_flag_lo, peer_flag_hi, stream=0):
        self._ck(self._L.scab_p2p_signal(self._h, local_seq, peer_flag_lo or None, peer_flag_hi or None, stream))

    def p2p_wait(self, local_expect, flag_from_lo, flag_from_hi, stream=0):
        self._ck(self._L.scab_p2p_wait(self._h, local_expect, flag_from_lo or None, flag_from_hi or None, stream))

    def p2p_wait_next(self, local_seq, flag_from_lo, flag_from_hi):
        self._ck(self._L.scab_p2p_wait_next(self._h, local_seq, flag_from_lo or None, flag_from_hi or None))

    def halo_counts(self, stream=0):
        out = (C.c_int32 * 2)()
        self._ck(self._L.scab_halo_counts(self._h, out, stream))
        return int(out[0]), int(out[1])
