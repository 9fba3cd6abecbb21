// tests/hostsim/lp_host.cpp -- TEST INFRASTRUCTURE: the device code of sca_b200/csrc/scab_policy_lp.cu (ORCA planes, fp32
// plane pre-filter, linearProgram1-4, action, kinematic update: every __device__ function and the kernel body itself),
// compiled for the host and run one thread at a time.  tests/test_hostsim_lp.py drives it against the oracle.
#define HOSTSIM_SINGLE_THREAD 1
#include "cuda_host_shim.h"
#define SCAB_HOSTSIM 1
#include "../../sca_b200/csrc/scab_policy_lp.cu"

double lp_smem[sizeof(Planes) / sizeof(double)];

// Snapshot in identity order (sorted index == row); obstacle k is neighbour id n + k.
extern "C" int hostsim_policy_lp(int mode, const scab_state_ptrs *st, int n, int m, const double *posr, const float *vel4,
                                 const int32_t *nbr, const int32_t *nbr_cnt, const int32_t *qcoll, const double *obs_posr,
                                 double time_step, double time_horizon, double max_speed, int32_t *slow_list, int32_t *slow_count,
                                 unsigned long long *counters)
{
    static int32_t *ids = nullptr;
    static int ids_n = 0;
    if (ids_n < n) { delete[] ids; ids = new int32_t[n]; ids_n = n; for (int i = 0; i < n; i++) ids[i] = i; }
    StepArgs A;
    std::memset(&A, 0, sizeof(A));
    A.st = *st;
    A.n = n; A.m = m;
    A.move_bound = 1e300;
    A.row_count = -1;
    A.time_step = time_step; A.inv_time_step = 1.0 / time_step; A.inv_time_horizon = 1.0 / time_horizon; A.max_speed = max_speed;
    A.s_posr = (double4 *)posr; A.s_vel = (float4 *)vel4; A.s_id = ids; A.s_gid = ids;
    A.s_nbr = (int32_t *)nbr; A.s_nbr_cnt = (int32_t *)nbr_cnt; A.s_qcoll = (int32_t *)qcoll;
    A.o_posr = (const double4 *)obs_posr; A.o_posr_by_id = (const double4 *)obs_posr;
    A.slow_list = slow_list; A.slow_count = slow_count; A.counters = counters;
    threadIdx = HostDim(); blockIdx = HostDim();
    switch (mode) {
    case 0: k_policy_lp<0>(A); break;
    case 1: *slow_count = 0; k_policy_lp<1>(A); k_policy_lp<2>(A); break;     // what launch_policy_lp does by default
    case 3: *slow_count = 0; k_policy_lp<3>(A); k_policy_lp<2>(A); break;     // SCAB_LP_F64=1
    case 4: *slow_count = 0; k_policy_lp<1>(A); k_policy_lp<4>(A); break;     // SCAB_LP_COMPACT=1
    case 5: *slow_count = 0; k_policy_lp<5>(A); k_policy_lp<2>(A); break;     // SCAB_LP_PASS1_CTAS8=1
    default: return -1;
    }
    return 0;
}
