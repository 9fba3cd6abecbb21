// scab_p2p.cu -- the halo / migration exchange WITHOUT a communication library: the pack kernels of scab_halo.cu write
// straight into the slab neighbour's receive buffers over NVLink (peer memory opened through CUDA IPC), a one-thread
// kernel publishes "my data for exchange #seq is there", and the consumer's one-thread kernel waits for it before its unpack
// kernel runs.  Per exchange that is two ~3 us graph nodes instead of pack -> ncclSend/ncclRecv -> unpack (60-130 us at these
// message sizes).  The reference is single-process: no counterpart.
//
// Protocol (every rank runs the same step sequence, so sequence numbers need no negotiation):
//   producer   pack kernels write the peer's buffers;  k_p2p_signal: seq = ++local_seq[kind]; fence.sys; peer.flag[kind][side] = seq
//   consumer   k_p2p_wait: e = ++local_expect[kind]; spin until flag[kind][from_lo] >= e and flag[kind][from_hi] >= e
// Single buffers are enough: a rank sends exchange k+1 of a kind only after it received the peer's NEXT exchange of the
// other kind (the pre- and post-policy exchanges alternate), which the peer sends after it consumed exchange k.
// A wait that does not complete within ~4 s sets the sticky OVF_P2P_TIMEOUT bit and gives up (never hang the GPU).
#include "scab_common.cuh"

__global__ void k_p2p_signal(unsigned long long *local_seq, unsigned long long *peer_flag_lo, unsigned long long *peer_flag_hi)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long seq = *local_seq + 1ull;
    *local_seq = seq;
    __threadfence_system();                    // everything the pack kernels wrote to the peers is visible before the flag
    if (peer_flag_lo) { *(volatile unsigned long long *)peer_flag_lo = seq; }
    if (peer_flag_hi) { *(volatile unsigned long long *)peer_flag_hi = seq; }
    __threadfence_system();
}

__global__ void k_p2p_wait(unsigned long long *local_expect, const unsigned long long *flag_from_lo, const unsigned long long *flag_from_hi,
                           unsigned long long *counters)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long e = *local_expect + 1ull;
    *local_expect = e;
    const long long t0 = clock64();
    for (;;) {
        const bool lo_ok = !flag_from_lo || *(volatile const unsigned long long *)flag_from_lo >= e;
        const bool hi_ok = !flag_from_hi || *(volatile const unsigned long long *)flag_from_hi >= e;
        if (lo_ok && hi_ok) break;
        if (clock64() - t0 > 8000000000ll) { atomicOr(&counters[7], OVF_P2P_TIMEOUT); break; }
        __nanosleep(40);
    }
    __threadfence_system();                    // acquire: the unpack kernel that follows reads what the peer wrote
}

cudaError_t launch_p2p_signal(unsigned long long *local_seq, unsigned long long *peer_flag_lo, unsigned long long *peer_flag_hi, cudaStream_t s)
{
    k_p2p_signal<<<1, 32, 0, s>>>(local_seq, peer_flag_lo, peer_flag_hi);
    return cudaGetLastError();
}

cudaError_t launch_p2p_wait(const StepArgs &A, unsigned long long *local_expect, const unsigned long long *flag_from_lo,
                            const unsigned long long *flag_from_hi, cudaStream_t s)
{
    k_p2p_wait<<<1, 32, 0, s>>>(local_expect, flag_from_lo, flag_from_hi, A.counters);
    return cudaGetLastError();
}
