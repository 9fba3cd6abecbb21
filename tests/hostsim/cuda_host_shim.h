// tests/hostsim/cuda_host_shim.h -- TEST INFRASTRUCTURE, never part of the product.
//
// Just enough of the CUDA device environment to compile sca_b200/csrc/*.cu for the HOST and execute the kernels
// functionally: every CUDA thread of a block is a real host thread (hostsim_runtime.cpp keeps a pool), blocks run one after
// the other, `__shared__` is plain static storage (one block at a time), `__syncthreads` and the `*_sync` warp primitives
// are rendezvous points between the participating host threads.  The arithmetic intrinsics map to the IEEE operations
// they stand for; compile with -ffp-contract=off (the device build uses -fmad=false: the only fused multiply-adds are
// the explicit ones, std::fma here).  Two build flavours:
//   * HOSTSIM_SINGLE_THREAD (tests/hostsim/lp_host.cpp): one thread per block, no runtime -- thread-per-agent kernels only;
//   * default: the full emulator behind the same C ABI as libscab200.so (tests/hostsim/build.py).
// What this checks is kernel LOGIC (indexing, branches, warp-cooperative protocols, the C ABI's host code); it says nothing
// about performance and is not a substitute for the -m gpu tests, which run the compiled sm_100a code.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>   // vector types, cudaError_t, cudaStream_t and the API prototypes (defined in hostsim_runtime.cpp)

#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__

struct HostDim { unsigned x = 0, y = 0, z = 0; };

#ifdef HOSTSIM_SINGLE_THREAD
static HostDim threadIdx, blockIdx;
static HostDim blockDim{1, 1, 1}, gridDim{1, 1, 1};
#undef __shared__
#define __shared__
#else
extern thread_local HostDim threadIdx, blockIdx, blockDim, gridDim;
// one block at a time => block-shared memory is ordinary static storage
#undef __shared__
#define __shared__ static
#endif

using std::isnan;
using std::max;
using std::min;

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
static inline void __threadfence_system() { __sync_synchronize(); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

#ifdef HOSTSIM_SINGLE_THREAD
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { return T(0); }   // the other lanes add nothing
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
template <typename T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
static inline int atomicAdd(int32_t *p, int v) { int old = *p; *p = old + v; return old; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { unsigned long long old = *p; *p = old | v; return old; }
#else
namespace hostsim {
// all lanes named in `mask` deposit `v` and get everybody's deposits back (lanes that already left the kernel: 0)
void warp_exchange(unsigned mask, unsigned long long v, unsigned long long out[32], unsigned *live_mask);
void block_barrier();
int block_or(int pred);
struct Cfg { unsigned grid, block; size_t smem; };
static inline Cfg cfg(unsigned g, unsigned b, size_t sm = 0, cudaStream_t = 0) { return Cfg{g, b, sm}; }
void launch_impl(Cfg c, void (*thunk)(void *), void *ctx);
template <class F> static inline void launch(Cfg c, F f) { launch_impl(c, [](void *p) { (*static_cast<F *>(p))(); }, &f); }
template <typename T> static inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
static inline unsigned lane_id() { return threadIdx.x & 31u; }
}  // namespace hostsim

static inline void __syncthreads() { hostsim::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned long long o[32]; unsigned live; hostsim::warp_exchange(mask, 0, o, &live); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    unsigned long long o[32]; unsigned live, r = 0;
    hostsim::warp_exchange(mask, pred ? 1 : 0, o, &live);
    for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r |= (unsigned)(o[l] & 1u) << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    unsigned long long o[32]; unsigned live;
    hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    const int l = hostsim::lane_id(), base = l & ~(width - 1);
    return hostsim::from_bits<T>(o[base + (src & (width - 1))]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32)
{
    unsigned long long o[32]; unsigned live;
    hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    const int l = hostsim::lane_id(), src = l ^ lane_mask;
    return ((src & ~(width - 1)) == (l & ~(width - 1))) ? hostsim::from_bits<T>(o[src]) : v;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    unsigned long long o[32]; unsigned live;
    hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    const int l = hostsim::lane_id(), src = l - (int)delta;
    return (src >= (l & ~(width - 1))) ? hostsim::from_bits<T>(o[src]) : v;
}
#define HOSTSIM_REDUCE(NAME, T, INIT, OP)                                                        \
    static inline T NAME(unsigned mask, T v)                                                     \
    {                                                                                            \
        unsigned long long o[32]; unsigned live;                                                 \
        hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);                             \
        T r = INIT;                                                                              \
        for (int l = 0; l < 32; l++) if ((live >> l) & 1u) { T x = hostsim::from_bits<T>(o[l]); r = OP; } \
        return r;                                                                                \
    }
HOSTSIM_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (x < r ? x : r))
HOSTSIM_REDUCE(__reduce_max_sync, unsigned, 0u, (x > r ? x : r))
HOSTSIM_REDUCE(__reduce_add_sync, unsigned, 0u, (r + x))
static inline int __reduce_min_sync(unsigned mask, int v)
{
    unsigned long long o[32]; unsigned live; hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    int r = 0x7fffffff; for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::min(r, hostsim::from_bits<int>(o[l])); return r;
}
static inline int __reduce_max_sync(unsigned mask, int v)
{
    unsigned long long o[32]; unsigned live; hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    int r = -0x7fffffff - 1; for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r = std::max(r, hostsim::from_bits<int>(o[l])); return r;
}
static inline int __syncthreads_or(int pred) { return hostsim::block_or(pred); }
static inline int __reduce_add_sync(unsigned mask, int v)
{
    unsigned long long o[32]; unsigned live; hostsim::warp_exchange(mask, hostsim::to_bits(v), o, &live);
    int r = 0; for (int l = 0; l < 32; l++) if ((live >> l) & 1u) r += hostsim::from_bits<int>(o[l]); return r;
}
// kernel-handle overloads of the two runtime queries (cuda_runtime.h only has them under nvcc)
template <class T> static inline cudaError_t cudaFuncSetAttribute(T *, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, T *, int, size_t) { *n = 2; return cudaSuccess; }
// real atomics: the lanes are real threads
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicSub(int *p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
#endif
