// tests/hostsim/cuda_host_shim.h -- TEST INFRASTRUCTURE: just enough of the CUDA device environment to compile the
// thread-per-agent device code of sca_b200/csrc for the HOST and run it one thread at a time (gridDim = blockDim = 1).
// The arithmetic intrinsics map to the IEEE operations they stand for (compile with -ffp-contract=off: the device build
// uses -fmad=false, so the only fused multiply-adds are the explicit ones, std::fma here).  Warp-cooperative kernels
// (k_select*, the scan kernels) cannot be run this way and are not included.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>   // vector types, cudaError_t, cudaStream_t (host API headers; nothing is called)

#undef __launch_bounds__
#define __launch_bounds__(...)

struct HostDim { unsigned x = 0, y = 0, z = 0; };
static HostDim threadIdx, blockIdx;
static HostDim blockDim{1, 1, 1}, gridDim{1, 1, 1};

using std::max;
using std::min;
using std::isnan;

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
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
// one thread per "warp": the other 31 lanes contribute nothing
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int, int = 32) { return T(0); }
static inline unsigned __ballot_sync(unsigned, int pred) { return pred ? 1u : 0u; }
template <typename T> static inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
static inline int atomicAdd(int32_t *p, int v) { int old = *p; *p = old + v; return old; }
