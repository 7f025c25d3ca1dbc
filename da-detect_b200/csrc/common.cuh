// Shared helpers for the dadetect_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dadetect_b200.h"

namespace dd {

extern thread_local char g_err[512];
extern long long g_launches;
extern int g_sm_budget;   // SMs the persistent dense kernels may occupy (dd_set_sm_budget)

inline int fail(int code, const char* what, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "%s (%s:%d): %s", what, file, line,
           code > 0 ? cudaGetErrorString((cudaError_t)code) : "invalid argument");
  return code;
}

#define DD_CHECK_ARG(cond)                                               \
  do {                                                                   \
    if (!(cond)) return dd::fail(-1, "argument check failed: " #cond, __FILE__, __LINE__); \
  } while (0)

#define DD_CUDA(expr)                                                    \
  do {                                                                   \
    cudaError_t e__ = (expr);                                            \
    if (e__ != cudaSuccess) return dd::fail((int)e__, #expr, __FILE__, __LINE__); \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define DD_LAUNCHED()                                                    \
  do {                                                                   \
    ++dd::g_launches;                                                    \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) return dd::fail((int)e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200
// CTAs a persistent dense kernel launches at most: all SMs, or fewer while latency-bound few-CTA kernels (top-k, NMS
// scan, samplers) run beside it on another stream and must keep their SMs (a persistent CTA that cannot become
// resident runs as a second wave and doubles the kernel's duration).
inline int sm_budget() { return g_sm_budget > 0 && g_sm_budget < kNumSMs ? g_sm_budget : kNumSMs; }

inline int grid_for(long long work_items, int block, int max_waves = 8) {
  long long b = (work_items + block - 1) / block;
  long long cap = (long long)kNumSMs * max_waves * (2048 / block);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// Exclusive scan over the CTA (<= 1024 threads, all participating) of one int per thread: returns this thread's
// offset, the CTA total in `total`.  `sm32` must hold >= 32 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* sm32, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();
  if (lane == 31) sm32[wid] = incl;
  __syncthreads();
  int w = lane < (int)(blockDim.x >> 5) ? sm32[lane] : 0;
  int wincl = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, wincl, o);
    if (lane >= o) wincl += t;
  }
  total = __shfl_sync(0xffffffffu, wincl, 31);
  const int base = __shfl_sync(0xffffffffu, wincl - w, wid);
  return base + incl - v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace dd
