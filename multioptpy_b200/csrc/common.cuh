// Shared device helpers for the mop_b200 kernels (sm_100a, FP64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/mop_b200.h"

#define MOP_FULL_MASK 0xffffffffu

namespace mop {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MOP_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(MOP_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(MOP_FULL_MASK, v, o));
  return v;
}

// Block-wide sum, result broadcast to every thread.  scratch: >= 33 doubles of
// shared memory.  Deterministic (fixed order).  Contains two __syncthreads.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch from a previous use
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ double block_max(double v, double* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// Single-barrier block reductions.  `buf` is a [2][32*K] double scratch in shared
// memory and `parity` a per-thread counter toggled on every call: consecutive
// reductions alternate buffers, so one __syncthreads per reduction is enough (the
// barrier of call i+1 orders the reads of call i before the writes of call i+2).
// Every warp re-reduces the per-warp partials, so all threads get the result.
template <int K>
__device__ __forceinline__ void block_sum_k(double (&v)[K], double* buf, int& parity) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  double* b = buf + (parity & 1) * (32 * K);
  parity ^= 1;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) b[k * 32 + w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(lane < nw ? b[k * 32 + lane] : 0.0);
}

// ~1 ulp reciprocal without the IEEE slow path: MUFU seed + two Newton steps.
// Valid for normal, non-tiny |x| (callers clamp pivots away from zero).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ double sgn(double x) { return (x > 0.0) - (x < 0.0); }

// y[i] = sum_j A[i][j] * v[j] for a row-major n x n matrix (lda) in global or
// shared memory; one warp per row, lanes stride the row (coalesced).
__device__ __forceinline__ void block_matvec(const double* __restrict__ A, int lda, int n,
                                             const double* __restrict__ v, double* __restrict__ y) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = w; i < n; i += nw) {
    const double* row = A + (size_t)i * lda;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(row[j], v[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[i] = acc;
  }
}

}  // namespace mop

// host-side error plumbing (capi.cu)
void mop_set_error(const char* fmt, ...);
#define MOP_CHECK_CUDA(expr)                                                        \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mop_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MOP_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)
#define MOP_REQUIRE(cond, msg)       \
  do {                               \
    if (!(cond)) {                   \
      mop_set_error("%s", msg);      \
      return MOP_ERR_INVALID;        \
    }                                \
  } while (0)
