// Measurement helpers exported through the C ABI (used by bench.py only):
// an FP64 FMA peak probe (the denominator of the FP64 roofline; MEASURED_PEAKS.json
// carries no FP64 figure) and an L2 flush.
#include "common.cuh"

namespace mop {

__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* out) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keep the chain alive
}

__global__ void k_fill(double* p, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void k_fast_rcp(const double* x, double* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = fast_rcp(x[i]);
}

}  // namespace mop

// Launches blocks x 256 threads, each doing iters * 64 dependent-chain-interleaved DFMAs.
// FLOPs of one launch = blocks * 256 * iters * 64 * 2.
extern "C" int mop_bench_dfma(int blocks, int iters, double* out, void* stream) {
  MOP_REQUIRE(blocks > 0 && iters > 0 && out, "mop_bench_dfma: bad arguments");
  mop::k_dfma_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// Writes `count` doubles (use a buffer larger than the 126 MB L2 to flush it).
extern "C" int mop_bench_fill(double* buf, size_t count, double value, void* stream) {
  MOP_REQUIRE(buf && count > 0, "mop_bench_fill: bad arguments");
  mop::k_fill<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(buf, count, value);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// out[i] = fast_rcp(x[i]) (the division-free reciprocal used by the twisted factorisation)
extern "C" int mop_debug_fast_rcp(const double* x, double* out, size_t count, void* stream) {
  MOP_REQUIRE(x && out && count > 0, "mop_debug_fast_rcp: bad arguments");
  mop::k_fast_rcp<<<148, 256, 0, (cudaStream_t)stream>>>(x, out, count);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
