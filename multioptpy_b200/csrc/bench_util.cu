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

// dependent-chain latencies (cycles per op), one warp: out[0..7] =
// DFMA, DADD, DMUL, shared load (pointer chase), 64-bit shuffle, fast_rcp, sqrt, div
__global__ void k_latency(double* out) {
  __shared__ double sm[64];
  __shared__ int idx[64];
  const int lane = threadIdx.x;
  for (int i = lane; i < 64; i += 32) { sm[i] = 1.0 + 1e-9 * i; idx[i] = (i + 1) & 63; }
  __syncthreads();
  const int N = 2048;
  double x = 1.0 + lane * 1e-6, y = 1.0000001, z = 1e-9;
  long long t0, t1;
  double res[8];
#define T0() do { t0 = clock64(); x += (double)(t0 & 1) * 1e-300; p += (int)(t0 & 0); } while (0)
#define T1() do { t1 = clock64() + (long long)(x == 123.456) + (long long)(p == -7); } while (0)
  int p = lane;
  T0();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, z);
  T1(); res[0] = double(t1 - t0) / N;
  T0();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + z;
  T1(); res[1] = double(t1 - t0) / N;
  T0();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  T1(); res[2] = double(t1 - t0) / N;
  T0();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = idx[p];
  T1(); res[3] = double(t1 - t0) / N;
  x += p;
  T0();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  T1(); res[4] = double(t1 - t0) / N;
  x = 1.5 + lane * 1e-3;
  T0();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = fast_rcp(x);
  T1(); res[5] = double(t1 - t0) / N;
  x = 2.0 + lane;
  T0();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = sqrt(x) + 1.0;
  T1(); res[6] = double(t1 - t0) / N;
  T0();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 3.0 / x;
  T1(); res[7] = double(t1 - t0) / N;
  if (lane == 0) {
    for (int i = 0; i < 8; ++i) out[i] = res[i];
    out[8] = x;
  }
}

// barrier + reduce-publish-broadcast round-trip cost for a CTA of blockDim.x threads:
// out[0] = cycles per bare __syncthreads, out[1] = cycles per (warp_sum in warp 0..2, publish,
// barrier, everyone reads 3 partials), out[2] = same with a dependent sqrt + 2 reciprocals.
__global__ void k_barrier_latency(double* out) {
  __shared__ double rb[16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int N = 512;
  long long t0, t1;
  double x = 1.0 + threadIdx.x * 1e-6;
  t0 = clock64();
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = double(t1 - t0) / N;
  int par = 0;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    if (wid < 3) {
      double r = warp_sum(x);
      if (lane == 0) rb[par * 8 + wid] = r;
    }
    __syncthreads();
    x = (rb[par * 8] + rb[par * 8 + 1] + rb[par * 8 + 2]) * 1e-3 + 1.0;
    par ^= 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = double(t1 - t0) / N;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    if (wid < 3) {
      double r = warp_sum(x);
      if (lane == 0) rb[par * 8 + wid] = r;
    }
    __syncthreads();
    x = (rb[par * 8] + rb[par * 8 + 1] + rb[par * 8 + 2]) * 1e-3 + 1.0;
    x = sqrt(x) * fast_rcp(x + 1.0) + fast_rcp(x + 2.0);
    par ^= 1;
  }
  t1 = clock64();
  if (threadIdx.x == 0) { out[2] = double(t1 - t0) / N; out[3] = x; }
}

}  // namespace mop

// Launches blocks x 256 threads, each doing iters * 64 dependent-chain-interleaved DFMAs.
// FLOPs of one launch = blocks * 256 * iters * 64 * 2.
extern "C" int mop_priv_bench_dfma(int blocks, int iters, double* out, void* stream) {
  MOP_REQUIRE(blocks > 0 && iters > 0 && out, "mop_priv_bench_dfma: bad arguments");
  mop::k_dfma_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// Writes `count` doubles (use a buffer larger than the 126 MB L2 to flush it).
extern "C" int mop_priv_bench_fill(double* buf, size_t count, double value, void* stream) {
  MOP_REQUIRE(buf && count > 0, "mop_priv_bench_fill: bad arguments");
  mop::k_fill<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(buf, count, value);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// out[i] = fast_rcp(x[i]) (the division-free reciprocal used by the twisted factorisation)
extern "C" int mop_priv_fast_rcp(const double* x, double* out, size_t count, void* stream) {
  MOP_REQUIRE(x && out && count > 0, "mop_priv_fast_rcp: bad arguments");
  mop::k_fast_rcp<<<148, 256, 0, (cudaStream_t)stream>>>(x, out, count);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// out[0..7]: dependent-chain latency in cycles of DFMA, DADD, DMUL, LDS, SHFL.64, fast_rcp,
// sqrt(+add), div (one warp, nothing else running)
extern "C" int mop_priv_latency(double* out, void* stream) {
  MOP_REQUIRE(out, "mop_priv_latency: bad arguments");
  mop::k_latency<<<1, 32, 0, (cudaStream_t)stream>>>(out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_priv_barrier_latency(int threads, double* out, void* stream) {
  MOP_REQUIRE(out && threads >= 32 && threads <= 1024, "mop_priv_barrier_latency: bad arguments");
  mop::k_barrier_latency<<<1, threads, 0, (cudaStream_t)stream>>>(out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
