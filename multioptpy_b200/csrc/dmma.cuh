// FP64 tensor-core tile (DMMA, mma.sync.m8n8k4) and the sixteen-value single-barrier block reduction shared by the
// blocked tridiagonalisations (tridiag_blocked.cu: one CTA per structure, n <= 160; tridiag_cluster.cu: one
// thread-block cluster per matrix, n <= 1024).
#pragma once
#include "common.cuh"

namespace mop {

// D = A B + C on the FP64 tensor cores: A 8 x 4 (row), B 4 x 8 (col), C / D 8 x 8.  Lane (g = lane / 4,
// t = lane % 4) holds A(g, t), B(t, g) and C(g, 2 t), C(g, 2 t + 1).
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

// Block-wide sums of SIXTEEN values, one barrier.  A transposing butterfly (16 + 8 + 4 + 2 + 2 shuffles instead
// of 16 x 10) leaves slot j with lanes 2 j, 2 j + 1; the per-warp partials of every slot are summed by every
// warp in fixed order (deterministic) and handed to all lanes through the warp's own row of `tot`.
// red: [2][16][NW | 1] (double-buffered by `parity`, so one barrier per call is enough; the odd row stride keeps the
// sixteen slot readers on different banks for every NW), tot: [NW][16].
__host__ __device__ constexpr int red16_doubles(int nw) { return 2 * 16 * (nw | 1); }
template <int NW>
__device__ __forceinline__ void block_sum16(double (&r)[16], double* red, double* tot, int& parity, int lane,
                                            int wid, bool contributes) {
  constexpr int NWP = NW | 1;
  double* bq = red + (parity & 1) * (16 * NWP);
  parity ^= 1;
  if (contributes) {  // warp-uniform
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    double t8[8], t4[4], t2[2], t1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double send = h16 ? r[j] : r[j + 8];
      const double keep = h16 ? r[j + 8] : r[j];
      t8[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double send = h8 ? t8[j] : t8[j + 4];
      const double keep = h8 ? t8[j + 4] : t8[j];
      t4[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const double send = h4 ? t4[j] : t4[j + 2];
      const double keep = h4 ? t4[j + 2] : t4[j];
      t2[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 4);
    }
    {
      const double send = h2 ? t2[0] : t2[1];
      const double keep = h2 ? t2[1] : t2[0];
      t1 = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 2);
    }
    t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 1);
    const int slot = lane >> 1;  // (h16 ? 8 : 0) + (h8 ? 4 : 0) + (h4 ? 2 : 0) + (h2 ? 1 : 0)
    if ((lane & 1) == 0) bq[slot * NWP + wid] = t1;
  } else if (lane < 16) {
    bq[lane * NWP + wid] = 0.0;
  }
  __syncthreads();
  if (!contributes) return;  // a warp without live rows needs no totals
  double* tw = tot + wid * 16;
  if (lane < 16) {
    double t[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) t[w] = bq[lane * NWP + w];
    double acc = t[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) acc += t[w];
    tw[lane] = acc;
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 16; q += 2) {
    const double2 v = *reinterpret_cast<const double2*>(tw + q);
    r[q] = v.x;
    r[q + 1] = v.y;
  }
  __syncwarp();
}

}  // namespace mop
