// Sturm-sequence eigenvalue count for symmetric tridiagonal matrices, shared by the
// shared-memory eigensolver (eigh_tridiag.cu) and the large-n path (eigh_large.cu).
#pragma once
#include "common.cuh"

namespace mop {

constexpr double TRI_EPS = 2.220446049250313e-16;
constexpr double TRI_GAPTOL = 1e-3;  // cluster gap relative to ||T|| (LAPACK dstein ORTOL)

// number of eigenvalues of the unreduced block rows [s, t) that are < x
// (sign changes of p_k = (d_k - x) p_{k-1} - e_{k-1}^2 p_{k-2}); d, e2 scaled to ||T|| <= 1.
// The dependent chain is one DFMA per row: (d_k - x) and the loads are hoisted four rows
// ahead, signs are compared on the high word, and the magnitude is checked (and both
// iterates rescaled) once per four rows.  An exact zero needs no special case: whichever
// sign it is given, the pair (k-1, k+1) contributes exactly one sign change.
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2,
                                           int s, int t, double x) {
  double pm1 = 1.0;
  double p = d[s] - x;
  int cnt = (unsigned)__double2hiint(p) >> 31;
  int k = s + 1;
  for (; k + 3 < t; k += 4) {
    const double a0 = d[k] - x, a1 = d[k + 1] - x, a2 = d[k + 2] - x, a3 = d[k + 3] - x;
    const double b0 = e2[k - 1], b1 = e2[k], b2 = e2[k + 1], b3 = e2[k + 2];
    const double p0 = fma(a0, p, -(b0 * pm1));
    const double p1 = fma(a1, p0, -(b1 * p));
    const double p2 = fma(a2, p1, -(b2 * p0));
    const double p3 = fma(a3, p2, -(b3 * p1));
    const int h = __double2hiint(p), h0 = __double2hiint(p0), h1 = __double2hiint(p1),
              h2 = __double2hiint(p2), h3 = __double2hiint(p3);
    cnt += ((unsigned)(h ^ h0) >> 31) + ((unsigned)(h0 ^ h1) >> 31) + ((unsigned)(h1 ^ h2) >> 31) +
           ((unsigned)(h2 ^ h3) >> 31);
    pm1 = p2;
    p = p3;
    const unsigned ex = ((unsigned)h3 >> 20) & 0x7ffu;
    if (ex - 523u > 1000u) {  // |p| outside [2^-500, 2^500]: rare, even per warp
      const double a = fmax(fabs(p), fabs(pm1));
      if (a > 0.0 && a < INFINITY) {
        const int ea = (__double2hiint(a) >> 20) & 0x7ff;          // biased exponent of a
        const double sc = __hiloint2double((2046 - ea) << 20, 0);   // 2^(1023 - ea)
        p *= sc;
        pm1 *= sc;
      }
    }
  }
  for (; k < t; ++k) {
    const double pn = fma(d[k] - x, p, -(e2[k - 1] * pm1));
    cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31;
    pm1 = p;
    p = pn;
  }
  return cnt;
}

// The same count from the INTERLEAVED table de[k] = (d_k, e_{k-1}^2) (one 16-byte shared load per row),
// eight rows per magnitude check, also returning the last term of the sequence, p_n(x) = det(T - x I) up
// to the positive scale 2^pexp, as (pval, pexp): the characteristic polynomial value that drives the
// superlinear root-finder of k_spectrum_step.  sign(pval) == (-1)^count by construction.
// Magnitude window [2^-300, 2^400]: with |d|, |e| <= 1 and |x| <= 3 eight rows grow |p| by at most 5^8 < 2^19
// and shrink it by at most (eps^2)^4 > 2^-440 (off-diagonals below eps (|d_i| + |d_i+1|) were split off).
__device__ __forceinline__ int sturm_eval(const double2* __restrict__ de, int s, int t, double x, double* pval,
                                          int* pexp) {
  double pm1 = 1.0;
  double p = de[s].x - x;
  int cnt = (unsigned)__double2hiint(p) >> 31;
  int esum = 0;
  int k = s + 1;
  for (; k + 7 < t; k += 8) {
    // the nine signs (the iterate before the chunk, then the eight new ones) are shifted into one word - one
    // funnel shift per row - and the sign changes counted once per chunk: popc(w ^ (w >> 1)) over eight pairs
    unsigned sg = (unsigned)__double2hiint(p) >> 31;
    int hprev = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double2 q = de[k + u];
      const double pn = fma(q.x - x, p, -(q.y * pm1));
      hprev = __double2hiint(pn);
      sg = __funnelshift_l((unsigned)hprev, sg, 1);
      pm1 = p;
      p = pn;
    }
    cnt += __popc((sg ^ (sg >> 1)) & 0xffu);
    const unsigned ex = ((unsigned)hprev >> 20) & 0x7ffu;
    if (ex - 723u > 700u) {  // |p| outside [2^-300, 2^400]: rare, even per warp
      const double a = fmax(fabs(p), fabs(pm1));
      if (a > 0.0 && a < INFINITY) {
        const int ea = (__double2hiint(a) >> 20) & 0x7ff;          // biased exponent of a
        const double sc = __hiloint2double((2046 - ea) << 20, 0);   // 2^(1023 - ea)
        p *= sc;
        pm1 *= sc;
        esum += ea - 1023;
      }
    }
  }
  for (; k < t; ++k) {
    const double2 q = de[k];
    const double pn = fma(q.x - x, p, -(q.y * pm1));
    cnt += (unsigned)(__double2hiint(pn) ^ __double2hiint(p)) >> 31;
    pm1 = p;
    p = pn;
  }
  *pval = p;
  *pexp = esum;
  return cnt;
}

}  // namespace mop
