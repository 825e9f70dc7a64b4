// TR/ROT basis and gradient projection shared by the projection kernels (project.cu) and the fused
// update + projection + tridiagonalisation front end (tridiag_blocked.cu).
// References: Utils/calc_tools.py:249-316 (classical Gram-Schmidt with drop threshold), Optimizer/rsirfo.py:128-190
// and Optimizer/rsprfo.py:227-285 (reduced QR of the six raw vectors).
#pragma once
#include "common.cuh"

namespace mop {

// Build the TR/ROT basis of one structure into T[6][np]; returns rank k.
// Must be called by the whole CTA.  scratch >= 40 doubles.
static __device__ int build_trrot_basis(int n, const double* __restrict__ x, double* T, int np,
                                 double* raw, double* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = n / 3;
  // plain mean, sequential per component as calc_center does (calc_tools.py:138-145)
  __shared__ double cen[3];
  if (tid < 3) {
    double acc = 0.0;
    for (int a = 0; a < N; ++a) acc += x[3 * a + tid];
    cen[tid] = acc / N;
  }
  __syncthreads();
  for (int i = tid; i < n; i += nt) {
    const int a = i / 3, c = i - 3 * a;
    const double cx = x[3 * a] - cen[0], cy = x[3 * a + 1] - cen[1], cz = x[3 * a + 2] - cen[2];
    raw[0 * np + i] = (c == 0) ? 1.0 : 0.0;
    raw[1 * np + i] = (c == 1) ? 1.0 : 0.0;
    raw[2 * np + i] = (c == 2) ? 1.0 : 0.0;
    raw[3 * np + i] = (c == 0) ? 0.0 : (c == 1 ? -cz : cy);
    raw[4 * np + i] = (c == 0) ? cz : (c == 1 ? 0.0 : -cx);
    raw[5 * np + i] = (c == 0) ? -cy : (c == 1 ? cx : 0.0);
  }
  __syncthreads();
  int k = 0;
  for (int v = 0; v < 6; ++v) {
    // classical GS: coefficients from the ORIGINAL vector (calc_tools.py:252-256)
    double cf[6];
    for (int j = 0; j < k; ++j) {
      double p = 0.0;
      for (int i = tid; i < n; i += nt) p = fma(raw[v * np + i], T[j * np + i], p);
      cf[j] = block_sum(p, scratch);
    }
    double p2 = 0.0;
    for (int i = tid; i < n; i += nt) {
      double w = raw[v * np + i];
      for (int j = 0; j < k; ++j) w -= cf[j] * T[j * np + i];
      T[k * np + i] = w;
      p2 = fma(w, w, p2);
    }
    const double nrm = sqrt(block_sum(p2, scratch));
    if (nrm > 1e-10) {
      for (int i = tid; i < n; i += nt) T[k * np + i] /= nrm;
      ++k;
    }
    __syncthreads();
  }
  return k;
}

// Gradient projection for RANK-DEFICIENT TR/ROT sets, as the reference computes it: numpy.linalg.qr(A, 'reduced') of
// the 3N x 6 matrix of raw vectors is LAPACK dgeqrf (unblocked dgeqr2 for six columns) + dorgqr (dorg2r), i.e.
// K = min(3N, 6) Householder reflectors and ALWAYS K orthonormal columns.  A dependent column leaves a zero (or
// rounding-noise) residual: tau = 0 for an exact zero, and Q's column is the image of a unit vector under the earlier
// reflectors - not in the TR/ROT span.  rule 0 = RSIRFO._project_grad_tr_rot (rsirfo.py:172-188): every column is
// projected out (for two atoms Q is 6 x 6 orthogonal and the projected gradient is rounding noise); rule 1 =
// EnhancedRSPRFO._project_grad_tr_rot (rsprfo.py:244-285): fewer than three atoms -> gradient returned as is, columns
// with |R_jj| <= 1e-10 dropped.  A [6][np] holds the raw vectors and is destroyed.  Whole CTA; scratch >= 40 doubles.
static __device__ void project_grad_qr(int n, double* A, int np, const double* __restrict__ g, double* __restrict__ gp,
                                int rule, double* scratch) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (rule == 1 && n < 9) {
    for (int i = tid; i < n; i += nt) gp[i] = g[i];
    return;
  }
  const int K = n < 6 ? n : 6;
  double tau[6], rdiag[6];
  for (int j = 0; j < K; ++j) {  // dgeqr2: dlarfg on column j, then H_j applied to the columns to its right
    double* aj = A + (size_t)j * np;
    double p = 0.0;
    for (int i = j + 1 + tid; i < n; i += nt) p = fma(aj[i], aj[i], p);
    const double xnorm = sqrt(block_sum(p, scratch));
    const double alpha = aj[j];
    __syncthreads();
    double t = 0.0, beta = alpha;
    if (xnorm != 0.0) {
      beta = -copysign(hypot(alpha, xnorm), alpha);
      t = (beta - alpha) / beta;
      const double sc = 1.0 / (alpha - beta);
      for (int i = j + 1 + tid; i < n; i += nt) aj[i] *= sc;
    }
    tau[j] = t;
    rdiag[j] = beta;
    __syncthreads();
    if (t != 0.0)
      for (int c = j + 1; c < 6; ++c) {
        double* ac = A + (size_t)c * np;
        double q = 0.0;
        for (int i = j + 1 + tid; i < n; i += nt) q = fma(aj[i], ac[i], q);
        const double wc = t * (block_sum(q, scratch) + ac[j]);
        for (int i = j + 1 + tid; i < n; i += nt) ac[i] = fma(-wc, aj[i], ac[i]);
        __syncthreads();
        if (tid == 0) ac[j] -= wc;
        __syncthreads();
      }
  }
  for (int j = K - 1; j >= 0; --j) {  // dorg2r: Q = H_0 ... H_{K-1} (first K columns), built from the last reflector back
    double* aj = A + (size_t)j * np;
    const double t = tau[j];
    if (t != 0.0)
      for (int c = j + 1; c < K; ++c) {
        double* qc = A + (size_t)c * np;
        double q = 0.0;
        for (int i = j + 1 + tid; i < n; i += nt) q = fma(aj[i], qc[i], q);
        const double wc = t * (block_sum(q, scratch) + qc[j]);
        for (int i = j + 1 + tid; i < n; i += nt) qc[i] = fma(-wc, aj[i], qc[i]);
        __syncthreads();
        if (tid == 0) qc[j] -= wc;
        __syncthreads();
      }
    for (int i = tid; i < n; i += nt) aj[i] = (i > j) ? -t * aj[i] : (i == j ? 1.0 - t : 0.0);
    __syncthreads();
  }
  double cf[6];
  for (int j = 0; j < K; ++j) {
    double q = 0.0;
    for (int i = tid; i < n; i += nt) q = fma(A[(size_t)j * np + i], g[i], q);
    cf[j] = block_sum(q, scratch);
    if (rule == 1 && !(fabs(rdiag[j]) > 1e-10)) cf[j] = 0.0;
  }
  for (int i = tid; i < n; i += nt) {
    double part = 0.0;
    for (int j = 0; j < K; ++j) part = fma(A[(size_t)j * np + i], cf[j], part);
    gp[i] = g[i] - part;
  }
  __syncthreads();
}

}  // namespace mop
