// Constrained RS-I-RFO: the subspace projection of CRSIRFO.run, Optimizer/crsirfo.py:16-45,
// 88-100.  The reference takes U = the null space of the normalised constraint rows from a full SVD, steps in the
// (n - r)-dimensional subspace (U^T H U, U^T g) and lifts the step back with U.  Every quantity of that step is
// invariant under the choice of U, so the kernel stays in the full space:
//     Q  = the r leading left singular vectors of the normalised rows C^T (singular values above
//          max(svd_threshold, 1e-6 s_max) - exactly the reference's rank rule), from the k x k Gram matrix
//          C C^T = V diag(s^2) V^T  (cyclic Jacobi, k <= 12):  Q_j = C^T v_j / s_j
//     gp = g - Q Q^T g            (= U U^T g)
//     Hp = P S P + sigma Q Q^T,   S = sym(H),  P = I - Q Q^T  (rank-r form  S - Y Q^T - Q Y^T,  Y = S Q - 1/2 Q (Q^T S Q))
// The constrained directions get the eigenvalue sigma = ||S||_F + 1 >= every eigenvalue of the subspace Hessian and
// zero gradient: they contribute nothing to the RFO step, never become the minimum eigenvalue that drives the
// adaptive trust radius (rsirfo.py:660-803), and the spectrum of Hp is the reference's subspace spectrum followed
// by r copies of sigma.  mop_rsirfo_spectral_step then IS the rest of CRSIRFO.run (crsirfo.py:101-170).
#include "common.cuh"

namespace mop {

constexpr int CR_KMAX = 12;
constexpr int CR_THREADS = 256;

__global__ void __launch_bounds__(256) k_add_inplace(size_t total, double* __restrict__ dst, const double* __restrict__ src) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) dst[e] += src[e];
}

// shared memory: C [k][np] | Q [k][np] | W -> Y [k][np] | gfull [np] | G, V [k][k] each | sv [k] | scratch 64
__host__ __device__ inline size_t cr_smem_bytes(int n, int k) {
  const size_t np = (size_t)((n + 3) & ~3);
  return sizeof(double) * (3 * k * np + np + 2 * (size_t)k * k + k + 64);
}

__global__ void __launch_bounds__(CR_THREADS)
k_constraint_project(int n, int k, double svd_thr, const double* __restrict__ C_all, const double* __restrict__ H_all,
                     const double* __restrict__ g_all, const double* __restrict__ shake_all, double* __restrict__ Hp_all,
                     double* __restrict__ gp_all, int32_t* __restrict__ rank_out) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = CR_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* C = sm;                       // normalised constraint rows
  double* Q = C + (size_t)k * np;       // orthonormal basis of their span (r rows used)
  double* W = Q + (size_t)k * np;
  double* gf = W + (size_t)k * np;      // gradient incl. the SHAKE transport term
  double* G = gf + np;                  // k x k Gram matrix -> rotated in place
  double* V = G + k * k;                // eigenvectors (columns)
  double* sv = V + k * k;               // singular values
  double* scratch = sv + k;             // 64
  __shared__ int s_rank;
  __shared__ double s_sigma;
  const double* H = H_all + (size_t)b * n * n;
  const double* g = g_all + (size_t)b * n;

  // ---- normalised rows (crsirfo.py:25-27: norms below 1e-12 divide by 1) ----
  for (int e = tid; e < k * np; e += CR_THREADS) {
    const int a = e / np, i = e - a * np;
    C[e] = i < n ? C_all[((size_t)b * k + a) * n + i] : 0.0;
    Q[e] = 0.0;
    W[e] = 0.0;
  }
  __syncthreads();
  for (int a = w; a < k; a += nw) {
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(C[a * np + i], C[a * np + i], p);
    p = sqrt(warp_sum(p));
    const double inv = 1.0 / (p < 1e-12 ? 1.0 : p);
    for (int i = lane; i < n; i += 32) C[a * np + i] *= inv;
  }
  __syncthreads();
  for (int e = w; e < k * k; e += nw) {
    const int a = e / k, c = e - a * k;
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(C[a * np + i], C[c * np + i], p);
    p = warp_sum(p);
    if (lane == 0) G[a * k + c] = p;
  }
  __syncthreads();
  // ---- k x k symmetric eigenproblem: cyclic Jacobi on one thread (k <= 12) ----
  if (tid == 0) {
    for (int a = 0; a < k; ++a)
      for (int c = 0; c < k; ++c) V[a * k + c] = a == c ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
      double off = 0.0;
      for (int p = 0; p < k; ++p)
        for (int q = p + 1; q < k; ++q) off += G[p * k + q] * G[p * k + q];
      if (!(off > 1e-300)) break;
      for (int p = 0; p < k; ++p)
        for (int q = p + 1; q < k; ++q) {
          const double apq = G[p * k + q];
          if (apq == 0.0) continue;
          const double th = (G[q * k + q] - G[p * k + p]) / (2.0 * apq);
          const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
          const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
          for (int r = 0; r < k; ++r) {
            const double grp = G[r * k + p], grq = G[r * k + q];
            G[r * k + p] = cs * grp - sn * grq;
            G[r * k + q] = sn * grp + cs * grq;
          }
          for (int r = 0; r < k; ++r) {
            const double gpr = G[p * k + r], gqr = G[q * k + r];
            G[p * k + r] = cs * gpr - sn * gqr;
            G[q * k + r] = sn * gpr + cs * gqr;
          }
          for (int r = 0; r < k; ++r) {
            const double vrp = V[r * k + p], vrq = V[r * k + q];
            V[r * k + p] = cs * vrp - sn * vrq;
            V[r * k + q] = sn * vrp + cs * vrq;
          }
        }
    }
    double smax = 0.0;
    for (int a = 0; a < k; ++a) {
      sv[a] = sqrt(fmax(G[a * k + a], 0.0));
      smax = fmax(smax, sv[a]);
    }
    const double thr = fmax(svd_thr, smax * 1e-6);   // crsirfo.py:31-33 (max_s = 1 for an empty set)
    int r = 0;
    for (int a = 0; a < k; ++a)
      if (sv[a] > thr) {
        // compact the retained (value, vector) pairs to the front
        sv[r] = sv[a];
        for (int c = 0; c < k; ++c) V[c * k + r] = V[c * k + a];
        ++r;
      }
    s_rank = r;
  }
  __syncthreads();
  const int r = s_rank;
  if (tid == 0 && rank_out) rank_out[b] = r;
  // Q_j = sum_a V[a][j] C_a / s_j
  for (int e = tid; e < r * np; e += CR_THREADS) {
    const int j = e / np, i = e - j * np;
    double acc = 0.0;
    for (int a = 0; a < k; ++a) acc = fma(V[a * k + j], C[a * np + i], acc);
    Q[e] = acc / sv[j];
  }
  // ---- gradient: SHAKE transport term g + H delta (crsirfo.py:70-80: only when |delta| > 1e-6), then projection ----
  bool use_shake = false;
  if (shake_all) {
    const double* dl = shake_all + (size_t)b * n;
    double p = 0.0;
    for (int i = tid; i < n; i += CR_THREADS) p = fma(dl[i], dl[i], p);
    use_shake = sqrt(block_sum(p, scratch)) > 1e-6;
  }
  for (int i = tid; i < n; i += CR_THREADS) gf[i] = g[i];
  __syncthreads();
  if (use_shake) {
    const double* dl = shake_all + (size_t)b * n;
    for (int i = w; i < n; i += nw) {
      double p = 0.0;
      for (int j = lane; j < n; j += 32) p = fma(H[(size_t)i * n + j], dl[j], p);
      p = warp_sum(p);
      if (lane == 0) gf[i] += p;
    }
    __syncthreads();
  }
  {
    double cf[CR_KMAX];
    for (int j = 0; j < r; ++j) {
      double p = 0.0;
      for (int i = tid; i < n; i += CR_THREADS) p = fma(Q[j * np + i], gf[i], p);
      cf[j] = block_sum(p, scratch);
    }
    for (int i = tid; i < n; i += CR_THREADS) {
      double part = 0.0;
      for (int j = 0; j < r; ++j) part = fma(Q[j * np + i], cf[j], part);
      gp_all[(size_t)b * n + i] = gf[i] - part;
    }
  }
  if (!Hp_all) return;
  __syncthreads();
  // ---- W = S Q (row pass + column pass of H), ||S||_F ----
  double fro = 0.0;
  for (int i = w; i < n; i += nw) {
    double acc[CR_KMAX];
#pragma unroll
    for (int v = 0; v < CR_KMAX; ++v) acc[v] = 0.0;
    for (int j = lane; j < n; j += 32) {
      const double a = H[(size_t)i * n + j];
      const double s = 0.5 * (a + H[(size_t)j * n + i]);
      fro = fma(s, s, fro);
#pragma unroll
      for (int v = 0; v < CR_KMAX; ++v)
        if (v < r) acc[v] = fma(s, Q[v * np + j], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < CR_KMAX; ++v)
      if (v < r) {
        const double t = warp_sum(acc[v]);
        if (lane == 0) W[v * np + i] = t;
      }
  }
  fro = block_sum(fro, scratch);
  if (tid == 0) s_sigma = sqrt(fro) + 1.0;
  __syncthreads();
  // M = Q^T W (r x r) in G, Y = W - 1/2 Q sym(M)
  for (int e = w; e < r * r; e += nw) {
    const int a = e / r, c = e - a * r;
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(Q[a * np + i], W[c * np + i], p);
    p = warp_sum(p);
    if (lane == 0) G[a * k + c] = p;
  }
  __syncthreads();
  for (int i = tid; i < n; i += CR_THREADS) {
    double qv[CR_KMAX];
    for (int a = 0; a < r; ++a) qv[a] = Q[a * np + i];
    for (int c = 0; c < r; ++c) {
      double corr = 0.0;
      for (int a = 0; a < r; ++a) corr = fma(qv[a], 0.5 * (G[a * k + c] + G[c * k + a]), corr);
      W[c * np + i] -= 0.5 * corr;
    }
  }
  __syncthreads();
  // ---- Hp = S - Y Q^T - Q Y^T + sigma Q Q^T (bit-symmetric) ----
  const double sigma = s_sigma;
  double* Hp = Hp_all + (size_t)b * n * n;
  for (int i = w; i < n; i += nw)
    for (int j = lane; j < n; j += 32) {
      double v = 0.5 * (H[(size_t)i * n + j] + H[(size_t)j * n + i]);
      for (int a = 0; a < r; ++a) {
        v -= __dadd_rn(__dmul_rn(W[a * np + i], Q[a * np + j]), __dmul_rn(Q[a * np + i], W[a * np + j]));
        v = __dadd_rn(v, __dmul_rn(sigma, __dmul_rn(Q[a * np + i], Q[a * np + j])));
      }
      Hp[(size_t)i * n + j] = v;
    }
}

// The explicit convergence test of the subspace gradient (crsirfo.py:108-118): below the threshold the reference
// returns a zero step BEFORE its trust-radius / energy bookkeeping and only records the current point as "previous".
__global__ void __launch_bounds__(128)
k_crsirfo_finalize(int n, double thr, const double* __restrict__ gp_all, const double* __restrict__ Be,
                   const double* __restrict__ state_before, double* __restrict__ state, double* __restrict__ move,
                   double* __restrict__ pred, int32_t* __restrict__ status) {
  __shared__ double scratch[40];
  const int b = blockIdx.x, tid = threadIdx.x;
  double p = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double v = gp_all[(size_t)b * n + i];
    p = fma(v, v, p);
  }
  const double nrm = sqrt(block_sum(p, scratch));
  if (!(nrm < thr)) return;
  for (int i = tid; i < n; i += blockDim.x) move[(size_t)b * n + i] = 0.0;
  for (int i = tid; i < MOP_RSIRFO_STATE; i += blockDim.x) {
    double v = state_before[(size_t)b * MOP_RSIRFO_STATE + i];
    if (i == MOP_RS_HAVE_PREV) v = 1.0;
    if (i == MOP_RS_PREV_ENERGY) v = Be ? Be[b] : 0.0;
    if (i == MOP_RS_HAVE_ENERGY) v = 1.0;
    state[(size_t)b * MOP_RSIRFO_STATE + i] = v;
  }
  if (tid == 0) {
    if (pred) pred[b] = 0.0;
    if (status) status[b] = MOP_ST_CONSTR_CONVERGED;
  }
}

}  // namespace mop

extern "C" int mop_add_inplace(size_t count, double* dst, const double* src, void* stream) {
  if (count == 0) return MOP_OK;
  MOP_REQUIRE(dst && src, "mop_add_inplace: null pointer");
  const size_t blocks = (count + 255) / 256;
  mop::k_add_inplace<<<(int)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(count, dst, src);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_constraint_project(int B, int n, int k, double svd_threshold, const double* C, const double* H,
                                      const double* g, const double* shake, double* Hp_out, double* gp_out,
                                      int32_t* rank_out, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && k >= 1 && k <= mop::CR_KMAX, "mop_constraint_project: 1 <= k <= 12 constraint rows required");
  if (B == 0) return MOP_OK;   // (an empty batch has no buffers)
  MOP_REQUIRE(C && H && g && gp_out, "mop_constraint_project: C, H, g, gp_out required");
  const size_t smem = mop::cr_smem_bytes(n, k);
  if (smem > 220 * 1024) {
    mop_set_error("mop_constraint_project: n = %d with k = %d rows needs %zu bytes of shared memory", n, k, smem);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_constraint_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_constraint_project<<<B, mop::CR_THREADS, smem, (cudaStream_t)stream>>>(n, k, svd_threshold, C, H, g, shake, Hp_out,
                                                                              gp_out, rank_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_crsirfo_finalize(int B, int n, double grad_threshold, const double* gp, const double* Be,
                                    const double* state_before, double* state, double* move, double* pred,
                                    int32_t* status, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_crsirfo_finalize: bad arguments");
  if (B == 0) return MOP_OK;
  MOP_REQUIRE(gp && state_before && state && move, "mop_crsirfo_finalize: null pointer");
  mop::k_crsirfo_finalize<<<B, 128, 0, (cudaStream_t)stream>>>(n, grad_threshold, gp, Be, state_before, state, move, pred, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
