// Batched symmetric eigensolver, robust path: two-sided cyclic Jacobi with a
// round-robin (tournament) parallel ordering.  One CTA per matrix; the matrix
// lives in shared memory (n <= JAC_SMEM_MAX_N) or in the caller's workspace.
// Replaces numpy.linalg.eigh at Optimizer/rsirfo.py:606,626,652 when the fast
// tridiagonal path (eigh_tridiag.cu) flags a structure, and serves as the
// cross-check eigensolver in the tests.  FP64-pipe / shared-memory bound.
#include "common.cuh"

namespace mop {

constexpr int JAC_THREADS = 512;
constexpr int JAC_MAX_SWEEPS = 40;

// Pair k of round r in a tournament over m (even) players: player m-1 is fixed.
__device__ __forceinline__ void tournament_pair(int m, int r, int k, int& p, int& q) {
  const int mm = m - 1;
  int a, b;
  if (k == 0) {
    a = mm;
    b = r % mm;
  } else {
    a = (r + k) % mm;
    b = (r - k + mm) % mm;
  }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

// A: m x lda working matrix (shared or global), Vt: n x n (global), row k = vector k.
__global__ void __launch_bounds__(JAC_THREADS)
k_eigh_jacobi(int n, int use_smem, const double* __restrict__ Ain, double* __restrict__ Awork_all,
              double* __restrict__ Vwork_all, double* __restrict__ evals_all,
              double* __restrict__ evecs_all, int32_t* __restrict__ status,
              const int32_t* __restrict__ only_flagged) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (only_flagged && !(only_flagged[b] & MOP_ST_EIG_FALLBACK)) return;
  const int m = (n + 1) & ~1;
  const int half = m >> 1;
  const int lda = m | 1;
  double* scratch = sm;            // 40
  double* rc = scratch + 40;       // half
  double* rs = rc + half;          // half
  int* rp = (int*)(rs + half);     // half
  int* rq = rp + half;             // half
  int* rank_of = rq + half;        // m ints (m even -> 8-byte aligned end)
  double* A = use_smem ? (double*)(rank_of + m) : Awork_all + (size_t)b * m * lda;
  double* Vt = Vwork_all + (size_t)b * n * n;
  const double* Asrc = Ain + (size_t)b * n * n;
  __shared__ int s_rot;

  // load (symmetrised), zero the dummy row/column, V = I
  double pn = 0.0;
  for (int e = tid; e < m * m; e += JAC_THREADS) {
    const int i = e / m, j = e - i * m;
    double v = 0.0;
    if (i < n && j < n) v = 0.5 * (Asrc[(size_t)i * n + j] + Asrc[(size_t)j * n + i]);
    A[i * lda + j] = v;
    pn = fma(v, v, pn);
  }
  for (int e = tid; e < n * n; e += JAC_THREADS) {
    const int i = e / n, j = e - i * n;
    Vt[e] = (i == j) ? 1.0 : 0.0;
  }
  const double fro = sqrt(block_sum(pn, scratch));
  const double thr = fro * 1.1102230246251565e-16;  // 2^-53 ||A||_F
  int st = 0;
  bool finite = isfinite(fro);

  int sweep = 0;
  bool converged = !finite || fro == 0.0;
  while (!converged && sweep < JAC_MAX_SWEEPS) {
    int rotated_in_sweep = 0;
    for (int r = 0; r < m - 1; ++r) {
      if (tid == 0) s_rot = 0;
      __syncthreads();
      if (tid < half) {
        int p, q;
        tournament_pair(m, r, tid, p, q);
        const double apq = A[p * lda + q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > thr) {
          const double theta = (A[q * lda + q] - A[p * lda + p]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
          c = 1.0 / sqrt(fma(t, t, 1.0));
          s = t * c;
          s_rot = 1;
        }
        rc[tid] = c;
        rs[tid] = s;
        rp[tid] = p;
        rq[tid] = q;
      }
      __syncthreads();
      if (s_rot) {
        rotated_in_sweep = 1;
        // A <- J^T A J, one 2x2 block per work item
        for (int e = tid; e < half * half; e += JAC_THREADS) {
          const int I = e / half, J = e - I * half;
          const double cI = rc[I], sI = rs[I], cJ = rc[J], sJ = rs[J];
          if (sI == 0.0 && sJ == 0.0) continue;
          const int pI = rp[I], qI = rq[I], pJ = rp[J], qJ = rq[J];
          const double a = A[pI * lda + pJ], bb = A[pI * lda + qJ];
          const double cc = A[qI * lda + pJ], d = A[qI * lda + qJ];
          // rows: row_p' = c row_p - s row_q ; row_q' = s row_p + c row_q
          const double a1 = cI * a - sI * cc, b1 = cI * bb - sI * d;
          const double c1 = sI * a + cI * cc, d1 = sI * bb + cI * d;
          // cols: col_p' = c col_p - s col_q ; col_q' = s col_p + c col_q
          double a2 = cJ * a1 - sJ * b1, b2 = sJ * a1 + cJ * b1;
          double c2 = cJ * c1 - sJ * d1, d2 = sJ * c1 + cJ * d1;
          if (I == J) {
            b2 = 0.0;
            c2 = 0.0;
          }
          A[pI * lda + pJ] = a2;
          A[pI * lda + qJ] = b2;
          A[qI * lda + pJ] = c2;
          A[qI * lda + qJ] = d2;
        }
        // V <- V J  (rows of Vt)
        for (int e = tid; e < half * n; e += JAC_THREADS) {
          const int J = e / n, i = e - J * n;
          const double s = rs[J];
          if (s == 0.0) continue;
          const int p = rp[J], q = rq[J];
          if (q >= n) continue;
          const double c = rc[J];
          const double vp = Vt[(size_t)p * n + i], vq = Vt[(size_t)q * n + i];
          Vt[(size_t)p * n + i] = c * vp - s * vq;
          Vt[(size_t)q * n + i] = s * vp + c * vq;
        }
      }
      __syncthreads();
    }
    ++sweep;
    if (!rotated_in_sweep) converged = true;
  }
  if (!converged) st |= MOP_ST_EIG_NOCONV;

  // eigenvalues = diagonal; ascending rank sort (stable by index)
  __syncthreads();
  double* evals = evals_all + (size_t)b * n;
  double* evecs = evecs_all + (size_t)b * n * n;
  for (int k = tid; k < n; k += JAC_THREADS) {
    const double lk = A[k * lda + k];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double lj = A[j * lda + j];
      rank += (lj < lk) || (lj == lk && j < k);
    }
    if (!finite) rank = k;  // NaN/Inf input: keep positions, caller falls back
    evals[rank] = lk;
    rank_of[k] = rank;
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += JAC_THREADS) {
    const int k = e / n, i = e - k * n;
    evecs[(size_t)rank_of[k] * n + i] = Vt[e];
  }
  if (tid == 0 && status) {
    int s0 = status[b] & ~(MOP_ST_EIG_NOCONV);
    status[b] = s0 | st;
  }
}

}  // namespace mop

int mop_jacobi_smem_max_n() { return 164; }

size_t mop_jacobi_workspace_bytes(int B, int n) {
  const int m = (n + 1) & ~1, lda = m | 1;
  size_t bytes = (size_t)B * n * n * sizeof(double);  // V accumulation
  if (n > mop_jacobi_smem_max_n()) bytes += (size_t)B * m * lda * sizeof(double);
  return bytes;
}

// awork_ext (optional, >= B * m * (m | 1) doubles with m = n rounded up to even): keep the working matrix
// there instead of in shared memory.  The fallback launches that follow the fast eigensolvers pass it:
// with 180 KB of shared memory per CTA the (almost always empty) launch could not share an SM with the
// kernels of other streams and stalled its stream until whole SMs drained.
int mop_launch_eigh_jacobi_ext(int B, int n, const double* A, double* evals, double* evecs,
                               int32_t* status, const int32_t* only_flagged, void* work,
                               size_t work_bytes, double* awork_ext, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_jacobi_workspace_bytes(B, n) || !work) {
    mop_set_error("eigh (jacobi): workspace too small (%zu < %zu)", work_bytes,
                  mop_jacobi_workspace_bytes(B, n));
    return MOP_ERR_WORKSPACE;
  }
  const int m = (n + 1) & ~1, half = m >> 1, lda = m | 1;
  const int use_smem = n <= mop_jacobi_smem_max_n() && !awork_ext;
  size_t smem = sizeof(double) * (40 + 2 * (size_t)half) + sizeof(int) * (2 * (size_t)half + m);
  if (use_smem) smem += sizeof(double) * (size_t)m * lda;
  double* Vwork = (double*)work;
  double* Awork = use_smem ? nullptr : (awork_ext ? awork_ext : Vwork + (size_t)B * n * n);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_eigh_jacobi,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_eigh_jacobi<<<B, mop::JAC_THREADS, smem, stream>>>(n, use_smem, A, Awork, Vwork, evals,
                                                           evecs, status, only_flagged);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int mop_launch_eigh_jacobi(int B, int n, const double* A, double* evals, double* evecs,
                           int32_t* status, const int32_t* only_flagged, void* work,
                           size_t work_bytes, cudaStream_t stream) {
  return mop_launch_eigh_jacobi_ext(B, n, A, evals, evecs, status, only_flagged, work, work_bytes, nullptr, stream);
}
