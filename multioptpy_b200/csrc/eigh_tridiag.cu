// Fast batched symmetric eigensolver and the fused RS-I-RFO step built on it.
//
// One CTA per matrix, the matrix resident in shared memory (n <= TRI_MAX_N):
//   1. Householder tridiagonalisation  A = Q T Q^T  (LAPACK dsytd2 conventions),
//      reflectors kept in the rows of the working matrix and spilled to global
//      scratch (L2) once T is known;
//   2. T is scaled to unit norm and split at negligible off-diagonals; every
//      thread-row i owns eigenvalue (i - block_start) of its block and finds it by
//      multisection on the Sturm count (division-free p-recurrence with rescaling);
//   3. eigenvectors of T by twisted factorisation (Fernando / Parlett-Dhillon), one
//      thread per vector, pivots staged in the shared-memory matrix that then
//      holds Z (column i = vector i);
//   4. clusters (gap < 1e-3 ||T||) are re-orthogonalised (CGS2), one warp per cluster;
//      a vector that cancels against its cluster flags the structure for the
//      robust Jacobi path (MOP_ST_EIG_FALLBACK) — correctness never depends on luck;
//   5a. MOP eigh API: V = Q Z written out (rows = eigenvectors, ascending);
//   5b. fused RS-I-RFO: gamma = Z^T (Q^T gp), the step is solved in the eigenbasis
//       (rfo_core.cuh) and transformed back as Q (Z c); V is never formed, so the
//       O(n^3) work is the 4/3 n^3 of the reduction only.
// Replaces numpy.linalg.eigh (LAPACK dsyevd) at Optimizer/rsirfo.py:606,626,652.
#include "rfo_core.cuh"
#include "tri_sturm.cuh"

namespace mop {

constexpr int TRI_MAX_N = 160;
constexpr int TRI_GMAX = 16;         // max row-groups in the column-split symv / update

// size (in doubles) of the phase-aliased scratch region X
__host__ __device__ inline size_t tri_x_doubles(int n) {
  const size_t np = (size_t)((n + 3) & ~3);
  size_t x = (size_t)(TRI_GMAX + 2) * np + 48;                               // phase 1 / 2
  const size_t fused = 2 * np + rfo_core_smem_bytes(n) / sizeof(double) + 8;  // fused RFO arrays
  if (fused > x) x = fused;
  if (32 * 64 > x) x = 32 * 64;                                              // phase 4 dots (<= 32 warps)
  return (x + 1) & ~(size_t)1;
}

struct TriArgs {
  int n;
  int fused;  // 0: eigh (evals/evecs out)   1: fused RS-I-RFO step
  const double* A;   // [B][n][n] input matrix (projected Hessian in fused mode)
  double* Vh;        // [B][n][n] scratch: reflector rows
  double* Dm;        // [B][n][n] scratch: backward reciprocal pivots
  double* evals;     // [B][n] out (ascending)
  double* evecs;     // [B][n][n] out (eigh mode), row k = vector k
  int32_t* status;
  // fused mode
  const double* gp;  // [B][n] projected gradient
  const double* Bg;  // [B][n] raw biased gradient (norm only)
  const double* Be;  // [B]
  double* state;     // [B][MOP_RSIRFO_STATE]
  double* move;      // [B][n]
  double* pred;      // [B]
  int saddle_order, neb_mode;
  double tmin, tmax;
  long long* dbg;    // optional [B][8] phase clocks (diagnostics)
  int ablate;        // diagnostics only: bit0 skip symv loop, bit1 skip update loop (results invalid)
  // prefactored mode: T, tau, Q^T gp and the reflector rows (in Vh) come from k_tridiag_packed
  const double* pf_d;
  const double* pf_e;
  const double* pf_tau;
  const double* pf_gq;
  const int* pf_flag;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_eigh_tridiag(TriArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int np = (n + 3) & ~3;
  const int lds = n | 1;
  // ---- shared memory carve-up --------------------------------------------------
  double* S = sm;                       // n x lds : A, then reciprocal pivots, then Z
  double* d = S + (size_t)n * lds;      // np
  double* e = d + np;                   // np   e[k] couples k, k+1
  double* e2 = e + np;                  // np
  double* tau = e2 + np;                // np
  double* lam = tau + np;               // np   eigenvalue of thread-row i (scaled)
  double* gq = lam + np;                // np   Q^T gp, later y = Z c
  double* X = gq + np;                  // phase scratch (aliased):
  //   phase 1: v[np], w[np], part[TRI_GMAX*np]
  //   phase 2: lo[np], hi[np], cnt[3*np] ints
  //   phase 3: nrm2 parts [2*np], twist index [np] ints
  //   phase 4: dots (NW * 64 doubles)
  //   fused  : lam_s[np], gam_s[np], RfoArrays
  int* blk_s = (int*)(X + tri_x_doubles(n));  // np
  int* blk_e = blk_s + np;              // np
  int* cl_s = blk_e + np;               // np   cluster start of i
  int* rank = cl_s + np;                // np   ascending rank of thread-row i
  int* inv = rank + np;                 // np   inverse permutation
  __shared__ double s_red[40];
  __shared__ double s_rbuf[2 * 32 * 2];   // reduction (C)
  __shared__ double s_rbuf1[2 * 32];      // reduction (E)
  __shared__ double s_tnorm;
  __shared__ int s_fallback;
  int parity = 0;

  const double* Ain = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;
  double* Dm = a.Dm + (size_t)b * n * n;
  int st_in = a.status ? a.status[b] : 0;
  st_in &= ~(MOP_ST_EIG_FALLBACK | MOP_ST_EIG_NOCONV);

  long long t_prev = clock64();
  int t_slot = 0;
#define TRI_MARK()                                                      \
  do {                                                                  \
    if (a.dbg && tid == 0) {                                            \
      const long long t_now = clock64();                                \
      a.dbg[(size_t)b * 16 + (t_slot++)] = t_now - t_prev;               \
      t_prev = t_now;                                                   \
    }                                                                   \
  } while (0)

  // ---- load.  Fused mode: the projection kernel writes a bit-symmetric matrix, read it
  // row-wise (coalesced); eigh mode symmetrises arbitrary input. -------------------------
  const bool prefactored = a.pf_d != nullptr;
  double pn = 0.0;
  if (prefactored) {
    for (int i = tid; i < n; i += THREADS) {
      d[i] = a.pf_d[(size_t)b * n + i];
      e[i] = a.pf_e[(size_t)b * n + i];
      tau[i] = a.pf_tau[(size_t)b * n + i];
      gq[i] = a.pf_gq[(size_t)b * n + i];
    }
  } else if (a.fused) {
    for (int idx = tid; idx < n * n; idx += THREADS) {
      const int i = idx / n, j = idx - i * n;
      const double v = Ain[idx];
      S[i * lds + j] = v;
      pn = fma(v, v, pn);
    }
  } else {
    for (int idx = tid; idx < n * n; idx += THREADS) {
      const int i = idx / n, j = idx - i * n;
      const double v = 0.5 * (Ain[idx] + Ain[(size_t)j * n + i]);
      S[i * lds + j] = v;
      pn = fma(v, v, pn);
    }
  }
  if (a.fused && !prefactored)
    for (int i = tid; i < n; i += THREADS) gq[i] = a.gp[(size_t)b * n + i];
  if (tid == 0) s_fallback = 0;
  double fro = sqrt(block_sum(pn, s_red));
  if (prefactored) {
    const int fl = a.pf_flag[b];
    fro = fl == 2 ? NAN : (fl == 1 ? 0.0 : 1.0);
  }
  const bool finite_in = isfinite(fro);
  bool identity = !finite_in;   // non-finite input: rsirfo.py:365-369 identity fallback
  const bool trivial = identity || fro == 0.0;

  double* v = X;
  double* w = X + np;
  double* part = X + 2 * np;
  TRI_MARK();  // 0: load

  // ---- phase 1: tridiagonalisation -------------------------------------------------------
  // Column step k: v from row k, p = tau A22 v, w = p - tau/2 (p.v) v, A22 -= v w^T + w v^T.
  // Thread (jc, q) owns the two columns jA = k+1+jc, jB = jA + cols2 and the rows
  // i = k+1+q, +G, ... ; block reductions are done by the warps that hold the q == 0
  // threads only (<= 3 warps), everyone else just adds their published partials.
  if (!trivial && n > 2 && !prefactored) {
    double* rb = s_rbuf;  // [2][8] partials: (value, warp)
    int par = 0;
    double xn2;
    {
      double r = 0.0;
      for (int j = 2 + tid; j < n; j += THREADS) r = fma(S[j], S[j], r);
      xn2 = block_sum(r, s_red);
    }
    int cols2 = -1, G = 1, jc = 0, q = 0;
    long long seg[5] = {0, 0, 0, 0, 0}, ts = clock64();
#define SEG(i)                              \
  do {                                      \
    if (a.dbg) {                            \
      const long long tn_ = clock64();      \
      seg[i] += tn_ - ts;                   \
      ts = tn_;                             \
    }                                       \
  } while (0)
    double* P = w;  // p = tau A22 v (w_i = p_i + alpha2 v_i is formed on the fly)
    for (int k = 0; k < n - 2; ++k) {
      double* ak = S + k * lds;
      const double alpha = ak[k + 1];   // stays in place: the reflector's unit entry is implicit
      double beta = alpha, tk = 0.0, scal = 0.0;
      if (xn2 > 0.0) {
        beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
        tk = (beta - alpha) * fast_rcp(beta);
        scal = fast_rcp(alpha - beta);
      }
      const int m = n - k - 1;
      const int c2 = (((m + 1) >> 1) + 31) & ~31;
      if (c2 != cols2) {
        cols2 = c2;
        G = THREADS / cols2;
        if (G > TRI_GMAX) G = TRI_GMAX;
        jc = tid % cols2;
        q = tid / cols2;
      }
      const int nredw = cols2 >> 5;
      const int jA = k + 1 + jc, jB = jA + cols2;
      const bool actA = (q < G) && (jA < n), actB = (q < G) && (jB < n);
      for (int j = k + 1 + tid; j < n; j += THREADS) {
        if (j == k + 1) {
          v[j] = 1.0;
        } else {
          const double vj = ak[j] * scal;
          v[j] = vj;
          ak[j] = vj;  // reflector kept in row k (columns > k+1)
        }
      }
      if (tid == 0) {
        d[k] = ak[k];
        e[k] = beta;
        tau[k] = tk;
      }
      __syncthreads();  // (A) v complete
      SEG(0);
      double nx = 0.0;
      if (tk != 0.0) {
        const int rstep = G * lds;
        // ---- partial sums of A22 v over this thread's rows (4 rows x 2 columns in flight) ----
        if (actA && !(a.ablate & 1)) {
          const double* pa = S + (k + 1 + q) * lds + jA;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
          int i = k + 1 + q;
          for (; i + 3 * G < n; i += 4 * G, pa += 4 * rstep) {
            const double v0 = v[i], v1 = v[i + G], v2 = v[i + 2 * G], v3 = v[i + 3 * G];
            const double sa0 = pa[0], sa1 = pa[rstep], sa2 = pa[2 * rstep], sa3 = pa[3 * rstep];
            a0 = fma(sa0, v0, a0);
            a1 = fma(sa1, v1, a1);
            a2 = fma(sa2, v2, a2);
            a3 = fma(sa3, v3, a3);
            if (actB) {
              const double sb0 = pa[cols2], sb1 = pa[rstep + cols2], sb2 = pa[2 * rstep + cols2],
                           sb3 = pa[3 * rstep + cols2];
              b0 = fma(sb0, v0, b0);
              b1 = fma(sb1, v1, b1);
              b2 = fma(sb2, v2, b2);
              b3 = fma(sb3, v3, b3);
            }
          }
          for (; i < n; i += G, pa += rstep) {
            const double v0 = v[i];
            a0 = fma(pa[0], v0, a0);
            if (actB) b0 = fma(pa[cols2], v0, b0);
          }
          part[q * np + jA] = (a0 + a1) + (a2 + a3);
          if (actB) part[q * np + jB] = (b0 + b1) + (b2 + b3);
        }
        __syncthreads();  // (B) partial sums complete
        SEG(1);
        // ---- q == 0 threads: p_j, published; reduction (C): p.v and (fused) v.gq ----
        if (wid < nredw) {
          double r0 = 0.0, r1 = 0.0;
          if (actA) {
            double pj = 0.0, pk = 0.0;
            int qq = 0;
            for (; qq + 1 < G; qq += 2) {
              pj += part[qq * np + jA];
              pk += part[(qq + 1) * np + jA];
            }
            if (qq < G) pj += part[qq * np + jA];
            pj = (pj + pk) * tk;
            P[jA] = pj;
            const double vj = v[jA];
            r0 = pj * vj;
            if (a.fused) r1 = vj * gq[jA];
          }
          if (actB) {
            double pj = 0.0, pk = 0.0;
            int qq = 0;
            for (; qq + 1 < G; qq += 2) {
              pj += part[qq * np + jB];
              pk += part[(qq + 1) * np + jB];
            }
            if (qq < G) pj += part[qq * np + jB];
            pj = (pj + pk) * tk;
            P[jB] = pj;
            const double vj = v[jB];
            r0 = fma(pj, vj, r0);
            if (a.fused) r1 = fma(vj, gq[jB], r1);
          }
          r0 = warp_sum(r0);
          r1 = warp_sum(r1);
          if (lane == 0) {
            rb[par * 8 + wid] = r0;
            rb[par * 8 + 4 + wid] = r1;
          }
        }
        __syncthreads();  // (C) p and the partial dot products are visible
        SEG(2);
        double pv = 0.0, gd = 0.0;
        for (int ww = 0; ww < nredw; ++ww) {
          pv += rb[par * 8 + ww];
          gd += rb[par * 8 + 4 + ww];
        }
        par ^= 1;
        const double alpha2 = -0.5 * tk * pv;
        // ---- A22 -= v w^T + w v^T, w = p + alpha2 v ; next column norm on the fly ----
        if (actA) {
          const double vA = v[jA], wA = fma(alpha2, vA, P[jA]);
          double vB = 0.0, wB = 0.0;
          if (actB) {
            vB = v[jB];
            wB = fma(alpha2, vB, P[jB]);
          }
          if (q == 0 && a.fused) {  // gq <- H_k gq
            gq[jA] = fma(-tk * gd, vA, gq[jA]);
            if (actB) gq[jB] = fma(-tk * gd, vB, gq[jB]);
          }
          double* pa = S + (k + 1 + q) * lds + jA;
          int i = k + 1 + q;
          if (q == 0) {  // first row of A22: its tail is the next Householder column
            const double v0 = v[i], w0 = fma(alpha2, v0, P[i]);
            const double na = pa[0] - fma(v0, wA, w0 * vA);
            pa[0] = na;
            if (jA >= k + 3) nx = na * na;
            if (actB) {
              const double nb = pa[cols2] - fma(v0, wB, w0 * vB);
              pa[cols2] = nb;
              nx = fma(nb, nb, nx);  // jB >= k + 33
            }
            i += G;
            pa += rstep;
          }
          if (a.ablate & 2) i = n;
          for (; i + 3 * G < n; i += 4 * G, pa += 4 * rstep) {
            const double v0 = v[i], v1 = v[i + G], v2 = v[i + 2 * G], v3 = v[i + 3 * G];
            const double w0 = fma(alpha2, v0, P[i]), w1 = fma(alpha2, v1, P[i + G]),
                         w2 = fma(alpha2, v2, P[i + 2 * G]), w3 = fma(alpha2, v3, P[i + 3 * G]);
            const double sa0 = pa[0], sa1 = pa[rstep], sa2 = pa[2 * rstep], sa3 = pa[3 * rstep];
            if (actB) {
              const double sb0 = pa[cols2], sb1 = pa[rstep + cols2], sb2 = pa[2 * rstep + cols2],
                           sb3 = pa[3 * rstep + cols2];
              pa[cols2] = sb0 - fma(v0, wB, w0 * vB);
              pa[rstep + cols2] = sb1 - fma(v1, wB, w1 * vB);
              pa[2 * rstep + cols2] = sb2 - fma(v2, wB, w2 * vB);
              pa[3 * rstep + cols2] = sb3 - fma(v3, wB, w3 * vB);
            }
            pa[0] = sa0 - fma(v0, wA, w0 * vA);
            pa[rstep] = sa1 - fma(v1, wA, w1 * vA);
            pa[2 * rstep] = sa2 - fma(v2, wA, w2 * vA);
            pa[3 * rstep] = sa3 - fma(v3, wA, w3 * vA);
          }
          for (; i < n; i += G, pa += rstep) {
            const double v0 = v[i], w0 = fma(alpha2, v0, P[i]);
            if (actB) pa[cols2] = pa[cols2] - fma(v0, wB, w0 * vB);
            pa[0] = pa[0] - fma(v0, wA, w0 * vA);
          }
        }
      } else {
        if (q == 0) {
          if (actA && jA >= k + 3) nx = S[(k + 1) * lds + jA] * S[(k + 1) * lds + jA];
          if (actB) nx = fma(S[(k + 1) * lds + jB], S[(k + 1) * lds + jB], nx);
        }
      }
      // ---- reduction (E): next column norm; also publishes the updated A22 ----
      if (wid < nredw) {
        nx = warp_sum(nx);
        if (lane == 0) rb[par * 8 + wid] = nx;
      }
      SEG(3);
      __syncthreads();  // (E)
      SEG(4);
      xn2 = 0.0;
      for (int ww = 0; ww < nredw; ++ww) xn2 += rb[par * 8 + ww];
      par ^= 1;
    }
    if (a.dbg && (tid == 0 || tid == THREADS - 32))
      for (int i = 0; i < 5; ++i) a.dbg[(size_t)b * 16 + 8 + (tid == 0 ? 0 : 5) + i] = seg[i];
#undef SEG
  }
  if (!trivial && !prefactored) {
    if (tid == 0) {
      if (n >= 2) {
        d[n - 2] = S[(n - 2) * lds + (n - 2)];
        e[n - 2] = S[(n - 2) * lds + (n - 1)];
        tau[n - 2] = 0.0;
      }
      d[n - 1] = S[(n - 1) * lds + (n - 1)];
      e[n - 1] = 0.0;
      tau[n - 1] = 0.0;
    }
    __syncthreads();
  }
  TRI_MARK();  // 1: tridiagonalisation (+ Q^T gp)

  // spill the reflector rows (needed after S is recycled)
  if (!trivial && !prefactored)
    for (int idx = tid; idx < n * n; idx += THREADS) {
      const int i = idx / n, j = idx - i * n;
      if (j > i) Vh[idx] = S[i * lds + j];
    }

  // ---- scale, split ------------------------------------------------------------------------
  double tn = 0.0;
  if (!trivial)
    for (int i = tid; i < n; i += THREADS) tn = fmax(tn, fmax(fabs(d[i]), fabs(e[i])));
  tn = block_max(tn, s_red);
  if (tid == 0) s_tnorm = tn;
  const bool zero_t = trivial || tn == 0.0 || !isfinite(tn);
  if (!trivial && !isfinite(tn)) identity = true;
  __syncthreads();
  if (!zero_t) {
    const double inv_tn = 1.0 / tn;
    for (int i = tid; i < n; i += THREADS) {
      d[i] *= inv_tn;
      e[i] *= inv_tn;
    }
    __syncthreads();
    for (int i = tid; i < n - 1; i += THREADS)
      if (fabs(e[i]) <= TRI_EPS * (fabs(d[i]) + fabs(d[i + 1]))) e[i] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) e2[i] = e[i] * e[i];
    if (tid == 0) {
      int s0 = 0;
      for (int i = 0; i < n; ++i) {
        blk_s[i] = s0;
        if (i == n - 1 || e[i] == 0.0) {
          for (int r = s0; r <= i; ++r) blk_e[r] = i + 1;
          s0 = i + 1;
        }
      }
    }
    __syncthreads();
    TRI_MARK();  // 2: spill, scale, split

    // ---- phase 2: multisection on the Sturm count --------------------------------------------
    double* lo = X;
    double* hi = X + np;
    int* cnts = (int*)(X + 2 * np);  // [P][np]
    const int P = (3 * n <= THREADS) ? 3 : ((2 * n <= THREADS) ? 2 : 1);
    for (int i = tid; i < n; i += THREADS) {
      const int s = blk_s[i], t = blk_e[i];
      double gl = INFINITY, gu = -INFINITY;
      for (int r = s; r < t; ++r) {
        const double rad = (r > s ? fabs(e[r - 1]) : 0.0) + (r < t - 1 ? fabs(e[r]) : 0.0);
        gl = fmin(gl, d[r] - rad);
        gu = fmax(gu, d[r] + rad);
      }
      const double pad = 4.0 * TRI_EPS * n * fmax(fabs(gl), fabs(gu)) + 1e-300;
      lo[i] = gl - pad;
      hi[i] = gu + pad;
    }
    __syncthreads();
    const int i_own = tid % n, jpt = tid / n;   // valid when tid < P * n
    const bool worker = tid < P * n;
    for (int round = 0; round < 80; ++round) {
      int active = 0;
      if (worker) {
        const double l = lo[i_own], h = hi[i_own];
        const double width = h - l;
        if (width > 2.0 * TRI_EPS * fmax(fabs(l), fabs(h)) + 4e-3 * TRI_EPS) {  // abs floor ~1e-18 ||T||
          const double x = l + width * ((double)(jpt + 1) / (double)(P + 1));
          if (x > l && x < h) {
            active = 1;
            cnts[jpt * np + i_own] = sturm_count(d, e2, blk_s[i_own], blk_e[i_own], x);
          }
        }
        if (!active) cnts[jpt * np + i_own] = -1;
      }
      if (!__syncthreads_or(active)) break;
      if (worker && jpt == 0) {
        const int i = i_own;
        const int want = i - blk_s[i] + 1;
        const double l = lo[i], h = hi[i];
        const double width = h - l;
        double nl = l, nh = h;
        for (int jj = 0; jj < P; ++jj) {
          const int c = cnts[jj * np + i];
          if (c < 0) continue;
          const double xj = l + width * ((double)(jj + 1) / (double)(P + 1));
          if (c >= want) {
            nh = fmin(nh, xj);
            break;
          }
          nl = fmax(nl, xj);
        }
        lo[i] = nl;
        hi[i] = nh;
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += THREADS) lam[i] = 0.5 * (lo[i] + hi[i]);
    __syncthreads();
    TRI_MARK();  // 3: multisection

    // ---- phase 3: twisted-factorisation eigenvectors; column i of S <- vector i --------------
    // Reciprocal pivots: forward 1/D+ in S (shared), backward 1/D- in Dm (global, L2).
    // Thread i runs the forward sweep, thread n + i the backward sweep (same thread when
    // the CTA has fewer than 2n threads).
    {
      const double piv = TRI_EPS * 1e-3;
      const bool two = (2 * n <= THREADS);
      double* nrm_up = X;          // np
      double* nrm_dn = X + np;     // np
      int* twist = (int*)(X + 2 * np);
      for (int tt = tid; tt < (two ? 2 * n : n); tt += THREADS) {
        const int i = tt < n ? tt : tt - n;
        const int s = blk_s[i], t = blk_e[i];
        const double l = lam[i];
        if (tt < n) {  // forward: q_k = (d_k - l) - e2_{k-1} / q_{k-1}
          for (int k = 0; k < n; ++k)
            if (k < s || k >= t) S[k * lds + i] = 0.0;
          double q = d[s] - l;
          for (int k = s; k < t; ++k) {
            if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
            const double r = fast_rcp(q);
            S[k * lds + i] = r;
            if (k + 1 < t) q = fma(-e2[k], r, d[k + 1] - l);
          }
        }
        if (tt >= n || !two) {  // backward: q_k = (d_k - l) - e2_k / q_{k+1}
          double q = d[t - 1] - l;
          for (int k = t - 1; k >= s; --k) {
            if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
            const double r = fast_rcp(q);
            Dm[(size_t)k * n + i] = r;
            if (k > s) q = fma(-e2[k - 1], r, d[k - 1] - l);
          }
        }
      }
      __syncthreads();
      // twist index: gamma_k = D+_k - e2_k / D-_{k+1}   (Dm loads batched 8 deep: L2 latency)
      for (int i = tid; i < n; i += THREADS) {
        const int s = blk_s[i], t = blk_e[i];
        const double l = lam[i];
        double best = INFINITY;
        int r = s;
        for (int k0 = s; k0 < t; k0 += 8) {
          double rm[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) rm[u] = (k0 + u + 1 < t) ? Dm[(size_t)(k0 + u + 1) * n + i] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + u;
            if (k < t) {
              double dp = d[k] - l;
              if (k > s) dp = fma(-e2[k - 1], S[(k - 1) * lds + i], dp);
              const double gam = fabs(fma(-e2[k], rm[u], dp));  // e2[t-1] == 0 closes the block
              if (gam < best) {
                best = gam;
                r = k;
              }
            }
          }
        }
        twist[i] = r;
      }
      __syncthreads();
      // z_r = 1; z_k = -e_k z_{k+1} / D+_k (k < r); z_k = -e_{k-1} z_{k-1} / D-_k (k > r)
      for (int tt = tid; tt < (two ? 2 * n : n); tt += THREADS) {
        const int i = tt < n ? tt : tt - n;
        const int s = blk_s[i], t = blk_e[i];
        const int r = twist[i];
        if (tt < n) {
          double z = 1.0, acc = 0.0;
          for (int k = r - 1; k >= s; --k) {
            z = -(e[k] * S[k * lds + i]) * z;
            S[k * lds + i] = z;
            acc = fma(z, z, acc);
          }
          nrm_up[i] = acc;
        }
        if (tt >= n || !two) {
          double z = 1.0, acc = 0.0;
          for (int k0 = r + 1; k0 < t; k0 += 8) {
            double f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = (k0 + u < t) ? -(e[k0 + u - 1] * Dm[(size_t)(k0 + u) * n + i]) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (k0 + u < t) {
                z *= f[u];
                S[(k0 + u) * lds + i] = z;
                acc = fma(z, z, acc);
              }
            }
          }
          nrm_dn[i] = acc;
        }
      }
      __syncthreads();
      for (int i = tid; i < n; i += THREADS) {
        const int s = blk_s[i], t = blk_e[i];
        S[twist[i] * lds + i] = 1.0;
        const double sc = 1.0 / sqrt(1.0 + nrm_up[i] + nrm_dn[i]);
        for (int k = s; k < t; ++k) S[k * lds + i] *= sc;
        if (!isfinite(sc) || sc == 0.0) s_fallback = 1;
      }
    }
    // clusters: consecutive rows of one block whose eigenvalues are closer than GAPTOL.
    // Fused mode: modes that the RFO step filters anyway (|lambda| < 1e-7 absolute, rsirfo.py:30
    // drops < 1e-6) are left alone and break clusters (the TR/ROT null space lands here).
    if (tid == 0) {
      int cs = 0;
      const double dead = a.fused ? 1e-7 / s_tnorm : -1.0;
      for (int i = 0; i < n; ++i) {
        const bool chain = i > 0 && blk_s[i] == blk_s[i - 1] && (lam[i] - lam[i - 1]) < TRI_GAPTOL &&
                           !(fabs(lam[i]) < dead) && !(fabs(lam[i - 1]) < dead);
        if (!chain) cs = i;
        cl_s[i] = cs;
      }
    }
    __syncthreads();
    TRI_MARK();  // 4: twisted vectors

    // ---- phase 4: CGS2 inside clusters, one warp per cluster -------------------------------------
    double* dots = X;  // NW * 64
    for (int c0 = wid; c0 < n; c0 += NW) {
      if (cl_s[c0] != c0) continue;  // this warp owns clusters starting at c0
      int cend = c0 + 1;
      while (cend < n && cl_s[cend] == c0) ++cend;
      if (cend - c0 < 2) continue;
      const int s = blk_s[c0], t = blk_e[c0];
      double* dw = dots + wid * 64;
      for (int c = c0 + 1; c < cend; ++c) {
        double nfirst = 1.0;
        for (int rep = 0; rep < 2; ++rep) {
          for (int p0 = c0; p0 < c; p0 += 64) {
            const int pe = min(c, p0 + 64);
            for (int p = p0; p < pe; ++p) {
              double dt = 0.0;
              for (int k = s + lane; k < t; k += 32) dt = fma(S[k * lds + p], S[k * lds + c], dt);
              dt = warp_sum(dt);
              if (lane == 0) dw[p - p0] = dt;
            }
            __syncwarp();
            for (int k = s + lane; k < t; k += 32) {
              double zc = S[k * lds + c];
              for (int p = p0; p < pe; ++p) zc = fma(-dw[p - p0], S[k * lds + p], zc);
              S[k * lds + c] = zc;
            }
            __syncwarp();
          }
          double nn = 0.0;
          for (int k = s + lane; k < t; k += 32) nn = fma(S[k * lds + c], S[k * lds + c], nn);
          nn = sqrt(warp_sum(nn));
          if (rep == 0) nfirst = nn;
          if (!(nn > 1e-2)) {
            if (lane == 0) s_fallback = 1;  // vector (nearly) inside the span of its cluster
          }
          const double sc = nn > 0.0 ? 1.0 / nn : 0.0;
          for (int k = s + lane; k < t; k += 32) S[k * lds + c] *= sc;
          __syncwarp();
          if (rep == 0 && nfirst > 0.7) break;  // "twice is enough" only when needed
        }
      }
    }
    __syncthreads();
  } else {
    // zero (or non-finite) matrix: spectrum 0 (identity vectors)
    for (int i = tid; i < n; i += THREADS) {
      lam[i] = 0.0;
      blk_s[i] = i;
      blk_e[i] = i + 1;
      for (int k = 0; k < n; ++k) S[k * lds + i] = (k == i) ? 1.0 : 0.0;
    }
    __syncthreads();
    TRI_MARK();
    TRI_MARK();
    TRI_MARK();
  }

  TRI_MARK();  // 5: cluster re-orthogonalisation
  if (s_fallback) {  // robust path will redo this structure; leave state untouched
    if (tid == 0 && a.status) a.status[b] = st_in | MOP_ST_EIG_FALLBACK;
    return;
  }

  // ---- ascending order over all blocks -------------------------------------------------------------
  for (int i = tid; i < n; i += THREADS) {
    const double li = lam[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (lam[j] < li) || (lam[j] == li && j < i);
    rank[i] = r;
    inv[r] = i;
  }
  __syncthreads();
  const double tnorm = zero_t ? 0.0 : s_tnorm;
  double* evals = a.evals ? a.evals + (size_t)b * n : nullptr;

  if (!a.fused) {
    // ---- 5a: V = Q Z, thread i transforms column i (reflectors from L2) ----------------------------
    if (!trivial) {
      for (int i = tid; i < n; i += THREADS) {
        for (int k = n - 3; k >= 0; --k) {
          const double tk = tau[k];
          if (tk == 0.0) continue;
          const double* vk = Vh + (size_t)k * n;
          double dot = S[(k + 1) * lds + i];
          for (int j = k + 2; j < n; ++j) dot = fma(vk[j], S[j * lds + i], dot);
          dot *= tk;
          S[(k + 1) * lds + i] -= dot;
          for (int j = k + 2; j < n; ++j) S[j * lds + i] = fma(-dot, vk[j], S[j * lds + i]);
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += THREADS) evals[rank[i]] = lam[i] * tnorm;
    double* evecs = a.evecs + (size_t)b * n * n;
    for (int idx = tid; idx < n * n; idx += THREADS) {
      const int r = idx / n, k = idx - r * n;
      evecs[idx] = S[k * lds + inv[r]];
    }
    TRI_MARK();  // 6: back-transform + output
    if (tid == 0 && a.status) a.status[b] = st_in;
    return;
  }

  // ---- 5b: fused RS-I-RFO step in the eigenbasis ------------------------------------------------------
  double* lam_s = X;            // ascending spectrum (unscaled)
  double* gam_s = X + np;       // gamma in the same order
  RfoArrays R = rfo_carve(X + 2 * np, n);
  double pg = 0.0;
  for (int i = tid; i < n; i += THREADS) {
    const double g = a.Bg[(size_t)b * n + i];
    pg = fma(g, g, pg);
  }
  const double gnorm_raw = sqrt(block_sum(pg, s_red));
  for (int i = tid; i < n; i += THREADS) {
    double acc = 0.0;
    const int s = blk_s[i], t = blk_e[i];
    for (int k = s; k < t; ++k) acc = fma(S[k * lds + i], gq[k], acc);
    const int r = rank[i];
    lam_s[r] = identity ? 1.0 : lam[i] * tnorm;
    gam_s[r] = acc;
  }
  __syncthreads();
  if (evals)
    for (int i = tid; i < n; i += THREADS) evals[i] = lam_s[i];
  double* stp = a.state + (size_t)b * MOP_RSIRFO_STATE;
  int flags = identity ? MOP_ST_EIG_NONFINITE : 0;
  flags |= rfo_core(n, a.saddle_order, a.neb_mode, a.tmin, a.tmax, lam_s, gam_s, identity, gnorm_raw,
                    a.Be ? a.Be[b] : 0.0, stp, R, a.pred ? a.pred + b : nullptr);
  TRI_MARK();  // 6: eigenbasis RFO
  // y = Z c  (thread k)
  double* y = gq;
  for (int k = tid; k < n; k += THREADS) {
    double acc = 0.0;
    for (int r = 0; r < n; ++r) {
      const double c = R.coef[r];
      if (c != 0.0) acc = fma(S[k * lds + inv[r]], c, acc);
    }
    y[k] = acc;
  }
  __syncthreads();
  // step = Q y = H_0 ... H_{n-3} y.  Z is dead now: the reflector rows come back from L2
  // into S (coalesced), then ONE warp applies four reflectors per reduction round:
  // with u_i = v_i^T y and G_ij = v_i^T v_j all taken from the SAME y, the sequential
  // coefficients are c_3 = t_3 u_3, c_2 = t_2 (u_2 - c_3 G_23), ... (applied high k first).
  if (!trivial) {
    for (int idx = tid; idx < n * n; idx += THREADS) {
      const int i = idx / n, j = idx - i * n;
      if (j > i) S[i * lds + j] = Vh[idx];
    }
  }
  __syncthreads();
  if (wid == 0) {
    if (!trivial) {
      constexpr int MAXJ = (TRI_MAX_N + 31) / 32;
      int k = n - 3;
      for (; k >= 3; k -= 4) {
        // reflectors k (index 3) .. k-3 (index 0); v_i lives on rows > k-3+i
        const double* v0 = S + (k - 3) * lds;
        const double* v1 = S + (k - 2) * lds;
        const double* v2 = S + (k - 1) * lds;
        const double* v3 = S + k * lds;
        double r[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        double A0[MAXJ], A1[MAXJ], A2[MAXJ], A3[MAXJ], Y[MAXJ];
#pragma unroll
        for (int u = 0; u < MAXJ; ++u) {
          const int j = k - 2 + lane + 32 * u;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, yj = 0.0;
          if (j < n) {
            yj = y[j];
            a0 = (j == k - 2) ? 1.0 : v0[j];
            a1 = (j <= k - 2) ? 0.0 : ((j == k - 1) ? 1.0 : v1[j]);
            a2 = (j <= k - 1) ? 0.0 : ((j == k) ? 1.0 : v2[j]);
            a3 = (j <= k) ? 0.0 : ((j == k + 1) ? 1.0 : v3[j]);
          }
          A0[u] = a0; A1[u] = a1; A2[u] = a2; A3[u] = a3; Y[u] = yj;
          r[0] = fma(a0, yj, r[0]);
          r[1] = fma(a1, yj, r[1]);
          r[2] = fma(a2, yj, r[2]);
          r[3] = fma(a3, yj, r[3]);
          r[4] = fma(a0, a1, r[4]);  // G01
          r[5] = fma(a0, a2, r[5]);  // G02
          r[6] = fma(a0, a3, r[6]);  // G03
          r[7] = fma(a1, a2, r[7]);  // G12
          r[8] = fma(a1, a3, r[8]);  // G13
          r[9] = fma(a2, a3, r[9]);  // G23
        }
#pragma unroll
        for (int q = 0; q < 10; ++q) r[q] = warp_sum(r[q]);
        const double c3 = tau[k] * r[3];
        const double c2 = tau[k - 1] * (r[2] - c3 * r[9]);
        const double c1 = tau[k - 2] * (r[1] - c3 * r[8] - c2 * r[7]);
        const double c0 = tau[k - 3] * (r[0] - c3 * r[6] - c2 * r[5] - c1 * r[4]);
#pragma unroll
        for (int u = 0; u < MAXJ; ++u) {
          const int j = k - 2 + lane + 32 * u;
          if (j < n) y[j] = Y[u] - (c0 * A0[u] + c1 * A1[u] + c2 * A2[u] + c3 * A3[u]);
        }
        __syncwarp();
      }
      for (; k >= 0; --k) {
        const double tk = tau[k];
        if (tk == 0.0) continue;
        const double* vk = S + k * lds;
        double dot = 0.0;
        for (int j = k + 1 + lane; j < n; j += 32) dot = fma(j == k + 1 ? 1.0 : vk[j], y[j], dot);
        dot = warp_sum(dot) * tk;
        for (int j = k + 1 + lane; j < n; j += 32) y[j] = fma(-dot, j == k + 1 ? 1.0 : vk[j], y[j]);
        __syncwarp();
      }
    }
    for (int j = lane; j < n; j += 32) a.move[(size_t)b * n + j] = -y[j];
  }
  TRI_MARK();  // 7: back-transform
  if (tid == 0 && a.status) {
    const int keep = st_in & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV |
                              MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF);
    a.status[b] = keep | flags;
  }
#undef TRI_MARK
}

}  // namespace mop

// ---------------------------------------------------------------------------------------------
static long long* g_tri_dbg = nullptr;
// diagnostics: device buffer [B][8] receiving per-phase clock counts of the next launches
extern "C" int mop_debug_tri_timing(void* buf) {
  g_tri_dbg = (long long*)buf;
  return MOP_OK;
}

static int g_tri_threads_override = 0;
static int g_tri_ablate = 0;
extern "C" int mop_debug_tri_ablate(int mask) {
  g_tri_ablate = mask;
  return MOP_OK;
}
// diagnostics / tuning: force the CTA size of the tridiagonal kernels (0 = automatic)
extern "C" int mop_debug_tri_threads(int threads) {
  g_tri_threads_override = threads;
  return MOP_OK;
}
static int tri_threads(int n) {
  if (g_tri_threads_override == 128 || g_tri_threads_override == 256 || g_tri_threads_override == 512 ||
      g_tri_threads_override == 1024)
    return g_tri_threads_override;
  if (3 * n <= 128) return 128;
  if (3 * n <= 256) return 256;
  return 512;
}
static size_t tri_smem_bytes(int n) {
  const int np = (n + 3) & ~3, lds = n | 1;
  const size_t x_doubles = mop::tri_x_doubles(n);
  return sizeof(double) * ((size_t)n * lds + 6 * (size_t)np + x_doubles) + sizeof(int) * 5 * (size_t)np;
}

int mop_tridiag_supported(int n) { return n >= 1 && n <= mop::TRI_MAX_N && tri_smem_bytes(n) <= 227 * 1024; }
// Vh | Dm | d, e, tau, gq | flag  (the last five feed the prefactored mode)
size_t mop_tridiag_workspace_bytes(int B, int n) {
  return 2 * sizeof(double) * (size_t)B * n * n + 4 * sizeof(double) * (size_t)B * n + sizeof(int) * (size_t)B + 64;
}
int mop_launch_tridiag_packed(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                              double* tau, double* gq, int* flag, cudaStream_t stream);
int mop_spectrum_step_supported(int n);
int mop_launch_spectrum_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                             const double* Vh, double* Z, double* Dm, const double* pd, const double* pe,
                             const double* ptau, const double* pgq, const int* pflag, const double* Bg,
                             const double* Be, double* state, double* move, double* evals_out, double* pred,
                             int32_t* status, cudaStream_t stream);
static int g_tri_spectrum = 1;
// tuning: 1 (default) = k_spectrum_step (Z in global memory, 7 CTAs per SM) after the packed
// tridiagonalisation, 0 = k_eigh_tridiag in prefactored mode (Z in shared memory, 1 CTA per SM)
extern "C" int mop_debug_tri_spectrum(int on) {
  g_tri_spectrum = on;
  return MOP_OK;
}
static int g_tri_packed = 1;
// tuning: 1 (default) = packed two-CTA-per-SM tridiagonalisation feeding the fused kernel, 0 = single kernel
extern "C" int mop_debug_tri_packed(int on) {
  g_tri_packed = on;
  return MOP_OK;
}

template <int T>
static int launch_tri(int B, const mop::TriArgs& a, size_t smem, cudaStream_t stream) {
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_eigh_tridiag<T>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_eigh_tridiag<T><<<B, T, smem, stream>>>(a);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

static int launch_tri_any(int B, const mop::TriArgs& a, cudaStream_t stream) {
  const size_t smem = tri_smem_bytes(a.n);
  switch (tri_threads(a.n)) {
    case 128: return launch_tri<128>(B, a, smem, stream);
    case 256: return launch_tri<256>(B, a, smem, stream);
    case 1024: return launch_tri<1024>(B, a, smem, stream);
    default: return launch_tri<512>(B, a, smem, stream);
  }
}

int mop_launch_eigh_tridiag(int B, int n, const double* A, double* evals, double* evecs,
                            int32_t* status, void* work, size_t work_bytes, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_supported(n)) {
    mop_set_error("tridiagonal eigensolver: n = %d not supported (max %d)", n, mop::TRI_MAX_N);
    return MOP_ERR_UNSUPPORTED;
  }
  if (!work || work_bytes < mop_tridiag_workspace_bytes(B, n)) {
    mop_set_error("tridiagonal eigensolver: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  mop::TriArgs a{};
  a.n = n;
  a.fused = 0;
  a.A = A;
  a.Vh = (double*)work;
  a.Dm = (double*)work + (size_t)B * n * n;
  a.evals = evals;
  a.evecs = evecs;
  a.status = status;
  a.dbg = g_tri_dbg;
  a.ablate = g_tri_ablate;
  return launch_tri_any(B, a, stream);
}

int mop_launch_rsirfo_fused(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                            const double* Hp, const double* gp, const double* Bg, const double* Be,
                            double* state, double* move, double* evals_out, double* pred,
                            int32_t* status, void* work, size_t work_bytes, double* zbuf, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_supported(n)) {
    mop_set_error("fused RS-I-RFO kernel: n = %d not supported (max %d)", n, mop::TRI_MAX_N);
    return MOP_ERR_UNSUPPORTED;
  }
  if (!work || work_bytes < mop_tridiag_workspace_bytes(B, n)) {
    mop_set_error("fused RS-I-RFO kernel: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  mop::TriArgs a{};
  a.n = n;
  a.fused = 1;
  a.A = Hp;
  a.Vh = (double*)work;
  a.Dm = (double*)work + (size_t)B * n * n;
  a.evals = evals_out;
  a.evecs = nullptr;
  a.status = status;
  a.gp = gp;
  a.Bg = Bg;
  a.Be = Be;
  a.state = state;
  a.move = move;
  a.pred = pred;
  a.saddle_order = saddle_order;
  a.neb_mode = neb_mode;
  a.tmin = tmin;
  a.tmax = tmax;
  a.dbg = g_tri_dbg;
  a.ablate = g_tri_ablate;
  if (g_tri_packed && n > 2) {
    double* pf = (double*)work + 2 * (size_t)B * n * n;
    double* pd = pf;
    double* pe = pf + (size_t)B * n;
    double* pt = pf + 2 * (size_t)B * n;
    double* pg = pf + 3 * (size_t)B * n;
    int* pflag = (int*)(pf + 4 * (size_t)B * n);
    int rc = mop_launch_tridiag_packed(B, n, Hp, gp, a.Vh, pd, pe, pt, pg, pflag, stream);
    if (rc != MOP_OK) return rc;
    // spectrum + step with Z in global memory, seven structures per SM (spectrum_step.cu); zbuf is a
    // [B][n][n] slab the caller does not need until this launch has finished
    if (zbuf && g_tri_spectrum && mop_spectrum_step_supported(n))
      return mop_launch_spectrum_step(B, n, saddle_order, neb_mode, tmin, tmax, a.Vh, zbuf, a.Dm, pd, pe, pt, pg,
                                      pflag, Bg, Be, state, move, evals_out, pred, status, stream);
    a.pf_d = pd;
    a.pf_e = pe;
    a.pf_tau = pt;
    a.pf_gq = pg;
    a.pf_flag = pflag;
  }
  return launch_tri_any(B, a, stream);
}
