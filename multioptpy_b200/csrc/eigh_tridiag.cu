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

namespace mop {

constexpr int TRI_MAX_N = 160;
constexpr int TRI_GMAX = 8;          // max row-groups in the column-split symv / update
constexpr double TRI_EPS = 2.220446049250313e-16;
constexpr double TRI_GAPTOL = 1e-3;  // cluster gap relative to ||T|| (LAPACK dstein ORTOL)

// size (in doubles) of the phase-aliased scratch region X
__host__ __device__ inline size_t tri_x_doubles(int n) {
  const size_t np = (size_t)((n + 3) & ~3);
  size_t x = (size_t)(TRI_GMAX + 2) * np + 48;                               // phase 1 / 2
  const size_t fused = 2 * np + rfo_core_smem_bytes(n) / sizeof(double) + 8;  // fused RFO arrays
  if (fused > x) x = fused;
  if (16 * 64 > x) x = 16 * 64;                                              // phase 4 dots
  return (x + 1) & ~(size_t)1;
}

struct TriArgs {
  int n;
  int fused;  // 0: eigh (evals/evecs out)   1: fused RS-I-RFO step
  const double* A;   // [B][n][n] input matrix (projected Hessian in fused mode)
  double* Vh;        // [B][n][n] scratch: reflector rows
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
};

// number of eigenvalues of the unreduced block rows [s, t) that are < x
// (sign changes of p_k = (d_k - x) p_{k-1} - e_{k-1}^2 p_{k-2}); d, e2 scaled to ||T|| <= 1.
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2,
                                           int s, int t, double x) {
  double pm1 = 1.0;
  double p = d[s] - x;
  if (p == 0.0) p = -1e-300;
  int cnt = p < 0.0;
  for (int k = s + 1; k < t; ++k) {
    double pn = fma(d[k] - x, p, -(e2[k - 1] * pm1));
    if (pn == 0.0) pn = (p < 0.0) ? 1e-300 * fabs(p) + 1e-320 : -(1e-300 * fabs(p) + 1e-320);
    cnt += ((pn < 0.0) != (p < 0.0));
    pm1 = p;
    p = pn;
    const double a = fabs(p);
    if (!(a < 1e100 && a > 1e-100)) {  // rescale both (also catches inf/nan -> stays nan)
      const double sc = 1.0 / a;
      p *= sc;
      pm1 *= sc;
    }
  }
  return cnt;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_eigh_tridiag(TriArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = THREADS / 32;
  const int np = (n + 3) & ~3;
  const int lds = n | 1;
  // ---- shared memory carve-up --------------------------------------------------
  double* S = sm;                       // n x lds : A, then pivots, then Z
  double* d = S + (size_t)n * lds;      // np
  double* e = d + np;                   // np   e[k] couples k, k+1
  double* e2 = e + np;                  // np
  double* tau = e2 + np;                // np
  double* lam = tau + np;               // np   eigenvalue of thread-row i (scaled)
  double* gq = lam + np;                // np   Q^T gp, later y = Z c
  double* X = gq + np;                  // phase scratch (aliased):
  //   phase 1: v[np], w[np], part[TRI_GMAX*np]
  //   phase 2: lo[np], hi[np], cnt[3*np] ints
  //   phase 4: dots (NW * 64 doubles)
  //   fused  : lam_s[np], gam_s[np], RfoArrays
  int* blk_s = (int*)(X + tri_x_doubles(n));  // np
  int* blk_e = blk_s + np;              // np
  int* cl_s = blk_e + np;               // np   cluster start of i
  int* rank = cl_s + np;                // np   ascending rank of thread-row i
  int* inv = rank + np;                 // np   inverse permutation
  __shared__ double s_red[40];
  __shared__ double s_tnorm;
  __shared__ int s_fallback;

  const double* Ain = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;
  int st_in = a.status ? a.status[b] : 0;
  st_in &= ~(MOP_ST_EIG_FALLBACK | MOP_ST_EIG_NOCONV);

  long long t_prev = clock64();
  int t_slot = 0;
#define TRI_MARK()                                                      \
  do {                                                                  \
    if (a.dbg && tid == 0) {                                            \
      const long long t_now = clock64();                                \
      a.dbg[(size_t)b * 8 + (t_slot++)] = t_now - t_prev;               \
      t_prev = t_now;                                                   \
    }                                                                   \
  } while (0)
  // ---- load (symmetrised) ---------------------------------------------------------
  double pn = 0.0;
  for (int idx = tid; idx < n * n; idx += THREADS) {
    const int i = idx / n, j = idx - i * n;
    const double v = 0.5 * (Ain[idx] + Ain[(size_t)j * n + i]);
    S[i * lds + j] = v;
    pn = fma(v, v, pn);
  }
  if (tid == 0) s_fallback = 0;
  const double fro = sqrt(block_sum(pn, s_red));
  const bool finite_in = isfinite(fro);
  bool identity = !finite_in;   // non-finite input: rsirfo.py:365-369 identity fallback
  const bool trivial = identity || fro == 0.0;

  double* v = X;
  double* w = X + np;
  double* part = X + 2 * np;
  TRI_MARK();  // 0: load

  // ---- phase 1: tridiagonalisation ---------------------------------------------------
  if (!trivial) {
    for (int k = 0; k < n - 2; ++k) {
      double* ak = S + k * lds;
      double ps = 0.0;
      for (int j = k + 2 + tid; j < n; j += THREADS) ps = fma(ak[j], ak[j], ps);
      const double xn2 = block_sum(ps, s_red);
      const double alpha = ak[k + 1];
      double beta = alpha, tk = 0.0, scal = 0.0;
      if (xn2 > 0.0) {
        beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
        tk = (beta - alpha) / beta;
        scal = 1.0 / (alpha - beta);
      }
      __syncthreads();  // everyone has read ak[k+1]
      for (int j = k + 1 + tid; j < n; j += THREADS) {
        const double vj = (j == k + 1) ? 1.0 : ak[j] * scal;
        v[j] = vj;
        ak[j] = vj;  // reflector kept in row k
      }
      if (tid == 0) {
        d[k] = ak[k];
        e[k] = beta;
        tau[k] = tk;
      }
      __syncthreads();
      if (tk != 0.0) {
        const int m = n - k - 1;
        const int cols = (m + 31) & ~31;
        int G = THREADS / cols;
        if (G > TRI_GMAX) G = TRI_GMAX;
        if (G < 1) G = 1;
        const int jc = tid % cols, q = tid / cols;
        const int j = k + 1 + jc;
        const bool act = (q < G) && (jc < m);
        // p = tau * A22 v   (column split: thread (j, q) sums rows i = k+1+q, +G, ...)
        if (act) {
          double acc = 0.0;
          for (int i = k + 1 + q; i < n; i += G) acc = fma(S[i * lds + j], v[i], acc);
          part[q * np + j] = acc;
        }
        __syncthreads();
        double pv = 0.0;
        for (int jj = k + 1 + tid; jj < n; jj += THREADS) {
          double s = 0.0;
          for (int qq = 0; qq < G; ++qq) s += part[qq * np + jj];
          s *= tk;
          w[jj] = s;
          pv = fma(s, v[jj], pv);
        }
        pv = block_sum(pv, s_red);
        const double alpha2 = -0.5 * tk * pv;
        for (int jj = k + 1 + tid; jj < n; jj += THREADS) w[jj] = fma(alpha2, v[jj], w[jj]);
        __syncthreads();
        // A22 -= v w^T + w v^T
        if (act) {
          const double vj = v[j], wj = w[j];
          for (int i = k + 1 + q; i < n; i += G) {
            double* pa = S + i * lds + j;
            *pa = *pa - fma(v[i], wj, w[i] * vj);
          }
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      if (n >= 2) {
        d[n - 2] = S[(n - 2) * lds + (n - 2)];
        e[n - 2] = S[(n - 2) * lds + (n - 1)];
        tau[n - 2] = 0.0;
      }
      d[n - 1] = S[(n - 1) * lds + (n - 1)];
      e[n - 1] = 0.0;
      tau[n - 1] = 0.0;
      if (n == 1) tau[0] = 0.0;
    }
    __syncthreads();
  }

  TRI_MARK();  // 1: tridiagonalisation
  // ---- fused: gq = Q^T gp by the last warp (overlaps with the spill below) ----------------
  if (a.fused) {
    for (int i = tid; i < n; i += THREADS) gq[i] = a.gp[(size_t)b * n + i];
    __syncthreads();
    if (!trivial && wid == NW - 1) {
      for (int k = 0; k < n - 2; ++k) {
        const double tk = tau[k];
        if (tk == 0.0) continue;
        const double* vk = S + k * lds;
        double dot = 0.0;
        for (int j = k + 1 + lane; j < n; j += 32) dot = fma(vk[j], gq[j], dot);
        dot = warp_sum(dot) * tk;
        for (int j = k + 1 + lane; j < n; j += 32) gq[j] = fma(-dot, vk[j], gq[j]);
        __syncwarp();
      }
    }
  }
  // spill the reflector rows (needed after S is recycled)
  if (!trivial)
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

    TRI_MARK();  // 2: Q^T g, spill, scale, split
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
      double x = 0.0;
      if (worker) {
        const double l = lo[i_own], h = hi[i_own];
        const double width = h - l;
        if (width > 2.0 * TRI_EPS * fmax(fabs(l), fabs(h)) + 4e-3 * TRI_EPS) {  // abs floor ~1e-18 ||T||
          x = l + width * ((double)(jpt + 1) / (double)(P + 1));
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
        for (int j = 0; j < P; ++j) {
          const int c = cnts[j * np + i];
          if (c < 0) continue;
          const double xj = l + width * ((double)(j + 1) / (double)(P + 1));
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
    // ---- phase 3: twisted-factorisation eigenvectors, thread i -> column i of S ----------------
    for (int i = tid; i < n; i += THREADS) {
      const int s = blk_s[i], t = blk_e[i];
      for (int k = 0; k < n; ++k)
        if (k < s || k >= t) S[k * lds + i] = 0.0;
      if (t - s == 1) {
        S[s * lds + i] = 1.0;
        continue;
      }
      const double l = lam[i];
      const double piv = TRI_EPS * 1e-3;
      // forward pivots D+ (stored), backward pivots D- (streamed) -> twist index r
      double q = d[s] - l;
      S[s * lds + i] = q;
      for (int k = s + 1; k < t; ++k) {
        if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
        q = (d[k] - l) - e2[k - 1] / q;
        S[k * lds + i] = q;
      }
      double dm = d[t - 1] - l;
      double best = fabs(S[(t - 1) * lds + i]);  // gamma_{t-1} = D+_{t-1}
      int r = t - 1;
      for (int k = t - 2; k >= s; --k) {
        if (fabs(dm) < piv) dm = (dm <= 0.0) ? -piv : piv;
        dm = (d[k] - l) - e2[k] / dm;
        const double gam = fabs(S[k * lds + i] + dm - (d[k] - l));
        if (gam < best) {
          best = gam;
          r = k;
        }
      }
      // recompute D- for k > r and store it over D+ (no longer needed there)
      if (r < t - 1) {
        dm = d[t - 1] - l;
        S[(t - 1) * lds + i] = dm;
        for (int k = t - 2; k > r; --k) {
          if (fabs(dm) < piv) dm = (dm <= 0.0) ? -piv : piv;
          dm = (d[k] - l) - e2[k] / dm;
          S[k * lds + i] = dm;
        }
      }
      // z_r = 1, outward recurrences, in place
      double z = 1.0, nrm2 = 1.0;
      for (int k = r - 1; k >= s; --k) {
        double dp = S[k * lds + i];
        if (fabs(dp) < piv) dp = (dp <= 0.0) ? -piv : piv;
        z = -(e[k] / dp) * z;
        S[k * lds + i] = z;
        nrm2 = fma(z, z, nrm2);
      }
      z = 1.0;
      for (int k = r + 1; k < t; ++k) {
        double dq = S[k * lds + i];
        if (fabs(dq) < piv) dq = (dq <= 0.0) ? -piv : piv;
        z = -(e[k - 1] / dq) * z;
        S[k * lds + i] = z;
        nrm2 = fma(z, z, nrm2);
      }
      S[r * lds + i] = 1.0;
      const double sc = 1.0 / sqrt(nrm2);
      for (int k = s; k < t; ++k) S[k * lds + i] *= sc;
      if (!isfinite(sc) || sc == 0.0) s_fallback = 1;
    }
    // clusters: consecutive rows of one block whose eigenvalues are closer than GAPTOL
    if (tid == 0) {
      int cs = 0;
      for (int i = 0; i < n; ++i) {
        if (i > 0 && blk_s[i] == blk_s[i - 1] && (lam[i] - lam[i - 1]) < TRI_GAPTOL) {
          // same cluster
        } else {
          cs = i;
        }
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
          if (!(nn > 1e-3) ) {
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
  // y = Z c  (thread k), then step = Q y by one warp, move = -step
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
  if (wid == 0) {
    if (!trivial) {
      for (int k = n - 3; k >= 0; --k) {
        const double tk = tau[k];
        if (tk == 0.0) continue;
        const double* vk = Vh + (size_t)k * n;
        double dot = 0.0;
        for (int j = k + 1 + lane; j < n; j += 32) dot = fma(j == k + 1 ? 1.0 : vk[j], y[j], dot);
        dot = warp_sum(dot) * tk;
        for (int j = k + 1 + lane; j < n; j += 32) y[j] = fma(-dot, j == k + 1 ? 1.0 : vk[j], y[j]);
        __syncwarp();
      }
    }
    for (int j = lane; j < n; j += 32) a.move[(size_t)b * n + j] = -y[j];
  }
  TRI_MARK();  // 6: eigenbasis RFO + back-transform
  if (tid == 0 && a.status) {
    const int keep = st_in & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV |
                              MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF);
    a.status[b] = keep | flags;
  }
}

}  // namespace mop

// ---------------------------------------------------------------------------------------------
static long long* g_tri_dbg = nullptr;
// diagnostics: device buffer [B][8] receiving per-phase clock counts of the next launches
extern "C" int mop_debug_tri_timing(void* buf) {
  g_tri_dbg = (long long*)buf;
  return MOP_OK;
}

static int tri_threads(int n) {
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
size_t mop_tridiag_workspace_bytes(int B, int n) { return sizeof(double) * (size_t)B * n * n; }

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
  a.evals = evals;
  a.evecs = evecs;
  a.status = status;
  a.dbg = g_tri_dbg;
  return launch_tri_any(B, a, stream);
}

int mop_launch_rsirfo_fused(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                            const double* Hp, const double* gp, const double* Bg, const double* Be,
                            double* state, double* move, double* evals_out, double* pred,
                            int32_t* status, void* work, size_t work_bytes, cudaStream_t stream) {
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
  return launch_tri_any(B, a, stream);
}
