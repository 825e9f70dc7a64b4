// Translation/rotation projection of Hessian and gradient.
//
// Reference: Calculationtools.project_out_hess_tr_and_rot_for_coord
// (Utils/calc_tools.py:249-316) forms the dense P = I - sum t t^T and two n^3
// matmuls.  Here the rank-6 structure is used:
//     Hp = S - Y T^T - T Y^T,   S = sym(H + Hbias),  W = S T,  Y = W - 1/2 T (T^T W)
// which is O(n^2) and streams the matrix twice (W pass + output pass).
// T (n x k, k <= 6) is the reference's CLASSICAL Gram-Schmidt basis (drop 1e-10).
// The gradient projection g - Q Q^T g (Optimizer/rsirfo.py:128-190, reduced QR)
// spans the same space when the six vectors are independent.  For rank-deficient
// sets (two atoms, linear molecules; MOP_ST_TRROT_RANKDEF) the reference's Q keeps
// SIX orthonormal columns - Householder QR never drops one, the column of the
// dependent vector is whatever the earlier reflectors make of a unit vector - so
// the kernel replays LAPACK's dgeqr2 + dorg2r there (project_grad_qr below).
#include "trrot.cuh"

namespace mop {

constexpr int PRJ_THREADS = 256;
constexpr int PT = 32;

__global__ void __launch_bounds__(PRJ_THREADS)
k_project_trrot(int n, int grad_rule, const double* __restrict__ Hall, const double* __restrict__ Hb_all,
                const double* __restrict__ x_all, const double* __restrict__ g_all,
                double* __restrict__ Hp_all, double* __restrict__ gp_all,
                int32_t* __restrict__ status, const int32_t* __restrict__ only_flagged) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = PRJ_THREADS >> 5;
  if (only_flagged && !(only_flagged[b] & MOP_ST_EIG_FALLBACK)) return;  // fallback re-projection of flagged structures
  const int np = (n + 3) & ~3;
  double* T = sm;                 // 6 x np
  double* raw = T + 6 * np;       // 6 x np   (later reused as W -> Y)
  double* scratch = raw + 6 * np; // 40
  double* M = scratch + 40;       // 36 : T^T W
  double* tA = M + 36;            // 32 x 33
  double* tB = tA + PT * (PT + 1);
  const double* H = Hall + (size_t)b * n * n;
  const double* Hb = Hb_all ? Hb_all + (size_t)b * n * n : nullptr;
  const double* x = x_all + (size_t)b * n;

  const int k = build_trrot_basis(n, x, T, np, raw, scratch);
  if (tid == 0 && status) {
    int st = status[b] & ~MOP_ST_TRROT_RANKDEF;
    if (k < 6) st |= MOP_ST_TRROT_RANKDEF;
    status[b] = st;
  }

  // ---- gradient: gp = g - T (T^T g) -----------------------------------------
  if (g_all && gp_all) {
    const double* g = g_all + (size_t)b * n;
    if (k < 6) {  // block-uniform: the reference's Householder Q differs from the Gram-Schmidt span here
      project_grad_qr(n, raw, np, g, gp_all + (size_t)b * n, grad_rule, scratch);
    } else {
      double cf[6];
      for (int j = 0; j < k; ++j) {
        double p = 0.0;
        for (int i = tid; i < n; i += PRJ_THREADS) p = fma(T[j * np + i], g[i], p);
        cf[j] = block_sum(p, scratch);
      }
      for (int i = tid; i < n; i += PRJ_THREADS) {
        double part = 0.0;
        for (int j = 0; j < k; ++j) part = fma(T[j * np + i], cf[j], part);
        gp_all[(size_t)b * n + i] = g[i] - part;
      }
    }
  }
  if (!Hp_all) return;
  __syncthreads();

  // ---- W = S T, S = 1/2 (M0 + M0^T): row pass (M0 T) + column pass (M0^T T) ----
  double* W = raw;  // raw vectors are dead now
  for (int i = tid; i < 6 * np; i += PRJ_THREADS) W[i] = 0.0;
  __syncthreads();
  for (int i = w; i < n; i += nw) {  // (M0 T)_i = sum_j M0[i][j] T[j]
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int j = lane; j < n; j += 32) {
      double a = H[(size_t)i * n + j];
      if (Hb) a += Hb[(size_t)i * n + j];
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[v] = fma(a, T[v * np + j], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < 6; ++v) acc[v] = warp_sum(acc[v]);
    if (lane == 0)
      for (int v = 0; v < k; ++v) W[v * np + i] = 0.5 * acc[v];
  }
  __syncthreads();
  for (int j = tid; j < n; j += PRJ_THREADS) {  // (M0^T T)_j = sum_i M0[i][j] T[i]
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      double a = H[(size_t)i * n + j];
      if (Hb) a += Hb[(size_t)i * n + j];
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[v] = fma(a, T[v * np + i], acc[v]);
    }
    for (int v = 0; v < k; ++v) W[v * np + j] += 0.5 * acc[v];
  }
  __syncthreads();
  // M = T^T W (k x k), then Y = W - 1/2 T M  (in place in W)
  for (int e = w; e < k * k; e += nw) {
    const int a = e / k, c = e - a * k;
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(T[a * np + i], W[c * np + i], p);
    p = warp_sum(p);
    if (lane == 0) M[a * 6 + c] = p;
  }
  __syncthreads();
  for (int i = tid; i < n; i += PRJ_THREADS) {
    double tv[6];
    for (int a = 0; a < k; ++a) tv[a] = T[a * np + i];
    for (int c = 0; c < k; ++c) {
      double corr = 0.0;
      for (int a = 0; a < k; ++a) corr = fma(tv[a], 0.5 * (M[a * 6 + c] + M[c * 6 + a]), corr);
      W[c * np + i] -= 0.5 * corr;
    }
  }
  __syncthreads();
  const double* Y = W;

  // ---- output pass: Hp = S - Y T^T - T Y^T, tile pairs ------------------------
  double* Hp = Hp_all + (size_t)b * n * n;
  const int TT = (n + PT - 1) / PT;
  for (int I = 0; I < TT; ++I) {
    for (int J = I; J < TT; ++J) {
      const int i0 = I * PT, j0 = J * PT;
      for (int e = tid; e < PT * PT; e += PRJ_THREADS) {
        const int r = e >> 5, c = e & 31;
        int gi = i0 + r, gj = j0 + c;
        double a = 0.0;
        if (gi < n && gj < n) {
          a = H[(size_t)gi * n + gj];
          if (Hb) a += Hb[(size_t)gi * n + gj];
        }
        tA[r * (PT + 1) + c] = a;
        gi = j0 + r;
        gj = i0 + c;
        a = 0.0;
        if (gi < n && gj < n) {
          a = H[(size_t)gi * n + gj];
          if (Hb) a += Hb[(size_t)gi * n + gj];
        }
        tB[r * (PT + 1) + c] = a;
      }
      __syncthreads();
      for (int e = tid; e < PT * PT; e += PRJ_THREADS) {
        const int r = e >> 5, c = e & 31;
        {
          const int gi = i0 + r, gj = j0 + c;
          if (gi < n && gj < n) {
            double v = 0.5 * (tA[r * (PT + 1) + c] + tB[c * (PT + 1) + r]);
            for (int a = 0; a < k; ++a)
              v -= __dadd_rn(__dmul_rn(Y[a * np + gi], T[a * np + gj]),
                             __dmul_rn(T[a * np + gi], Y[a * np + gj]));  // bit-symmetric
            Hp[(size_t)gi * n + gj] = v;
          }
        }
        if (J != I) {
          const int gi = j0 + r, gj = i0 + c;
          if (gi < n && gj < n) {
            double v = 0.5 * (tB[r * (PT + 1) + c] + tA[c * (PT + 1) + r]);
            for (int a = 0; a < k; ++a)
              v -= __dadd_rn(__dmul_rn(Y[a * np + gi], T[a * np + gj]),
                             __dmul_rn(T[a * np + gi], Y[a * np + gj]));  // bit-symmetric
            Hp[(size_t)gi * n + gj] = v;
          }
        }
      }
      __syncthreads();
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Multi-CTA path (inside the fused optimizer steps, scratch from the caller): same arithmetic, spread
// over row / tile blocks so that small batches fill the GPU and the three passes over one structure's
// Hessian are not serialised inside one CTA.
//   k_prj_basis  grid B            T, rank, projected gradient
//   k_prj_w      grid (n/32, B)    W = 1/2 (M0 T + M0^T T) for a block of 32 indices (row + column pass)
//   k_prj_y      grid B            M = T^T W, Y = W - 1/2 T sym(M)
//   k_prj_out    grid (tile rows, B)  Hp = S - Y T^T - T Y^T, tile pairs (I, J >= I), register prefetch
// Scratch per structure: T [6][np] | W -> Y [6][np] | header (rank) 8 doubles.
__host__ __device__ inline size_t prj_scratch_doubles(int n) { return 12 * (size_t)((n + 3) & ~3) + 8; }

__global__ void __launch_bounds__(PRJ_THREADS)
k_prj_basis(int n, int grad_rule, size_t sstride, const double* __restrict__ x_all, const double* __restrict__ g_all,
            double* __restrict__ gp_all, double* __restrict__ scratch, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int np = (n + 3) & ~3;
  double* T = sm;
  double* raw = T + 6 * np;
  double* red = raw + 6 * np;
  double* scr = scratch + (size_t)b * sstride;
  const int k = build_trrot_basis(n, x_all + (size_t)b * n, T, np, raw, red);
  if (tid == 0) {
    scr[12 * np] = (double)k;
    if (status) {
      int st = status[b] & ~MOP_ST_TRROT_RANKDEF;
      if (k < 6) st |= MOP_ST_TRROT_RANKDEF;
      status[b] = st;
    }
  }
  for (int i = tid; i < 6 * np; i += PRJ_THREADS) scr[i] = (i / np < k) ? T[i] : 0.0;
  if (g_all && gp_all) {
    const double* g = g_all + (size_t)b * n;
    if (k < 6) {
      __syncthreads();
      project_grad_qr(n, raw, np, g, gp_all + (size_t)b * n, grad_rule, red);
      return;
    }
    double cf[6];
    for (int j = 0; j < k; ++j) {
      double p = 0.0;
      for (int i = tid; i < n; i += PRJ_THREADS) p = fma(T[j * np + i], g[i], p);
      cf[j] = block_sum(p, red);
    }
    for (int i = tid; i < n; i += PRJ_THREADS) {
      double part = 0.0;
      for (int j = 0; j < k; ++j) part = fma(T[j * np + i], cf[j], part);
      gp_all[(size_t)b * n + i] = g[i] - part;
    }
  }
}

__global__ void __launch_bounds__(PRJ_THREADS)
k_prj_w(int n, size_t sstride, const double* __restrict__ Hall, const double* __restrict__ Hb_all,
        double* __restrict__ scratch) {
  extern __shared__ double sm[];  // T [6][np] | colacc [8][6][32]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int nw = PRJ_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* T = sm;
  double* cacc = T + 6 * np;
  double* scr = scratch + (size_t)b * sstride;
  double* W = scr + 6 * np;
  for (int i = tid; i < 6 * np; i += PRJ_THREADS) T[i] = scr[i];
  __syncthreads();
  const double* H = Hall + (size_t)b * n * n;
  const double* Hb = Hb_all ? Hb_all + (size_t)b * n * n : nullptr;
  const int i0 = blockIdx.x * 32;
  // column pass: lane owns column i0 + lane, warp w the rows j = w, w + 8, ...
  double ca[6] = {0, 0, 0, 0, 0, 0};
  const int col = i0 + lane;
  if (col < n) {
    for (int j = w; j < n; j += nw) {
      double a = H[(size_t)j * n + col];
      if (Hb) a += Hb[(size_t)j * n + col];
#pragma unroll
      for (int v = 0; v < 6; ++v) ca[v] = fma(a, T[v * np + j], ca[v]);
    }
  }
#pragma unroll
  for (int v = 0; v < 6; ++v) cacc[(w * 6 + v) * 32 + lane] = ca[v];
  // row pass: warp w owns rows i0 + 4 w .. + 3
  double ra[4][6];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int v = 0; v < 6; ++v) ra[q][v] = 0.0;
  const int r0 = i0 + 4 * w;
  for (int j = lane; j < n; j += 32) {
    double t[6];
#pragma unroll
    for (int v = 0; v < 6; ++v) t[v] = T[v * np + j];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + q;
      if (r < n) {
        double a = H[(size_t)r * n + j];
        if (Hb) a += Hb[(size_t)r * n + j];
#pragma unroll
        for (int v = 0; v < 6; ++v) ra[q][v] = fma(a, t[v], ra[q][v]);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = r0 + q;
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      const double rsum = warp_sum(ra[q][v]);
      if (lane == 0 && r < n) {
        double csum = 0.0;
        for (int ww = 0; ww < nw; ++ww) csum += cacc[(ww * 6 + v) * 32 + (r - i0)];
        W[v * np + r] = 0.5 * rsum + 0.5 * csum;
      }
    }
  }
}

// One-read W pass: CTA (slab, b) streams the 32 rows of its slab ONCE.  Element a = M0[r][j] adds a T[j] to
// the row sums of r (complete inside the CTA) and a T[r] to the column sums of j (partial: this slab only);
// the partials go to scratch [slab][6][np] and k_prj_y adds them in slab order (deterministic).
// shared: T [6][np] | cpart [8 warps][6][np]
__global__ void __launch_bounds__(PRJ_THREADS)
k_prj_w1(int n, size_t sstride, const double* __restrict__ Hall, const double* __restrict__ Hb_all,
         double* __restrict__ scratch) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int nw = PRJ_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* T = sm;
  double* cpart = T + 6 * np;
  double* scr = scratch + (size_t)b * sstride;
  double* W = scr + 6 * np;
  double* Wc = scr + 12 * np + 8 + (size_t)blockIdx.x * 6 * np;
  for (int i = tid; i < 6 * np; i += PRJ_THREADS) T[i] = scr[i];
  __syncthreads();
  const double* H = Hall + (size_t)b * n * n;
  const double* Hb = Hb_all ? Hb_all + (size_t)b * n * n : nullptr;
  const int i0 = blockIdx.x * 32, r0 = i0 + 4 * w;
  double tr[4][6], ra[4][6];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      tr[q][v] = (r0 + q < n) ? T[v * np + r0 + q] : 0.0;
      ra[q][v] = 0.0;
    }
  for (int j0 = 0; j0 < n; j0 += 64) {  // two column chunks per pass: eight loads in flight per lane
    double a[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = j0 + 32 * h + lane;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q;
        double x = 0.0;
        if (r < n && j < n) {
          x = H[(size_t)r * n + j];
          if (Hb) x += Hb[(size_t)r * n + j];
        }
        a[h][q] = x;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = j0 + 32 * h + lane;
      if (j < n) {
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          const double t = T[v * np + j];
          double c = 0.0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ra[q][v] = fma(a[h][q], t, ra[q][v]);
            c = fma(a[h][q], tr[q][v], c);
          }
          cpart[(size_t)(w * 6 + v) * np + j] = c;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      const double rsum = warp_sum(ra[q][v]);
      if (lane == 0 && r0 + q < n) W[v * np + r0 + q] = rsum;
    }
  __syncthreads();
  for (int e = tid; e < 6 * np; e += PRJ_THREADS) {
    const int v = e / np, j = e - v * np;
    double csum = 0.0;
    if (j < n)
      for (int ww = 0; ww < nw; ++ww) csum += cpart[(size_t)(ww * 6 + v) * np + j];
    Wc[e] = csum;
  }
}

__global__ void __launch_bounds__(PRJ_THREADS)
k_prj_y(int n, size_t sstride, int nslab, double* __restrict__ scratch) {
  __shared__ double M[36];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = PRJ_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* scr = scratch + (size_t)b * sstride;
  const double* T = scr;
  double* W = scr + 6 * np;
  const int k = (int)scr[12 * np];
  if (nslab > 0) {  // one-read W pass: W = 1/2 (row sums + column partials of every row slab, fixed order)
    const double* Wc = scr + 12 * np + 8;
    for (int e = tid; e < 6 * np; e += PRJ_THREADS) {
      double csum = 0.0;
      for (int sl = 0; sl < nslab; ++sl) csum += Wc[(size_t)sl * 6 * np + e];
      W[e] = 0.5 * W[e] + 0.5 * csum;
    }
    __syncthreads();
  }
  for (int e = w; e < k * k; e += nw) {
    const int a = e / k, c = e - a * k;
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(T[a * np + i], W[c * np + i], p);
    p = warp_sum(p);
    if (lane == 0) M[a * 6 + c] = p;
  }
  __syncthreads();
  for (int i = tid; i < n; i += PRJ_THREADS) {
    double tv[6];
    for (int a = 0; a < k; ++a) tv[a] = T[a * np + i];
    for (int c = 0; c < k; ++c) {
      double corr = 0.0;
      for (int a = 0; a < k; ++a) corr = fma(tv[a], 0.5 * (M[a * 6 + c] + M[c * 6 + a]), corr);
      W[c * np + i] -= 0.5 * corr;
    }
  }
}

__global__ void __launch_bounds__(PRJ_THREADS)
k_prj_out(int n, size_t sstride, int TT, const double* __restrict__ Hall, const double* __restrict__ Hb_all,
          const double* __restrict__ scratch, double* __restrict__ Hp_all) {
  __shared__ double tA[PT * (PT + 1)], tB[PT * (PT + 1)];
  __shared__ double ti[12][PT], tj[12][PT];  // rows 0-5: T, 6-11: Y, for the I and J index blocks
  const int b = blockIdx.y, tid = threadIdx.x, I = blockIdx.x;
  const int np = (n + 3) & ~3;
  const double* scr = scratch + (size_t)b * sstride;
  const int k = (int)scr[12 * np];
  const double* H = Hall + (size_t)b * n * n;
  const double* Hb = Hb_all ? Hb_all + (size_t)b * n * n : nullptr;
  double* Hp = Hp_all + (size_t)b * n * n;
  const int i0 = I * PT;
  for (int e = tid; e < 12 * PT; e += PRJ_THREADS) {
    const int a = e / PT, r = e - a * PT;
    ti[a][r] = (i0 + r < n) ? scr[(size_t)a * np + i0 + r] : 0.0;
  }
  constexpr int EPT = PT * PT / PRJ_THREADS;
  double ra[EPT], rb[EPT];
  auto load = [&](int J) {
    const int j0 = J * PT;
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * PRJ_THREADS, r = e >> 5, c = e & 31;
      int gi = i0 + r, gj = j0 + c;
      double a = 0.0;
      if (gi < n && gj < n) {
        a = H[(size_t)gi * n + gj];
        if (Hb) a += Hb[(size_t)gi * n + gj];
      }
      ra[u] = a;
      gi = j0 + r;
      gj = i0 + c;
      a = 0.0;
      if (gi < n && gj < n) {
        a = H[(size_t)gi * n + gj];
        if (Hb) a += Hb[(size_t)gi * n + gj];
      }
      rb[u] = a;
    }
  };
  load(I);
  for (int J = I; J < TT; ++J) {
    const int j0 = J * PT;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * PRJ_THREADS, r = e >> 5, c = e & 31;
      tA[r * (PT + 1) + c] = ra[u];
      tB[r * (PT + 1) + c] = rb[u];
    }
    for (int e = tid; e < 12 * PT; e += PRJ_THREADS) {
      const int a = e / PT, r = e - a * PT;
      tj[a][r] = (j0 + r < n) ? scr[(size_t)a * np + j0 + r] : 0.0;
    }
    __syncthreads();
    if (J + 1 < TT) load(J + 1);
    for (int e = tid; e < PT * PT; e += PRJ_THREADS) {
      const int r = e >> 5, c = e & 31;
      {
        const int gi = i0 + r, gj = j0 + c;
        if (gi < n && gj < n) {
          double v = 0.5 * (tA[r * (PT + 1) + c] + tB[c * (PT + 1) + r]);
          for (int a = 0; a < k; ++a)
            v -= __dadd_rn(__dmul_rn(ti[6 + a][r], tj[a][c]), __dmul_rn(ti[a][r], tj[6 + a][c]));  // bit-symmetric
          Hp[(size_t)gi * n + gj] = v;
        }
      }
      if (J != I) {
        const int gi = j0 + r, gj = i0 + c;
        if (gi < n && gj < n) {
          double v = 0.5 * (tB[r * (PT + 1) + c] + tA[c * (PT + 1) + r]);
          for (int a = 0; a < k; ++a)
            v -= __dadd_rn(__dmul_rn(tj[6 + a][r], ti[a][c]), __dmul_rn(tj[a][r], ti[6 + a][c]));  // bit-symmetric
          Hp[(size_t)gi * n + gj] = v;
        }
      }
    }
  }
}

}  // namespace mop

// grad_rule: 0 = RSIRFO's gradient projection, 1 = EnhancedRSPRFO's (they differ for rank-deficient TR/ROT sets only)
int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                             const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                             cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  const int np = (n + 3) & ~3;
  const size_t smem = sizeof(double) * (12 * (size_t)np + 40 + 36 + 2 * mop::PT * (mop::PT + 1));
  if (smem > 200 * 1024) {
    mop_set_error("n = %d too large for the projection kernel", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_project_trrot,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_project_trrot<<<B, mop::PRJ_THREADS, smem, stream>>>(n, grad_rule, H, Hbias, x, g, Hp_out, gp_out,
                                                            status, nullptr);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// The same projection for the structures whose status carries MOP_ST_EIG_FALLBACK only (the others exit at once):
// the fused front end never writes the projected Hessian, the robust eigensolver path needs it.
int mop_launch_project_trrot_flagged(int B, int n, const double* H, const double* Hbias, const double* x,
                                     double* Hp_out, const int32_t* flags, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  const int np = (n + 3) & ~3;
  const size_t smem = sizeof(double) * (12 * (size_t)np + 40 + 36 + 2 * mop::PT * (mop::PT + 1));
  if (smem > 200 * 1024) {
    mop_set_error("n = %d too large for the projection kernel", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_project_trrot,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_project_trrot<<<B, mop::PRJ_THREADS, smem, stream>>>(n, 0, H, Hbias, x, nullptr, Hp_out, nullptr, nullptr,
                                                            flags);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_project_trrot(int B, int n, const double* H, const double* Hbias,
                                 const double* x, const double* g, double* Hp_out, double* gp_out,
                                 int32_t* status, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && n % 3 == 0, "mop_project_trrot: n must be a positive multiple of 3");
  MOP_REQUIRE(x, "mop_project_trrot: x must be a device pointer");
  MOP_REQUIRE((H && Hp_out) || (g && gp_out), "mop_project_trrot: nothing to project");
  MOP_REQUIRE(!Hp_out || H, "mop_project_trrot: H required with Hp_out");
  return mop_launch_project_trrot(B, n, H, Hbias, x, g, Hp_out, gp_out, status, 0,
                                  (cudaStream_t)stream);
}

// scratch layouts: two-read W pass 12 np + 8 doubles per structure; one-read W pass adds [nslab][6][np]
static size_t prj_stride(int n, int one_read) {
  const size_t np = (size_t)((n + 3) & ~3);
  return mop::prj_scratch_doubles(n) + (one_read ? (size_t)((n + 31) / 32) * 6 * np : 0);
}
static size_t prj_w1_smem(int n) { return sizeof(double) * (size_t)((n + 3) & ~3) * (6 + 8 * 6); }
size_t mop_project_scratch_bytes(int B, int n) { return sizeof(double) * (size_t)B * prj_stride(n, 0); }
// preferred scratch: enough for the one-read W pass (the launcher falls back to two reads with less)
size_t mop_project_scratch_bytes_pref(int B, int n) {
  return sizeof(double) * (size_t)B * prj_stride(n, prj_w1_smem(n) <= 100 * 1024);
}

// Same contract as mop_launch_project_trrot, four multi-CTA kernels, `scratch` from the caller.
int mop_launch_project_trrot_split(int B, int n, const double* H, const double* Hbias, const double* x,
                                   const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  const int np = (n + 3) & ~3;
  if (!scratch || scratch_bytes < mop_project_scratch_bytes(B, n)) {
    mop_set_error("projection: scratch too small");
    return MOP_ERR_WORKSPACE;
  }
  const size_t smem0 = sizeof(double) * (12 * (size_t)np + 40);
  if (smem0 > 200 * 1024) {
    mop_set_error("n = %d too large for the projection kernel", n);
    return MOP_ERR_UNSUPPORTED;
  }
  const int nslab = (n + 31) / 32;
  const int one_read = prj_w1_smem(n) <= 100 * 1024 && scratch_bytes >= sizeof(double) * (size_t)B * prj_stride(n, 1);
  const size_t sstride = prj_stride(n, one_read);
  double* scr = (double*)scratch;
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_prj_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
  mop::k_prj_basis<<<B, mop::PRJ_THREADS, smem0, stream>>>(n, grad_rule, sstride, x, g, gp_out, scr, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  if (!Hp_out) return MOP_OK;
  dim3 grid(nslab, B);
  if (one_read) {
    const size_t smem = prj_w1_smem(n);
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_prj_w1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_prj_w1<<<grid, mop::PRJ_THREADS, smem, stream>>>(n, sstride, H, Hbias, scr);
  } else {
    const size_t smem = sizeof(double) * (6 * (size_t)np + 8 * 6 * 32);
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_prj_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_prj_w<<<grid, mop::PRJ_THREADS, smem, stream>>>(n, sstride, H, Hbias, scr);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  mop::k_prj_y<<<B, mop::PRJ_THREADS, 0, stream>>>(n, sstride, one_read ? nslab : 0, scr);
  MOP_CHECK_CUDA(cudaGetLastError());
  {
    const int TT = (n + mop::PT - 1) / mop::PT;
    dim3 grid2(TT, B);
    mop::k_prj_out<<<grid2, mop::PRJ_THREADS, 0, stream>>>(n, sstride, TT, H, Hbias, scr, Hp_out);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  return MOP_OK;
}
