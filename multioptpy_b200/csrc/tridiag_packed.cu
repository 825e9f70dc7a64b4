// Householder tridiagonalisation on PACKED lower-triangular storage (n <= 160).
//
// The full-storage shared-memory kernel (eigh_tridiag.cu) needs 8 n^2 bytes, i.e. one CTA per SM at
// n = 150, and its column step is latency bound (four barriers, two block reductions and the
// Householder scalar chain per column: issue slots are ~60 % idle).  Here the matrix is kept as the
// packed lower triangle, 4 n (n+1) bytes = 90.6 KB at n = 150, so TWO CTAs (two structures) share an
// SM and fill each other's stalls.  LAPACK dsytd2 (lower) conventions; the reflectors, T and Q^T g are
// handed to k_eigh_tridiag (prefactored mode) through global memory (L2).
//
// Column step k, m = n-k-1 trailing rows, S = THREADS / m_pad threads per index:
//   symv   thread (i, s): every S-th term of  p_i = sum_{j<=i} L(i,j) v_j + sum_{r>i} L(r,i) v_r
//          (row part contiguous in the packed row, column part conflict-free across threads),
//          partial sums combined by shuffles inside the S-group - no shared-memory reduction;
//   update thread t owns the row pair (k+1+t, n-1-t): every thread touches m+1 elements.
#include "common.cuh"

namespace mop {

struct PkArgs {
  int n;
  const double* A;   // [B][n][n] symmetric input (projected Hessian)
  const double* gp;  // [B][n] or null
  double* Vh;        // [B][n][n] reflector k in row k, columns k+2.. (unit entry implicit)
  double* dd;        // [B][n]
  double* ee;        // [B][n]
  double* tau;       // [B][n]
  double* gq;        // [B][n] Q^T gp
  int* flag;         // [B] 0 normal, 1 zero matrix, 2 non-finite input
  long long* dbg;    // optional [B][8] phase cycles
};

// Padded packed lower triangle: row i (i + 1 entries) is allotted 16 ceil(i / 16) + 1 doubles, so consecutive
// rows start one double-word bank apart (mod 16): the sixteen lanes of a shared-memory wavefront that read the
// same column of sixteen consecutive rows - the access pattern of the symv row part, the rank-2 update and the
// Householder column - hit sixteen different banks.  (The plain triangle T(i) = i (i + 1) / 2 is a permutation
// of the banks only for some i: ncu counted 39 % excess wavefronts.)  10 % more shared memory: 99.8 KB at
// n = 150, still two CTAs per SM.
__device__ __host__ __forceinline__ int pk_row(int i) {  // start of row i; i = 0 falls out of the formula (q = -1)
  const int q = (i - 1) >> 4, rem = (i - 1) & 15;
  return i + 16 * (q + 1) * (8 * q + rem);
}
__device__ __forceinline__ int pk_len(int r) { return ((r + 15) & ~15) + 1; }  // pk_row(r + 1) - pk_row(r)
__device__ __forceinline__ int pk(int i, int j) { return pk_row(i) + j; }  // j <= i; n <= 160: fits int
// smallest power of two S <= 32 with 2 S m > threads, as log2 (the thread-per-index split of the column step)
__device__ __forceinline__ int split_log2(int threads, int m) {
  const int lg = 32 - __clz(threads / (2 * m));
  return lg < 5 ? lg : 5;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_tridiag_packed(PkArgs a) {
  extern __shared__ double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int np = (n + 3) & ~3;
  double* L = sm;                                  // pk_row(n) doubles
  double* v = L + (((size_t)pk_row(n) + 3) & ~(size_t)3);  // np
  double* w = v + np;                              // np
  double* gq = w + np;                             // np
  __shared__ double s_rb[2 * 32 * 2];  // every reduction is a block_sum_k<2>: the two parity buffers stay disjoint
  int parity = 0;
  const double* Ain = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;

  double pn[2] = {0.0, 0.0};
  for (int idx = tid; idx < n * n; idx += THREADS) {
    const int i = idx / n, j = idx - i * n;
    const double x = Ain[idx];
    if (j <= i) L[pk(i, j)] = x;
    pn[0] = fma(x, x, pn[0]);
  }
  for (int i = tid; i < n; i += THREADS) gq[i] = a.gp ? a.gp[(size_t)b * n + i] : 0.0;
  block_sum_k<2>(pn, s_rb, parity);
  const double fro = sqrt(pn[0]);
  const bool nonfinite = !isfinite(fro), trivial = nonfinite || fro == 0.0;
  if (tid == 0) a.flag[b] = nonfinite ? 2 : (fro == 0.0 ? 1 : 0);
  if (trivial || n <= 2) {
    for (int i = tid; i < n; i += THREADS) {
      a.dd[(size_t)b * n + i] = nonfinite ? NAN : (trivial ? 0.0 : L[pk(i, i)]);
      a.ee[(size_t)b * n + i] = (!trivial && i + 1 < n) ? L[pk(i + 1, i)] : 0.0;
      a.tau[(size_t)b * n + i] = 0.0;
      a.gq[(size_t)b * n + i] = gq[i];
    }
    return;
  }
  // norm^2 of column 0 below the sub-diagonal
  double xn[2] = {0.0, 0.0};
  for (int i = 2 + tid; i < n; i += THREADS) xn[0] = fma(L[pk(i, 0)], L[pk(i, 0)], xn[0]);
  block_sum_k<2>(xn, s_rb, parity);
  double xn2 = xn[0];

  long long seg[6] = {0, 0, 0, 0, 0, 0}, ts = clock64();
#define PSEG(i)                             \
  do {                                      \
    if (a.dbg) {                            \
      const long long tn_ = clock64();      \
      seg[i] += tn_ - ts;                   \
      ts = tn_;                             \
    }                                       \
  } while (0)
  for (int k = 0; k < n - 2; ++k) {
    const int m = n - k - 1;  // trailing order; indices k+1 .. n-1
    const double alpha = L[pk(k + 1, k)];
    double beta = alpha, tk = 0.0, scal = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
      tk = (beta - alpha) * fast_rcp(beta);
      scal = fast_rcp(alpha - beta);
    }
    // v (unit entry explicit in shared memory), reflector row to global
    for (int i = k + 1 + tid; i < n; i += THREADS) {
      const double vi = (i == k + 1) ? 1.0 : L[pk(i, k)] * scal;
      v[i] = vi;
      Vh[(size_t)k * n + i] = vi;  // unit entry written too (the large-n consumers read it, the fused kernel ignores it)
    }
    if (tid == 0) {
      a.dd[(size_t)b * n + k] = L[pk(k, k)];
      a.ee[(size_t)b * n + k] = beta;
      a.tau[(size_t)b * n + k] = tk;
    }
    __syncthreads();  // (A) v complete
    PSEG(0);
    double red[2] = {0.0, 0.0};
    double pi = 0.0;
    // S threads per index, S a power of two <= 32 (a group never straddles a warp).  Lanes are PART major: a
    // half-warp (one LDS.64 wavefront) holds W = min(32 / S, 16) consecutive indices; for S <= 2 (m > 64,
    // where the time goes) these are sixteen consecutive indices of ONE part, and with the padded rows every
    // wavefront below is conflict free.  Part s takes the terms s, s + S, ... (balanced to one term).
    const int lgS = split_log2(THREADS, m), S = 1 << lgS;
    const int per = 32 >> lgS, lgW = (5 - lgS) < 4 ? (5 - lgS) : 4, W = 1 << lgW;
    const int t_idx = (tid >> 5) * per + (lane & (per - 1)), s = lane >> (5 - lgS);
    const bool act = t_idx < m;
    const int i = k + 1 + t_idx;
    if (tk != 0.0) {
      if (act) {
        // row part: terms q = 0 .. i-k-1 -> L(i, k+1+q) v_{k+1+q}
        const int nrow = i - k;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double* Li = L + pk_row(i) + k + 1;
        const double* vq = v + k + 1;
        int q = s;
        for (; q + 3 * S < nrow; q += 4 * S) {  // four independent loads in flight
          const double l0 = Li[q], l1 = Li[q + S], l2 = Li[q + 2 * S], l3 = Li[q + 3 * S];
          const double u0 = vq[q], u1 = vq[q + S], u2 = vq[q + 2 * S], u3 = vq[q + 3 * S];
          a0 = fma(l0, u0, a0);
          a1 = fma(l1, u1, a1);
          a2 = fma(l2, u2, a2);
          a3 = fma(l3, u3, a3);
        }
        for (; q < nrow; q += S) a0 = fma(Li[q], vq[q], a0);
        // column part: rows r > i -> L(r, i) v_r.  The W indices of a group walk the SAME rows (from the
        // group's first index + 1, the rows up to the own index masked out): L(r, i .. i+W-1) is contiguous
        // and v_r a broadcast.
        int r = i - (lane & (W - 1)) + 1 + s;
        if (S == 1) {  // m > 128: consecutive rows, the offset advances by the row allotment
          int off = pk_row(r) + i;
          for (; r + 3 < n; r += 4) {
            const int o1 = off + pk_len(r), o2 = o1 + pk_len(r + 1), o3 = o2 + pk_len(r + 2);
            const double l0 = L[off], l1 = L[o1], l2 = L[o2], l3 = L[o3];
            const double u0 = v[r], u1 = v[r + 1], u2 = v[r + 2], u3 = v[r + 3];
            a0 = fma(r > i ? l0 : 0.0, u0, a0);
            a1 = fma(r + 1 > i ? l1 : 0.0, u1, a1);
            a2 = fma(r + 2 > i ? l2 : 0.0, u2, a2);
            a3 = fma(r + 3 > i ? l3 : 0.0, u3, a3);
            off = o3 + pk_len(r + 3);
          }
          for (; r < n; ++r) {
            a0 = fma(r > i ? L[off] : 0.0, v[r], a0);
            off += pk_len(r);
          }
        } else {
          for (; r + 3 * S < n; r += 4 * S) {
            const double l0 = L[pk_row(r) + i], l1 = L[pk_row(r + S) + i], l2 = L[pk_row(r + 2 * S) + i],
                         l3 = L[pk_row(r + 3 * S) + i];
            const double u0 = v[r], u1 = v[r + S], u2 = v[r + 2 * S], u3 = v[r + 3 * S];
            a0 = fma(r > i ? l0 : 0.0, u0, a0);
            a1 = fma(r + S > i ? l1 : 0.0, u1, a1);
            a2 = fma(r + 2 * S > i ? l2 : 0.0, u2, a2);
            a3 = fma(r + 3 * S > i ? l3 : 0.0, u3, a3);
          }
          for (; r < n; r += S) a0 = fma(r > i ? L[pk_row(r) + i] : 0.0, v[r], a0);
        }
        a0 += a2;
        a1 += a3;
        pi = a0 + a1;
      }
      PSEG(1);
      for (int o = per; o < 32; o <<= 1) pi += __shfl_xor_sync(MOP_FULL_MASK, pi, o);  // the S parts of an index
      pi *= tk;
      if (act && s == 0) {
        const double vi = v[i];
        red[0] = pi * vi;
        red[1] = vi * gq[i];
      }
      block_sum_k<2>(red, s_rb, parity);  // (C)
      PSEG(2);
      const double alpha2 = -0.5 * tk * red[0];
      if (act && s == 0) {
        const double vi = v[i];
        w[i] = fma(alpha2, vi, pi);
        gq[i] = fma(-tk * red[1], vi, gq[i]);
      }
      __syncthreads();  // (D) w complete
      PSEG(3);
      // rank-2 update of the lower triangle, row pairs (k+1+t, n-1-t): m+1 elements per pair
      double nx[2] = {0.0, 0.0};
      {
        const int npair = (m + 1) >> 1;
        const int lgS2 = split_log2(THREADS, npair), S2 = 1 << lgS2;
        const int per2 = 32 >> lgS2;
        const int t2 = (tid >> 5) * per2 + (lane & (per2 - 1)), s2 = lane >> (5 - lgS2);
        if (t2 < npair) {
          const int rows[2] = {k + 1 + t2, n - 1 - t2};
          const int nr = (rows[0] == rows[1]) ? 1 : 2;
          for (int h = 0; h < nr; ++h) {
            const int r = rows[h];
            const double vr = v[r], wr = w[r];
            double* Lr = L + pk_row(r) + k + 1;
            const int len = r - k;  // columns k+1 .. r
            const double* wj = w + k + 1;
            const double* vj = v + k + 1;
            int c = s2;
            if (c == 0) {  // first column of the trailing block: the next Householder column
              const double x = Lr[0] - fma(vr, wj[0], wr * vj[0]);
              Lr[0] = x;
              if (r >= k + 3) nx[0] = fma(x, x, nx[0]);
              c += S2;
            }
            for (; c + 3 * S2 < len; c += 4 * S2) {  // loads first: the compiler cannot prove Lr, w, v disjoint
              const double l0 = Lr[c], l1 = Lr[c + S2], l2 = Lr[c + 2 * S2], l3 = Lr[c + 3 * S2];
              const double w0 = wj[c], w1 = wj[c + S2], w2 = wj[c + 2 * S2], w3 = wj[c + 3 * S2];
              const double u0 = vj[c], u1 = vj[c + S2], u2 = vj[c + 2 * S2], u3 = vj[c + 3 * S2];
              Lr[c] = l0 - fma(vr, w0, wr * u0);
              Lr[c + S2] = l1 - fma(vr, w1, wr * u1);
              Lr[c + 2 * S2] = l2 - fma(vr, w2, wr * u2);
              Lr[c + 3 * S2] = l3 - fma(vr, w3, wr * u3);
            }
            for (; c < len; c += S2) Lr[c] = Lr[c] - fma(vr, wj[c], wr * vj[c]);
          }
        }
      }
      PSEG(4);
      block_sum_k<2>(nx, s_rb, parity);  // (E) also publishes the updated triangle
      PSEG(5);
      xn2 = nx[0];
    } else {
      double nx[2] = {0.0, 0.0};
      for (int r = k + 3 + tid; r < n; r += THREADS) nx[0] = fma(L[pk(r, k + 1)], L[pk(r, k + 1)], nx[0]);
      block_sum_k<2>(nx, s_rb, parity);
      xn2 = nx[0];
    }
  }
  if (tid == 0) {
    a.dd[(size_t)b * n + n - 2] = L[pk(n - 2, n - 2)];
    a.ee[(size_t)b * n + n - 2] = L[pk(n - 1, n - 2)];
    a.tau[(size_t)b * n + n - 2] = 0.0;
    a.dd[(size_t)b * n + n - 1] = L[pk(n - 1, n - 1)];
    a.ee[(size_t)b * n + n - 1] = 0.0;
    a.tau[(size_t)b * n + n - 1] = 0.0;
  }
  for (int i2 = tid; i2 < n; i2 += THREADS) a.gq[(size_t)b * n + i2] = gq[i2];
  if (a.dbg && (tid == 0 || tid == 96))
    for (int q = 0; q < 6; ++q) a.dbg[(size_t)b * 16 + (tid == 0 ? 0 : 8) + q] = seg[q];
}


// ------------------------------------------------------------------------------------------------
// Row-per-warp mapping (k_tridiag_rwf below).  ncu on k_tridiag_packed: the shared-memory pipe is ~70 % busy and every
// element of the triangle is read twice by the symv (row part + column gather) with one more load of v or
// w per FMA.  Here a WARP owns a row and its LANES the columns, for the symv and for the rank-2 update:
//   symv    x = L(i, c) is read once, contiguously; it feeds the row sum (x v_c, v_c in a register, one
//           butterfly per row) and the lane's column accumulator (x v_i, v_i a broadcast) - the transposed
//           contribution p_c += L(i, c) v_i - which stays in registers over all rows of the warp and is
//           combined across the eight warps through shared memory;
//   update  L(r, c) -= v_r w_c + w_r v_c with v_c, w_c in registers: one load and one store per element.
// Plain triangle T(i) = i (i + 1) / 2 (row accesses are contiguous, no padding needed); shared memory:
// L | v | w | gq | u | prow | part[8][np] = 106 KB at n = 150, two CTAs per SM.
__device__ __forceinline__ int pk0(int i, int j) { return ((i * (i + 1)) >> 1) + j; }

// ------------------------------------------------------------------------------------------------
// Fused variant (default): the rank-2 update of reflector k-1 and the symv of reflector k in ONE pass over
// the triangle.  The symv needs v_k, which needs the norm of the updated column k - but the symv is linear:
// with u = the raw updated column k (rows > k) and v_k = s u + (1 - s alpha) e_{k+1}  (s = 1 / (alpha - beta)),
//     A' v_k = s (A' u) + (1 - s alpha) A'(:, k+1),
// so the pass accumulates q = A' u while it writes A' = A - v w^T - w v^T, and the scalars join afterwards.
// Per column: (b) the fused pass  (P)  (c) q, the sums S1..S4 and three scalars through ONE block reduction
// (C), then beta, t, s, p, w, v AND the raw next column u' from the values already in registers  (D): three
// barriers and one read + one write of every element (separate passes: five barriers, two reads + one write).
// Row blocks of FOUR consecutive rows per warp (register budget: v, w, u and the column sums of 5 column
// chunks per lane plus four rows in flight), dealt longest first in serpentine order.
//
// Block-wide sums of EIGHT values for 8-warp CTAs, one barrier (double-buffered like block_sum_k): a
// transposing butterfly leaves value j with the lanes whose bits 4..2 spell j (9 shuffles instead of 40), the
// eight per-warp partials of each value are combined by every warp and handed to all lanes.
__device__ __forceinline__ void block_sum8(double (&r)[8], double* buf, int& parity, int lane, int wid) {
  double* bq = buf + (parity & 1) * 64;
  parity ^= 1;
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  double t4[4], t2[2], t1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double send = h16 ? r[j] : r[j + 4];
    const double keep = h16 ? r[j + 4] : r[j];
    t4[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const double send = h8 ? t4[j] : t4[j + 2];
    const double keep = h8 ? t4[j + 2] : t4[j];
    t2[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 8);
  }
  {
    const double send = h4 ? t2[0] : t2[1];
    const double keep = h4 ? t2[1] : t2[0];
    t1 = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 4);
  }
  t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 2);
  t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 1);
  const int j = lane >> 2;  // value index held by this lane: (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0)
  if ((lane & 3) == 0) bq[j * 8 + wid] = t1;
  __syncthreads();
  double t = bq[j * 8 + (lane & 3)] + bq[j * 8 + (lane & 3) + 4];  // warps sub and sub + 4 of value j
  t += __shfl_xor_sync(MOP_FULL_MASK, t, 1);
  t += __shfl_xor_sync(MOP_FULL_MASK, t, 2);
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = __shfl_sync(MOP_FULL_MASK, t, 4 * q);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) k_tridiag_rwf(PkArgs a) {
  constexpr int NW = THREADS / 32;
  constexpr int MAXU = 5;  // n <= 160 columns over 32 lanes
  extern __shared__ double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  double* L = sm;                                                 // n (n + 1) / 2
  double* v = L + (((size_t)n * (n + 1) / 2 + 3) & ~(size_t)3);   // np   reflector being applied
  double* w = v + np;                                             // np
  double* gq = w + np;                                            // np
  double* uu = gq + np;                                           // np   raw next column
  double* prow = uu + np;                                         // np   row sums of the symv
  double* part = prow + np;                                       // NW * np  column sums per warp
  __shared__ double s_rb[2 * 32 * 2];
  __shared__ double s_r8[2 * 8 * 8];
  int parity = 0;
  const double* Ain = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;

  double pn[2] = {0.0, 0.0};
  for (int idx = tid; idx < n * n; idx += THREADS) {
    const int i = idx / n, j = idx - i * n;
    const double x = Ain[idx];
    if (j <= i) L[pk0(i, j)] = x;
    pn[0] = fma(x, x, pn[0]);
  }
  for (int i = tid; i < n; i += THREADS) {
    gq[i] = a.gp ? a.gp[(size_t)b * n + i] : 0.0;
    v[i] = 0.0;  // "reflector -1": nothing to apply in the first pass
    w[i] = 0.0;
  }
  block_sum_k<2>(pn, s_rb, parity);
  const double fro = sqrt(pn[0]);
  const bool nonfinite = !isfinite(fro), trivial = nonfinite || fro == 0.0;
  if (tid == 0) a.flag[b] = nonfinite ? 2 : (fro == 0.0 ? 1 : 0);
  if (trivial || n <= 2) {
    for (int i = tid; i < n; i += THREADS) {
      a.dd[(size_t)b * n + i] = nonfinite ? NAN : (trivial ? 0.0 : L[pk0(i, i)]);
      a.ee[(size_t)b * n + i] = (!trivial && i + 1 < n) ? L[pk0(i + 1, i)] : 0.0;
      a.tau[(size_t)b * n + i] = 0.0;
      a.gq[(size_t)b * n + i] = gq[i];
    }
    return;
  }

  // raw column 0 and d_0; "reflector -1" is zero
  for (int i = tid; i < n; i += THREADS) uu[i] = i > 0 ? L[pk0(i, 0)] : 0.0;
  if (tid == 0) a.dd[(size_t)b * n] = L[0];
  __syncthreads();

  long long seg[6] = {0, 0, 0, 0, 0, 0}, ts = clock64();
#define FSEG(i)                             \
  do {                                      \
    if (a.dbg) {                            \
      const long long tn_ = clock64();      \
      seg[i] += tn_ - ts;                   \
      ts = tn_;                             \
    }                                       \
  } while (0)
  // Invariant at the top of iteration k: v, w = reflector k-1 (not yet applied to L), uu = the raw updated
  // column k (rows > k; uu[k] = 0), d_k written.
  for (int k = 0; k < n - 2; ++k) {
    const int m = n - k;  // order of the block the pass updates (rows / columns k .. n-1)
    // ---- (b) fused pass: warp = four consecutive rows, lane = column ---------------------------------
    double vreg[MAXU], wreg[MAXU], ureg[MAXU], cacc[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int c = k + lane + 32 * u;
      const bool in = c < n;
      vreg[u] = in ? v[c] : 0.0;
      wreg[u] = in ? w[c] : 0.0;
      ureg[u] = in ? uu[c] : 0.0;
      cacc[u] = 0.0;
    }
    const int nblk = (m + 3) >> 2;
    for (int pb = 0; pb * NW < nblk; ++pb) {
      const int blk = nblk - 1 - (pb * NW + ((pb & 1) ? NW - 1 - wid : wid));  // longest blocks first, serpentine
      if (blk < 0) continue;  // warp-uniform
      const int rb = k + 4 * blk;
      const int nu = ((min(rb + 3, n - 1) - k) >> 5) + 1;  // chunks of the longest row of the block
      double vr[4], wr[4], ur[4], rs[4];
      int ro[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rb + j;
        const bool in = r < n;
        vr[j] = in ? v[r] : 0.0;
        wr[j] = in ? w[r] : 0.0;
        ur[j] = in ? uu[r] : 0.0;
        rs[j] = 0.0;
        ro[j] = pk0(in ? r : n - 1, 0);
      }
      const bool full4 = rb + 3 < n;  // warp-uniform: all four rows exist
#pragma unroll
      for (int u = 0; u < MAXU; ++u) {
        if (u < nu) {  // warp-uniform
          const int c = k + lane + 32 * u;
          double x[4];
          if (full4 && k + 32 * u + 31 <= rb) {
            // interior chunk: every lane's column lies at or below the diagonal of all four rows - no masks
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = L[ro[j] + c];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              x[j] -= fma(vr[j], wreg[u], wr[j] * vreg[u]);
              L[ro[j] + c] = x[j];
              rs[j] = fma(x[j], ureg[u], rs[j]);
              cacc[u] = fma(x[j], ur[j], cacc[u]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (c <= rb + j && rb + j < n) ? L[ro[j] + c] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool in = c <= rb + j && rb + j < n;
              x[j] -= fma(vr[j], wreg[u], wr[j] * vreg[u]);
              if (in) L[ro[j] + c] = x[j];
              const double xm = in ? x[j] : 0.0;
              // the diagonal element also lands in the column accumulator; (c) takes it out again
              rs[j] = fma(xm, ureg[u], rs[j]);
              cacc[u] = fma(xm, ur[j], cacc[u]);
            }
          }
        }
      }
      // transposing butterfly: 4 row sums over 32 lanes in 6 shuffles
      const bool h16 = lane & 16, h8 = lane & 8;
      double t2[2], t1;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double send = h16 ? rs[j] : rs[j + 2];
        const double keep = h16 ? rs[j + 2] : rs[j];
        t2[j] = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 16);
      }
      {
        const double send = h8 ? t2[0] : t2[1];
        const double keep = h8 ? t2[1] : t2[0];
        t1 = keep + __shfl_xor_sync(MOP_FULL_MASK, send, 8);
      }
      t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 4);
      t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 2);
      t1 += __shfl_xor_sync(MOP_FULL_MASK, t1, 1);
      if ((lane & 7) == 0) {  // lanes with bit 4 set hold rows 2-3, bit 3 adds 1
        const int i = rb + (h16 ? 2 : 0) + (h8 ? 1 : 0);
        if (i < n) prow[i] = t1;
      }
    }
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      const int c = k + lane + 32 * u;
      if (c < n) part[wid * np + c] = cacc[u];
    }
    __syncthreads();  // (P) updated triangle, row sums and column partials complete
    FSEG(1);
    // ---- (c) q = A' u; every scalar of the step from ONE block reduction -------------------------------
    //   S1 = sum q_i u_i, S2 = sum c_i u_i, S3 = sum u_i gq_i, S4 = sum u_i^2 over i >= k+2 (c = A'(:, k+1)),
    //   and q, c, gq at i = k+1.  Then beta, t, s;  p = t (s q + (1 - s alpha) c),  v = s u (v_{k+1} = 1),
    //   p.v = t [(s q0 + ca c0) + s (s S1 + ca S2)],  v.gq = gq0 + s S3,  w = p - t/2 (p.v) v,
    //   and the raw NEXT column  u'_i = c_i - v_i w_{k+1} - w_i  comes from the same registers.
    const int i = k + 1 + tid;
    const double alpha = uu[k + 1];
    double ui = 0.0, qi = 0.0, ci = 0.0, gi = 0.0;
    double rd[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < n) {
      ui = uu[i];
      qi = fma(-L[pk0(i, i)], ui, prow[i]);  // the diagonal term was counted in both sums
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) qi += part[ww * np + i];
      ci = L[pk0(i, k + 1)];
      gi = gq[i];
      if (tid == 0) {
        rd[4] = qi;
        rd[5] = ci;
        rd[6] = gi;
      } else {
        rd[0] = qi * ui;
        rd[1] = ci * ui;
        rd[2] = ui * gi;
        rd[3] = ui * ui;
      }
    }
    block_sum8(rd, s_r8, parity, lane, wid);  // (C)
    FSEG(2);
    const double xn2 = rd[3];
    double beta = alpha, tk = 0.0, scal = 0.0;
    if (xn2 > 0.0) {
      beta = -copysign(sqrt(fma(alpha, alpha, xn2)), alpha);
      tk = (beta - alpha) * fast_rcp(beta);
      scal = fast_rcp(alpha - beta);
    }
    const double ca = 1.0 - scal * alpha;
    const double p0 = tk * fma(scal, rd[4], ca * rd[5]);
    const double pv = p0 + tk * scal * fma(scal, rd[0], ca * rd[1]);
    const double vg = fma(scal, rd[2], rd[6]);
    const double alpha2 = -0.5 * tk * pv;
    const double w0 = p0 + alpha2;  // w_{k+1}
    if (i < n) {
      const double vi = (tid == 0) ? 1.0 : ui * scal;
      const double pi = tk * fma(scal, qi, ca * ci);
      const double wi = fma(alpha2, vi, pi);
      const double un = ci - fma(vi, w0, wi);  // raw column k+1 after reflector k
      v[i] = vi;
      w[i] = wi;
      gq[i] = fma(-tk * vg, vi, gi);
      uu[i] = tid == 0 ? 0.0 : un;
      Vh[(size_t)k * n + i] = vi;
      if (tid == 0) {
        a.ee[(size_t)b * n + k] = beta;
        a.tau[(size_t)b * n + k] = tk;
        a.dd[(size_t)b * n + k + 1] = un;
      }
    }
    __syncthreads();  // (D) v, w of reflector k and the raw column k+1 complete
    FSEG(3);
  }
  // reflector n-3 is still to be applied to the last diagonal element; e_{n-2} is the raw column n-2
  if (tid == 0) {
    const int i0 = n - 2, i1 = n - 1;
    a.ee[(size_t)b * n + i0] = uu[i1];
    a.tau[(size_t)b * n + i0] = 0.0;
    a.dd[(size_t)b * n + i1] = L[pk0(i1, i1)] - 2.0 * v[i1] * w[i1];
    a.ee[(size_t)b * n + i1] = 0.0;
    a.tau[(size_t)b * n + i1] = 0.0;
  }
  for (int i2 = tid; i2 < n; i2 += THREADS) a.gq[(size_t)b * n + i2] = gq[i2];
  if (a.dbg && (tid == 0 || tid == 96))
    for (int q = 0; q < 6; ++q) a.dbg[(size_t)b * 16 + (tid == 0 ? 0 : 8) + q] = seg[q];
#undef FSEG
}

}  // namespace mop

size_t mop_tridiag_packed_smem(int n) {
  const int np = (n + 3) & ~3;
  return sizeof(double) * ((((size_t)mop::pk_row(n) + 3) & ~(size_t)3) + 3 * (size_t)np);
}

size_t mop_tridiag_rw_smem(int n) {
  const int np = (n + 3) & ~3;
  return sizeof(double) * ((((size_t)n * (n + 1) / 2 + 3) & ~(size_t)3) + (5 + 8) * (size_t)np);
}
int mop_launch_tridiag_blk(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                           double* tau, double* gq, int* flag, cudaStream_t stream);
static int g_pk_blocked = 1;
// tuning: 1 (default) = k_tridiag_blk (tridiag_blocked.cu: dlatrd panels, thread-per-row symv, DMMA trailing updates)
extern "C" int mop_debug_packed_blocked(int on) {
  g_pk_blocked = on;
  return MOP_OK;
}
static int g_pk_rowwarp = 1;
// tuning: 1 (default) = k_tridiag_rwf (warp per row block, lanes over columns, update + symv fused),
// 0 = k_tridiag_packed (thread groups per index, separate symv and update passes)
extern "C" int mop_debug_packed_rowwarp(int on) {
  g_pk_rowwarp = on;
  return MOP_OK;
}
static int g_pk_threads = 256;
static long long* g_pk_dbg = nullptr;
extern "C" int mop_debug_packed_timing(void* buf) {
  g_pk_dbg = (long long*)buf;
  return MOP_OK;
}
extern "C" int mop_debug_packed_threads(int t) {
  g_pk_threads = t;
  return MOP_OK;
}

// d, e, tau, gq: [B][n]; Vh: [B][n][n]; flag: [B]
int mop_launch_tridiag_packed(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                              double* tau, double* gq, int* flag, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (g_pk_blocked) return mop_launch_tridiag_blk(B, n, A, gp, Vh, dd, ee, tau, gq, flag, stream);
  mop::PkArgs a{n, A, gp, Vh, dd, ee, tau, gq, flag, g_pk_dbg};
  if (g_pk_rowwarp) {
    const size_t smem_rw = mop_tridiag_rw_smem(n);
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_rwf<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rw));
    mop::k_tridiag_rwf<256><<<B, 256, smem_rw, stream>>>(a);
    MOP_CHECK_CUDA(cudaGetLastError());
    return MOP_OK;
  }
  const size_t smem = mop_tridiag_packed_smem(n);
  if (g_pk_threads == 320) {
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_packed<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_tridiag_packed<320><<<B, 320, smem, stream>>>(a);
  } else if (g_pk_threads == 512) {
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_packed<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_tridiag_packed<512><<<B, 512, smem, stream>>>(a);
  } else {
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_packed<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_tridiag_packed<256><<<B, 256, smem, stream>>>(a);
  }
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
