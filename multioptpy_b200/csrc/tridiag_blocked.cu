// Blocked Householder tridiagonalisation on the packed lower triangle in shared memory (n <= 160):
// LAPACK dlatrd panels, FP64 tensor-core (DMMA, mma.sync.m8n8k4) rank-2NB trailing updates.
//
// Why: the unblocked fused kernel (k_tridiag_rwf, tridiag_packed.cu) reads AND writes every element of the
// trailing triangle in every column step and pays a transposing butterfly per four rows plus a cross-warp
// combine of column sums: ncu counted 1.26e9 warp instructions per 1024 structures at n = 150, nine times
// the element work, issue-bound at 59 % of the slots.  Here
//   * the column step only READS the triangle (symv on the panel-start matrix, LAPACK dlatrd); the rank-2
//     updates of a panel of NB = 6 reflectors are applied once per panel as a rank-12 update, 8 x 8 tiles on
//     the FP64 tensor cores (three m8n8k4 DMMAs per tile instead of 24 DFMAs per element);
//   * the symv is THREAD PER ROW: thread i owns p_i = sum_j A(i, j) u_j completely (row part to the left of
//     the diagonal, column part below it), so there is no butterfly and no cross-warp combine; the row part
//     is walked by a permuted lane -> row map (even rows in the low half-warp, odd rows in the high one),
//     which makes the plain triangle T(i) = i (i + 1) / 2 bank-conflict free without padding, because
//     T(2k) mod 16 and T(2k + 1) mod 16 are permutations over any 16 consecutive k;
//   * as in k_tridiag_rwf the symv runs on the RAW updated column u (v = s u + (1 - s alpha) e_{k+1} is
//     linear in it), so it does not wait for the Householder norm, and EVERY scalar product of the column
//     step - norm, p.v, v.Q^T g and the 2 (NB - 1) panel products V^T u, W^T u of dlatrd's correction
//     w -= V (W^T v) + W (V^T v) - goes through ONE 16-value block reduction (transposing butterfly, one
//     barrier).  Two barriers per column.
// The reflectors are kept where LAPACK keeps them (column k of the triangle, unit entry explicit), the W
// panel in a [n][10] array (rows 16-byte aligned, conflict-free 128-bit row reads).  Shared memory at
// n = 150: 90.6 KB triangle + 12 KB panel + 6 KB vectors = 108 KB, two CTAs per SM.
//
// Outputs are those of mop_launch_tridiag_packed (LAPACK dsytd2 conventions): d, e, tau, the reflector rows Vh
// and Q^T g.  Replaces the reduction stage of numpy.linalg.eigh at Optimizer/rsirfo.py:606,626,652.
#include "common.cuh"

namespace mop {

struct PkArgs {
  int n;
  const double* A;   // [B][n][n] symmetric input (projected Hessian); only the lower triangle is read
  const double* gp;  // [B][n] or null
  double* Vh;        // [B][n][n] reflector k in row k, columns k+1.. (unit entry written)
  double* dd;        // [B][n]
  double* ee;        // [B][n]
  double* tau;       // [B][n]
  double* gq;        // [B][n] Q^T gp
  int* flag;         // [B] 0 normal, 1 zero matrix, 2 non-finite input
  long long* dbg;    // optional [B][16] phase cycles
};

constexpr int TB_NB = 6;   // reflectors per panel: 4 scalars + 2 * NB panel products = the 16 slots of one reduction
constexpr int TB_WS = 10;  // doubles per row of the W panel: 80-byte rows, eight 128-bit row reads hit 32 banks once

__device__ __forceinline__ int tri0(int i) { return (i * (i + 1)) >> 1; }

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
// red: [2][16][NW] (double-buffered by `parity`, so one barrier per call is enough), tot: [NW][16].
template <int NW>
__device__ __forceinline__ void block_sum16(double (&r)[16], double* red, double* tot, int& parity, int lane,
                                            int wid, bool contributes) {
  double* bq = red + (parity & 1) * (16 * NW);
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
    if ((lane & 1) == 0) bq[slot * NW + wid] = t1;
  } else if (lane < 16) {
    bq[lane * NW + wid] = 0.0;
  }
  __syncthreads();
  if (!contributes) return;  // a warp without live rows needs no totals
  double* tw = tot + wid * 16;
  if (lane < 16) {
    double t[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) t[w] = bq[lane * NW + w];
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

// Rank-2NB update of the trailing triangle with one finished panel, A(i, j) -= sum_l V(i, l) W(j, l) + W(i, l) V(j, l)
// for i >= j >= kn, as C + (-P) Q^T with P = [V | W], Q = [W | V] (K = 12): 8 x 8 tiles, three DMMAs each.
// V(i, l) = L(i, k0 + l) (the eliminated columns), W(i, l) = Wp[i][l].  A warp owns whole tile rows (its three
// A fragments stay in registers), dealt longest first in serpentine order, and walks a row four tiles at a
// time so that four independent DMMA chains are in flight.
template <int NW>
__device__ __forceinline__ void trailing_update_dmma(double* L, const double* Wp, int n, int k0, int kn, int lane,
                                                     int wid) {
  const int g = lane >> 2, t = lane & 3;
  const int mt = (n - kn + 7) >> 3;
  for (int rr = wid; rr < mt; rr += NW) {
    const int round = rr / NW, pos = rr - round * NW;
    const int R = mt - round * NW < NW ? mt - round * NW : NW;
    const int I = mt - 1 - (round * NW + ((round & 1) ? R - 1 - pos : pos));
    const int ri = kn + 8 * I + g;
    const int ric = ri < n ? ri : n - 1;  // clamped for addressing
    const int Tr = tri0(ric);
    double a[3];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      const int c = 4 * ks + t;  // column of P
      const double pa = c < TB_NB ? L[Tr + k0 + c] : Wp[ric * TB_WS + c - TB_NB];
      a[ks] = ri < n ? -pa : 0.0;
    }
    for (int J0 = 0; J0 <= I; J0 += 4) {
      double bq[4][3], c0[4], c1[4];
      bool in0[4], in1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int J = J0 + u;
        const int rj = kn + 8 * J + g;
        const bool jin = J <= I && rj < n;
        const int rjc = jin ? rj : n - 1;
        const int Tj = tri0(rjc);
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const int c = 4 * ks + t;  // column of Q
          const double qb = c < TB_NB ? Wp[rjc * TB_WS + c] : L[Tj + k0 + c - TB_NB];
          bq[u][ks] = jin ? qb : 0.0;
        }
        const int cj = kn + 8 * J + 2 * t;
        in0[u] = J <= I && ri < n && cj <= ri;
        in1[u] = J <= I && ri < n && cj + 1 <= ri;
        c0[u] = in0[u] ? L[Tr + cj] : 0.0;
        c1[u] = in1[u] ? L[Tr + cj + 1] : 0.0;
      }
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (J0 + u <= I) dmma884(c0[u], c1[u], a[ks], bq[u][ks], c0[u], c1[u]);  // warp-uniform predicate
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cj = kn + 8 * (J0 + u) + 2 * t;
        if (in0[u]) L[Tr + cj] = c0[u];
        if (in1[u]) L[Tr + cj + 1] = c1[u];
      }
    }
  }
}

// The 32 x 32 diagonal block of warp RW0 / 32 in the symv: index j of the block pairs with L(i, RW0 + j) for
// j <= lane (row access) and with L(RW0 + j, i) below the diagonal (column access); RW0 is a template parameter
// so that every offset of the column access is an immediate.
template <int RW0>
__device__ __forceinline__ void symv_diag32(const double* __restrict__ L, const double* __restrict__ uu, int Ti, int ii,
                                            int lane, int n, double (&acc)[4]) {
  const double* A0 = L + Ti + RW0;
  const double* A1 = L + tri0(RW0) + ii;
#pragma unroll
  for (int jb = 0; jb < 32; jb += 8) {
    if (RW0 + jb >= n) break;  // warp-uniform; rows past n-1 of the last batch read the W panel, times u = 0
    double l[8];
    double2 u2[4];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = jb + q;
      l[q] = (j <= lane) ? A0[j] : A1[j * RW0 + ((j * (j + 1)) >> 1)];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) u2[q] = *reinterpret_cast<const double2*>(uu + RW0 + jb + 2 * q);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[q & 1] = fma(l[2 * q], u2[q].x, acc[q & 1]);
      acc[2 + (q & 1)] = fma(l[2 * q + 1], u2[q].y, acc[2 + (q & 1)]);
    }
  }
}

template <int NW>
struct TbMinBlocks {
  static constexpr int value = NW >= 5 ? 2 : (NW == 4 ? 3 : (NW == 3 ? 5 : (NW == 2 ? 8 : 16)));
};

// shared memory: L | Wp | uu | zrow | gq | red | tot | pub | block_sum_k scratch (doubles)
__host__ __device__ inline size_t tb_smem_doubles(int n, int nw) {
  const size_t np = (size_t)((n + 3) & ~3);
  const size_t nl = ((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1;
  return nl + (size_t)n * TB_WS + (32 * (size_t)nw + 16) + 2 * np + 2 * 16 * (size_t)nw + 16 * (size_t)nw + 4 + 64;
}

template <int NW, bool DBG>
__global__ void __launch_bounds__(32 * NW, TbMinBlocks<NW>::value) k_tridiag_blk(PkArgs a) {
  constexpr int THREADS = 32 * NW;
  constexpr int NB = TB_NB, WS = TB_WS;
  constexpr int NU = 32 * NW + 16;  // uu is zero outside (k, n): the symv loops run in whole batches of eight
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  const size_t nl = ((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1;
  double* L = sm;                   // packed lower triangle, row i at i (i + 1) / 2
  double* Wp = L + nl;              // [n][WS] W panel (directly behind L: symv reads past row n-1 land here, times 0)
  double* uu = Wp + (size_t)n * WS; // [NU] raw updated column (16-byte aligned: nl and n WS are even)
  double* zrow = uu + NU;           // [np] row-part sums of the symv (permuted lane map)
  double* gq = zrow + np;           // [np] Q^T g
  double* red = gq + np;            // [2][16][NW]
  double* tot = red + 2 * 16 * NW;  // [NW][16]
  double* pub = tot + 16 * NW;      // [4]  z_{k+1}, c_{k+1}
  double* s_rb = pub + 4;           // [64] block_sum_k<2> scratch
  int parity = 0, parity2 = 0;
  const double* Ain = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;

  // everything behind the triangle starts finite (the batched symv multiplies a few words of it by zero)
  for (int i = tid; i < (int)(tb_smem_doubles(n, NW) - nl); i += THREADS) Wp[i] = 0.0;
  // ---- load the lower triangle (row segments, coalesced), Frobenius norm of it for the trivial cases ----
  double pn[2] = {0.0, 0.0};
  for (int i = wid; i < n; i += NW) {
    const double* row = Ain + (size_t)i * n;
    double* Lr = L + tri0(i);
    for (int j = lane; j <= i; j += 32) {
      const double x = row[j];
      Lr[j] = x;
      pn[0] = fma(x, x, pn[0]);
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += THREADS) gq[i] = a.gp ? a.gp[(size_t)b * n + i] : 0.0;
  block_sum_k<2>(pn, s_rb, parity2);
  const double fro = sqrt(pn[0]);
  const bool nonfinite = !isfinite(fro), trivial = nonfinite || fro == 0.0;
  if (tid == 0) a.flag[b] = nonfinite ? 2 : (fro == 0.0 ? 1 : 0);
  if (trivial || n <= 2) {
    for (int i = tid; i < n; i += THREADS) {
      a.dd[(size_t)b * n + i] = nonfinite ? NAN : (trivial ? 0.0 : L[tri0(i) + i]);
      a.ee[(size_t)b * n + i] = (!trivial && i + 1 < n) ? L[tri0(i + 1) + i] : 0.0;
      a.tau[(size_t)b * n + i] = 0.0;
      a.gq[(size_t)b * n + i] = gq[i];
    }
    return;
  }
  for (int i = tid; i < n; i += THREADS) uu[i] = i > 0 ? L[tri0(i)] : 0.0;  // raw column 0
  if (tid == 0) a.dd[(size_t)b * n] = L[0];
  __syncthreads();

  long long seg[5] = {0, 0, 0, 0, 0}, ts = DBG ? clock64() : 0;
#define BSEG(q)                             \
  do {                                      \
    if (DBG) {                              \
      const long long tn_ = clock64();      \
      seg[q] += tn_ - ts;                   \
      ts = tn_;                             \
    }                                       \
  } while (0)

  const int i = tid;                        // the row this thread owns
  const int ii = i < n ? i : n - 1;         // clamped for addressing
  const int Ti = tri0(ii);
  const int rw0 = 32 * wid;                 // first row of the warp
  const int iR = rw0 + 2 * (lane & 15) + (lane >> 4);  // permuted row of the off-diagonal row part
  const int iRc = iR < n ? iR : n - 1;
  const double* LR = L + tri0(iRc);
  int k0 = 0;                               // first column of the current panel
  // rows of the current panel owned by this thread, kept in registers across its columns: V(i, l), W(i, l)
  double Vi[NB], Wi[NB];
#pragma unroll
  for (int l = 0; l < NB; ++l) Vi[l] = Wi[l] = 0.0;
  // Invariant at the top of column k: uu = raw updated column k (rows > k, zero elsewhere), d_k written, reflectors
  // k0 .. k-1 of the current panel in L(:, k0 .. k-1) / Wp(:, 0 .. k-k0-1) (and Vi / Wi), the triangle (columns > k)
  // as it was at the panel start.
#pragma unroll 1
  for (int k = 0; k < n - 2; ++k) {
    const int jj = k - k0;
    const bool act = i > k && i < n;
    const bool wact = rw0 + 31 > k && rw0 < n;  // warp-uniform: the warp owns a row > k
    double rd[16];
    double alpha = 0.0, gq0 = 0.0, ui = 0.0, gi = 0.0, ci = 0.0, zi = 0.0;
    if (wact) {
      // ---- (a) c = updated column k+1; panel products with u (rows k+1 of the panel are broadcast loads) --------
      double Vk[NB], Wk[NB];
      {
        const double* wk = Wp + (k + 1) * WS;
        const double* vk = L + tri0(k + 1) + k0;
#pragma unroll
        for (int l = 0; l < NB; l += 2) {
          const double2 y = *reinterpret_cast<const double2*>(wk + l);
          Wk[l] = l < jj ? y.x : 0.0;
          Wk[l + 1] = l + 1 < jj ? y.y : 0.0;
        }
#pragma unroll
        for (int l = 0; l < NB; ++l) Vk[l] = l < jj ? vk[l] : 0.0;
      }
      alpha = uu[k + 1];
      gq0 = gq[k + 1];
      ui = act ? uu[i] : 0.0;
      gi = act ? gq[i] : 0.0;
      ci = act ? L[Ti + k + 1] : 0.0;
      {
        double t[NB];
#pragma unroll
        for (int l = 0; l < NB; ++l) {
          t[l] = fma(Vi[l], Wk[l], Wi[l] * Vk[l]);
          rd[4 + l] = Vi[l] * ui;
          rd[4 + NB + l] = Wi[l] * ui;
        }
        ci -= ((t[0] + t[1]) + (t[2] + t[3])) + (t[4] + t[5]);  // tree: the column step is a latency chain
      }
      BSEG(0);
      // ---- (b) z = A u on the panel-start triangle, rows / columns > k ----------------------------------------
      double ar[4] = {0.0, 0.0, 0.0, 0.0}, ac[4] = {0.0, 0.0, 0.0, 0.0};
      // row part left of the warp's diagonal block: columns < rw0 of row iR (permuted map: no bank conflicts);
      // whole batches of eight from the last multiple of eight at or below k+1 (u is zero up to k)
#pragma unroll 1
      for (int c = (k + 1) & ~7; c < rw0; c += 8) {
        double l[8];
        double2 u2[4];
#pragma unroll
        for (int q = 0; q < 8; ++q) l[q] = LR[c + q];
#pragma unroll
        for (int q = 0; q < 4; ++q) u2[q] = *reinterpret_cast<const double2*>(uu + c + 2 * q);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ar[q & 1] = fma(l[2 * q], u2[q].x, ar[q & 1]);
          ar[2 + (q & 1)] = fma(l[2 * q + 1], u2[q].y, ar[2 + (q & 1)]);
        }
      }
      switch (wid) {  // the warp's diagonal block (offsets are immediates)
        case 0: symv_diag32<0>(L, uu, Ti, ii, lane, n, ac); break;
        case 1: symv_diag32<32>(L, uu, Ti, ii, lane, n, ac); break;
        case 2: symv_diag32<64>(L, uu, Ti, ii, lane, n, ac); break;
        case 3: symv_diag32<96>(L, uu, Ti, ii, lane, n, ac); break;
        default: symv_diag32<128>(L, uu, Ti, ii, lane, n, ac); break;
      }
      {  // column part below the warp's block: rows >= rw0+32 of column i (contiguous across the lanes); rows
         // past n-1 of the last batch read the W panel behind the triangle, times u = 0
        const double* p = L + tri0(rw0 + 32) + ii;
#pragma unroll 1
        for (int r = rw0 + 32; r < n; r += 8) {
          double l[8];
          double2 u2[4];
          l[0] = p[0];
          l[1] = p[r + 1];
          l[2] = p[2 * r + 3];
          l[3] = p[3 * r + 6];
          l[4] = p[4 * r + 10];
          l[5] = p[5 * r + 15];
          l[6] = p[6 * r + 21];
          l[7] = p[7 * r + 28];
#pragma unroll
          for (int q = 0; q < 4; ++q) u2[q] = *reinterpret_cast<const double2*>(uu + r + 2 * q);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ac[q & 1] = fma(l[2 * q], u2[q].x, ac[q & 1]);
            ac[2 + (q & 1)] = fma(l[2 * q + 1], u2[q].y, ac[2 + (q & 1)]);
          }
          p += 8 * r + 36;
        }
      }
      if (iR < n) zrow[iR] = (ar[0] + ar[1]) + (ar[2] + ar[3]);
      __syncwarp();
      zi = ((ac[0] + ac[1]) + (ac[2] + ac[3])) + zrow[ii];
      BSEG(1);
      // ---- (c) every scalar product of the step in ONE reduction -----------------------------------------------
      const bool in2 = act && i >= k + 2;
      rd[0] = in2 ? zi * ui : 0.0;   // S1' = sum_{i >= k+2} z_i u_i
      rd[1] = in2 ? ci * ui : 0.0;   // S2  = sum c_i u_i
      rd[2] = in2 ? ui * gi : 0.0;   // S3  = sum u_i (Q^T g)_i
      rd[3] = in2 ? ui * ui : 0.0;   // ||u(k+2:)||^2
      if (i == k + 1) {
        pub[0] = zi;
        pub[1] = ci;
      }
    }
    block_sum16<NW>(rd, red, tot, parity, lane, wid, wact);
    BSEG(2);
    // ---- (d) Householder scalars, w, v, the raw next column -------------------------------------------------------
    if (wact) {
      double Vk[NB], Wk[NB];  // reloaded (uniform loads) rather than kept live across the symv
      {
        const double* wk = Wp + (k + 1) * WS;
        const double* vk = L + tri0(k + 1) + k0;
#pragma unroll
        for (int l = 0; l < NB; l += 2) {
          const double2 y = *reinterpret_cast<const double2*>(wk + l);
          Wk[l] = l < jj ? y.x : 0.0;
          Wk[l + 1] = l + 1 < jj ? y.y : 0.0;
        }
#pragma unroll
        for (int l = 0; l < NB; ++l) Vk[l] = l < jj ? vk[l] : 0.0;
      }
      const double zk1 = pub[0], ck1 = pub[1];
      const double xn2 = rd[3];
      const bool refl = xn2 > 0.0;
      const double nrm = sqrt(fma(alpha, alpha, xn2));
      const double beta = refl ? -copysign(nrm, alpha) : alpha;
      const double tk = refl ? (beta - alpha) * fast_rcp(beta) : 0.0;
      const double scal = refl ? fast_rcp(alpha - beta) : 0.0;
      const double ca = 1.0 - scal * alpha;
      // dlatrd's correction: (A - V W^T - W V^T) u = z - V (W^T u) - W (V^T u); tree sums (latency)
      double tq[NB], tk1[NB], ts1[NB];
#pragma unroll
      for (int l = 0; l < NB; ++l) {
        const double vtu = rd[4 + l], wtu = rd[4 + NB + l];
        tk1[l] = fma(Vk[l], wtu, Wk[l] * vtu);
        tq[l] = fma(Vi[l], wtu, Wi[l] * vtu);
        ts1[l] = fma(wtu, vtu - alpha * Vk[l], vtu * (wtu - alpha * Wk[l]));
      }
      const double qk1 = zk1 - (((tk1[0] + tk1[1]) + (tk1[2] + tk1[3])) + (tk1[4] + tk1[5]));
      const double qi = zi - (((tq[0] + tq[1]) + (tq[2] + tq[3])) + (tq[4] + tq[5]));
      const double S1 = rd[0] - (((ts1[0] + ts1[1]) + (ts1[2] + ts1[3])) + (ts1[4] + ts1[5]));
      const double p0 = tk * fma(scal, qk1, ca * ck1);
      const double pv = p0 + tk * scal * fma(scal, S1, ca * rd[1]);
      const double vg = fma(scal, rd[2], gq0);
      const double alpha2 = -0.5 * tk * pv;
      const double w0 = p0 + alpha2;  // w_{k+1}
      const bool first = i == k + 1;
      const double vi = act ? (first ? 1.0 : ui * scal) : 0.0;
      const double pi = tk * fma(scal, qi, ca * ci);
      const double wi = act ? fma(alpha2, vi, pi) : 0.0;
#pragma unroll
      for (int l = 0; l < NB; ++l)
        if (l == jj) {
          Vi[l] = vi;
          Wi[l] = wi;
        }
      if (act) {
        const double un = ci - fma(vi, w0, wi);  // raw column k+1 after reflector k (d_{k+1} for the first row)
        Wp[i * WS + jj] = wi;
        L[Ti + k] = vi;
        Vh[(size_t)k * n + i] = vi;
        gq[i] = fma(-tk * vg, vi, gi);
        uu[i] = first ? 0.0 : un;
        if (first) {
          a.ee[(size_t)b * n + k] = beta;
          a.tau[(size_t)b * n + k] = tk;
          a.dd[(size_t)b * n + k + 1] = un;
        }
      }
    }
    __syncthreads();
    BSEG(3);
    if (jj == NB - 1 && k + 1 < n - 2) {  // panel complete: rank-12 update of the trailing triangle on the tensor cores
      trailing_update_dmma<NW>(L, Wp, n, k0, k + 1, lane, wid);
      k0 = k + 1;
#pragma unroll
      for (int l = 0; l < NB; ++l) Vi[l] = Wi[l] = 0.0;
      __syncthreads();
      BSEG(4);
    }
  }
  // e_{n-2} is the raw column n-2; the last diagonal element still lacks the reflectors of the open panel
  if (tid == 0) {
    const int i0 = n - 2, i1 = n - 1, jj = i0 - k0;
    double dl = L[tri0(i1) + i1];
    for (int l = 0; l < jj; ++l) dl -= 2.0 * L[tri0(i1) + k0 + l] * Wp[i1 * WS + l];
    a.ee[(size_t)b * n + i0] = uu[i1];
    a.tau[(size_t)b * n + i0] = 0.0;
    a.dd[(size_t)b * n + i1] = dl;
    a.ee[(size_t)b * n + i1] = 0.0;
    a.tau[(size_t)b * n + i1] = 0.0;
  }
  for (int i2 = tid; i2 < n; i2 += THREADS) a.gq[(size_t)b * n + i2] = gq[i2];
  if (DBG && a.dbg && (tid == 0 || tid == 96))
    for (int q = 0; q < 5; ++q) a.dbg[(size_t)b * 16 + (tid == 0 ? 0 : 8) + q] = seg[q];
#undef BSEG
}

}  // namespace mop

static long long* g_tb_dbg = nullptr;
extern "C" int mop_priv_tridiag_blk_timing(void* buf) {
  g_tb_dbg = (long long*)buf;
  return MOP_OK;
}

int mop_tridiag_blk_supported(int n) { return n >= 1 && n <= 160; }

template <int NW, bool DBG>
static int launch_blk(int B, const mop::PkArgs& a, cudaStream_t stream) {
  const size_t smem = sizeof(double) * mop::tb_smem_doubles(a.n, NW);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_blk<NW, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_tridiag_blk<NW, DBG><<<B, 32 * NW, smem, stream>>>(a);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// d, e, tau, gq: [B][n]; Vh: [B][n][n]; flag: [B]
int mop_launch_tridiag_blk(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                           double* tau, double* gq, int* flag, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_blk_supported(n)) {
    mop_set_error("blocked tridiagonalisation: n = %d not supported (max 160)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  mop::PkArgs a{n, A, gp, Vh, dd, ee, tau, gq, flag, g_tb_dbg};
  const int nw = (n + 31) / 32;
  switch (nw) {
    case 1: return launch_blk<1, false>(B, a, stream);
    case 2: return launch_blk<2, false>(B, a, stream);
    case 3: return launch_blk<3, false>(B, a, stream);
    case 4: return launch_blk<4, false>(B, a, stream);
    default: return a.dbg ? launch_blk<5, true>(B, a, stream) : launch_blk<5, false>(B, a, stream);
  }
}
