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
#include "dmma.cuh"
#include "trrot.cuh"
#include "update_coef.cuh"

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
  // Staged reduction (see mop_launch_front_tridiag_blk): this launch reduces the trailing block of rows / columns
  // row0 .. nfull-1 (n = nfull - row0 is the LOCAL size; d, e, tau, gq, Vh are indexed in the full matrix), starting
  // from the state a previous stage left in `hin` (null: from A / the fused front end) and - when `hout` is set -
  // stopping at the panel boundary `kstop` (local column, a multiple of TB_NB), where it leaves the trailing triangle,
  // the raw next column and Q^T g of the rows >= kstop in `hout` for the next, smaller and more densely resident stage.
  int nfull, row0, kstop;
  const double* hin;   // [B][hstride]: packed triangle (n (n + 1) / 2, rounded up to even) | uu [n] | gq [n]
  double* hout;
  size_t hstride;
};

// Fused front end (FUSED kernels): the Hessian update, its write-back and the TR/ROT projection run on the triangle in
// shared memory before the reduction starts, so H is read once and written once and the projected Hessian never
// exists in HBM (RSIRFO.run steps 1-2: Optimizer/rsirfo.py:308-358, update_hessian :1316-1372,
// Utils/calc_tools.py:249-316).
struct FrontArgs {
  double* H;             // [B][n][n] in / out (written only when the update is applied)
  const double* Hbias;   // [B][n][n] or null
  const double* x;       // [B][n]
  const double* xp;      // [B][n] previous geometry or null
  const double* g;       // [B][n] gradient of the update (raw for RSIRFO)
  const double* gprev;   // [B][n] or null
  const double* Bg;      // [B][n] gradient to project
  const double* state;   // [B][state_stride] or null (MOP_RS_HAVE_PREV)
  int state_stride;
  int method, guards, grad_rule;
  const int32_t* method_per;  // [B] per-structure update method (overrides `method`) or null: NEB chains mix FSB / Bofill
  int packed;            // != 0: H and Hbias are packed lower triangles [B][n (n + 1) / 2] (row i at i (i + 1) / 2)
  double* gp_out;        // [B][n] projected gradient
  int32_t* status;       // [B]
};

constexpr int TB_STAGE_MIN = 24;  // staged reduction: a stage is only split off while at least this many rows remain behind it
constexpr int TB_NB = 6;   // reflectors per panel: 4 scalars + 2 * NB panel products = the 16 slots of one reduction
constexpr int TB_WS = 10;  // doubles per row of the W panel: 80-byte rows, eight 128-bit row reads hit 32 banks once

__device__ __forceinline__ int tri0(int i) { return (i * (i + 1)) >> 1; }

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


// One half-pass over a row-major n x n matrix in global memory: rows are dealt to the warps four at a time (twenty
// loads in flight per lane), LOWER = the row segments j <= i, otherwise j > i.  Element (i, j) adds w x to the packed
// triangle at (max, min) (w = 1/2 off the diagonal, so two half-passes leave sym(M) there; INIT stores instead of
// adding) and x v_j to the row sums acc0 / acc1 (M s and M y of the update), which go to out0 / out1.
template <int NW, bool LOWER, bool INIT, int NVEC>
__device__ __forceinline__ void front_half_pass(const double* __restrict__ M, int n, double* L, const double* v0,
                                                const double* v1, double* out0, double* out1, int lane, int wid) {
  for (int i0 = 4 * wid; i0 < n; i0 += 4 * NW) {
    double x[4][5];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int j = lane + 32 * q;
        const bool in = i < n && j < n && (LOWER ? j <= i : j > i);
        x[r][q] = in ? M[(size_t)i * n + j] : 0.0;
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int j = lane + 32 * q;
        const bool in = i < n && j < n && (LOWER ? j <= i : j > i);
        if (in) {
          const double w = (j == i) ? x[r][q] : 0.5 * x[r][q];
          double* dst = LOWER ? L + tri0(i) + j : L + tri0(j) + i;
          if (INIT) *dst = w;
          else *dst += w;
          if (NVEC > 0) a0 = fma(x[r][q], v0[j], a0);
          if (NVEC > 1) a1 = fma(x[r][q], v1[j], a1);
        }
      }
      if (NVEC > 0) {
        a0 = warp_sum(a0);
        if (NVEC > 1) a1 = warp_sum(a1);
        if (lane == 0 && i < n) {
          if (LOWER) {
            out0[i] = a0;
            if (NVEC > 1) out1[i] = a1;
          } else {
            out0[i] += a0;
            if (NVEC > 1) out1[i] += a1;
          }
        }
      }
    }
  }
}

// Rank-2k update of the whole triangle with the projection vectors, Hp = S - Y T^T - T Y^T, as C + (-P) Q^T with
// P = [Y | T], Q = [T | Y] (K = 12; vectors k .. 5 are zero): the same 8 x 8 DMMA tiles as the trailing updates.
// T, Y: [6][np].
template <int NW>
__device__ __forceinline__ void project_update_dmma(double* L, const double* T, const double* Y, int n, int np,
                                                    int lane, int wid) {
  const int g = lane >> 2, t = lane & 3;
  const int mt = (n + 7) >> 3;
  for (int I = mt - 1 - wid; I >= 0; I -= NW) {
    const int ri = 8 * I + g;
    const int ric = ri < n ? ri : n - 1;
    const int Tr = tri0(ric);
    double a[3];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      const int c = 4 * ks + t;
      const double pa = c < 6 ? Y[c * np + ric] : T[(c - 6) * np + ric];
      a[ks] = ri < n ? -pa : 0.0;
    }
    for (int J0 = 0; J0 <= I; J0 += 4) {
      double bq[4][3], c0[4], c1[4];
      bool in0[4], in1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int J = J0 + u;
        const int rj = 8 * J + g;
        const bool jin = J <= I && rj < n;
        const int rjc = jin ? rj : n - 1;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const int c = 4 * ks + t;
          const double qb = c < 6 ? T[c * np + rjc] : Y[(c - 6) * np + rjc];
          bq[u][ks] = jin ? qb : 0.0;
        }
        const int cj = 8 * J + 2 * t;
        in0[u] = J <= I && ri < n && cj <= ri;
        in1[u] = J <= I && ri < n && cj + 1 <= ri;
        c0[u] = in0[u] ? L[Tr + cj] : 0.0;
        c1[u] = in1[u] ? L[Tr + cj + 1] : 0.0;
      }
#pragma unroll
      for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (J0 + u <= I) dmma884(c0[u], c1[u], a[ks], bq[u][ks], c0[u], c1[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cj = 8 * (J0 + u) + 2 * t;
        if (in0[u]) L[Tr + cj] = c0[u];
        if (in1[u]) L[Tr + cj + 1] = c1[u];
      }
    }
  }
}

// ---- TMA bulk copies of the packed triangle (cp.async.bulk, SASS UBLKCP): HBM and shared memory hold the same layout, so
// the 90.6 KB of a structure (n = 150) move as ONE asynchronous copy each way instead of 71 load / store rounds per
// thread.  The copy engine needs 16-byte aligned addresses and sizes; n (n + 1) / 2 is odd for n = 1, 2 (mod 4), so every
// other structure starts on an 8-byte boundary: the kernel then places the triangle one double into its (even-sized)
// shared-memory slot, which gives source and destination the same phase, and the one element in front of the first
// 16-byte boundary / behind the last travels by a plain load or store.
__device__ __forceinline__ unsigned tb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tb_bulk_load(double* dst, const double* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tb_smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tb_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(tb_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tb_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = tb_smem_addr(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tb_bulk_store(double* dst, const double* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tb_smem_addr(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the source may be overwritten from here on
}
// split of a packed triangle of ntri doubles at address p into [head | 16-byte aligned body of an even count | tail]
__device__ __forceinline__ void tb_bulk_split(const void* p, int ntri, int* head, int* body) {
  *head = (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1);
  *body = (ntri - *head) & ~1;
}

// The fused front end.  On return L holds the projected effective Hessian sym(P^T (H' + Hbias) P) (H' = the updated
// Hessian, already written back), *gp_mine the projected gradient entry of row tid.  `fr` is the region behind the
// triangle (free until the reduction starts; left dirty).  Whole CTA.
template <int NW>
__device__ void fused_front(const FrontArgs& f, int n, int np, int b, double* L, double* fr, double* gp_mine,
                            long long* dbg) {
  constexpr int THREADS = 32 * NW;
  // optional phase clocks (diagnostics, tools/tb_segments.py front): dbg [16] per structure
  long long fmark = dbg ? clock64() : 0;
  int fslot = 0;
#define FMARK()                                   \
  do {                                            \
    if (dbg && threadIdx.x == 0) {                \
      const long long tn_ = clock64();            \
      dbg[fslot++] = tn_ - fmark;                 \
      fmark = tn_;                                \
    }                                             \
  } while (0)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const size_t nn = (size_t)n * n;
  double* H = f.H + (size_t)b * nn;   // (full-square layout; the packed layout is addressed where it is used)
  double* vs = fr;            // s
  double* vy = vs + np;       // y (after damping)
  double* vu = vy + np;       // u = H s
  double* vr = vu + np;       // r = y - u
  double* hy = vr + np;       // H y (flowchart)         -- later Tm rows 0 .. 3 live in hy .. hy + 4 np
  double* scratch = fr + 12 * np;  // 48 doubles
  __shared__ UpdCoef coef;
  int st = f.status ? f.status[b] : 0;
  st &= ~(MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV | MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY |
          MOP_ST_TRROT_RANKDEF);

  // ---- s, y, guards, damping (RSIRFO.update_hessian, rsirfo.py:1316-1340) ----------------------------------------
  const int method_b = f.method_per ? f.method_per[b] : f.method;
  const bool asked = method_b != MOP_UPD_NONE && f.xp != nullptr && f.gprev != nullptr;
  const bool have_prev = asked && (f.state == nullptr || f.state[(size_t)b * f.state_stride + MOP_RS_HAVE_PREV] != 0.0);
  bool upd = have_prev;
  int m = method_b;
  double ss = 0.0, sy = 0.0, yy = 0.0;
  if (!have_prev) {
    if (asked) st |= MOP_ST_NO_HISTORY;
  } else {
    for (int i = tid; i < n; i += THREADS) {
      vs[i] = f.x[(size_t)b * n + i] - f.xp[(size_t)b * n + i];
      vy[i] = f.g[(size_t)b * n + i] - f.gprev[(size_t)b * n + i];
    }
    __syncthreads();
    double pss = 0, psy = 0, pyy = 0;
    for (int i = tid; i < n; i += THREADS) {
      pss = fma(vs[i], vs[i], pss);
      psy = fma(vs[i], vy[i], psy);
      pyy = fma(vy[i], vy[i], pyy);
    }
    ss = block_sum(pss, scratch);
    sy = block_sum(psy, scratch);
    yy = block_sum(pyy, scratch);
    if (f.guards) {
      int skip = 0;
      if (sqrt(ss) < 1e-10 || sqrt(yy) < 1e-10) skip = MOP_ST_UPD_SKIP_SMALL;
      else if (f.guards == 1 && sy <= 0.0) skip = MOP_ST_UPD_SKIP_CURV;
      if (skip) {
        st |= skip;
        upd = false;
      }
    }
    if (upd && method_has_dd(m)) {
      bool active = true;
      if (m == MOP_UPD_BLOCK_BFGS_DD && !(sqrt(ss) > 1e-8)) active = false;
      if (active) {
        const double th = dd_theta(ss, sy, method_dd_thr(m));
        if (th != 1.0) {
          for (int i = tid; i < n; i += THREADS) vy[i] = th * vy[i] + (1.0 - th) * vs[i];
          __syncthreads();
          double p = 0;
          for (int i = tid; i < n; i += THREADS) p = fma(vs[i], vy[i], p);
          sy = block_sum(p, scratch);
        }
      }
    }
  }
  __syncthreads();
  FMARK();  // 0: s, y, guards

  // ---- one read of H: sym(H) into the triangle, u = H s (and H y) from the same loads ----------------------------
  const int ntri = (n * (n + 1)) >> 1;
  if (f.packed) {
    // packed lower triangle in HBM = the shared-memory layout: a straight coalesced copy, then u = H s as a
    // thread-per-row symv on the triangle
    const double* Hpk = f.H + (size_t)b * ntri;
    __shared__ __align__(8) unsigned long long s_mbar;
    const bool bulk = ((reinterpret_cast<uintptr_t>(Hpk) ^ reinterpret_cast<uintptr_t>(L)) & 15) == 0;   // same 16-byte phase
    if (bulk) {
      int head, body;
      tb_bulk_split(Hpk, ntri, &head, &body);
      if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tb_smem_addr(&s_mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();
      if (tid == 0) tb_bulk_load(L + head, Hpk + head, (unsigned)body * 8u, &s_mbar);
      if (tid == 32 % THREADS && head) L[0] = Hpk[0];
      if (tid == 64 % THREADS && head + body < ntri) L[ntri - 1] = Hpk[ntri - 1];
      tb_mbar_wait(&s_mbar, 0);
    } else {
      for (int e = tid; e < ntri; e += THREADS) L[e] = Hpk[e];
    }
    __syncthreads();
    if (upd && tid < n) {
      const int i = tid;
      const bool two = m == MOP_UPD_FLOWCHART;
      double a0 = 0.0, a1 = 0.0;
      const double* Li = L + tri0(i);
      for (int j = 0; j <= i; ++j) {
        const double x = Li[j];
        a0 = fma(x, vs[j], a0);
        if (two) a1 = fma(x, vy[j], a1);
      }
      const double* p = L + tri0(i + 1) + i;
      for (int r = i + 1; r < n; ++r) {
        const double x = p[0];
        a0 = fma(x, vs[r], a0);
        if (two) a1 = fma(x, vy[r], a1);
        p += r + 1;
      }
      vu[i] = a0;
      if (two) hy[i] = a1;
    }
  } else if (upd) {
    if (m == MOP_UPD_FLOWCHART) {
      front_half_pass<NW, true, true, 2>(H, n, L, vs, vy, vu, hy, lane, wid);
      __syncthreads();
      front_half_pass<NW, false, false, 2>(H, n, L, vs, vy, vu, hy, lane, wid);
    } else {
      front_half_pass<NW, true, true, 1>(H, n, L, vs, vy, vu, hy, lane, wid);
      __syncthreads();
      front_half_pass<NW, false, false, 1>(H, n, L, vs, vy, vu, hy, lane, wid);
    }
  } else {
    front_half_pass<NW, true, true, 0>(H, n, L, vs, vy, vu, hy, lane, wid);
    __syncthreads();
    front_half_pass<NW, false, false, 0>(H, n, L, vs, vy, vu, hy, lane, wid);
  }
  __syncthreads();
  FMARK();  // 1: read of H

  if (upd) {
    // ---- scalars, coefficient matrix (hessian_update.py / block_hessian_update.py through update_coef.cuh) -------
    if (m == MOP_UPD_FLOWCHART) {
      double pzz = 0, pzs = 0;
      for (int i = tid; i < n; i += THREADS) {
        const double z = vy[i] - hy[i];
        pzz = fma(z, z, pzz);
        pzs = fma(z, vs[i], pzs);
      }
      const double zz = block_sum(pzz, scratch);
      const double zs = block_sum(pzs, scratch);
      m = flowchart_select(ss, yy, sy, zz, zs);
    }
    double psu = 0, prs = 0, prr = 0;
    for (int i = tid; i < n; i += THREADS) {
      const double r = vy[i] - vu[i];
      vr[i] = r;
      psu = fma(vs[i], vu[i], psu);
      prs = fma(r, vs[i], prs);
      prr = fma(r, r, prr);
    }
    UpdScalars q;
    q.ss = ss;
    q.sy = sy;
    q.su = block_sum(psu, scratch);
    q.rs = block_sum(prs, scratch);
    q.rr = block_sum(prr, scratch);
    if (tid == 0) update_coefficients(m, q, coef);
    __syncthreads();
    st |= MOP_ST_UPDATED | coef.flags;
    // Tm[a][j] = sum_b C[a][b] v_b[j]: delta_ij = sum_a v_a[i] Tm[a][j]  (the fma order of coef_delta)
    double* Tm = hy;  // 4 np (H y is dead)
    for (int j = tid; j < n; j += THREADS) {
      const double vj[4] = {vs[j], vy[j], vu[j], vr[j]};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) t = fma(coef.c[a][c], vj[c], t);
        Tm[a * np + j] = t;
      }
    }
    __syncthreads();
    // ---- H' = sym(H) + 1/2 (delta + delta^T) on the triangle: thread per row ------------------------------------
    if (tid < n) {
      const int i = tid;
      const double vi[4] = {vs[i], vy[i], vu[i], vr[i]};
      const double ti[4] = {Tm[i], Tm[np + i], Tm[2 * np + i], Tm[3 * np + i]};
      double* Li = L + tri0(i);
      for (int j = 0; j <= i; ++j) {
        double d0 = 0.0, d1 = 0.0;
        const double vj[4] = {vs[j], vy[j], vu[j], vr[j]};
        const double tj[4] = {Tm[j], Tm[np + j], Tm[2 * np + j], Tm[3 * np + j]};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          d0 = fma(vi[a], tj[a], d0);
          d1 = fma(vj[a], ti[a], d1);
        }
        Li[j] += 0.5 * (d0 + d1);
      }
    }
    __syncthreads();
    FMARK();  // 2: coefficients + update on the triangle
    // ---- write H' back (the only write of the Hessian): full rows from the triangle, or the triangle itself ------
    if (f.packed) {
      double* Hpk = f.H + (size_t)b * ntri;
      if (((reinterpret_cast<uintptr_t>(Hpk) ^ reinterpret_cast<uintptr_t>(L)) & 15) == 0) {
        int head, body;
        tb_bulk_split(Hpk, ntri, &head, &body);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the update above -> visible to the copy engine
        __syncthreads();
        if (tid == 0) tb_bulk_store(Hpk + head, L + head, (unsigned)body * 8u);
        if (tid == 32 % THREADS && head) Hpk[0] = L[0];
        if (tid == 64 % THREADS && head + body < ntri) Hpk[ntri - 1] = L[ntri - 1];
      } else {
        for (int e = tid; e < ntri; e += THREADS) Hpk[e] = L[e];
      }
    } else {
      for (int i = wid; i < n; i += NW) {
        double* row = H + (size_t)i * n;
        for (int j = lane; j < n; j += 32) row[j] = (j <= i) ? L[tri0(i) + j] : L[tri0(j) + i];
      }
    }
  }
  // ---- effective Hessian: + sym(Hbias) (rsirfo.py:349-353) -----------------------------------------------------------
  if (f.Hbias && f.packed) {
    const double* Hb = f.Hbias + (size_t)b * ntri;
    __syncthreads();
    for (int e = tid; e < ntri; e += THREADS) L[e] += Hb[e];
  } else if (f.Hbias) {
    const double* Hb = f.Hbias + (size_t)b * nn;
    __syncthreads();
    front_half_pass<NW, true, false, 0>(Hb, n, L, vs, vy, vu, hy, lane, wid);
    __syncthreads();
    front_half_pass<NW, false, false, 0>(Hb, n, L, vs, vy, vu, hy, lane, wid);
  }
  __syncthreads();

  FMARK();  // 3: write-back (+ bias)
  // ---- TR/ROT basis (classical Gram-Schmidt with drop, calc_tools.py:250-259), projected gradient -----------------
  double* T = fr;             // [6][np]
  double* Y = fr + 6 * np;    // [6][np]: raw vectors, then W = S T, then Y
  const int k = build_trrot_basis(n, f.x + (size_t)b * n, T, np, Y, scratch);
  if (k < 6) st |= MOP_ST_TRROT_RANKDEF;
  {
    const double* g = f.Bg + (size_t)b * n;
    double* gpo = f.gp_out + (size_t)b * n;
    if (k < 6) {  // block-uniform: the reference's Householder Q differs from the Gram-Schmidt span here
      project_grad_qr(n, Y, np, g, gpo, f.grad_rule, scratch);
      __syncthreads();
      *gp_mine = tid < n ? gpo[tid] : 0.0;
    } else {
      double cf[6];
      for (int j = 0; j < 6; ++j) {
        double p = 0.0;
        for (int i = tid; i < n; i += THREADS) p = fma(T[j * np + i], g[i], p);
        cf[j] = block_sum(p, scratch);
      }
      double mine = 0.0;
      for (int i = tid; i < n; i += THREADS) {
        double part = 0.0;
        for (int j = 0; j < 6; ++j) part = fma(T[j * np + i], cf[j], part);
        const double v = g[i] - part;
        gpo[i] = v;
        if (i == tid) mine = v;
      }
      *gp_mine = mine;
    }
  }
  __syncthreads();
  for (int e = tid; e < 6 * np; e += THREADS) {
    if (e / np >= k) T[e] = 0.0;  // unused basis slots (rejected candidates leave their residual there)
    Y[e] = 0.0;
  }
  __syncthreads();
  FMARK();  // 4: basis + projected gradient
  // ---- W = S T: thread per row on the triangle (row part + column part), six vectors at once -------------------------
  if (tid < n) {
    const int i = tid;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double* Li = L + tri0(i);
    for (int j = 0; j <= i; ++j) {
      const double x = Li[j];
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[v] = fma(x, T[v * np + j], acc[v]);
    }
    const double* p = L + tri0(i + 1) + i;
    for (int r = i + 1; r < n; ++r) {
      const double x = p[0];
#pragma unroll
      for (int v = 0; v < 6; ++v) acc[v] = fma(x, T[v * np + r], acc[v]);
      p += r + 1;
    }
#pragma unroll
    for (int v = 0; v < 6; ++v)
      if (v < k) Y[v * np + i] = acc[v];
  }
  __syncthreads();
  // M = T^T W (k x k), Y = W - 1/2 T sym(M)
  double* M = scratch;  // 36
  for (int e = wid; e < k * k; e += NW) {
    const int a = e / k, c = e - a * k;
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = fma(T[a * np + i], Y[c * np + i], p);
    p = warp_sum(p);
    if (lane == 0) M[a * 6 + c] = p;
  }
  __syncthreads();
  for (int i = tid; i < n; i += THREADS) {
    double tv[6];
    for (int a = 0; a < k; ++a) tv[a] = T[a * np + i];
    for (int c = 0; c < k; ++c) {
      double corr = 0.0;
      for (int a = 0; a < k; ++a) corr = fma(tv[a], 0.5 * (M[a * 6 + c] + M[c * 6 + a]), corr);
      Y[c * np + i] -= 0.5 * corr;
    }
  }
  __syncthreads();
  FMARK();  // 5: W = S T, M, Y
  // ---- Hp = S - Y T^T - T Y^T: rank-12 update on the tensor cores ---------------------------------------------------
  project_update_dmma<NW>(L, T, Y, n, np, lane, wid);
  __syncthreads();
  FMARK();  // 6: rank-12 projection update
#undef FMARK
  if (tid == 0 && f.status) f.status[b] = st;
  __syncthreads();
}

template <int NW>
struct TbMinBlocks {
  static constexpr int value = NW >= 5 ? 2 : (NW == 4 ? 3 : (NW == 3 ? 5 : (NW == 2 ? 8 : 16)));
};

// shared memory: L | Wp | uu | zrow | gq | red | tot | pub | block_sum_k scratch (doubles)
__host__ __device__ inline size_t tb_smem_doubles(int n, int nw) {
  const size_t np = (size_t)((n + 3) & ~3);
  const size_t nl = ((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1;
  size_t fr = (size_t)n * TB_WS + (32 * (size_t)nw + 16) + 2 * np + (size_t)red16_doubles(nw) + 16 * (size_t)nw + 4 + 64;
  const size_t front = 12 * np + 48;  // fused front end: T | Y (or s, y, u, r, H y, coefficient rows) + reduction scratch
  if (fr < front) fr = front;
  return nl + fr;
}

template <int NW, bool DBG, bool FUSED>
__global__ void __launch_bounds__(32 * NW, TbMinBlocks<NW>::value) k_tridiag_blk(PkArgs a, FrontArgs f) {
  constexpr int THREADS = 32 * NW;
  constexpr int NB = TB_NB, WS = TB_WS;
  constexpr int NU = 32 * NW + 16;  // uu is zero outside (k, n): the symv loops run in whole batches of eight
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  const size_t nl = ((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1;
  double* L = sm;                   // packed lower triangle, row i at i (i + 1) / 2
  if (FUSED && f.packed && ((n * (n + 1) / 2) & 1) &&
      (reinterpret_cast<uintptr_t>(f.H + (size_t)b * (n * (n + 1) / 2)) & 15))
    L = sm + 1;                     // odd triangle on an 8-byte boundary: same 16-byte phase as its HBM image (TMA bulk copies)
  double* Wp = sm + nl;             // [n][WS] W panel (directly behind L: symv reads past row n-1 land here, times 0)
  double* uu = Wp + (size_t)n * WS; // [NU] raw updated column (16-byte aligned: nl and n WS are even)
  double* zrow = uu + NU;           // [np] row-part sums of the symv (permuted lane map)
  double* gq = zrow + np;           // [np] Q^T g
  double* red = gq + np;            // [2][16][NW]
  double* tot = red + red16_doubles(NW);  // [NW][16]
  double* pub = tot + 16 * NW;      // [4]  z_{k+1}, c_{k+1}
  double* s_rb = pub + 4;           // [64] block_sum_k<2> scratch
  int parity = 0, parity2 = 0;
  const int nf = a.nfull ? a.nfull : n, r0 = a.row0;   // full size and offset of this stage in the full matrix
  const size_t ob = (size_t)b * nf + r0;               // d, e, tau, gq of local row i live at ob + i
  const bool resume = !FUSED && a.hin != nullptr;
  const double* Ain = (FUSED || resume) ? nullptr : a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * nf * nf + (size_t)r0 * nf + r0;   // reflector k in row k: Vh[k * nf + i]

  double pn[2] = {0.0, 0.0};
  double gp_mine = 0.0;
  if (FUSED) {
    // update + write-back + projection on the triangle (one read and one write of H, no projected Hessian in HBM)
    fused_front<NW>(f, n, np, b, L, Wp, &gp_mine, a.dbg ? a.dbg + (size_t)b * 16 : nullptr);
    for (int i = tid; i < (int)(tb_smem_doubles(n, NW) - nl); i += THREADS) Wp[i] = 0.0;
    for (int e = tid; e < n * (n + 1) / 2; e += THREADS) pn[0] = fma(L[e], L[e], pn[0]);
  } else if (resume) {
    // ---- continue a reduction an earlier stage handed over (its flagged structures are finished already) ----
    if (a.flag[b] != 0) return;
    for (int i = tid; i < (int)(tb_smem_doubles(n, NW) - nl); i += THREADS) Wp[i] = 0.0;
    __syncthreads();
    const double* h = a.hin + (size_t)b * a.hstride;
    if ((reinterpret_cast<uintptr_t>(h) & 15) == 0) {   // the hand-over triangle (even-sized slot, 16-byte aligned): one TMA bulk copy
      __shared__ __align__(8) unsigned long long s_hbar;
      if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tb_smem_addr(&s_hbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      __syncthreads();
      if (tid == 0) tb_bulk_load(L, h, (unsigned)nl * 8u, &s_hbar);
      for (int i = tid; i < n; i += THREADS) {
        uu[i] = h[nl + i];
        gq[i] = h[nl + n + i];
      }
      tb_mbar_wait(&s_hbar, 0);
    } else {
      for (int e = tid; e < n * (n + 1) / 2; e += THREADS) L[e] = h[e];
      for (int i = tid; i < n; i += THREADS) {
        uu[i] = h[nl + i];
        gq[i] = h[nl + n + i];
      }
    }
    __syncthreads();
  } else {
    // everything behind the triangle starts finite (the batched symv multiplies a few words of it by zero)
    for (int i = tid; i < (int)(tb_smem_doubles(n, NW) - nl); i += THREADS) Wp[i] = 0.0;
    // ---- load the lower triangle (row segments, coalesced), Frobenius norm of it for the trivial cases ----
    for (int i = wid; i < n; i += NW) {
      const double* row = Ain + (size_t)i * n;
      double* Lr = L + tri0(i);
      for (int j = lane; j <= i; j += 32) {
        const double x = row[j];
        Lr[j] = x;
        pn[0] = fma(x, x, pn[0]);
      }
    }
  }
  if (!resume) {
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) gq[i] = FUSED ? (i == tid ? gp_mine : 0.0) : (a.gp ? a.gp[(size_t)b * n + i] : 0.0);
    block_sum_k<2>(pn, s_rb, parity2);
    const double fro = sqrt(pn[0]);
    const bool nonfinite = !isfinite(fro), trivial = nonfinite || fro == 0.0;
    if (tid == 0) a.flag[b] = nonfinite ? 2 : (fro == 0.0 ? 1 : 0);
    if (trivial || n <= 2) {
      for (int i = tid; i < n; i += THREADS) {
        a.dd[ob + i] = nonfinite ? NAN : (trivial ? 0.0 : L[tri0(i) + i]);
        a.ee[ob + i] = (!trivial && i + 1 < n) ? L[tri0(i + 1) + i] : 0.0;
        a.tau[ob + i] = 0.0;
        a.gq[ob + i] = gq[i];
      }
      return;
    }
    for (int i = tid; i < n; i += THREADS) uu[i] = i > 0 ? L[tri0(i)] : 0.0;  // raw column 0
    if (tid == 0) a.dd[ob] = L[0];
    __syncthreads();
  }

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
        Vh[(size_t)k * nf + i] = vi;
        gq[i] = fma(-tk * vg, vi, gi);
        uu[i] = first ? 0.0 : un;
        if (first) {
          a.ee[ob + k] = beta;
          a.tau[ob + k] = tk;
          a.dd[ob + k + 1] = un;
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
      if (a.hout && k0 == a.kstop) {
        // ---- hand the trailing block over to the next stage: triangle of the rows / columns >= k0, raw column k0,
        // Q^T g; the finished rows of Q^T g go to their final place ----
        const int ks = k0, m = n - ks;
        const size_t nlm = ((size_t)m * (m + 1) / 2 + 1) & ~(size_t)1;
        double* h = a.hout + (size_t)b * a.hstride;
        for (int r = ks + wid; r < n; r += NW) {
          const double* src = L + tri0(r) + ks;
          double* dst = h + tri0(r - ks);
          for (int j = lane; j <= r - ks; j += 32) dst[j] = src[j];
        }
        for (int r = ks + tid; r < n; r += THREADS) {
          h[nlm + r - ks] = uu[r];
          h[nlm + m + r - ks] = gq[r];
        }
        for (int r = tid; r < ks; r += THREADS) a.gq[ob + r] = gq[r];
        return;
      }
    }
  }
  // e_{n-2} is the raw column n-2; the last diagonal element still lacks the reflectors of the open panel
  if (tid == 0) {
    const int i0 = n - 2, i1 = n - 1, jj = i0 - k0;
    double dl = L[tri0(i1) + i1];
    for (int l = 0; l < jj; ++l) dl -= 2.0 * L[tri0(i1) + k0 + l] * Wp[i1 * WS + l];
    a.ee[ob + i0] = uu[i1];
    a.tau[ob + i0] = 0.0;
    a.dd[ob + i1] = dl;
    a.ee[ob + i1] = 0.0;
    a.tau[ob + i1] = 0.0;
  }
  for (int i2 = tid; i2 < n; i2 += THREADS) a.gq[ob + i2] = gq[i2];
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

template <int NW, bool DBG, bool FUSED>
static int launch_blk(int B, const mop::PkArgs& a, const mop::FrontArgs& f, cudaStream_t stream) {
  const size_t smem = sizeof(double) * mop::tb_smem_doubles(a.n, NW);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_tridiag_blk<NW, DBG, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  mop::k_tridiag_blk<NW, DBG, FUSED><<<B, 32 * NW, smem, stream>>>(a, f);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

template <bool FUSED>
static int dispatch_blk(int B, const mop::PkArgs& a, const mop::FrontArgs& f, cudaStream_t stream) {
  switch ((a.n + 31) / 32) {
    case 1: return launch_blk<1, false, FUSED>(B, a, f, stream);
    case 2: return launch_blk<2, false, FUSED>(B, a, f, stream);
    case 3: return launch_blk<3, false, FUSED>(B, a, f, stream);
    case 4: return launch_blk<4, false, FUSED>(B, a, f, stream);
    default: return (a.dbg && !FUSED) ? launch_blk<5, true, false>(B, a, f, stream) : launch_blk<5, false, FUSED>(B, a, f, stream);
  }
}

// Staged reduction.  The column step is a latency chain whose cost barely depends on how many CTAs share the SM, and
// the trailing matrix shrinks: whenever what is left fits a denser launch (m <= 120: three CTAs per SM, <= 88: five,
// <= 64: eight, <= 32: sixteen) the reduction is handed over at the next panel boundary, through the two halves of an
// n^2-double slab per structure (`hand`, [B][n][n]; it may alias the structure's own input matrix A: a CTA reads its A
// before it writes its hand-over).  n = 150: columns 0-29 | 30-65 | 66-89 | 90-119 | 120-147.
static int tb_stage_cols(int m) {  // columns a stage of local size m reduces before handing over (0: it finishes the job)
  static const int fit[] = {120, 88, 64, 32};
  for (int t : fit)
    if (t < m) {
      const int cols = ((m - t + mop::TB_NB - 1) / mop::TB_NB) * mop::TB_NB;
      return (m - cols >= mop::TB_STAGE_MIN) ? cols : 0;
    }
  return 0;
}

template <bool FUSED>
static int staged_blk(int B, mop::PkArgs a, const mop::FrontArgs& f, double* hand, cudaStream_t stream) {
  const int n = a.n;
  // a batch that leaves SMs idle anyway (<= two CTAs per SM: a NEB chain of 64 images) gains nothing from denser stages
  // and would pay their launches and hand-overs
  if (!hand || B <= 2 * 148 || tb_stage_cols(n) == 0) return dispatch_blk<FUSED>(B, a, f, stream);
  double* buf[2] = {hand, hand + (((size_t)n * n / 2) & ~(size_t)1)};
  int c = 0, s = 0;
  for (;; ++s) {
    const int m = n - c, cols = tb_stage_cols(m);
    const bool last = cols == 0;
    a.n = m;
    a.row0 = c;
    a.kstop = cols;
    a.hin = s ? buf[(s + 1) & 1] : nullptr;
    a.hout = last ? nullptr : buf[s & 1];
    a.hstride = (size_t)n * n;
    const int rc = s ? dispatch_blk<false>(B, a, f, stream) : dispatch_blk<FUSED>(B, a, f, stream);
    if (rc != MOP_OK || last) return rc;
    c += cols;
  }
}

// Continue a reduction another kernel (the cluster tridiagonalisation, tridiag_cluster.cu) handed over: rows / columns
// row0 .. nfull-1, state in region 0 of hand [B][hstride]; regions 1 and 2 (13312 doubles each) alternate between the
// stages.  flag [B] must be zero, gq [B][nfull] is scratch (the rows >= row0 receive Q^T 0).
int mop_launch_tridiag_blk_resume(int B, int nfull, int row0, double* hand, size_t hstride, double* Vh, double* dd,
                                  double* ee, double* tau, double* gq, int* flag, cudaStream_t stream) {
  constexpr size_t REGION = 13312;
  if (B == 0) return MOP_OK;
  if (nfull - row0 > 160 || nfull - row0 < 3 || row0 % mop::TB_NB != 0) {
    mop_set_error("blocked tridiagonalisation: cannot resume %d rows at row %d", nfull - row0, row0);
    return MOP_ERR_UNSUPPORTED;
  }
  mop::PkArgs a{nfull - row0, nullptr, nullptr, Vh, dd, ee, tau, gq, flag, nullptr, nfull, row0, 0, hand, nullptr, hstride};
  mop::FrontArgs f{};
  int c = row0;
  for (int s = 0;; ++s) {
    const int m = nfull - c, cols = tb_stage_cols(m);
    const bool last = cols == 0;
    a.n = m;
    a.row0 = c;
    a.kstop = cols;
    a.hout = last ? nullptr : hand + REGION * (1 + (s & 1));
    const int rc = dispatch_blk<false>(B, a, f, stream);
    if (rc != MOP_OK || last) return rc;
    a.hin = a.hout;
    c += cols;
  }
}

// launches the staged reduction of an n x n matrix takes (1: not staged)
extern "C" int mop_tridiag_stage_count(int n) {
  int s = 1;
  for (int m = n, cols; (cols = tb_stage_cols(m)) != 0; m -= cols) ++s;
  return s;
}

// d, e, tau, gq: [B][n]; Vh: [B][n][n]; flag: [B]
// hand: [B][n][n] scratch for the staged reduction (may be A itself; null: one launch)
int mop_launch_tridiag_blk(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                           double* tau, double* gq, int* flag, double* hand, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_blk_supported(n)) {
    mop_set_error("blocked tridiagonalisation: n = %d not supported (max 160)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  mop::PkArgs a{n, A, gp, Vh, dd, ee, tau, gq, flag, g_tb_dbg, n, 0, 0, nullptr, nullptr, 0};
  mop::FrontArgs f{};
  return staged_blk<false>(B, a, f, a.dbg ? nullptr : hand, stream);
}

// Steps 1-3a of RSIRFO.run in one kernel: Hessian update (method, guards as mop_launch_hessian_update with mode 1),
// write-back of H, TR/ROT projection of gradient (-> gp_out) and effective Hessian, tridiagonalisation of the latter.
// packed != 0: H and Hbias are packed lower triangles [B][n (n + 1) / 2].  hand: [B][n][n] doubles of scratch for the
// staged reduction (null: one launch does the whole reduction).
int mop_launch_front_tridiag_blk(int B, int n, int method, const int32_t* method_per, int guards, int grad_rule,
                                 int packed, double* H, const double* Hbias,
                                 const double* x, const double* xp, const double* g, const double* gprev,
                                 const double* Bg, const double* state, int state_stride, double* gp_out,
                                 int32_t* status, double* Vh, double* dd, double* ee, double* tau, double* gq, int* flag,
                                 double* hand, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_blk_supported(n) || n < 3) {
    mop_set_error("fused update + projection + tridiagonalisation: n = %d not supported (3 .. 160)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  mop::PkArgs a{n, nullptr, nullptr, Vh, dd, ee, tau, gq, flag, g_tb_dbg, n, 0, 0, nullptr, nullptr, 0};
  mop::FrontArgs f{H, Hbias, x, xp, g, gprev, Bg, state, state_stride, method, guards, grad_rule, method_per, packed, gp_out, status};
  return staged_blk<true>(B, a, f, hand, stream);
}
