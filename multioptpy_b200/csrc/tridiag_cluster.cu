// Blocked Householder tridiagonalisation for matrices that do not fit one SM (160 < n <= 1024; BASELINE config 5:
// P-RFO at N = 200 atoms, n = 600): LAPACK dlatrd panels, FP64 tensor-core (DMMA, mma.sync.m8n8k4) rank-2NB
// trailing updates, one thread-block CLUSTER per matrix.
//
// The unblocked cluster kernel (k_lg_tridiag2, eigh_large.cu) reads AND writes the whole trailing matrix in every
// column step - 16 n^3 / 3 bytes through L2 per matrix, 1.15 GB at n = 600 - and pays four CTA barriers and a chain of
// dependent reductions around every cluster barrier.  Here
//   * a column step only READS the trailing matrix (symv on the panel-start matrix A0, dlatrd): rows are dealt to the
//     warps of the cluster, every warp pushes its row sums into the shared memory of all CTAs (DSMEM), ONE cluster
//     barrier per column; the panel of NB = 6 reflectors is applied once per panel as a rank-12 update of the full
//     trailing square, 8 x 8 tiles on the FP64 tensor cores, the tiles dealt over all warps of the cluster;
//   * everything else of the column step - the corrected columns k and k+1, the Householder scalars, dlatrd's
//     correction w -= V (W^T v) + W (V^T v), the next raw column - is O(n NB) work that EVERY CTA repeats on its own
//     copy of the panel (V, W in shared memory), bit-identically, instead of exchanging it: the only cluster traffic
//     is the symv result;
//   * as in k_tridiag_blk the symv runs on the RAW updated column u (v = s u + (1 - s alpha) e_{k+1} is linear in
//     it), so it does not wait for the Householder norm, and the norm, c.u and the 2 (NB - 1) panel products travel
//     through one 16-value block reduction while the symv loads are in flight.
// The matrix is read with ld.global.cg (L2 only): rows updated by another CTA at a panel end are never served from a
// stale L1 line.  Outputs as k_lg_tridiag2 (LAPACK dsytd2 conventions): d, e, tau, reflector k in row k of Vh with
// the unit entry explicit.  Replaces the reduction stage of numpy.linalg.eigh at Optimizer/rsprfo.py:783,798,1141
// and Optimizer/rsirfo.py:606 for large systems.
#include <cooperative_groups.h>

#include "dmma.cuh"

namespace cg = cooperative_groups;

namespace mop {

constexpr int TC_NB = 6;        // reflectors per panel (2 + 2 NB values fill the 16-slot reduction)
constexpr int TC_THREADS = 512;
constexpr int TC_NW = TC_THREADS / 32;
constexpr int TC_NPT = 2;       // rows per thread of the replicated element-wise work (n <= 1024)

// Matrix loads of the symv: ld.global.cg (L2 only).  Rows updated by another CTA at a panel end must never be served
// from a stale L1 line, and a CTA's share of the matrix does not fit L1 anyway; the intrinsic (not volatile asm)
// leaves the compiler free to hoist the next block's loads above the arithmetic of the current one.
__device__ __forceinline__ double tc_ld(const double* p) { return __ldcg(p); }

// ---- symv exchange: remote shared-memory stores that signal an mbarrier of the destination CTA ----------------------
// st.async.shared::cluster...complete_tx delivers 8 bytes into a peer's shared memory and counts them on that peer's
// mbarrier; the consumer arms the barrier with the bytes it expects for the column and spins on try_wait.  No fence,
// no cluster barrier: a column step costs one DSMEM latency instead of MEMBAR.GPU + CCTL.IVALL + barrier.cluster.
__device__ __forceinline__ unsigned tc_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = tc_smem_addr(bar);
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  }
}
// value -> peer `rank`'s copy of *dst (a shared-memory address of THIS CTA), counted on the peer's copy of *bar
__device__ __forceinline__ void tc_push(double* dst, unsigned long long* bar, unsigned rank, double value) {
  unsigned rd, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(tc_smem_addr(dst)), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(tc_smem_addr(bar)), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rd),
               "l"(__double_as_longlong(value)), "r"(rb)
               : "memory");
}

struct TcArgs {
  int n;
  double* A;    // [B][n][n] symmetric working copy (destroyed)
  double* Vh;   // [B][n][n] reflector k in row k, columns k+1.. (unit entry written)
  double* dd;   // [B][n]
  double* ee;   // [B][n]
  double* tau;  // [B][n]
  long long* dbg;  // optional [B][8] phase cycles (diagnostics)
  int ablate;      // diagnostics: bit0 no matrix loads, bit1 no column sums, bit2 no butterfly (results invalid)
  // Hand-over to the shared-memory kernel: at the panel boundary kstop (n - kstop <= 160) the trailing lower triangle
  // and the raw next column go to hout [B][hstride] in the layout k_tridiag_blk resumes from (tridiag_blocked.cu:
  // packed triangle, rounded up to even | uu [m] | Q^T g [m] = 0) and the cluster exits.  0 / null: reduce to the end.
  int kstop;
  double* hout;
  size_t hstride;
};

// sym_cl > 0: the symmetric variant with sym_cl CTAs per cluster (z slots per source CTA, row sums, per-warp column sums)
__host__ __device__ inline size_t tc_smem_doubles(int n, int sym_cl) {
  const size_t np = (size_t)((n + 3) & ~3);
  const size_t z = sym_cl > 0 ? 2 * (size_t)sym_cl + 2 + TC_NW : 2;
  return (2 * TC_NB + 1 + z) * np + red16_doubles(TC_NW) + 16 * TC_NW + 8 + 64;
}

// z_r = sum_{j > k} A0[r][j] u_j for RL rows of one warp (first, first + rs, ...), pushed into every CTA's z.
// RL * UNR = 24 independent loads per lane are in flight.
template <int RL>
__device__ __forceinline__ void tc_symv_rows(int CL, const double* A, int n, int k, int first, int W, int lane,
                                             const double* __restrict__ uu, double* zdst, unsigned long long* bar) {
  constexpr int UNR = 24 / RL;
  double acc[RL];
#pragma unroll
  for (int q = 0; q < RL; ++q) acc[q] = 0.0;
  const double* base = A + (size_t)first * n;
  const size_t rs = (size_t)W * n;
#pragma unroll 1
  for (int j0 = (k + 1) & ~31; j0 < n; j0 += 32 * UNR) {
    double x[RL][UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + 32 * u + lane;
#pragma unroll
      for (int q = 0; q < RL; ++q) x[q][u] = j < n ? tc_ld(base + (size_t)q * rs + j) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = j0 + 32 * u + lane;
      const double uj = j < n ? uu[j] : 0.0;  // zero up to k
#pragma unroll
      for (int q = 0; q < RL; ++q) acc[q] = fma(x[q][u], uj, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < RL; ++q) acc[q] = warp_sum(acc[q]);
  for (int idx = lane; idx < RL * CL; idx += 32) {
    const int q = idx / CL, t = idx - q * CL;
    double val = 0.0;
#pragma unroll
    for (int qq = 0; qq < RL; ++qq)
      if (qq == q) val = acc[qq];
    tc_push(zdst + first + q * W, bar, (unsigned)t, val);
  }
}

// Symmetric symv: only the LOWER triangle is read (half the L2 traffic).  Element x = A0[r][j], j <= r, of an owned row
// adds x u_j to the row sum of r and x u_r to the column sum of j (the diagonal element lands in both; the caller
// takes it out of the column sum again, diag[] in shared memory).  A warp owns groups of FOUR ADJACENT rows
// 4 (gw + W t) .. + 3, so a group meets its diagonal in a single 128-column block: every block to the left of it
// runs without masks.  16-byte loads (n even), two per row and block: eight independent 128-bit loads per lane in
// flight.  The loops are ROLLED on purpose - the fully unrolled version was 250 KB of straight-line code run once per
// column, i.e. instruction-fetch bound.  The four row sums of a block go through a transposing butterfly (six
// shuffles) into one accumulator per lane class; the column sums of the lane's columns accumulate in the warp's own
// shared-memory vector.  rowp: [np] row sums (written by the owner warp), colw: this warp's [np] column sums.
__device__ __forceinline__ double2 tc_ld2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

__device__ __forceinline__ double tc_rows4_sum(const double (&ra)[4], bool h16, bool h8) {
  const double t0 = (h16 ? ra[2] : ra[0]) + __shfl_xor_sync(MOP_FULL_MASK, h16 ? ra[0] : ra[2], 16);
  const double t1 = (h16 ? ra[3] : ra[1]) + __shfl_xor_sync(MOP_FULL_MASK, h16 ? ra[1] : ra[3], 16);
  double sres = (h8 ? t1 : t0) + __shfl_xor_sync(MOP_FULL_MASK, h8 ? t0 : t1, 8);
  sres += __shfl_xor_sync(MOP_FULL_MASK, sres, 4);
  sres += __shfl_xor_sync(MOP_FULL_MASK, sres, 2);
  sres += __shfl_xor_sync(MOP_FULL_MASK, sres, 1);
  return sres;  // lanes of class 2 (bit 4) + (bit 3) = q hold the sum of row q
}

__device__ __forceinline__ void tc_symv_sym(const double* A, int n, int k, int gw, int W, int lane,
                                            const double* __restrict__ uu, double* rowp, double* colw, int abl,
                                            long long* prof) {
  for (int j = 2 * lane; j < n; j += 64) *reinterpret_cast<double2*>(colw + j) = make_double2(0.0, 0.0);
  const bool h16 = lane & 16, h8 = lane & 8;
  const int cb0 = (k + 1) >> 7;  // 128-column blocks to the left hold dead columns only
  int t = 0;
  if (k + 1 > 4 * gw + 3) t = (k + 1 - 4 * gw - 3 + 4 * W - 1) / (4 * W);  // first group with a live row
#pragma unroll 1
  for (;; ++t) {
    const int base = 4 * (gw + W * t);
    if (base >= n) break;
    const bool full = base > k && base + 3 < n;
    double ur[4];
    const double* p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = base + q;
      const bool live = r > k && r < n;
      ur[q] = live ? uu[r] : 0.0;
      p[q] = A + (size_t)(live ? r : (base < n - 1 ? base : n - 1)) * n + 2 * lane;
    }
    const int rmax = base + 3 < n ? base + 3 : n - 1;
    double racc = 0.0;
    int cb = cb0;
    long long tq0 = prof ? clock64() : 0;
    if (full) {
#pragma unroll 1
      for (; 128 * (cb + 1) <= base; ++cb) {  // every column of the block is left of every row's diagonal
        double2 x[4][2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q][u] = tc_ld2(p[q] + 128 * cb + 64 * u);
        double ra[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int j = 128 * cb + 64 * u + 2 * lane;
          const double2 uj = *reinterpret_cast<const double2*>(uu + j);
          double2 ca = *reinterpret_cast<const double2*>(colw + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ra[q] = fma(x[q][u].x, uj.x, ra[q]);
            ra[q] = fma(x[q][u].y, uj.y, ra[q]);
            ca.x = fma(x[q][u].x, ur[q], ca.x);
            ca.y = fma(x[q][u].y, ur[q], ca.y);
          }
          *reinterpret_cast<double2*>(colw + j) = ca;
        }
        racc += tc_rows4_sum(ra, h16, h8);
        if (prof) prof[2] += 1;
      }
    }
    if (prof) {
      const long long t_ = clock64();
      prof[0] += t_ - tq0;
      tq0 = t_;
    }
#pragma unroll 1
    for (; 128 * cb <= rmax; ++cb) {  // the block(s) with the diagonals (or an incomplete group): masked
      if (prof) prof[3] += 1;
      double2 x[4][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 128 * cb + 64 * u + 2 * lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int r = base + q;
          const bool live = r > k && r < n;
          x[q][u] = (live && j <= r && !(abl & 1)) ? tc_ld2(p[q] + 128 * cb + 64 * u) : make_double2(0.0, 0.0);
          if (j + 1 > r) x[q][u].y = 0.0;  // the pair straddles the diagonal
        }
      }
      double ra[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 128 * cb + 64 * u + 2 * lane;
        if (j < n) {  // (n even: j + 1 < n as well)
          const double2 uj = *reinterpret_cast<const double2*>(uu + j);
          double2 ca = *reinterpret_cast<const double2*>(colw + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ra[q] = fma(x[q][u].x, uj.x, ra[q]);
            ra[q] = fma(x[q][u].y, uj.y, ra[q]);
            ca.x = fma(x[q][u].x, ur[q], ca.x);
            ca.y = fma(x[q][u].y, ur[q], ca.y);
          }
          *reinterpret_cast<double2*>(colw + j) = ca;
        }
      }
      racc += tc_rows4_sum(ra, h16, h8);
    }
    if (prof) prof[1] += clock64() - tq0;
    if ((lane & 7) == 0) {  // lanes 0, 8, 16, 24 hold rows 0, 1, 2, 3 of the group
      const int r = base + (lane >> 3);
      if (r > k && r < n) rowp[r] = racc;
    }
  }
}

// A[i][j] -= sum_l V(i, l) W(j, l) + W(i, l) V(j, l) on the square i, j >= kn (LOWER: tiles on and below the diagonal
// only): C + (-P) Q^T with P = [V | W], Q = [W | V] (K = 12).  Strips of four 8 x 8 tiles are dealt to the warps of the
// cluster; the C tiles of the next strip are loaded before the DMMAs of the current one are issued.
template <bool LOWER>
__device__ __forceinline__ void tc_trailing_update(double* A, const double* Vp, const double* Wp, int n, int np, int kn,
                                                   int lane, int gw, int W) {
  constexpr int NB = TC_NB;
  const int g = lane >> 2, t = lane & 3;
  const int mt = (n - kn + 7) >> 3;
  const int njg = (mt + 3) >> 2;
  const bool vec = ((n | kn) & 1) == 0;  // (row * n + kn + 8 J + 2 t) even: 16-byte accesses
  // strip s -> (I, Jg).  LOWER: row I has (I >> 2) + 1 strips; enumerated through the full grid, skipping Jg > I / 4
  auto strip = [&](int s, int& I, int& Jg) {
    I = s / njg;
    Jg = s - I * njg;
    return s < mt * njg && (!LOWER || 4 * Jg <= I);
  };
  auto load_c = [&](int I, int Jg, double (&c0)[4], double (&c1)[4]) {
    const int ri = kn + 8 * I + g;
    const double* row = A + (size_t)(ri < n ? ri : n - 1) * n;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int J = 4 * Jg + u, cj = kn + 8 * J + 2 * t;
      const bool jt = J < mt && (!LOWER || J <= I) && ri < n;
      c0[u] = c1[u] = 0.0;
      if (jt && vec && cj + 1 < n) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(row + cj));
        c0[u] = v.x;
        c1[u] = v.y;
      } else if (jt) {
        if (cj < n) c0[u] = __ldcg(row + cj);
        if (cj + 1 < n) c1[u] = __ldcg(row + cj + 1);
      }
    }
  };
  int s = gw, I = 0, Jg = 0;
  while (s < mt * njg && !strip(s, I, Jg)) s += W;
  double c0[4], c1[4];
  if (s < mt * njg) load_c(I, Jg, c0, c1);
  while (s < mt * njg) {
    int sn = s + W, In = 0, Jn = 0;
    while (sn < mt * njg && !strip(sn, In, Jn)) sn += W;
    double d0[4], d1[4];
    if (sn < mt * njg) load_c(In, Jn, d0, d1);  // next strip's tiles in flight during this strip's DMMAs
    const int ri = kn + 8 * I + g;
    const int ric = ri < n ? ri : n - 1;
    double a[3];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
      const int c = 4 * ks + t;
      const double pa = c < NB ? Vp[c * np + ric] : Wp[(c - NB) * np + ric];
      a[ks] = ri < n ? -pa : 0.0;
    }
    double* row = A + (size_t)ric * n;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int J = 4 * Jg + u;
      const bool jt = J < mt && (!LOWER || J <= I);  // warp-uniform
      if (jt) {
        const int rj = kn + 8 * J + g;
        const bool jin = rj < n;
        const int rjc = jin ? rj : n - 1;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const int c = 4 * ks + t;
          const double qb = c < NB ? Wp[c * np + rjc] : Vp[(c - NB) * np + rjc];
          dmma884(c0[u], c1[u], a[ks], jin ? qb : 0.0, c0[u], c1[u]);
        }
        const int cj = kn + 8 * J + 2 * t;
        if (ri < n) {
          if (vec && cj + 1 < n) {
            *reinterpret_cast<double2*>(row + cj) = make_double2(c0[u], c1[u]);
          } else {
            if (cj < n) row[cj] = c0[u];
            if (cj + 1 < n) row[cj + 1] = c1[u];
          }
        }
      }
    }
    s = sn;
    I = In;
    Jg = Jn;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      c0[u] = d0[u];
      c1[u] = d1[u];
    }
  }
}

// SYM: the symv reads the lower triangle only (clusters of at most four CTAs); every CTA sums the row and
// column parts of its warps into one vector and pushes it to its peers, which add the CL vectors in rank order.
template <bool SYM, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1) k_lg_tridiag_blk(TcArgs a) {
  constexpr int THREADS = TC_THREADS, NW = TC_NW, NB = TC_NB, NPT = TC_NPT;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
  const int b = blockIdx.x / CL;
  const int n = a.n, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  extern __shared__ __align__(16) double sm[];
  double* Vp = sm;                       // [NB][np] panel reflectors (replicated in every CTA)
  double* Wp = Vp + (size_t)NB * np;     // [NB][np]
  double* uu = Wp + (size_t)NB * np;     // [np] raw updated column (zero up to k)
  double* zz = uu + np;                  // [2][np] symv results, pushed by the owners, double-buffered by column parity
                                         // (SYM: [2][CL][np], one vector per source CTA)
  double* rowp = zz + (size_t)(SYM ? 2 * CL : 2) * np;  // SYM: [np] row sums of the owned rows
  double* diag = rowp + (SYM ? np : 0);                 // SYM: [np] diagonal of the panel-start matrix
  double* colp = diag + (SYM ? np : 0);                 // SYM: [NW][np] column sums of every warp
  double* red = colp + (SYM ? (size_t)NW * np : 0);     // [2][16][NW]
  double* tot = red + red16_doubles(NW);  // [NW][16]
  double* pub = tot + 16 * NW;           // [8]
  double* s_rb = pub + 8;                // [64] block_sum_k<1> scratch
  __shared__ __align__(8) unsigned long long mbar[2];  // z exchange, one per column parity
  // (one spare np keeps the carve of tc_smem_doubles simple)
  int parity = 0, parity2 = 0;
  double* A = a.A + (size_t)b * n * n;
  double* Vh = a.Vh + (size_t)b * n * n;
  const int W = CL * NW, gw = wid * CL + cr;

  for (int i = tid; i < (int)tc_smem_doubles(n, SYM ? CL : 0); i += THREADS) sm[i] = 0.0;
  __syncthreads();
  for (int i = tid; i < n; i += THREADS) uu[i] = i > 0 ? __ldcg(A + (size_t)i * n) : 0.0;  // raw column 0
  if (SYM)
    for (int i = tid; i < n; i += THREADS) diag[i] = __ldcg(A + (size_t)i * n + i);
  if (cr == 0 && tid == 0) a.dd[(size_t)b * n] = __ldcg(A);
  if (tid == 0) {
    tc_mbar_init(mbar, 1);
    tc_mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();  // every CTA of the cluster is running, zeroed and its barriers initialised before the first remote store

  long long seg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ts = clock64();
  long long prof[4] = {0, 0, 0, 0};
#define TSEG(q)                          \
  do {                                   \
    if (DBG && a.dbg) {                  \
      const long long tn_ = clock64();   \
      seg[q] += tn_ - ts;                \
      ts = tn_;                          \
    }                                    \
  } while (0)
  int k0 = 0;
#pragma unroll 1
  for (int k = 0; k < n - 2; ++k) {
    const int jj = k - k0;
    // z of rows k+1 .. n-1: one value per row from its owner, SYM: one per row from every CTA
    if (tid == 0) tc_mbar_expect(mbar + (k & 1), 8u * (unsigned)(n - k - 1) * (SYM ? (unsigned)CL : 1u));
    // ---- row k+1 of the panel-start matrix (= its column k+1): issued first, used after the symv ----
    double crow[NPT];
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * THREADS;
      crow[e] = (i > k && i < n && !(DBG && (a.ablate & 8))) ? __ldcg(A + (size_t)i * n + k + 1) : 0.0;  // (column access: the lower triangle is
                                                                              // the one every variant keeps current)
    }
    // ---- z = A0 u: the warp's live rows (> k) ----
    {
      const int q0 = (k + 1 > gw) ? (k + 1 - gw + W - 1) / W : 0;
      const int first = gw + q0 * W;
      const int rl = first < n ? (n - 1 - first) / W + 1 : 0;
      unsigned long long* bar = mbar + (k & 1);
      if (SYM) {
        if (!DBG || !(a.ablate & 16))
          tc_symv_sym(A, n, k, gw, W, lane, uu, rowp, colp + (size_t)wid * np, DBG ? a.ablate : 0,
                      (DBG && a.dbg && tid == 0 && cr == 0) ? prof : nullptr);
        TSEG(0);
        __syncthreads();
        TSEG(7);
        double* zdst = zz + ((size_t)(k & 1) * CL + cr) * np;  // this CTA's slot in every peer
#pragma unroll
        for (int e = 0; e < NPT; ++e) {
          const int i = tid + e * THREADS;
          if (i > k && i < n) {
            // owned rows: the row sum, minus the diagonal term that the symv also put into the column sum
            double acc = ((i >> 2) % CL == cr) ? fma(-diag[i], uu[i], rowp[i]) : 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) acc += colp[(size_t)w * np + i];
            for (int t = 0; t < CL; ++t) tc_push(zdst + i, bar, (unsigned)t, acc);
          }
        }
      } else {
        double* zdst = zz + (size_t)(k & 1) * np;
        int q = 0;
#pragma unroll 1
        for (; q + 4 <= rl; q += 4) tc_symv_rows<4>(CL, A, n, k, first + q * W, W, lane, uu, zdst, bar);
#pragma unroll 1
        for (; q < rl; ++q) tc_symv_rows<1>(CL, A, n, k, first + q * W, W, lane, uu, zdst, bar);
      }
    }
    TSEG(0);
    // ---- c = current column k+1, the panel products with u, the norm: one reduction ----
    double Vk[NB], Wk[NB];
#pragma unroll
    for (int l = 0; l < NB; ++l) {
      Vk[l] = l < jj ? Vp[l * np + k + 1] : 0.0;
      Wk[l] = l < jj ? Wp[l * np + k + 1] : 0.0;
    }
    const double alpha = uu[k + 1];
    double ui[NPT], ci[NPT], rd[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) rd[q] = 0.0;
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * THREADS;
      const bool act = i > k && i < n;
      ui[e] = act ? uu[i] : 0.0;
      ci[e] = 0.0;
      if (act) {
        double t[NB];
#pragma unroll
        for (int l = 0; l < NB; ++l) {
          const double vi = l < jj ? Vp[l * np + i] : 0.0, wi = l < jj ? Wp[l * np + i] : 0.0;
          t[l] = fma(vi, Wk[l], wi * Vk[l]);
          rd[4 + l] = fma(vi, ui[e], rd[4 + l]);
          rd[4 + NB + l] = fma(wi, ui[e], rd[4 + NB + l]);
        }
        ci[e] = crow[e] - (((t[0] + t[1]) + (t[2] + t[3])) + (t[4] + t[5]));
        if (i >= k + 2) {
          rd[1] = fma(ci[e], ui[e], rd[1]);
          rd[3] = fma(ui[e], ui[e], rd[3]);
        }
        if (i == k + 1) pub[1] = ci[e];
      }
    }
    TSEG(1);
    block_sum16<NW>(rd, red, tot, parity, lane, wid, true);
    TSEG(2);
    tc_mbar_wait(mbar + (k & 1), (unsigned)(k >> 1) & 1u);  // z of this column is complete in this CTA
    TSEG(3);
    const double* z = zz + (size_t)(k & 1) * (SYM ? CL : 1) * np;
    auto zval = [&](int i) {
      double v = z[i];
      if (SYM)
        for (int c = 1; c < CL; ++c) v += z[(size_t)c * np + i];  // rank order: identical in every CTA
      return v;
    };
    double s1[1] = {0.0};
    double zi[NPT];
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * THREADS;
      zi[e] = (i > k && i < n) ? zval(i) : 0.0;
      if (i >= k + 2 && i < n) s1[0] = fma(zi[e], ui[e], s1[0]);
    }
    block_sum_k<1>(s1, s_rb, parity2);
    TSEG(4);
    // ---- Householder scalars, w, v, the raw next column (every CTA, identically) ----
    {
      const double zk1 = zval(k + 1), ck1 = pub[1];
      const double xn2 = rd[3];
      const bool refl = xn2 > 0.0;
      const double nrm = sqrt(fma(alpha, alpha, xn2));
      const double beta = refl ? -copysign(nrm, alpha) : alpha;
      const double tk = refl ? (beta - alpha) / beta : 0.0;
      const double scal = refl ? 1.0 / (alpha - beta) : 0.0;
      const double ca = 1.0 - scal * alpha;
      double tk1[NB], ts1[NB];
#pragma unroll
      for (int l = 0; l < NB; ++l) {
        const double vtu = rd[4 + l], wtu = rd[4 + NB + l];
        tk1[l] = fma(Vk[l], wtu, Wk[l] * vtu);
        ts1[l] = fma(wtu, vtu - alpha * Vk[l], vtu * (wtu - alpha * Wk[l]));
      }
      const double qk1 = zk1 - (((tk1[0] + tk1[1]) + (tk1[2] + tk1[3])) + (tk1[4] + tk1[5]));
      const double S1 = s1[0] - (((ts1[0] + ts1[1]) + (ts1[2] + ts1[3])) + (ts1[4] + ts1[5]));
      const double p0 = tk * fma(scal, qk1, ca * ck1);
      const double pv = p0 + tk * scal * fma(scal, S1, ca * rd[1]);
      const double alpha2 = -0.5 * tk * pv;
      const double w0 = p0 + alpha2;  // w_{k+1}
#pragma unroll
      for (int e = 0; e < NPT; ++e) {
        const int i = tid + e * THREADS;
        if (i > k && i < n) {
          const bool first = i == k + 1;
          double tq[NB];
#pragma unroll
          for (int l = 0; l < NB; ++l) {
            const double vi = l < jj ? Vp[l * np + i] : 0.0, wi = l < jj ? Wp[l * np + i] : 0.0;
            tq[l] = fma(vi, rd[4 + NB + l], wi * rd[4 + l]);
          }
          const double qi = zi[e] - (((tq[0] + tq[1]) + (tq[2] + tq[3])) + (tq[4] + tq[5]));
          const double vi = first ? 1.0 : ui[e] * scal;
          const double pi = tk * fma(scal, qi, ca * ci[e]);
          const double wi = fma(alpha2, vi, pi);
          const double un = ci[e] - fma(vi, w0, wi);  // raw column k+1 after reflector k
          Vp[jj * np + i] = vi;
          Wp[jj * np + i] = wi;
          uu[i] = first ? 0.0 : un;
          if (cr == k % CL) Vh[(size_t)k * n + i] = vi;
          if (first && cr == 0) {
            a.ee[(size_t)b * n + k] = beta;
            a.tau[(size_t)b * n + k] = tk;
            a.dd[(size_t)b * n + k + 1] = un;
          }
        }
      }
    }
    __syncthreads();
    TSEG(5);
    if (jj == NB - 1 && k + 1 < n - 2) {  // panel complete: rank-12 update of the trailing square on the tensor cores
      tc_trailing_update<SYM>(A, Vp, Wp, n, np, k + 1, lane, gw, W);
      k0 = k + 1;
      cluster.sync();  // the updated rows are in L2 before anyone reads them
      if (a.hout && k0 == a.kstop) {
        const int m = n - k0;
        const size_t nlm = ((size_t)m * (m + 1) / 2 + 1) & ~(size_t)1;
        double* h = a.hout + (size_t)b * a.hstride;
        for (int r = k0 + gw; r < n; r += W) {
          const double* src = A + (size_t)r * n + k0;
          double* dst = h + (((size_t)(r - k0) * (r - k0 + 1)) >> 1);
          for (int j = lane; j <= r - k0; j += 32) dst[j] = __ldcg(src + j);
        }
        if (cr == 0)
          for (int r = k0 + tid; r < n; r += THREADS) {
            h[nlm + r - k0] = uu[r];
            h[nlm + m + r - k0] = 0.0;
          }
        cluster.sync();  // nobody leaves while a peer may still store into its shared memory
        return;
      }
      if (SYM) {
        for (int i = k0 + tid; i < n; i += THREADS) diag[i] = __ldcg(A + (size_t)i * n + i);
        __syncthreads();
      }
      TSEG(6);
    }
  }
  if (DBG && a.dbg && tid == 0 && cr == 0) {
    for (int q = 0; q < 8; ++q) a.dbg[(size_t)b * 16 + q] = seg[q];
    for (int q = 0; q < 4; ++q) a.dbg[(size_t)b * 16 + 8 + q] = prof[q];
  }
#undef TSEG
  // e_{n-2} is the raw column n-2; the last diagonal element still lacks the reflectors of the open panel
  if (cr == 0 && tid == 0) {
    const int i0 = n - 2, i1 = n - 1, jj = i0 - k0;
    double dl = __ldcg(A + (size_t)i1 * n + i1);
    for (int l = 0; l < jj; ++l) dl -= 2.0 * Vp[l * np + i1] * Wp[l * np + i1];
    a.ee[(size_t)b * n + i0] = uu[i1];
    a.tau[(size_t)b * n + i0] = 0.0;
    a.dd[(size_t)b * n + i1] = dl;
    a.ee[(size_t)b * n + i1] = 0.0;
    a.tau[(size_t)b * n + i1] = 0.0;
    Vh[(size_t)i0 * n + i1] = 1.0;
  }
  cluster.sync();  // nobody leaves while a peer may still store into its shared memory
}

}  // namespace mop

// tridiag_blocked.cu: continue a reduction from the state the cluster kernel left at hand [B][hstride] (region 0)
constexpr size_t MOP_TB_RESUME_REGION = 13312;  // doubles per hand-over region: triangle of 160 rows + 2 x 160, rounded
int mop_launch_tridiag_blk_resume(int B, int nfull, int row0, double* hand, size_t hstride, double* Vh, double* dd,
                                  double* ee, double* tau, double* gq, int* flag, cudaStream_t stream);

static long long* g_tc_dbg = nullptr;
static int g_tc_ablate = 0;
extern "C" int mop_priv_tridiag_cluster_ablate(int mask) {
  g_tc_ablate = mask;
  return MOP_OK;
}
static int g_tc_sym = 1;
extern "C" int mop_priv_tridiag_cluster_sym(int on) {
  g_tc_sym = on;
  return MOP_OK;
}
extern "C" int mop_priv_tridiag_cluster_timing(void* buf) {
  g_tc_dbg = (long long*)buf;
  return MOP_OK;
}

int mop_tridiag_cluster_supported(int n) { return n >= 3 && n <= 1024; }

// A: [B][n][n] symmetric working copies (destroyed); Vh: [B][n][n]; dd, ee, tau: [B][n].  cluster_ctas: 8 (or 4, 2, 1).
// gq_scratch [B][n] doubles + flag_scratch [B] ints (or null: no hand-over to the shared-memory kernel)
int mop_launch_tridiag_cluster(int B, int n, double* A, double* Vh, double* dd, double* ee, double* tau,
                               int cluster_ctas, double* gq_scratch, int* flag_scratch, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_cluster_supported(n)) {
    mop_set_error("blocked cluster tridiagonalisation: n = %d not supported (3 .. 1024)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  // Cluster size: the column step is a latency chain, so throughput wants MANY small clusters (measured at n = 600,
  // 256 matrices: 102 / 75 / 69 ms per mop_eigh batch with 8 / 4 / 2 CTAs per matrix) and a small batch wants its
  // matrices spread over the whole GPU.
  // A batch that does not fill the GPU takes the LARGEST cluster that still runs it as ONE wave (per-matrix latency
  // 7.2 / 10.9 / 17 ms at 8 / 4 / 2 CTAs, n = 600): 32 matrices on 8-CTA clusters were two waves of 18 (14.4 ms), on
  // 4-CTA clusters they are one (10.9 ms); 64 matrices: one wave of 2-CTA clusters instead of two of 4-CTA clusters.
  int CL = cluster_ctas;
  if (CL != 1 && CL != 2 && CL != 4 && CL != 8) CL = 8 * B <= 148 ? 8 : (4 * B <= 148 ? 4 : 2);
  mop::TcArgs a{n, A, Vh, dd, ee, tau, g_tc_dbg, g_tc_ablate, 0, nullptr, 0};
  // The last <= 160 rows go to the staged shared-memory kernel (2 - 16 matrices per SM instead of one cluster of SMs per
  // matrix; every column there costs the cluster a fixed ~12 k cycles).  The hand-over buffers are the top rows of the
  // matrix's own slab - finished, never read again - so the slab must be big enough that they stay clear of the rows
  // >= kstop the copy reads.
  int kstop = 0;
  if (gq_scratch && flag_scratch && !g_tc_dbg && !g_tc_ablate) {
    kstop = ((n - 160 + mop::TC_NB - 1) / mop::TC_NB) * mop::TC_NB;
    if (kstop < mop::TC_NB || n - kstop < 24 || (size_t)kstop * n < 3 * MOP_TB_RESUME_REGION) kstop = 0;
  }
  if (kstop) {
    a.kstop = kstop;
    a.hout = A;
    a.hstride = (size_t)n * n;
  }
  // lower-triangle symv (half the L2 traffic) whenever the lane-private column sums fit the registers and the
  // cluster is small enough for the all-to-all of the per-CTA vectors
  // (measured at n = 600: 2 CTAs 68.7 ms with / 73.9 without, 4 CTAs 80.4 with / 74.9 without per 256 matrices;
  // 32 matrices on 4 CTAs 12.0 / 11.3 ms: the all-to-all of four per-CTA vectors costs more than the halved loads save)
  const bool sym = g_tc_sym && CL <= 2 && n % 2 == 0;
  const size_t smem = sizeof(double) * mop::tc_smem_doubles(n, sym ? CL : 0);
  if (smem > 227 * 1024) {
    mop_set_error("blocked cluster tridiagonalisation: n = %d needs %zu bytes of shared memory", n, smem);
    return MOP_ERR_UNSUPPORTED;
  }
  auto kern = g_tc_dbg ? (sym ? mop::k_lg_tridiag_blk<true, true> : mop::k_lg_tridiag_blk<false, true>)
                       : (sym ? mop::k_lg_tridiag_blk<true, false> : mop::k_lg_tridiag_blk<false, false>);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CL));
  cfg.blockDim = dim3(mop::TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MOP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  if (!kstop) return MOP_OK;
  MOP_CHECK_CUDA(cudaMemsetAsync(flag_scratch, 0, sizeof(int) * (size_t)B, stream));
  return mop_launch_tridiag_blk_resume(B, n, kstop, A, (size_t)n * n, Vh, dd, ee, tau, gq_scratch, flag_scratch, stream);
}
