// Batched symmetric eigensolver for matrices that do not fit in one SM's shared memory
// (160 < n <= 1024; BASELINE config 5: P-RFO at N = 200 atoms, n = 600).
//
// The matrix stays in global memory (a batch slice that is being reduced is L2-resident) and
// one thread-block CLUSTER reduces one matrix, so that a single matrix is streamed by the
// L2 ports of several SMs and the per-column synchronisation is a hardware cluster barrier:
//   1. k_lg_tridiag_blk (tridiag_cluster.cu): blocked Householder tridiagonalisation A = Q T Q^T (dlatrd panels,
//      DMMA trailing updates); n <= 160: k_tridiag_blk (tridiag_blocked.cu), one CTA per matrix.
//   2. k_lg_trieig: eigenvalues of T by bisection on the Sturm count, eigenvectors by twisted
//      factorisation, CGS2 inside clusters - the algorithm of eigh_tridiag.cu with Z in
//      global memory (column i = vector i).
//   3. k_lg_backtransform: V = Q Z, one warp per eigenvector held in registers, reflectors
//      staged through shared memory once per CTA of 32 vectors; rows written in ascending order.
// Structures whose clusters cancel are flagged MOP_ST_EIG_FALLBACK and redone by the Jacobi
// kernel (global-memory variant).  Replaces numpy.linalg.eigh at Optimizer/rsprfo.py:783,798
// and Optimizer/rsirfo.py:606 for large systems.
#include <cooperative_groups.h>

#include <type_traits>

#include "tri_sturm.cuh"

namespace cg = cooperative_groups;

namespace mop {

constexpr int LG_MAX_N = 1024;
constexpr int LG_EIG_THREADS = 1024;
constexpr int LG_BT_THREADS = 512;
// Re-orthogonalisation threshold on eigenvalue gaps relative to ||T||.  A twisted-factorisation
// vector is accurate to about eps ||T|| / gap, so neighbours further apart than this are orthogonal
// to ~1e-12 already; dstein's 1e-3 would chain most of a dense 600-eigenvalue spectrum.
constexpr double LG_GAPTOL = 3e-4;

struct LgArgs {
  int n;
  double* A;     // [B][n][n] working copy (symmetric, destroyed)
  double* Vh;    // [B][n][n] reflector k in row k, columns k+1.. (v[k+1] = 1 stored)
  double* Z;     // [B][n][n] eigenvectors of T, column i = vector i
  double* Dm;    // [B][n][n] backward pivots
  double* dd;    // [B][n]
  double* ee;    // [B][n]
  double* tau;   // [B][n]
  double* pbuf;  // [B][2][n] A v, double-buffered by column parity
  int* rank;     // [B][n] ascending rank of vector i
  double* evals; // [B][n] out
  double* evecs; // [B][n][n] out
  int32_t* status;
  long long* dbg;
  int ablate;  // diagnostics: bit0 skip the trailing stores, bit1 skip the trailing loads (results invalid)
};

// A_out = 1/2 (A + A^T), 32 x 32 tiles
__global__ void __launch_bounds__(256) k_lg_symcopy(int n, const double* __restrict__ Ain, double* __restrict__ Aout) {
  __shared__ double t[32][33];
  const size_t off = (size_t)blockIdx.z * n * n;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = j0 + r, j = i0 + tx;  // transposed tile
    t[r][tx] = (i < n && j < n) ? Ain[off + (size_t)i * n + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    if (i < n && j < n) Aout[off + (size_t)i * n + j] = 0.5 * (Ain[off + (size_t)i * n + j] + t[tx][r]);
  }
}

// ---- inverse iteration for a cluster vector that cancelled (LAPACK dstein's method) -----------------
// Twisted-factorisation vectors of (numerically) multiple eigenvalues are nearly parallel, so
// orthogonalising them against each other leaves rounding noise.  Such a vector is rebuilt by inverse
// iteration on the unreduced block [s, t): random start, (T - lam I)^{-1} by Gaussian elimination with
// partial pivoting (dgttrf / dgtts2 recurrences, tiny pivots replaced), re-orthogonalisation against the
// cluster vectors found so far after every solve.  One warp; lane 0 runs the O(m) recurrences in shared
// memory.  fix: 6 arrays of np doubles (x | dl | dd | du | du2 | piv).  prev: rows [0, nprev) of R
// (stride lds) are the finished cluster vectors.  Result in fix[0 .. m).
__device__ void lg_inverse_iteration(const double* d, const double* e, int s, int t, double lam, const double* R, size_t lds,
                                     int nprev, double* fix, int np, int lane, unsigned seed) {
  const int m = t - s;
  double* x = fix;
  double* dl = fix + np;
  double* dd = fix + 2 * np;
  double* du = fix + 3 * np;
  double* du2 = fix + 4 * np;
  double* pv = fix + 5 * np;
  const double tiny = TRI_EPS;  // T is scaled to unit norm
  for (int i = lane; i < m; i += 32) {
    unsigned h = (seed + 0x9e3779b9u * (unsigned)(i + 1));
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    x[i] = (double)(h & 0xffffffu) / 8388608.0 - 1.0;
  }
  __syncwarp();
  if (lane == 0) {  // factorise T - lam I
    for (int i = 0; i < m; ++i) {
      dd[i] = d[s + i] - lam;
      if (i < m - 1) {
        du[i] = e[s + i];
        dl[i] = e[s + i];
      }
      du2[i] = 0.0;
      pv[i] = 0.0;
    }
    for (int i = 0; i < m - 1; ++i) {
      if (fabs(dd[i]) >= fabs(dl[i])) {
        if (fabs(dd[i]) < tiny) dd[i] = dd[i] < 0.0 ? -tiny : tiny;
        const double f = dl[i] / dd[i];
        dl[i] = f;
        dd[i + 1] -= f * du[i];
      } else {
        const double f = dd[i] / dl[i];
        dd[i] = dl[i];
        dl[i] = f;
        const double tmp = du[i];
        du[i] = dd[i + 1];
        dd[i + 1] = tmp - f * dd[i + 1];
        if (i < m - 2) {
          du2[i] = du[i + 1];
          du[i + 1] = -f * du[i + 1];
        }
        pv[i] = 1.0;
      }
    }
    if (fabs(dd[m - 1]) < tiny) dd[m - 1] = dd[m - 1] < 0.0 ? -tiny : tiny;
  }
  __syncwarp();
  for (int it = 0; it < 4; ++it) {
    // orthogonalise against the finished cluster vectors (twice), normalise
    for (int rep = 0; rep < 2; ++rep)
      for (int p = 0; p < nprev; ++p) {
        const double* rp = R + (size_t)p * lds + s;
        double dt = 0.0;
        for (int i = lane; i < m; i += 32) dt = fma(rp[i], x[i], dt);
        dt = warp_sum(dt);
        for (int i = lane; i < m; i += 32) x[i] = fma(-dt, rp[i], x[i]);
        __syncwarp();
      }
    double nn = 0.0;
    for (int i = lane; i < m; i += 32) nn = fma(x[i], x[i], nn);
    nn = sqrt(warp_sum(nn));
    const double sc = nn > 0.0 ? 1.0 / nn : 0.0;
    for (int i = lane; i < m; i += 32) x[i] *= sc;
    __syncwarp();
    if (it == 3) break;
    if (lane == 0) {  // x <- (T - lam I)^{-1} x
      for (int i = 0; i < m - 1; ++i) {
        if (pv[i] == 0.0) {
          x[i + 1] -= dl[i] * x[i];
        } else {
          const double tmp = x[i];
          x[i] = x[i + 1];
          x[i + 1] = tmp - dl[i] * x[i];
        }
      }
      x[m - 1] /= dd[m - 1];
      if (m > 1) x[m - 2] = (x[m - 2] - du[m - 2] * x[m - 1]) / dd[m - 2];
      for (int i = m - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
      double mx = 0.0;  // keep the iterate finite
      for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(x[i]));
      if (mx > 0.0 && isfinite(mx)) {
        const double r = 1.0 / mx;
        for (int i = 0; i < m; ++i) x[i] *= r;
      }
    }
    __syncwarp();
  }
}

// ---- 2. eigenpairs of T, one CTA per matrix ---------------------------------------------------
// THREADS = 1024 for large n (three-way multisection, forward and backward sweeps on separate threads); 160 for
// n <= 160, where one thread per eigenpair is all the phases can use and seven CTAs per SM overlap their chains
// (P-RFO at n = 150, 1024 structures: 2.2 ms at 1024 threads, 1.3 ms at 256 x 4 CTAs, 1.1 ms at 160 x 7)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 160 ? 7 : (THREADS <= 256 ? 4 : 1))) k_lg_trieig(LgArgs a) {
  constexpr int NW = THREADS / 32;
  extern __shared__ double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  double* d = sm;
  double* e = d + np;
  double* e2 = e + np;
  double* lam = e2 + np;
  double* X = lam + np;          // lo | hi, then nrm_up | nrm_dn, then dots (NW * 64 <= 2 np needs np >= 1024: separate)
  double* dots = X + 2 * np;     // NW * 64
  int* blk_s = (int*)(dots + NW * 64);
  int* blk_e = blk_s + np;
  int* cl_s = blk_e + np;
  int* twist = cl_s + np;
  int* cnts = twist + np;        // P * np (P <= 3)
  double* fixb = (double*)(cnts + 3 * np + (np & 1));  // 6 np: inverse-iteration scratch, one user at a time
  __shared__ double s_red[40];
  __shared__ double s_tnorm;
  __shared__ int s_fallback, s_fixlock;
  double* S = a.Z + (size_t)b * n * n;
  double* Dm = a.Dm + (size_t)b * n * n;
  const size_t lds = n;
  int st_in = a.status ? a.status[b] : 0;
  st_in &= ~(MOP_ST_EIG_FALLBACK | MOP_ST_EIG_NOCONV);
  if (tid == 0) {
    s_fallback = 0;
    s_fixlock = 0;
  }
  long long tq = clock64();
  int tslot = 0;
#define LG_EMARK()                                                                    \
  do {                                                                                \
    if (a.dbg && tid == 0) {                                                          \
      const long long tnow = clock64();                                               \
      a.dbg[(size_t)gridDim.x * 4 + (size_t)b * 4 + (tslot++)] = tnow - tq;           \
      tq = tnow;                                                                      \
    }                                                                                 \
  } while (0)
  double tn = 0.0;
  for (int i = tid; i < n; i += THREADS) {
    d[i] = a.dd[(size_t)b * n + i];
    e[i] = (i < n - 1) ? a.ee[(size_t)b * n + i] : 0.0;
    tn = fmax(tn, fmax(fabs(d[i]), fabs(e[i])));
    if (!isfinite(d[i]) || !isfinite(e[i])) tn = INFINITY;
  }
  tn = block_max(tn, s_red);
  if (tid == 0) s_tnorm = tn;
  const bool zero_t = tn == 0.0 || !isfinite(tn);
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
    // ---- eigenvalues: grid bracket, then a barrier-free safeguarded secant (Illinois) iteration per eigenvalue ----
    // (as k_spectrum_step: one Sturm evaluation per thread on a uniform grid over the block's Gershgorin interval brackets
    // every eigenvalue; the bracket then moves on the Sturm count alone, exactly as in bisection, and p_n(x) - the last
    // term of the same recurrence - only proposes the next abscissa once the bracket isolates one eigenvalue, with a
    // forced bisection whenever two evaluations fail to halve it.  The predecessor bisected all eigenvalues in lock
    // step from the Gershgorin interval: ~52 evaluations and 104 barriers per eigenvalue instead of ~20 and none.)
    // scratch: the inverse-iteration buffer (idle until the cluster repair) and the count table
    static_assert(THREADS >= 160, "one thread per eigenvalue");
    {
      double2* de = (double2*)fixb;            // (d_k, e_{k-1}^2): the Sturm table, np pairs
      double* gpv = fixb + 2 * np;             // p_n at the grid points
      double* gbl = fixb + 3 * np;             // Gershgorin bounds, stored at the block start
      double* gbu = fixb + 4 * np;
      int* gcnt = cnts;
      int* gexp = cnts + np;
      for (int i = tid; i < n; i += THREADS) de[i] = make_double2(d[i], i > 0 ? e2[i - 1] : 0.0);
      for (int i = tid; i < n; i += THREADS)
        if (blk_s[i] == i) {
          const int s0 = i, t0 = blk_e[i];
          double gl = INFINITY, gu = -INFINITY;
          for (int r = s0; r < t0; ++r) {
            const double rad = (r > s0 ? fabs(e[r - 1]) : 0.0) + (r < t0 - 1 ? fabs(e[r]) : 0.0);
            gl = fmin(gl, d[r] - rad);
            gu = fmax(gu, d[r] + rad);
          }
          const double pad = 4.0 * TRI_EPS * n * fmax(fabs(gl), fabs(gu)) + 1e-300;
          gbl[i] = gl - pad;
          gbu[i] = gu + pad;
        }
      __syncthreads();
      for (int i = tid; i < n; i += THREADS) {
        const int bs = blk_s[i], bt = blk_e[i], m = bt - bs, j = i - bs;
        const double gl = gbl[bs], gu = gbu[bs];
        double pv;
        int pe;
        gcnt[i] = sturm_eval(de, bs, bt, gl + (gu - gl) * ((double)(j + 1) / (double)(m + 1)), &pv, &pe);
        gpv[i] = pv;
        gexp[i] = pe;
      }
      __syncthreads();
      for (int i = tid; i < n; i += THREADS) {
        const int bs = blk_s[i], bt = blk_e[i], m = bt - bs, j = i - bs, want = j + 1;
        const double gl = gbl[bs], gu = gbu[bs];
        int qlo = -1, qhi = m;   // first grid index with count >= want (sentinels: -1 -> gl, m -> gu)
        while (qhi - qlo > 1) {
          const int q = (qlo + qhi) >> 1;
          if (gcnt[bs + q] >= want) qhi = q;
          else qlo = q;
        }
        double l = gl, h = gu, pl = 0.0, ph = 0.0;
        int el = 0, eh = 0, cl = 0, ch = m;
        bool okl = false, okh = false;
        if (qlo >= 0) {
          l = gl + (gu - gl) * ((double)(qlo + 1) / (double)(m + 1));
          cl = gcnt[bs + qlo]; pl = gpv[bs + qlo]; el = gexp[bs + qlo]; okl = true;
        }
        if (qhi < m) {
          h = gl + (gu - gl) * ((double)(qhi + 1) / (double)(m + 1));
          ch = gcnt[bs + qhi]; ph = gpv[bs + qhi]; eh = gexp[bs + qhi]; okh = true;
        }
        double wa = INFINITY, wb = INFINITY;
        int side = 0;
        for (int round = 0; round < 200; ++round) {
          const double width = h - l;
          const double tolw = 2.0 * TRI_EPS * fmax(fabs(l), fabs(h)) + 4e-3 * TRI_EPS;
          if (!(width > tolw)) break;
          const bool force = width > 0.5 * wa;
          wa = wb;
          wb = width;
          double x = l + 0.5 * width;
          if (!force && okl && okh && ch - cl == 1 && pl != 0.0 && ph != 0.0) {
            int dex = eh - el;
            dex = dex < -1000 ? -1000 : (dex > 1000 ? 1000 : dex);
            const double rho = (ph / pl) * __hiloint2double((1023 + dex) << 20, 0);  // f(h) / f(l) < 0
            const double frac = 1.0 / (1.0 - rho);
            if (frac > 0.0 && frac < 1.0) {
              const double ms = 0.5 * tolw;
              x = fmin(fmax(l + width * frac, l + ms), h - ms);
            }
          }
          if (!(x > l && x < h)) break;
          double pv;
          int pe;
          const int c = sturm_eval(de, bs, bt, x, &pv, &pe);
          if (c >= want) {
            h = x; ph = pv; eh = pe; ch = c; okh = true;
            if (side > 0) pl *= 0.5;   // Illinois: the retained end keeps losing weight
            side = 1;
          } else {
            l = x; pl = pv; el = pe; cl = c; okl = true;
            if (side < 0) ph *= 0.5;
            side = -1;
          }
        }
        lam[i] = 0.5 * (l + h);
      }
    }
    __syncthreads();
    LG_EMARK();

    // ---- twisted-factorisation eigenvectors, thread i owns column i of S ----
    {
      const double piv = TRI_EPS * 1e-3;
      double* nrm_up = X;
      double* nrm_dn = X + np;
      for (int i = tid; i < n; i += THREADS) {
        const int s = blk_s[i], t = blk_e[i];
        const double l = lam[i];
        for (int k = 0; k < n; ++k)
          if (k < s || k >= t) S[k * lds + i] = 0.0;
        double q = d[s] - l;
        for (int k = s; k < t; ++k) {
          if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
          const double r = fast_rcp(q);
          S[k * lds + i] = r;
          if (k + 1 < t) q = fma(-e2[k], r, d[k + 1] - l);
        }
        // backward sweep with the twist search fused in: gamma_k = |D+_k - e2_k / D-_{k+1}| needs the forward reciprocal
        // of row k-1 (eight loads in flight) and the backward reciprocal of row k+1, which is the previous iteration's r
        double best = INFINITY;
        int rr = s;
        {
          double q = d[t - 1] - l, rm = 0.0;
          for (int k0 = t - 1; k0 >= s; k0 -= 8) {
            double sp[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) sp[u] = (k0 - u > s) ? S[(k0 - u - 1) * lds + i] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int k = k0 - u;
              if (k >= s) {
                double dp = d[k] - l;
                if (k > s) dp = fma(-e2[k - 1], sp[u], dp);
                const double gam = fabs(fma(-e2[k], rm, dp));   // e2[t-1] == 0 closes the block
                if (gam <= best) {   // descending scan, ties to the smaller index: the first minimum of an ascending scan
                  best = gam;
                  rr = k;
                }
                if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
                const double r = fast_rcp(q);
                Dm[k * lds + i] = r;
                rm = r;
                if (k > s) q = fma(-e2[k - 1], r, d[k - 1] - l);
              }
            }
          }
        }
        twist[i] = rr;
        // z_r = 1; z_k = -e_k z_{k+1} / D+_k (k < r); z_k = -e_{k-1} z_{k-1} / D-_k (k > r): the recurrences run twice - a
        // read-only pass for the norm, then the pass that stores the NORMALISED vector (one write of S instead of a
        // write, a read and a second write)
        double au = 0.0, ad = 0.0, sc = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
          double z = 1.0;
          for (int k0 = rr - 1; k0 >= s; k0 -= 8) {
            double f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = (k0 - u >= s) ? -(e[k0 - u] * S[(k0 - u) * lds + i]) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (k0 - u >= s) {
                z *= f[u];
                if (pass) S[(k0 - u) * lds + i] = z * sc;
                else au = fma(z, z, au);
              }
            }
          }
          z = 1.0;
          for (int k0 = rr + 1; k0 < t; k0 += 8) {
            double f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = (k0 + u < t) ? -(e[k0 + u - 1] * Dm[(k0 + u) * lds + i]) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (k0 + u < t) {
                z *= f[u];
                if (pass) S[(k0 + u) * lds + i] = z * sc;
                else ad = fma(z, z, ad);
              }
            }
          }
          if (!pass) sc = 1.0 / sqrt(1.0 + au + ad);
        }
        S[rr * lds + i] = sc;
        if (!isfinite(sc) || sc == 0.0) s_fallback = 1;
        nrm_up[i] = au;
        nrm_dn[i] = ad;
      }
    }
    if (tid == 0) {
      int cs = 0;
      for (int i = 0; i < n; ++i) {
        const bool chain = i > 0 && blk_s[i] == blk_s[i - 1] && (lam[i] - lam[i - 1]) < LG_GAPTOL;
        if (!chain) cs = i;
        cl_s[i] = cs;
      }
    }
    __syncthreads();
    LG_EMARK();
    // ---- CGS2 inside clusters, one warp per cluster ----
    // The chain's vectors are columns of S (stride n between consecutive elements): they are copied
    // once into rows of the dead pivot buffer Dm (contiguous), orthogonalised there with coalesced
    // accesses, and copied back.
    for (int c0 = wid; c0 < n; c0 += NW) {
      if (cl_s[c0] != c0) continue;
      int cend = c0 + 1;
      while (cend < n && cl_s[cend] == c0) ++cend;
      if (cend - c0 < 2) continue;
      const int s = blk_s[c0], t = blk_e[c0];
      double* dw = dots + wid * 64;
      double* R = Dm + (size_t)c0 * lds;  // rows c0 .. cend-1 of Dm belong to this chain
      for (int p = c0; p < cend; ++p)
        for (int k = s + lane; k < t; k += 32) R[(size_t)(p - c0) * lds + k] = S[k * lds + p];
      __syncwarp();
      for (int c = c0 + 1; c < cend; ++c) {
        double* rc = R + (size_t)(c - c0) * lds;
        double nfirst = 1.0;
        for (int rep = 0; rep < 2; ++rep) {
          for (int p0 = c0; p0 < c; p0 += 64) {
            const int pe = min(c, p0 + 64);
            for (int p = p0; p < pe; ++p) {
              const double* rp = R + (size_t)(p - c0) * lds;
              double dt = 0.0;
              for (int k = s + lane; k < t; k += 32) dt = fma(rp[k], rc[k], dt);
              dt = warp_sum(dt);
              if (lane == 0) dw[p - p0] = dt;
            }
            __syncwarp();
            for (int k = s + lane; k < t; k += 32) {
              double zc = rc[k];
              for (int p = p0; p < pe; ++p) zc = fma(-dw[p - p0], R[(size_t)(p - c0) * lds + k], zc);
              rc[k] = zc;
            }
            __syncwarp();
          }
          double nn = 0.0;
          for (int k = s + lane; k < t; k += 32) nn = fma(rc[k], rc[k], nn);
          nn = sqrt(warp_sum(nn));
          if (rep == 0) nfirst = nn;
          if (!(nn > 1e-2)) {
            // the vector lies (numerically) in the span of its cluster: rebuild it by inverse iteration
            if (lane == 0)
              while (atomicCAS(&s_fixlock, 0, 1) != 0) {
              }
            __syncwarp();
            __threadfence_block();
            lg_inverse_iteration(d, e, s, t, lam[c], R, lds, c - c0, fixb, np, lane, (unsigned)(b * 7919 + c));
            double chk = 0.0;
            for (int k = s + lane; k < t; k += 32) {
              const double xv = fixb[k - s];
              rc[k] = xv;
              chk = fma(xv, xv, chk);
            }
            chk = warp_sum(chk);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
              if (!(chk > 0.5) || !isfinite(chk)) s_fallback = 1;  // not even inverse iteration produced a vector
              atomicExch(&s_fixlock, 0);
            }
            break;  // orthonormal against the cluster by construction
          }
          const double sc = nn > 0.0 ? 1.0 / nn : 0.0;
          for (int k = s + lane; k < t; k += 32) rc[k] *= sc;
          __syncwarp();
          if (rep == 0 && nfirst > 0.7) break;
        }
      }
      for (int p = c0 + 1; p < cend; ++p)
        for (int k = s + lane; k < t; k += 32) S[k * lds + p] = R[(size_t)(p - c0) * lds + k];
    }
    __syncthreads();
  } else {
    for (int i = tid; i < n; i += THREADS) {
      lam[i] = isfinite(tn) ? 0.0 : NAN;
      for (int k = 0; k < n; ++k) S[k * lds + i] = (k == i) ? 1.0 : 0.0;
    }
    __syncthreads();
  }
  LG_EMARK();
  // ascending rank over all blocks
  int* rank = a.rank + (size_t)b * n;
  const double tnorm = zero_t ? 1.0 : s_tnorm;
  for (int i = tid; i < n; i += THREADS) {
    const double li = lam[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (lam[j] < li) || (lam[j] == li && j < i) || (li != li && j < i);
    rank[i] = r;
    a.evals[(size_t)b * n + r] = li * tnorm;
  }
  if (tid == 0 && a.status) a.status[b] = st_in | (s_fallback ? MOP_ST_EIG_FALLBACK : 0);
}

// ---- 3. V = Q Z, one warp per eigenvector -------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(LG_BT_THREADS, 1) k_lg_backtransform(LgArgs a) {
  constexpr int NW = LG_BT_THREADS / 32;
  constexpr int CH = 4;  // reflectors staged per barrier
  constexpr int RS = 32 * NPL;    // staged reflector row: zero beyond n and up to the unit entry, so the dot and
                                  // the update below need no masks (ncu: 132 instructions per reflector and warp
                                  // with them, half of the kernel)
  extern __shared__ double sm[];  // [2][CH][RS] | tau [np]
  const int n = a.n, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  const int i = blockIdx.x * NW + wid;  // vector (column of Z)
  const bool live = i < n;
  const double* S = a.Z + (size_t)b * n * n;
  const double* Vh = a.Vh + (size_t)b * n * n;
  const double* tau = a.tau + (size_t)b * n;
  if (a.status && (a.status[b] & MOP_ST_EIG_FALLBACK)) return;  // the robust path redoes it
  double z[NPL];
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    z[q] = (live && j < n) ? S[(size_t)j * n + i] : 0.0;
  }
  // reflectors n-3 .. 0 in chunks of CH (descending); the next chunk travels from global memory to
  // registers while the current one is applied from shared memory
  const int last = n - 3;
  // staging map: reflector u of the chunk, element j = tid (+ THREADS for n > THREADS) - no index division, and at
  // small n only the first n threads load (the flat (u, j) = idx / n map spent ~1 k instructions per thread and
  // chunk on divisions and masked slots: 7.0 ms of the 10.4 ms of mop_eigh at n = 150)
  constexpr int JT = (LG_MAX_N + LG_BT_THREADS - 1) / LG_BT_THREADS;  // elements per thread and reflector
  double pf[CH][JT];
  auto prefetch = [&](int kc) {
#pragma unroll
    for (int u = 0; u < CH; ++u) {
      const int k = kc - u;
#pragma unroll
      for (int t = 0; t < JT; ++t) {
        const int j = tid + t * LG_BT_THREADS;
        pf[u][t] = (k >= 0 && j < n && j >= k + 1) ? Vh[(size_t)k * n + j] : 0.0;
      }
    }
  };
  if (last >= 0) prefetch(last);
  double* tau_s = sm + 2 * (size_t)CH * RS;  // a global load per reflector would sit on the critical path
  for (int j = tid; j < n; j += LG_BT_THREADS) tau_s[j] = tau[j];
  for (int j = tid; j < 2 * CH * RS; j += LG_BT_THREADS) sm[j] = 0.0;  // the slots j >= n stay zero
  __syncthreads();
  int buf = 0;
  for (int kc = last; kc >= 0; kc -= CH) {
    double* vb = sm + (size_t)buf * CH * RS;
#pragma unroll
    for (int u = 0; u < CH; ++u)
#pragma unroll
      for (int t = 0; t < JT; ++t) {
        const int j = tid + t * LG_BT_THREADS;
        if (j < n) vb[u * RS + j] = pf[u][t];
      }
    __syncthreads();
    if (kc - CH >= 0) prefetch(kc - CH);
    for (int u = 0; u < CH; ++u) {
      const int k = kc - u;
      if (k < 0) break;
      const double tk = tau_s[k];
      if (tk == 0.0) continue;
      const double* vk = vb + u * RS + lane;
      double vq[NPL], d0 = 0.0, d1 = 0.0;
#pragma unroll
      for (int q = 0; q < NPL; ++q) vq[q] = vk[32 * q];
#pragma unroll
      for (int q = 0; q < NPL; q += 2) {
        d0 = fma(vq[q], z[q], d0);
        if (q + 1 < NPL) d1 = fma(vq[q + 1], z[q + 1], d1);
      }
      const double dot = warp_sum(d0 + d1) * tk;
#pragma unroll
      for (int q = 0; q < NPL; ++q) z[q] = fma(-dot, vq[q], z[q]);
    }
    buf ^= 1;
  }
  if (!live) return;
  const int r = a.rank[(size_t)b * n + i];
  double* out = a.evecs + ((size_t)b * n + r) * n;
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    if (j < n) out[j] = z[q];
  }
}


// ---- 4. factored form: Zt (rows = eigenvectors of T, ascending) and x <- Q x / Q^T x -------------
// Consumers that only need V^T g and V c (the RFO / P-RFO steps) rotate their vectors into the
// basis of T instead of forming V = Q Z: O(n^2) per vector instead of the 2 n^3 back-transform.
__global__ void __launch_bounds__(256) k_lg_transpose_sorted(int n, const double* __restrict__ Z,
                                                             const int* __restrict__ rank, double* __restrict__ Zt) {
  __shared__ double t[32][33];
  const size_t off = (size_t)blockIdx.z * n * n;
  const int* rk = rank + (size_t)blockIdx.z * n;
  const int k0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, i = i0 + tx;
    t[r][tx] = (k < n && i < n) ? Z[off + (size_t)k * n + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, k = k0 + tx;
    if (i < n && k < n) Zt[off + (size_t)rk[i] * n + k] = t[tx][r];
  }
}

struct LgVecs {
  double* x[4];  // each [B][n]
  int m;
};

// warp w of CTA b transforms x[w][b]: trans = 1: Q^T x (reflectors ascending), 0: Q x (descending)
template <int NPL>
__global__ void __launch_bounds__(128) k_lg_apply_q(int n, int trans, const double* __restrict__ Vh_all,
                                                    const double* __restrict__ tau_all, LgVecs X) {
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (wid >= X.m) return;
  const double* Vh = Vh_all + (size_t)b * n * n;
  const double* tau = tau_all + (size_t)b * n;
  double* x = X.x[wid] + (size_t)b * n;
  double z[NPL], vc[NPL], vn[NPL];
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    z[q] = j < n ? x[j] : 0.0;
  }
  const int nref = n - 2;  // reflectors 0 .. n-3
  if (nref <= 0) return;
  auto load = [&](int k, double* dst) {
#pragma unroll
    for (int q = 0; q < NPL; ++q) {
      const int j = lane + 32 * q;
      dst[q] = (j >= k + 1 && j < n) ? Vh[(size_t)k * n + j] : 0.0;
    }
  };
  int k = trans ? 0 : nref - 1;
  const int step = trans ? 1 : -1;
  load(k, vc);
  for (int it = 0; it < nref; ++it) {
    const int kn = k + step;
    if (it + 1 < nref) load(kn, vn);
    const double tk = tau[k];
    if (tk != 0.0) {
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < NPL; ++q) dot = fma(vc[q], z[q], dot);
      dot = warp_sum(dot) * tk;
#pragma unroll
      for (int q = 0; q < NPL; ++q) z[q] = fma(-dot, vc[q], z[q]);
    }
#pragma unroll
    for (int q = 0; q < NPL; ++q) vc[q] = vn[q];
    k = kn;
  }
#pragma unroll
  for (int q = 0; q < NPL; ++q) {
    const int j = lane + 32 * q;
    if (j < n) x[j] = z[q];
  }
}

// structures redone by the Jacobi kernel get Q = I (their vectors are in the original basis)
__global__ void k_lg_clear_tau_flagged(int n, const int32_t* __restrict__ status, double* __restrict__ tau) {
  const int b = blockIdx.x;
  if (!(status[b] & MOP_ST_EIG_FALLBACK)) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) tau[(size_t)b * n + i] = 0.0;
}

}  // namespace mop

// ------------------------------------------------------------------------------------------------
static size_t lg_al(size_t x) { return (x + 255) & ~(size_t)255; }
static int g_lg_cluster = 0;
// tuning (csrc/mop_private.h): CTAs per matrix of the cluster tridiagonalisation (1, 2, 4, 8; 0 = by batch size)
extern "C" int mop_priv_large_cluster(int cl) {
  g_lg_cluster = cl;
  return MOP_OK;
}
int mop_launch_tridiag_cluster(int B, int n, double* A, double* Vh, double* dd, double* ee, double* tau,
                               int cluster_ctas, double* gq_scratch, int* flag_scratch, cudaStream_t stream);
int mop_launch_tridiag_blk(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                           double* tau, double* gq, int* flag, double* hand, cudaStream_t stream);

// diagnostics (csrc/mop_private.h): device buffer [2][B][4] receiving the phase clocks of k_lg_trieig (second half)
static long long* g_lg_dbg = nullptr;
extern "C" int mop_priv_large_timing(void* buf) {
  g_lg_dbg = (long long*)buf;
  return MOP_OK;
}

int mop_large_supported(int n) { return n >= 3 && n <= mop::LG_MAX_N; }

size_t mop_large_workspace_bytes(int B, int n) {
  const size_t nn = lg_al(sizeof(double) * (size_t)B * n * n), nv = lg_al(sizeof(double) * (size_t)B * n);
  return 4 * nn + 5 * nv + lg_al(sizeof(int) * (size_t)B * n);
}

template <int NPL>
static int lg_launch_bt(int B, const mop::LgArgs& a, cudaStream_t stream) {
  const int np = (a.n + 3) & ~3;
  const size_t smem = sizeof(double) * (2 * 4 * (size_t)(32 * NPL) + (size_t)np);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_lg_backtransform<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.n + (mop::LG_BT_THREADS / 32) - 1) / (mop::LG_BT_THREADS / 32), B);
  mop::k_lg_backtransform<NPL><<<grid, mop::LG_BT_THREADS, smem, stream>>>(a);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

static void lg_carve(int B, int n, void* work, mop::LgArgs& a) {
  const size_t nn = lg_al(sizeof(double) * (size_t)B * n * n), nv = lg_al(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  a.n = n;
  a.A = (double*)w;
  a.Vh = (double*)(w + nn);
  a.Z = (double*)(w + 2 * nn);
  a.Dm = (double*)(w + 3 * nn);
  a.dd = (double*)(w + 4 * nn);
  a.ee = (double*)(w + 4 * nn + nv);
  a.tau = (double*)(w + 4 * nn + 2 * nv);
  a.pbuf = (double*)(w + 4 * nn + 3 * nv);  // 2 nv
  a.rank = (int*)(w + 4 * nn + 5 * nv);
}

// A = Q T Q^T and the eigenpairs of T: leaves reflectors (Vh, tau), Z (columns) and rank in `work`.
static int lg_factor(int B, int n, const double* A, double* evals, int32_t* status, void* work, size_t work_bytes,
                     mop::LgArgs& a, cudaStream_t stream) {
  if (!mop_large_supported(n)) {
    mop_set_error("large-n eigensolver: n = %d not supported (3..%d)", n, mop::LG_MAX_N);
    return MOP_ERR_UNSUPPORTED;
  }
  if (!work || work_bytes < mop_large_workspace_bytes(B, n)) {
    mop_set_error("large-n eigensolver: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  lg_carve(B, n, work, a);
  a.evals = evals;
  a.evecs = nullptr;
  a.status = status;
  a.dbg = g_lg_dbg;
  a.ablate = 0;
  {
    dim3 grid((n + 31) / 32, (n + 31) / 32, B);
    mop::k_lg_symcopy<<<grid, 256, 0, stream>>>(n, A, a.A);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  const int np = (n + 3) & ~3;
  if (n <= 160 && g_lg_cluster == 0) {
    // the matrix fits one SM: the blocked shared-memory tridiagonalisation (two structures per SM) replaces the
    // cluster kernel; same outputs (reflector rows with explicit unit entries, d, e, tau)
    double* gq_dummy = a.pbuf;
    int* flag = (int*)(a.pbuf + (size_t)B * n);
    int rc = mop_launch_tridiag_blk(B, n, a.A, nullptr, a.Vh, a.dd, a.ee, a.tau, gq_dummy, flag, a.A /* staged hand-over in place */, stream);
    if (rc != MOP_OK) return rc;
  } else {
    int rc = mop_launch_tridiag_cluster(B, n, a.A, a.Vh, a.dd, a.ee, a.tau, g_lg_cluster, a.pbuf,
                                        (int*)(a.pbuf + (size_t)B * n), stream);
    if (rc != MOP_OK) return rc;
  }
  {
    // (a 640-thread instantiation with two CTAs per SM - 256 matrices of order 600 as one wave instead of two - was
    // measured slower: 48 registers spill the Sturm loop, 10.0 ms against 9.2 ms; tools/trieig_phases.py)
    const int thr = n <= 160 ? 160 : mop::LG_EIG_THREADS;
    const size_t smem = sizeof(double) * (12 * (size_t)np + (thr / 32) * 64 + 2) + sizeof(int) * 7 * (size_t)np;
    if (thr == 160) {
      MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_lg_trieig<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      mop::k_lg_trieig<160><<<B, 160, smem, stream>>>(a);
    } else {
      MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_lg_trieig<mop::LG_EIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      mop::k_lg_trieig<mop::LG_EIG_THREADS><<<B, mop::LG_EIG_THREADS, smem, stream>>>(a);
    }
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  return MOP_OK;
}

// eigh for 160 < n <= 1024: evals ascending, evecs rows = eigenvectors.  Flags structures whose
// clusters cancelled with MOP_ST_EIG_FALLBACK (caller runs the Jacobi kernel on those).
int mop_launch_eigh_large(int B, int n, const double* A, double* evals, double* evecs, int32_t* status,
                          void* work, size_t work_bytes, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  mop::LgArgs a{};
  int rc = lg_factor(B, n, A, evals, status, work, work_bytes, a, stream);
  if (rc != MOP_OK) return rc;
  a.evecs = evecs;
  const int npl = (n + 31) / 32;
  if (npl <= 5) return lg_launch_bt<5>(B, a, stream);
  if (npl <= 8) return lg_launch_bt<8>(B, a, stream);
  if (npl <= 12) return lg_launch_bt<12>(B, a, stream);
  if (npl <= 16) return lg_launch_bt<16>(B, a, stream);
  if (npl <= 20) return lg_launch_bt<20>(B, a, stream);
  if (npl <= 24) return lg_launch_bt<24>(B, a, stream);
  if (npl <= 28) return lg_launch_bt<28>(B, a, stream);
  return lg_launch_bt<32>(B, a, stream);
}

// Factored eigendecomposition: evals ascending, Zt [B][n][n] rows = eigenvectors of T in the same
// order; A = Q (Zt^T diag(evals) Zt) Q^T with Q kept in `work` for mop_launch_large_apply_q.
// Structures flagged MOP_ST_EIG_FALLBACK get Q = I here; the caller runs the Jacobi kernel on them
// with Zt as its eigenvector output.
int mop_launch_eigh_large_factored(int B, int n, const double* A, double* evals, double* Zt, int32_t* status,
                                   void* work, size_t work_bytes, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  mop::LgArgs a{};
  int rc = lg_factor(B, n, A, evals, status, work, work_bytes, a, stream);
  if (rc != MOP_OK) return rc;
  dim3 grid((n + 31) / 32, (n + 31) / 32, B);
  mop::k_lg_transpose_sorted<<<grid, 256, 0, stream>>>(n, a.Z, a.rank, Zt);
  MOP_CHECK_CUDA(cudaGetLastError());
  mop::k_lg_clear_tau_flagged<<<B, 256, 0, stream>>>(n, status, a.tau);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

template <int NPL>
static int lg_launch_apply(int B, int n, int trans, const mop::LgArgs& a, const mop::LgVecs& X, cudaStream_t stream) {
  mop::k_lg_apply_q<NPL><<<B, 128, 0, stream>>>(n, trans, a.Vh, a.tau, X);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// x_i <- Q^T x_i (trans = 1) or Q x_i (trans = 0) for up to four [B][n] vectors, Q from the
// last mop_launch_eigh_large_factored on the same workspace.
int mop_launch_large_apply_q(int B, int n, int trans, void* work, double* x0, double* x1, double* x2, double* x3,
                             cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  mop::LgArgs a{};
  lg_carve(B, n, work, a);
  mop::LgVecs X{};
  double* xs[4] = {x0, x1, x2, x3};
  X.m = 0;
  for (int i = 0; i < 4; ++i)
    if (xs[i]) X.x[X.m++] = xs[i];
  if (X.m == 0) return MOP_OK;
  const int npl = (n + 31) / 32;
  if (npl <= 8) return lg_launch_apply<8>(B, n, trans, a, X, stream);
  if (npl <= 16) return lg_launch_apply<16>(B, n, trans, a, X, stream);
  if (npl <= 24) return lg_launch_apply<24>(B, n, trans, a, X, stream);
  return lg_launch_apply<32>(B, n, trans, a, X, stream);
}
