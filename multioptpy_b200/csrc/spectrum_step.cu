// Spectrum + RS-I-RFO step of a tridiagonalised Hessian, SEVEN structures per SM (n <= 160).
//
// The first-generation kernel kept the eigenvectors Z of T in shared memory: one CTA per SM, and
// every phase after the tridiagonalisation is a long dependent chain run by a few hundred threads
// (Sturm bisection 53 %, twisted factorisation 20 %, the one-warp secular solve and back-transform the
// rest) - the SM idles.  Here Z lives in global memory (one [n][n] slab per structure, column i owned by
// thread i, so every access is coalesced across the CTA), the CTA shrinks to 160 threads and 26 KB of
// shared memory, and seven CTAs share an SM: the whole C2 batch (1024 structures) is resident at once and
// the chains of different structures overlap.
//
//   1. scale T to unit norm, split at negligible off-diagonals (as k_eigh_tridiag);
//   2. eigenvalues: one Sturm count per thread on a uniform grid over the block's Gershgorin interval
//      brackets every eigenvalue to (range / m), then each thread bisects ITS eigenvalue without any
//      barrier;
//   3. eigenvectors of T by twisted factorisation, thread i = vector i: backward pivots -> Dm, forward
//      pivots -> Z with the twist index found on the fly, then the two z recurrences; vectors stay
//      unnormalised in Z, the scale is kept per column (zsc);
//   4. clusters (gap < 1e-3 ||T||): CGS2, one warp per cluster; a vector that cancels flags the
//      structure for the Jacobi path (MOP_ST_EIG_FALLBACK), exactly as the shared-memory kernel does;
//   5. gamma = Z^T (Q^T gp), the RS-I-RFO step in the eigenbasis (rfo_core.cuh, rsirfo.py:360-490),
//      y = Z c by warps over rows, step = Q y with the reflector rows read from L2.
//
// Inputs come from k_tridiag_packed (T, tau, Q^T gp, reflector rows).  Replaces numpy.linalg.eigh at
// Optimizer/rsirfo.py:606,626,652 and the step algebra of rsirfo.py:360-490 on the fused path.
#include "rfo_core.cuh"
#include "tri_sturm.cuh"

namespace mop {

constexpr int SP_MAX_N = 160;
// The CTA has one thread per row / eigenvalue, rounded up to whole warps (NW = ceil(n / 32) <= 5), and as many CTAs
// share an SM as 64 registers per thread allow (58 at NW = 5): small structures (config 4: n = 72, n = 24) are
// latency chains just like the large ones, only more of them fit.
template <int NW>
struct SpMinBlocks {
  static constexpr int value = NW >= 5 ? 7 : (NW == 4 ? 8 : (NW == 3 ? 10 : (NW == 2 ? 16 : 32)));
};

struct SpArgs {
  int n;
  double* Z;          // [B][n][n] scratch: eigenvectors of T, column i = vector i (row-major rows k)
  double* Dm;         // [B][n][n] scratch: backward reciprocal pivots
  const double* Vh;   // [B][n][n] reflector k in row k, columns k+2.. (k_tridiag_packed)
  const double* pf_d;
  const double* pf_e;
  const double* pf_tau;
  const double* pf_gq;
  const int* pf_flag;
  double* evals;      // [B][n] out (ascending) or null
  int32_t* status;
  const double* Bg;   // [B][n] raw biased gradient (norm only)
  const double* Be;   // [B] or null
  double* state;      // [B][MOP_RSIRFO_STATE]
  double* move;       // [B][n]
  double* pred;       // [B] or null
  int saddle_order, neb_mode;
  double tmin, tmax;
  long long* dbg;     // optional [B][16] phase clocks (diagnostics)
};

// Shared-memory plan (doubles; np = n rounded up to 4):
//   tau | lam | gq | zsc | e2     5 np   live to the end (e2 is recycled as the coefficients of y = Z c)
//   Y:  d | e | de (2 np) | X     T itself, dead once the eigenvectors exist; X = grid-phase scratch (4 np)
// and from the cluster phase on the whole of Y is recycled: CGS2 pair buffers + dots, then lam_s | gam_s |
// RfoArrays, then the reflector ring.  Seven CTAs stay under the 164 KB carve-out (92 KB of L1 left for the
// local-memory spills and the strided cluster accesses).
__host__ __device__ inline size_t sp_y_doubles(int n, int SP_NW) {
  const size_t np = (size_t)((n + 3) & ~3);
  size_t y = 2 * np + rfo_core_smem_bytes(n) / sizeof(double) + 8;           // lam_s, gam_s, RfoArrays; ring 12 np
  if (y < 8 * np) y = 8 * np;                                                 // d, e, de + grid-phase scratch
  if (y < (size_t)SP_NW * (2 * np + 64)) y = (size_t)SP_NW * (2 * np + 64);   // CGS2 pair buffers + dots
  return (y + 1) & ~(size_t)1;
}

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}

// per-warp slot for the CGS2 dot products: behind the five 2-column pair buffers in X
__device__ __forceinline__ double* gq_dots(double* X, int np, int wid, int nw) { return X + (size_t)nw * 2 * np + wid * 64; }

template <int NW>
__global__ void __launch_bounds__(32 * NW, SpMinBlocks<NW>::value) k_spectrum_step(SpArgs a) {
  constexpr int THREADS = 32 * NW;
  extern __shared__ __align__(16) double sm[];
  const int n = a.n, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int np = (n + 3) & ~3;
  double* tau = sm;            // np
  double* lam = tau + np;      // np   eigenvalue of thread-row i (scaled)
  double* gq = lam + np;       // np   Q^T gp, later y
  double* zsc = gq + np;       // np   normalisation factor of column i of Z
  double* e2 = zsc + np;       // np   later: c_i * zsc_i (column coefficients of y = Z c)
  double* Y = e2 + np;         // recycled region (see sp_y_doubles)
  double* d = Y;               // np
  double* e = d + np;          // np   e[k] couples k, k+1
  double2* de = (double2*)(e + np);  // np pairs (d_k, e_{k-1}^2): the Sturm table (16-byte aligned: np % 4 == 0)
  double* X = e + 3 * np;      // grid-phase scratch, 4 np
  unsigned char* blk_s = (unsigned char*)(Y + sp_y_doubles(n, NW));  // index tables: n <= 160 fits a byte
  unsigned char* blk_e = blk_s + np;
  unsigned char* cl_s = blk_e + np;
  unsigned char* rank = cl_s + np;
  unsigned char* inv = rank + np;
  __shared__ double s_red[40];
  __shared__ double s_tnorm;
  __shared__ int s_fallback;

  double* Z = a.Z + (size_t)b * n * n;
  double* Dm = a.Dm + (size_t)b * n * n;
  const double* Vh = a.Vh + (size_t)b * n * n;
  int st_in = a.status ? a.status[b] : 0;
  st_in &= ~(MOP_ST_EIG_FALLBACK | MOP_ST_EIG_NOCONV);

  long long t_prev = clock64();
  int t_slot = 0;
#define SP_MARK()                                                       \
  do {                                                                  \
    if (a.dbg && tid == 0) {                                            \
      const long long t_now = clock64();                                \
      a.dbg[(size_t)b * 16 + (t_slot++)] = t_now - t_prev;               \
      t_prev = t_now;                                                   \
    }                                                                   \
  } while (0)

  for (int i = tid; i < n; i += THREADS) {
    d[i] = a.pf_d[(size_t)b * n + i];
    e[i] = a.pf_e[(size_t)b * n + i];
    tau[i] = a.pf_tau[(size_t)b * n + i];
    gq[i] = a.pf_gq[(size_t)b * n + i];
    zsc[i] = 1.0;
  }
  if (tid == 0) s_fallback = 0;
  const int fl = a.pf_flag[b];
  bool identity = fl == 2;                 // non-finite input: rsirfo.py:365-369 identity fallback
  const bool trivial = identity || fl == 1;

  // ---- scale, split --------------------------------------------------------------------------
  __syncthreads();
  double tn = 0.0;
  if (!trivial)
    for (int i = tid; i < n; i += THREADS) tn = fmax(tn, fmax(fabs(d[i]), fabs(e[i])));
  tn = block_max(tn, s_red);
  if (tid == 0) s_tnorm = tn;
  const bool zero_t = trivial || tn == 0.0 || !isfinite(tn);
  if (!trivial && !isfinite(tn)) identity = true;
  __syncthreads();
  double gdot = 0.0;  // column i's z . (Q^T gp), accumulated while the twisted vector is generated
  if (!zero_t) {
    const double inv_tn = 1.0 / tn;
    for (int i = tid; i < n; i += THREADS) {
      d[i] *= inv_tn;
      e[i] *= inv_tn;
    }
    __syncthreads();
    double enew = 0.0;
    if (tid < n) {
      enew = e[tid];
      if (tid < n - 1 && fabs(enew) <= TRI_EPS * (fabs(d[tid]) + fabs(d[tid + 1]))) enew = 0.0;
      if (tid == n - 1) enew = 0.0;
    }
    __syncthreads();
    if (tid < n) {
      e[tid] = enew;
      e2[tid] = enew * enew;
    }
    __syncthreads();
    if (tid < n) de[tid] = make_double2(d[tid], tid > 0 ? e2[tid - 1] : 0.0);
    // block of row i: [blk_s, blk_e); every thread scans outwards from its own row
    if (tid < n) {
      int s = tid, t = tid;
      while (s > 0 && e[s - 1] != 0.0) --s;
      while (t < n - 1 && e[t] != 0.0) ++t;
      blk_s[tid] = s;
      blk_e[tid] = t + 1;
    }
    __syncthreads();
    SP_MARK();  // 0: load, scale, split

    // ---- eigenvalues: grid bracket, then a barrier-free safeguarded secant (Illinois) iteration ------
    // Every evaluation returns the Sturm count AND p_n(x) = det(T - x I); the bracket [l, h] moves on the
    // count alone (count(l) < want <= count(h), exactly as in bisection), the polynomial values only
    // choose the next abscissa once the bracket isolates one eigenvalue, so a bad proposal costs time,
    // never correctness.  A step that fails to halve the bracket within two evaluations forces a bisection.
    double* gpv = X + np;                 // p_n at the grid points
    int* gcnt = (int*)(X + 2 * np);       // counts at the grid points
    int* gexp = gcnt + np;                // binary exponents of gpv
    double gl = INFINITY, gu = -INFINITY;
    int bs = 0, bt = 0;
    // Gershgorin interval of every block: the block's first row scans it once, the others read the result
    double* gbl = X;            // [np] lower bound, stored at the block start
    double* gbu = X + 3 * np;   // [np] upper bound (X holds at least 4 np doubles)
    if (tid < n && blk_s[tid] == tid) {
      const int s0 = tid, t0 = blk_e[tid];
      double a = INFINITY, c = -INFINITY;
      for (int r = s0; r < t0; ++r) {
        const double rad = (r > s0 ? fabs(e[r - 1]) : 0.0) + (r < t0 - 1 ? fabs(e[r]) : 0.0);
        a = fmin(a, d[r] - rad);
        c = fmax(c, d[r] + rad);
      }
      const double pad = 4.0 * TRI_EPS * n * fmax(fabs(a), fabs(c)) + 1e-300;
      gbl[tid] = a - pad;
      gbu[tid] = c + pad;
    }
    __syncthreads();
    if (tid < n) {
      bs = blk_s[tid];
      bt = blk_e[tid];
      gl = gbl[bs];
      gu = gbu[bs];
      const int m = bt - bs, j = tid - bs;
      const double x = gl + (gu - gl) * ((double)(j + 1) / (double)(m + 1));
      double pv;
      int pe;
      gcnt[tid] = sturm_eval(de, bs, bt, x, &pv, &pe);
      gpv[tid] = pv;
      gexp[tid] = pe;
    }
    __syncthreads();
    if (tid < n) {
      const int m = bt - bs, j = tid - bs, want = j + 1;
      // first grid index q in [0, m] with count(x_q) >= want (sentinels: -1 -> gl, m -> gu)
      int qlo = -1, qhi = m;
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
        cl = gcnt[bs + qlo];
        pl = gpv[bs + qlo];
        el = gexp[bs + qlo];
        okl = true;
      }
      if (qhi < m) {
        h = gl + (gu - gl) * ((double)(qhi + 1) / (double)(m + 1));
        ch = gcnt[bs + qhi];
        ph = gpv[bs + qhi];
        eh = gexp[bs + qhi];
        okh = true;
      }
      double wa = INFINITY, wb = INFINITY;
      int side = 0;
      const double dead_s = 1e-7 / s_tnorm;  // in units of ||T||
      for (int round = 0; round < 200; ++round) {
        const double width = h - l;
        // relative 2 eps plus the absolute floor eps ||T|| / 2 (dstebz: abstol = eps ||T||): below it the
        // computed p_n is rounding noise and the tridiagonal itself carries errors of that size
        const double tolw = 2.0 * TRI_EPS * fmax(fabs(l), fabs(h)) + 0.5 * TRI_EPS;
        if (!(width > tolw)) break;
        // modes the RFO step discards (the whole bracket inside |lambda| < 1e-7, rsirfo.py:30 filters < 1e-6) -
        // in practice the six TR/ROT null modes, one unresolvable cluster at 1e-17 ||T|| - need no more than
        // 1e-13 ||T||: their vectors are never used and 1e-13 is far below every threshold applied to eigenvalues
        if (width < 1e-13 && h < dead_s && l > -dead_s) break;
        const bool force = width > 0.5 * wa;
        wa = wb;
        wb = width;
        double x = l + 0.5 * width;
        if (!force && okl && okh && ch - cl == 1 && pl != 0.0 && ph != 0.0) {
          int de = eh - el;
          de = de < -1000 ? -1000 : (de > 1000 ? 1000 : de);
          const double rho = (ph / pl) * __hiloint2double((1023 + de) << 20, 0);  // f(h) / f(l) < 0
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
          h = x;
          ph = pv;
          eh = pe;
          ch = c;
          okh = true;
          if (side > 0) pl *= 0.5;  // Illinois: the retained end keeps losing weight
          side = 1;
        } else {
          l = x;
          pl = pv;
          el = pe;
          cl = c;
          okl = true;
          if (side < 0) ph *= 0.5;
          side = -1;
        }
      }
      lam[tid] = 0.5 * (l + h);
    }
    __syncthreads();
    SP_MARK();  // 1: eigenvalues

    // clusters: consecutive rows of one block whose eigenvalues are closer than GAPTOL; modes the RFO
    // step filters anyway (|lambda| < 1e-7 absolute, rsirfo.py:30 drops < 1e-6) break clusters
    if (tid < n) {
      const double dead = 1e-7 / s_tnorm;
      int cs = tid;
      while (cs > 0 && blk_s[cs] == blk_s[cs - 1] && (lam[cs] - lam[cs - 1]) < TRI_GAPTOL &&
             !(fabs(lam[cs]) < dead) && !(fabs(lam[cs - 1]) < dead))
        --cs;
      cl_s[tid] = cs;
    }

    // ---- eigenvectors by twisted factorisation; thread i owns column i of Z ----------------------
    if (tid < n) {
      const int i = tid, s = bs, t = bt;
      const double l = lam[i];
      const double piv = TRI_EPS * 1e-3;
      for (int k = 0; k < s; ++k) Z[(size_t)k * n + i] = 0.0;
      for (int k = t; k < n; ++k) Z[(size_t)k * n + i] = 0.0;
      {  // backward: q_k = (d_k - l) - e2_k / q_{k+1};  Dm[k] = 1 / q_k
        double q = d[t - 1] - l;
        for (int k = t - 1; k >= s; --k) {
          if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
          const double r = fast_rcp(q);
          Dm[(size_t)k * n + i] = r;
          if (k > s) q = fma(-e2[k - 1], r, d[k - 1] - l);
        }
      }
      // forward: q_k = (d_k - l) - e2_{k-1} / q_{k-1};  Z[k] = 1 / q_k;  twist at min |q_k - e2_k / D-_{k+1}|
      int r_tw = s;
      {
        double best = INFINITY;
        double q = d[s] - l;
        for (int k0 = s; k0 < t; k0 += 8) {
          double rm[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) rm[u] = (k0 + u + 1 < t) ? Dm[(size_t)(k0 + u + 1) * n + i] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + u;
            if (k < t) {
              const double gam = fabs(fma(-e2[k], rm[u], q));  // e2[t-1] == 0 closes the block
              if (gam < best) {
                best = gam;
                r_tw = k;
              }
              if (fabs(q) < piv) q = (q <= 0.0) ? -piv : piv;
              const double r = fast_rcp(q);
              Z[(size_t)k * n + i] = r;
              if (k + 1 < t) q = fma(-e2[k], r, d[k + 1] - l);
            }
          }
        }
      }
      // z_r = 1; z_k = -e_k z_{k+1} / D+_k (k < r); z_k = -e_{k-1} z_{k-1} / D-_k (k > r)
      double acc = 0.0;
      {
        double z = 1.0;
        for (int k0 = r_tw - 1; k0 >= s; k0 -= 8) {
          double f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = (k0 - u >= s) ? -(e[k0 - u] * Z[(size_t)(k0 - u) * n + i]) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (k0 - u >= s) {
              z *= f[u];
              Z[(size_t)(k0 - u) * n + i] = z;
              acc = fma(z, z, acc);
              gdot = fma(z, gq[k0 - u], gdot);
            }
          }
        }
        z = 1.0;
        for (int k0 = r_tw + 1; k0 < t; k0 += 8) {
          double f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = (k0 + u < t) ? -(e[k0 + u - 1] * Dm[(size_t)(k0 + u) * n + i]) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (k0 + u < t) {
              z *= f[u];
              Z[(size_t)(k0 + u) * n + i] = z;
              acc = fma(z, z, acc);
              gdot = fma(z, gq[k0 + u], gdot);
            }
          }
        }
        Z[(size_t)r_tw * n + i] = 1.0;
        gdot += gq[r_tw];
      }
      const double sc = 1.0 / sqrt(1.0 + acc);
      zsc[i] = sc;
      if (!isfinite(sc) || sc == 0.0) s_fallback = 1;
    }
    __syncthreads();
    SP_MARK();  // 2: twisted vectors

    // ---- CGS2 inside clusters, one warp per cluster -------------------------------------------------
    // The cluster's columns are gathered (scaled to unit norm) into a contiguous column-major copy - the
    // warp's slice of shared memory for pairs, the structure's Dm slab (dead after the twisted sweeps)
    // for larger clusters - orthogonalised there, and columns 1.. written back; column 0 is untouched.
    for (int c0 = wid; c0 < n; c0 += NW) {
      if (cl_s[c0] != c0) continue;
      int cend = c0 + 1;
      while (cend < n && cl_s[cend] == c0) ++cend;
      const int cs = cend - c0;
      if (cs < 2) continue;
      const int s = blk_s[c0], t = blk_e[c0], m = t - s;
      double* buf = (cs == 2) ? Y + (size_t)wid * 2 * np : Dm + (size_t)c0 * n;  // buf[col * m + row]
      for (int i0 = 0; i0 < cs * m; i0 += 32 * 8) {
        double zz[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + 32 * u + lane;
          zz[u] = 0.0;
          if (idx < cs * m) {
            const int row = idx / cs, col = idx - row * cs;
            zz[u] = Z[(size_t)(s + row) * n + c0 + col] * zsc[c0 + col];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int idx = i0 + 32 * u + lane;
          if (idx < cs * m) {
            const int row = idx / cs, col = idx - row * cs;
            buf[(size_t)col * m + row] = zz[u];
          }
        }
      }
      __syncwarp();
      // Column c lives in registers (rows lane + 32 u); the dots against the finished columns are taken four
      // at a time (twenty independent loads in flight instead of one dependent round trip per column - the
      // low end of a dense spectrum can chain 10 - 15 eigenvalues into one cluster).
      constexpr int MAXJ = NW;
      double* dts = gq_dots(Y, np, wid, NW);
      for (int c = 1; c < cs; ++c) {
        double* zc_ = buf + (size_t)c * m;
        double zr[MAXJ];
#pragma unroll
        for (int u = 0; u < MAXJ; ++u) zr[u] = (lane + 32 * u < m) ? zc_[lane + 32 * u] : 0.0;
        double nfirst = 1.0;
        for (int rep = 0; rep < 2; ++rep) {
          for (int p0 = 0; p0 < c; p0 += 64) {  // classical Gram-Schmidt: all dots of a batch from the same vector
            const int pe = min(c, p0 + 64);
            for (int p = p0; p < pe; p += 4) {
              double d4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (p + q < pe) {  // warp-uniform
                  const double* zp = buf + (size_t)(p + q) * m;
#pragma unroll
                  for (int u = 0; u < MAXJ; ++u)
                    if (lane + 32 * u < m) d4[q] = fma(zp[lane + 32 * u], zr[u], d4[q]);
                }
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) d4[q] += __shfl_xor_sync(MOP_FULL_MASK, d4[q], o);
              }
              if (lane < 4 && p + lane < pe) dts[p + lane - p0] = lane == 0 ? d4[0] : (lane == 1 ? d4[1] : (lane == 2 ? d4[2] : d4[3]));
            }
            __syncwarp();
            for (int p = p0; p < pe; p += 4) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (p + q < pe) {  // warp-uniform
                  const double* zp = buf + (size_t)(p + q) * m;
                  const double dq = dts[p + q - p0];
#pragma unroll
                  for (int u = 0; u < MAXJ; ++u)
                    if (lane + 32 * u < m) zr[u] = fma(-dq, zp[lane + 32 * u], zr[u]);
                }
              }
            }
            __syncwarp();
          }
          double nn = 0.0;
#pragma unroll
          for (int u = 0; u < MAXJ; ++u) nn = fma(zr[u], zr[u], nn);
          nn = sqrt(warp_sum(nn));
          if (rep == 0) nfirst = nn;
          if (!(nn > 1e-2)) {
            if (lane == 0) s_fallback = 1;  // vector (nearly) inside the span of its cluster
          }
          const double sc = nn > 0.0 ? 1.0 / nn : 0.0;
#pragma unroll
          for (int u = 0; u < MAXJ; ++u) zr[u] *= sc;
          if (rep == 0 && nfirst > 0.7) break;  // "twice is enough" only when needed
        }
#pragma unroll
        for (int u = 0; u < MAXJ; ++u)
          if (lane + 32 * u < m) zc_[lane + 32 * u] = zr[u];
        __syncwarp();
      }
      for (int c = 1; c < cs; ++c) {
        for (int k = lane; k < m; k += 32) Z[(size_t)(s + k) * n + c0 + c] = buf[(size_t)c * m + k];
        if (lane == 0) zsc[c0 + c] = 1.0;
      }
    }
    __syncthreads();
  } else {
    // zero (or non-finite) matrix: spectrum 0 (identity vectors)
    if (tid < n) {
      lam[tid] = 0.0;
      blk_s[tid] = tid;
      blk_e[tid] = tid + 1;
      for (int k = 0; k < n; ++k) Z[(size_t)k * n + tid] = (k == tid) ? 1.0 : 0.0;
    }
    __syncthreads();
    SP_MARK();
    SP_MARK();
    SP_MARK();
  }
  SP_MARK();  // 3: cluster re-orthogonalisation
  if (s_fallback) {  // robust path will redo this structure; leave state untouched
    if (tid == 0 && a.status) a.status[b] = st_in | MOP_ST_EIG_FALLBACK;
    return;
  }

  // ---- ascending order over all blocks ---------------------------------------------------------
  if (tid < n) {
    const double li = lam[tid];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (lam[j] < li) || (lam[j] == li && j < tid);
    rank[tid] = r;
    inv[r] = tid;
  }
  __syncthreads();
  const double tnorm = zero_t ? 0.0 : s_tnorm;
  double* evals = a.evals ? a.evals + (size_t)b * n : nullptr;

  // ---- gamma = Z^T (Q^T gp), step in the eigenbasis ----------------------------------------------
  double* lam_s = Y;
  double* gam_s = Y + np;
  RfoArrays R = rfo_carve(Y + 2 * np, n);
  double pg = 0.0;
  for (int i = tid; i < n; i += THREADS) {
    const double g = a.Bg[(size_t)b * n + i];
    pg = fma(g, g, pg);
  }
  const double gnorm_raw = sqrt(block_sum(pg, s_red));
  // Columns the cluster phase rewrote (members 1.. of a cluster) and the identity fallback re-read Z; every other
  // column already has its dot product from the twisted sweep (one pass over Z less: 184 MB per 1024 structures).
  if (tid < n && !zero_t && cl_s[tid] == tid) {
    const int r = rank[tid];
    lam_s[r] = identity ? 1.0 : lam[tid] * tnorm;
    gam_s[r] = gdot * zsc[tid];
  } else if (tid < n) {
    const int i = tid, s = blk_s[i], t = blk_e[i];
    double acc0 = 0.0, acc1 = 0.0;
    int k = s;
    for (; k + 7 < t; k += 8) {
      double z[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) z[u] = Z[(size_t)(k + u) * n + i];
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        acc0 = fma(z[u], gq[k + u], acc0);
        acc1 = fma(z[u + 1], gq[k + u + 1], acc1);
      }
    }
    for (; k < t; ++k) acc0 = fma(Z[(size_t)k * n + i], gq[k], acc0);
    const int r = rank[i];
    lam_s[r] = identity ? 1.0 : lam[i] * tnorm;
    gam_s[r] = (acc0 + acc1) * zsc[i];
  }
  __syncthreads();
  if (evals && tid < n) evals[tid] = lam_s[tid];
  double* stp = a.state + (size_t)b * MOP_RSIRFO_STATE;
  int flags = identity ? MOP_ST_EIG_NONFINITE : 0;
  flags |= rfo_core<NW>(n, a.saddle_order, a.neb_mode, a.tmin, a.tmax, lam_s, gam_s, identity, gnorm_raw,
                    a.Be ? a.Be[b] : 0.0, stp, R, a.pred ? a.pred + b : nullptr);
  SP_MARK();  // 4: gamma + eigenbasis RFO

  // ---- y = Z c: warps over rows (coalesced), four rows per pass -------------------------------------
  double* cz = e2;
  if (tid < n) cz[tid] = R.coef[rank[tid]] * zsc[tid];
  __syncthreads();
  double* y = gq;
  {
    constexpr int MAXJ = NW;
    double cc[MAXJ];
#pragma unroll
    for (int u = 0; u < MAXJ; ++u) cc[u] = (lane + 32 * u < n) ? cz[lane + 32 * u] : 0.0;
    for (int k0 = 4 * wid; k0 < n; k0 += 4 * NW) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int u = 0; u < MAXJ; ++u) {
        const int i = lane + 32 * u;
        if (i < n) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (k0 + r < n) acc[r] = fma(Z[(size_t)(k0 + r) * n + i], cc[u], acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = warp_sum(acc[r]);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (k0 + r < n) y[k0 + r] = acc[r];
      }
    }
  }
  __syncthreads();
  SP_MARK();  // 5: Z c

  // ---- step = Q y = H_0 ... H_{n-3} y: one warp, two reflectors per reduction round.  The reflector rows
  // stream from L2 / HBM into a six-stage shared-memory ring (cp.async, five rounds ahead; the ring
  // aliases the RFO arrays, dead by now), so a round costs its reductions, not a memory round trip.
  // With u_i = v_i^T y and G = v_0^T v_1 taken from the same y: c_1 = t_1 u_1, c_0 = t_0 (u_0 - c_1 G).
  if (wid == 0) {
    if (!trivial) {
      constexpr int NS = 6;
      double* ring = Y;  // [NS][2][np]
      const int nround = (n - 3 >= 1) ? (n - 3 + 1) / 2 : 0;  // rounds r = 0.. handle k = n-3-2r >= 1
      auto issue = [&](int r) {
        if (r < nround) {
          const int k = n - 3 - 2 * r;
          const double* v0 = Vh + (size_t)(k - 1) * n;
          const double* v1 = Vh + (size_t)k * n;
          double* r0 = ring + (size_t)(r % NS) * 2 * np;
          double* r1 = r0 + np;
          for (int j = k + lane; j < n; j += 32) {
            cp_async8(r0 + j, v0 + j);
            cp_async8(r1 + j, v1 + j);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      for (int r = 0; r < NS - 1; ++r) issue(r);
      for (int r = 0; r < nround; ++r) {
        issue(r + NS - 1);  // overwrites the stage read in round r-1 (ordered by the __syncwarp below)
        asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
        __syncwarp();
        const int k = n - 3 - 2 * r;
        const double* r0 = ring + (size_t)(r % NS) * 2 * np;
        const double* r1 = r0 + np;
        double u0 = 0.0, u1 = 0.0, g01 = 0.0;
        for (int j = k + lane; j < n; j += 32) {
          const double a0 = (j == k) ? 1.0 : r0[j];
          const double a1 = (j == k) ? 0.0 : ((j == k + 1) ? 1.0 : r1[j]);
          const double yj = y[j];
          u0 = fma(a0, yj, u0);
          u1 = fma(a1, yj, u1);
          g01 = fma(a0, a1, g01);
        }
        u0 = warp_sum(u0);
        u1 = warp_sum(u1);
        g01 = warp_sum(g01);
        const double c1 = tau[k] * u1;
        const double c0 = tau[k - 1] * (u0 - c1 * g01);
        for (int j = k + lane; j < n; j += 32) {
          const double a0 = (j == k) ? 1.0 : r0[j];
          const double a1 = (j == k) ? 0.0 : ((j == k + 1) ? 1.0 : r1[j]);
          y[j] -= c0 * a0 + c1 * a1;
        }
        __syncwarp();
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      for (int k = n - 3 - 2 * nround; k >= 0; --k) {
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
  SP_MARK();  // 6: back-transform
  if (tid == 0 && a.status) {
    const int keep = st_in & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV |
                              MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF);
    a.status[b] = keep | flags;
  }
#undef SP_MARK
}

}  // namespace mop

static long long* g_sp_dbg = nullptr;
// diagnostics (include/../csrc/mop_private.h): device buffer [B][16] receiving per-phase clock counts of the next launches
extern "C" int mop_priv_spectrum_timing(void* buf) {
  g_sp_dbg = (long long*)buf;
  return MOP_OK;
}

template <int NW>
static int launch_spectrum(int B, const mop::SpArgs& a, cudaStream_t stream) {
  const int np = (a.n + 3) & ~3;
  const size_t smem = sizeof(double) * (5 * (size_t)np + mop::sp_y_doubles(a.n, NW)) + 5 * (size_t)np;
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_spectrum_step<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_spectrum_step<NW><<<B, 32 * NW, smem, stream>>>(a);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

int mop_spectrum_step_supported(int n) { return n >= 3 && n <= mop::SP_MAX_N; }

// T, tau, Q^T gp, flag and the reflector rows Vh come from mop_launch_tridiag_packed; Z and Dm are
// [B][n][n] scratch slabs
int mop_launch_spectrum_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                             const double* Vh, double* Z, double* Dm, const double* pd, const double* pe,
                             const double* ptau, const double* pgq, const int* pflag, const double* Bg,
                             const double* Be, double* state, double* move, double* evals_out, double* pred,
                             int32_t* status, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  mop::SpArgs a{};
  a.n = n;
  a.Z = Z;
  a.Dm = Dm;
  a.Vh = Vh;
  a.pf_d = pd;
  a.pf_e = pe;
  a.pf_tau = ptau;
  a.pf_gq = pgq;
  a.pf_flag = pflag;
  a.evals = evals_out;
  a.status = status;
  a.Bg = Bg;
  a.Be = Be;
  a.state = state;
  a.move = move;
  a.pred = pred;
  a.saddle_order = saddle_order;
  a.neb_mode = neb_mode;
  a.tmin = tmin;
  a.tmax = tmax;
  a.dbg = g_sp_dbg;
  switch ((n + 31) / 32) {
    case 1: return launch_spectrum<1>(B, a, stream);
    case 2: return launch_spectrum<2>(B, a, stream);
    case 3: return launch_spectrum<3>(B, a, stream);
    case 4: return launch_spectrum<4>(B, a, stream);
    default: return launch_spectrum<5>(B, a, stream);
  }
}

// ---- the shared-memory spectral path, 3 <= n <= 160: blocked tridiagonalisation + spectrum / step ------------------
int mop_launch_tridiag_blk(int B, int n, const double* A, const double* gp, double* Vh, double* dd, double* ee,
                           double* tau, double* gq, int* flag, double* hand, cudaStream_t stream);

int mop_tridiag_supported(int n) { return n >= 3 && n <= mop::SP_MAX_N; }
// Vh | Dm | d, e, tau, gq | flag
size_t mop_tridiag_workspace_bytes(int B, int n) {
  return 2 * sizeof(double) * (size_t)B * n * n + 4 * sizeof(double) * (size_t)B * n + sizeof(int) * (size_t)B + 64;
}

// RS-I-RFO step from an already projected Hessian (mop_rsirfo_spectral_step): k_tridiag_blk, then k_spectrum_step.
// zbuf: a [B][n][n] slab for the eigenvectors of T.
int mop_launch_rsirfo_fused(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                            const double* Hp, const double* gp, const double* Bg, const double* Be,
                            double* state, double* move, double* evals_out, double* pred,
                            int32_t* status, void* work, size_t work_bytes, double* zbuf, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_supported(n) || !zbuf) {
    mop_set_error("spectral RS-I-RFO step: n = %d not supported (3 .. %d)", n, mop::SP_MAX_N);
    return MOP_ERR_UNSUPPORTED;
  }
  if (!work || work_bytes < mop_tridiag_workspace_bytes(B, n)) {
    mop_set_error("spectral RS-I-RFO step: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  double* Vh = (double*)work;
  double* Dm = Vh + (size_t)B * n * n;
  double* pd = Vh + 2 * (size_t)B * n * n;
  double* pe = pd + (size_t)B * n;
  double* pt = pd + 2 * (size_t)B * n;
  double* pg = pd + 3 * (size_t)B * n;
  int* pflag = (int*)(pd + 4 * (size_t)B * n);
  int rc = mop_launch_tridiag_blk(B, n, Hp, gp, Vh, pd, pe, pt, pg, pflag, Dm /* free until the spectrum kernel */, stream);
  if (rc != MOP_OK) return rc;
  return mop_launch_spectrum_step(B, n, saddle_order, neb_mode, tmin, tmax, Vh, zbuf, Dm, pd, pe, pt, pg, pflag, Bg, Be,
                                  state, move, evals_out, pred, status, stream);
}
