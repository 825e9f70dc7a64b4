// Fused quasi-Newton Hessian update, generic-n streaming path.
//
// One CTA per structure.  Pass 1 streams H once for u = H s (and H y for the
// flowchart selector), pass 2 streams it again (L2-resident: the CTA's own
// 8 n^2 bytes) applying  H <- 1/2 (H + H^T) + sum_ab C_ab v_a v_b^T  tile pair by
// tile pair, so every global access is a coalesced 256-byte row segment.
// HBM-bound: algorithmic traffic 16 n^2 bytes per structure (read + write H).
#include "update_coef.cuh"

namespace mop {

constexpr int UPD_THREADS = 256;
constexpr int TILE = 32;

// smem: v[4][np] (s, y, u, r), hy[np], scratch[40], tiles 2 x 32 x 33
__global__ void __launch_bounds__(UPD_THREADS)
k_hessian_update(int n, int method, int mode, int guards, double* __restrict__ Hall,
                 const double* __restrict__ s_all, const double* __restrict__ y_all,
                 const double* __restrict__ x_all, const double* __restrict__ xp_all,
                 const double* __restrict__ g_all, const double* __restrict__ gp_all,
                 const double* __restrict__ state, int state_stride,
                 double* __restrict__ delta_all, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int np = (n + 3) & ~3;
  double* vs = sm;
  double* vy = vs + np;
  double* vu = vy + np;
  double* vr = vu + np;
  double* hy = vr + np;
  double* scratch = hy + np;             // 40
  double* tA = scratch + 40;             // 32 x 33
  double* tB = tA + TILE * (TILE + 1);   // 32 x 33
  __shared__ UpdCoef coef;
  __shared__ int s_apply, s_flags, s_method;

  double* H = Hall + (size_t)b * n * n;
  int st = status ? status[b] : 0;
  st &= ~(MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV | MOP_ST_UPD_TERM_ZEROED |
          MOP_ST_NO_HISTORY);

  // ---- s, y ---------------------------------------------------------------
  const bool from_points = (x_all != nullptr);
  bool have_prev = true;
  if (from_points) {
    have_prev = (xp_all != nullptr) && (gp_all != nullptr) &&
                (state == nullptr || state[(size_t)b * state_stride + MOP_RS_HAVE_PREV] != 0.0);
  }
  if (!have_prev) {
    if (tid == 0 && status) status[b] = st | MOP_ST_NO_HISTORY;
    return;
  }
  for (int i = tid; i < n; i += UPD_THREADS) {
    double s, y;
    if (from_points) {
      s = x_all[(size_t)b * n + i] - xp_all[(size_t)b * n + i];
      y = g_all[(size_t)b * n + i] - gp_all[(size_t)b * n + i];
    } else {
      s = s_all[(size_t)b * n + i];
      y = y_all[(size_t)b * n + i];
    }
    vs[i] = s;
    vy[i] = y;
  }
  __syncthreads();
  double pss = 0, psy = 0, pyy = 0;
  for (int i = tid; i < n; i += UPD_THREADS) {
    pss = fma(vs[i], vs[i], pss);
    psy = fma(vs[i], vy[i], psy);
    pyy = fma(vy[i], vy[i], pyy);
  }
  const double ss = block_sum(pss, scratch);
  double sy = block_sum(psy, scratch);
  const double yy = block_sum(pyy, scratch);

  if (guards) {  // 1: RSIRFO.update_hessian (rsirfo.py:1326,1333); 2: EnhancedRSPRFO (rsprfo.py:1210)
    int skip = 0;
    if (sqrt(ss) < 1e-10 || sqrt(yy) < 1e-10) skip = MOP_ST_UPD_SKIP_SMALL;
    else if (guards == 1 && sy <= 0.0) skip = MOP_ST_UPD_SKIP_CURV;
    if (skip) {
      if (tid == 0 && status) status[b] = st | skip;
      return;
    }
  }

  // ---- Powell damping (replaces y) -----------------------------------------
  if (method_has_dd(method)) {
    bool active = true;
    if (method == MOP_UPD_BLOCK_BFGS_DD && !(sqrt(ss) > 1e-8)) active = false;  // rank guard first
    if (active) {
      const double th = dd_theta(ss, sy, method_dd_thr(method));
      if (th != 1.0) {
        for (int i = tid; i < n; i += UPD_THREADS) vy[i] = th * vy[i] + (1.0 - th) * vs[i];
        __syncthreads();
        double p = 0;
        for (int i = tid; i < n; i += UPD_THREADS) p = fma(vs[i], vy[i], p);
        sy = block_sum(p, scratch);
      }
    }
  }

  // ---- u = H s (and H y for the flowchart) ----------------------------------
  block_matvec(H, n, n, vs, vu);
  if (method == MOP_UPD_FLOWCHART) block_matvec(H, n, n, vy, hy);
  __syncthreads();
  int m = method;
  if (method == MOP_UPD_FLOWCHART) {
    double pzz = 0, pzs = 0;
    for (int i = tid; i < n; i += UPD_THREADS) {
      const double z = vy[i] - hy[i];
      pzz = fma(z, z, pzz);
      pzs = fma(z, vs[i], pzs);
    }
    const double zz = block_sum(pzz, scratch);
    const double zs = block_sum(pzs, scratch);
    m = flowchart_select(ss, yy, sy, zz, zs);
  }
  double psu = 0, prs = 0, prr = 0;
  for (int i = tid; i < n; i += UPD_THREADS) {
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
  if (tid == 0) {
    update_coefficients(m, q, coef);
    s_flags = coef.flags;
  }
  __syncthreads();

  // ---- apply, tile pair by tile pair ----------------------------------------
  const int T = (n + TILE - 1) / TILE;
  double* D = (mode == 0) ? delta_all + (size_t)b * n * n : nullptr;
  for (int I = 0; I < T; ++I) {
    for (int J = I; J < T; ++J) {
      const int i0 = I * TILE, j0 = J * TILE;
      if (mode == 1) {
        for (int e = tid; e < TILE * TILE; e += UPD_THREADS) {
          const int r = e >> 5, c = e & 31;
          const int gi = i0 + r, gj = j0 + c;
          tA[r * (TILE + 1) + c] = (gi < n && gj < n) ? H[(size_t)gi * n + gj] : 0.0;
          const int hi = j0 + r, hj = i0 + c;
          tB[r * (TILE + 1) + c] = (hi < n && hj < n) ? H[(size_t)hi * n + hj] : 0.0;
        }
        __syncthreads();
      }
      for (int e = tid; e < TILE * TILE; e += UPD_THREADS) {
        const int r = e >> 5, c = e & 31;
        {  // element (i0 + r, j0 + c)
          const int gi = i0 + r, gj = j0 + c;
          if (gi < n && gj < n) {
            const double vi[4] = {vs[gi], vy[gi], vu[gi], vr[gi]};
            const double vj[4] = {vs[gj], vy[gj], vu[gj], vr[gj]};
            const double d = 0.5 * (coef_delta(coef, vi, vj) + coef_delta(coef, vj, vi));
            if (mode == 1)
              H[(size_t)gi * n + gj] = 0.5 * (tA[r * (TILE + 1) + c] + tB[c * (TILE + 1) + r]) + d;
            else
              D[(size_t)gi * n + gj] = d;
          }
        }
        if (J != I) {  // mirrored element (j0 + r, i0 + c)
          const int gi = j0 + r, gj = i0 + c;
          if (gi < n && gj < n) {
            const double vi[4] = {vs[gi], vy[gi], vu[gi], vr[gi]};
            const double vj[4] = {vs[gj], vy[gj], vu[gj], vr[gj]};
            const double d = 0.5 * (coef_delta(coef, vi, vj) + coef_delta(coef, vj, vi));
            if (mode == 1)
              H[(size_t)gi * n + gj] = 0.5 * (tB[r * (TILE + 1) + c] + tA[c * (TILE + 1) + r]) + d;
            else
              D[(size_t)gi * n + gj] = d;
          }
        }
      }
      if (mode == 1) __syncthreads();
    }
  }
  if (tid == 0 && status) status[b] = st | MOP_ST_UPDATED | s_flags;
}


// ------------------------------------------------------------------------------------------------
// Multi-CTA path (used inside the fused optimizer steps, where scratch is available): the
// one-CTA-per-structure kernel above is latency-bound (two dependent passes over one structure's
// Hessian by 256 threads reach ~20 % of the HBM bandwidth) and leaves the GPU idle for small batches.
//   k_upd_matvec   grid (ceil(n/32), B): u = H s (and H y for the flowchart), four rows per warp in flight
//   k_upd_scalars  grid B, one warp's worth of arithmetic: guards, damping, scalars, coefficient matrix
//   k_upd_apply    grid (tile pairs, B): H <- 1/2 (H + H^T) + delta for one 32 x 32 tile pair
// Scratch per structure: s, y (damped), u, H y  (4 n doubles) + 24 doubles of coefficients / flags.
constexpr int UPD_SCR_HDR = 24;
__host__ __device__ inline size_t upd_scratch_doubles(int n) { return 4 * (size_t)n + UPD_SCR_HDR; }

__device__ __forceinline__ bool upd_have_prev(const double* x_all, const double* xp_all, const double* gp_all,
                                              const double* state, int state_stride, int b) {
  if (x_all == nullptr) return true;
  return (xp_all != nullptr) && (gp_all != nullptr) &&
         (state == nullptr || state[(size_t)b * state_stride + MOP_RS_HAVE_PREV] != 0.0);
}

__global__ void __launch_bounds__(256)
k_upd_matvec(int n, int method, const double* __restrict__ Hall, const double* __restrict__ s_all,
             const double* __restrict__ y_all, const double* __restrict__ x_all, const double* __restrict__ xp_all,
             const double* __restrict__ g_all, const double* __restrict__ gp_all, const double* __restrict__ state,
             int state_stride, double* __restrict__ scratch) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (!upd_have_prev(x_all, xp_all, gp_all, state, state_stride, b)) return;
  const int np = (n + 3) & ~3;
  double* vs = sm;
  double* vy = sm + np;
  const bool flow = method == MOP_UPD_FLOWCHART;
  for (int i = tid; i < n; i += 256) {
    const size_t e = (size_t)b * n + i;
    vs[i] = x_all ? x_all[e] - xp_all[e] : s_all[e];
    if (flow) vy[i] = x_all ? g_all[e] - gp_all[e] : y_all[e];
  }
  __syncthreads();
  const double* H = Hall + (size_t)b * n * n;
  double* scr = scratch + (size_t)b * upd_scratch_doubles(n);
  double* u = scr + UPD_SCR_HDR + 2 * (size_t)n;
  double* hy = u + n;
  const int r0 = blockIdx.x * 32 + w * 4;
  double au[4] = {0, 0, 0, 0}, ah[4] = {0, 0, 0, 0};
  for (int j = lane; j < n; j += 32) {
    const double sj = vs[j];
    const double yj = flow ? vy[j] : 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + q;
      if (r < n) {
        const double h = H[(size_t)r * n + j];
        au[q] = fma(h, sj, au[q]);
        if (flow) ah[q] = fma(h, yj, ah[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double a = warp_sum(au[q]);
    const double hsum = flow ? warp_sum(ah[q]) : 0.0;
    if (lane == 0 && r0 + q < n) {
      u[r0 + q] = a;
      if (flow) hy[r0 + q] = hsum;
    }
  }
}

// scratch header: [0..15] coefficient matrix, [16] flags, [17] apply (0 / 1)
__global__ void __launch_bounds__(128)
k_upd_scalars(int n, int method, int guards, const double* __restrict__ s_all, const double* __restrict__ y_all,
              const double* __restrict__ x_all, const double* __restrict__ xp_all, const double* __restrict__ g_all,
              const double* __restrict__ gp_all, const double* __restrict__ state, int state_stride,
              double* __restrict__ scratch, int32_t* __restrict__ status) {
  __shared__ double red[40];
  const int b = blockIdx.x, tid = threadIdx.x;
  double* scr = scratch + (size_t)b * upd_scratch_doubles(n);
  double* vs = scr + UPD_SCR_HDR;
  double* vy = vs + n;
  const double* vu = vy + n;
  const double* hy = vu + n;
  int st = status ? status[b] : 0;
  st &= ~(MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV | MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY);
  if (!upd_have_prev(x_all, xp_all, gp_all, state, state_stride, b)) {
    if (tid == 0) {
      scr[17] = 0.0;
      if (status) status[b] = st | MOP_ST_NO_HISTORY;
    }
    return;
  }
  double pss = 0, psy = 0, pyy = 0;
  for (int i = tid; i < n; i += 128) {
    const size_t e = (size_t)b * n + i;
    const double s = x_all ? x_all[e] - xp_all[e] : s_all[e];
    const double y = x_all ? g_all[e] - gp_all[e] : y_all[e];
    vs[i] = s;
    vy[i] = y;
    pss = fma(s, s, pss);
    psy = fma(s, y, psy);
    pyy = fma(y, y, pyy);
  }
  const double ss = block_sum(pss, red);
  double sy = block_sum(psy, red);
  const double yy = block_sum(pyy, red);
  if (guards) {
    int skip = 0;
    if (sqrt(ss) < 1e-10 || sqrt(yy) < 1e-10) skip = MOP_ST_UPD_SKIP_SMALL;
    else if (guards == 1 && sy <= 0.0) skip = MOP_ST_UPD_SKIP_CURV;
    if (skip) {
      if (tid == 0) {
        scr[17] = 0.0;
        if (status) status[b] = st | skip;
      }
      return;
    }
  }
  if (method_has_dd(method)) {
    bool active = true;
    if (method == MOP_UPD_BLOCK_BFGS_DD && !(sqrt(ss) > 1e-8)) active = false;
    if (active) {
      const double th = dd_theta(ss, sy, method_dd_thr(method));
      if (th != 1.0) {
        double p = 0;
        for (int i = tid; i < n; i += 128) {
          const double yt = th * vy[i] + (1.0 - th) * vs[i];
          vy[i] = yt;
          p = fma(vs[i], yt, p);
        }
        sy = block_sum(p, red);
      }
    }
  }
  int m = method;
  if (method == MOP_UPD_FLOWCHART) {
    double pzz = 0, pzs = 0;
    for (int i = tid; i < n; i += 128) {
      const double z = vy[i] - hy[i];
      pzz = fma(z, z, pzz);
      pzs = fma(z, vs[i], pzs);
    }
    const double zz = block_sum(pzz, red);
    const double zs = block_sum(pzs, red);
    m = flowchart_select(ss, yy, sy, zz, zs);
  }
  double psu = 0, prs = 0, prr = 0;
  for (int i = tid; i < n; i += 128) {
    const double r = vy[i] - vu[i];
    psu = fma(vs[i], vu[i], psu);
    prs = fma(r, vs[i], prs);
    prr = fma(r, r, prr);
  }
  UpdScalars q;
  q.ss = ss;
  q.sy = sy;
  q.su = block_sum(psu, red);
  q.rs = block_sum(prs, red);
  q.rr = block_sum(prr, red);
  if (tid == 0) {
    UpdCoef coef;
    update_coefficients(m, q, coef);
    for (int a = 0; a < 4; ++a)
      for (int c = 0; c < 4; ++c) scr[4 * a + c] = coef.c[a][c];
    scr[16] = (double)coef.flags;
    scr[17] = 1.0;
    if (status) status[b] = st | MOP_ST_UPDATED | coef.flags;
  }
}

// CTA (I, b) walks the tile pairs (I, J >= I) of one structure; the tiles of pair J+1 are loaded into
// registers while pair J is computed from shared memory.  With Cs = 1/2 (C + C^T) the symmetrised update of
// element (i, j) is v_i^T Cs v_j = sum_a v_i[a] t_j[a], t_j = Cs v_j staged once per column (4 FMAs per
// element); every value is computed once and its mirror image written from shared memory, so the result is
// exactly symmetric and both global writes are coalesced.
__global__ void __launch_bounds__(UPD_THREADS)
k_upd_apply(int n, int T, int mode, double* __restrict__ Hall, const double* __restrict__ scratch,
            double* __restrict__ delta_all) {
  __shared__ double tA[TILE * (TILE + 1)], tB[TILE * (TILE + 1)];
  __shared__ double vi_[4][TILE], tj_[4][TILE];
  __shared__ double cs[4][4];
  const int b = blockIdx.y, tid = threadIdx.x, I = blockIdx.x;
  const double* scr = scratch + (size_t)b * upd_scratch_doubles(n);
  if (scr[17] == 0.0) return;
  const int i0 = I * TILE;
  const double* vs = scr + UPD_SCR_HDR;
  auto vec = [&](int a, int gidx) -> double {  // a: s, y, u, r = y - u
    if (gidx >= n) return 0.0;
    return a < 3 ? vs[(size_t)a * n + gidx] : vs[(size_t)n + gidx] - vs[2 * (size_t)n + gidx];
  };
  if (tid < 16) cs[tid >> 2][tid & 3] = 0.5 * (scr[tid] + scr[4 * (tid & 3) + (tid >> 2)]);
  if (tid < 4 * TILE) vi_[tid / TILE][tid % TILE] = vec(tid / TILE, i0 + tid % TILE);
  double* H = Hall + (size_t)b * n * n;
  double* D = (mode == 0) ? delta_all + (size_t)b * n * n : nullptr;
  double* O = mode == 1 ? H : D;
  constexpr int EPT = TILE * TILE / UPD_THREADS;  // elements per thread and tile
  double ra[EPT], rb[EPT];
  auto load = [&](int J) {
    const int j0 = J * TILE;
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * UPD_THREADS, r = e >> 5, c = e & 31;
      const int gi = i0 + r, gj = j0 + c, hi = j0 + r, hj = i0 + c;
      ra[u] = (gi < n && gj < n) ? H[(size_t)gi * n + gj] : 0.0;
      rb[u] = (hi < n && hj < n) ? H[(size_t)hi * n + hj] : 0.0;
    }
  };
  if (mode == 1) load(I);
  for (int J = I; J < T; ++J) {
    const int j0 = J * TILE;
    __syncthreads();  // previous pair is done with the tiles and t_j
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * UPD_THREADS, r = e >> 5, c = e & 31;
      tA[r * (TILE + 1) + c] = mode == 1 ? ra[u] : 0.0;
      tB[r * (TILE + 1) + c] = mode == 1 ? rb[u] : 0.0;
    }
    if (tid < TILE) {  // t_j = Cs v_j for column j0 + tid
      double vj[4];
      for (int a2 = 0; a2 < 4; ++a2) vj[a2] = vec(a2, j0 + tid);
      for (int a2 = 0; a2 < 4; ++a2) {
        double t = 0.0;
        for (int b2 = 0; b2 < 4; ++b2) t = fma(cs[a2][b2], vj[b2], t);
        tj_[a2][tid] = t;
      }
    }
    __syncthreads();
    if (mode == 1 && J + 1 < T) load(J + 1);
    // value of element (i0 + r, j0 + c), kept in tA[r][c]; the diagonal tile computes r <= c only
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * UPD_THREADS, r = e >> 5, c = e & 31;
      if (J == I && r > c) continue;
      double d = 0.0;
#pragma unroll
      for (int a2 = 0; a2 < 4; ++a2) d = fma(vi_[a2][r], tj_[a2][c], d);
      const double val = 0.5 * (tA[r * (TILE + 1) + c] + tB[c * (TILE + 1) + r]) + d;
      tA[r * (TILE + 1) + c] = val;
      if (J == I) tA[c * (TILE + 1) + r] = val;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
      const int e = tid + u * UPD_THREADS, r = e >> 5, c = e & 31;
      {
        const int gi = i0 + r, gj = j0 + c;
        if (gi < n && gj < n) O[(size_t)gi * n + gj] = tA[r * (TILE + 1) + c];
      }
      if (J != I) {  // mirror image, row j0 + r
        const int gi = j0 + r, gj = i0 + c;
        if (gi < n && gj < n) O[(size_t)gi * n + gj] = tA[c * (TILE + 1) + r];
      }
    }
  }
}

}  // namespace mop

static size_t upd_smem_bytes(int n) {
  const int np = (n + 3) & ~3;
  return sizeof(double) * (5 * (size_t)np + 40 + 2 * mop::TILE * (mop::TILE + 1));
}

// Internal launcher shared with mop_rsirfo_step (s, y formed from the points).
int mop_launch_hessian_update(int B, int n, int method, int mode, int guards, double* H,
                              const double* s, const double* y, const double* x, const double* xp,
                              const double* g, const double* gp, const double* state, int state_stride,
                              double* delta_out, int32_t* status, cudaStream_t stream) {
  if (method == MOP_UPD_PCFD_BOFILL) {
    mop_set_error("pcfd_bofill (O(n^4) null-space perturbation) is not implemented on the device");
    return MOP_ERR_UNSUPPORTED;
  }
  if (method < 0 || method > MOP_UPD_MSP) {
    mop_set_error("unknown Hessian update method id %d", method);
    return MOP_ERR_INVALID;
  }
  if (method == MOP_UPD_NONE || B == 0) return MOP_OK;
  const size_t smem = upd_smem_bytes(n);
  if (smem > 200 * 1024) {
    mop_set_error("n = %d too large for the update kernel's vector staging", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_hessian_update,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_hessian_update<<<B, mop::UPD_THREADS, smem, stream>>>(
      n, method, mode, guards, H, s, y, x, xp, g, gp, state, state_stride, delta_out, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

size_t mop_hessian_update_scratch_bytes(int B, int n) { return sizeof(double) * (size_t)B * mop::upd_scratch_doubles(n); }

// Same contract as mop_launch_hessian_update, three multi-CTA kernels, `scratch` from the caller.
int mop_launch_hessian_update_split(int B, int n, int method, int mode, int guards, double* H, const double* s,
                                    const double* y, const double* x, const double* xp, const double* g,
                                    const double* gp, const double* state, int state_stride, double* delta_out,
                                    int32_t* status, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  if (method == MOP_UPD_PCFD_BOFILL) {
    mop_set_error("pcfd_bofill (O(n^4) null-space perturbation) is not implemented on the device");
    return MOP_ERR_UNSUPPORTED;
  }
  if (method < 0 || method > MOP_UPD_MSP) {
    mop_set_error("unknown Hessian update method id %d", method);
    return MOP_ERR_INVALID;
  }
  if (method == MOP_UPD_NONE || B == 0) return MOP_OK;
  if (!scratch || scratch_bytes < mop_hessian_update_scratch_bytes(B, n)) {
    mop_set_error("hessian update: scratch too small");
    return MOP_ERR_WORKSPACE;
  }
  const int np = (n + 3) & ~3;
  double* scr = (double*)scratch;
  {
    dim3 grid((n + 31) / 32, B);
    mop::k_upd_matvec<<<grid, 256, sizeof(double) * 2 * (size_t)np, stream>>>(n, method, H, s, y, x, xp, g, gp, state,
                                                                            state_stride, scr);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  mop::k_upd_scalars<<<B, 128, 0, stream>>>(n, method, guards, s, y, x, xp, g, gp, state, state_stride, scr, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  {
    const int T = (n + mop::TILE - 1) / mop::TILE;
    dim3 grid(T, B);
    mop::k_upd_apply<<<grid, mop::UPD_THREADS, 0, stream>>>(n, T, mode, H, scr, delta_out);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  return MOP_OK;
}

extern "C" size_t mop_hessian_update_workspace_bytes(int B, int n) {
  return (B <= 0 || n <= 0) ? 0 : mop_hessian_update_scratch_bytes(B, n);
}

extern "C" int mop_hessian_update(int B, int n, int method, int mode, int rsirfo_guards, double* H,
                                  const double* s, const double* y, double* delta_out,
                                  int32_t* status, void* work, size_t work_bytes, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_hessian_update: B >= 0 and n > 0 required");
  MOP_REQUIRE(H && s && y, "mop_hessian_update: H, s, y must be device pointers");
  MOP_REQUIRE(mode == 0 || mode == 1, "mop_hessian_update: mode must be 0 (delta) or 1 (in place)");
  MOP_REQUIRE(mode == 1 || delta_out, "mop_hessian_update: delta_out required in mode 0");
  if (work)
    return mop_launch_hessian_update_split(B, n, method, mode, rsirfo_guards, H, s, y, nullptr, nullptr, nullptr,
                                           nullptr, nullptr, 0, delta_out, status, work, work_bytes,
                                           (cudaStream_t)stream);
  return mop_launch_hessian_update(B, n, method, mode, rsirfo_guards, H, s, y, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, 0, delta_out, status,
                                   (cudaStream_t)stream);
}
