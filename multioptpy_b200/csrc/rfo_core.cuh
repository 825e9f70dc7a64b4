// RS-I-RFO step in the eigenbasis, shared by the generic
// (rfo_step.cu) and the fused shared-memory (eigh_tridiag.cu) kernels.
//
// Input: ascending spectrum lam[n] of the TR/ROT-projected Hessian Hp and the
// components gam[n] = V^T gp of the projected gradient.  Output: coef[n], the RFO
// step expressed in the same eigenbasis (step = sum_k coef[k] v_k), the predicted
// energy change and the updated per-structure RSIRFO state.  Follows
// Optimizer/rsirfo.py:360-490; the second eigendecomposition of the image Hessian
// (rsirfo.py:423-427) is derived analytically: H* = V diag(lam') V^T with the first
// `saddle_order` eigenvalues of |lam| > 1e-10 sign-flipped (zeroed in NEB mode) and
// the matching gradient components negated (zeroed).
#pragma once
#include "rfo_secular.cuh"

namespace mop {

// check_hessian_conditioning (rsirfo.py:492-551) -> ill-conditioned?  Whole CTA.
__device__ __forceinline__ bool spectrum_ill_conditioned(const double* lam, int n, double* scratch) {
  if (n < 2) return false;
  double mx = 0.0, mn = INFINITY, cnt = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fabs(lam[i]);
    if (a > 1e-10) {
      mx = fmax(mx, a);
      mn = fmin(mn, a);
      cnt += 1.0;
    }
  }
  cnt = block_sum(cnt, scratch);
  mx = block_max(mx, scratch);
  mn = -block_max(-mn, scratch);
  if (cnt < 2.0) return true;
  if (mn < 1e-15) return true;
  return (mx / mn) > 1e8;
}

// adjust_trust_radius(+_adaptive), rsirfo.py:660-887
__device__ __forceinline__ double adjust_trust(double trust, double actual, double predicted,
                                               double min_eig, double gnorm, int saddle_order,
                                               double tmin, double tmax) {
  if (fabs(predicted) < 1e-10) return trust;
  const double ratio = actual / predicted;
  if (gnorm < 1e-2) {
    const double a = fabs(min_eig);
    double cf = a > 1e-6 ? fmin(2.5, 1.0 / fmax(a, 0.1)) : 1.5;
    if (saddle_order > 0 && min_eig < -1e-6) cf *= 0.8;
    if (ratio > 0.75) trust = fmin(trust * fmin(1.5 * cf, 2.5), tmax);
    else if (ratio > 0.5) trust = fmin(trust * fmin(1.1 * cf, 1.5), tmax);
    else if (ratio > 0.25) { if (cf > 1.2) trust = fmin(trust * 1.05, tmax); }
    else if (ratio > 0.1) trust = fmax(trust * 0.5, tmin);
    else trust = fmax(trust * 0.25, tmin);
    return fmin(fmax(trust, tmin), tmax);
  }
  if (ratio > 0.75) trust = fmin(trust * 1.2, tmax);
  else if (ratio < 0.25) trust = fmax(trust * 0.5, tmin);
  return trust;
}

// Shared-memory arrays the core needs: 9 doubles x np + 2 ints x np + 40 scratch.
struct RfoArrays {
  double *lams, *gams, *lamk, *gamk, *stepk, *w1, *w2, *w3, *coef;
  int *ord, *ordk;
  double* scratch;
};
__host__ __device__ inline size_t rfo_core_smem_bytes(int n) {
  const int np = (n + 3) & ~3;
  return sizeof(double) * (9 * (size_t)np + 40) + sizeof(int) * 2 * (size_t)np;
}
__device__ __forceinline__ RfoArrays rfo_carve(double* base, int n) {
  const int np = (n + 3) & ~3;
  RfoArrays a;
  a.lams = base;
  a.gams = a.lams + np;
  a.lamk = a.gams + np;
  a.gamk = a.lamk + np;
  a.stepk = a.gamk + np;
  a.w1 = a.stepk + np;
  a.w2 = a.w1 + np;
  a.w3 = a.w2 + np;
  a.coef = a.w3 + np;
  a.scratch = a.coef + np;
  a.ord = (int*)(a.scratch + 40);
  a.ordk = a.ord + np;
  return a;
}

// lam/gam: [n] in shared memory (lam may be modified by the level-shift emulation).
// identity: the spectrum was non-finite and has been replaced by (1, gp) by the caller.
// Returns status flags; writes R.coef (indexed like lam), *pred_out, state.
// MAXJ > 0 (n <= 32 MAXJ): the register-resident secular solver (rfo_secular.cuh), bracket probes on warps 1 and 2.
template <int MAXJ = 0>
static __device__ int rfo_core(int n, int saddle_order, int neb_mode, double tmin, double tmax,
                        double* lam, const double* gam, bool identity, double gnorm_raw, double Be,
                        double* st, const RfoArrays& R, double* pred_out) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
  __shared__ int s_k, s_flags;
  __shared__ double s_trust, s_probe[2];
  int flags = 0;

  // level shift emulation (rsirfo.py:602-631): eigh(H + 1e-5 I) - 1e-5
  if (!identity && spectrum_ill_conditioned(lam, n, R.scratch)) {
    flags |= MOP_ST_LEVEL_SHIFT;
    for (int i = tid; i < n; i += nt) lam[i] = __dadd_rn(__dadd_rn(lam[i], 1e-5), -1e-5);
  }
  __syncthreads();

  // inner trust radius bookkeeping (rsirfo.py:381-398)
  if (tid == 0) {
    double trust = st[MOP_RS_TRUST];
    if (st[MOP_RS_HAVE_ENERGY] != 0.0) {
      const double actual = Be - st[MOP_RS_PREV_ENERGY];
      int na = (int)st[MOP_RS_NACT];
      if (na >= 3) {
        st[MOP_RS_ACT0] = st[MOP_RS_ACT0 + 1];
        st[MOP_RS_ACT0 + 1] = st[MOP_RS_ACT0 + 2];
        na = 2;
      }
      st[MOP_RS_ACT0 + na] = actual;
      st[MOP_RS_NACT] = na + 1;
      const int npred = (int)st[MOP_RS_NPRED];
      if (npred > 0)
        trust = adjust_trust(trust, actual, st[MOP_RS_PRED0 + npred - 1], lam[0], gnorm_raw,
                             saddle_order, tmin, tmax);
    }
    st[MOP_RS_TRUST] = trust;
    s_trust = trust;
  }

  // image function (rsirfo.py:408-425)
  for (int i = tid; i < n; i += nt) {
    R.lams[i] = lam[i];
    R.gams[i] = gam[i];
  }
  __syncthreads();
  if (tid == 0 && saddle_order > 0) {
    int found = 0;
    for (int i = 0; i < n && found < saddle_order; ++i) {
      if (fabs(lam[i]) > 1e-10) {
        if (neb_mode) {
          R.lams[i] = 0.0;
          R.gams[i] = 0.0;
        } else {
          R.lams[i] = -lam[i];
          R.gams[i] = -gam[i];
        }
        ++found;
      }
    }
  }
  __syncthreads();
  if (saddle_order > 0) {  // second "eigh": ascending order of the image spectrum
    for (int i = tid; i < n; i += nt) {
      const double li = R.lams[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (R.lams[j] < li) || (R.lams[j] == li && j < i);
      R.ord[rank] = i;
    }
    __syncthreads();
    for (int r = tid; r < n; r += nt) {
      R.w1[r] = R.lams[R.ord[r]];
      R.w2[r] = R.gams[R.ord[r]];
    }
    __syncthreads();
    for (int r = tid; r < n; r += nt) {
      R.lams[r] = R.w1[r];
      R.gams[r] = R.w2[r];
    }
    __syncthreads();
    if (!identity && spectrum_ill_conditioned(R.lams, n, R.scratch)) {
      flags |= MOP_ST_LEVEL_SHIFT;
      for (int i = tid; i < n; i += nt) R.lams[i] = __dadd_rn(__dadd_rn(R.lams[i], 1e-5), -1e-5);
    }
    __syncthreads();
  } else {
    for (int i = tid; i < n; i += nt) R.ord[i] = i;
    __syncthreads();
  }

  // small-eigenvalue filter (rsirfo.py:265-283,440), order preserved
  if (tid == 0) {
    int k = 0;
    for (int r = 0; r < n; ++r) {
      if (!(fabs(R.lams[r]) < 1e-6)) {
        R.lamk[k] = R.lams[r];
        R.gamk[k] = R.gams[r];
        R.ordk[k] = R.ord[r];
        ++k;
      }
    }
    s_k = k;
  }
  __syncthreads();
  const int kk = s_k;
  const double trust = s_trust;

  // RS step in the eigenbasis (rsirfo.py:924-985)
  if (MAXJ > 0) {
    const int nw = nt >> 5;
    RfoTerms<(MAXJ > 0 ? MAXJ : 1)> w;
    w.lam = R.lamk;
    w.gam = R.gamk;
    w.k = kk;
    double stepr[(MAXJ > 0 ? MAXJ : 1)];
    double mu0 = 0.0, n0 = 0.0;
    bool hard = false;
    if (wid == 0) {
      double n2;
      mu0 = solve_rfo_r(w, 1.0, lane, &hard, stepr, &n2);
      n0 = sqrt(n2);
      if (lane == 0) s_flags = (!(n0 <= trust) ? MOP_ST_ALPHA_SEARCH : 0) | (hard ? MOP_ST_HARD_CASE : 0);
    }
    __syncthreads();
    const int sf0 = s_flags;
    __syncthreads();
    if (sf0 & MOP_ST_ALPHA_SEARCH) {  // block-uniform
      // probes of the Brent bracket on the warps that would idle (warp 0 itself when the CTA has no others)
      const int w_lo = nw > 1 ? 1 : 0, w_hi = nw > 2 ? 2 : w_lo;
      int f = 0;
      if (wid == 0) f = alpha_newton_r(w, trust, mu0, n0, stepr, R.w3, lane);
      if (wid == w_lo) {
        const double o = alpha_probe_r<(MAXJ > 0 ? MAXJ : 1)>(R.lamk, R.gamk, kk, 1e-6, trust, lane);
        if (lane == 0) s_probe[0] = o;
      }
      if (wid == w_hi) {
        const double o = alpha_probe_r<(MAXJ > 0 ? MAXJ : 1)>(R.lamk, R.gamk, kk, 1000.0, trust, lane);
        if (lane == 0) s_probe[1] = o;
      }
      if (wid == 0 && lane == 0) s_flags |= f;
      __syncthreads();
      if (tid == 0 && s_probe[0] * s_probe[1] < 0.0) s_flags |= MOP_ST_BRENT_BRACKET;  // Brent branch not replayed
    }
    if (wid == 0) {
#pragma unroll
      for (int u = 0; u < (MAXJ > 0 ? MAXJ : 1); ++u)
        if (lane + 32 * u < kk) R.stepk[lane + 32 * u] = stepr[u];
    }
  } else if (wid == 0) {
    RfoWork w{R.lamk, R.gamk, R.w1, R.w2, R.stepk, kk};
    bool hard = false;
    int f = 0;
    solve_rfo(w, 1.0, lane, &hard);
    const double n0 = sqrt(warp_norm2(R.stepk, kk, lane));
    if (!(n0 <= trust)) {
      f |= MOP_ST_ALPHA_SEARCH;
      f |= alpha_search(w, trust, R.w3, lane);
    }
    if (hard) f |= MOP_ST_HARD_CASE;
    if (lane == 0) s_flags = f;
  }
  for (int i = tid; i < n; i += nt) R.coef[i] = 0.0;
  __syncthreads();
  flags |= s_flags;
  double bad = 0.0;
  for (int k = tid; k < kk; k += nt) {
    const double c = R.stepk[k];
    R.coef[R.ordk[k]] = c;
    if (!isfinite(c)) bad = 1.0;
  }
  bad = block_sum(bad, R.scratch);
  if (bad > 0.0) {  // rsirfo.py:456-462: steepest descent on the projected gradient
    flags |= MOP_ST_STEP_NAN_SD;
    double p = 0.0;
    for (int i = tid; i < n; i += nt) p = fma(gam[i], gam[i], p);
    const double nrm = sqrt(block_sum(p, R.scratch));
    const double sc = nrm > trust ? trust / nrm : 1.0;
    for (int i = tid; i < n; i += nt) R.coef[i] = -gam[i] * sc;
    __syncthreads();
  }

  // predicted energy change gp.s + 1/2 s^T Hp s in the eigenbasis (rsirfo.py:469,1717-1720):
  // s = sum c_k v_k  ->  sum_k c_k (gam_k + 1/2 lam_k c_k)
  double pe = 0.0;
  for (int i = tid; i < n; i += nt) {
    const double c = R.coef[i];
    pe += c * fma(0.5 * lam[i], c, gam[i]);
  }
  const double pred = block_sum(pe, R.scratch);
  if (tid == 0) {
    int npred = (int)st[MOP_RS_NPRED];
    if (npred >= 3) {
      st[MOP_RS_PRED0] = st[MOP_RS_PRED0 + 1];
      st[MOP_RS_PRED0 + 1] = st[MOP_RS_PRED0 + 2];
      npred = 2;
    }
    st[MOP_RS_PRED0 + npred] = pred;
    st[MOP_RS_NPRED] = npred + 1;
    st[MOP_RS_HAVE_PREV] = 1.0;
    st[MOP_RS_PREV_ENERGY] = Be;
    st[MOP_RS_HAVE_ENERGY] = 1.0;
    st[MOP_RS_ITER] += 1.0;
    if (pred_out) *pred_out = pred;
  }
  __syncthreads();
  return flags;
}

}  // namespace mop
