// RS-I-RFO step from a finished eigendecomposition, generic path.
//
// One CTA per structure.  Inputs: ascending eigenvalues and eigenvectors (row k =
// vector k) of the TR/ROT-projected Hessian Hp, the projected gradient gp, the raw
// biased gradient Bg (norm only), Hp itself (predicted energy change) and the
// per-structure RSIRFO state.  Follows Optimizer/rsirfo.py:360-490 step by step;
// the second eigendecomposition of the image Hessian H* (rsirfo.py:423-427) is
// derived from the first one: H* = V diag(lambda') V^T with the `saddle_order`
// lowest |lambda| > 1e-10 eigenvalues sign-flipped (zeroed in NEB mode), and
// g* has the matching components negated (zeroed).
#include "rfo_secular.cuh"

namespace mop {

constexpr int RFO_THREADS = 256;

// check_hessian_conditioning (rsirfo.py:492-551) -> ill-conditioned?
// lam in shared memory, whole CTA participates.
__device__ bool spectrum_ill_conditioned(const double* lam, int n, double* scratch) {
  if (n < 2) return false;
  double mx = 0.0, mn = INFINITY;
  double cnt = 0.0;
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
__device__ double adjust_trust(double trust, double actual, double predicted, double min_eig,
                               double gnorm, int saddle_order, double tmin, double tmax) {
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

__global__ void __launch_bounds__(RFO_THREADS)
k_rfo_step(int n, int saddle_order, int neb_mode, double tmin, double tmax,
           const double* __restrict__ evals_all, const double* __restrict__ evecs_all,
           const double* __restrict__ gp_all, const double* __restrict__ Bg_all,
           const double* __restrict__ Hp_all, const double* __restrict__ Be_all,
           double* __restrict__ state_all, double* __restrict__ move_all,
           double* __restrict__ evals_out, double* __restrict__ pred_all,
           int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = RFO_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* lam = sm;            // spectrum of Hp
  double* gam = lam + np;      // V^T gp
  double* gp = gam + np;
  double* lams = gp + np;      // image spectrum, sorted
  double* gams = lams + np;    // image gradient components, same order
  double* lamk = gams + np;    // kept (|lambda*| >= 1e-6), ascending
  double* gamk = lamk + np;
  double* stepk = gamk + np;
  double* w1 = stepk + np;     // scratch lamp
  double* w2 = w1 + np;        // scratch g2
  double* w3 = w2 + np;        // scratch best / full step (n)
  double* scratch = w3 + np;   // 40
  int* ord = (int*)(scratch + 40);  // mode index of sorted position
  int* ordk = ord + np;             // mode index of kept position
  __shared__ int s_k, s_flags, s_identity;
  __shared__ double s_trust;

  const double* V = evecs_all + (size_t)b * n * n;
  double* st = state_all + (size_t)b * MOP_RSIRFO_STATE;
  int flags = 0;

  // ---- spectrum, projected gradient, finiteness (rsirfo.py:360-369) -----------
  double bad = 0.0, pg = 0.0;
  for (int i = tid; i < n; i += RFO_THREADS) {
    const double l = evals_all[(size_t)b * n + i];
    lam[i] = l;
    gp[i] = gp_all[(size_t)b * n + i];
    if (!isfinite(l)) bad = 1.0;
    const double g = Bg_all[(size_t)b * n + i];
    pg = fma(g, g, pg);
  }
  const double gnorm_raw = sqrt(block_sum(pg, scratch));
  __syncthreads();
  for (int k = wid; k < n; k += nw) {
    const double* vk = V + (size_t)k * n;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = vk[i];
      if (!isfinite(v)) bad = 1.0;
      acc = fma(v, gp[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) gam[k] = acc;
  }
  bad = block_sum(bad, scratch);
  const bool identity = bad > 0.0;
  if (identity) {
    flags |= MOP_ST_EIG_NONFINITE;
    for (int i = tid; i < n; i += RFO_THREADS) {
      lam[i] = 1.0;
      gam[i] = gp[i];
    }
  }
  __syncthreads();
  if (evals_out)
    for (int i = tid; i < n; i += RFO_THREADS) evals_out[(size_t)b * n + i] = lam[i];

  // ---- level shift emulation (rsirfo.py:602-631): eigh(H + 1e-5 I) - 1e-5 -------
  if (!identity && spectrum_ill_conditioned(lam, n, scratch)) {
    flags |= MOP_ST_LEVEL_SHIFT;
    for (int i = tid; i < n; i += RFO_THREADS) lam[i] = __dadd_rn(__dadd_rn(lam[i], 1e-5), -1e-5);
    __syncthreads();
  }

  // ---- inner trust radius bookkeeping (rsirfo.py:381-398) ----------------------
  if (tid == 0) {
    double trust = st[MOP_RS_TRUST];
    const double Be = Be_all ? Be_all[b] : 0.0;
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
  __syncthreads();

  // ---- image function: flip (zero) the first `saddle_order` modes (rsirfo.py:408-425)
  for (int i = tid; i < n; i += RFO_THREADS) {
    lams[i] = lam[i];
    gams[i] = gam[i];
  }
  __syncthreads();
  if (tid == 0 && saddle_order > 0) {
    int found = 0;
    for (int i = 0; i < n && found < saddle_order; ++i) {
      if (fabs(lam[i]) > 1e-10) {
        if (neb_mode) {
          lams[i] = 0.0;
          gams[i] = 0.0;
        } else {
          lams[i] = -lam[i];
          gams[i] = -gam[i];
        }
        ++found;
      }
    }
  }
  __syncthreads();
  // second "eigh": ascending order of the image spectrum (stable by mode index)
  if (saddle_order > 0) {
    for (int i = tid; i < n; i += RFO_THREADS) {
      const double li = lams[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (lams[j] < li) || (lams[j] == li && j < i);
      ord[rank] = i;
    }
    __syncthreads();
    for (int r = tid; r < n; r += RFO_THREADS) {
      w1[r] = lams[ord[r]];
      w2[r] = gams[ord[r]];
    }
    __syncthreads();
    for (int r = tid; r < n; r += RFO_THREADS) {
      lams[r] = w1[r];
      gams[r] = w2[r];
    }
    __syncthreads();
    if (!identity && spectrum_ill_conditioned(lams, n, scratch)) {
      flags |= MOP_ST_LEVEL_SHIFT;
      for (int i = tid; i < n; i += RFO_THREADS) lams[i] = __dadd_rn(__dadd_rn(lams[i], 1e-5), -1e-5);
      __syncthreads();
    }
  } else {
    for (int i = tid; i < n; i += RFO_THREADS) ord[i] = i;
    __syncthreads();
  }

  // ---- small-eigenvalue filter (rsirfo.py:265-283,440), order preserved ---------
  if (tid == 0) {
    int k = 0;
    for (int r = 0; r < n; ++r) {
      if (!(fabs(lams[r]) < 1e-6)) {
        lamk[k] = lams[r];
        gamk[k] = gams[r];
        ordk[k] = ord[r];
        ++k;
      }
    }
    s_k = k;
  }
  __syncthreads();
  const int kk = s_k;
  const double trust = s_trust;

  // ---- RS step in the eigenbasis: warp 0 (rsirfo.py:924-985) ---------------------
  if (wid == 0) {
    RfoWork w{lamk, gamk, w1, w2, stepk, kk};
    bool hard = false;
    int f = 0;
    solve_rfo(w, 1.0, lane, &hard);
    const double n0 = sqrt(warp_norm2(stepk, kk, lane));
    if (!(n0 <= trust)) {
      f |= MOP_ST_ALPHA_SEARCH;
      f |= alpha_search(w, trust, w3, lane);
    }
    if (hard) f |= MOP_ST_HARD_CASE;
    if (lane == 0) s_flags = f;
  }
  __syncthreads();
  flags |= s_flags;

  // ---- back-transform: step = V* step_k ------------------------------------------
  double* full = w3;
  double nonfinite = 0.0;
  for (int i = tid; i < n; i += RFO_THREADS) {
    double acc = 0.0;
    if (identity) {
      // V = I, kept position k <-> coordinate ordk[k]
      for (int k = 0; k < kk; ++k)
        if (ordk[k] == i) acc = stepk[k];
    } else {
      for (int k = 0; k < kk; ++k) acc = fma(V[(size_t)ordk[k] * n + i], stepk[k], acc);
    }
    full[i] = acc;
    if (!isfinite(acc)) nonfinite = 1.0;
  }
  nonfinite = block_sum(nonfinite, scratch);
  if (nonfinite > 0.0) {  // rsirfo.py:456-462
    flags |= MOP_ST_STEP_NAN_SD;
    double p = 0.0;
    for (int i = tid; i < n; i += RFO_THREADS) p = fma(gp[i], gp[i], p);
    const double nrm = sqrt(block_sum(p, scratch));
    const double sc = nrm > trust ? trust / nrm : 1.0;
    for (int i = tid; i < n; i += RFO_THREADS) full[i] = -gp[i] * sc;
  }
  __syncthreads();

  // ---- predicted energy change  gp.s + 1/2 s^T Hp s  (rsirfo.py:469,1717-1720) ------
  const double* Hp = Hp_all + (size_t)b * n * n;
  double pe = 0.0;
  for (int i = wid; i < n; i += nw) {
    const double* row = Hp + (size_t)i * n;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(row[j], full[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) pe += full[i] * (gp[i] + 0.5 * acc);
  }
  const double pred = block_sum(pe, scratch);

  for (int i = tid; i < n; i += RFO_THREADS) move_all[(size_t)b * n + i] = -full[i];
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
    st[MOP_RS_PREV_ENERGY] = Be_all ? Be_all[b] : 0.0;
    st[MOP_RS_HAVE_ENERGY] = 1.0;
    st[MOP_RS_ITER] += 1.0;
    if (pred_all) pred_all[b] = pred;
    if (status) {
      const int keep = status[b] & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV |
                                    MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF |
                                    MOP_ST_EIG_NOCONV | MOP_ST_EIG_FALLBACK);
      status[b] = keep | flags;
    }
  }
}

__global__ void k_clamp_and_move(int n, const double* __restrict__ x, double* __restrict__ move,
                                 const double* __restrict__ trust, double* __restrict__ xnew) {
  __shared__ double scratch[40];
  const int b = blockIdx.x;
  double p = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = move[(size_t)b * n + i];
    p = fma(m, m, p);
  }
  const double nrm = sqrt(block_sum(p, scratch));
  const double tr = trust[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double m = move[(size_t)b * n + i];
    if (nrm > tr) {
      m = tr * m / nrm;  // optimizer.py:793
      move[(size_t)b * n + i] = m;
    }
    if (xnew) xnew[(size_t)b * n + i] = (x[(size_t)b * n + i] - m) * 0.52917721067;
  }
}

}  // namespace mop

int mop_launch_rfo_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                        const double* evals, const double* evecs, const double* gp,
                        const double* Bg, const double* Hp, const double* Be, double* state,
                        double* move, double* evals_out, double* pred, int32_t* status,
                        cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  const int np = (n + 3) & ~3;
  const size_t smem = sizeof(double) * (11 * (size_t)np + 40) + sizeof(int) * 2 * (size_t)np;
  if (smem > 200 * 1024) {
    mop_set_error("n = %d too large for the RFO step kernel", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_rfo_step, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  mop::k_rfo_step<<<B, mop::RFO_THREADS, smem, stream>>>(n, saddle_order, neb_mode, tmin, tmax,
                                                       evals, evecs, gp, Bg, Hp, Be, state, move,
                                                       evals_out, pred, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_clamp_and_move(int B, int n, const double* x, double* move,
                                  const double* trust_outer, double* x_new_ang, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_clamp_and_move: B >= 0 and n > 0 required");
  MOP_REQUIRE(move && trust_outer, "mop_clamp_and_move: move and trust_outer required");
  MOP_REQUIRE(!x_new_ang || x, "mop_clamp_and_move: x required with x_new_ang");
  if (B == 0) return MOP_OK;
  mop::k_clamp_and_move<<<B, 128, 0, (cudaStream_t)stream>>>(n, x, move, trust_outer, x_new_ang);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
