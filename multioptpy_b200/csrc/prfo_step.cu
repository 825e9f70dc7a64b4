// RS-P-RFO saddle-search step: EnhancedRSPRFO.run,
// Optimizer/rsprfo.py:713-886, for a batch of structures.
//
// Sequence per call (host side below): TR/ROT projection of gradient and Hessian ->
// k_prfo_prev (reduction ratio of the previous step and Nocedal-Wright trust update,
// rsprfo.py:908-962,421-512) -> Hessian update with the BIASED gradients (:1190-1260) ->
// eigendecomposition of H + H_bias (NOT projected, :780-783) -> k_prfo_step: eigenvalue
// shifting (:287-355; the second eigh of the rebuilt Hessian is a re-sort), mode following
// (:964-1071), P-RFO step = extreme eigenpairs of the two arrowhead matrices (:1097-1168),
// trust / gradient scaling (:357-419,854-865), predicted energy change, state.
// The alpha micro-cycles (:514-662) divide eigenvalues AND gradient by alpha, so every cycle
// reproduces the alpha = 1 step and the loop returns it scaled to the effective
// trust radius; the kernel evaluates it once.  The arrowhead eigenpairs are obtained from the
// secular equation  nu + sum g_i^2 / (lambda_i - nu) = 0  solved to machine precision in the
// pole-relative variable t = lambda_pole - nu (the reference calls LAPACK eigh).
#include "common.cuh"

namespace mop {

constexpr int PRFO_THREADS = 256;

// smallest root of  nu + sum g2_i / (lam_i - nu) = 0  for ascending lam[0..k): returns t > 0
// with nu = lam[pole] - t, pole = first index with g2 > 0 (-1: all gradient components zero).
// One warp; every lane returns the same value.
__device__ double secular_min_root(const double* lam, const double* g2, int k, int lane, int* pole_out) {
  int pole = 0x7fffffff;
  double gs = 0.0;
  for (int i = lane; i < k; i += 32) {
    if (g2[i] > 0.0 && i < pole) pole = i;
    gs += g2[i];
  }
  gs = warp_sum(gs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pole = min(pole, __shfl_xor_sync(MOP_FULL_MASK, pole, o));
  if (pole == 0x7fffffff) {
    *pole_out = -1;
    return 0.0;
  }
  *pole_out = pole;
  const double lp = lam[pole];
  // F(t) = f(lp - t) = lp - t + sum g2_i / ((lam_i - lp) + t); decreasing in t on (0, inf)
  auto F = [&](double t, double* dF) {
    double a = 0.0, d = 0.0;
    for (int i = lane; i < k; i += 32) {
      if (g2[i] == 0.0) continue;
      const double den = (lam[i] - lp) + t;
      const double q = g2[i] / den;
      a += q;
      d += q / den;
    }
    a = warp_sum(a);
    d = warp_sum(d);
    *dF = -1.0 - d;
    return (lp - t) + a;
  };
  // bracket: F(0+) = +inf, F(t_hi) <= 0 with the Baker bound t_hi = lp - guess
  double t_hi = lp - 0.5 * (lp - sqrt(fmax(0.0, lp * lp + 4.0 * gs)));
  if (!(t_hi > 0.0)) t_hi = fmax(fabs(lp), 1.0) * 1e-300 + sqrt(gs) + 1e-300;
  double dF;
  int guard = 0;
  while (F(t_hi, &dF) > 0.0 && guard++ < 200) t_hi *= 2.0;
  double t_lo = 0.0;
  double t = t_hi;
  for (int it = 0; it < 200; ++it) {
    const double f = F(t, &dF);
    if (f == 0.0) break;
    if (f > 0.0) t_lo = t; else t_hi = t;
    double tn = t - f / dF;
    if (!(tn > t_lo && tn < t_hi)) tn = 0.5 * (t_lo + t_hi);
    if (fabs(tn - t) <= 4.0 * 2.220446049250313e-16 * fabs(tn)) {
      t = tn;
      break;
    }
    t = tn;
    if (t_hi - t_lo <= 4.0 * 2.220446049250313e-16 * t_hi) break;
  }
  return t;
}

#define MOP_PS_TRUST 0
#define MOP_PS_FIRST_DONE 1
#define MOP_PS_HAVE_ENERGY 2
#define MOP_PS_PREV_ENERGY 3
#define MOP_PS_HAVE_PRED 4
#define MOP_PS_LAST_PRED 5
#define MOP_PS_HAVE_TS 6
#define MOP_PS_ITER 7

// _process_previous_step + compute_reduction_ratio + adjust_trust_radius
__global__ void __launch_bounds__(256)
k_prfo_prev(int n, const double* __restrict__ Hp_all, const double* __restrict__ prev_grad,
            const double* __restrict__ prev_move, const double* __restrict__ pre_move_arg,
            const double* __restrict__ Be, double* __restrict__ state_all, double tmin, double tmax) {
  extern __shared__ double sm[];
  __shared__ double scratch[40];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  double* st = state_all + (size_t)b * MOP_PRFO_STATE;
  if (st[MOP_PS_FIRST_DONE] == 0.0 || st[MOP_PS_HAVE_ENERGY] == 0.0 || st[MOP_PS_HAVE_PRED] == 0.0) return;
  double* s = sm;
  for (int i = tid; i < n; i += blockDim.x) s[i] = prev_move[(size_t)b * n + i];
  __syncthreads();
  const double* H = Hp_all + (size_t)b * n * n;
  double part = 0.0;
  for (int i = w; i < n; i += nw) {
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(H[(size_t)i * n + j], s[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) part += s[i] * (prev_grad[(size_t)b * n + i] + 0.5 * acc);
  }
  const double model = block_sum(part, scratch);  // g.s + 1/2 s^T H s
  double p2 = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double v = pre_move_arg ? pre_move_arg[(size_t)b * n + i] : s[i];
    p2 = fma(v, v, p2);
  }
  const double psn = sqrt(block_sum(p2, scratch));
  if (tid != 0) return;
  const double actual = Be[b] - st[MOP_PS_PREV_ENERGY];
  const double pred_red = -model;
  double ratio;
  if (fabs(pred_red) < 1e-14) ratio = fabs(actual) < 1e-14 ? 1.0 : 0.0;
  else {
    ratio = actual / pred_red;
    if (!isfinite(ratio)) ratio = 0.0;
  }
  double trust = st[MOP_PS_TRUST];
  const bool at_boundary = psn >= trust * 0.95;
  if (ratio < 0.25) trust = fmax(0.25 * psn, tmin);
  else if (ratio > 0.75 && at_boundary) trust = fmin(2.0 * trust, tmax);
  st[MOP_PS_TRUST] = trust;
}

// Update rejection of EnhancedRSPRFO.update_hessian (rsprfo.py:1239-1250): the reference computes eigvalsh of the
// updated Hessian and keeps the OLD Hessian when max |lambda| > 1e6.  ||H||_2 <= ||H||_F <= sqrt(n) ||H||_2, so the
// Frobenius norm settles every structure outside the band 1e6 < ||H||_F <= 1e6 sqrt(n); structures inside it are
// flagged for the Jacobi eigensolver (flag = MOP_ST_EIG_FALLBACK, the bit its only_flagged filter tests).  A
// non-finite Hessian is accepted, as in the reference (eigvalsh returns NaN, and NaN > 1e6 is false).
__global__ void __launch_bounds__(512) k_prfo_guard(int n, const double* __restrict__ H, const int32_t* __restrict__ status,
                                                    int32_t* __restrict__ flag) {
  __shared__ double scratch[40];
  const size_t b = blockIdx.x, nn = (size_t)n * n;
  if (!(status[b] & MOP_ST_UPDATED)) {  // block-uniform
    if (threadIdx.x == 0) flag[b] = 0;
    return;
  }
  double acc = 0.0;
  for (size_t e = threadIdx.x; e < nn; e += blockDim.x) {
    const double x = H[b * nn + e];
    acc = fma(x, x, acc);
  }
  const double fro = sqrt(block_sum(acc, scratch));
  if (threadIdx.x == 0) {
    int f = 0;
    if (isfinite(fro) && fro > 1e6) f = (fro > 1e6 * sqrt((double)n)) ? 2 : MOP_ST_EIG_FALLBACK;
    flag[b] = f;
  }
}

// flag 2: reject; flag MOP_ST_EIG_FALLBACK: reject when the Jacobi spectrum of the updated Hessian exceeds 1e6
__global__ void k_prfo_revert(int n, double* __restrict__ H, const double* __restrict__ Hold,
                              const int32_t* __restrict__ flag, const double* __restrict__ evals,
                              int32_t* __restrict__ status) {
  const size_t b = blockIdx.y, nn = (size_t)n * n;
  const int f = flag[b];
  if (f == 0) return;
  if (f != 2) {
    double mx = 0.0;
    for (int i = 0; i < n; ++i) mx = fmax(mx, fabs(evals[b * n + i]));  // NaN-free by construction (fmax drops NaN)
    if (!(mx > 1e6)) return;
  }
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
    H[b * nn + e] = Hold[b * nn + e];
  if (blockIdx.x == 0 && threadIdx.x == 0)
    status[b] = (status[b] & ~(MOP_ST_UPDATED | MOP_ST_UPD_TERM_ZEROED)) | MOP_ST_UPD_REJECTED;
}

__global__ void k_sum_sym(int n, const double* __restrict__ H, const double* __restrict__ Hb,
                          double* __restrict__ out) {
  const size_t b = blockIdx.y;
  const size_t nn = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x)
    out[b * nn + e] = H[b * nn + e] + (Hb ? Hb[b * nn + e] : 0.0);
}

__global__ void __launch_bounds__(PRFO_THREADS, 1)
k_prfo_step(int n, int so, double tmin, double tmax, const double* __restrict__ evals_all,
            const double* __restrict__ evecs_all, const double* __restrict__ gp_all,
            const double* __restrict__ Bg_all, const double* __restrict__ Be_all, double* __restrict__ state_all,
            double* __restrict__ prev_grad, double* __restrict__ prev_move, double* __restrict__ ts_all,
            double* __restrict__ move_all, double* __restrict__ evals_out, double* __restrict__ pred_all,
            int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = PRFO_THREADS >> 5;
  const int np = (n + 3) & ~3;
  double* lam = sm;            // shifted spectrum, eigh index
  double* gam = lam + np;      // V^T g
  double* ov = gam + np;       // |V^T ts|
  double* gp = ov + np;
  double* cf = gp + np;        // step coefficients, eigh index
  double* la = cf + np;        // subspace work: lam
  double* g2 = la + np;        //                g^2
  double* scratch = g2 + np;   // 40
  int* ord = (int*)(scratch + 40);  // ascending order of the shifted spectrum
  int* sub = ord + np;              // eigh indices of the current subspace
  unsigned char* ismax = (unsigned char*)(sub + np);
  __shared__ int s_shift, s_best, s_kmin, s_kmax, s_flags;
  __shared__ double s_tmin, s_tmax_root;
  __shared__ int s_pole_min, s_pole_max;

  const double* V = evecs_all + (size_t)b * n * n;
  double* st = state_all + (size_t)b * MOP_PRFO_STATE;
  double* ts = ts_all + (size_t)b * n;
  int flags = 0;
  if (tid == 0) s_flags = 0;

  double bad = 0.0;
  for (int i = tid; i < n; i += PRFO_THREADS) {
    const double l = evals_all[(size_t)b * n + i];
    lam[i] = l;
    gp[i] = gp_all[(size_t)b * n + i];
    if (!isfinite(l)) bad = 1.0;
    ismax[i] = 0;
  }
  const bool have_ts = st[MOP_PS_HAVE_TS] != 0.0;
  for (int k = wid; k < n; k += nw) {
    const double* vk = V + (size_t)k * n;
    double a = 0.0, o = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = vk[i];
      if (!isfinite(v)) bad = 1.0;
      a = fma(v, gp_all[(size_t)b * n + i], a);
      if (have_ts) o = fma(v, ts[i], o);
    }
    a = warp_sum(a);
    o = warp_sum(o);
    if (lane == 0) {
      gam[k] = a;
      ov[k] = fabs(o);
    }
  }
  bad = block_sum(bad, scratch);
  const bool identity = bad > 0.0;
  if (identity) {
    flags |= MOP_ST_EIG_NONFINITE;
    for (int i = tid; i < n; i += PRFO_THREADS) {
      lam[i] = 1.0;
      gam[i] = gp[i];
      ov[i] = have_ts ? fabs(ts[i]) : 0.0;
    }
  }
  __syncthreads();
  double pg = 0.0;
  for (int i = tid; i < n; i += PRFO_THREADS) pg = fma(gp[i], gp[i], pg);
  const double gnorm = sqrt(block_sum(pg, scratch));

  // ---- eigenvalue shifting (rsprfo.py:287-355); lam is ascending on entry ------------------
  if (tid == 0) {
    int sh = 0;
    if (so == 0) {
      if (lam[0] < 0.001) {
        const double d = 0.001 - lam[0];
        for (int i = 0; i < n; ++i) lam[i] += d;
        sh = 1;
      }
    } else {
      for (int i = 0; i < so && i < n; ++i)
        if (lam[i] > -0.001) {
          lam[i] = -0.001;
          sh = 1;
        }
      for (int i = so; i < n; ++i)
        if (lam[i] < 1e-6) {
          lam[i] = 0.001;
          sh = 1;
        }
    }
    s_shift = sh;
  }
  __syncthreads();
  if (s_shift) flags |= MOP_ST_LEVEL_SHIFT;
  // the second eigh of V diag(lam') V^T: ascending order of the shifted values
  for (int i = tid; i < n; i += PRFO_THREADS) {
    const double li = lam[i];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (lam[j] < li) || (lam[j] == li && j < i);
    ord[r] = i;
  }
  __syncthreads();
  if (evals_out)
    for (int r = tid; r < n; r += PRFO_THREADS) evals_out[(size_t)b * n + r] = lam[ord[r]];

  // ---- mode selection (rsprfo.py:964-1071), thread 0 ------------------------------------
  if (tid == 0) {
    int best = -1;
    if (so > 0) {
      if (!have_ts) {
        best = ord[0];
        for (int r = 0; r < so && r < n; ++r) ismax[ord[r]] = 1;
      } else {
        int bi = ord[0];
        double bo = ov[ord[0]];
        for (int r = 1; r < n; ++r)
          if (ov[ord[r]] > bo) {
            bo = ov[ord[r]];
            bi = ord[r];
          }
        bool pick_lowest = false;
        if (bo > 0.5) {
          best = bi;
        } else {  // _handle_mode_mixing
          double bw = -1.0;
          int bj = -1;
          for (int r = 0; r < n; ++r) {
            const int i = ord[r];
            if (ov[i] > 0.3) {
              const double wgt = ov[i] * ov[i] * (lam[i] < 0.0 ? 1.0 : 0.1);
              if (wgt > bw) {
                bw = wgt;
                bj = i;
              }
            }
          }
          if (bj < 0) pick_lowest = true; else best = bj;
        }
        if (pick_lowest) {
          best = ord[0];
          for (int r = 0; r < so && r < n; ++r) ismax[ord[r]] = 1;
        } else {
          ismax[best] = 1;
          int need = so - 1;
          for (int r = 0; r < n && need > 0; ++r)
            if (ord[r] != best) {
              ismax[ord[r]] = 1;
              --need;
            }
        }
      }
    }
    s_best = best;
  }
  __syncthreads();
  for (int i = tid; i < n; i += PRFO_THREADS) cf[i] = 0.0;

  // ---- P-RFO step: extreme roots of the two arrowhead matrices (rsprfo.py:1097-1168) -----------
  for (int pass = 0; pass < 2; ++pass) {  // 0: min subspace (ascending), 1: max subspace (negated, ascending)
    __syncthreads();
    if (tid == 0) {
      int k = 0;
      if (pass == 0) {
        for (int r = 0; r < n; ++r) {
          const int i = ord[r];
          if (!ismax[i]) {
            sub[k] = i;
            la[k] = lam[i];
            g2[k] = gam[i] * gam[i];
            ++k;
          }
        }
        s_kmin = k;
      } else {
        for (int r = n - 1; r >= 0; --r) {
          const int i = ord[r];
          if (ismax[i]) {
            sub[k] = i;
            la[k] = -lam[i];
            g2[k] = gam[i] * gam[i];
            ++k;
          }
        }
        s_kmax = k;
      }
    }
    __syncthreads();
    const int k = pass == 0 ? s_kmin : s_kmax;
    if (k == 0) continue;
    if (wid == 0) {
      int pole;
      const double t = secular_min_root(la, g2, k, lane, &pole);
      if (pole < 0) {
        if (lane == 0) atomicOr(&s_flags, MOP_ST_HARD_CASE);
      } else {
        const double lp = la[pole];
        // step_i = g_i / (lambda_i - nu);  min: la = lam, nu = lp - t;  max: la = -lam, nu = -(lp - t)
        for (int q = lane; q < k; q += 32) {
          const int i = sub[q];
          const double den = (la[q] - lp) + t;  // = (+-)(lambda_i - nu)
          cf[i] = (pass == 0 ? 1.0 : -1.0) * gam[i] / den;
        }
      }
    }
  }
  __syncthreads();
  flags |= s_flags;

  // ---- trust scaling in the eigenbasis, NaN fallback (rsprfo.py:514-662,833-846) ----------------
  double trust = st[MOP_PS_TRUST];
  double eff = trust;
  if (gnorm < 1e-3) eff = fmin(fmax(0.5 * gnorm / 1e-3 * tmax, tmin), trust);
  double p2 = 0.0, nf = 0.0;
  for (int i = tid; i < n; i += PRFO_THREADS) {
    p2 = fma(cf[i], cf[i], p2);
    if (!isfinite(cf[i])) nf = 1.0;
  }
  double nrm = sqrt(block_sum(p2, scratch));
  nf = block_sum(nf, scratch);
  if (nf > 0.0 || !isfinite(nrm)) {
    flags |= MOP_ST_STEP_NAN_SD;
    double q2 = 0.0;
    for (int i = tid; i < n; i += PRFO_THREADS) q2 = fma(gam[i], gam[i], q2);
    const double sdn = sqrt(block_sum(q2, scratch));
    const double tgt = fmin(sdn, trust);
    for (int i = tid; i < n; i += PRFO_THREADS) cf[i] = sdn > 1e-12 ? -gam[i] * (tgt / sdn) : 0.0;
    nrm = sdn > 1e-12 ? tgt : 0.0;
  } else if (nrm > eff) {
    const double sc = eff / nrm;
    for (int i = tid; i < n; i += PRFO_THREADS) cf[i] *= sc;
    nrm = eff;
    flags |= MOP_ST_ALPHA_SEARCH;  // step limited by the (effective) trust radius
  }
  __syncthreads();
  // gradient-based scaling and the 1.01 trust rule act on ||V c|| = ||c|| (V orthonormal)
  double scale = 1.0;
  if (!(gnorm < 1e-10 || nrm < 1e-10)) {
    const double r = nrm / gnorm;
    if (r > 50.0) scale = fmax(50.0 / r, 0.1);
  }
  double snorm = nrm * scale;
  if (snorm > eff * 1.01) {
    scale *= eff / snorm;
    snorm = eff;
  }
  if (scale != 1.0)
    for (int i = tid; i < n; i += PRFO_THREADS) cf[i] *= scale;
  __syncthreads();

  // ---- back-transform, prediction, state --------------------------------------------------------
  for (int i = tid; i < n; i += PRFO_THREADS) {
    double acc = 0.0;
    if (identity) acc = cf[i];
    else
      for (int k = 0; k < n; ++k) {
        const double c = cf[k];
        if (c != 0.0) acc = fma(V[(size_t)k * n + i], c, acc);
      }
    move_all[(size_t)b * n + i] = acc;
    prev_move[(size_t)b * n + i] = acc;
    prev_grad[(size_t)b * n + i] = Bg_all[(size_t)b * n + i];
    if (s_best >= 0) ts[i] = identity ? (i == s_best ? 1.0 : 0.0) : V[(size_t)s_best * n + i];
  }
  double pe = 0.0;
  for (int i = tid; i < n; i += PRFO_THREADS) pe += cf[i] * fma(0.5 * lam[i], cf[i], gam[i]);
  const double pred = block_sum(pe, scratch);
  if (tid == 0) {
    st[MOP_PS_FIRST_DONE] = 1.0;
    st[MOP_PS_HAVE_ENERGY] = 1.0;
    st[MOP_PS_PREV_ENERGY] = Be_all ? Be_all[b] : 0.0;
    st[MOP_PS_HAVE_PRED] = 1.0;
    st[MOP_PS_LAST_PRED] = pred;
    if (s_best >= 0) st[MOP_PS_HAVE_TS] = 1.0;
    st[MOP_PS_ITER] += 1.0;
    if (pred_all) pred_all[b] = pred;
    if (status) {
      const int keep = status[b] & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_TERM_ZEROED |
                                    MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF | MOP_ST_EIG_NOCONV | MOP_ST_EIG_FALLBACK |
                                    MOP_ST_UPD_REJECTED);
      status[b] = keep | flags;
    }
  }
}

}  // namespace mop

// launchers implemented elsewhere
int mop_launch_hessian_update(int B, int n, int method, int mode, int guards, double* H,
                              const double* s, const double* y, const double* x, const double* xp,
                              const double* g, const double* gp, const double* state, int state_stride,
                              double* delta_out, int32_t* status, cudaStream_t stream);
int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                             const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                             cudaStream_t stream);
size_t mop_project_scratch_bytes(int B, int n);
size_t mop_hessian_update_scratch_bytes(int B, int n);
int mop_launch_project_trrot_split(int B, int n, const double* H, const double* Hbias, const double* x,
                                   const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream);
int mop_launch_hessian_update_split(int B, int n, int method, int mode, int guards, double* H, const double* s,
                                    const double* y, const double* x, const double* xp, const double* g,
                                    const double* gp, const double* state, int state_stride, double* delta_out,
                                    int32_t* status, void* scratch, size_t scratch_bytes, cudaStream_t stream);
extern "C" size_t mop_eigh_workspace_bytes(int B, int n, int algo);
extern "C" int mop_eigh(int B, int n, int algo, const double* A, double* evals, double* evecs,
                        int32_t* status, void* work, size_t work_bytes, void* stream);

size_t mop_jacobi_workspace_bytes(int B, int n);
int mop_launch_eigh_jacobi(int B, int n, const double* A, double* evals, double* evecs, int32_t* status,
                           const int32_t* only_flagged, void* work, size_t work_bytes, cudaStream_t stream);
int mop_tridiag_supported(int n);
int mop_large_supported(int n);
int mop_launch_eigh_large_factored(int B, int n, const double* A, double* evals, double* Zt, int32_t* status,
                                   void* work, size_t work_bytes, cudaStream_t stream);
int mop_launch_large_apply_q(int B, int n, int trans, void* work, double* x0, double* x1, double* x2, double* x3,
                             cudaStream_t stream);

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
// large systems: the step is taken in the basis of the tridiagonal matrix (eigh_large.cu, part 4)
static bool prfo_factored(int n, int eigh_algo) {
  if (eigh_algo == MOP_EIGH_LARGE) return mop_large_supported(n) != 0;
  return eigh_algo == MOP_EIGH_AUTO && mop_large_supported(n);  // small n: packed front end (eigh_large.cu)
}

// workspace: Hp/A | evecs | evals | gp | 4 rotated vectors | eigh work
extern "C" size_t mop_rsprfo_workspace_bytes(int B, int n, int eigh_algo) {
  if (B <= 0 || n <= 0) return 0;
  const size_t nn = al256(sizeof(double) * (size_t)B * n * n), nv = al256(sizeof(double) * (size_t)B * n);
  return 2 * nn + 6 * nv + mop_eigh_workspace_bytes(B, n, prfo_factored(n, eigh_algo) ? MOP_EIGH_LARGE : eigh_algo);
}

extern "C" int mop_rsprfo_step(int B, int n, int method, int saddle_order, int eigh_algo, double trust_min,
                               double trust_max, double* H, const double* Hbias, const double* x,
                               const double* Bg, const double* x_prev, const double* Bg_prev,
                               const double* pre_move, const double* Be, double* state, double* prev_grad,
                               double* prev_move, double* ts_vec, double* move_out, double* eigvals_out,
                               double* pred_out, int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0 && n % 3 == 0, "mop_rsprfo_step: n must be a positive multiple of 3");
  MOP_REQUIRE(H && x && Bg && Be && state && prev_grad && prev_move && ts_vec && move_out && status && work,
              "mop_rsprfo_step: required device pointer is NULL");
  MOP_REQUIRE(saddle_order >= 0 && saddle_order < n, "mop_rsprfo_step: bad saddle_order");
  MOP_REQUIRE((x_prev == nullptr) == (Bg_prev == nullptr), "mop_rsprfo_step: x_prev and Bg_prev go together");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_rsprfo_workspace_bytes(B, n, eigh_algo)) {
    mop_set_error("mop_rsprfo_step: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = al256(sizeof(double) * (size_t)B * n * n), nv = al256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  double* A = (double*)w;  // projected Hessian first, then H + Hbias
  double* evecs = (double*)(w + nn);
  double* evals = (double*)(w + 2 * nn);
  double* gp = (double*)(w + 2 * nn + nv);
  double* Xg = (double*)(w + 2 * nn + 2 * nv);
  double* Xts = (double*)(w + 2 * nn + 3 * nv);
  double* Xmv = (double*)(w + 2 * nn + 4 * nv);
  double* Xpm = (double*)(w + 2 * nn + 5 * nv);
  char* ework = w + 2 * nn + 6 * nv;
  const size_t ebytes = work_bytes - (2 * nn + 6 * nv);
  const bool factored = prfo_factored(n, eigh_algo);
  MOP_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, stream));
  // projected gradient (current geometry) and projected pre-update Hessian for the reduction ratio
  // small batches: the multi-CTA projection fills the GPU; large ones: one CTA per structure is as fast
  const bool prj_split = B <= 2 * 148 && nn >= mop_project_scratch_bytes(B, n);  // (tiny n: scratch exceeds the slab)
  int rc = prj_split ? mop_launch_project_trrot_split(B, n, H, Hbias, x, Bg, A, gp, status, 1, evecs, nn, stream)
                     : mop_launch_project_trrot(B, n, H, Hbias, x, Bg, A, gp, status, 1, stream);
  if (rc != MOP_OK) return rc;
  {
    const size_t smem = sizeof(double) * (size_t)n;
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_prfo_prev, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_prfo_prev<<<B, 256, smem, stream>>>(n, A, prev_grad, prev_move, pre_move, Be, state, trust_min, trust_max);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  if (x_prev && method != MOP_UPD_NONE) {  // biased gradients, small-change skip only (rsprfo.py:1203-1213)
    // the pre-update Hessian is kept in A (dead since k_prfo_prev) until the rejection test has passed
    MOP_CHECK_CUDA(cudaMemcpyAsync(A, H, sizeof(double) * (size_t)B * n * n, cudaMemcpyDeviceToDevice, stream));
    // scratch: the eigenvector buffer is not live yet
    rc = nn >= mop_hessian_update_scratch_bytes(B, n)
             ? mop_launch_hessian_update_split(B, n, method, 1, 2, H, nullptr, nullptr, x, x_prev, Bg, Bg_prev, state,
                                               MOP_PRFO_STATE, nullptr, status, evecs, nn, stream)
             : mop_launch_hessian_update(B, n, method, 1, 2, H, nullptr, nullptr, x, x_prev, Bg, Bg_prev, state,
                                         MOP_PRFO_STATE, nullptr, status, stream);
    if (rc != MOP_OK) return rc;
    // rsprfo.py:1239-1250: revert an update whose spectrum exceeds 1e6
    int32_t* rflag = (int32_t*)Xg;  // [B] ints; the rotated-vector slots are not live yet
    mop::k_prfo_guard<<<B, 512, 0, stream>>>(n, H, status, rflag);
    MOP_CHECK_CUDA(cudaGetLastError());
    const size_t jac = al256(mop_jacobi_workspace_bytes(B, n));
    rc = mop_launch_eigh_jacobi(B, n, H, evals, evecs, nullptr, rflag, ework, jac, stream);  // almost always empty
    if (rc != MOP_OK) return rc;
    dim3 rgrid(32, B);
    mop::k_prfo_revert<<<rgrid, 256, 0, stream>>>(n, H, A, rflag, evals, status);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  {
    dim3 grid(148, B);
    mop::k_sum_sym<<<grid, 256, 0, stream>>>(n, H, Hbias, A);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  const size_t vbytes = sizeof(double) * (size_t)B * n;
  if (factored) {
    const size_t jac = al256(mop_jacobi_workspace_bytes(B, n));
    rc = mop_launch_eigh_large_factored(B, n, A, evals, evecs, status, ework + jac, ebytes - jac, stream);
    if (rc != MOP_OK) return rc;
    rc = mop_launch_eigh_jacobi(B, n, A, evals, evecs, status, status, ework, jac, stream);
    if (rc != MOP_OK) return rc;
    MOP_CHECK_CUDA(cudaMemcpyAsync(Xg, gp, vbytes, cudaMemcpyDeviceToDevice, stream));
    MOP_CHECK_CUDA(cudaMemcpyAsync(Xts, ts_vec, vbytes, cudaMemcpyDeviceToDevice, stream));
    rc = mop_launch_large_apply_q(B, n, 1, ework + jac, Xg, Xts, nullptr, nullptr, stream);
    if (rc != MOP_OK) return rc;
  } else {
    rc = mop_eigh(B, n, eigh_algo, A, evals, evecs, status, ework, ebytes, stream);
    if (rc != MOP_OK) return rc;
  }
  const int np = (n + 3) & ~3;
  const size_t smem = sizeof(double) * (7 * (size_t)np + 40) + sizeof(int) * 2 * (size_t)np + (size_t)np + 16;
  if (smem > 200 * 1024) {
    mop_set_error("mop_rsprfo_step: n = %d too large", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_prfo_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (factored) {
    mop::k_prfo_step<<<B, mop::PRFO_THREADS, smem, stream>>>(n, saddle_order, trust_min, trust_max, evals, evecs, Xg,
                                                           Bg, Be, state, prev_grad, Xpm, Xts, Xmv, eigvals_out,
                                                           pred_out, status);
    MOP_CHECK_CUDA(cudaGetLastError());
    const size_t jac = al256(mop_jacobi_workspace_bytes(B, n));
    rc = mop_launch_large_apply_q(B, n, 0, ework + jac, Xmv, Xts, nullptr, nullptr, stream);
    if (rc != MOP_OK) return rc;
    MOP_CHECK_CUDA(cudaMemcpyAsync(move_out, Xmv, vbytes, cudaMemcpyDeviceToDevice, stream));
    MOP_CHECK_CUDA(cudaMemcpyAsync(prev_move, Xmv, vbytes, cudaMemcpyDeviceToDevice, stream));
    MOP_CHECK_CUDA(cudaMemcpyAsync(ts_vec, Xts, vbytes, cudaMemcpyDeviceToDevice, stream));
    return MOP_OK;
  }
  mop::k_prfo_step<<<B, mop::PRFO_THREADS, smem, stream>>>(n, saddle_order, trust_min, trust_max, evals, evecs, gp,
                                                         Bg, Be, state, prev_grad, prev_move, ts_vec, move_out,
                                                         eigvals_out, pred_out, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
