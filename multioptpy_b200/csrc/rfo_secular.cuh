// RFO secular-equation machinery of RSIRFO (Optimizer/rsirfo.py:924-1313,
// 1374-1575, 1688-1715), executed by ONE WARP: every lane runs the same scalar
// control flow on identical values (butterfly reductions give all lanes the
// same sums), the O(k) term loops are strided over the lanes.
//
// The reference stops its iterations at loose tolerances and the alpha loop
// returns the step of whichever alpha it ended on, so the control flow is
// replayed iteration for iteration (SURVEY hazards H3, H8).
#pragma once
#include "common.cuh"

namespace mop {

struct RfoWork {
  const double* lam;  // kept eigenvalues, ascending, [k]
  const double* gam;  // gradient components in the kept eigenbasis, [k]
  double* lamp;       // scratch [k]: lam / alpha
  double* g2;         // scratch [k]: (gam / alpha)^2
  double* step;       // out [k]
  int k;
};

__device__ __forceinline__ double safe_den(double den, double tiny) {
  // np.where(|den| < tiny, sign(den) * tiny, den); exact zeros -> +tiny
  if (fabs(den) < tiny) den = sgn(den) * tiny;
  if (den == 0.0) den = tiny;
  return den;
}

__device__ __forceinline__ double sec_f(const RfoWork& w, double lmd, int lane) {
  double acc = 0.0;
  for (int i = lane; i < w.k; i += 32) acc += w.g2[i] / safe_den(w.lamp[i] - lmd, 1e-30);
  return lmd + warp_sum(acc);
}
__device__ __forceinline__ double sec_fp(const RfoWork& w, double lmd, int lane) {
  double acc = 0.0;
  for (int i = lane; i < w.k; i += 32) {
    const double d = safe_den(w.lamp[i] - lmd, 1e-30);
    acc += w.g2[i] / (d * d);
  }
  return 1.0 + warp_sum(acc);
}

// _solve_secular_safeguarded, rsirfo.py:1374-1503
__device__ __forceinline__ double secular_safeguarded(const RfoWork& w, double pole, double guess,
                                                     double gsum, int lane) {
  double b = pole, a = guess;
  double fa = sec_f(w, a, lane);
  const double gnorm = sqrt(gsum);
  int limit = 10;
  while (fa > 0.0 && limit > 0) {
    a = a - fmax(gnorm, fmax(fabs(a) * 0.1, 1e-8));
    fa = sec_f(w, a, lane);
    --limit;
  }
  if (fa > 0.0) return guess;
  double lk = guess;
  if (lk <= a || lk >= b) lk = (a + b) / 2.0;
  const double tol = 1e-10 * fabs(pole) + 1e-12;
  for (int it = 0; it < 250; ++it) {
    const double f = sec_f(w, lk, lane);
    if (fabs(f) < tol) return lk;
    const double fp = sec_fp(w, lk, lane);
    const double dn = fabs(fp) > 1e-20 ? -f / fp : 0.0;
    const double ln = lk + dn;
    const double lb = (a + b) / 2.0;
    const double nxt = (dn != 0.0 && ln > a && ln < b) ? ln : lb;
    if (f > 0.0) b = lk; else a = lk;
    lk = nxt;
    if (fabs(b - a) < tol) return (a + b) / 2.0;
  }
  return (a + b) / 2.0;
}

// solve_rfo + _solve_secular_more_sorensen, rsirfo.py:1505-1575,1688-1715.
// Writes w.step, returns lambda_aug (in the 1/alpha-scaled frame); *hard set when
// every gradient component is (numerically) zero.
__device__ __forceinline__ double solve_rfo(const RfoWork& w, double alpha, int lane, bool* hard) {
  double gs = 0.0;
  int first = 0x7fffffff;
  for (int i = lane; i < w.k; i += 32) {
    const double lp = w.lam[i] / alpha;
    const double gp = w.gam[i] / alpha;
    const double q = gp * gp;
    w.lamp[i] = lp;
    w.g2[i] = q;
    gs += q;
    if (q > 1e-20 && i < first) first = i;
  }
  __syncwarp();
  gs = warp_sum(gs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(MOP_FULL_MASK, first, o));
  double mu;
  if (first == 0x7fffffff) {
    mu = w.k > 0 ? w.lamp[0] : 0.0;
    if (hard) *hard = true;
  } else {
    const double pole = w.lamp[first];
    const double guess = 0.5 * (pole - sqrt(fmax(0.0, pole * pole + 4.0 * gs)));
    mu = secular_safeguarded(w, pole, guess, gs, lane);
  }
  for (int i = lane; i < w.k; i += 32) {
    const double den = safe_den(w.lam[i] / alpha - mu, 1e-20);
    w.step[i] = -(w.gam[i] / alpha) / den;
  }
  __syncwarp();
  return mu;
}

__device__ __forceinline__ double warp_norm2(const double* v, int k, int lane) {
  double acc = 0.0;
  for (int i = lane; i < k; i += 32) acc = fma(v[i], v[i], acc);
  return warp_sum(acc);
}

// get_step_derivative, rsirfo.py:1250-1313
__device__ __forceinline__ double step_derivative(const RfoWork& w, double alpha, double mu, int lane) {
  double acc = 0.0;
  int any_valid = 0;
  for (int i = lane; i < w.k; i += 32) {
    double den = w.lam[i] - mu * alpha;
    if (fabs(den) < 1e-8) den = sgn(den) * fmax(1e-8, fabs(den));  // exact zero stays zero
    const double d3 = den * den * den;
    if (fabs(d3) > 1e-10) {
      any_valid = 1;
      double t = (w.gam[i] * w.gam[i]) / d3;
      if (fabs(t) > 1e20) t = sgn(t) * 1e20;
      acc += t;
    }
  }
  any_valid = __any_sync(MOP_FULL_MASK, any_valid);
  if (!any_valid) return 1e-8;
  double d = 2.0 * mu * warp_sum(acc);
  if (!isfinite(d) || fabs(d) > 1e20) d = (d != 0.0) ? sgn(d) * 1e20 : 1e-8;
  return d;
}

// compute_rsprfo_step, rsirfo.py:986-1248.  Leaves the returned step in w.step.
// `best` is scratch [k].  Returns status bits to OR in.
//
// The step does not depend on alpha analytically (solve_rfo scales eigenvalues and gradient alike): the
// norms the loop sees differ by rounding only and every exit test (converged at once, alpha clipped to a bound, three
// equal norms) is decided with a wide margin - the replay is exact.  The exception: a secular root within ~1e-6 |pole|
// of its pole.  The loose stopping rule of the root finder (rsirfo.py:1437-1503) then lets the step norm jump between
// bisection histories by far more than the loop's 1e-6 test, and the micro-cycle the reference stops on is decided by
// last-bit differences of its BLAS / numpy sums (machine dependent).  The replay is still carried out, but a loop
// whose norms spread by more than 1e-7, or that converges / runs out after the first cycle, reports
// MOP_ST_ALPHA_UNSTABLE: the step is not reproducible to 1e-10, in the reference itself.
__device__ __forceinline__ int alpha_search(const RfoWork& w, double trust, double* best, int lane) {
  const double alpha0 = 1.0, alpha_max = 1000.0, alpha_step_max = 10.0, step_tol = 1e-3;
  const int max_micro = 40;
  int flags = 0;
  const double r2 = trust * trust;
  double alpha = alpha0;
  {
    solve_rfo(w, 1e-6, lane, nullptr);
    const double nlo = sqrt(warp_norm2(w.step, w.k, lane));
    solve_rfo(w, alpha_max, lane, nullptr);
    const double nhi = sqrt(warp_norm2(w.step, w.k, lane));
    const double olo = nlo * nlo - r2, ohi = nhi * nhi - r2;
    if (olo * ohi < 0.0) flags |= MOP_ST_BRENT_BRACKET;  // Brent branch not replayed (rounding-only event)
  }
  double hist0 = 0.0, hist1 = 0.0;  // last two recorded norms
  int nhist = 0;
  bool have_best = false;
  double best_diff = INFINITY;
  bool has_left = false, has_right = false;
  double a_left = 0.0, a_right = 0.0;
  double nmin = INFINITY, nmax = 0.0;
  for (int it = 0; it < max_micro; ++it) {
    const double mu = solve_rfo(w, alpha, lane, nullptr);
    const double nrm = sqrt(warp_norm2(w.step, w.k, lane));
    nmin = fmin(nmin, nrm);
    nmax = fmax(nmax, nrm);
    const int spread = (nmax - nmin > 1e-7) ? MOP_ST_ALPHA_UNSTABLE : 0;
    const double diff = fabs(nrm - trust);
    if (diff < best_diff) {
      for (int i = lane; i < w.k; i += 32) best[i] = w.step[i];
      best_diff = diff;
      have_best = true;
    }
    const double obj = nrm * nrm - r2;
    if (obj < 0.0 && (!has_left || alpha > a_left)) {
      a_left = alpha;
      has_left = true;
    } else if (obj > 0.0 && (!has_right || alpha < a_right)) {
      a_right = alpha;
      has_right = true;
    }
    if (fabs(obj) < 1e-8 || diff < step_tol) return it == 0 ? flags : (flags | MOP_ST_ALPHA_UNSTABLE);  // crossed the tolerance by rounding
    // history (fixed-size array in the reference; max_micro entries always fit)
    const double prev0 = hist0, prev1 = hist1;
    hist0 = hist1;
    hist1 = nrm;
    ++nhist;
    const double d = step_derivative(w, alpha, mu, lane);
    double a_new;
    if (fabs(d) < 1e-10) {
      if (has_left && has_right) a_new = (a_left + a_right) / 2.0;
      else if (obj > 0.0) a_new = fmax(alpha / 2.0, 1e-6);
      else a_new = fmin(alpha * 2.0, alpha_max);
    } else {
      const double a_step = fmin(alpha_step_max, fmax(-alpha_step_max, -obj / d));
      a_new = alpha + a_step;
      if (has_left && has_right) a_new = fmax(fmin(a_new, a_right * 0.99), a_left * 1.01);
    }
    alpha = fmin(fmax(a_new, 1e-6), alpha_max);
    if (alpha == alpha_max || alpha == 1e-6) return flags | spread;
    if (nhist >= 3 && fabs(hist1 - hist0) < 1e-6 && fabs(prev1 - prev0) < 1e-6) return flags | spread;
  }
  flags |= MOP_ST_ALPHA_UNSTABLE;
  // micro-cycles exhausted (rsirfo.py:1213-1246)
  if (have_best) {
    const double bn = sqrt(warp_norm2(best, w.k, lane));
    if (fabs(bn - trust) < step_tol * 1.1) {
      for (int i = lane; i < w.k; i += 32) w.step[i] = best[i];
      __syncwarp();
      return flags;
    }
  }
  const double gn = sqrt(warp_norm2(w.gam, w.k, lane));
  for (int i = lane; i < w.k; i += 32) w.step[i] = gn > 1e-10 ? -w.gam[i] / gn * trust : 0.0;
  __syncwarp();
  return flags;
}

// ------------------------------------------------------------------------------------------------------------------
// Register-resident variant for k <= 32 MAXJ kept modes (the fused shared-memory kernels, n <= 160): the scaled terms
// lam / alpha, gam / alpha of lane's modes i = lane + 32 u live in registers, f and f' of the secular function come
// from ONE pass and one pair of interleaved butterflies, nothing is staged through shared memory between the
// iterations.  Term order and reduction order are those of the functions above, so the two variants agree bit for
// bit.  Because a warp needs no scratch, the two bracket probes of compute_rsprfo_step (alpha = 1e-6, alpha_max;
// rsirfo.py:1023-1034 - they only decide whether Brent's method would have run) can run on other warps while warp 0
// replays the Newton loop, and the loop's first micro-cycle reuses the alpha0 = 1 solve of get_rs_step
// (rsirfo.py:931, :1087 - the same call on the same arguments).
template <int MAXJ>
struct RfoTerms {
  const double* lam;  // kept eigenvalues, ascending, [k] (shared memory)
  const double* gam;  // gradient components, [k]
  int k;
  double lp[MAXJ], gp[MAXJ];  // lam / alpha, gam / alpha of this lane's modes
};

template <int MAXJ, bool WITH_FP>
__device__ __forceinline__ double sec_ffp_r(const RfoTerms<MAXJ>& w, double lmd, int lane, double* fp) {
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int u = 0; u < MAXJ; ++u) {
    if (lane + 32 * u < w.k) {
      const double d = safe_den(w.lp[u] - lmd, 1e-30);
      const double q = w.gp[u] * w.gp[u];
      a0 += q / d;
      if (WITH_FP) a1 += q / (d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(MOP_FULL_MASK, a0, o);
    if (WITH_FP) a1 += __shfl_xor_sync(MOP_FULL_MASK, a1, o);
  }
  if (WITH_FP) *fp = 1.0 + a1;
  return lmd + a0;
}

template <int MAXJ>
__device__ __forceinline__ double secular_safeguarded_r(const RfoTerms<MAXJ>& w, double pole, double guess, double gsum,
                                                       int lane) {
  double b = pole, a = guess;
  double fa = sec_ffp_r<MAXJ, false>(w, a, lane, nullptr);
  const double gnorm = sqrt(gsum);
  int limit = 10;
  while (fa > 0.0 && limit > 0) {
    a = a - fmax(gnorm, fmax(fabs(a) * 0.1, 1e-8));
    fa = sec_ffp_r<MAXJ, false>(w, a, lane, nullptr);
    --limit;
  }
  if (fa > 0.0) return guess;
  double lk = guess;
  if (lk <= a || lk >= b) lk = (a + b) / 2.0;
  const double tol = 1e-10 * fabs(pole) + 1e-12;
  for (int it = 0; it < 250; ++it) {
    double fp;
    const double f = sec_ffp_r<MAXJ, true>(w, lk, lane, &fp);
    if (fabs(f) < tol) return lk;
    const double dn = fabs(fp) > 1e-20 ? -f / fp : 0.0;
    const double ln = lk + dn;
    const double lb = (a + b) / 2.0;
    const double nxt = (dn != 0.0 && ln > a && ln < b) ? ln : lb;
    if (f > 0.0) b = lk; else a = lk;
    lk = nxt;
    if (fabs(b - a) < tol) return (a + b) / 2.0;
  }
  return (a + b) / 2.0;
}

// solve_rfo: step[u] = the step component of mode lane + 32 u, *nrm2 = ||step||^2; returns lambda_aug.
template <int MAXJ>
__device__ __forceinline__ double solve_rfo_r(RfoTerms<MAXJ>& w, double alpha, int lane, bool* hard,
                                              double (&step)[MAXJ], double* nrm2) {
  double gs = 0.0;
  int first = 0x7fffffff;
#pragma unroll
  for (int u = 0; u < MAXJ; ++u) {
    const int i = lane + 32 * u;
    w.lp[u] = 0.0;
    w.gp[u] = 0.0;
    if (i < w.k) {
      w.lp[u] = w.lam[i] / alpha;
      w.gp[u] = w.gam[i] / alpha;
      const double q = w.gp[u] * w.gp[u];
      gs += q;
      if (q > 1e-20 && i < first) first = i;
    }
  }
  gs = warp_sum(gs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(MOP_FULL_MASK, first, o));
  double mu;
  if (first == 0x7fffffff) {
    mu = w.k > 0 ? w.lam[0] / alpha : 0.0;
    if (hard) *hard = true;
  } else {
    const double pole = w.lam[first] / alpha;
    const double guess = 0.5 * (pole - sqrt(fmax(0.0, pole * pole + 4.0 * gs)));
    mu = secular_safeguarded_r<MAXJ>(w, pole, guess, gs, lane);
  }
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < MAXJ; ++u) {
    step[u] = 0.0;
    if (lane + 32 * u < w.k) {
      const double den = safe_den(w.lp[u] - mu, 1e-20);
      step[u] = -w.gp[u] / den;
      acc = fma(step[u], step[u], acc);
    }
  }
  *nrm2 = warp_sum(acc);
  return mu;
}

template <int MAXJ>
__device__ __forceinline__ double step_derivative_r(const RfoTerms<MAXJ>& w, double alpha, double mu, int lane) {
  double acc = 0.0;
  int any_valid = 0;
#pragma unroll
  for (int u = 0; u < MAXJ; ++u) {
    const int i = lane + 32 * u;
    if (i < w.k) {
      double den = w.lam[i] - mu * alpha;
      if (fabs(den) < 1e-8) den = sgn(den) * fmax(1e-8, fabs(den));  // exact zero stays zero
      const double d3 = den * den * den;
      if (fabs(d3) > 1e-10) {
        any_valid = 1;
        double t = (w.gam[i] * w.gam[i]) / d3;
        if (fabs(t) > 1e20) t = sgn(t) * 1e20;
        acc += t;
      }
    }
  }
  any_valid = __any_sync(MOP_FULL_MASK, any_valid);
  if (!any_valid) return 1e-8;
  double d = 2.0 * mu * warp_sum(acc);
  if (!isfinite(d) || fabs(d) > 1e20) d = (d != 0.0) ? sgn(d) * 1e20 : 1e-8;
  return d;
}

// One bracket probe of compute_rsprfo_step: ||step(alpha)||^2 - trust^2 (any warp).
template <int MAXJ>
__device__ __forceinline__ double alpha_probe_r(const double* lam, const double* gam, int k, double alpha, double trust,
                                                int lane) {
  RfoTerms<MAXJ> w;
  w.lam = lam;
  w.gam = gam;
  w.k = k;
  double step[MAXJ], n2;
  solve_rfo_r<MAXJ>(w, alpha, lane, nullptr, step, &n2);
  const double nrm = sqrt(n2);
  return nrm * nrm - trust * trust;
}

// The Newton loop of compute_rsprfo_step (alpha_search above without the two probes).  step / mu0 / nrm0: the
// alpha0 = 1 solve, reused as the first micro-cycle; on return step holds the step to use.  best: scratch [k].
template <int MAXJ>
__device__ __forceinline__ int alpha_newton_r(RfoTerms<MAXJ>& w, double trust, double mu0, double nrm0,
                                              double (&step)[MAXJ], double* best, int lane) {
  const double alpha0 = 1.0, alpha_max = 1000.0, alpha_step_max = 10.0, step_tol = 1e-3;
  const int max_micro = 40;
  int flags = 0;
  const double r2 = trust * trust;
  double alpha = alpha0;
  double hist0 = 0.0, hist1 = 0.0;
  int nhist = 0;
  bool have_best = false;
  double best_diff = INFINITY;
  bool has_left = false, has_right = false;
  double a_left = 0.0, a_right = 0.0;
  double nmin = INFINITY, nmax = 0.0;
  for (int it = 0; it < max_micro; ++it) {
    double mu = mu0, nrm = nrm0;
    if (it > 0) {
      double n2;
      mu = solve_rfo_r<MAXJ>(w, alpha, lane, nullptr, step, &n2);
      nrm = sqrt(n2);
    }
    nmin = fmin(nmin, nrm);
    nmax = fmax(nmax, nrm);
    const int spread = (nmax - nmin > 1e-7) ? MOP_ST_ALPHA_UNSTABLE : 0;
    const double diff = fabs(nrm - trust);
    if (diff < best_diff) {
#pragma unroll
      for (int u = 0; u < MAXJ; ++u)
        if (lane + 32 * u < w.k) best[lane + 32 * u] = step[u];
      best_diff = diff;
      have_best = true;
    }
    const double obj = nrm * nrm - r2;
    if (obj < 0.0 && (!has_left || alpha > a_left)) {
      a_left = alpha;
      has_left = true;
    } else if (obj > 0.0 && (!has_right || alpha < a_right)) {
      a_right = alpha;
      has_right = true;
    }
    if (fabs(obj) < 1e-8 || diff < step_tol) return it == 0 ? flags : (flags | MOP_ST_ALPHA_UNSTABLE);
    const double prev0 = hist0, prev1 = hist1;
    hist0 = hist1;
    hist1 = nrm;
    ++nhist;
    const double d = step_derivative_r<MAXJ>(w, alpha, mu, lane);
    double a_new;
    if (fabs(d) < 1e-10) {
      if (has_left && has_right) a_new = (a_left + a_right) / 2.0;
      else if (obj > 0.0) a_new = fmax(alpha / 2.0, 1e-6);
      else a_new = fmin(alpha * 2.0, alpha_max);
    } else {
      const double a_step = fmin(alpha_step_max, fmax(-alpha_step_max, -obj / d));
      a_new = alpha + a_step;
      if (has_left && has_right) a_new = fmax(fmin(a_new, a_right * 0.99), a_left * 1.01);
    }
    alpha = fmin(fmax(a_new, 1e-6), alpha_max);
    if (alpha == alpha_max || alpha == 1e-6) return flags | spread;
    if (nhist >= 3 && fabs(hist1 - hist0) < 1e-6 && fabs(prev1 - prev0) < 1e-6) return flags | spread;
  }
  flags |= MOP_ST_ALPHA_UNSTABLE;
  // micro-cycles exhausted (rsirfo.py:1213-1246)
  __syncwarp();
  if (have_best) {
    const double bn = sqrt(warp_norm2(best, w.k, lane));
    if (fabs(bn - trust) < step_tol * 1.1) {
#pragma unroll
      for (int u = 0; u < MAXJ; ++u)
        if (lane + 32 * u < w.k) step[u] = best[lane + 32 * u];
      return flags;
    }
  }
  const double gn = sqrt(warp_norm2(w.gam, w.k, lane));
#pragma unroll
  for (int u = 0; u < MAXJ; ++u)
    if (lane + 32 * u < w.k) step[u] = gn > 1e-10 ? -w.gam[lane + 32 * u] / gn * trust : 0.0;
  return flags;
}

}  // namespace mop
