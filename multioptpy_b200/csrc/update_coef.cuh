// Quasi-Newton Hessian-update algebra in fused rank-k form.
//
// Every update of Optimizer/hessian_update.py and the (always depth-1)
// Optimizer/block_hessian_update.py is  delta = sum_ab C[a][b] * v_a v_b^T  over
// the four vectors  v0 = s, v1 = y (after optional Powell damping), v2 = u = H s,
// v3 = r = y - u.  This header turns the scalar products into the symmetric
// 4x4 coefficient matrix C, reproducing every guard of the reference.
#pragma once
#include "common.cuh"

namespace mop {

struct UpdScalars {
  double ss, sy, su, rs, rr;  // s.s, s.y, s.(Hs), r.s, r.r   (y already damped)
};

struct UpdCoef {
  double c[4][4];
  int flags;  // MOP_ST_UPD_TERM_ZEROED when a guarded term was dropped
};

__device__ __forceinline__ void coef_zero(UpdCoef& k) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) k.c[a][b] = 0.0;
  k.flags = 0;
}

// hessian_update.py:35-65 (tau = 1e-10)
__device__ __forceinline__ void add_bfgs(UpdCoef& k, const UpdScalars& q, double w) {
  if (fabs(q.sy) >= 1e-10) k.c[1][1] += w / q.sy; else k.flags |= MOP_ST_UPD_TERM_ZEROED;
  if (fabs(q.su) >= 1e-10) k.c[2][2] -= w / q.su; else k.flags |= MOP_ST_UPD_TERM_ZEROED;
}
// hessian_update.py:67-85 with A = cf * r
__device__ __forceinline__ void add_sr1(UpdCoef& k, const UpdScalars& q, double cf, double w) {
  const double den = cf * q.rs;
  if (fabs(den) >= 1e-10) k.c[3][3] += w * cf * cf / den; else k.flags |= MOP_ST_UPD_TERM_ZEROED;
}
// hessian_update.py:87-104
__device__ __forceinline__ void add_psb(UpdCoef& k, const UpdScalars& q, double w) {
  if (fabs(q.ss) >= 1e-10) {
    k.c[0][0] -= w * q.rs / (q.ss * q.ss);
    k.c[0][3] += w / q.ss;
    k.c[3][0] += w / q.ss;
  } else {
    k.flags |= MOP_ST_UPD_TERM_ZEROED;
  }
}
// hessian_update.py:106-130 with A = cf * r
__device__ __forceinline__ double phi2(const UpdScalars& q, double cf, int& flags) {
  const double as = cf * q.rs;
  const double den = (cf * cf * q.rr) * q.ss;
  if (fabs(den) >= 1e-10) return (as * as) / den;
  flags |= MOP_ST_UPD_TERM_ZEROED;
  return 0.0;
}

// 1x1 numpy.linalg.inv with the safe_inv fallback (block_hessian_update.py:12-21)
__device__ __forceinline__ double inv1(double x) { return x == 0.0 ? 1.0 / (x + 1e-10) : 1.0 / x; }

// block_hessian_update.py:75-118 (q = 1); `a` scales s, y, u ("weighted subspace").
__device__ __forceinline__ void add_blk_bfgs(UpdCoef& k, const UpdScalars& q, double a, bool guard,
                                             double w) {
  if (!(fabs(a) * sqrt(q.ss) > 1e-8)) return;
  if (guard && (a * a * q.sy) <= 1e-12) return;
  k.c[2][2] -= w * a * a * inv1(a * a * q.su);
  k.c[1][1] += w * a * a * inv1(a * a * q.sy);
}
// block_hessian_update.py:159-184 (q = 1)
__device__ __forceinline__ void add_blk_sr1(UpdCoef& k, const UpdScalars& q, double cf, double a,
                                            double w) {
  k.c[3][3] += w * (a * cf) * (a * cf) * inv1(a * a * cf * q.rs);
}
// block_hessian_update.py:120-157 (q = 1, threshold 1e-8)
__device__ __forceinline__ void add_blk_psb(UpdCoef& k, const UpdScalars& q, double a, double w) {
  if (!(fabs(a) * sqrt(q.ss) > 1e-8)) return;
  const double ss = a * a * q.ss;
  if (fabs(ss) >= 1e-8) {
    const double a2 = a * a;
    k.c[0][0] -= w * (a2 * q.rs) * a2 / (ss * ss);
    k.c[0][3] += w * a2 / ss;
    k.c[3][0] += w * a2 / ss;
  }
}
// block_hessian_update.py:190-231 (q = 1)
__device__ __forceinline__ double blk_weight(const UpdScalars& q, bool cfd) {
  const double cf = cfd ? 2.0 : 1.0;
  const double as = cf * q.rs;
  const double den = (cf * cf * q.rr) * q.ss;
  double c = fabs(den) > 1e-12 ? (as * as) / den : 0.0;
  if (c != c) c = 0.0;
  return fmax(0.0, fmin(1.0, c));
}

// Powell damping with B = I (hessian_update.py:200-242, block_..py:565-595).
// Returns theta; y_tilde = theta*y + (1-theta)*s; theta == 1 means "inactive".
__device__ __forceinline__ double dd_theta(double ss, double sy, double thr) {
  if (sy < 0.2 * ss) {
    const double den = ss - sy;
    double th = fabs(den) < thr ? 0.1 : 0.8 * ss / den;
    return fmax(0.0, fmin(1.0, th));
  }
  return 1.0;
}
__device__ __forceinline__ bool method_has_dd(int m) {
  return m == MOP_UPD_BFGS_DD || m == MOP_UPD_FSB_DD || m == MOP_UPD_CFD_FSB_DD ||
         m == MOP_UPD_BLOCK_BFGS_DD || m == MOP_UPD_BLOCK_FSB_DD || m == MOP_UPD_BLOCK_CFD_FSB_DD;
}
__device__ __forceinline__ double method_dd_thr(int m) {
  return (m == MOP_UPD_BFGS_DD || m == MOP_UPD_FSB_DD || m == MOP_UPD_CFD_FSB_DD) ? 1e-10 : 1e-12;
}

// The coefficient matrix for a concrete (non-flowchart) method.
// msp_arg: s.A / (|A||s|) pre-clipped input is derived here from q.
__device__ __forceinline__ void update_coefficients(int m, const UpdScalars& q, UpdCoef& k) {
  coef_zero(k);
  switch (m) {
    case MOP_UPD_BFGS:
    case MOP_UPD_BFGS_DD:
      add_bfgs(k, q, 1.0);
      break;
    case MOP_UPD_SR1:
      add_sr1(k, q, 1.0, 1.0);
      break;
    case MOP_UPD_PSB:
      add_psb(k, q, 1.0);
      break;
    case MOP_UPD_FSB:
    case MOP_UPD_FSB_DD:
    case MOP_UPD_CFD_FSB:
    case MOP_UPD_CFD_FSB_DD: {
      const double cf = (m == MOP_UPD_CFD_FSB || m == MOP_UPD_CFD_FSB_DD) ? 2.0 : 1.0;
      // reference order: SR1 delta, BFGS delta, then the Bofill constant
      UpdCoef t;
      coef_zero(t);
      const double phi = sqrt(phi2(q, cf, k.flags));
      add_sr1(k, q, cf, phi);
      add_bfgs(k, q, 1.0 - phi);
      break;
    }
    case MOP_UPD_BOFILL:
    case MOP_UPD_CFD_BOFILL: {
      const double cf = m == MOP_UPD_CFD_BOFILL ? 2.0 : 1.0;
      const double p2 = phi2(q, cf, k.flags);
      add_psb(k, q, 1.0 - p2);
      add_sr1(k, q, cf, p2);
      break;
    }
    case MOP_UPD_MSP: {  // hessian_update.py:345-368
      const double den = sqrt(q.rr) * sqrt(q.ss);
      double arg = 0.0;
      if (den >= 1e-10) arg = fmax(-1.0, fmin(1.0, q.rs / den));
      const double phi = 1.0 - arg * arg;
      add_psb(k, q, phi);
      add_sr1(k, q, 1.0, 1.0 - phi);
      break;
    }
    case MOP_UPD_BLOCK_BFGS:
      add_blk_bfgs(k, q, 1.0, true, 1.0);
      break;
    case MOP_UPD_BLOCK_BFGS_DD:  // rank guard, damping, no curvature guard (:619-641)
      add_blk_bfgs(k, q, 1.0, false, 1.0);
      break;
    case MOP_UPD_BLOCK_FSB:
    case MOP_UPD_BLOCK_FSB_DD:
    case MOP_UPD_BLOCK_CFD_FSB:
    case MOP_UPD_BLOCK_CFD_FSB_DD: {
      const bool cfd = (m == MOP_UPD_BLOCK_CFD_FSB || m == MOP_UPD_BLOCK_CFD_FSB_DD);
      const double c = blk_weight(q, cfd);
      const double w = cfd ? c : sqrt(c);  // CFD-FSB mixes with c, FSB with sqrt(c)
      add_blk_sr1(k, q, cfd ? 2.0 : 1.0, 1.0, w);
      add_blk_bfgs(k, q, 1.0, true, 1.0 - w);
      break;
    }
    case MOP_UPD_BLOCK_BOFILL:
    case MOP_UPD_BLOCK_CFD_BOFILL: {
      const bool cfd = m == MOP_UPD_BLOCK_CFD_BOFILL;
      const double w = blk_weight(q, cfd);
      add_blk_sr1(k, q, cfd ? 2.0 : 1.0, 1.0, w);
      add_blk_psb(k, q, 1.0, 1.0 - w);
      break;
    }
    case MOP_UPD_BLOCK_FSB_WEIGHTED:
    case MOP_UPD_BLOCK_CFD_FSB_WEIGHTED:
    case MOP_UPD_BLOCK_BOFILL_WEIGHTED:
    case MOP_UPD_BLOCK_CFD_BOFILL_WEIGHTED: {  // block_hessian_update.py:319-437
      const bool cfd = (m == MOP_UPD_BLOCK_CFD_FSB_WEIGHTED || m == MOP_UPD_BLOCK_CFD_BOFILL_WEIGHTED);
      const bool fsb = (m == MOP_UPD_BLOCK_FSB_WEIGHTED || m == MOP_UPD_BLOCK_CFD_FSB_WEIGHTED);
      const double c = blk_weight(q, cfd);
      const double w = (m == MOP_UPD_BLOCK_FSB_WEIGHTED) ? sqrt(c) : c;
      add_blk_sr1(k, q, cfd ? 2.0 : 1.0, w, 1.0);
      if (fsb) add_blk_bfgs(k, q, 1.0 - w, true, 1.0);
      else add_blk_psb(k, q, 1.0 - w, 1.0);
      break;
    }
    default:
      break;
  }
}

// Flowchart selection (hessian_update.py:163-194): z = y - H y (sic).
__device__ __forceinline__ int flowchart_select(double ss, double yy, double sy, double zz, double zs) {
  double zden = sqrt(ss) * sqrt(zz);
  if (fabs(zden) < 1e-10) zden += 1e-10;
  double yden = sqrt(ss) * sqrt(yy);
  if (fabs(yden) < 1e-10) yden += 1e-10;
  if (zs / zden < -0.1) return MOP_UPD_SR1;
  if (sy / yden > 0.1) return MOP_UPD_BFGS;
  return MOP_UPD_FSB;
}

__device__ __forceinline__ double coef_delta(const UpdCoef& k, const double vi[4], const double vj[4]) {
  double d = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) t = fma(k.c[a][b], vj[b], t);
    d = fma(vi[a], t, d);
  }
  return d;
}

}  // namespace mop
