// Hyper-dual numbers (f, df/da, df/db, d2f/da db): one evaluation of a scalar function with the seeds
// a, b set on two input coordinates yields one entry of its Hessian (and two of its gradient) exactly.
// Used where the reference differentiates small closed-form expressions with torch.func (RIC second
// derivatives, restraint bias potentials).
#pragma once
#include "common.cuh"

namespace mop {

struct HD {
  double f, a, b, ab;
};
__device__ __forceinline__ HD hd_const(double x) { return HD{x, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ HD operator+(HD x, HD y) { return HD{x.f + y.f, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
__device__ __forceinline__ HD operator-(HD x, HD y) { return HD{x.f - y.f, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
__device__ __forceinline__ HD operator*(HD x, HD y) {
  return HD{x.f * y.f, x.a * y.f + x.f * y.a, x.b * y.f + x.f * y.b, x.ab * y.f + x.a * y.b + x.b * y.a + x.f * y.ab};
}
__device__ __forceinline__ HD hd_unary(HD x, double g, double g1, double g2) {
  return HD{g, g1 * x.a, g1 * x.b, g1 * x.ab + g2 * x.a * x.b};
}
__device__ __forceinline__ HD hd_recip(HD x) { const double r = 1.0 / x.f; return hd_unary(x, r, -r * r, 2.0 * r * r * r); }
__device__ __forceinline__ HD operator/(HD x, HD y) { return x * hd_recip(y); }
__device__ __forceinline__ HD hd_sqrt(HD x) { const double s = sqrt(x.f); return hd_unary(x, s, 0.5 / s, -0.25 / (s * x.f)); }
__device__ __forceinline__ HD hd_acos(HD x) {
  const double om = 1.0 - x.f * x.f, s = sqrt(om);
  return hd_unary(x, acos(x.f), -1.0 / s, -x.f / (s * om));
}
// atan2(y, x): f_y = x / r2, f_x = -y / r2, f_yy = -2 x y / r2^2 = -f_xx, f_xy = (y^2 - x^2) / r2^2
__device__ __forceinline__ HD hd_atan2(HD y, HD x) {
  const double r2 = x.f * x.f + y.f * y.f, ir2 = 1.0 / r2;
  const double fy = x.f * ir2, fx = -y.f * ir2;
  const double fyy = -2.0 * x.f * y.f * ir2 * ir2, fxy = (y.f * y.f - x.f * x.f) * ir2 * ir2;
  return HD{atan2(y.f, x.f), fy * y.a + fx * x.a, fy * y.b + fx * x.b,
            fy * y.ab + fx * x.ab + fyy * (y.a * y.b - x.a * x.b) + fxy * (y.a * x.b + x.a * y.b)};
}
__device__ __forceinline__ HD hd_abs(HD x) { return x.f < 0.0 ? HD{-x.f, -x.a, -x.b, -x.ab} : x; }
__device__ __forceinline__ HD hd_dot(const HD* u, const HD* v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; }
__device__ __forceinline__ void hd_cross(const HD* u, const HD* v, HD* c) {
  c[0] = u[1] * v[2] - u[2] * v[1];
  c[1] = u[2] * v[0] - u[0] * v[2];
  c[2] = u[0] * v[1] - u[1] * v[0];
}

__device__ __forceinline__ HD operator*(double c, HD x) { return HD{c * x.f, c * x.a, c * x.b, c * x.ab}; }
__device__ __forceinline__ HD hd_clamp_min(HD x, double lo) { return x.f < lo ? hd_const(lo) : x; }   // torch.clamp(min=)
__device__ __forceinline__ HD hd_clamp(HD x, double lo, double hi) { return x.f < lo ? hd_const(lo) : (x.f > hi ? hd_const(hi) : x); }

}  // namespace mop
