// Restraint bias potentials: energy, gradient and Hessian of
//   kind 1  StructKeepPotential     (Potential/keep_potential.py:5-61)   E = 1/2 k (|x_i - x_j| - r0)^2
//   kind 2  StructKeepPotentialv2   (keep_potential.py:64-116)           same between two fragment centroids
//   kind 3  StructKeepAnglePotential (keep_angle_potential.py:7-229)     E = 1/2 k (theta - theta0)^2 with the
//           reference's fifth-order expansions of acos^2 within 1e-3 rad of 0 and pi and its three
//           theta0 branches
//   kind 4  StructKeepDihedralAnglePotential (keep_dihedral_angle_potential.py:6-154)  E = 1/2 k wrap(phi - phi0)^2
//           S(|n1|^2) S(|n2|^2) with the smoothstep collinearity switch; phi0 arrives in RADIANS (the reference
//           converts degrees in float32 or float64 depending on the caller - the host mirror reproduces that)
//   kind 5  one atom pair of LJRepulsivePotentialScale / Value (LJ_repulsive_potential.py:9-114)
//           E = eps (-2 (sigma / r)^6 + (sigma / r)^12); the host expands the fragment product and derives eps, sigma
//           (the "scale" unit in the reference's float32 arithmetic)
//   kind 6  StructAnharmonicKeepPotential (anharmonic_keep_potential.py:14-27)  E = D (1 - exp(-sqrt(k / 2D) (r - r0)))^2
//   kind 7  WellPotential (switching_potential.py:5-67)  flat-bottomed well between two fragment centroids with
//           quintic switching walls: limits a < b < c < d
//   kind 8  StructKeepOutofPlainAnglePotential (keep_outofplain_angle_potential.py:6-146)  E = 1/2 k (phi - phi0)^2,
//           phi = atan2(a1 . n, sqrt(|a1|^2 - (a1 . n)^2)) the elevation of a1 over the plane (a2, a3); zero when the
//           plane is undefined (|a2 x a3|^2 < 1e-8)
//   kind 9  StructKeepAnglePotentialv2 (keep_angle_potential.py:226-478)  the angle between three fragment CENTROIDS;
//           within 1e-3 rad of 0 / pi it continues 1/2 k (theta - theta0)^2 by a quadratic in cos(theta) matched in value
//           and slope at the cut (Gauss-Newton curvature), except at an exactly linear / collapsed theta0 where the
//           fifth-order expansion of kind 3 is the energy itself
//   kind 10 StructKeepDihedralAnglePotentialv2 (keep_dihedral_angle_potential.py:156-257)  kind 4 on four centroids
//   kind 11 StructKeepOutofPlainAnglePotentialv2 (keep_outofplain_angle_potential.py:148-290)  kind 8 on four centroids
//           (kinds 9-11: q[g] = number of atoms of fragment g, the fragments one after the other in `atoms`)
//   kind 12 one target atom of WellPotentialVP (switching_potential.py:121-170): the well of kind 7 in the distance to a
//           fixed point;  kind 13 one target atom of WellPotentialWall (:69-119): the well in |x_axis| (n2 = axis);
//           WellPotentialAround (:172-224) is kind 7 per target atom (fragment 1 = the atom, fragment 2 = the centre atoms)
// The reference differentiates calc_energy with torch.func.jacrev / hessian on the CPU
// (Potential/potential.py:127-137); here thread (term, coordinate pair) evaluates the same expression once in
// hyper-dual arithmetic.  Results are ADDED to E, grad, hess (the aggregator sums all bias terms).
#include "hyperdual.cuh"

namespace mop {

constexpr int BIAS_MAXA = 64;  // atoms per term (both fragments together)

struct BiasTerm {
  int kind;
  int n1, n2;             // atoms in fragment 1 / 2 (kind 1: 1, 1; kind 3 / 4: atoms i, j, k (, l) in `atoms`, n1 = 3 / 4)
  int atoms[BIAS_MAXA];   // 0-based
  double k, p;            // spring constant; r0 in Angstrom (kinds 1, 2), theta0 in degrees (kind 3), phi0 in radians (kind 4)
  double q[4];            // kind 6: q[0] = well depth; kinds 7, 12, 13: k = wall energy (Hartree), q = a, b, c, d (Bohr);
                          // kind 5: k = eps (Hartree), p = sigma (Bohr); kind 8: p = phi0 in radians
  double r3[3];           // kind 12: the void point (Bohr, already rounded to float32 as the reference stores it)
  double pad_;
};

__device__ __forceinline__ HD hd_exp(HD x) { const double e = exp(x.f); return hd_unary(x, e, e, e); }
__device__ __forceinline__ HD hd_pow_int(HD x, int n) {
  HD r = hd_const(1.0);
  for (int i = 0; i < n; ++i) r = r * x;
  return r;
}

__device__ HD bias_energy(const BiasTerm& t, const double* xyz, int ca, int cb) {
  // coordinate index c = 3 * (position in t.atoms) + component; seeds on ca, cb
  auto X = [&](int pos, int comp) -> HD {
    const int c = 3 * pos + comp;
    return HD{xyz[3 * t.atoms[pos] + comp], c == ca ? 1.0 : 0.0, c == cb ? 1.0 : 0.0, 0.0};
  };
  const double BOHR2ANG = 0.52917721067;
  // point g of an angle / dihedral / out-of-plane term: atom g (kinds 3, 4, 8) or the centroid of fragment g (9-11)
  const bool grouped = t.kind >= 9 && t.kind <= 11;
  int goff[5] = {0, 0, 0, 0, 0};
  if (grouped)
    for (int g = 0; g < 4; ++g) goff[g + 1] = goff[g] + (int)t.q[g];
  auto P = [&](int g, int comp) -> HD {
    if (!grouped) return X(g, comp);
    HD sum = hd_const(0.0);
    for (int a = goff[g]; a < goff[g + 1]; ++a) sum = sum + X(a, comp);
    return (1.0 / (goff[g + 1] - goff[g])) * sum;
  };
  if (t.kind == 1 || t.kind == 2) {
    HD v[3];
    for (int c = 0; c < 3; ++c) {
      HD s1 = hd_const(0.0), s2 = hd_const(0.0);
      for (int a = 0; a < t.n1; ++a) s1 = s1 + X(a, c);
      for (int a = 0; a < t.n2; ++a) s2 = s2 + X(t.n1 + a, c);
      v[c] = (1.0 / t.n1) * s1 - (1.0 / t.n2) * s2;
    }
    const HD d = hd_clamp_min(hd_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12);
    const HD diff = d - hd_const(t.p / BOHR2ANG);
    return (0.5 * t.k) * (diff * diff);
  }
  if (t.kind == 5 || t.kind == 6) {  // atom pair
    HD v[3];
    for (int c = 0; c < 3; ++c) v[c] = X(0, c) - X(1, c);
    const HD r = hd_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (t.kind == 5) {
      const HD x = hd_const(t.p) / r;
      const HD x2 = x * x, x6 = x2 * x2 * x2;
      return t.k * (x6 * x6 - 2.0 * x6);
    }
    const HD ex = hd_exp((-sqrt(t.k / (2.0 * t.q[0]))) * (r - hd_const(t.p / BOHR2ANG)));
    const HD om = hd_const(1.0) - ex;
    return t.q[0] * (om * om);
  }
  if (t.kind == 7 || t.kind == 12 || t.kind == 13) {
    HD r;
    if (t.kind == 7) {
      HD v[3];
      for (int c = 0; c < 3; ++c) {
        HD s1 = hd_const(0.0), s2 = hd_const(0.0);
        for (int a = 0; a < t.n1; ++a) s1 = s1 + X(a, c);
        for (int a = 0; a < t.n2; ++a) s2 = s2 + X(t.n1 + a, c);
        v[c] = (1.0 / t.n1) * s1 - (1.0 / t.n2) * s2;
      }
      r = hd_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    } else if (t.kind == 12) {
      HD v[3];
      for (int c = 0; c < 3; ++c) v[c] = X(0, c) - hd_const(t.r3[c]);
      r = hd_sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    } else {
      const HD x = X(0, t.n2);
      r = x.f < 0.0 ? hd_const(0.0) - x : x;   // |x|: torch.linalg.norm of a scalar
    }
    const double a = t.q[0], b = t.q[1], c = t.q[2], d = t.q[3];
    const HD xs = (0.5 / (b - a)) * r + hd_const(1.0 - 0.5 * b / (b - a));
    const HD xl = (0.5 / (c - d)) * r + hd_const(1.0 - 0.5 * c / (c - d));
    auto wall = [](HD x) -> HD {  // 2 - 20 x^3 + 30 x^4 - 12 x^5
      const HD x3 = x * x * x;
      return hd_const(2.0) - 20.0 * x3 + 30.0 * (x3 * x) - 12.0 * (x3 * x * x);
    };
    if (r.f <= a) return t.k * (hd_const(2.875) - 3.75 * xs);
    if (r.f <= b) return t.k * wall(xs);
    if (r.f < c) return hd_const(0.0);
    if (r.f < d) return t.k * wall(xl);
    return t.k * (hd_const(2.875) - 3.75 * xl);
  }
  if (t.kind == 8 || t.kind == 11) {
    HD a1[3], a2[3], a3[3], nn[3];
    for (int c = 0; c < 3; ++c) {
      const HD p0 = P(0, c);
      a1[c] = P(1, c) - p0;
      a2[c] = P(2, c) - p0;
      a3[c] = P(3, c) - p0;
    }
    hd_cross(a2, a3, nn);
    const HD nsq = hd_dot(nn, nn);
    if (nsq.f < 1e-8) return hd_const(0.0);  // torch.where(is_undefined_plane, 0, .)
    const HD inn = hd_recip(hd_clamp_min(hd_sqrt(nsq), 1e-12));
    const HD h = hd_dot(a1, nn) * inn;
    const HD a1sq = hd_dot(a1, a1);          // (|a1|)^2 of the reference: sqrt then square, same value to rounding
    const HD rp = hd_sqrt(hd_clamp_min(a1sq - h * h, 0.0));
    const HD diff = hd_atan2(h, rp) - hd_const(t.p);
    return (0.5 * t.k) * (diff * diff);
  }
  const double PI = 3.141592653589793;
  if (t.kind == 4 || t.kind == 10) {
    HD b1[3], b2[3], b3[3], n1[3], n2[3], m1[3];
    for (int c = 0; c < 3; ++c) {
      const HD p1 = P(1, c), p2 = P(2, c);
      b1[c] = p1 - P(0, c);
      b2[c] = p2 - p1;
      b3[c] = P(3, c) - p2;
    }
    hd_cross(b1, b2, n1);
    hd_cross(b2, b3, n2);
    const HD n1sq = hd_dot(n1, n1), n2sq = hd_dot(n2, n2);
    auto sw = [](HD val) -> HD {  // smoothstep on [1e-10, 1e-8]
      HD tt = hd_clamp((1.0 / (1e-8 - 1e-10)) * (val - hd_const(1e-10)), 0.0, 1.0);
      return tt * tt * (hd_const(3.0) - 2.0 * tt);
    };
    const HD s1 = sw(n1sq), s2 = sw(n2sq);
    const HD in1 = hd_recip(hd_clamp_min(hd_sqrt(n1sq), 1e-12)), in2 = hd_recip(hd_clamp_min(hd_sqrt(n2sq), 1e-12));
    const HD ib2 = hd_recip(hd_clamp_min(hd_sqrt(hd_dot(b2, b2)), 1e-12));
    HD n1h[3], n2h[3], b2h[3];
    for (int c = 0; c < 3; ++c) {
      n1h[c] = n1[c] * in1;
      n2h[c] = n2[c] * in2;
      b2h[c] = b2[c] * ib2;
    }
    const HD x = hd_dot(n1h, n2h);
    hd_cross(n1h, n2h, m1);
    const HD y = hd_dot(m1, b2h);
    HD diff = hd_atan2(y, x) - hd_const(t.p);
    diff = diff - hd_const(2.0 * PI * rint(diff.f / (2.0 * PI)));  // wrap to [-pi, pi]; torch.round = half to even
    return ((0.5 * t.k) * (diff * diff)) * s1 * s2;
  }
  // kinds 3, 9
  const double theta0 = t.p * (PI / 180.0);
  HD v1[3], v2[3];
  for (int c = 0; c < 3; ++c) {
    const HD pv = P(1, c);
    v1[c] = P(0, c) - pv;
    v2[c] = P(2, c) - pv;
  }
  const HD n1 = hd_sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
  const HD n2 = hd_sqrt(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);
  const HD n12 = hd_clamp_min(n1 * n2, 1e-12);
  HD u = (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]) / n12;
  u = hd_clamp(u, -1.0, 1.0);
  const double ucp = cos(1e-3), ucn = cos(PI - 1e-3);
  auto taylor = [](HD delta) -> HD {  // (acos(1 - delta))^2, Horner as the reference
    HD term = hd_const(128.0 / 1575.0);
    term = hd_const(4.0 / 35.0) + delta * term;
    term = hd_const(8.0 / 45.0) + delta * term;
    term = hd_const(1.0 / 3.0) + delta * term;
    term = hd_const(2.0) + delta * term;
    return delta * term;
  };
  const bool near0 = u.f > ucp, nearpi = u.f < ucn;
  if (t.kind == 9) {
    auto quad = [&](double th_cut, double ucut) -> HD {  // value + slope at the cut, curvature k (dtheta/du)^2
      const double dth = -1.0 / sin(th_cut);
      const double val = 0.5 * t.k * (th_cut - theta0) * (th_cut - theta0);
      const double d1 = t.k * (th_cut - theta0) * dth, d2 = t.k * (dth * dth);
      const HD du = u - hd_const(ucut);
      return hd_const(val) + d1 * du + (0.5 * d2) * (du * du);
    };
    if (fabs(theta0) < 1e-8) {
      if (near0) return (0.5 * t.k) * taylor(hd_const(1.0) - u);
      if (nearpi) return quad(PI - 1e-3, ucn);
      const HD th = hd_acos(hd_clamp(u, -1.0, ucp));
      return (0.5 * t.k) * (th * th);
    }
    if (fabs(theta0 - PI) < 1e-8) {
      if (nearpi) return (0.5 * t.k) * taylor(hd_const(1.0) + u);
      if (near0) return quad(1e-3, ucp);
      const HD d = hd_acos(hd_clamp(u, ucn, 1.0)) - hd_const(theta0);
      return (0.5 * t.k) * (d * d);
    }
    if (near0) return quad(1e-3, ucp);
    if (nearpi) return quad(PI - 1e-3, ucn);
    const HD d = hd_acos(u) - hd_const(theta0);
    return (0.5 * t.k) * (d * d);
  }
  HD theta_minus;  // theta - theta0 (or its stand-in), energy = 1/2 k (.)^2
  if (fabs(theta0) < 1e-8) {                 // branch A
    if (near0) return (0.5 * t.k) * taylor(hd_const(1.0) - u);
    if (nearpi) {
      const HD th = hd_const(PI) - hd_sqrt(hd_clamp_min(taylor(hd_const(1.0) + u), 1e-30));
      return (0.5 * t.k) * (th * th);
    }
    const HD th = hd_acos(hd_clamp(u, ucn, ucp));
    return (0.5 * t.k) * (th * th);
  }
  if (fabs(theta0 - PI) < 1e-8) {            // branch B
    if (nearpi) return (0.5 * t.k) * taylor(hd_const(1.0) + u);
    if (near0) theta_minus = hd_sqrt(hd_clamp_min(taylor(hd_const(1.0) - u), 1e-30)) - hd_const(PI);
    else theta_minus = hd_acos(hd_clamp(u, ucn, ucp)) - hd_const(PI);
    return (0.5 * t.k) * (theta_minus * theta_minus);
  }
  if (near0) theta_minus = hd_sqrt(hd_clamp_min(taylor(hd_const(1.0) - u), 1e-30)) - hd_const(theta0);
  else if (nearpi) theta_minus = (hd_const(PI) - hd_sqrt(hd_clamp_min(taylor(hd_const(1.0) + u), 1e-30))) - hd_const(theta0);
  else theta_minus = hd_acos(u) - hd_const(theta0);
  return (0.5 * t.k) * (theta_minus * theta_minus);
}

__global__ void __launch_bounds__(128) k_bias_terms(int N, const BiasTerm* __restrict__ terms, const double* __restrict__ xyz_all,
                                                    double* __restrict__ E_all, double* __restrict__ g_all,
                                                    double* __restrict__ H_all) {
  __shared__ BiasTerm t;
  const int b = blockIdx.y, n = 3 * N;
  if (threadIdx.x == 0) t = terms[blockIdx.x];
  __syncthreads();
  const int m = (t.kind == 12 || t.kind == 13) ? 1 : t.kind >= 9 ? (int)(t.q[0] + t.q[1] + t.q[2] + t.q[3])
                            : (t.kind == 3 ? 3 : ((t.kind == 4 || t.kind == 8) ? 4 : ((t.kind == 5 || t.kind == 6) ? 2 : t.n1 + t.n2)));
  const int nc = 3 * m;
  const double* xyz = xyz_all + (size_t)b * n;
  for (int w = threadIdx.x; w < nc * nc; w += blockDim.x) {
    const int ca = w / nc, cb = w - ca * nc;
    if (cb < ca) continue;
    const HD e = bias_energy(t, xyz, ca, cb);
    const int ga = 3 * t.atoms[ca / 3] + ca % 3, gb = 3 * t.atoms[cb / 3] + cb % 3;
    if (H_all) {
      atomicAdd(&H_all[(size_t)b * n * n + (size_t)ga * n + gb], e.ab);
      if (ca != cb) atomicAdd(&H_all[(size_t)b * n * n + (size_t)gb * n + ga], e.ab);
    }
    if (ca == cb) {
      if (g_all) atomicAdd(&g_all[(size_t)b * n + ga], e.a);
      if (ca == 0 && E_all) atomicAdd(&E_all[b], e.f);
    }
  }
}

}  // namespace mop

// terms: device array of nterm records {int32 kind, n1, n2, atoms[64]; double k, p} (264 + 16 bytes, see
// mop_bias_term_bytes); E [B], grad [B][3 natoms], hess [B][3 natoms][3 natoms] are accumulated into.
extern "C" size_t mop_bias_term_bytes(void) { return sizeof(mop::BiasTerm); }

extern "C" int mop_bias_terms(int B, int natoms, int nterm, const void* terms, const double* xyz, double* E, double* grad,
                              double* hess, void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 0 && nterm >= 0 && xyz && (nterm == 0 || terms), "mop_bias_terms: bad arguments");
  if (B == 0 || nterm == 0) return MOP_OK;
  dim3 grid(nterm, B);
  mop::k_bias_terms<<<grid, 128, 0, (cudaStream_t)stream>>>(natoms, (const mop::BiasTerm*)terms, xyz, E, grad, hess);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
