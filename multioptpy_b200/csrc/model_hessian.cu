// Model-Hessian assembly: connectivity tables, Fischer and Lindh
// model Hessians as sums of  k_t b_t b_t^T  over internal coordinates.
//
// One CTA per structure.  Phase 1: connectivity (connectivity.cuh, bit-exact tables);
// phase 2: one thread per internal coordinate evaluates its force constant and Wilson
// b-vectors (restating ModelHessian/calc_params.py stretch2/bend2/torsion2) into an L2
// scratch record; phase 3: one thread per ATOM-PAIR block gathers the records that touch
// both atoms IN TABLE ORDER (bonds, angles, dihedrals) — deterministic, no atomics, the
// reference's accumulation order — and writes the 3x3 block and its mirror.
// The TR/ROT projection that ends every reference model (calc_tools.py:249) is the
// separate k_project_trrot launch.  HBM-bound: 8 n^2 bytes written per structure.
#include "connectivity.cuh"

namespace mop {

constexpr int MH_THREADS = 256;
constexpr double PI_D = 3.141592653589793;

struct ICRec {      // one internal coordinate
  int atom[4];      // -1 padded
  double k;         // force constant (0: skipped)
  double b[12];     // Wilson vectors, b[3*p + c] for atom slot p
};

// stretch2 (calc_params.py:220-227): b0 = -(x0 - x1)/r, b1 = +(x0 - x1)/r
__device__ __forceinline__ double stretch2(const double* x0, const double* x1, double b0[3], double b1[3]) {
  const double d0 = x0[0] - x1[0], d1 = x0[1] - x1[1], d2 = x0[2] - x1[2];
  const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2)));
  b0[0] = -1.0 * d0 / r; b0[1] = -1.0 * d1 / r; b0[2] = -1.0 * d2 / r;
  b1[0] = d0 / r; b1[1] = d1 / r; b1[2] = d2 / r;
  return r;
}

// bend2 (calc_params.py:183-218) for atoms (0, 1, 2), centre 1.  Returns the angle.
__device__ __forceinline__ double bend2(const double* x0, const double* x1, const double* x2,
                                        double bf[9], double* r01, double* r12) {
  double bij0[3], bij1[3], bjk0[3], bjk1[3];
  const double rij = stretch2(x0, x1, bij0, bij1);
  const double rjk = stretch2(x1, x2, bjk0, bjk1);
  double co = 0.0, crap = 0.0;
  for (int i = 0; i < 3; ++i) {
    co += bij0[i] * bjk1[i];
    crap += bij0[i] * bij0[i];
    crap += bjk1[i] * bjk1[i];
  }
  double fir, si;
  if (sqrt(crap) < 1e-12) {
    fir = PI_D - asin(sqrt(crap));
    si = sqrt(crap);
  } else {
    fir = acos(co);
    si = sqrt(1.0 - co * co);
  }
  if (fabs(fir - PI_D) < 1e-12) fir = PI_D;
  const double den1 = rij * si, den2 = rjk * si;
  for (int i = 0; i < 3; ++i) {
    bf[i] = den1 < 1e-12 ? 0.0 : (co * bij0[i] - bjk1[i]) / den1;
    bf[6 + i] = den2 < 1e-12 ? 0.0 : (co * bjk1[i] - bij0[i]) / den2;
    bf[3 + i] = -1.0 * (bf[i] + bf[6 + i]);
  }
  if (r01) *r01 = rij;
  if (r12) *r12 = rjk;
  return fir;
}

// torsion2 (calc_params.py:137-181): b-vectors of the dihedral 0-1-2-3.
__device__ __forceinline__ void torsion2(const double* x0, const double* x1, const double* x2,
                                         const double* x3, double bt[12]) {
  double brij0[3], brij1[3], brjk0[3], brjk1[3], brkl0[3], brkl1[3], tmp[9];
  const double r1 = stretch2(x0, x1, brij0, brij1);
  const double r2 = stretch2(x1, x2, brjk0, brjk1);
  const double r3 = stretch2(x2, x3, brkl0, brkl1);
  const double fi2 = bend2(x0, x1, x2, tmp, nullptr, nullptr);
  const double fi3 = bend2(x1, x2, x3, tmp, nullptr, nullptr);
  const double s2 = sin(fi2), s3 = sin(fi3), c2 = cos(fi2), c3 = cos(fi3);
  for (int ix = 1; ix <= 3; ++ix) {
    int iy = ix + 1;
    if (iy > 3) iy -= 3;
    int iz = iy + 1;
    if (iz > 3) iz -= 3;
    const double t0 = (brij1[iy - 1] * brjk1[iz - 1] - brij1[iz - 1] * brjk1[iy - 1]) / (r1 * (s2 * s2));
    const double t3 = (brkl0[iy - 1] * brjk0[iz - 1] - brkl0[iz - 1] * brjk0[iy - 1]) / (r3 * (s3 * s3));
    const double t1 = -1.0 * ((r2 - r1 * c2) * t0 + r3 * c3 * t3) / r2;
    bt[ix - 1] = t0;
    bt[9 + ix - 1] = t3;
    bt[3 + ix - 1] = t1;
    bt[6 + ix - 1] = -1.0 * (t0 + t1 + t3);
  }
}

// sin^2 of the angle a-b-c (fischer.py:155-165)
__device__ __forceinline__ double sin_sq_angle(const double* xa, const double* xb, const double* xc) {
  const double v1[3] = {xa[0] - xb[0], xa[1] - xb[1], xa[2] - xb[2]};
  const double v2[3] = {xc[0] - xb[0], xc[1] - xb[1], xc[2] - xb[2]};
  const double cx = v1[1] * v2[2] - v1[2] * v2[1], cy = v1[2] * v2[0] - v1[0] * v2[2],
               cz = v1[0] * v2[1] - v1[1] * v2[0];
  const double cs = cx * cx + cy * cy + cz * cz;
  const double n1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  const double n2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  if (n1 * n2 < 1e-12) return 0.0;
  return cs / (n1 * n2);
}

// ---------------------------------------------------------------------------------------
// D3(BJ) pair block of FischerD3ApproxHessianOld.d3_hessian_contribution (fischerd3old.py:85-128):
// h_proj u u^T + h_perp (1 - u u^T) with the reference's "simplified" second derivative.  pi / pj: {cov radius, C6, r4r2,
// vdW radius} of the two atoms; d3c = (s6, s8, a1, a2).
__device__ __forceinline__ void d3_pair_block(const double* xi, const double* xj, const double* pi, const double* pj,
                                              const double d3c[4], double blk[9]) {
  const double dx = xi[0] - xj[0], dy = xi[1] - xj[1], dz = xi[2] - xj[2];
  const double r = np_dist(xi, xj);
  const double c6 = sqrt(pi[1] * pj[1]);
  const double c8 = 3.0 * c6 * sqrt(pi[2] * pj[2]);
  const double r0 = pi[3] + pj[3];
  const double s6 = d3c[0], s8 = d3c[1], a1 = d3c[2], a2 = d3c[3];
  const double r2 = r * r, r4 = r2 * r2, r5 = r4 * r, r6 = r4 * r2, r7 = r6 * r, r8 = r4 * r4, r9 = r8 * r;
  const double q6 = a1 * r0 + a2, q8 = a1 * r0 + (a2 + 2.0);
  const double q62 = q6 * q6, q82 = q8 * q8, q84 = q82 * q82;
  const double d6 = r6 + q62 * q62 * q62, d8 = r8 + q84 * q84;
  const double f6 = r6 / d6, f8 = r8 / d8;
  const double df6 = 6.0 * r5 / d6 - 6.0 * (r6 * r6) / (d6 * d6);
  const double df8 = 8.0 * r7 / d8 - 8.0 * (r8 * r8) / (d8 * d8);
  const double g6 = -s6 * c6 * ((-6.0 / r7) * f6 + (1.0 / r6) * df6);
  const double g8 = -s8 * c8 * ((-8.0 / r9) * f8 + (1.0 / r8) * df8);
  const double hpar = s6 * c6 / r8 * (42.0 * f6 - r * df6) + s8 * c8 / (r8 * r2) * (72.0 * f8 - r * df8);
  const double hperp = (g6 + g8) / r;
  const double u[3] = {dx / r, dy / r, dz / r};
  for (int p = 0; p < 3; ++p)
    for (int m = 0; m < 3; ++m) {
      const double P = u[p] * u[m];
      blk[3 * p + m] = hpar * P + hperp * ((p == m ? 1.0 : 0.0) - P);
    }
}

// kind 0: connectivity tables only; kind 1: Fischer (ModelHessian/fischer.py); kind 2: Fischer + D3, old variant
// (ModelHessian/fischerd3old.py: linear-angle skips, sin^2-damped torsions, D3(BJ) blocks for the non-bonded pairs;
// rad_all then holds FOUR doubles per atom: covalent radius, D2 C6, D3 r4r2, D2 vdW radius); kind 3: Fischer + "dynamic"
// D3 (ModelHessian/fischerd3.py: C6 scaled by the fractional coordination number against a reference valence, the
// 1.1-factor connectivity for the torsion bond count and the non-bonded mask; FIVE doubles per atom, + reference CN)
__global__ void __launch_bounds__(MH_THREADS, 4)
k_model_hessian(int kind, int N, const double* __restrict__ xyz_all, const double* __restrict__ rad_all,
                int rad_stride, double factor, int capB, int capA, int capD, int* __restrict__ bonds_all,
                int* __restrict__ angles_all, int* __restrict__ dihs_all, int* __restrict__ counts_all,
                ICRec* __restrict__ rec_all, double* __restrict__ H_all, int32_t* __restrict__ status, double s6,
                double s8, double a1, double a2) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  double* xyz = sm;              // 3N
  double* rad = xyz + 3 * N;     // N
  int* cnt = (int*)(rad + N);    // 4
  int* wtot = cnt + 4;           // 36
  int* nb13 = wtot + 36;         // N : neighbour count with the 1.3 factor (Fischer bond_sum)
  unsigned char* bm = (unsigned char*)(nb13 + N + (N & 1));  // N*N
  for (int i = tid; i < 3 * N; i += MH_THREADS) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  const int pw = kind == 2 ? 4 : (kind == 3 ? 5 : 1);  // doubles per atom in rad_all
  const double* prm4 = rad_all + (size_t)b * rad_stride * pw;
  for (int i = tid; i < N; i += MH_THREADS) rad[i] = prm4[(size_t)i * pw];
  __syncthreads();
  bond_matrix(N, xyz, rad, factor, bm);
  ConnTables T;
  T.bonds = bonds_all + (size_t)b * capB * 2;
  T.angles = angles_all + (size_t)b * capA * 3;
  T.dihs = dihs_all + (size_t)b * capD * 4;
  T.capB = capB; T.capA = capA; T.capD = capD;
  enumerate_tables(N, bm, T, cnt, wtot);
  if (tid == 0) {
    counts_all[3 * b] = T.nb;
    counts_all[3 * b + 1] = T.na;
    counts_all[3 * b + 2] = T.nd;
    if (status) status[b] = T.overflow ? 1 : 0;
  }
  if (kind == 0) return;

  // ---- Fischer: records -------------------------------------------------------------
  ICRec* rec = rec_all + (size_t)b * (capB + capA + capD);
  const int nrec = T.nb + T.na + T.nd;
  // bond_sum uses a SECOND connectivity with factor 1.3 (fischer.py:17,63)
  double* cscale = (double*)(bm + (((size_t)N * N + 7) & ~(size_t)7));  // kind 3: [N] CN scaling of C6
  for (int i = tid; i < N; i += MH_THREADS) {
    int c = 0;
    double cn = 0.0;
    for (int j = 0; j < N; ++j) {
      if (kind == 3) {
        // calc_coordination_numbers (fischerd3.py:47-62): 1 / (1 + exp(clip(-16 (4/3 r / rcov - 1), -100, 100))),
        // the diagonal (r = inf) included, which adds 1 / (1 + e^-100) to every atom
        double term = -100.0;
        if (j != i) {
          const double d = np_dist(xyz + 3 * i, xyz + 3 * j);
          term = -16.0 * ((4.0 / 3.0) * (d / __dadd_rn(rad[i], rad[j])) - 1.0);
          term = fmin(fmax(term, -100.0), 100.0);
        }
        cn += 1.0 / (1.0 + exp(term));
        if (j != i) c += bm[i * N + j];   // neighbor_counts = bond_mat.sum(axis=1) (fischerd3.py:138)
        continue;
      }
      if (j == i) continue;
      const int lo = i < j ? i : j, hi = i < j ? j : i;  // dist = ||coord[lo] - coord[hi]||
      const double d = np_dist(xyz + 3 * lo, xyz + 3 * hi);
      const double cs = __dadd_rn(rad[lo], rad[hi]);
      c += d <= __dmul_rn(cs, 1.3);
    }
    nb13[i] = c;
    if (kind == 3) cscale[i] = fmin(fmax(1.0 - 0.05 * (cn - prm4[(size_t)i * 5 + 4]), 0.75), 1.25);
  }
  __syncthreads();
  for (int t = tid; t < nrec; t += MH_THREADS) {
    ICRec r;
    for (int q = 0; q < 4; ++q) r.atom[q] = -1;
    for (int q = 0; q < 12; ++q) r.b[q] = 0.0;
    r.k = 0.0;
    if (t < T.nb) {  // fischer_bond (fischer.py:73-98)
      const int i = T.bonds[2 * t], j = T.bonds[2 * t + 1];
      r.atom[0] = i; r.atom[1] = j;
      const double rij = np_dist(xyz + 3 * i, xyz + 3 * j);
      const double rcov = __dadd_rn(rad[i], rad[j]);
      r.k = (kind == 3 && rij < 0.1) ? 0.0 : 0.3601 * exp(-1.944 * (rij - rcov));
      stretch2(xyz + 3 * i, xyz + 3 * j, r.b, r.b + 3);
    } else if (t < T.nb + T.na) {  // fischer_angle (:100-131)
      const int* a = T.angles + 3 * (t - T.nb);
      const int i = a[0], j = a[1], k = a[2];
      r.atom[0] = i; r.atom[1] = j; r.atom[2] = k;
      const double rij = np_dist(xyz + 3 * i, xyz + 3 * j), rjk = np_dist(xyz + 3 * j, xyz + 3 * k);
      const double cij = __dadd_rn(rad[i], rad[j]), cjk = __dadd_rn(rad[j], rad[k]);
      const double val = cij * cjk;
      r.k = fabs(val) < 1e-10 ? 0.0
                              : 0.089 + 0.11 / pow(val, -0.42) * exp(-0.44 * (rij + rjk - cij - cjk));
      bool skip = false;
      if (kind >= 2) {  // fischerd3old.py:195-210, fischerd3.py:104-117: overlapping atoms and (anti)parallel arms are skipped
        const double* xi = xyz + 3 * i; const double* xj = xyz + 3 * j; const double* xk = xyz + 3 * k;
        const double dt = (xi[0] - xj[0]) * (xk[0] - xj[0]) + (xi[1] - xj[1]) * (xk[1] - xj[1]) +
                          (xi[2] - xj[2]) * (xk[2] - xj[2]);
        skip = rij < 0.1 || rjk < 0.1 || fabs(dt / (rij * rjk)) > 0.9999;
      }
      if (skip) r.k = 0.0;
      else bend2(xyz + 3 * i, xyz + 3 * j, xyz + 3 * k, r.b, nullptr, nullptr);
    } else {  // fischer_dihedral (:133-210)
      const int* a = T.dihs + 4 * (t - T.nb - T.na);
      const int i = a[0], j = a[1], k = a[2], l = a[3];
      r.atom[0] = i; r.atom[1] = j; r.atom[2] = k; r.atom[3] = l;
      double s1, s2, damp = 1.0;
      bool ok;
      if (kind >= 2) {  // fischerd3old.py:258-300 (cut-offs 1e-8 / 1e-4), fischerd3.py:143-163 (0.1 / 1e-3)
        const double* xi = xyz + 3 * i; const double* xj = xyz + 3 * j; const double* xk = xyz + 3 * k; const double* xl = xyz + 3 * l;
        const double nji = np_dist(xi, xj), njk = np_dist(xk, xj), nkl = np_dist(xl, xk);
        const double nmin = kind == 3 ? 0.1 : 1e-8, smin = kind == 3 ? 1e-3 : 1e-4;
        ok = !(nji < nmin || njk < nmin || nkl < nmin);
        if (ok) {
          const double d1 = (xi[0] - xj[0]) * (xk[0] - xj[0]) + (xi[1] - xj[1]) * (xk[1] - xj[1]) + (xi[2] - xj[2]) * (xk[2] - xj[2]);
          const double d2 = -((xk[0] - xj[0]) * (xl[0] - xk[0]) + (xk[1] - xj[1]) * (xl[1] - xk[1]) + (xk[2] - xj[2]) * (xl[2] - xk[2]));
          const double c1 = d1 / (nji * njk), c2 = d2 / (njk * nkl);
          s1 = 1.0 - fmin(c1 * c1, 1.0);
          s2 = 1.0 - fmin(c2 * c2, 1.0);
          ok = !(s1 < smin || s2 < smin);
          damp = s1 * s2;
        }
      } else {
        s1 = sin_sq_angle(xyz + 3 * i, xyz + 3 * j, xyz + 3 * k);
        s2 = sin_sq_angle(xyz + 3 * j, xyz + 3 * k, xyz + 3 * l);
        ok = !(s1 < 1.0e-3 || s2 < 1.0e-3);
      }
      if (ok) {
        const double rjk = np_dist(xyz + 3 * j, xyz + 3 * k);
        const double cjk = __dadd_rn(rad[j], rad[k]);
        const int bond_sum = nb13[j] + nb13[k] - 2;
        const double val = rjk * cjk;
        r.k = fabs(val) < 1e-10
                  ? 0.0
                  : 0.0015 + 14.0 * pow((double)max(bond_sum, 0), 0.57) / pow(val, 4.0) * exp(-2.85 * (rjk - cjk));
        r.k *= damp;
        torsion2(xyz + 3 * i, xyz + 3 * j, xyz + 3 * k, xyz + 3 * l, r.b);
      }
    }
    rec[t] = r;
  }
  __syncthreads();

  // ---- gather per atom-pair block (a <= c), table order --------------------------------
  double* H = H_all + (size_t)b * 9 * N * N;
  const int n = 3 * N;
  for (int e = tid; e < N * N; e += MH_THREADS) {
    const int a = e / N, c = e - a * N;
    if (a > c) continue;
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < nrec; ++t) {
      const ICRec* r = rec + t;
      int pa = -1, pc = -1;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int at = r->atom[q];
        if (at == a) pa = q;
        if (at == c) pc = q;
      }
      if (pa < 0 || pc < 0) continue;
      const double k = r->k;
      if (k == 0.0) continue;  // skipped dihedral (reference `continue`s before accumulating)
      if (kind >= 2) {  // force_const * np.outer(b_m, b_n) (fischerd3old.py:168-229)
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int m = 0; m < 3; ++m) acc[3 * p + m] += k * (r->b[3 * pa + p] * r->b[3 * pc + m]);
      } else {
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int m = 0; m < 3; ++m) acc[3 * p + m] += k * r->b[3 * pa + p] * r->b[3 * pc + m];
      }
    }
    if (kind >= 2) {  // d3_dispersion_hessian (fischerd3old.py:322-352): non-bonded (factor 1.3) pairs, r >= 0.1;
                      // fischerd3.py:205-214: not bonded in the table connectivity, r > 0.1, C6 scaled by the CN factors
      const double d3c[4] = {s6, s8, a1, a2};
      for (int o = (a == c ? 0 : c); o < (a == c ? N : c + 1); ++o) {
        if (o == a) continue;
        const int hi = a > o ? a : o, lo = a > o ? o : a;  // the reference's pair (i > j)
        const double d = np_dist(xyz + 3 * hi, xyz + 3 * lo);
        double ph[4], pl[4];
        if (kind == 2) {
          if (d <= __dmul_rn(__dadd_rn(rad[hi], rad[lo]), 1.3) || d < 0.1) continue;
          for (int q = 0; q < 4; ++q) { ph[q] = prm4[4 * (size_t)hi + q]; pl[q] = prm4[4 * (size_t)lo + q]; }
        } else {
          if (bm[hi * N + lo] || !(d > 0.1)) continue;
          for (int q = 0; q < 4; ++q) { ph[q] = prm4[5 * (size_t)hi + q]; pl[q] = prm4[5 * (size_t)lo + q]; }
          ph[1] *= cscale[hi];
          pl[1] *= cscale[lo];
        }
        double blk[9];
        d3_pair_block(xyz + 3 * hi, xyz + 3 * lo, ph, pl, d3c, blk);
        for (int q = 0; q < 9; ++q) acc[q] += (a == c) ? blk[q] : -blk[q];
      }
    }
    // upper triangle is authoritative: cart_hess[i, j] = cart_hess[j, i] for j < i (fischer.py:229-231)
    for (int p = 0; p < 3; ++p)
      for (int m = 0; m < 3; ++m) {
        const int row = 3 * a + p, col = 3 * c + m;
        if (row <= col) {
          H[(size_t)row * n + col] = acc[3 * p + m];
          H[(size_t)col * n + row] = acc[3 * p + m];
        }
      }
  }
}


// ---------------------------------------------------------------------------------------
// Lindh model Hessian (ModelHessian/lindh.py:79-165): diagonal force constants k_p over ALL
// atom pairs p = (i < j) (itertools.combinations order), then  H = B^T diag(k) B  with the
// all-pairs distance B matrix (Coordinate/redundant_coordinate.py:15-43), i.e. block (i,i)
// += k e e^T, block (i,j) = -k e e^T with e = (x_i - x_j)/r.  The reference adds a K term
// built from an internal-coordinate gradient obtained by solving the SINGULAR system
// (B B^T) q = B g and indexed inconsistently; it is numerically ill-posed in the
// reference itself, so it is not part of this kernel (kdiag_out exposes k for the
// decomposed parity check).
// atom parameters prm[a][6] = {cov radius, period index 0/1/2, mass, UFF distance, UFF well
// depth, UFF effective charge}.
__global__ void __launch_bounds__(MH_THREADS, 4)
k_lindh(int N, const double* __restrict__ xyz_all, const double* __restrict__ prm_all, int prm_stride,
        int capB, int capA, int capD, int* __restrict__ bonds_all, int* __restrict__ angles_all,
        int* __restrict__ dihs_all, int* __restrict__ counts_all, double* __restrict__ fc_all,
        double* __restrict__ kdiag_all, double* __restrict__ H_all, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int M = N * (N - 1) / 2;
  double* xyz = sm;               // 3N
  double* prm = xyz + 3 * N;      // 6N
  double* rad = prm + 6 * N;      // N
  double* kd = rad + N;           // M
  int* cnt = (int*)(kd + M);      // 4
  int* wtot = cnt + 4;            // 36
  unsigned char* bm = (unsigned char*)(wtot + 36);  // N*N
  for (int i = tid; i < 3 * N; i += MH_THREADS) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  for (int i = tid; i < 6 * N; i += MH_THREADS) prm[i] = prm_all[(size_t)b * prm_stride * 6 + i];
  __syncthreads();
  for (int i = tid; i < N; i += MH_THREADS) rad[i] = prm[6 * i];
  for (int i = tid; i < M; i += MH_THREADS) kd[i] = 0.0;
  __syncthreads();
  bond_matrix(N, xyz, rad, 1.1, bm);
  ConnTables T;
  T.bonds = bonds_all + (size_t)b * capB * 2;
  T.angles = angles_all + (size_t)b * capA * 3;
  T.dihs = dihs_all + (size_t)b * capD * 4;
  T.capB = capB; T.capA = capA; T.capD = capD;
  enumerate_tables(N, bm, T, cnt, wtot);
  if (tid == 0) {
    counts_all[3 * b] = T.nb; counts_all[3 * b + 1] = T.na; counts_all[3 * b + 2] = T.nd;
    if (status) status[b] = T.overflow ? 1 : 0;
  }
  const double alpha_tab[3][3] = {{1.0000, 0.3949, 0.3949}, {0.3949, 0.2800, 0.2800}, {0.3949, 0.2800, 0.2800}};
  const int nrec = T.nb + T.na + T.nd;
  double* fc = fc_all + (size_t)b * (capB + capA + capD);
  // force constant of every internal coordinate (lindh.py:89-98)
  for (int t = tid; t < nrec; t += MH_THREADS) {
    int at[4], len;
    double f;
    if (t < T.nb) { len = 2; f = 0.45; at[0] = T.bonds[2 * t]; at[1] = T.bonds[2 * t + 1]; }
    else if (t < T.nb + T.na) { len = 3; f = 0.15; const int* a = T.angles + 3 * (t - T.nb); at[0] = a[0]; at[1] = a[1]; at[2] = a[2]; }
    else { len = 4; f = 0.005; const int* a = T.dihs + 4 * (t - T.nb - T.na); at[0] = a[0]; at[1] = a[1]; at[2] = a[2]; at[3] = a[3]; }
    for (int q = 0; q + 1 < len; ++q) {
      const int i = at[q], j = at[q + 1];
      const double cR = __dadd_rn(rad[i], rad[j]);
      const double al = alpha_tab[(int)prm[6 * i + 1]][(int)prm[6 * j + 1]];
      const double R = np_dist(xyz + 3 * i, xyz + 3 * j);
      f *= exp(al * (cR * cR - R * R));
    }
    fc[t] = f;
  }
  __syncthreads();
  // accumulate into the pair diagonal: one thread per internal coordinate, FP64 shared-memory atomics
  // (a single thread walking the tables in global memory cost ~1e6 cycles per structure; the summation
  // order differs from the reference's by rounding only)
  {
    auto pidx = [N](int i, int j) { if (i > j) { const int t = i; i = j; j = t; } return i * N - i * (i + 1) / 2 + (j - i - 1); };
    for (int t = tid; t < nrec; t += MH_THREADS) {
      const double f = fc[t];
      if (t < T.nb) {
        const int i = T.bonds[2 * t], j = T.bonds[2 * t + 1];
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const double m1 = prm[6 * lo + 2], m2 = prm[6 * hi + 2];
        atomicAdd(&kd[pidx(i, j)], f / ((m1 * m2) / (m1 + m2)));
      } else if (t < T.nb + T.na) {
        const int* a = T.angles + 3 * (t - T.nb);
        atomicAdd(&kd[pidx(a[0], a[1])], f);
        atomicAdd(&kd[pidx(a[1], a[2])], f);
      } else {
        const int* a = T.dihs + 4 * (t - T.nb - T.na);
        atomicAdd(&kd[pidx(a[0], a[1])], f);
        atomicAdd(&kd[pidx(a[1], a[2])], f);
        atomicAdd(&kd[pidx(a[2], a[3])], f);
      }
    }
  }
  __syncthreads();
  // non-bonded pairs: Lennard-Jones + electrostatic force constants (lindh.py:19-41,133-137)
  for (int e = tid; e < N * N; e += MH_THREADS) {
    const int i = e / N, j = e - i * N;
    if (i >= j || bm[e] == 1) continue;
    const double d = np_dist(xyz + 3 * i, xyz + 3 * j);
    const double eps = sqrt(prm[6 * i + 4] * prm[6 * j + 4]);
    const double sig = sqrt(prm[6 * i + 3] * prm[6 * j + 3]);
    const double lj = -12.0 * eps * (-7.0 * (pow(sig, 6.0) / pow(d, 8.0)) + 13.0 * (pow(sig, 12.0) / pow(d, 14.0)));
    const double q = prm[6 * i + 5] * prm[6 * j + 5];
    const double es = 664.12 * (q / pow(d, 3.0)) * (0.52917721067 * 0.52917721067 / 627.509);
    const int p = i * N - i * (i + 1) / 2 + (j - i - 1);
    kd[p] = (kd[p] + lj) + es;
  }
  __syncthreads();
  if (kdiag_all)
    for (int p = tid; p < M; p += MH_THREADS) kdiag_all[(size_t)b * M + p] = kd[p];
  // H = B^T diag(k) B, one thread per atom-pair block (a <= c)
  double* H = H_all + (size_t)b * 9 * N * N;
  const int n = 3 * N;
  for (int e = tid; e < N * N; e += MH_THREADS) {
    const int a = e / N, c = e - a * N;
    if (a > c) continue;
    double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (a != c) {
      const double dx = xyz[3 * a] - xyz[3 * c], dy = xyz[3 * a + 1] - xyz[3 * c + 1], dz = xyz[3 * a + 2] - xyz[3 * c + 2];
      const double r = np_dist(xyz + 3 * a, xyz + 3 * c);
      const double ev[3] = {dx / r, dy / r, dz / r};
      const double k = kd[a * N - a * (a + 1) / 2 + (c - a - 1)];
      for (int p = 0; p < 3; ++p)
        for (int m = 0; m < 3; ++m) blk[3 * p + m] = -k * ev[p] * ev[m];
    } else {
      for (int o = 0; o < N; ++o) {
        if (o == a) continue;
        const int lo = a < o ? a : o, hi = a < o ? o : a;
        const double dx = xyz[3 * lo] - xyz[3 * hi], dy = xyz[3 * lo + 1] - xyz[3 * hi + 1], dz = xyz[3 * lo + 2] - xyz[3 * hi + 2];
        const double r = np_dist(xyz + 3 * lo, xyz + 3 * hi);
        const double ev[3] = {dx / r, dy / r, dz / r};
        const double k = kd[lo * N - lo * (lo + 1) / 2 + (hi - lo - 1)];
        for (int p = 0; p < 3; ++p)
          for (int m = 0; m < 3; ++m) blk[3 * p + m] += k * ev[p] * ev[m];
      }
    }
    for (int p = 0; p < 3; ++p)
      for (int m = 0; m < 3; ++m) {
        H[(size_t)(3 * a + p) * n + 3 * c + m] = blk[3 * p + m];
        H[(size_t)(3 * c + m) * n + 3 * a + p] = blk[3 * p + m];
      }
  }
}

}  // namespace mop

// forward declaration (project.cu)
int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                             const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                             cudaStream_t stream);

static size_t mh_smem(int N) {
  return sizeof(double) * (5 * (size_t)N + 2) + sizeof(int) * (40 + (size_t)N + 1) + (size_t)N * N + 16;
}

extern "C" int mop_connectivity(int B, int natoms, const double* xyz, const double* radii,
                                int radii_stride, double factor, int capB, int capA, int capD,
                                int32_t* bonds, int32_t* angles, int32_t* dihedrals, int32_t* counts,
                                int32_t* status, void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 0, "mop_connectivity: B >= 0 and natoms > 0 required");
  MOP_REQUIRE(xyz && radii && bonds && angles && dihedrals && counts,
              "mop_connectivity: xyz, radii, bonds, angles, dihedrals, counts must be device pointers");
  MOP_REQUIRE(radii_stride == 0 || radii_stride == natoms, "mop_connectivity: radii_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;
  const size_t smem = mh_smem(natoms);
  if (smem > 200 * 1024) {
    mop_set_error("mop_connectivity: natoms = %d too large", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_model_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_model_hessian<<<B, mop::MH_THREADS, smem, (cudaStream_t)stream>>>(
      0, natoms, xyz, radii, radii_stride, factor, capB, capA, capD, bonds, angles, dihedrals, counts,
      nullptr, nullptr, status, 0.0, 0.0, 0.0, 0.0);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

static void fischer_caps(int N, int* capB, int* capA, int* capD) {
  *capB = N * 8 < N * (N - 1) / 2 + 1 ? N * 8 : N * (N - 1) / 2 + 1;
  *capA = N * 28;
  *capD = N * 64;
}

extern "C" size_t mop_fischer_workspace_bytes(int B, int natoms) {
  if (B <= 0 || natoms <= 0) return 0;
  int cb, ca, cd;
  fischer_caps(natoms, &cb, &ca, &cd);
  size_t bytes = (size_t)B * (2 * cb + 3 * ca + 4 * cd + 4) * sizeof(int32_t);
  bytes = (bytes + 255) & ~(size_t)255;
  bytes += (size_t)B * (cb + ca + cd) * sizeof(mop::ICRec);
  bytes = (bytes + 255) & ~(size_t)255;
  bytes += (size_t)B * 9 * natoms * natoms * sizeof(double);  // unprojected Hessian
  return bytes;
}

// FischerApproxHessian.main (ModelHessian/fischer.py:212-236): H_out [B][3N][3N], TR/ROT projected.
extern "C" int mop_fischer_hessian(int B, int natoms, const double* xyz, const double* radii,
                                   int radii_stride, double* H_out, int32_t* counts_out,
                                   int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 0, "mop_fischer_hessian: B >= 0 and natoms > 0 required");
  MOP_REQUIRE(xyz && radii && H_out && work, "mop_fischer_hessian: xyz, radii, H_out, work must be device pointers");
  MOP_REQUIRE(radii_stride == 0 || radii_stride == natoms, "mop_fischer_hessian: radii_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_fischer_workspace_bytes(B, natoms)) {
    mop_set_error("mop_fischer_hessian: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  int cb, ca, cd;
  fischer_caps(natoms, &cb, &ca, &cd);
  char* w = (char*)work;
  int32_t* bonds = (int32_t*)w;
  int32_t* angles = bonds + (size_t)B * 2 * cb;
  int32_t* dihs = angles + (size_t)B * 3 * ca;
  int32_t* counts = dihs + (size_t)B * 4 * cd;
  size_t off = ((size_t)B * (2 * cb + 3 * ca + 4 * cd + 4) * sizeof(int32_t) + 255) & ~(size_t)255;
  mop::ICRec* rec = (mop::ICRec*)(w + off);
  off += ((size_t)B * (cb + ca + cd) * sizeof(mop::ICRec) + 255) & ~(size_t)255;
  double* Hraw = (double*)(w + off);
  const size_t smem = mh_smem(natoms);
  if (smem > 200 * 1024) {
    mop_set_error("mop_fischer_hessian: natoms = %d too large", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_model_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_model_hessian<<<B, mop::MH_THREADS, smem, stream>>>(1, natoms, xyz, radii, radii_stride, 1.1, cb, ca,
                                                           cd, bonds, angles, dihs, counts, rec, Hraw,
                                                           status, 0.0, 0.0, 0.0, 0.0);
  MOP_CHECK_CUDA(cudaGetLastError());
  if (counts_out)
    MOP_CHECK_CUDA(cudaMemcpyAsync(counts_out, counts, sizeof(int32_t) * 3 * (size_t)B,
                                   cudaMemcpyDeviceToDevice, stream));
  return mop_launch_project_trrot(B, 3 * natoms, Hraw, nullptr, xyz, nullptr, H_out, nullptr, nullptr, 0, stream);
}


// FischerD3ApproxHessianOld.main (ModelHessian/fischerd3old.py:355-381): H_out [B][3N][3N], TR/ROT projected.
// atom_params [B or 1][natoms][4] = {covalent radius (Bohr), D2 C6 (hartree bohr^6), D3 r4r2, D2 vdW radius (Bohr)};
// d3 = (s6, s8, a1, a2).  Workspace: mop_fischer_workspace_bytes.
static int fischer_d3_common(int kind, int B, int natoms, const double* xyz, const double* atom_params,
                             int param_stride, double s6, double s8, double a1, double a2, double* H_out,
                             int32_t* counts_out, int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 0, "mop_fischer_d3old_hessian: B >= 0 and natoms > 0 required");
  MOP_REQUIRE(xyz && atom_params && H_out && work, "mop_fischer_d3old_hessian: xyz, atom_params, H_out, work required");
  MOP_REQUIRE(param_stride == 0 || param_stride == natoms, "mop_fischer_d3old_hessian: param_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_fischer_workspace_bytes(B, natoms)) {
    mop_set_error("mop_fischer_d3old_hessian: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  int cb, ca, cd;
  fischer_caps(natoms, &cb, &ca, &cd);
  char* w = (char*)work;
  int32_t* bonds = (int32_t*)w;
  int32_t* angles = bonds + (size_t)B * 2 * cb;
  int32_t* dihs = angles + (size_t)B * 3 * ca;
  int32_t* counts = dihs + (size_t)B * 4 * cd;
  size_t off = ((size_t)B * (2 * cb + 3 * ca + 4 * cd + 4) * sizeof(int32_t) + 255) & ~(size_t)255;
  mop::ICRec* rec = (mop::ICRec*)(w + off);
  off += ((size_t)B * (cb + ca + cd) * sizeof(mop::ICRec) + 255) & ~(size_t)255;
  double* Hraw = (double*)(w + off);
  const size_t smem = mh_smem(natoms);
  if (smem > 200 * 1024) {
    mop_set_error("mop_fischer_d3old_hessian: natoms = %d too large", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_model_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_model_hessian<<<B, mop::MH_THREADS, smem, stream>>>(kind, natoms, xyz, atom_params, param_stride, 1.1, cb, ca, cd,
                                                           bonds, angles, dihs, counts, rec, Hraw, status, s6, s8, a1, a2);
  MOP_CHECK_CUDA(cudaGetLastError());
  if (counts_out)
    MOP_CHECK_CUDA(cudaMemcpyAsync(counts_out, counts, sizeof(int32_t) * 3 * (size_t)B, cudaMemcpyDeviceToDevice, stream));
  return mop_launch_project_trrot(B, 3 * natoms, Hraw, nullptr, xyz, nullptr, H_out, nullptr, nullptr, 0, stream);
}

extern "C" int mop_fischer_d3old_hessian(int B, int natoms, const double* xyz, const double* atom_params,
                                         int param_stride, double s6, double s8, double a1, double a2, double* H_out,
                                         int32_t* counts_out, int32_t* status, void* work, size_t work_bytes,
                                         void* stream_) {
  return fischer_d3_common(2, B, natoms, xyz, atom_params, param_stride, s6, s8, a1, a2, H_out, counts_out, status, work,
                           work_bytes, stream_);
}

// FischerD3ApproxHessian.main (ModelHessian/fischerd3.py:186-304, the variant the AutoTS configurations select):
// atom_params [B or 1][natoms][5] = {covalent radius, D2 C6, D3 r4r2, D2 vdW radius, reference coordination number}.
extern "C" int mop_fischer_d3_hessian(int B, int natoms, const double* xyz, const double* atom_params,
                                      int param_stride, double s6, double s8, double a1, double a2, double* H_out,
                                      int32_t* counts_out, int32_t* status, void* work, size_t work_bytes,
                                      void* stream_) {
  return fischer_d3_common(3, B, natoms, xyz, atom_params, param_stride, s6, s8, a1, a2, H_out, counts_out, status, work,
                           work_bytes, stream_);
}

static size_t lindh_smem(int N) {
  const size_t M = (size_t)N * (N - 1) / 2;
  return sizeof(double) * (10 * (size_t)N + M) + sizeof(int) * 40 + (size_t)N * N + 16;
}

extern "C" size_t mop_lindh_workspace_bytes(int B, int natoms) {
  if (B <= 0 || natoms <= 0) return 0;
  int cb, ca, cd;
  fischer_caps(natoms, &cb, &ca, &cd);
  size_t bytes = (size_t)B * (2 * cb + 3 * ca + 4 * cd + 4) * sizeof(int32_t);
  bytes = (bytes + 255) & ~(size_t)255;
  bytes += (size_t)B * (cb + ca + cd) * sizeof(double);
  bytes = (bytes + 255) & ~(size_t)255;
  bytes += (size_t)B * 9 * natoms * natoms * sizeof(double);
  return bytes;
}

// LindhApproxHessian.main without the ill-posed K term (ModelHessian/lindh.py:145-165):
// H_out = project(B^T diag(k) B).  atom_params [B or 1][natoms][6]; kdiag_out (optional)
// [B][natoms (natoms - 1) / 2] = the diagonal RIC force constants of guess_lindh_hessian.
extern "C" int mop_lindh_hessian(int B, int natoms, const double* xyz, const double* atom_params,
                                 int param_stride, double* H_out, double* kdiag_out, int32_t* counts_out,
                                 int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 1, "mop_lindh_hessian: B >= 0 and natoms > 1 required");
  MOP_REQUIRE(xyz && atom_params && H_out && work, "mop_lindh_hessian: xyz, atom_params, H_out, work required");
  MOP_REQUIRE(param_stride == 0 || param_stride == natoms, "mop_lindh_hessian: param_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_lindh_workspace_bytes(B, natoms)) {
    mop_set_error("mop_lindh_hessian: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  const size_t smem = lindh_smem(natoms);
  if (smem > 200 * 1024) {
    mop_set_error("mop_lindh_hessian: natoms = %d too large for the shared-memory pair table", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  int cb, ca, cd;
  fischer_caps(natoms, &cb, &ca, &cd);
  char* w = (char*)work;
  int32_t* bonds = (int32_t*)w;
  int32_t* angles = bonds + (size_t)B * 2 * cb;
  int32_t* dihs = angles + (size_t)B * 3 * ca;
  int32_t* counts = dihs + (size_t)B * 4 * cd;
  size_t off = ((size_t)B * (2 * cb + 3 * ca + 4 * cd + 4) * sizeof(int32_t) + 255) & ~(size_t)255;
  double* fc = (double*)(w + off);
  off += ((size_t)B * (cb + ca + cd) * sizeof(double) + 255) & ~(size_t)255;
  double* Hraw = (double*)(w + off);
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_lindh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_lindh<<<B, mop::MH_THREADS, smem, stream>>>(natoms, xyz, atom_params, param_stride, cb, ca, cd, bonds,
                                                   angles, dihs, counts, fc, kdiag_out, Hraw, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  if (counts_out)
    MOP_CHECK_CUDA(cudaMemcpyAsync(counts_out, counts, sizeof(int32_t) * 3 * (size_t)B,
                                   cudaMemcpyDeviceToDevice, stream));
  return mop_launch_project_trrot(B, 3 * natoms, Hraw, nullptr, xyz, nullptr, H_out, nullptr, nullptr, 0, stream);
}
