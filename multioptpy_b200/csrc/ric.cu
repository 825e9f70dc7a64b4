// Redundant internal coordinates: Coordinate/redundant_coordinate.py.
//   k_ric_bmatrix       all-pairs distance B matrix, rows in itertools.combinations order (:15-43)
//   k_ric_partial_rows  stretch / bend / torsion Wilson rows with the linear / planar branches (:150-320)
//   k_ric_grad_to_cart  B^T q without forming B (:47-50)
//   k_ric_hb / k_ric_bthb   Wilson back-transformation B^T H B + K for a dense or diagonal RIC Hessian (:145)
//   k_ric_kmatrix       K = sum_t q_t d2(coordinate_t)/dx2 over the bond / angle / dihedral tables (:63-143);
//                       the reference differentiates TorchDerivatives.{distance, angle, dihedral_angle}
//                       with torch.func.hessian, here every second derivative is one hyper-dual evaluation
//   k_ric_gram / k_ric_pinv_apply   calc_int_grad_from_pBmat / calc_cart_grad_from_pBmat (:377-439): the
//                       SVD of the symmetric G = pB^T pB is its eigendecomposition (batched eigensolver)
// All kernels are HBM-bound streaming / gather kernels; one CTA (or warp) per structure, row or term.
#include "common.cuh"
#include "hyperdual.cuh"

namespace mop {

__device__ __forceinline__ int pair_index(int i, int j, int N) {  // i < j
  return i * N - i * (i + 1) / 2 + (j - i - 1);
}
__device__ __forceinline__ double dist3(const double* a, const double* b) {
  const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return sqrt(dx * dx + dy * dy + dz * dz);
}

// ---- B matrix -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ric_bmatrix(int N, const double* __restrict__ xyz_all, double* __restrict__ Bm_all) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int n = 3 * N, M = N * (N - 1) / 2;
  const double* xyz = xyz_all + (size_t)b * n;
  double* Bm = Bm_all + (size_t)b * M * n;
  for (int i = blockIdx.x; i < N - 1; i += gridDim.x) {
    for (int j = i + 1 + w; j < N; j += nw) {
      const double r = dist3(xyz + 3 * i, xyz + 3 * j);
      double* row = Bm + (size_t)pair_index(i, j, N) * n;
      for (int c = lane; c < n; c += 32) {
        const int a = c / 3, k = c - 3 * a;
        double v = 0.0;
        if (a == i) v = (xyz[3 * i + k] - xyz[3 * j + k]) / r;
        else if (a == j) v = -1.0 * (xyz[3 * i + k] - xyz[3 * j + k]) / r;
        row[c] = v;
      }
    }
  }
}

// ---- partial rows ----------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3r(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// parts[m][3] for labels (0-based) at[0..m)
__device__ void ric_row_parts(const double* xyz, const int* at, int m, double (*parts)[3]) {
  const double PI = 3.141592653589793;
  if (m == 2) {
    const double* xi = xyz + 3 * at[0];
    const double* xj = xyz + 3 * at[1];
    const double r = dist3(xi, xj);
    for (int c = 0; c < 3; ++c) {
      parts[0][c] = (xi[c] - xj[c]) / r;
      parts[1][c] = -1.0 * (xi[c] - xj[c]) / r;
    }
  } else if (m == 3) {
    double u[3], w[3];
    for (int c = 0; c < 3; ++c) {
      u[c] = xyz[3 * at[0] + c] - xyz[3 * at[1] + c];
      w[c] = xyz[3 * at[2] + c] - xyz[3 * at[1] + c];
    }
    const double lu = sqrt(dot3(u, u)), lw = sqrt(dot3(w, w));
    double cs = dot3(u, w) / (lu * lw);
    cs = fmin(fmax(cs, -1.0), 1.0);
    const double th = acos(cs);
    if (fabs(th) > PI - 1e-6) {
      for (int c = 0; c < 3; ++c) {
        parts[0][c] = (PI - th) / (2.0 * lu * lu) * u[c];
        parts[1][c] = (1.0 / lu - 1.0 / lw) * (PI - th) / (2.0 * lu) * u[c];
        parts[2][c] = (PI - th) / (2.0 * lw * lw) * w[c];
      }
    } else {
      const double ct = 1.0 / tan(th), st = sin(th);
      for (int c = 0; c < 3; ++c) {
        parts[0][c] = ct * u[c] / (lu * lu) - w[c] / (lu * lw * st);
        parts[1][c] = (u[c] + w[c]) / (lu * lw * st) - ct * (u[c] / (lu * lu) + w[c] / (lw * lw));
        parts[2][c] = ct * w[c] / (lw * lw) - u[c] / (lu * lw * st);
      }
    }
  } else {
    double vij[3], vlk[3], vkj[3], ukj[3], a1[3], a2[3];
    for (int c = 0; c < 3; ++c) {
      vij[c] = xyz[3 * at[0] + c] - xyz[3 * at[1] + c];
      vlk[c] = xyz[3 * at[3] + c] - xyz[3 * at[2] + c];
      vkj[c] = xyz[3 * at[2] + c] - xyz[3 * at[1] + c];
    }
    const double nkj = sqrt(dot3(vkj, vkj));
    for (int c = 0; c < 3; ++c) ukj[c] = vkj[c] / nkj;
    const double dij = dot3(vij, ukj), dlk = dot3(vlk, ukj);
    for (int c = 0; c < 3; ++c) {
      a1[c] = vij[c] - dij * ukj[c];
      a2[c] = vlk[c] - dlk * ukj[c];
    }
    const double n1 = sqrt(dot3(a1, a1)), n2 = sqrt(dot3(a2, a2));
    double cr[3];
    cross3r(vij, vkj, cr);
    const double det = dot3(vlk, cr);  // det [vlk; vij; vkj]
    const double sg = det > 0.0 ? 1.0 : (det < 0.0 ? -1.0 : 1.0);
    double cs = dot3(a1, a2) / (n1 * n2);
    cs = fmin(fmax(cs, -1.0), 1.0);
    const double phi = acos(cs) * sg;
    const double A = dij / nkj, Bc = dlk / nkj;
    const bool near_pi = fabs(phi) > PI - 1e-6, near_0 = fabs(phi) < 1e-6;
    if (near_pi || near_0) {
      double G[3];
      cross3r(vkj, a1, G);
      const double nG = sqrt(dot3(G, G));
      for (int c = 0; c < 3; ++c) {
        const double uG = G[c] / nG;
        parts[0][c] = uG / n1;
        parts[1][c] = -((1.0 - A) / n1 - Bc / n2) * uG;
        parts[2][c] = -((1.0 + Bc) / n2 + A / n1) * uG;
        parts[3][c] = near_pi ? uG / n2 : -1.0 * uG / n2;
      }
    } else {
      const double ct = 1.0 / tan(phi), st = sin(phi);
      for (int c = 0; c < 3; ++c) {
        parts[0][c] = ct * a1[c] / (n1 * n1) - a2[c] / (n1 * n2 * st);
        parts[1][c] = ((1.0 - A) * a2[c] - Bc * a1[c]) / (n1 * n2 * st) -
                      ct * ((1.0 - A) * a1[c] / (n1 * n1) - Bc * a2[c] / (n2 * n2));
        parts[2][c] = ((1.0 + Bc) * a1[c] + A * a2[c]) / (n1 * n2 * st) -
                      ct * ((1.0 + Bc) * a2[c] / (n2 * n2) + A * a1[c] / (n1 * n1));
        parts[3][c] = ct * a2[c] / (n2 * n2) - a1[c] / (n1 * n2 * st);
      }
    }
  }
}

// one warp per (structure, row); labels [nrows][4], 1-based, 0 = unused
__global__ void __launch_bounds__(128) k_ric_partial_rows(int N, int nrows, const double* __restrict__ xyz_all,
                                                          const int32_t* __restrict__ labels, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (gwarp >= nrows) return;
  const int n = 3 * N;
  const double* xyz = xyz_all + (size_t)b * n;
  int at[4], m = 0;
  for (int q = 0; q < 4; ++q) {
    const int l = labels[4 * gwarp + q];
    if (l > 0) at[m++] = l - 1;
  }
  double parts[4][3];
  ric_row_parts(xyz, at, m, parts);
  double* row = out + ((size_t)b * nrows + gwarp) * n;
  for (int c = lane; c < n; c += 32) {
    const int a = c / 3, k = c - 3 * a;
    double v = 0.0;
    for (int q = m - 1; q >= 0; --q)  // the first matching label wins, as the reference's if / elif chain
      if (a == at[q]) v = parts[q][k];
    row[c] = v;
  }
}

// ---- gradient back-transformation: g_i = sum_{j != i} (x_i - x_j) / r_ij * q_{pair(i,j)} ---------
__global__ void __launch_bounds__(128) k_ric_grad_to_cart(int N, const double* __restrict__ xyz_all,
                                                          const double* __restrict__ q_all, double* __restrict__ g_all) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= N) return;
  const int n = 3 * N, M = N * (N - 1) / 2;
  const double* xyz = xyz_all + (size_t)b * n;
  const double* q = q_all + (size_t)b * M;
  double g[3] = {0.0, 0.0, 0.0};
  for (int j = lane; j < N; j += 32) {
    if (j == i) continue;
    const double r = dist3(xyz + 3 * i, xyz + 3 * j);
    const double qq = q[i < j ? pair_index(i, j, N) : pair_index(j, i, N)];
    for (int c = 0; c < 3; ++c) g[c] = fma((xyz[3 * i + c] - xyz[3 * j + c]) / r, qq, g[c]);
  }
  for (int c = 0; c < 3; ++c) g[c] = warp_sum(g[c]);
  if (lane == 0)
    for (int c = 0; c < 3; ++c) g_all[(size_t)b * n + 3 * i + c] = g[c];
}

// ---- Wilson back-transformation ------------------------------------------------------------------
// T = Hric B (M x n): T[p][3 bb + d] = sum_{o != bb} Hric[p][pair(bb,o)] * (x_bb - x_o)_d / r
__global__ void __launch_bounds__(256) k_ric_hb(int N, const double* __restrict__ xyz_all, const double* __restrict__ Hric_all,
                                                double* __restrict__ T_all) {
  extern __shared__ double sm[];  // xyz
  const int b = blockIdx.y, n = 3 * N, M = N * (N - 1) / 2;
  const double* Hric = Hric_all + (size_t)b * M * M;
  double* T = T_all + (size_t)b * M * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = xyz_all[(size_t)b * n + i];
  __syncthreads();
  for (int p = blockIdx.x; p < M; p += gridDim.x) {
    const double* hrow = Hric + (size_t)p * M;
    for (int bb = threadIdx.x; bb < N; bb += blockDim.x) {
      double acc[3] = {0.0, 0.0, 0.0};
      for (int o = 0; o < N; ++o) {
        if (o == bb) continue;
        const double r = dist3(sm + 3 * bb, sm + 3 * o);
        const double h = hrow[bb < o ? pair_index(bb, o, N) : pair_index(o, bb, N)];
        for (int d = 0; d < 3; ++d) acc[d] = fma(h, (sm[3 * bb + d] - sm[3 * o + d]) / r, acc[d]);
      }
      for (int d = 0; d < 3; ++d) T[(size_t)p * n + 3 * bb + d] = acc[d];
    }
  }
}

// H[3a + c][:] = sum_{o != a} (x_a - x_o)_c / r * T[pair(a,o)][:]  (+ K); diag != 0: Hric is a
// diagonal (T row p = hd[p] * B row p, formed on the fly)
__global__ void __launch_bounds__(256) k_ric_bthb(int N, int diag, const double* __restrict__ xyz_all,
                                                  const double* __restrict__ T_all, const double* __restrict__ hd_all,
                                                  const double* __restrict__ K_all, double* __restrict__ H_all) {
  extern __shared__ double sm[];
  const int b = blockIdx.y, a = blockIdx.x, n = 3 * N, M = N * (N - 1) / 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = xyz_all[(size_t)b * n + i];
  __syncthreads();
  double* H = H_all + (size_t)b * n * n;
  const double* K = K_all ? K_all + (size_t)b * n * n : nullptr;
  for (int col = threadIdx.x; col < n; col += blockDim.x) {
    double acc[3] = {0.0, 0.0, 0.0};
    if (!diag) {
      const double* T = T_all + (size_t)b * M * n;
      for (int o = 0; o < N; ++o) {
        if (o == a) continue;
        const double r = dist3(sm + 3 * a, sm + 3 * o);
        const double t = T[(size_t)(a < o ? pair_index(a, o, N) : pair_index(o, a, N)) * n + col];
        for (int c = 0; c < 3; ++c) acc[c] = fma((sm[3 * a + c] - sm[3 * o + c]) / r, t, acc[c]);
      }
    } else {
      const double* hd = hd_all + (size_t)b * M;
      const int bb = col / 3, d = col - 3 * bb;
      if (bb != a) {  // only the pair (a, bb) couples the two atoms
        const double r = dist3(sm + 3 * a, sm + 3 * bb);
        const double h = hd[a < bb ? pair_index(a, bb, N) : pair_index(bb, a, N)];
        const double ed = (sm[3 * a + d] - sm[3 * bb + d]) / r;
        for (int c = 0; c < 3; ++c) acc[c] = -h * ((sm[3 * a + c] - sm[3 * bb + c]) / r) * ed;
      } else {
        for (int o = 0; o < N; ++o) {
          if (o == a) continue;
          const double r = dist3(sm + 3 * a, sm + 3 * o);
          const double h = hd[a < o ? pair_index(a, o, N) : pair_index(o, a, N)];
          const double ed = (sm[3 * a + d] - sm[3 * o + d]) / r;
          for (int c = 0; c < 3; ++c) acc[c] = fma(h * ((sm[3 * a + c] - sm[3 * o + c]) / r), ed, acc[c]);
        }
      }
    }
    for (int c = 0; c < 3; ++c) {
      const size_t e = (size_t)(3 * a + c) * n + col;
      H[e] = acc[c] + (K ? K[e] : 0.0);
    }
  }
}

// ---- K matrix: hyper-dual second derivatives --------------------------------------------------------
// TorchDerivatives.distance / angle / dihedral_angle (:442-477) on m atoms
__device__ HD ric_coordinate(const HD (*c)[3], int m) {
  if (m == 2) {
    HD d[3];
    for (int k = 0; k < 3; ++k) d[k] = c[0][k] - c[1][k];
    return hd_sqrt(hd_dot(d, d));
  }
  if (m == 3) {
    HD v1[3], v2[3];
    for (int k = 0; k < 3; ++k) {
      v1[k] = c[0][k] - c[1][k];
      v2[k] = c[2][k] - c[1][k];
    }
    const HD den = hd_sqrt(hd_dot(v1, v1)) * hd_sqrt(hd_dot(v2, v2)) + hd_const(1e-15);
    return hd_acos(hd_dot(v1, v2) / den);
  }
  HD a1[3], a2[3], a3[3], v1[3], v2[3];
  for (int k = 0; k < 3; ++k) {
    a1[k] = c[1][k] - c[0][k];
    a2[k] = c[2][k] - c[1][k];
    a3[k] = c[3][k] - c[2][k];
  }
  hd_cross(a1, a2, v1);
  hd_cross(a2, a3, v2);
  const HD i1 = hd_recip(hd_sqrt(hd_dot(v1, v1))), i2 = hd_recip(hd_sqrt(hd_dot(v2, v2)));
  for (int k = 0; k < 3; ++k) {
    v1[k] = v1[k] * i1;
    v2[k] = v2[k] * i2;
  }
  const HD s2 = hd_dot(v2, v2);
  HD den = hd_const(0.0);
  for (int k = 0; k < 3; ++k) den = den + (v1[k] * v1[k] * s2 + hd_const(1e-15));
  return hd_abs(hd_acos(hd_dot(v1, v2) / hd_sqrt(den)));
}

// thread per (term, coordinate pair a <= b); terms = bonds, angles, dihedrals in table order
__global__ void __launch_bounds__(128) k_ric_kmatrix(int N, const double* __restrict__ xyz_all,
                                                     const int32_t* __restrict__ bonds, const int32_t* __restrict__ angles,
                                                     const int32_t* __restrict__ dihs, const int32_t* __restrict__ counts,
                                                     int capB, int capA, int capD, int table_stride,
                                                     const double* __restrict__ q_all, int Mq, double* __restrict__ K_all) {
  const int b = blockIdx.y, n = 3 * N;
  const int32_t* cnt = counts + (size_t)(table_stride ? b : 0) * 3;
  const int nb = cnt[0], na = cnt[1], nd = cnt[2];
  const int32_t* tb = bonds + (size_t)(table_stride ? b : 0) * capB * 2;
  const int32_t* ta = angles + (size_t)(table_stride ? b : 0) * capA * 3;
  const int32_t* td = dihs + (size_t)(table_stride ? b : 0) * capD * 4;
  const double* xyz = xyz_all + (size_t)b * n;
  double* K = K_all + (size_t)b * n * n;
  const int nterm = nb + na + nd;
  for (int work = blockIdx.x * blockDim.x + threadIdx.x; work < nterm * 78; work += gridDim.x * blockDim.x) {
    const int t = work / 78, slot = work - t * 78;
    int at[4], m;
    if (t < nb) { m = 2; at[0] = tb[2 * t]; at[1] = tb[2 * t + 1]; }
    else if (t < nb + na) { m = 3; const int32_t* r = ta + 3 * (t - nb); at[0] = r[0]; at[1] = r[1]; at[2] = r[2]; }
    else { m = 4; const int32_t* r = td + 4 * (t - nb - na); at[0] = r[0]; at[1] = r[1]; at[2] = r[2]; at[3] = r[3]; }
    const int nc = 3 * m;
    if (slot >= nc * (nc + 1) / 2 || t >= Mq) continue;
    int ca = 0, rem = slot;  // slot -> (ca <= cb)
    while (rem >= nc - ca) {
      rem -= nc - ca;
      ++ca;
    }
    const int cb = ca + rem;
    HD c[4][3];
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < 3; ++k) {
        const int ci = 3 * i + k;
        c[i][k] = HD{xyz[3 * at[i] + k], ci == ca ? 1.0 : 0.0, ci == cb ? 1.0 : 0.0, 0.0};
      }
    const double h = ric_coordinate(c, m).ab * q_all[(size_t)b * Mq + t];
    const int ga = 3 * at[ca / 3] + ca % 3, gb = 3 * at[cb / 3] + cb % 3;
    atomicAdd(&K[(size_t)ga * n + gb], h);
    if (ca != cb) atomicAdd(&K[(size_t)gb * n + ga], h);
  }
}

// ---- pseudo-inverse gradient transforms --------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ric_gram(int n, int m, const double* __restrict__ pB_all, double* __restrict__ G_all) {
  const int b = blockIdx.y;
  const double* pB = pB_all + (size_t)b * m * n;
  double* G = G_all + (size_t)b * n * n;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int i = e / n, j = e - i * n;
    double acc = 0.0;
    for (int r = 0; r < m; ++r) acc = fma(pB[(size_t)r * n + i], pB[(size_t)r * n + j], acc);
    G[e] = acc;
  }
}

// int_grad = pB (V f(L) V^T g), f(s) = 1/s for s > 1e-6 else s (calc_inv_G_mat :381-394)
__global__ void __launch_bounds__(256) k_ric_pinv_apply(int n, int m, const double* __restrict__ pB_all,
                                                        const double* __restrict__ evals_all, const double* __restrict__ V_all,
                                                        const double* __restrict__ g_all, double* __restrict__ out_all) {
  extern __shared__ double sm[];
  double* y = sm;       // n
  double* z = sm + n;   // n
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const double* V = V_all + (size_t)b * n * n;
  const double* g = g_all + (size_t)b * n;
  for (int k = w; k < n; k += nw) {
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(V[(size_t)k * n + i], g[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const double s = fabs(evals_all[(size_t)b * n + k]);  // singular value of the PSD Gram matrix
      y[k] = (s > 1e-6 ? 1.0 / s : s) * acc;
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc = fma(V[(size_t)k * n + i], y[k], acc);
    z[i] = acc;
  }
  __syncthreads();
  const double* pB = pB_all + (size_t)b * m * n;
  for (int r = w; r < m; r += nw) {
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(pB[(size_t)r * n + i], z[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) out_all[(size_t)b * m + r] = acc;
  }
}

// cart_grad = pB^T int_grad
__global__ void __launch_bounds__(256) k_ric_pb_t(int n, int m, const double* __restrict__ pB_all,
                                                  const double* __restrict__ q_all, double* __restrict__ g_all) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < m; ++r) acc = fma(pB_all[((size_t)b * m + r) * n + i], q_all[(size_t)b * m + r], acc);
    g_all[(size_t)b * n + i] = acc;
  }
}

}  // namespace mop

extern "C" size_t mop_eigh_workspace_bytes(int B, int n, int algo);
extern "C" int mop_eigh(int B, int n, int algo, const double* A, double* evals, double* evecs, int32_t* status,
                        void* work, size_t work_bytes, void* stream);
static size_t ric_al(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int mop_ric_bmatrix(int B, int natoms, const double* xyz, double* Bmat, void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 1 && xyz && Bmat, "mop_ric_bmatrix: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid(natoms - 1, B);
  mop::k_ric_bmatrix<<<grid, 256, 0, (cudaStream_t)stream>>>(natoms, xyz, Bmat);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_ric_partial_rows(int B, int natoms, const double* xyz, int nrows, const int32_t* labels, double* rows_out,
                                    void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 1 && nrows >= 0 && xyz && labels && rows_out, "mop_ric_partial_rows: bad arguments");
  if (B == 0 || nrows == 0) return MOP_OK;
  dim3 grid((nrows + 3) / 4, B);
  mop::k_ric_partial_rows<<<grid, 128, 0, (cudaStream_t)stream>>>(natoms, nrows, xyz, labels, rows_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_ric_grad_to_cart(int B, int natoms, const double* xyz, const double* ric_grad, double* cart_grad,
                                    void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 1 && xyz && ric_grad && cart_grad, "mop_ric_grad_to_cart: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid((natoms + 3) / 4, B);
  mop::k_ric_grad_to_cart<<<grid, 128, 0, (cudaStream_t)stream>>>(natoms, xyz, ric_grad, cart_grad);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" size_t mop_ric_hess_workspace_bytes(int B, int natoms, int diagonal) {
  if (B <= 0 || natoms <= 1 || diagonal) return 0;
  return sizeof(double) * (size_t)B * (natoms * (natoms - 1) / 2) * 3 * natoms;
}

extern "C" int mop_ric_hess_to_cart(int B, int natoms, const double* xyz, const double* ric_hess, int diagonal,
                                    const double* K, double* cart_hess, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 1 && xyz && ric_hess && cart_hess, "mop_ric_hess_to_cart: bad arguments");
  if (B == 0) return MOP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t smem = sizeof(double) * 3 * (size_t)natoms;
  const int M = natoms * (natoms - 1) / 2;
  if (!diagonal) {
    if (!work || work_bytes < mop_ric_hess_workspace_bytes(B, natoms, 0)) {
      mop_set_error("mop_ric_hess_to_cart: workspace too small");
      return MOP_ERR_WORKSPACE;
    }
    dim3 g1(M < 1024 ? M : 1024, B);
    mop::k_ric_hb<<<g1, 256, smem, stream>>>(natoms, xyz, ric_hess, (double*)work);
    MOP_CHECK_CUDA(cudaGetLastError());
  }
  dim3 g2(natoms, B);
  mop::k_ric_bthb<<<g2, 256, smem, stream>>>(natoms, diagonal ? 1 : 0, xyz, (const double*)work, ric_hess, K, cart_hess);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_ric_kmatrix(int B, int natoms, const double* xyz, const int32_t* bonds, const int32_t* angles,
                               const int32_t* dihedrals, const int32_t* counts, int cap_bonds, int cap_angles,
                               int cap_dihedrals, int tables_per_structure, const double* ric_grad, int ric_len,
                               double* K_out, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 1 && xyz && bonds && angles && dihedrals && counts && ric_grad && K_out,
              "mop_ric_kmatrix: bad arguments");
  if (B == 0) return MOP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t n = 3 * (size_t)natoms;
  MOP_CHECK_CUDA(cudaMemsetAsync(K_out, 0, sizeof(double) * (size_t)B * n * n, stream));
  const int maxterm = cap_bonds + cap_angles + cap_dihedrals;
  int gx = (maxterm * 78 + 127) / 128;
  if (gx < 1) gx = 1;
  if (gx > 4096) gx = 4096;
  dim3 grid(gx, B);
  mop::k_ric_kmatrix<<<grid, 128, 0, stream>>>(natoms, xyz, bonds, angles, dihedrals, counts, cap_bonds, cap_angles,
                                              cap_dihedrals, tables_per_structure ? 1 : 0, ric_grad, ric_len, K_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// workspace: G | V | evals | eigh work
extern "C" size_t mop_ric_pb_workspace_bytes(int B, int n) {
  if (B <= 0 || n <= 0) return 0;
  return 2 * ric_al(sizeof(double) * (size_t)B * n * n) + ric_al(sizeof(double) * (size_t)B * n) +
         mop_eigh_workspace_bytes(B, n, MOP_EIGH_AUTO);
}

extern "C" int mop_ric_pb_int_grad(int B, int n, int m, const double* pB, const double* cart_grad, double* int_grad,
                                   int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0 && m > 0 && pB && cart_grad && int_grad && status && work, "mop_ric_pb_int_grad: bad arguments");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_ric_pb_workspace_bytes(B, n)) {
    mop_set_error("mop_ric_pb_int_grad: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = ric_al(sizeof(double) * (size_t)B * n * n), nv = ric_al(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  double* G = (double*)w;
  double* V = (double*)(w + nn);
  double* ev = (double*)(w + 2 * nn);
  MOP_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, stream));
  dim3 g1((n * n + 255) / 256 < 64 ? (n * n + 255) / 256 : 64, B);
  mop::k_ric_gram<<<g1, 256, 0, stream>>>(n, m, pB, G);
  MOP_CHECK_CUDA(cudaGetLastError());
  int rc = mop_eigh(B, n, MOP_EIGH_AUTO, G, ev, V, status, w + 2 * nn + nv, work_bytes - (2 * nn + nv), stream);
  if (rc != MOP_OK) return rc;
  mop::k_ric_pinv_apply<<<B, 256, sizeof(double) * 2 * (size_t)n, stream>>>(n, m, pB, ev, V, cart_grad, int_grad);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_ric_pb_cart_grad(int B, int n, int m, const double* pB, const double* int_grad, double* cart_grad,
                                    void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && m > 0 && pB && int_grad && cart_grad, "mop_ric_pb_cart_grad: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid((n + 255) / 256, B);
  mop::k_ric_pb_t<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, pB, int_grad, cart_grad);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
