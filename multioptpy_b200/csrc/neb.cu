// NEB pieces around the per-image quasi-Newton step.
//
// Images are a contiguous block [first, first + nloc) of the nimg-image chain; every
// per-image quantity that needs a neighbour reads it from a HALO layout: arrays of
// nloc + 2 entries whose slots 0 and nloc + 1 hold images first-1 and first+nloc
// (filled by the NCCL halo exchange when the chain is sharded over GPUs; unused at the
// chain ends).  One CTA per image.
//   mop_bneb_force    : CaluculationBNEB.calc_force (MEP/pathopt_bneb_force.py:33-117)
//   mop_neb_ayala     : calculate_gamma + H += gamma t t^T (pathopt_bneb_force.py:161-222,
//                       Optimizer/rfo_neb.py:43-73)
//   mop_neb_limit_tr  : _limit_step_size + TR_NEB.TR_calc (rfo_neb.py:76-83,
//                       Optimizer/trust_radius_neb.py:17-98)
#include "common.cuh"

namespace mop {

constexpr int NEB_THREADS = 128;

// projection along the per-atom unit vectors from image `a` to image `bimg`:
// returns sum_atoms u (u . g) for this thread's atoms into out (accumulated with weight w).
// B-matrix rows u_i = (x_b,i - x_a,i) / (|x_a,i - x_b,i| + 1e-15); through the SVD
// pseudo-inverse of B^T B (redundant_coordinate.py:381-400,432-439) the projected part of
// the gradient is  u_hat (u_hat . g)  per atom (nothing when the atoms coincide).
__device__ __forceinline__ void tangent_projection(int N, const double* xa, const double* xb,
                                                   const double* g, double wgt, double* proj) {
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const double dx = xb[3 * a] - xa[3 * a], dy = xb[3 * a + 1] - xa[3 * a + 1], dz = xb[3 * a + 2] - xa[3 * a + 2];
    const double nrm = sqrt(dx * dx + dy * dy + dz * dz);
    const double den = nrm + 1e-15;
    const double ux = dx / den, uy = dy / den, uz = dz / den;
    const double s = ux * ux + uy * uy + uz * uz;  // singular value of the 3x3 block u u^T
    if (s > 1e-6) {
      const double c = (ux * g[3 * a] + uy * g[3 * a + 1] + uz * g[3 * a + 2]) / s;  // int_grad
      proj[3 * a] -= wgt * c * ux;       // calc_cart_grad_from_pBmat(-w * int_grad, B)
      proj[3 * a + 1] -= wgt * c * uy;
      proj[3 * a + 2] -= wgt * c * uz;
    }
  }
}

__global__ void __launch_bounds__(NEB_THREADS)
k_bneb_force(int nimg, int first, int n, const double* __restrict__ xh, const double* __restrict__ Eh,
             const double* __restrict__ g_all, double* __restrict__ force, double* __restrict__ tau) {
  extern __shared__ double sm[];
  const int l = blockIdx.x, i = first + l, tid = threadIdx.x, N = n / 3;
  const double* g = g_all + (size_t)l * n;
  double* proj = sm;  // n : projection_grad (= tangent_grad)
  for (int k = tid; k < n; k += NEB_THREADS) proj[k] = 0.0;
  __syncthreads();
  if (i == 0 || i == nimg - 1) {  // endpoints: force = -g, tau = 0 (:40-47)
    for (int k = tid; k < n; k += NEB_THREADS) {
      force[(size_t)l * n + k] = -g[k];
      tau[(size_t)l * n + k] = 0.0;
    }
    return;
  }
  const double* x1 = xh + (size_t)l * n;        // image i-1
  const double* x2 = xh + (size_t)(l + 1) * n;  // image i
  const double* x3 = xh + (size_t)(l + 2) * n;  // image i+1
  const double e0 = Eh[l], e1 = Eh[l + 1], e2 = Eh[l + 2];
  if (e0 < e1 && e1 < e2) {
    tangent_projection(N, x2, x3, g, 1.0, proj);
  } else if (e0 > e1 && e1 > e2) {
    tangent_projection(N, x1, x2, g, 1.0, proj);
  } else {
    const double mx = fmax(fabs(e2 - e1), fabs(e1 - e0)), mn = fmin(fabs(e2 - e1), fabs(e1 - e0));
    const double a = mx / (mx + mn + 1e-8), b = mn / (mx + mn + 1e-8);
    if (e0 < e2) {
      tangent_projection(N, x2, x3, g, a, proj);
      tangent_projection(N, x1, x2, g, b, proj);
    } else {
      tangent_projection(N, x2, x3, g, b, proj);
      tangent_projection(N, x1, x2, g, a, proj);
    }
  }
  __syncthreads();
  for (int k = tid; k < n; k += NEB_THREADS) {
    force[(size_t)l * n + k] = -(g[k] + proj[k]);  // total_force = -proj_grad (:62)
    tau[(size_t)l * n + k] = proj[k];
  }
}

// 6x6 dense solve with partial pivoting (numpy.linalg.solve / LAPACK dgesv semantics)
__device__ bool solve6(double A[6][6], double b[6]) {
  for (int c = 0; c < 6; ++c) {
    int p = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < 6; ++r)
      if (fabs(A[r][c]) > best) {
        best = fabs(A[r][c]);
        p = r;
      }
    if (best == 0.0) return false;
    if (p != c) {
      for (int k = 0; k < 6; ++k) {
        const double t = A[c][k];
        A[c][k] = A[p][k];
        A[p][k] = t;
      }
      const double t = b[c];
      b[c] = b[p];
      b[p] = t;
    }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r][c] / A[c][c];
      if (f != 0.0) {
        for (int k = c + 1; k < 6; ++k) A[r][k] -= f * A[c][k];
        b[r] -= f * b[c];
      }
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < 6; ++k) s -= A[r][k] * b[k];
    b[r] = s / A[r][r];
  }
  return true;
}

__global__ void __launch_bounds__(256)
k_neb_ayala(int nimg, int first, int n, const double* __restrict__ xh, const double* __restrict__ Eh,
            const double* __restrict__ gh, const double* __restrict__ tau_all, double* __restrict__ H_all,
            double* __restrict__ gamma_out) {
  __shared__ double scratch[40];
  __shared__ double s_gamma;
  const int l = blockIdx.x, i = first + l, tid = threadIdx.x;
  if (i == 0 || i == nimg - 1) {  // endpoints keep their Hessian (rfo_neb.py:50-51)
    if (tid == 0 && gamma_out) gamma_out[l] = 0.0;
    return;
  }
  const double* qp = xh + (size_t)l * n;
  const double* qc = xh + (size_t)(l + 1) * n;
  const double* qn = xh + (size_t)(l + 2) * n;
  const double* gp = gh + (size_t)l * n;
  const double* gc = gh + (size_t)(l + 1) * n;
  const double* gn = gh + (size_t)(l + 2) * n;
  const double* t = tau_all + (size_t)l * n;
  double dp2 = 0, dn2 = 0, a_gp = 0, a_gc = 0, a_gn = 0;
  for (int k = tid; k < n; k += blockDim.x) {
    const double dpk = qc[k] - qp[k], dnk = qn[k] - qc[k];
    dp2 = fma(dpk, dpk, dp2);
    dn2 = fma(dnk, dnk, dn2);
    a_gp = fma(gp[k], dpk, a_gp);
    a_gn = fma(gn[k], dnk, a_gn);
    a_gc = fma(gc[k], t[k], a_gc);
  }
  dp2 = block_sum(dp2, scratch);
  dn2 = block_sum(dn2, scratch);
  a_gp = block_sum(a_gp, scratch);
  a_gn = block_sum(a_gn, scratch);
  a_gc = block_sum(a_gc, scratch);
  if (tid == 0) {
    const double dprev = sqrt(dp2), dnext = sqrt(dn2);
    double gamma = 0.0;
    if (!(dprev < 1e-6 || dnext < 1e-6)) {
      const double sp = -dprev, sc = 0.0, sn = dnext;
      const double s3[3] = {sp, sc, sn};
      double A[6][6], b[6];
      for (int r = 0; r < 3; ++r) {
        double pw = 1.0;
        for (int c = 0; c < 6; ++c) {
          A[r][c] = pw;
          pw *= s3[r];
        }
        A[3 + r][0] = 0.0;
        pw = 1.0;
        for (int c = 1; c < 6; ++c) {
          A[3 + r][c] = c * pw;
          pw *= s3[r];
        }
      }
      b[0] = Eh[l]; b[1] = Eh[l + 1]; b[2] = Eh[l + 2];
      b[3] = a_gp / dprev;  // g_prev . (q_curr - q_prev)/dist_prev
      b[4] = a_gc;          // g_curr . tangent (as given, not normalised)
      b[5] = a_gn / dnext;
      if (solve6(A, b)) gamma = 2.0 * b[2];
    }
    s_gamma = gamma;
    if (gamma_out) gamma_out[l] = gamma;
  }
  __syncthreads();
  const double gamma = s_gamma;
  if (gamma == 0.0) return;
  double* H = H_all + (size_t)l * n * n;  // H += gamma |t><t|
  for (size_t e = tid; e < (size_t)n * n; e += blockDim.x) {
    const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
    H[e] += gamma * (t[r] * t[c]);
  }
}

__global__ void __launch_bounds__(NEB_THREADS)
k_neb_limit_tr(int nimg, int first, int n, int fix_init, int fix_end, int step_limit, const double* __restrict__ xh,
               const double* __restrict__ g_all, double* __restrict__ delta_all) {
  __shared__ double scratch[40];
  const int l = blockIdx.x, i = first + l, tid = threadIdx.x;
  double* d = delta_all + (size_t)l * n;
  const bool endpoint = (i == 0 || i == nimg - 1);
  double p = 0.0;
  for (int k = tid; k < n; k += NEB_THREADS) p = fma(d[k], d[k], p);
  double nrm = sqrt(block_sum(p, scratch));
  // _limit_step_size (rfo_neb.py:76-83)
  double scale = 1.0;
  if (step_limit && nrm > 1e-8) scale = fmin(endpoint ? 0.2 : 0.1, nrm) / nrm;
  nrm *= scale;
  if (endpoint) {  // TR_calc ends (:18-27, :85-93)
    double f = scale;
    if ((i == 0 && fix_init) || (i == nimg - 1 && fix_end) || nrm < 1e-15) f = 0.0;
    else f = scale * fmin(0.5, nrm) / nrm;
    for (int k = tid; k < n; k += NEB_THREADS) d[k] *= f;
    return;
  }
  const double* x1 = xh + (size_t)l * n;
  const double* x2 = xh + (size_t)(l + 1) * n;
  const double* x3 = xh + (size_t)(l + 2) * n;
  const double* g = g_all + (size_t)l * n;
  double d1 = 0, d2 = 0, c1 = 0, c2 = 0, fd = 0, ff = 0;
  for (int k = tid; k < n; k += NEB_THREADS) {
    const double a = x1[k] - x2[k], b = x3[k] - x2[k], dk = d[k] * scale;
    d1 = fma(a, a, d1);
    d2 = fma(b, b, d2);
    c1 = fma(a, dk, c1);
    c2 = fma(b, dk, c2);
    fd = fma(g[k], dk, fd);
    ff = fma(g[k], g[k], ff);
  }
  d1 = sqrt(block_sum(d1, scratch));
  d2 = sqrt(block_sum(d2, scratch));
  c1 = block_sum(c1, scratch);
  c2 = block_sum(c2, scratch);
  fd = block_sum(fd, scratch);
  ff = sqrt(block_sum(ff, scratch));
  const double tr1 = d1 / 2.0, tr2 = d2 / 2.0;
  const double cos1 = c1 / ((d1 + 1e-15) * nrm), cos2 = c2 / ((d2 + 1e-15) * nrm);
  const double fcos = fd / (ff * nrm);
  double f = scale;
  if (fcos >= 0.0) {
    if ((cos1 > 0 && cos2 < 0) || (cos1 < 0 && cos2 > 0)) {
      if (nrm > tr1 && cos1 > 0) f = scale * tr1 / nrm;
      else if (nrm > tr2 && cos2 > 0) f = scale * tr2 / nrm;
    } else if (cos1 < 0 && cos2 < 0) {
      // keep
    } else {
      if (nrm > tr1) f = scale * tr1 / nrm;
      else if (nrm > tr2) f = scale * tr2 / nrm;
    }
  } else {
    f = 0.0;  // "no displacements"
  }
  for (int k = tid; k < n; k += NEB_THREADS) d[k] *= f;
}


// ---- FIRE optimizer of the NEB driver (Optimizer/fire_neb.py:38-92) -----------------------------
// blend: per atom  v <- (1 - a) v + a |v| / |F| F  (kept when |F| <= 1e-10), and the power
// P = sum v_prev . F accumulated into one device scalar (the caller all-reduces it over ranks);
// advance: v_new = (reset ? 0 : v_blend) + dt F,  delta = dt (v_new + v_prev) or dt v_new.
__global__ void __launch_bounds__(256) k_neb_fire_blend(int natoms_total, double a, const double* __restrict__ F,
                                                        const double* __restrict__ V, const double* __restrict__ Vp,
                                                        double* __restrict__ Vneb, double* __restrict__ power) {
  __shared__ double scratch[40];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double p = 0.0;
  if (i < natoms_total) {
    const double fx = F[3 * i], fy = F[3 * i + 1], fz = F[3 * i + 2];
    const double vx = V[3 * i], vy = V[3 * i + 1], vz = V[3 * i + 2];
    const double fn = sqrt(fx * fx + fy * fy + fz * fz), vn = sqrt(vx * vx + vy * vy + vz * vz);
    double ox = vx, oy = vy, oz = vz;
    if (fn > 1e-10) {
      const double r = a * (vn / fn);
      ox = (1.0 - a) * vx + r * fx;
      oy = (1.0 - a) * vy + r * fy;
      oz = (1.0 - a) * vz + r * fz;
    }
    Vneb[3 * i] = ox; Vneb[3 * i + 1] = oy; Vneb[3 * i + 2] = oz;
    if (Vp) p = Vp[3 * i] * fx + Vp[3 * i + 1] * fy + Vp[3 * i + 2] * fz;
  }
  p = block_sum(p, scratch);
  if (threadIdx.x == 0 && power && Vp) atomicAdd(power, p);
}

__global__ void __launch_bounds__(256) k_neb_fire_advance(size_t total, double dt, int reset, const double* __restrict__ Vneb,
                                                          const double* __restrict__ F, const double* __restrict__ Vp,
                                                          double* __restrict__ Vnew, double* __restrict__ delta) {
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const double v = (reset ? 0.0 * Vneb[e] : Vneb[e]) + dt * F[e];
    Vnew[e] = v;
    delta[e] = Vp ? dt * (v + Vp[e]) : dt * v;
  }
}

// Image redistribution at equal arc length (Interpolation/linear_interpolation.py:308-336 distribute_geometry with
// Utils/calc_tools.py:853-862 calc_path_length_list; the `align_distances` strategy of NEB._align_geometries,
// neb.py:649-760).  One CTA for the whole chain: warp per segment for the centroid-free segment lengths, thread 0 for
// the running path length (the reference's own summation order), warp per output image for the segment search
// (first j with s_j <= i L / (M - 1) <= s_j+1) and the linear interpolation.  Writes images [first, first + nloc).
__global__ void __launch_bounds__(256) k_neb_redistribute(int nimg, int natoms, int first, int nloc,
                                                          const double* __restrict__ x, double* __restrict__ xout,
                                                          double* __restrict__ plen_out) {
  extern __shared__ double pl[];  // nimg running path lengths
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  const int n = 3 * natoms;
  for (int i = wid; i < nimg - 1; i += nw) {
    const double* xi = x + (size_t)i * n;
    const double* xj = xi + n;
    double si[3] = {0, 0, 0}, sj[3] = {0, 0, 0};
    for (int a = lane; a < natoms; a += 32)
      for (int c = 0; c < 3; ++c) {
        si[c] += xi[3 * a + c];
        sj[c] += xj[3 * a + c];
      }
    for (int c = 0; c < 3; ++c) {
      si[c] = warp_sum(si[c]) / natoms;
      sj[c] = warp_sum(sj[c]) / natoms;
    }
    double q = 0.0;
    for (int a = lane; a < natoms; a += 32)
      for (int c = 0; c < 3; ++c) {
        const double d = (xj[3 * a + c] - sj[c]) - (xi[3 * a + c] - si[c]);
        q = fma(d, d, q);
      }
    q = warp_sum(q);
    if (lane == 0) pl[i + 1] = sqrt(q);
  }
  __syncthreads();
  if (tid == 0) {
    pl[0] = 0.0;
    for (int i = 1; i < nimg; ++i) pl[i] = pl[i - 1] + pl[i];
  }
  __syncthreads();
  if (plen_out)
    for (int i = tid; i < nimg; i += blockDim.x) plen_out[i] = pl[i];
  const double total = pl[nimg - 1];
  const double node_dist = total / (nimg - 1);
  for (int i = first + wid; i < first + nloc; i += nw) {
    double* o = xout + (size_t)(i - first) * n;
    int j = -1;       // source segment; -1: copy image `src`
    int src = i;
    if (!(total < 1e-8) && i > 0 && i < nimg - 1) {
      const double dist = i * node_dist;
      src = nimg - 1;  // "not found" safeguard of the reference
      for (int j0 = 0; j0 < nimg - 1 && j < 0; j0 += 32) {
        const int jj = j0 + lane;
        const bool hit = jj < nimg - 1 && pl[jj] <= dist && dist <= pl[jj + 1];
        const unsigned m = __ballot_sync(MOP_FULL_MASK, hit);
        if (m) j = j0 + __ffs(m) - 1;
      }
      if (j >= 0) {
        const double dt = (dist - pl[j]) / (pl[j + 1] - pl[j]);
        const double* xa = x + (size_t)j * n;
        const double* xb = xa + n;
        for (int e = lane; e < n; e += 32) o[e] = xa[e] + (xb[e] - xa[e]) * dt;
      }
    }
    if (j < 0) {
      const double* xs = x + (size_t)src * n;
      for (int e = lane; e < n; e += 32) o[e] = xs[e];
    }
  }
}

}  // namespace mop

extern "C" int mop_bneb_force(int nimg, int first, int nloc, int n, const double* x_halo,
                              const double* E_halo, const double* g, double* force, double* tau,
                              void* stream) {
  MOP_REQUIRE(nimg >= 2 && nloc >= 0 && first >= 0 && first + nloc <= nimg && n > 0 && n % 3 == 0,
              "mop_bneb_force: bad image range or n");
  MOP_REQUIRE(x_halo && E_halo && g && force && tau, "mop_bneb_force: null pointer");
  if (nloc == 0) return MOP_OK;
  const size_t smem = sizeof(double) * (size_t)n;
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_bneb_force, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_bneb_force<<<nloc, mop::NEB_THREADS, smem, (cudaStream_t)stream>>>(nimg, first, n, x_halo, E_halo, g,
                                                                          force, tau);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_neb_ayala(int nimg, int first, int nloc, int n, const double* x_halo,
                             const double* E_halo, const double* g_halo, const double* tau, double* H,
                             double* gamma_out, void* stream) {
  MOP_REQUIRE(nimg >= 2 && nloc >= 0 && first >= 0 && first + nloc <= nimg && n > 0,
              "mop_neb_ayala: bad image range or n");
  MOP_REQUIRE(x_halo && E_halo && g_halo && tau && H, "mop_neb_ayala: null pointer");
  if (nloc == 0) return MOP_OK;
  mop::k_neb_ayala<<<nloc, 256, 0, (cudaStream_t)stream>>>(nimg, first, n, x_halo, E_halo, g_halo, tau, H,
                                                          gamma_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_neb_limit_tr(int nimg, int first, int nloc, int n, int fix_init_edge, int fix_end_edge,
                                int apply_step_limit, const double* x_halo, const double* g, double* delta,
                                void* stream) {
  MOP_REQUIRE(nimg >= 2 && nloc >= 0 && first >= 0 && first + nloc <= nimg && n > 0,
              "mop_neb_limit_tr: bad image range or n");
  MOP_REQUIRE(x_halo && g && delta, "mop_neb_limit_tr: null pointer");
  if (nloc == 0) return MOP_OK;
  mop::k_neb_limit_tr<<<nloc, mop::NEB_THREADS, 0, (cudaStream_t)stream>>>(nimg, first, n, fix_init_edge,
                                                                         fix_end_edge, apply_step_limit, x_halo, g,
                                                                         delta);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_neb_fire_blend(int nloc, int natoms, double a, const double* force, const double* velocity,
                                  const double* prev_velocity, double* vneb_out, double* power_accum, void* stream) {
  MOP_REQUIRE(nloc >= 0 && natoms > 0 && force && velocity && vneb_out, "mop_neb_fire_blend: bad arguments");
  if (nloc == 0) return MOP_OK;
  const int tot = nloc * natoms;
  mop::k_neb_fire_blend<<<(tot + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tot, a, force, velocity, prev_velocity,
                                                                           vneb_out, power_accum);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_neb_fire_advance(int nloc, int n, double dt, int reset, const double* vneb, const double* force,
                                    const double* prev_velocity, double* velocity_out, double* delta_out, void* stream) {
  MOP_REQUIRE(nloc >= 0 && n > 0 && vneb && force && velocity_out && delta_out, "mop_neb_fire_advance: bad arguments");
  if (nloc == 0) return MOP_OK;
  const size_t tot = (size_t)nloc * n;
  const int grid = (int)((tot + 255) / 256 < 1184 ? (tot + 255) / 256 : 1184);
  mop::k_neb_fire_advance<<<grid, 256, 0, (cudaStream_t)stream>>>(tot, dt, reset, vneb, force, prev_velocity,
                                                                 velocity_out, delta_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_neb_redistribute(int nimg, int natoms, int first, int nloc, const double* x_chain, double* x_out,
                                    double* path_length_out, void* stream) {
  MOP_REQUIRE(nimg >= 2 && natoms > 0 && nloc >= 0 && first >= 0 && first + nloc <= nimg,
              "mop_neb_redistribute: bad image range or natoms");
  MOP_REQUIRE(nimg <= 16384, "mop_neb_redistribute: at most 16384 images");
  if (nloc == 0 && !path_length_out) return MOP_OK;   // (a rank without images has no output buffer)
  MOP_REQUIRE(x_chain && (nloc == 0 || (x_out && x_chain != x_out)), "mop_neb_redistribute: x_chain and a distinct x_out required");
  mop::k_neb_redistribute<<<1, 256, sizeof(double) * (size_t)nimg, (cudaStream_t)stream>>>(nimg, natoms, first, nloc, x_chain,
                                                                                          x_out, path_length_out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
