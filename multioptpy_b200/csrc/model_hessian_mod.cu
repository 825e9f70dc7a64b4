// Eigenvalue-based modifiers of a model Hessian (ModelHessian/approx_hessian.py:95-110), applied after the base model:
//   "ts"   TransitionStateHessian.create_ts_hessian (ModelHessian/tshess.py:14-40): reflect the Hessian through the
//          lowest non-zero mode, sym((1 - 2 v v^T) H), unless a negative eigenvalue exists already;
//   "clip" eigenvalue smoothing (approx_hessian.py:103-126): V diag(smooth(lambda)) V^T with
//          smooth(x) = sign(x) (2 - |x|^-0.1) for |x| >= 1.
// Both take the eigendecomposition from mop_eigh (rows of `evecs` = eigenvectors, ascending).  One CTA per structure.
#include "common.cuh"

namespace mop {

__global__ void __launch_bounds__(256) k_ts_modify(int n, const double* __restrict__ H_all,
                                                   const double* __restrict__ evals_all,
                                                   const double* __restrict__ evecs_all, double* __restrict__ out_all,
                                                   int32_t* __restrict__ modified) {
  extern __shared__ double sm[];
  double* v = sm;      // [n] target eigenvector
  double* u = v + n;   // [n] H^T v
  __shared__ int s_neg, s_count;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const double* H = H_all + (size_t)b * n * n;
  const double* ev = evals_all + (size_t)b * n;
  double* out = out_all + (size_t)b * n * n;
  if (tid == 0) {
    int neg = 0, count = 0;
    for (int i = 0; i < n; ++i) neg |= ev[i] < -1e-8;          // tshess.py:19
    for (int i = 0; i < n; ++i) {                               // :23-28: leading (numerically) zero modes
      if (fabs(ev[i]) < 1e-8) ++count;
      else break;
    }
    s_neg = neg;
    s_count = count;
  }
  __syncthreads();
  const bool keep = s_neg || s_count >= n;  // (all modes zero: the reference indexes out of range; left unchanged here)
  if (tid == 0 && modified) modified[b] = keep ? 0 : 1;
  if (keep) {
    for (size_t e = tid; e < (size_t)n * n; e += blockDim.x) out[e] = H[e];
    return;
  }
  const double* vec = evecs_all + ((size_t)b * n + s_count) * n;
  for (int i = tid; i < n; i += blockDim.x) {
    v[i] = vec[i];
    u[i] = 0.0;
  }
  __syncthreads();
  // u_j = sum_k v_k H_kj: warps over row blocks, lanes over columns (coalesced); per-warp partials combined through atomics
  // would not be deterministic - every thread owns columns instead
  for (int j = tid; j < n; j += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc = fma(v[k], H[(size_t)k * n + j], acc);
    u[j] = acc;
  }
  __syncthreads();
  (void)lane; (void)w; (void)nw;
  // ts = 1/2 (M + M^T), M = H - 2 v u^T  (tshess.py:33-38)
  for (size_t e = tid; e < (size_t)n * n; e += blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    const double mij = H[e] - 2.0 * v[i] * u[j];
    const double mji = H[(size_t)j * n + i] - 2.0 * v[j] * u[i];
    out[e] = 0.5 * (mij + mji);
  }
}

__device__ __forceinline__ double smooth_eigval(double x) {  // approx_hessian.py:119-126, alpha = 0.1
  const double a = fabs(x);
  if (a >= 1.0) return sgn(x) * (2.0 - 1.0 / pow(a, 0.1));
  return x;
}

// out = V diag(smooth(lambda)) V^T; evecs rows = eigenvectors.  32 x 32 output tiles, k-loop through shared memory.
__global__ void __launch_bounds__(256) k_clip_recompose(int n, const double* __restrict__ evals_all,
                                                        const double* __restrict__ evecs_all,
                                                        double* __restrict__ out_all) {
  __shared__ double A[32][33], Bt[32][33], lam[32];
  const int b = blockIdx.z, ti = blockIdx.y * 32, tj = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const double* V = evecs_all + (size_t)b * n * n;
  const double* ev = evals_all + (size_t)b * n;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < n; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {  // A[k][i] = V[k0 + k][ti + i], Bt[k][j] = V[k0 + k][tj + j]
      const int k = k0 + r;
      A[r][tx] = (k < n && ti + tx < n) ? V[(size_t)k * n + ti + tx] : 0.0;
      Bt[r][tx] = (k < n && tj + tx < n) ? V[(size_t)k * n + tj + tx] : 0.0;
    }
    if (threadIdx.x < 32) lam[threadIdx.x] = k0 + threadIdx.x < n ? smooth_eigval(ev[k0 + threadIdx.x]) : 0.0;
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double bj = Bt[k][tx] * lam[k];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(A[k][ty + 8 * q], bj, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ti + ty + 8 * q, j = tj + tx;
    if (i < n && j < n) out_all[(size_t)b * n * n + (size_t)i * n + j] = acc[q];
  }
}

}  // namespace mop

// TransitionStateHessian.create_ts_hessian for a batch: H, out [B][n][n] (out may not alias H), evals [B][n] and evecs
// [B][n][n] from mop_eigh(H); modified (optional) [B]: 1 where the reflection was applied.
extern "C" int mop_hessian_ts_modify(int B, int n, const double* H, const double* evals, const double* evecs, double* out,
                                     int32_t* modified, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && H && evals && evecs && out && out != H, "mop_hessian_ts_modify: bad arguments");
  if (B == 0) return MOP_OK;
  mop::k_ts_modify<<<B, 256, sizeof(double) * 2 * n, (cudaStream_t)stream>>>(n, H, evals, evecs, out, modified);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// The "clip" modifier: out = V diag(smooth(lambda)) V^T from mop_eigh's (evals, evecs).
extern "C" int mop_hessian_clip_eigvals(int B, int n, const double* evals, const double* evecs, double* out,
                                        void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && evals && evecs && out && out != evecs, "mop_hessian_clip_eigvals: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid((n + 31) / 32, (n + 31) / 32, B);
  mop::k_clip_recompose<<<grid, 256, 0, (cudaStream_t)stream>>>(n, evals, evecs, out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// ---- effective Hessian for fixed atoms (optimization.py:1325-1343,1358-1362) --------------------------------------------
// H -= H[:, f] pinv(H[f, f] + 1e-10 I) H[f, :] for the coordinate set f of the fixed atoms (HessianManager.
// calc_eff_hess_for_fix_atoms_and_set_hess applies it to the bias Hessian and to the model Hessian).  The pseudo-inverse
// comes from the eigendecomposition of the small symmetric block (mop_eigh on the gathered blocks): numpy.linalg.pinv
// drops singular values <= 1e-15 max|lambda|.
namespace mop {

__global__ void __launch_bounds__(256) k_gather_fix_block(int n, int m, const int* __restrict__ fix,
                                                          const double* __restrict__ H_all, double* __restrict__ F_all) {
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int p = e / m, q = e - p * m;
    F_all[(size_t)b * m * m + e] = H_all[(size_t)b * n * n + (size_t)fix[p] * n + fix[q]] + (p == q ? 1e-10 : 0.0);
  }
}

// evals / evecs (rows = eigenvectors) of the gathered blocks.  Shared memory: P [m][m] | W [m][n] | C [n][m].
__global__ void __launch_bounds__(256) k_schur_fix(int n, int m, const int* __restrict__ fix,
                                                   const double* __restrict__ evals_all,
                                                   const double* __restrict__ evecs_all, double* __restrict__ H_all) {
  extern __shared__ double sm[];
  double* P = sm;
  double* W = P + (size_t)m * m;
  double* C = W + (size_t)m * n;
  const int b = blockIdx.x, tid = threadIdx.x;
  double* H = H_all + (size_t)b * n * n;
  const double* ev = evals_all + (size_t)b * m;
  const double* V = evecs_all + (size_t)b * m * m;
  double mx = 0.0;
  for (int k = 0; k < m; ++k) mx = fmax(mx, fabs(ev[k]));
  const double cut = 1e-15 * mx;
  for (int e = tid; e < m * m; e += blockDim.x) {
    const int p = e / m, q = e - p * m;
    double acc = 0.0;
    for (int k = 0; k < m; ++k)
      if (fabs(ev[k]) > cut) acc += V[(size_t)k * m + p] * V[(size_t)k * m + q] / ev[k];
    P[e] = acc;
  }
  for (int e = tid; e < n * m; e += blockDim.x) {
    const int i = e / m, a = e - i * m;
    C[e] = H[(size_t)i * n + fix[a]];          // H[:, f]
  }
  __syncthreads();
  for (int e = tid; e < m * n; e += blockDim.x) {  // W = P H[f, :]
    const int a = e / n, j = e - a * n;
    double acc = 0.0;
    for (int c = 0; c < m; ++c) acc += P[a * m + c] * H[(size_t)fix[c] * n + j];
    W[e] = acc;
  }
  __syncthreads();
  for (size_t e = tid; e < (size_t)n * n; e += blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    double acc = 0.0;
    for (int a = 0; a < m; ++a) acc += C[i * m + a] * W[a * n + j];
    H[e] -= acc;
  }
}

}  // namespace mop

extern "C" int mop_fix_atoms_gather(int B, int n, int m, const int32_t* fix_coords, const double* H, double* blocks,
                                    void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && m > 0 && m <= n && fix_coords && H && blocks, "mop_fix_atoms_gather: bad arguments");
  if (B == 0) return MOP_OK;
  mop::k_gather_fix_block<<<B, 256, 0, (cudaStream_t)stream>>>(n, m, fix_coords, H, blocks);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

// H [B][n][n] in place; fix_coords [m] device int32 (3 (a - 1) + c of every fixed atom); evals [B][m], evecs [B][m][m]:
// mop_eigh of the blocks from mop_fix_atoms_gather.
extern "C" int mop_fix_atoms_schur(int B, int n, int m, const int32_t* fix_coords, const double* evals,
                                   const double* evecs, double* H, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && m > 0 && m <= n && fix_coords && evals && evecs && H, "mop_fix_atoms_schur: bad arguments");
  if (B == 0) return MOP_OK;
  const size_t smem = sizeof(double) * ((size_t)m * m + 2 * (size_t)m * n);
  if (smem > 200 * 1024) {
    mop_set_error("mop_fix_atoms_schur: %d fixed coordinates of %d need %zu bytes of shared memory", m, n, smem);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_schur_fix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_schur_fix<<<B, 256, smem, (cudaStream_t)stream>>>(n, m, fix_coords, evals, evecs, H);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// "sr": ShortRangeCorrectionHessian (ModelHessian/shortrange.py:9-346) - second derivatives of the short-range Coulomb
// kernel (1 - erf(w r)) / r between NON-bonded atom pairs (BondConnectivity, 1.1 x the covalent radii) with
// electronegativity charges, TR/ROT-projected, added to the base Hessian and symmetrised.  One CTA per structure, one
// thread per 3 x 3 block: an off-diagonal block is minus the pair block, a diagonal block the sum of the atom's pair
// blocks in ascending partner order (the order the reference's double loop adds them in).
#include "connectivity.cuh"

namespace mop {

__device__ __forceinline__ void sr_pair_block(const double* xi, const double* xj, double ri, double rj, double qf,
                                              double omega, double cutoff, double* blk) {
  for (int e = 0; e < 9; ++e) blk[e] = 0.0;
  const double dist = np_dist(xj, xi);
  if (dist <= __dmul_rn(__dadd_rn(rj, ri), 1.1)) return;   // bonded (bond_connectivity.py:34-40)
  double rv[3] = {xj[0] - xi[0], xj[1] - xi[1], xj[2] - xi[2]};
  const double r = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  if (r > cutoff) return;
  const double PI = 3.141592653589793;
  double d1, d2;
  if (r < 1e-10) {
    d1 = -2.0 * omega * omega * omega / (3.0 * sqrt(PI));
    d2 = 0.0;
  } else {
    const double ef = erf(omega * r), ex = exp(-(omega * r) * (omega * r));
    d1 = 2.0 * omega * ex / (sqrt(PI) * r) + (ef - 1.0) / (r * r);
    const double xf = ex / sqrt(PI);
    d2 = 2.0 * (2.0 * ef - 1.0) / (r * r * r) + 4.0 * omega * xf / (r * r) + 2.0 * (omega * omega * omega) * xf;
  }
  const double u[3] = {rv[0] / r, rv[1] / r, rv[2] / r};
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) {
      const double o = u[a] * u[c];
      blk[3 * a + c] = qf * (d2 * o + d1 / r * ((a == c ? 1.0 : 0.0) - o));
    }
}

__global__ void __launch_bounds__(256)
k_sr_correction(int N, const double* __restrict__ xyz_all, const double* __restrict__ rad_all, int rad_stride,
                const double* __restrict__ chg_all, int chg_stride, double omega, double cfac, double cutoff,
                double* __restrict__ C_all) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, n = 3 * N;
  double* xyz = sm;
  double* rad = xyz + 3 * N;
  double* chg = rad + N;
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    rad[i] = rad_all[(size_t)b * rad_stride + i];
    chg[i] = chg_all[(size_t)b * chg_stride + i];
  }
  __syncthreads();
  double* C = C_all + (size_t)b * n * n;
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    double blk[9], acc[9];
    if (i != j) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;   // the reference evaluates the pair (lo, hi)
      sr_pair_block(xyz + 3 * lo, xyz + 3 * hi, rad[lo], rad[hi], chg[lo] * chg[hi] * cfac, omega, cutoff, blk);
      for (int q = 0; q < 9; ++q) acc[q] = 0.0 - blk[q];
    } else {
      for (int q = 0; q < 9; ++q) acc[q] = 0.0;
      for (int k = 0; k < N; ++k) {
        if (k == i) continue;
        const int lo = i < k ? i : k, hi = i < k ? k : i;
        sr_pair_block(xyz + 3 * lo, xyz + 3 * hi, rad[lo], rad[hi], chg[lo] * chg[hi] * cfac, omega, cutoff, blk);
        for (int q = 0; q < 9; ++q) acc[q] += blk[q];
      }
    }
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) C[(size_t)(3 * i + a) * n + 3 * j + c] = acc[3 * a + c];
  }
}

// out = 1/2 ((H + C) + (H + C)^T)
__global__ void __launch_bounds__(256) k_add_sym(int n, const double* __restrict__ H, const double* __restrict__ C,
                                                 double* __restrict__ out) {
  const size_t b = blockIdx.y, nn = (size_t)n * n;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / n, j = e - i * n, t = j * n + i;
    out[b * nn + e] = 0.5 * ((H[b * nn + e] + C[b * nn + e]) + (H[b * nn + t] + C[b * nn + t]));
  }
}

}  // namespace mop

int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x, const double* g,
                             double* Hp_out, double* gp_out, int32_t* status, int grad_rule, cudaStream_t stream);

// ShortRangeCorrectionHessian.main for a batch: out = sym(H + P^T C P).  radii: covalent radii (Bohr) [natoms] (stride 0) or
// [B][natoms]; charges: 0.2 (mean electronegativity - electronegativity) per atom, same strides; work: 2 B n^2 doubles.
extern "C" size_t mop_hessian_sr_workspace_bytes(int B, int natoms) {
  return B > 0 && natoms > 0 ? sizeof(double) * 2 * (size_t)B * 9 * natoms * natoms : 0;
}

extern "C" int mop_hessian_sr_correction(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                                         const double* charges, int charges_stride, double omega, double cx_sr,
                                         double scaling_factor, const double* H, double* out, void* work,
                                         size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 0, "mop_hessian_sr_correction: bad arguments");
  MOP_REQUIRE((radii_stride == 0 || radii_stride == natoms) && (charges_stride == 0 || charges_stride == natoms),
              "mop_hessian_sr_correction: strides must be 0 or natoms");
  if (B == 0) return MOP_OK;   // (an empty batch has no buffers)
  MOP_REQUIRE(xyz && radii && charges && H && out, "mop_hessian_sr_correction: null pointer");
  if (!work || work_bytes < mop_hessian_sr_workspace_bytes(B, natoms)) {
    mop_set_error("mop_hessian_sr_correction: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int n = 3 * natoms;
  double* C = (double*)work;
  double* Cp = C + (size_t)B * n * n;
  const size_t smem = sizeof(double) * 5 * (size_t)natoms;
  mop::k_sr_correction<<<B, 256, smem, stream>>>(natoms, xyz, radii, radii_stride, charges, charges_stride, omega,
                                                 cx_sr * scaling_factor, 15.0, C);
  MOP_CHECK_CUDA(cudaGetLastError());
  int rc = mop_launch_project_trrot(B, n, C, nullptr, xyz, nullptr, Cp, nullptr, nullptr, 0, stream);
  if (rc != MOP_OK) return rc;
  dim3 grid(64, B);
  mop::k_add_sym<<<grid, 256, 0, stream>>>(n, H, Cp, out);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
