// AFIR bias potential: energy, gradient and Hessian in closed form.
//
// Reference: AFIRPotential.calc_energy (Potential/AFIR_potential.py:18-55) gives
//   E = alpha * A / B,  A = sum_p w_p r_p,  B = sum_p w_p,  w_p = ((R_i + R_j) / r_p)^6
// over all fragment-1 x fragment-2 atom pairs p = (i, j); gradient and Hessian come from
// torch.func.jacrev / hessian in Potential/potential.py:130-135.  Here they are analytic:
//   u_p = a'_p - Ebar w'_p  (a = w r, Ebar = A/B),  U = sum u_p grad r_p,  W = sum w'_p grad r_p
//   grad E = alpha U / B
//   H = alpha [ sum_p (u_p/B) hess r_p + sum_p ((a''_p - Ebar w''_p)/B) grad r_p grad r_p^T
//               - (U W^T + W U^T)/B^2 ]
// The covalent radii reach the reference as FLOAT32 tensors and are added in float32
// (SURVEY H1); radii_f32 carries exactly those values and the sum is formed in float.
// One CTA per structure; fragment pairs are tiled over the threads, per-atom accumulation
// in shared memory in a fixed order (deterministic).
#include "common.cuh"

namespace mop {

constexpr int AFIR_THREADS = 256;

__global__ void __launch_bounds__(AFIR_THREADS)
k_afir(int N, const double* __restrict__ xyz_all, int n1, const int* __restrict__ frag1, int n2,
       const int* __restrict__ frag2, const float* __restrict__ rad_f32, const double* __restrict__ gamma_all,
       double* __restrict__ E_all, double* __restrict__ grad_all, double* __restrict__ H_all) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = 3 * N;
  double* xyz = sm;          // 3N
  double* U = xyz + n;       // 3N
  double* W = U + n;         // 3N
  double* scratch = W + n;   // 40
  for (int i = tid; i < n; i += AFIR_THREADS) {
    xyz[i] = xyz_all[(size_t)b * n + i];
    U[i] = 0.0;
    W[i] = 0.0;
  }
  __syncthreads();
  // alpha (AFIR_potential.py:33-36); gamma in kJ/mol
  const double gam = gamma_all[b];
  const double hartree2kjmol = 2625.5, bohr2ang = 0.52917721067;
  const double R0 = 3.8164 / bohr2ang, EPS = 1.0061 / hartree2kjmol;
  double alpha = 0.0;
  if (gam > 0.0 || gam < 0.0) {
    const double gh = gam / hartree2kjmol;
    alpha = gh / ((pow(2.0, -1.0 / 6.0) - pow(1.0 + sqrt(1.0 + fabs(gh) / EPS), -1.0 / 6.0)) * R0);
  }
  const int np_ = n1 * n2;
  // pass 1: A, B
  double pa = 0.0, pb = 0.0;
  for (int p = tid; p < np_; p += AFIR_THREADS) {
    const int i = frag1[p / n2], j = frag2[p % n2];
    const double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1], dz = xyz[3 * i + 2] - xyz[3 * j + 2];
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const double Rs = (double)(rad_f32[i] + rad_f32[j]);  // float32 addition, then promoted
    const double w = pow(Rs / r, 6.0);
    pa += w * r;
    pb += w;
  }
  const double A = block_sum(pa, scratch);
  const double Bs = block_sum(pb, scratch);
  const double Ebar = A / Bs;
  if (tid == 0 && E_all) E_all[b] = alpha * Ebar;
  // zero the Hessian
  double* H = H_all ? H_all + (size_t)b * n * n : nullptr;
  if (H)
    for (size_t e = tid; e < (size_t)n * n; e += AFIR_THREADS) H[e] = 0.0;
  __syncthreads();
  // pass 2: U, W and the pair-local Hessian terms.  Atom i of fragment 1 is owned by one
  // thread at a time (loop over its partners j), so shared/global accumulation is race free
  // for the i-side; the j-side is accumulated in a second sweep with roles swapped.
  for (int side = 0; side < 2; ++side) {
    const int na = side == 0 ? n1 : n2, nbp = side == 0 ? n2 : n1;
    const int* fa = side == 0 ? frag1 : frag2;
    const int* fb = side == 0 ? frag2 : frag1;
    for (int ia = tid; ia < na; ia += AFIR_THREADS) {
      const int i = fa[ia];
      // duplicates of atom i inside the same fragment would race: handled by ownership below
      bool first = true;
      for (int q = 0; q < ia; ++q)
        if (fa[q] == i) first = false;
      if (!first) continue;
      double u3[3] = {0, 0, 0}, w3[3] = {0, 0, 0};
      double hd[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int ia2 = ia; ia2 < na; ++ia2) {
        if (fa[ia2] != i) continue;
        for (int jb = 0; jb < nbp; ++jb) {
          const int j = fb[jb];
          const double dx = xyz[3 * i] - xyz[3 * j], dy = xyz[3 * i + 1] - xyz[3 * j + 1],
                       dz = xyz[3 * i + 2] - xyz[3 * j + 2];
          const double r = sqrt(dx * dx + dy * dy + dz * dz);
          const double Rs = (double)(rad_f32[i] + rad_f32[j]);
          const double w = pow(Rs / r, 6.0);
          const double wp = -6.0 * w / r, ap = -5.0 * w;              // w', a'
          const double wpp = 42.0 * w / (r * r), app = 30.0 * w / r;  // w'', a''
          const double up = ap - Ebar * wp;
          const double e[3] = {dx / r, dy / r, dz / r};               // d r / d x_i
          const double c1 = alpha * (up / Bs) / r;                    // hess r = (I - e e^T)/r
          const double c2 = alpha * (app - Ebar * wpp) / Bs;
          for (int c = 0; c < 3; ++c) {
            u3[c] += up * e[c];
            w3[c] += wp * e[c];
          }
          double blk[9];
          for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d) blk[3 * c + d] = c1 * ((c == d ? 1.0 : 0.0) - e[c] * e[d]) + c2 * e[c] * e[d];
          for (int q = 0; q < 9; ++q) hd[q] += blk[q];
          if (H && side == 0 && i != j) {  // off-diagonal blocks (i, j) and (j, i) = -blk, written once
            for (int c = 0; c < 3; ++c)
              for (int d = 0; d < 3; ++d) {
                atomicAdd(&H[(size_t)(3 * i + c) * n + 3 * j + d], -blk[3 * c + d]);
                atomicAdd(&H[(size_t)(3 * j + d) * n + 3 * i + c], -blk[3 * c + d]);
              }
          }
        }
      }
      for (int c = 0; c < 3; ++c) {
        atomicAdd(&U[3 * i + c], u3[c]);   // an atom may sit in both fragments
        atomicAdd(&W[3 * i + c], w3[c]);
      }
      if (H)
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d) atomicAdd(&H[(size_t)(3 * i + c) * n + 3 * i + d], hd[3 * c + d]);
    }
    __syncthreads();
  }
  // gradient and the rank-2 term
  if (grad_all)
    for (int i = tid; i < n; i += AFIR_THREADS) grad_all[(size_t)b * n + i] = alpha * U[i] / Bs;
  if (H) {
    const double f = alpha / (Bs * Bs);
    for (size_t e = tid; e < (size_t)n * n; e += AFIR_THREADS) {
      const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
      const double ur = U[r], wr = W[r], uc = U[c], wc = W[c];
      if ((ur != 0.0 || wr != 0.0) && (uc != 0.0 || wc != 0.0)) H[e] -= f * (ur * wc + wr * uc);
    }
  }
}

}  // namespace mop

// AFIR energy / gradient / Hessian.  frag1/frag2: 0-based atom indices (device int32),
// radii_f32: covalent radii [natoms] rounded to float32 (Bohr), gamma [B] in kJ/mol.
extern "C" int mop_afir(int B, int natoms, const double* xyz, int n1, const int32_t* frag1, int n2,
                        const int32_t* frag2, const float* radii_f32, const double* gamma, double* E,
                        double* grad, double* H, void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 0 && n1 > 0 && n2 > 0, "mop_afir: B >= 0, natoms, n1, n2 > 0 required");
  MOP_REQUIRE(xyz && frag1 && frag2 && radii_f32 && gamma, "mop_afir: xyz, frag1, frag2, radii_f32, gamma required");
  if (B == 0) return MOP_OK;
  const size_t smem = sizeof(double) * (9 * (size_t)natoms + 40);
  if (smem > 200 * 1024) {
    mop_set_error("mop_afir: natoms = %d too large", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_afir, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_afir<<<B, mop::AFIR_THREADS, smem, (cudaStream_t)stream>>>(natoms, xyz, n1, frag1, n2, frag2,
                                                                 radii_f32, gamma, E, grad, H);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
