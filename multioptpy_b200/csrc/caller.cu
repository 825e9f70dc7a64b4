// Caller-side pieces of the optimizer step: the composite outer trust
// radius of Optimizer/trust_radius.py:120-206, evaluated for a whole batch, with its
// quadratic model  Ce = g_prev.m_prev + 1/2 m_prev^T (H + H_bias) m_prev  streamed from HBM
// (one CTA per structure, coalesced row reads).  The norm clamp and geometry update of
// optimizer.py:792-798 live in rfo_step.cu (mop_clamp_and_move).
#include "common.cuh"

namespace mop {

// per-structure state of TrustRadius: [MOP_TR_STATE] doubles
//  0 iteration_count, 1 n_ratios (<=5 kept), 2..6 energy_ratios (oldest first),
//  7 n_changes (<=3 kept), 8..10 energy_changes (oldest first)
__global__ void __launch_bounds__(256)
k_outer_trust_radius(int n, const double* __restrict__ H_all, const double* __restrict__ Hb_all,
                     const double* __restrict__ pre_Bg, const double* __restrict__ pre_move,
                     const double* __restrict__ Be, const double* __restrict__ pre_Be,
                     double* __restrict__ trust, double* __restrict__ state_all, double tmin, double tmax) {
  extern __shared__ double sm[];
  __shared__ double scratch[40];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  double* st = state_all + (size_t)b * MOP_TR_STATE;
  if (st[0] == 0.0) {  // trust_radius.py:147-149: first call only counts
    if (tid == 0) st[0] = 1.0;
    return;
  }
  double* m = sm;  // n
  for (int i = tid; i < n; i += blockDim.x) m[i] = pre_move[(size_t)b * n + i];
  __syncthreads();
  const double* H = H_all + (size_t)b * n * n;
  const double* Hb = Hb_all ? Hb_all + (size_t)b * n * n : nullptr;
  double part = 0.0;
  for (int i = w; i < n; i += nw) {
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) {
      double a = H[(size_t)i * n + j];
      if (Hb) a += Hb[(size_t)i * n + j];
      acc = fma(a, m[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) part += m[i] * (pre_Bg[(size_t)b * n + i] + 0.5 * acc);
  }
  double Ce = block_sum(part, scratch);
  double nrm2 = 0.0;
  for (int i = tid; i < n; i += blockDim.x) nrm2 = fma(m[i], m[i], nrm2);
  const double mnorm = sqrt(block_sum(nrm2, scratch));
  if (tid != 0) return;
  const double eps = 1e-8;
  if (fabs(Ce) < eps) {  // :160-163
    Ce += sgn(Ce) * eps;
    if (fabs(Ce) < eps) Ce = eps;
  }
  const double dE = pre_Be[b] - Be[b];
  const double r = dE / Ce;
  // histories (:169-171); the reference keeps them all, only the last 5 / 3 are ever read
  int nr = (int)st[1];
  if (nr >= 5) {
    for (int i = 0; i < 4; ++i) st[2 + i] = st[3 + i];
    nr = 4;
  }
  st[2 + nr] = r;
  st[1] = nr + 1;
  int nc = (int)st[7];
  if (nc >= 3) {
    st[8] = st[9];
    st[9] = st[10];
    nc = 2;
  }
  st[8 + nc] = dE;
  st[7] = nc + 1;
  nr += 1;
  nc += 1;
  // adaptive factor (:79-103): 2 exp(-var(last <= 5 ratios)), x0.8 when approaching convergence
  double var = 0.0;
  if (nr > 1) {
    double mean = 0.0;
    for (int i = 0; i < nr; ++i) mean += st[2 + i];
    mean /= nr;
    for (int i = 0; i < nr; ++i) var += (st[2 + i] - mean) * (st[2 + i] - mean);
    var /= nr;
  }
  double f = 2.0 * exp(-var);
  if (nc >= 2) {  // is_approaching_convergence (:105-117)
    bool all_small = true;
    double mean = 0.0;
    for (int i = 0; i < nc; ++i) {
      const double a = fabs(st[8 + i]);
      all_small = all_small && (a < 0.01);
      mean += a;
    }
    if (all_small && mean / nc < 0.005) f *= 0.8;
  }
  f = fmax(1.1, fmin(f, 3.0));
  double tr = trust[b];
  if (r <= 0.25 || r >= 1.75) tr /= f;
  else if (r >= 0.75 && r <= 1.25) {
    if (fabs(mnorm - tr) < eps) tr *= sqrt(f);
  }
  st[0] += 1.0;
  trust[b] = fmin(fmax(tr, tmin), tmax);
}

}  // namespace mop

extern "C" int mop_outer_trust_radius(int B, int n, const double* H, const double* Hbias,
                                      const double* pre_Bg, const double* pre_move, const double* Be,
                                      const double* pre_Be, double* trust, double* state,
                                      double trust_min, double trust_max, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_outer_trust_radius: B >= 0 and n > 0 required");
  MOP_REQUIRE(H && pre_Bg && pre_move && Be && pre_Be && trust && state,
              "mop_outer_trust_radius: H, pre_Bg, pre_move, Be, pre_Be, trust, state must be device pointers");
  if (B == 0) return MOP_OK;
  const size_t smem = sizeof(double) * (size_t)n;
  if (smem > 200 * 1024) {
    mop_set_error("mop_outer_trust_radius: n = %d too large", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_outer_trust_radius, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_outer_trust_radius<<<B, 256, smem, (cudaStream_t)stream>>>(n, H, Hbias, pre_Bg, pre_move, Be, pre_Be,
                                                                  trust, state, trust_min, trust_max);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
