// RS-I-RFO step from a finished eigendecomposition, generic path.
//
// One CTA per structure.  Inputs: ascending eigenvalues and eigenvectors (row k =
// vector k) of the TR/ROT-projected Hessian Hp, the projected gradient gp, the raw
// biased gradient Bg (norm only), Hp itself (predicted energy change) and the
// per-structure RSIRFO state.  Follows Optimizer/rsirfo.py:360-490 step by step;
// the second eigendecomposition of the image Hessian H* (rsirfo.py:423-427) is
// derived from the first one: H* = V diag(lambda') V^T with the `saddle_order`
// lowest |lambda| > 1e-10 eigenvalues sign-flipped (zeroed in NEB mode), and
// g* has the matching components negated (zeroed).
#include "rfo_core.cuh"

namespace mop {

constexpr int RFO_THREADS = 256;

__global__ void __launch_bounds__(RFO_THREADS, 1)
k_rfo_step(int n, int saddle_order, int neb_mode, double tmin, double tmax,
           const double* __restrict__ evals_all, const double* __restrict__ evecs_all,
           const double* __restrict__ gp_all, const double* __restrict__ Bg_all,
           const double* __restrict__ Be_all, double* __restrict__ state_all,
           double* __restrict__ move_all, double* __restrict__ evals_out,
           double* __restrict__ pred_all, int32_t* __restrict__ status, int only_flagged) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = RFO_THREADS >> 5;
  if (only_flagged && !(status[b] & MOP_ST_EIG_FALLBACK)) return;
  const int np = (n + 3) & ~3;
  double* lam = sm;        // spectrum of Hp
  double* gam = lam + np;  // V^T gp
  double* gp = gam + np;
  RfoArrays R = rfo_carve(gp + np, n);
  const double* V = evecs_all + (size_t)b * n * n;
  double* st = state_all + (size_t)b * MOP_RSIRFO_STATE;
  int flags = 0;

  // spectrum, projected gradient, finiteness (rsirfo.py:360-369)
  double bad = 0.0, pg = 0.0;
  for (int i = tid; i < n; i += RFO_THREADS) {
    const double l = evals_all[(size_t)b * n + i];
    lam[i] = l;
    gp[i] = gp_all[(size_t)b * n + i];
    if (!isfinite(l)) bad = 1.0;
    const double g = Bg_all[(size_t)b * n + i];
    pg = fma(g, g, pg);
  }
  const double gnorm_raw = sqrt(block_sum(pg, R.scratch));
  for (int k = wid; k < n; k += nw) {
    const double* vk = V + (size_t)k * n;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = vk[i];
      if (!isfinite(v)) bad = 1.0;
      acc = fma(v, gp[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) gam[k] = acc;
  }
  bad = block_sum(bad, R.scratch);
  const bool identity = bad > 0.0;
  if (identity) {
    flags |= MOP_ST_EIG_NONFINITE;
    for (int i = tid; i < n; i += RFO_THREADS) {
      lam[i] = 1.0;
      gam[i] = gp[i];
    }
  }
  __syncthreads();
  if (evals_out)
    for (int i = tid; i < n; i += RFO_THREADS) evals_out[(size_t)b * n + i] = lam[i];

  flags |= rfo_core(n, saddle_order, neb_mode, tmin, tmax, lam, gam, identity, gnorm_raw,
                    Be_all ? Be_all[b] : 0.0, st, R, pred_all ? pred_all + b : nullptr);

  // back-transform: step = sum_k coef[k] v_k ; the reference returns minus the step
  for (int i = tid; i < n; i += RFO_THREADS) {
    double acc = 0.0;
    if (identity) {
      acc = R.coef[i];
    } else {
      for (int k = 0; k < n; ++k) {
        const double c = R.coef[k];
        if (c != 0.0) acc = fma(V[(size_t)k * n + i], c, acc);
      }
    }
    move_all[(size_t)b * n + i] = -acc;
  }
  if (tid == 0 && status) {
    const int keep = status[b] & (MOP_ST_UPDATED | MOP_ST_UPD_SKIP_SMALL | MOP_ST_UPD_SKIP_CURV |
                                  MOP_ST_UPD_TERM_ZEROED | MOP_ST_NO_HISTORY | MOP_ST_TRROT_RANKDEF |
                                  MOP_ST_EIG_NOCONV | MOP_ST_EIG_FALLBACK);
    status[b] = keep | flags;
  }
}

__global__ void k_clamp_and_move(int n, const double* __restrict__ x, double* __restrict__ move,
                                 const double* __restrict__ trust, double* __restrict__ xnew) {
  __shared__ double scratch[40];
  const int b = blockIdx.x;
  double p = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = move[(size_t)b * n + i];
    p = fma(m, m, p);
  }
  const double nrm = sqrt(block_sum(p, scratch));
  const double tr = trust[b];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double m = move[(size_t)b * n + i];
    if (nrm > tr) {
      m = tr * m / nrm;  // optimizer.py:793
      move[(size_t)b * n + i] = m;
    }
    if (xnew) xnew[(size_t)b * n + i] = (x[(size_t)b * n + i] - m) * 0.52917721067;
  }
}

}  // namespace mop

int mop_launch_rfo_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                        const double* evals, const double* evecs, const double* gp,
                        const double* Bg, const double* Be, double* state,
                        double* move, double* evals_out, double* pred, int32_t* status,
                        int only_flagged, cudaStream_t stream) {
  if (B == 0) return MOP_OK;
  const int np = (n + 3) & ~3;
  const size_t smem = sizeof(double) * 3 * (size_t)np + mop::rfo_core_smem_bytes(n);
  if (smem > 200 * 1024) {
    mop_set_error("n = %d too large for the RFO step kernel", n);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_rfo_step, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  mop::k_rfo_step<<<B, mop::RFO_THREADS, smem, stream>>>(n, saddle_order, neb_mode, tmin, tmax,
                                                       evals, evecs, gp, Bg, Be, state, move,
                                                       evals_out, pred, status, only_flagged);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_clamp_and_move(int B, int n, const double* x, double* move,
                                  const double* trust_outer, double* x_new_ang, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_clamp_and_move: B >= 0 and n > 0 required");
  MOP_REQUIRE(move && trust_outer, "mop_clamp_and_move: move and trust_outer required");
  MOP_REQUIRE(!x_new_ang || x, "mop_clamp_and_move: x required with x_new_ang");
  if (B == 0) return MOP_OK;
  mop::k_clamp_and_move<<<B, 128, 0, (cudaStream_t)stream>>>(n, x, move, trust_outer, x_new_ang);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
