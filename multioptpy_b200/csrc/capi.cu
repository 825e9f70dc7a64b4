// Host side of the C ABI (include/mop_b200.h): argument checks, workspace
// carving and kernel sequencing.  No allocation, no synchronisation.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

// launchers implemented next to their kernels
int mop_launch_hessian_update(int B, int n, int method, int mode, int guards, double* H,
                              const double* s, const double* y, const double* x, const double* xp,
                              const double* g, const double* gp, const double* state, int state_stride,
                              double* delta_out, int32_t* status, cudaStream_t stream);
int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                             const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                             cudaStream_t stream);
size_t mop_project_scratch_bytes(int B, int n);
int mop_launch_project_trrot_split(int B, int n, const double* H, const double* Hbias, const double* x,
                                   const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                                   void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mop_hessian_update_scratch_bytes(int B, int n);
int mop_launch_hessian_update_split(int B, int n, int method, int mode, int guards, double* H, const double* s,
                                    const double* y, const double* x, const double* xp, const double* g,
                                    const double* gp, const double* state, int state_stride, double* delta_out,
                                    int32_t* status, void* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t mop_jacobi_workspace_bytes(int B, int n);
int mop_launch_eigh_jacobi_ext(int B, int n, const double* A, double* evals, double* evecs,
                               int32_t* status, const int32_t* only_flagged, void* work,
                               size_t work_bytes, double* awork_ext, cudaStream_t stream);
int mop_launch_eigh_jacobi(int B, int n, const double* A, double* evals, double* evecs,
                           int32_t* status, const int32_t* only_flagged, void* work,
                           size_t work_bytes, cudaStream_t stream);
size_t mop_tridiag_workspace_bytes(int B, int n);
int mop_tridiag_supported(int n);
size_t mop_large_workspace_bytes(int B, int n);
int mop_large_supported(int n);
int mop_launch_eigh_large(int B, int n, const double* A, double* evals, double* evecs, int32_t* status,
                          void* work, size_t work_bytes, cudaStream_t stream);
int mop_launch_eigh_large_factored(int B, int n, const double* A, double* evals, double* Zt, int32_t* status,
                                   void* work, size_t work_bytes, cudaStream_t stream);
int mop_launch_large_apply_q(int B, int n, int trans, void* work, double* x0, double* x1, double* x2, double* x3,
                             cudaStream_t stream);
int mop_launch_rfo_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                        const double* evals, const double* evecs, const double* gp,
                        const double* Bg, const double* Be, double* state,
                        double* move, double* evals_out, double* pred, int32_t* status,
                        int only_flagged, cudaStream_t stream);
int mop_launch_rsirfo_fused(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                            const double* Hp, const double* gp, const double* Bg, const double* Be,
                            double* state, double* move, double* evals_out, double* pred,
                            int32_t* status, void* work, size_t work_bytes, double* zbuf, cudaStream_t stream);

int mop_launch_project_trrot_flagged(int B, int n, const double* H, const double* Hbias, const double* x,
                                     double* Hp_out, const int32_t* flags, cudaStream_t stream);
int mop_launch_front_tridiag_blk(int B, int n, int method, const int32_t* method_per, int guards, int grad_rule,
                                 int packed, double* H, const double* Hbias,
                                 const double* x, const double* xp, const double* g, const double* gprev,
                                 const double* Bg, const double* state, int state_stride, double* gp_out,
                                 int32_t* status, double* Vh, double* dd, double* ee, double* tau, double* gq, int* flag,
                                 double* hand, cudaStream_t stream);
int mop_spectrum_step_supported(int n);
int mop_launch_spectrum_step(int B, int n, int saddle_order, int neb_mode, double tmin, double tmax,
                             const double* Vh, double* Z, double* Dm, const double* pd, const double* pe,
                             const double* ptau, const double* pgq, const int* pflag, const double* Bg,
                             const double* Be, double* state, double* move, double* evals_out, double* pred,
                             int32_t* status, cudaStream_t stream);

extern "C" size_t mop_rsirfo_spectral_workspace_bytes(int B, int n);
extern "C" int mop_rsirfo_spectral_step(int B, int n, int saddle_order, int neb_mode, double trust_min,
                                        double trust_max, const double* Hp, const double* gp,
                                        const double* Bg, const double* Be, double* state,
                                        double* move_out, double* eigvals_out, double* pred_out,
                                        int32_t* status, void* work, size_t work_bytes, void* stream_);

static thread_local char g_err[512] = "";

void mop_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int mop_version(void) { return MOP_VERSION; }
extern "C" const char* mop_last_error(void) { return g_err; }

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

namespace mop {
// Packed lower triangle <-> full symmetric matrix (row i of the triangle at i (i + 1) / 2).  only_flagged: structures
// without MOP_ST_EIG_FALLBACK in flags[b] are skipped (the robust path of the packed step).
__global__ void k_unpack_lower(int n, const double* __restrict__ P, double* __restrict__ H,
                               const int32_t* __restrict__ only_flagged) {
  const size_t b = blockIdx.y, nn = (size_t)n * n, nt = (size_t)n * (n + 1) / 2;
  if (only_flagged && !(only_flagged[b] & MOP_ST_EIG_FALLBACK)) return;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    const int hi = i > j ? i : j, lo = i > j ? j : i;
    H[b * nn + e] = P[b * nt + ((size_t)hi * (hi + 1) >> 1) + lo];
  }
}
__global__ void k_pack_lower(int n, const double* __restrict__ H, double* __restrict__ P) {
  const size_t b = blockIdx.y, nn = (size_t)n * n, nt = (size_t)n * (n + 1) / 2;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < nn; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    if (j <= i) P[b * nt + ((size_t)i * (i + 1) >> 1) + j] = H[b * nn + e];
  }
}
}  // namespace mop

extern "C" int mop_pack_lower(int B, int n, const double* H, double* packed, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && H && packed, "mop_pack_lower: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid(64, B);
  mop::k_pack_lower<<<grid, 256, 0, (cudaStream_t)stream>>>(n, H, packed);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
extern "C" int mop_unpack_lower(int B, int n, const double* packed, double* H, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && H && packed, "mop_unpack_lower: bad arguments");
  if (B == 0) return MOP_OK;
  dim3 grid(64, B);
  mop::k_unpack_lower<<<grid, 256, 0, (cudaStream_t)stream>>>(n, packed, H, nullptr);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

static int pick_algo(int algo, int n) {
  if (algo == MOP_EIGH_AUTO)
    return mop_tridiag_supported(n) ? MOP_EIGH_TRIDIAG : (mop_large_supported(n) ? MOP_EIGH_LARGE : MOP_EIGH_JACOBI);
  return algo;
}

static size_t eigh_work_bytes(int B, int n, int algo) {
  algo = pick_algo(algo, n);
  size_t jac = align256(mop_jacobi_workspace_bytes(B, n));  // also the fallback of the fast path
  if (algo == MOP_EIGH_TRIDIAG) {
    const size_t a = align256(mop_tridiag_workspace_bytes(B, n)), b = align256(mop_large_workspace_bytes(B, n));
    return jac + (a > b ? a : b);
  }
  if (algo == MOP_EIGH_LARGE) return jac + align256(mop_large_workspace_bytes(B, n));
  return jac;
}

extern "C" size_t mop_eigh_workspace_bytes(int B, int n, int algo) {
  if (B <= 0 || n <= 0) return 0;
  return eigh_work_bytes(B, n, algo);
}

static int run_eigh(int B, int n, int algo, const double* A, double* evals, double* evecs,
                    int32_t* status, void* work, size_t work_bytes, cudaStream_t stream) {
  algo = pick_algo(algo, n);
  if (work_bytes < eigh_work_bytes(B, n, algo)) {
    mop_set_error("eigh: workspace too small (%zu < %zu bytes)", work_bytes,
                  eigh_work_bytes(B, n, algo));
    return MOP_ERR_WORKSPACE;
  }
  const size_t jac = align256(mop_jacobi_workspace_bytes(B, n));
  if (algo == MOP_EIGH_JACOBI)
    return mop_launch_eigh_jacobi(B, n, A, evals, evecs, status, nullptr, work, jac, stream);
  if (algo == MOP_EIGH_TRIDIAG) {
    if (!mop_tridiag_supported(n)) {
      mop_set_error("eigh: tridiagonal path does not support n = %d", n);
      return MOP_ERR_UNSUPPORTED;
    }
    if (!status) {
      mop_set_error("eigh: the tridiagonal path needs a status array (fallback flags)");
      return MOP_ERR_INVALID;
    }
    // blocked shared-memory tridiagonalisation + global-memory spectrum + register back-transform (eigh_large.cu)
    int rc = mop_launch_eigh_large(B, n, A, evals, evecs, status, (char*)work + jac, work_bytes - jac, stream);
    if (rc != MOP_OK) return rc;
    // robust fallback for structures the fast path flagged (no host sync: CTAs of
    // unflagged structures exit immediately)
    return mop_launch_eigh_jacobi(B, n, A, evals, evecs, status, status, work, jac, stream);
  }
  if (algo == MOP_EIGH_LARGE) {
    if (!status) {
      mop_set_error("eigh: the large-n path needs a status array (fallback flags)");
      return MOP_ERR_INVALID;
    }
    int rc = mop_launch_eigh_large(B, n, A, evals, evecs, status, (char*)work + jac, work_bytes - jac, stream);
    if (rc != MOP_OK) return rc;
    return mop_launch_eigh_jacobi(B, n, A, evals, evecs, status, status, work, jac, stream);
  }
  mop_set_error("eigh: unknown algorithm id %d", algo);
  return MOP_ERR_INVALID;
}

extern "C" int mop_eigh(int B, int n, int algo, const double* A, double* evals, double* evecs,
                        int32_t* status, void* work, size_t work_bytes, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_eigh: B >= 0 and n > 0 required");
  MOP_REQUIRE(A && evals && evecs && work, "mop_eigh: A, evals, evecs, work must be device pointers");
  if (B == 0) return MOP_OK;
  return run_eigh(B, n, algo, A, evals, evecs, status, work, work_bytes, (cudaStream_t)stream);
}

// workspace of mop_rsirfo_spectral_step: evecs | evals | jacobi work | tridiag work
extern "C" size_t mop_rsirfo_spectral_workspace_bytes(int B, int n) {
  if (B <= 0 || n <= 0) return 0;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  return nn + nv + align256(mop_jacobi_workspace_bytes(B, n)) + align256(mop_tridiag_workspace_bytes(B, n));
}

extern "C" int mop_rsirfo_spectral_step(int B, int n, int saddle_order, int neb_mode, double trust_min,
                                        double trust_max, const double* Hp, const double* gp,
                                        const double* Bg, const double* Be, double* state,
                                        double* move_out, double* eigvals_out, double* pred_out,
                                        int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0, "mop_rsirfo_spectral_step: B >= 0 and n > 0 required");
  MOP_REQUIRE(Hp && gp && Bg && state && move_out && status && work,
              "mop_rsirfo_spectral_step: Hp, gp, Bg, state, move_out, status, work must be device pointers");
  if (B == 0) return MOP_OK;
  if (!mop_tridiag_supported(n)) {
    mop_set_error("mop_rsirfo_spectral_step: n = %d exceeds the shared-memory path", n);
    return MOP_ERR_UNSUPPORTED;
  }
  if (work_bytes < mop_rsirfo_spectral_workspace_bytes(B, n)) {
    mop_set_error("mop_rsirfo_spectral_step: workspace too small (%zu < %zu bytes)", work_bytes,
                  mop_rsirfo_spectral_workspace_bytes(B, n));
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  const size_t jac = align256(mop_jacobi_workspace_bytes(B, n));
  char* w = (char*)work;
  double* evecs = (double*)w;
  double* evals = (double*)(w + nn);
  void* jwork = w + nn + nv;
  void* twork = w + nn + nv + jac;
  // tridiagonalise, solve and step in one shared-memory-resident kernel; structures it
  // flags (tight eigenvalue clusters) are redone by the robust Jacobi path (no host sync:
  // CTAs of unflagged structures exit immediately).
  int rc = mop_launch_rsirfo_fused(B, n, saddle_order, neb_mode, trust_min, trust_max, Hp, gp, Bg, Be,
                                   state, move_out, eigvals_out, pred_out, status, twork,
                                   work_bytes - (nn + nv + jac), evecs /* free until the fallback runs */, stream);
  if (rc != MOP_OK) return rc;
  // the working matrices of the flagged structures go to the (now dead) reflector / pivot slabs, not to
  // shared memory: the launch is almost always empty and must not wait for whole SMs
  rc = mop_launch_eigh_jacobi_ext(B, n, Hp, evals, evecs, status, status, jwork, jac, (double*)twork, stream);
  if (rc != MOP_OK) return rc;
  return mop_launch_rfo_step(B, n, saddle_order, neb_mode, trust_min, trust_max, evals, evecs, gp, Bg,
                             Be, state, move_out, eigvals_out, pred_out, status, 1, stream);
}

// n <= 160: ONE kernel reads H, applies the update (raw gradients, rsirfo.py:308-309,1316-1372), writes H back,
// projects gradient and effective Hessian (rsirfo.py:337,349-358) and tridiagonalises on the triangle in shared
// memory - the projected Hessian never exists in HBM; the spectrum kernel finishes the step.  packed: H / Hbias are
// packed lower triangles.  Hp, gp, rest: the workspace carve of mop_rsirfo_step (status already zeroed).
// phase & 1: front end + tridiagonalisation of structures [b0, b0 + bc); phase & 2: spectrum, step and fallbacks of all B
// (mop_rsirfo_step_packed_begin / _finish run the two phases as separate calls: chunks of a batch that arrives from the
// host are reduced while the next chunk is still in flight, the spectrum kernel then runs once with the whole batch
// resident - it needs seven CTAs per SM to hide its dependent chains and cannot share an SM with the reduction).
static int rsirfo_step_fused(int B, int n, int method, const int32_t* method_per, int saddle_order, int neb_mode, double trust_min,
                             double trust_max, int packed, double* H, const double* Hbias, const double* x,
                             const double* Bg, const double* g, const double* x_prev, const double* g_prev,
                             const double* Be, double* state, double* move_out, double* eigvals_out, double* pred_out,
                             int32_t* status, double* Hp, double* gp, char* rest, cudaStream_t stream, int phase = 3,
                             int b0 = 0, int bc = -1) {
  if (bc < 0) bc = B;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  const size_t jac = align256(mop_jacobi_workspace_bytes(B, n));
  double* zbuf = (double*)rest;                    // evecs slab of the spectral layout
  double* evals2 = (double*)(rest + nn);
  void* jwork = rest + nn + nv;
  double* twork = (double*)(rest + nn + nv + jac);
  const size_t bnn = (size_t)B * n * n, bn = (size_t)B * n;
  double* Vh = twork;
  double* Dm = twork + bnn;
  double* pd = twork + 2 * bnn;
  double* pe = pd + bn;
  double* pt = pd + 2 * bn;
  double* pg = pd + 3 * bn;
  int* pflag = (int*)(pd + 4 * bn);
  int rc = MOP_OK;
  if (phase & 1) {
    const size_t hs = packed ? (size_t)n * (n + 1) / 2 : (size_t)n * n, o = (size_t)b0, on = o * n;
    rc = mop_launch_front_tridiag_blk(bc, n, x_prev ? method : MOP_UPD_NONE, (x_prev && method_per) ? method_per + o : nullptr, 1,
                                      0, packed, H + o * hs, Hbias ? Hbias + o * hs : nullptr, x + on,
                                      x_prev ? x_prev + on : nullptr, g + on, g_prev ? g_prev + on : nullptr, Bg + on,
                                      state + o * MOP_RSIRFO_STATE, MOP_RSIRFO_STATE, gp + on, status + o,
                                      Vh + o * n * n, pd + on, pe + on, pt + on, pg + on, pflag + o,
                                      // staged reduction through the pivot slab (free until the spectrum kernel runs) when the
                                      // whole batch is reduced at once; a chunk of a streamed batch (phase 1 alone) is about one
                                      // wave of the first stage and would leave the denser later stages mostly empty
                                      phase == 3 ? Dm + o * n * n : nullptr, stream);
    if (rc != MOP_OK) return rc;
  }
  if (!(phase & 2)) return MOP_OK;
  rc = mop_launch_spectrum_step(B, n, saddle_order, neb_mode, trust_min, trust_max, Vh, zbuf, Dm, pd, pe, pt, pg, pflag,
                                Bg, Be, state, move_out, eigvals_out, pred_out, status, stream);
  if (rc != MOP_OK) return rc;
  // robust path for the structures the spectrum kernel flagged (almost always none: empty launches, no host sync)
  const double* Hfull = H;
  const double* Hbfull = Hbias;
  if (packed) {  // the flagged structures' Hessians as full squares, in slabs that are dead by now
    dim3 grid(32, B);
    mop::k_unpack_lower<<<grid, 256, 0, stream>>>(n, H, Dm, status);
    if (Hbias) mop::k_unpack_lower<<<grid, 256, 0, stream>>>(n, Hbias, Vh, status);
    MOP_CHECK_CUDA(cudaGetLastError());
    Hfull = Dm;
    Hbfull = Hbias ? Vh : nullptr;
  }
  rc = mop_launch_project_trrot_flagged(B, n, Hfull, Hbfull, x, Hp, status, stream);
  if (rc != MOP_OK) return rc;
  // (packed: the Jacobi working matrices must not alias the unpacked inputs - they go to the eigenvector slab's twin)
  rc = mop_launch_eigh_jacobi_ext(B, n, Hp, evals2, zbuf, status, status, jwork, jac, packed ? nullptr : twork, stream);
  if (rc != MOP_OK) return rc;
  return mop_launch_rfo_step(B, n, saddle_order, neb_mode, trust_min, trust_max, evals2, zbuf, gp, Bg, Be, state,
                             move_out, eigvals_out, pred_out, status, 1, stream);
}

// workspace of mop_rsirfo_step: Hp | evecs | evals | gp | eigh work
extern "C" size_t mop_rsirfo_workspace_bytes(int B, int n, int algo) {
  if (B <= 0 || n <= 0) return 0;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  const size_t generic = nn + 3 * nv + eigh_work_bytes(B, n, algo);    // evecs | evals | 2 rotated vectors | eigh work
  const size_t fused = mop_rsirfo_spectral_workspace_bytes(B, n);
  return nn + nv + (generic > fused ? generic : fused);                  // Hp | gp | rest
}

extern "C" int mop_rsirfo_step(int B, int n, int method, int saddle_order, int neb_mode,
                               int eigh_algo, double trust_min, double trust_max, double* H,
                               const double* Hbias, const double* x, const double* Bg,
                               const double* g, const double* x_prev, const double* g_prev,
                               const double* Be, double* state, double* move_out,
                               double* eigvals_out, double* pred_out, int32_t* status, void* work,
                               size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0 && n % 3 == 0, "mop_rsirfo_step: n must be a positive multiple of 3");
  MOP_REQUIRE(H && x && Bg && g && state && move_out && status && work,
              "mop_rsirfo_step: H, x, Bg, g, state, move_out, status, work must be device pointers");
  MOP_REQUIRE(saddle_order >= 0 && saddle_order < n, "mop_rsirfo_step: bad saddle_order");
  MOP_REQUIRE((x_prev == nullptr) == (g_prev == nullptr),
              "mop_rsirfo_step: x_prev and g_prev must both be given or both be NULL");
  if (B == 0) return MOP_OK;
  if (work_bytes < mop_rsirfo_workspace_bytes(B, n, eigh_algo)) {
    mop_set_error("mop_rsirfo_step: workspace too small (%zu < %zu bytes)", work_bytes,
                  mop_rsirfo_workspace_bytes(B, n, eigh_algo));
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  // layout: Hp | gp | rest ; rest = spectral workspace (fused path) or evecs | evals | eigh work
  double* Hp = (double*)w;
  double* gp = (double*)(w + nn);
  char* rest = w + nn + nv;
  const size_t rest_bytes = work_bytes - (nn + nv);
  double* evecs = (double*)rest;
  double* evals = (double*)(rest + nn);
  double* Xg = (double*)(rest + nn + nv);
  double* Xmv = (double*)(rest + nn + 2 * nv);
  char* ework = rest + nn + 3 * nv;
  const size_t ebytes = rest_bytes - (nn + 3 * nv);

  MOP_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, stream));
  int rc = MOP_OK;
  if (pick_algo(eigh_algo, n) == MOP_EIGH_TRIDIAG && mop_spectrum_step_supported(n))
    return rsirfo_step_fused(B, n, method, nullptr, saddle_order, neb_mode, trust_min, trust_max, 0, H, Hbias, x, Bg, g, x_prev,
                             g_prev, Be, state, move_out, eigvals_out, pred_out, status, Hp, gp, rest, stream);
  // (1) Hessian update with RAW gradients (rsirfo.py:308-309,1316-1372) and (2) TR/ROT projection of gradient
  // and effective Hessian (rsirfo.py:337,349-358) for n > 160.  Small batches take the multi-CTA projection (it
  // fills the GPU), large ones one CTA per structure (as fast, fewer launches).  (Running the pair chunk by chunk so
  // that a chunk stays in L2 between its passes was measured SLOWER on B200: the short launches are latency bound.)
  const size_t n2 = (size_t)n * n;
  const bool chunked = false;
  const size_t slab = sizeof(double) * (size_t)B * n2;  // what a [B][n][n] slab really holds (nn is rounded up)
  const bool split_ok = slab >= mop_project_scratch_bytes(B, n) && slab >= mop_hessian_update_scratch_bytes(B, n) &&
                        (chunked || B <= 2 * 148);
  const int CH = B;
  for (int b0 = 0; b0 < B; b0 += CH) {
    const int bc = B - b0 < CH ? B - b0 : CH;
    double* Hc = H + b0 * n2;
    const double* Hbc = Hbias ? Hbias + b0 * n2 : nullptr;
    const size_t chunk_bytes = sizeof(double) * bc * n2;
    if (x_prev && method != MOP_UPD_NONE) {
      // scratch: the projected-Hessian buffer of this chunk is not live yet
      // (tiny n: the multi-CTA scratch, B (4 n + 24) doubles, does not fit the Hp slab - one CTA per structure)
      const size_t upd_scr = split_ok ? chunk_bytes : nn;
      rc = upd_scr >= mop_hessian_update_scratch_bytes(bc, n)
               ? mop_launch_hessian_update_split(bc, n, method, 1, 1, Hc, nullptr, nullptr, x + (size_t)b0 * n,
                                                 x_prev + (size_t)b0 * n, g + (size_t)b0 * n, g_prev + (size_t)b0 * n,
                                                 state + (size_t)b0 * MOP_RSIRFO_STATE, MOP_RSIRFO_STATE, nullptr,
                                                 status + b0, split_ok ? (void*)(Hp + b0 * n2) : (void*)Hp, upd_scr,
                                                 stream)
               : mop_launch_hessian_update(bc, n, method, 1, 1, Hc, nullptr, nullptr, x + (size_t)b0 * n,
                                           x_prev + (size_t)b0 * n, g + (size_t)b0 * n, g_prev + (size_t)b0 * n,
                                           state + (size_t)b0 * MOP_RSIRFO_STATE, MOP_RSIRFO_STATE, nullptr,
                                           status + b0, stream);
      if (rc != MOP_OK) return rc;
    }
    rc = split_ok ? mop_launch_project_trrot_split(bc, n, Hc, Hbc, x + (size_t)b0 * n, Bg + (size_t)b0 * n,
                                                   Hp + b0 * n2, gp + (size_t)b0 * n, status + b0, 0, evecs + b0 * n2,
                                                   chunk_bytes, stream)
                  : mop_launch_project_trrot(bc, n, Hc, Hbc, x + (size_t)b0 * n, Bg + (size_t)b0 * n, Hp + b0 * n2,
                                             gp + (size_t)b0 * n, status + b0, 0, stream);
    if (rc != MOP_OK) return rc;
  }
  if (pick_algo(eigh_algo, n) == MOP_EIGH_TRIDIAG)
    return mop_rsirfo_spectral_step(B, n, saddle_order, neb_mode, trust_min, trust_max, Hp, gp, Bg, Be,
                                    state, move_out, eigvals_out, pred_out, status, rest, rest_bytes,
                                    stream_);
  if (pick_algo(eigh_algo, n) == MOP_EIGH_LARGE) {
    // large systems: Hp = Q T Q^T, the step is taken in the basis of T (eigh_large.cu, part 4):
    // gp -> Q^T gp, eigenvectors of T instead of V = Q Z, move -> Q move
    const size_t jac = align256(mop_jacobi_workspace_bytes(B, n));
    rc = mop_launch_eigh_large_factored(B, n, Hp, evals, evecs, status, ework + jac, ebytes - jac, stream);
    if (rc != MOP_OK) return rc;
    rc = mop_launch_eigh_jacobi(B, n, Hp, evals, evecs, status, status, ework, jac, stream);
    if (rc != MOP_OK) return rc;
    const size_t vbytes = sizeof(double) * (size_t)B * n;
    MOP_CHECK_CUDA(cudaMemcpyAsync(Xg, gp, vbytes, cudaMemcpyDeviceToDevice, stream));
    rc = mop_launch_large_apply_q(B, n, 1, ework + jac, Xg, nullptr, nullptr, nullptr, stream);
    if (rc != MOP_OK) return rc;
    rc = mop_launch_rfo_step(B, n, saddle_order, neb_mode, trust_min, trust_max, evals, evecs, Xg, Bg, Be, state,
                             Xmv, eigvals_out, pred_out, status, 0, stream);
    if (rc != MOP_OK) return rc;
    rc = mop_launch_large_apply_q(B, n, 0, ework + jac, Xmv, nullptr, nullptr, nullptr, stream);
    if (rc != MOP_OK) return rc;
    MOP_CHECK_CUDA(cudaMemcpyAsync(move_out, Xmv, vbytes, cudaMemcpyDeviceToDevice, stream));
    return MOP_OK;
  }
  // (3) eigendecomposition (rsirfo.py:360)
  rc = run_eigh(B, n, eigh_algo, Hp, evals, evecs, status, ework, ebytes, stream);
  if (rc != MOP_OK) return rc;
  // (4) image function, secular solve, step, bookkeeping (rsirfo.py:365-490)
  return mop_launch_rfo_step(B, n, saddle_order, neb_mode, trust_min, trust_max, evals, evecs, gp,
                             Bg, Be, state, move_out, eigvals_out, pred_out, status, 0, stream);
}

// RSIRFO.run with the Hessians in PACKED lower-triangular storage, [B][n (n + 1) / 2] (row i at i (i + 1) / 2): half
// the bytes in HBM and over PCIe; otherwise mop_rsirfo_step.  n <= 160 (the fused shared-memory path) only.
extern "C" int mop_rsirfo_step_packed(int B, int n, int method, int saddle_order, int neb_mode, double trust_min,
                                      double trust_max, double* H_packed, const double* Hbias_packed, const double* x,
                                      const double* Bg, const double* g, const double* x_prev, const double* g_prev,
                                      const double* Be, double* state, double* move_out, double* eigvals_out,
                                      double* pred_out, int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0 && n % 3 == 0, "mop_rsirfo_step_packed: n must be a positive multiple of 3");
  MOP_REQUIRE(H_packed && x && Bg && g && state && move_out && status && work,
              "mop_rsirfo_step_packed: H, x, Bg, g, state, move_out, status, work must be device pointers");
  MOP_REQUIRE(saddle_order >= 0 && saddle_order < n, "mop_rsirfo_step_packed: bad saddle_order");
  MOP_REQUIRE((x_prev == nullptr) == (g_prev == nullptr),
              "mop_rsirfo_step_packed: x_prev and g_prev must both be given or both be NULL");
  if (B == 0) return MOP_OK;
  if (!mop_spectrum_step_supported(n) || !mop_tridiag_supported(n)) {
    mop_set_error("mop_rsirfo_step_packed: n = %d is outside the shared-memory path (3 .. 160)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  if (work_bytes < mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG)) {
    mop_set_error("mop_rsirfo_step_packed: workspace too small (%zu < %zu bytes)", work_bytes,
                  mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG));
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  MOP_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, stream));
  return rsirfo_step_fused(B, n, method, nullptr, saddle_order, neb_mode, trust_min, trust_max, 1, H_packed, Hbias_packed, x, Bg,
                           g, x_prev, g_prev, Be, state, move_out, eigvals_out, pred_out, status, (double*)w,
                           (double*)(w + nn), w + nn + nv, stream);
}

// The packed step in two calls, for a batch that is streamed in from the host chunk by chunk (see rsirfo_step_fused).
// Every pointer addresses the WHOLE batch of B structures; _begin works on structures [b0, b0 + bc) only and may be
// issued on a stream of its own per chunk, _finish follows once on a stream that waits for all of them.
static int packed_args_ok(const char* who, int B, int n, int saddle_order, const void* H, const void* x, const void* Bg,
                          const void* g, const void* state, const void* status, const void* work, const void* x_prev,
                          const void* g_prev, size_t work_bytes) {
  if (!(B >= 0 && n > 0 && n % 3 == 0)) { mop_set_error("%s: n must be a positive multiple of 3", who); return MOP_ERR_INVALID; }
  if (!(H && x && Bg && g && state && status && work)) { mop_set_error("%s: null device pointer", who); return MOP_ERR_INVALID; }
  if (!(saddle_order >= 0 && saddle_order < n)) { mop_set_error("%s: bad saddle_order", who); return MOP_ERR_INVALID; }
  if ((x_prev == nullptr) != (g_prev == nullptr)) { mop_set_error("%s: x_prev and g_prev must both be given or both be NULL", who); return MOP_ERR_INVALID; }
  if (B > 0 && (!mop_spectrum_step_supported(n) || !mop_tridiag_supported(n))) {
    mop_set_error("%s: n = %d is outside the shared-memory path (3 .. 160)", who, n);
    return MOP_ERR_UNSUPPORTED;
  }
  if (work_bytes < mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG)) {
    mop_set_error("%s: workspace too small (%zu < %zu bytes)", who, work_bytes, mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG));
    return MOP_ERR_WORKSPACE;
  }
  return MOP_OK;
}

extern "C" int mop_rsirfo_step_packed_begin(int B, int b0, int bc, int n, int method, double* H_packed,
                                            const double* Hbias_packed, const double* x, const double* Bg, const double* g,
                                            const double* x_prev, const double* g_prev, double* state, int32_t* status,
                                            void* work, size_t work_bytes, void* stream_) {
  int rc = packed_args_ok("mop_rsirfo_step_packed_begin", B, n, 0, H_packed, x, Bg, g, state, status, work, x_prev, g_prev, work_bytes);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(b0 >= 0 && bc >= 0 && b0 + bc <= B, "mop_rsirfo_step_packed_begin: chunk outside the batch");
  if (bc == 0) return MOP_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  MOP_CHECK_CUDA(cudaMemsetAsync(status + b0, 0, sizeof(int32_t) * (size_t)bc, stream));
  return rsirfo_step_fused(B, n, method, nullptr, 0, 0, 0.0, 0.0, 1, H_packed, Hbias_packed, x, Bg, g, x_prev, g_prev, nullptr,
                           state, nullptr, nullptr, nullptr, status, (double*)w, (double*)(w + nn), w + nn + nv, stream, 1, b0,
                           bc);
}

extern "C" int mop_rsirfo_step_packed_finish(int B, int n, int saddle_order, int neb_mode, double trust_min,
                                             double trust_max, double* H_packed, const double* Hbias_packed, const double* x,
                                             const double* Bg, const double* Be, double* state, double* move_out,
                                             double* eigvals_out, double* pred_out, int32_t* status, void* work,
                                             size_t work_bytes, void* stream_) {
  int rc = packed_args_ok("mop_rsirfo_step_packed_finish", B, n, saddle_order, H_packed, x, Bg, Bg, state, status, work, nullptr,
                          nullptr, work_bytes);
  if (rc != MOP_OK) return rc;
  MOP_REQUIRE(move_out, "mop_rsirfo_step_packed_finish: move_out required");
  if (B == 0) return MOP_OK;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  return rsirfo_step_fused(B, n, MOP_UPD_NONE, nullptr, saddle_order, neb_mode, trust_min, trust_max, 1, H_packed, Hbias_packed,
                           x, Bg, nullptr, nullptr, nullptr, Be, state, move_out, eigvals_out, pred_out, status, (double*)w,
                           (double*)(w + nn), w + nn + nv, (cudaStream_t)stream_, 2);
}

// RSIRFO.run for a batch whose structures use DIFFERENT update methods (method_per [B], device): a NEB chain runs
// rsirfo_block_fsb at its ends and rsirfo_block_bofill inside (Optimizer/rfo_neb.py:116-121) - one launch instead of
// one per group.  Full-square Hessians; 3 <= n <= 160; workspace = mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG).
extern "C" int mop_rsirfo_step_mixed(int B, int n, const int32_t* method_per, int saddle_order, int neb_mode,
                                     double trust_min, double trust_max, double* H, const double* Hbias, const double* x,
                                     const double* Bg, const double* g, const double* x_prev, const double* g_prev,
                                     const double* Be, double* state, double* move_out, double* eigvals_out,
                                     double* pred_out, int32_t* status, void* work, size_t work_bytes, void* stream_) {
  MOP_REQUIRE(B >= 0 && n > 0 && n % 3 == 0, "mop_rsirfo_step_mixed: n must be a positive multiple of 3");
  MOP_REQUIRE(method_per && H && x && Bg && g && state && move_out && status && work,
              "mop_rsirfo_step_mixed: method_per, H, x, Bg, g, state, move_out, status, work must be device pointers");
  MOP_REQUIRE(saddle_order >= 0 && saddle_order < n, "mop_rsirfo_step_mixed: bad saddle_order");
  MOP_REQUIRE((x_prev == nullptr) == (g_prev == nullptr),
              "mop_rsirfo_step_mixed: x_prev and g_prev must both be given or both be NULL");
  if (B == 0) return MOP_OK;
  if (!mop_spectrum_step_supported(n) || !mop_tridiag_supported(n)) {
    mop_set_error("mop_rsirfo_step_mixed: n = %d is outside the shared-memory path (3 .. 160)", n);
    return MOP_ERR_UNSUPPORTED;
  }
  if (work_bytes < mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG)) {
    mop_set_error("mop_rsirfo_step_mixed: workspace too small");
    return MOP_ERR_WORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t nn = align256(sizeof(double) * (size_t)B * n * n);
  const size_t nv = align256(sizeof(double) * (size_t)B * n);
  char* w = (char*)work;
  MOP_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, stream));
  return rsirfo_step_fused(B, n, MOP_UPD_FLOWCHART, method_per, saddle_order, neb_mode, trust_min, trust_max, 0, H, Hbias,
                           x, Bg, g, x_prev, g_prev, Be, state, move_out, eigvals_out, pred_out, status, (double*)w,
                           (double*)(w + nn), w + nn + nv, stream);
}
