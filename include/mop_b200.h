/*
 * mop_b200.h — C ABI of the B200-native optimizer-step hot path.
 *
 * Drop-in boundary for the per-structure optimizer step of ss0832/MultiOptPy
 * (pure Python; it has no FFI of its own, so every entry point names the
 * Python method(s) it replaces, paths relative to multioptpy/ in the
 * reference tree).  INTEGRATION.md shows the ctypes stubs a maintainer adds.
 *
 * Conventions
 *  - every pointer is a CALLER-OWNED DEVICE pointer (cudaMalloc / torch CUDA
 *    tensor) unless marked "host"; the library allocates nothing and keeps no
 *    global state, so calls are thread-safe and CUDA-graph capturable;
 *  - FP64 everywhere, row-major, batch-major: H is [B][n][n], vectors [B][n],
 *    n = 3 * natoms, geometry in Bohr, energies in Hartree;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*,
 *    NULL = legacy default stream);
 *  - return value: MOP_OK or a negative MOP_ERR_*; never throws.
 *    mop_last_error() (host, thread-local) describes the last failure;
 *  - data-dependent branches taken per structure are reported in the
 *    int32 status word (MOP_ST_* bits), [B] on the device.
 */
#ifndef MOP_B200_H
#define MOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOP_VERSION 100

#define MOP_OK 0
#define MOP_ERR_INVALID (-1)     /* bad argument (null pointer, n <= 0, ...) */
#define MOP_ERR_CUDA (-2)        /* CUDA runtime error, see mop_last_error() */
#define MOP_ERR_UNSUPPORTED (-3) /* method / size not implemented on device  */
#define MOP_ERR_WORKSPACE (-4)   /* work_bytes smaller than *_workspace_bytes */

/* Hessian-update method ids: the prioritised substring table of
 * Optimizer/rsirfo.py:208-251 (first hit wins; no hit -> FLOWCHART). */
enum {
  MOP_UPD_NONE = 0,
  MOP_UPD_FLOWCHART = 1,
  MOP_UPD_BLOCK_CFD_FSB_DD = 2,
  MOP_UPD_BLOCK_CFD_FSB_WEIGHTED = 3,
  MOP_UPD_BLOCK_CFD_FSB = 4,
  MOP_UPD_BLOCK_CFD_BOFILL_WEIGHTED = 5,
  MOP_UPD_BLOCK_CFD_BOFILL = 6,
  MOP_UPD_BLOCK_BFGS_DD = 7,
  MOP_UPD_BLOCK_BFGS = 8,
  MOP_UPD_BLOCK_FSB_DD = 9,
  MOP_UPD_BLOCK_FSB_WEIGHTED = 10,
  MOP_UPD_BLOCK_FSB = 11,
  MOP_UPD_BLOCK_BOFILL_WEIGHTED = 12,
  MOP_UPD_BLOCK_BOFILL = 13,
  MOP_UPD_BFGS_DD = 14,
  MOP_UPD_BFGS = 15,
  MOP_UPD_SR1 = 16,
  MOP_UPD_PCFD_BOFILL = 17, /* O(n^4) null-space variant: MOP_ERR_UNSUPPORTED */
  MOP_UPD_CFD_FSB_DD = 18,
  MOP_UPD_CFD_FSB = 19,
  MOP_UPD_CFD_BOFILL = 20,
  MOP_UPD_FSB_DD = 21,
  MOP_UPD_FSB = 22,
  MOP_UPD_BOFILL = 23,
  MOP_UPD_PSB = 24,
  MOP_UPD_MSP = 25
};

/* per-structure status bits */
#define MOP_ST_UPDATED (1 << 0)         /* Hessian update applied                         */
#define MOP_ST_UPD_SKIP_SMALL (1 << 1)  /* ||s|| or ||y|| < 1e-10 (rsirfo.py:1326)        */
#define MOP_ST_UPD_SKIP_CURV (1 << 2)   /* s.y <= 0 (rsirfo.py:1333)                      */
#define MOP_ST_UPD_TERM_ZEROED (1 << 3) /* a guarded denominator fired, term dropped      */
#define MOP_ST_LEVEL_SHIFT (1 << 4)     /* kappa > 1e8: +1e-5 level shift (rsirfo.py:618) */
#define MOP_ST_EIG_NONFINITE (1 << 5)   /* NaN/Inf spectrum -> identity (rsirfo.py:365)   */
#define MOP_ST_ALPHA_SEARCH (1 << 6)    /* ||step|| > inner trust: alpha loop ran         */
#define MOP_ST_STEP_NAN_SD (1 << 7)     /* NaN step -> steepest descent (rsirfo.py:456)   */
#define MOP_ST_HARD_CASE (1 << 8)       /* all gradient components zero (rsirfo.py:1559)  */
#define MOP_ST_TRROT_RANKDEF (1 << 9)   /* < 6 independent TR/ROT vectors (linear)        */
#define MOP_ST_BRENT_BRACKET (1 << 10)  /* alpha bracket seen; Brent branch not replayed  */
#define MOP_ST_EIG_NOCONV (1 << 11)     /* eigensolver hit its sweep / iteration limit    */
#define MOP_ST_EIG_FALLBACK (1 << 12)   /* robust Jacobi fallback produced the spectrum   */
#define MOP_ST_NO_HISTORY (1 << 13)     /* first call: no previous point, update skipped  */
#define MOP_ST_UPD_REJECTED (1 << 14)   /* P-RFO: updated spectrum > 1e6, update reverted (rsprfo.py:1247) */
#define MOP_ST_ALPHA_UNSTABLE (1 << 16) /* alpha loop did not end on its rounding-free exits: the secular root is ill-conditioned, the reference's own step is summation-order dependent (rfo_secular.cuh) */
#define MOP_ST_CONSTR_CONVERGED (1 << 17) /* CRSIRFO: subspace gradient below the threshold, zero step returned (crsirfo.py:108-118) */
#define MOP_ST_LINDH_NO_K (1 << 15)     /* Lindh: non-zero gradient but no internal gradient, K term omitted */

/* eigensolver selection */
#define MOP_EIGH_AUTO 0
#define MOP_EIGH_JACOBI 1  /* two-sided cyclic Jacobi, matrix resident in shared memory */
#define MOP_EIGH_TRIDIAG 2 /* Householder tridiagonalisation + bisection + inverse iteration */
#define MOP_EIGH_LARGE 3   /* same algorithm for 160 < n <= 1024: matrix streamed from L2 by a thread-block cluster */

/* RSIRFO per-structure state: [B][MOP_RSIRFO_STATE] doubles (RSIRFO attributes
 * of Optimizer/rsirfo.py:97-112 that survive between run() calls). */
#define MOP_RSIRFO_STATE 16
#define MOP_RS_TRUST 0        /* self.trust_radius                          */
#define MOP_RS_HAVE_PREV 1    /* prev_geometry/prev_gradient set (0/1)      */
#define MOP_RS_PREV_ENERGY 2  /* self.prev_energy                           */
#define MOP_RS_HAVE_ENERGY 3  /* prev_energy is not None (0/1)              */
#define MOP_RS_NPRED 4        /* len(predicted_energy_changes) <= 3         */
#define MOP_RS_PRED0 5        /* .. 7 : predicted_energy_changes, oldest first */
#define MOP_RS_NACT 8         /* len(actual_energy_changes) <= 3            */
#define MOP_RS_ACT0 9         /* .. 11                                      */
#define MOP_RS_ITER 12        /* self.iteration                             */

int mop_version(void);
const char* mop_last_error(void);

/* ---- (1) Hessian update ------------------------------------------------
 * Replaces ModelHessianUpdate.*_hessian_update (Optimizer/hessian_update.py:
 * 248-433) and BlockHessianUpdate.block_*_hessian_update
 * (Optimizer/block_hessian_update.py:443-709; history depth is always 1).
 * mode 0: delta_out[B][n][n] = delta_hess (the operator contract; H untouched)
 * mode 1: H <- 1/2 ((H + delta) + (H + delta)^T) in place, as
 *         RSIRFO.update_hessian does (Optimizer/rsirfo.py:1361-1372), with its
 *         skip rules (:1326,:1333) when rsirfo_guards != 0.
 * s, y: [B][n] displacement and gradient difference.
 * work (optional, mop_hessian_update_workspace_bytes): with it the update runs as three multi-CTA kernels
 * (row-block H s, per-structure coefficients, tile-row apply); with work = NULL as one CTA per structure. */
size_t mop_hessian_update_workspace_bytes(int B, int n);
int mop_hessian_update(int B, int n, int method, int mode, int rsirfo_guards, double* H,
                       const double* s, const double* y, double* delta_out, int32_t* status,
                       void* work, size_t work_bytes, void* stream);

/* ---- (2a) TR/ROT projection ---------------------------------------------
 * Replaces Calculationtools.project_out_hess_tr_and_rot_for_coord
 * (Utils/calc_tools.py:249-316) and RSIRFO._project_grad_tr_rot
 * (Optimizer/rsirfo.py:128-190).  Hp_out = sym(P^T (H + Hbias) P); Hbias,
 * g, gp_out may be NULL. */
int mop_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                      const double* g, double* Hp_out, double* gp_out, int32_t* status,
                      void* stream);

/* ---- (2b) batched symmetric eigensolver -----------------------------------
 * Replaces numpy.linalg.eigh at Optimizer/rsirfo.py:606,626,652.
 * evals[B][n] ascending; evecs[B][n][n] with ROW k = eigenvector k. */
size_t mop_eigh_workspace_bytes(int B, int n, int algo);
int mop_eigh(int B, int n, int algo, const double* A, double* evals, double* evecs,
             int32_t* status, void* work, size_t work_bytes, void* stream);

/* ---- (2c) one RS-I-RFO step ---------------------------------------------
 * Replaces RSIRFO.run (Optimizer/rsirfo.py:285-490): Hessian update from
 * (x - x_prev, g - g_prev) with RAW gradients, TR/ROT projection of gradient
 * and Hessian, eigendecomposition with conditional level shift, image
 * projection for `saddle_order` roots (factor 1 instead of 2 when neb_mode),
 * small-eigenvalue filter, secular-equation RFO solve, alpha loop when the
 * step exceeds the inner trust radius, NaN fallbacks, predicted energy change
 * and inner trust-radius bookkeeping.
 * move_out[B][n] = the value RSIRFO.run returns (minus the RFO step; the
 * caller computes x_new = x - move, optimizer.py:798).
 * x_prev / g_prev may be NULL (no history).  Hbias may be NULL. */
/* Kernel launches of the (staged) shared-memory tridiagonalisation inside mop_rsirfo_step / mop_eigh for n <= 160: the
 * reduction is handed to denser launches as the trailing block shrinks (5 at n = 150, 3 at n = 72, 1 at n <= 47); batches
 * of at most 296 structures (two CTAs per SM) are reduced in one launch. */
int mop_tridiag_stage_count(int n);
size_t mop_rsirfo_workspace_bytes(int B, int n, int algo);
int mop_rsirfo_step(int B, int n, int method, int saddle_order, int neb_mode, int eigh_algo,
                    double trust_min, double trust_max, double* H, const double* Hbias,
                    const double* x, const double* Bg, const double* g, const double* x_prev,
                    const double* g_prev, const double* Be, double* state, double* move_out,
                    double* eigvals_out, double* pred_out, int32_t* status, void* work,
                    size_t work_bytes, void* stream);

/* The same step with the Hessians in PACKED lower-triangular storage, [B][n (n + 1) / 2] with row i at
 * i (i + 1) / 2 (they are symmetric: RSIRFO symmetrises after every update, rsirfo.py:1372): half the bytes in HBM
 * and over PCIe.  3 <= n <= 160 (the shared-memory path); workspace = mop_rsirfo_workspace_bytes(B, n,
 * MOP_EIGH_TRIDIAG).  mop_pack_lower / mop_unpack_lower convert between the layouts on the device. */
int mop_rsirfo_step_packed(int B, int n, int method, int saddle_order, int neb_mode, double trust_min,
                           double trust_max, double* H_packed, const double* Hbias_packed, const double* x,
                           const double* Bg, const double* g, const double* x_prev, const double* g_prev,
                           const double* Be, double* state, double* move_out, double* eigvals_out, double* pred_out,
                           int32_t* status, void* work, size_t work_bytes, void* stream);
/* The packed step as TWO calls, for a batch that streams in from host memory chunk by chunk (the host-buffer
 * pipeline multioptpy_b200/host_pipeline.py; RSIRFO.run of Optimizer/rsirfo.py:285-490 for every structure).
 * All pointers address the whole batch of B structures.  _begin: Hessian update, write-back, TR/ROT projection and
 * tridiagonalisation of structures [b0, b0 + bc) - one call per chunk, each on the stream that carried the chunk's
 * copies.  _finish: spectrum, eigenvectors, RS-I-RFO step and the robust fallbacks of all B, on a stream that waits
 * for every _begin.  Same workspace (mop_rsirfo_workspace_bytes(B, n, MOP_EIGH_TRIDIAG)) for all calls of a step;
 * results identical to mop_rsirfo_step_packed. */
int mop_rsirfo_step_packed_begin(int B, int b0, int bc, int n, int method, double* H_packed, const double* Hbias_packed,
                                 const double* x, const double* Bg, const double* g, const double* x_prev,
                                 const double* g_prev, double* state, int32_t* status, void* work, size_t work_bytes,
                                 void* stream);
int mop_rsirfo_step_packed_finish(int B, int n, int saddle_order, int neb_mode, double trust_min, double trust_max,
                                  double* H_packed, const double* Hbias_packed, const double* x, const double* Bg,
                                  const double* Be, double* state, double* move_out, double* eigvals_out,
                                  double* pred_out, int32_t* status, void* work, size_t work_bytes, void* stream);
/* The same step for a batch whose structures use DIFFERENT update methods, method_per [B] int32 on the device
 * (a NEB chain: rsirfo_block_fsb at the ends, rsirfo_block_bofill inside, Optimizer/rfo_neb.py:116-121). */
int mop_rsirfo_step_mixed(int B, int n, const int32_t* method_per, int saddle_order, int neb_mode, double trust_min,
                          double trust_max, double* H, const double* Hbias, const double* x, const double* Bg,
                          const double* g, const double* x_prev, const double* g_prev, const double* Be, double* state,
                          double* move_out, double* eigvals_out, double* pred_out, int32_t* status, void* work,
                          size_t work_bytes, void* stream);
int mop_pack_lower(int B, int n, const double* H, double* packed, void* stream);
int mop_unpack_lower(int B, int n, const double* packed, double* H, void* stream);

/* ---- (2e) one RS-P-RFO step ----------------------------------------------------
 * Replaces EnhancedRSPRFO.run (Optimizer/rsprfo.py:713-886): reduction ratio of the previous
 * step and Nocedal-Wright trust-radius update, Hessian update with the BIASED gradients (small
 * change skip only), TR/ROT projection of the gradient, eigendecomposition of H + Hbias (not
 * projected), eigenvalue shifting, mode following, P-RFO step from the extreme eigenpairs of
 * the max / min arrowhead matrices, trust and gradient-based scaling, predicted energy change.
 * move_out [B][n] = the value run() returns (the caller computes x - move).
 * Per-structure state: state [B][MOP_PRFO_STATE] doubles (column 0 = trust radius, initialise
 * to 0.1 for saddle searches / 0.5 for minimisations, rest 0), prev_grad / prev_move / ts_vec
 * [B][n] scratch owned by the caller across calls.  pre_move, x_prev, Bg_prev, Hbias may be NULL.
 * An update whose spectrum exceeds 1e6 is reverted bit-exactly (rsprfo.py:1242-1250; MOP_ST_UPD_REJECTED): the
 * Frobenius norm decides almost every structure (||H||_F <= 1e6 accepts, ||H||_F > 1e6 sqrt(n) rejects), the
 * band between them goes through the Jacobi eigensolver. */
#define MOP_PRFO_STATE 8
size_t mop_rsprfo_workspace_bytes(int B, int n, int eigh_algo);
int mop_rsprfo_step(int B, int n, int method, int saddle_order, int eigh_algo, double trust_min,
                    double trust_max, double* H, const double* Hbias, const double* x, const double* Bg,
                    const double* x_prev, const double* Bg_prev, const double* pre_move, const double* Be,
                    double* state, double* prev_grad, double* prev_move, double* ts_vec, double* move_out,
                    double* eigvals_out, double* pred_out, int32_t* status, void* work, size_t work_bytes,
                    void* stream);

/* ---- (2d) RS-I-RFO step from an already projected Hessian ---------------------
 * The part of RSIRFO.run after the projections (Optimizer/rsirfo.py:358-490): one
 * shared-memory-resident kernel per structure that tridiagonalises Hp, finds the
 * spectrum, solves the RFO secular equation in the eigenbasis and transforms the step
 * back (n <= 160); structures with tight eigenvalue clusters are redone by the Jacobi
 * path.  Hp [B][n][n] must be symmetric; gp = projected gradient, Bg = raw biased
 * gradient (its norm drives the inner trust-radius rule, rsirfo.py:312,835). */
size_t mop_rsirfo_spectral_workspace_bytes(int B, int n);
int mop_rsirfo_spectral_step(int B, int n, int saddle_order, int neb_mode, double trust_min,
                             double trust_max, const double* Hp, const double* gp,
                             const double* Bg, const double* Be, double* state, double* move_out,
                             double* eigvals_out, double* pred_out, int32_t* status, void* work,
                             size_t work_bytes, void* stream);

/* ---- (2e) constrained RS-I-RFO: the subspace projection of CRSIRFO.run ---------------
 * Replaces CRSIRFO._get_null_space_basis + the projections of CRSIRFO.run (Optimizer/crsirfo.py:16-45,88-100).
 * C [B][k][n]: the raw constraint rows of constraints_obj._get_all_constraint_vectors (1 <= k <= 12); they are
 * normalised, their span truncated by the reference's singular-value rule (svd_threshold, default 1e-5) and
 * projected out of g (+ H shake when shake [B][n], the SHAKE displacement, is longer than 1e-6) and of sym(H):
 * gp_out [B][n], Hp_out [B][n][n] (the constrained directions carry the eigenvalue ||sym(H)||_F + 1 and no
 * gradient), rank_out [B].  mop_rsirfo_spectral_step(Hp, gp, Bg = gp, ...) is then the rest of CRSIRFO.run;
 * mop_crsirfo_finalize applies its explicit convergence test (|gp| < grad_threshold: zero step, state restored from
 * state_before with only the "previous point" fields set, status = MOP_ST_CONSTR_CONVERGED).
 * mop_add_inplace: dst += src (the reference adds the bias Hessian INTO self.hessian, crsirfo.py:76,86). */
int mop_constraint_project(int B, int n, int k, double svd_threshold, const double* C, const double* H, const double* g,
                           const double* shake, double* Hp_out, double* gp_out, int32_t* rank_out, void* stream);
int mop_crsirfo_finalize(int B, int n, double grad_threshold, const double* gp, const double* Be,
                         const double* state_before, double* state, double* move, double* pred, int32_t* status,
                         void* stream);
int mop_add_inplace(size_t count, double* dst, const double* src, void* stream);

/* ---- (3a) connectivity tables ------------------------------------------------
 * Replaces BondConnectivity.connectivity_table (Utils/bond_connectivity.py:7-134):
 * bonds [B][capB][2] (i <= j), angles [B][capA][3] as [j, i, n] with centre i, dihedrals
 * [B][capD][4]; counts [B][3].  Index values and ORDER are bit-exact with the reference.
 * radii: covalent radii in Bohr, [B][natoms] (radii_stride = natoms) or one molecule shared
 * by the batch (radii_stride = 0); factor = 1.1 in the reference.  status[b] = 1 when a
 * table overflowed its capacity. */
int mop_connectivity(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                     double factor, int capB, int capA, int capD, int32_t* bonds, int32_t* angles,
                     int32_t* dihedrals, int32_t* counts, int32_t* status, void* stream);

/* ---- (3b) Fischer model Hessian ------------------------------------------------
 * Replaces FischerApproxHessian.main (ModelHessian/fischer.py:212-236), i.e.
 * ApproxHessian().main(coord, element_list, cart_gradient, "fischer"): bond / angle /
 * dihedral force constants, sum k b b^T with the Wilson vectors of
 * ModelHessian/calc_params.py, upper-to-lower symmetrisation, TR/ROT projection.
 * H_out [B][3N][3N]; counts_out (optional) [B][3] = table sizes. */
size_t mop_fischer_workspace_bytes(int B, int natoms);
int mop_fischer_hessian(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                        double* H_out, int32_t* counts_out, int32_t* status, void* work,
                        size_t work_bytes, void* stream);

/* ---- (3b') Fischer + D3 model Hessian, old variant -----------------------------------------
 * Replaces FischerD3ApproxHessianOld.main (ModelHessian/fischerd3old.py:355-381), i.e.
 * ApproxHessian().main(coord, element_list, cart_gradient, "fischerd3old") - what a bare `-modelhess` selects
 * (interface.py:184-191): the Fischer terms with the linear-angle skips (:195-210) and the sin^2-damped torsion force
 * constants (:258-300), plus the reference's simplified D3(BJ) pair blocks (:85-128) for every pair that is not bonded
 * at 1.3 x the covalent radii (:322-352), symmetrisation, TR/ROT projection.
 * atom_params [B or 1][natoms][4] = {covalent radius (Bohr), D2 C6 (hartree bohr^6), D3 r4r2, D2 vdW radius (Bohr)}
 * (param_stride 0: one set for the batch); s6, s8, a1, a2: Parameters/d3.py (PBE0: 1.0, 0.7875, 0.4289, 4.4407).
 * Workspace: mop_fischer_workspace_bytes. */
int mop_fischer_d3old_hessian(int B, int natoms, const double* xyz, const double* atom_params, int param_stride,
                              double s6, double s8, double a1, double a2, double* H_out, int32_t* counts_out,
                              int32_t* status, void* work, size_t work_bytes, void* stream);

/* FischerD3ApproxHessian.main (ModelHessian/fischerd3.py:186-304; `fischerd3`, the AutoTS default, test/config.json):
 * as above with the table connectivity (factor 1.1) for the torsion bond count and the non-bonded mask, cut-offs 0.1 /
 * 1e-3, and C6 scaled per atom by clip(1 - 0.05 (CN - CN_ref), 0.75, 1.25) with the fractional coordination number of
 * :47-62.  atom_params [B or 1][natoms][5] = {..., reference coordination number}. */
int mop_fischer_d3_hessian(int B, int natoms, const double* xyz, const double* atom_params, int param_stride,
                           double s6, double s8, double a1, double a2, double* H_out, int32_t* counts_out,
                           int32_t* status, void* work, size_t work_bytes, void* stream);

/* ---- model Hessian modifiers (ModelHessian/approx_hessian.py:95-110) ---------------------------------------------
 * "ts": TransitionStateHessian.create_ts_hessian (ModelHessian/tshess.py:14-40) - unless an eigenvalue < -1e-8 exists,
 * reflect through the lowest mode with |lambda| >= 1e-8: out = sym((1 - 2 v v^T) H).  "clip": eigenvalue smoothing
 * (approx_hessian.py:103-126): out = V diag(s(lambda)) V^T, s(x) = sign(x) (2 - |x|^-0.1) for |x| >= 1.
 * evals [B][n], evecs [B][n][n] (rows = eigenvectors): the output of mop_eigh on the same Hessians.  out must not
 * alias the inputs; modified (optional) [B] = 1 where the reflection was applied. */
int mop_hessian_ts_modify(int B, int n, const double* H, const double* evals, const double* evecs, double* out,
                          int32_t* modified, void* stream);
int mop_hessian_clip_eigvals(int B, int n, const double* evals, const double* evecs, double* out, void* stream);

/* "sr": ShortRangeCorrectionHessian.main (ModelHessian/shortrange.py:9-346, approx_hessian.py:100-102): the second
 * derivatives of the short-range Coulomb kernel (1 - erf(omega r)) / r over the NON-bonded atom pairs (bonded:
 * distance <= 1.1 (R_i + R_j), BondConnectivity) within 15 Bohr, weighted q_i q_j cx_sr scaling_factor, TR/ROT-
 * projected, added to H and symmetrised: out = sym(H + P C P).  radii = covalent radii in Bohr, charges =
 * 0.2 (mean Pauling electronegativity - electronegativity) per atom ([natoms] with stride 0 or [B][natoms]);
 * defaults of the reference: omega 0.2, cx_sr 0.78, scaling_factor 0.5. */
size_t mop_hessian_sr_workspace_bytes(int B, int natoms);
int mop_hessian_sr_correction(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                              const double* charges, int charges_stride, double omega, double cx_sr,
                              double scaling_factor, const double* H, double* out, void* work, size_t work_bytes,
                              void* stream);

/* ---- effective Hessian for fixed atoms ---------------------------------------------------------------------------
 * Replaces HessianManager.calc_eff_hess_for_fix_atoms_and_set_hess (optimization.py:1325-1343,1358-1362):
 * H -= H[:, f] pinv(H[f, f] + 1e-10 I) H[f, :] with f the 3 n_fix coordinates of force_data["fix_atoms"], applied by the
 * caller to the bias Hessian and to the model Hessian.  mop_fix_atoms_gather writes the regularised blocks
 * [B][m][m]; mop_eigh diagonalises them; mop_fix_atoms_schur forms the pseudo-inverse (numpy.linalg.pinv cut-off,
 * 1e-15 max|lambda|) and updates H in place. */
int mop_fix_atoms_gather(int B, int n, int m, const int32_t* fix_coords, const double* H, double* blocks, void* stream);
int mop_fix_atoms_schur(int B, int n, int m, const int32_t* fix_coords, const double* evals, const double* evecs,
                        double* H, void* stream);

/* ---- restraint bias potentials --------------------------------------------------------------
 * Replaces calc_energy + torch.func.jacrev / hessian (Potential/potential.py:127-137) for StructKeepPotential
 * (kind 1), StructKeepPotentialv2 (kind 2; Potential/keep_potential.py), StructKeepAnglePotential (kind 3;
 * Potential/keep_angle_potential.py:7-229) and StructKeepDihedralAnglePotential (kind 4;
 * Potential/keep_dihedral_angle_potential.py:6-154).  terms: device array of nterm records of mop_bias_term_bytes()
 * bytes each: int32 kind, n1, n2, atoms[64] (0-based; kind 1: atoms i, j with n1 = n2 = 1; kind 2: the two
 * fragments back to back; kind 3: atoms i, j, k with n1 = 3; kind 4: atoms i, j, k, l with n1 = 4), then double
 * k (spring constant) and p (distance in Angstrom; angle in degrees for kind 3; phi0 in RADIANS for kind 4 - the
 * reference converts it in float32 or float64 depending on the caller, the host reproduces that).  E [B], grad [B][n], hess [B][n][n] are ADDED to (any may be NULL). */
size_t mop_bias_term_bytes(void);
int mop_bias_terms(int B, int natoms, int nterm, const void* terms, const double* xyz, double* E, double* grad,
                   double* hess, void* stream);

/* ---- step post-processing (either side of the optimizer step) -----------------------------
 * mop_kabsch replaces Calculationtools.kabsch_algorithm (Utils/calc_tools.py:412-425): P, Q [B][natoms][3];
 * P_aligned = P centred and rotated onto Q, Q_centred (optional) = Q minus its centroid (the reference
 * mutates both arguments in place).  status[b] = 1 for collinear structures (rotation undefined, identity used).
 * mop_check_convergence replaces ConvergenceChecker.check_convergence (optimization.py:1244-1289) without
 * the optimizer-instance override: grad, disp [B][n]; out (optional) [B][8] = {converged, max displacement
 * threshold, rms displacement threshold, max |g|, rms g, max |d|, rms d, 0}; converged (optional) [B]. */
int mop_kabsch(int B, int natoms, const double* P, const double* Q, double* P_aligned, double* Q_centred,
               int32_t* status, void* stream);
int mop_check_convergence(int B, int n, const double* grad, const double* disp, double max_force_thr,
                          double rms_force_thr, double max_disp_thr, double rms_disp_thr, double* out,
                          int32_t* converged, void* stream);

/* ---- redundant internal coordinates ------------------------------------------------------
 * Replaces Coordinate/redundant_coordinate.py: RedundantInternalCoordinates.B_matrix (:15-43),
 * RICgrad2cartgrad (:47-50), RIChess2carthess (:63-146), partial_stretch_B_matirx /
 * partial_bend_B_matrix / partial_torsion_B_matrix (:150-320), calc_int_grad_from_pBmat /
 * calc_cart_grad_from_pBmat with calc_inv_B_mat / calc_inv_G_mat (:377-439).
 * M = natoms (natoms - 1) / 2 atom pairs in itertools.combinations order, n = 3 natoms.
 *  mop_ric_bmatrix       Bmat [B][M][n]
 *  mop_ric_partial_rows  labels [nrows][4] int32, 1-based atom labels as in the reference, 0 = unused
 *                        (2 / 3 / 4 labels = stretch / bend / torsion); rows_out [B][nrows][n]
 *  mop_ric_grad_to_cart  cart_grad [B][n] = B^T ric_grad [B][M]
 *  mop_ric_hess_to_cart  cart_hess = B^T H B + K; ric_hess [B][M][M], or [B][M] when diagonal != 0;
 *                        K [B][n][n] or NULL
 *  mop_ric_kmatrix       K [B][n][n] = sum_t ric_grad[t] d2 q_t / dx2, t running over the bond, angle and
 *                        dihedral tables in that order (0-based atom indices, mop_connectivity layout;
 *                        tables_per_structure = 0: one set of tables shared by the batch);
 *                        ric_grad [B][ric_len], terms t >= ric_len are skipped.  The reference obtains
 *                        ric_grad from a singular solve (cartgrad2RICgrad): it is an input here.
 *  mop_ric_pb_int_grad   int_grad [B][m] = (G^+ pB^T)^T cart_grad, G = pB^T pB, pB [B][m][n]; the
 *                        pseudo-inverse keeps singular values <= 1e-6 as they are, as the reference does
 *  mop_ric_pb_cart_grad  cart_grad [B][n] = pB^T int_grad */
int mop_ric_bmatrix(int B, int natoms, const double* xyz, double* Bmat, void* stream);
int mop_ric_partial_rows(int B, int natoms, const double* xyz, int nrows, const int32_t* labels, double* rows_out,
                         void* stream);
int mop_ric_grad_to_cart(int B, int natoms, const double* xyz, const double* ric_grad, double* cart_grad, void* stream);
size_t mop_ric_hess_workspace_bytes(int B, int natoms, int diagonal);
int mop_ric_hess_to_cart(int B, int natoms, const double* xyz, const double* ric_hess, int diagonal, const double* K,
                         double* cart_hess, void* work, size_t work_bytes, void* stream);
int mop_ric_kmatrix(int B, int natoms, const double* xyz, const int32_t* bonds, const int32_t* angles,
                    const int32_t* dihedrals, const int32_t* counts, int cap_bonds, int cap_angles, int cap_dihedrals,
                    int tables_per_structure, const double* ric_grad, int ric_len, double* K_out, void* stream);
size_t mop_ric_pb_workspace_bytes(int B, int n);
int mop_ric_pb_int_grad(int B, int n, int m, const double* pB, const double* cart_grad, double* int_grad,
                        int32_t* status, void* work, size_t work_bytes, void* stream);
int mop_ric_pb_cart_grad(int B, int n, int m, const double* pB, const double* int_grad, double* cart_grad, void* stream);

/* ---- Swart model Hessian ----------------------------------------------------------
 * Replaces SwartApproxHessian.main (ModelHessian/swart.py:317-355): all-pairs screened stretch
 * terms, screened bend terms with the near-linear blending, non-finite fallback to stretches
 * only (status[b] = 1 when taken), then the TR/ROT projection.  radii [B or 1][natoms] = the
 * model's own table in Bohr (swart.py:10-25; 1.0 for unknown elements); radii_stride 0 shares one
 * row.  Hraw_out (optional) [B][n][n] receives the unprojected Hessian; when NULL the workspace
 * holds it. */
size_t mop_swart_workspace_bytes(int B, int natoms);
int mop_swart_hessian(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                      double* H_out, double* Hraw_out, int32_t* status, void* work, size_t work_bytes,
                      void* stream);

/* ---- (3b') Lindh model Hessian ---------------------------------------------------
 * Replaces LindhApproxHessian.main (ModelHessian/lindh.py:145-165) up to its K term:
 * H_out = project(B^T diag(k) B) with the all-pairs distance B matrix
 * (Coordinate/redundant_coordinate.py:15-43) and the diagonal force constants of
 * guess_lindh_hessian (lindh.py:79-143); kdiag_out (optional) returns them,
 * [B][N(N-1)/2] in itertools.combinations order.  The reference's K term multiplies
 * second derivatives by an internal gradient obtained from a singular solve and is
 * ill-posed in the reference itself; it is not added.
 * atom_params [B or 1][natoms][6] = {covalent radius (Bohr), period index 0/1/2
 * (H-He / Li-Ne / other), atomic mass, UFF VDW distance, UFF well depth, UFF charge}. */
size_t mop_lindh_workspace_bytes(int B, int natoms);
int mop_lindh_hessian(int B, int natoms, const double* xyz, const double* atom_params, int param_stride,
                      double* H_out, double* kdiag_out, int32_t* counts_out, int32_t* status, void* work,
                      size_t work_bytes, void* stream);

/* ---- (3c) AFIR bias potential ---------------------------------------------------
 * Replaces AFIRPotential.calc_energy (Potential/AFIR_potential.py:18-55) and the
 * torch.func.jacrev / hessian calls of BiasPotentialCalculation.main
 * (Potential/potential.py:127-152) for AFIR terms: E [B], grad [B][3N], H [B][3N][3N]
 * (any of them may be NULL).  frag1 / frag2: 0-based atom indices; radii_f32: covalent
 * radii in Bohr ROUNDED TO FLOAT32 (the reference builds float32 tensors); gamma [B] kJ/mol. */
int mop_afir(int B, int natoms, const double* xyz, int n1, const int32_t* frag1, int n2,
             const int32_t* frag2, const float* radii_f32, const double* gamma, double* E,
             double* grad, double* H, void* stream);

/* ---- (5) NEB: tangent projection, Ayala curvature update, step limits -----------------
 * Images [first, first + nloc) of an nimg-image chain live on this GPU.  *_halo arrays have
 * nloc + 2 entries: slot l + 1 = local image l, slots 0 and nloc + 1 = images first - 1 and
 * first + nloc (received over NCCL when the chain is sharded; ignored at the chain ends).
 * mop_bneb_force  replaces CaluculationBNEB.calc_force (MEP/pathopt_bneb_force.py:33-117):
 *                 force [nloc][n] = -(g + projection), tau [nloc][n] = projection (get_tau).
 * mop_neb_ayala   replaces calculate_gamma (pathopt_bneb_force.py:161-222) and the rank-1
 *                 update H += gamma t t^T of RFOOptimizer (Optimizer/rfo_neb.py:43-73).
 * mop_neb_limit_tr replaces _limit_step_size (rfo_neb.py:76-83; only when apply_step_limit != 0) and
 *                 TR_NEB.TR_calc (Optimizer/trust_radius_neb.py:17-98); delta [nloc][n] in place. */
int mop_bneb_force(int nimg, int first, int nloc, int n, const double* x_halo, const double* E_halo,
                   const double* g, double* force, double* tau, void* stream);
int mop_neb_ayala(int nimg, int first, int nloc, int n, const double* x_halo, const double* E_halo,
                  const double* g_halo, const double* tau, double* H, double* gamma_out, void* stream);
int mop_neb_limit_tr(int nimg, int first, int nloc, int n, int fix_init_edge, int fix_end_edge,
                     int apply_step_limit, const double* x_halo, const double* g, double* delta, void* stream);
/* Image redistribution at equal arc length: distribute_geometry (Interpolation/linear_interpolation.py:308-336) on
 * the path lengths of calc_path_length_list (Utils/calc_tools.py:853-862) - the `align_distances` strategy of
 * NEB._align_geometries (neb.py:649-760).  x_chain [nimg][natoms][3] is the WHOLE chain (a sharded chain is
 * all-gathered first: neb_halo.gather_chain), x_out [nloc][natoms][3] receives the new images first .. first+nloc-1,
 * path_length_out [nimg] (or null) the running path length.  x_out must not alias x_chain. */
int mop_neb_redistribute(int nimg, int natoms, int first, int nloc, const double* x_chain, double* x_out,
                         double* path_length_out, void* stream);
/* FIRE optimizer of the NEB driver (Optimizer/fire_neb.py:38-92).  mop_neb_fire_blend: per-atom velocity / force
 * blend vneb_out [nloc][natoms][3] and the power sum v_prev . F accumulated into the device scalar power_accum
 * (zeroed by the caller, all-reduced over ranks when the chain is sharded; prev_velocity = NULL on the first
 * iteration).  The (dt, a, n_reset) schedule is scalar host logic (host mirror FIREOptimizer).
 * mop_neb_fire_advance: velocity_out = (reset ? 0 : vneb) + dt F, delta_out = dt (velocity_out + prev_velocity)
 * or dt velocity_out; mop_neb_limit_tr(apply_step_limit = 0) then applies TR_calc. */
int mop_neb_fire_blend(int nloc, int natoms, double a, const double* force, const double* velocity,
                       const double* prev_velocity, double* vneb_out, double* power_accum, void* stream);
int mop_neb_fire_advance(int nloc, int n, double dt, int reset, const double* vneb, const double* force,
                         const double* prev_velocity, double* velocity_out, double* delta_out, void* stream);

/* ---- caller side: composite outer trust radius -----------------------------------
 * Replaces TrustRadius.update_trust_radii (Optimizer/trust_radius.py:120-206) as called
 * from CalculateMoveVector.update_trust_radius_conditionally (optimizer.py:534-553):
 * r = (pre_Be - Be) / (pre_Bg.pre_move + 1/2 pre_move^T (H + Hbias) pre_move), adaptive
 * factor from the ratio history, clip to [trust_min, trust_max].  trust [B] in/out;
 * state [B][MOP_TR_STATE] doubles, zero-initialised by the caller. */
#define MOP_TR_STATE 12
int mop_outer_trust_radius(int B, int n, const double* H, const double* Hbias, const double* pre_Bg,
                           const double* pre_move, const double* Be, const double* pre_Be,
                           double* trust, double* state, double trust_min, double trust_max,
                           void* stream);

/* ---- caller side: CalculateMoveVector.calc_move_vector clamp -------------
 * Replaces optimizer.py:792-798,812: scale move to trust_outer[B] if longer,
 * x_new_ang = (x - move) * 0.52917721067. */
int mop_clamp_and_move(int B, int n, const double* x, double* move, const double* trust_outer,
                       double* x_new_ang, void* stream);

/* Measurement probes and tuning / diagnostic hooks are NOT part of this interface: they are declared in
 * multioptpy_b200/csrc/mop_private.h (mop_priv_*), used by bench.py and tools/ only. */

#ifdef __cplusplus
}
#endif
#endif /* MOP_B200_H */
