// Step post-processing either side of the optimizer step:
//   k_kabsch       Calculationtools.kabsch_algorithm (Utils/calc_tools.py:412-425): centre both
//                  geometries, rotate P onto Q.  The reference takes R = V U^T from numpy's SVD of
//                  H = P^T Q and flips the last right-singular vector when det R < 0; that rotation is
//                  v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T for either sign of det H, with (u_k, v_k) the
//                  two leading singular pairs - obtained here from a 3 x 3 Jacobi eigensolve of H H^T.
//                  (Needs sigma_2 > 0, i.e. non-collinear structures; collinear input is flagged.)
//   k_convergence  ConvergenceChecker.check_convergence (optimization.py:1244-1289): max / filtered-rms
//                  of gradient and displacement against the force-relaxed displacement thresholds.
// One warp per structure; HBM-bound streaming of 2 x 24 N bytes in, 2 x 24 N out.
#include "common.cuh"

namespace mop {

__device__ void jacobi3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double dia = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * dia) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // columns
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // rows
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

__global__ void __launch_bounds__(128) k_kabsch(int B, int N, const double* __restrict__ P_all, const double* __restrict__ Q_all,
                                                double* __restrict__ Pout, double* __restrict__ Qout, int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const double* P = P_all + (size_t)b * 3 * N;
  const double* Q = Q_all + (size_t)b * 3 * N;
  double cp[3] = {0, 0, 0}, cq[3] = {0, 0, 0};
  for (int a = lane; a < N; a += 32)
    for (int c = 0; c < 3; ++c) {
      cp[c] += P[3 * a + c];
      cq[c] += Q[3 * a + c];
    }
  for (int c = 0; c < 3; ++c) {
    cp[c] = warp_sum(cp[c]) / N;
    cq[c] = warp_sum(cq[c]) / N;
  }
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // H = P^T Q (centred)
  for (int a = lane; a < N; a += 32) {
    double p[3], q[3];
    for (int c = 0; c < 3; ++c) {
      p[c] = P[3 * a + c] - cp[c];
      q[c] = Q[3 * a + c] - cq[c];
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) H[i][j] = fma(p[i], q[j], H[i][j]);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = warp_sum(H[i][j]);
  // leading singular pairs: H H^T = U S^2 U^T
  double A[3][3], U[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = H[i][0] * H[j][0] + H[i][1] * H[j][1] + H[i][2] * H[j][2];
  jacobi3(A, U, w);
  int o0 = 0, o1 = 1, o2 = 2;  // descending order of w
  if (w[o0] < w[o1]) { const int t = o0; o0 = o1; o1 = t; }
  if (w[o1] < w[o2]) { const int t = o1; o1 = o2; o2 = t; }
  if (w[o0] < w[o1]) { const int t = o0; o0 = o1; o1 = t; }
  double u1[3], u2[3], v1[3], v2[3], u3[3], v3[3];
  for (int i = 0; i < 3; ++i) {
    u1[i] = U[i][o0];
    u2[i] = U[i][o1];
  }
  double n1 = 0.0, n2 = 0.0;
  for (int j = 0; j < 3; ++j) {  // v = H^T u
    v1[j] = H[0][j] * u1[0] + H[1][j] * u1[1] + H[2][j] * u1[2];
    v2[j] = H[0][j] * u2[0] + H[1][j] * u2[1] + H[2][j] * u2[2];
    n1 = fma(v1[j], v1[j], n1);
    n2 = fma(v2[j], v2[j], n2);
  }
  n1 = sqrt(n1);
  n2 = sqrt(n2);
  const bool degenerate = !(n2 > 1e-12 * fmax(n1, 1e-300));
  double R[3][3];
  if (degenerate) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] = i == j ? 1.0 : 0.0;
  } else {
    for (int j = 0; j < 3; ++j) v1[j] /= n1;
    // re-orthogonalise v2 against v1 (H^T u2 carries the rounding of the eigenvectors)
    double d12 = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2];
    double nn = 0.0;
    for (int j = 0; j < 3; ++j) {
      v2[j] -= d12 * v1[j];
      nn = fma(v2[j], v2[j], nn);
    }
    nn = sqrt(nn);
    for (int j = 0; j < 3; ++j) v2[j] /= nn;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R[i][j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
  }
  if (lane == 0 && status) status[b] = degenerate ? 1 : 0;
  for (int a = lane; a < N; a += 32) {
    double p[3];
    for (int c = 0; c < 3; ++c) p[c] = P[3 * a + c] - cp[c];
    for (int i = 0; i < 3; ++i) Pout[(size_t)b * 3 * N + 3 * a + i] = R[i][0] * p[0] + R[i][1] * p[1] + R[i][2] * p[2];
    if (Qout)
      for (int c = 0; c < 3; ++c) Qout[(size_t)b * 3 * N + 3 * a + c] = Q[3 * a + c] - cq[c];
  }
}

// out[b][0..5] = converged (0/1), max_displacement_threshold, rms_displacement_threshold, max_force, rms_force,
//                max_displacement; out[b][6] = rms_displacement
__global__ void __launch_bounds__(128) k_convergence(int B, int n, const double* __restrict__ g_all, const double* __restrict__ d_all,
                                                     double tmf, double trf, double tmd, double trd, double* __restrict__ out,
                                                     int32_t* __restrict__ converged) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const double* g = g_all + (size_t)b * n;
  const double* d = d_all + (size_t)b * n;
  double mg = 0.0, sg = 0.0, cg = 0.0, md = 0.0, sd = 0.0, cd = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double x = g[i], y = d[i];
    mg = fmax(mg, fabs(x));
    md = fmax(md, fabs(y));
    if (fabs(x) > 1e-10) { sg = fma(x, x, sg); cg += 1.0; }
    if (fabs(y) > 1e-10) { sd = fma(y, y, sd); cd += 1.0; }
  }
  mg = warp_max(mg); md = warp_max(md);
  sg = warp_sum(sg); cg = warp_sum(cg); sd = warp_sum(sd); cd = warp_sum(cd);
  const double rg = cg > 0.0 ? sqrt(sg / cg) : 0.0, rd = cd > 0.0 ? sqrt(sd / cd) : 0.0;
  const double mdt = fmax(tmd, tmd + fmax(0.0, tmf - mg));
  const double rdt = fmax(trd, trd + fmax(0.0, trf - rg));
  const int ok = mg < tmf && rg < trf && md < mdt && rd < rdt;
  if (lane == 0) {
    if (converged) converged[b] = ok;
    if (out) {
      double* o = out + (size_t)b * 8;
      o[0] = ok; o[1] = mdt; o[2] = rdt; o[3] = mg; o[4] = rg; o[5] = md; o[6] = rd; o[7] = 0.0;
    }
  }
}

}  // namespace mop

extern "C" int mop_kabsch(int B, int natoms, const double* P, const double* Q, double* P_aligned, double* Q_centred,
                          int32_t* status, void* stream) {
  MOP_REQUIRE(B >= 0 && natoms > 0 && P && Q && P_aligned, "mop_kabsch: bad arguments");
  if (B == 0) return MOP_OK;
  mop::k_kabsch<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, natoms, P, Q, P_aligned, Q_centred, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}

extern "C" int mop_check_convergence(int B, int n, const double* grad, const double* disp, double max_force_thr,
                                     double rms_force_thr, double max_disp_thr, double rms_disp_thr, double* out,
                                     int32_t* converged, void* stream) {
  MOP_REQUIRE(B >= 0 && n > 0 && grad && disp && (out || converged), "mop_check_convergence: bad arguments");
  if (B == 0) return MOP_OK;
  mop::k_convergence<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, n, grad, disp, max_force_thr, rms_force_thr,
                                                                   max_disp_thr, rms_disp_thr, out, converged);
  MOP_CHECK_CUDA(cudaGetLastError());
  return MOP_OK;
}
