// Swart model Hessian: SwartApproxHessian.main, ModelHessian/swart.py:317-355.
//
// One CTA per structure.  Screening s_ij = exp(1 - r_ij / (R_i + R_j)) (:64-82) is evaluated on
// the fly; all-pairs stretch terms 0.35 s^3 e e^T (:83-107) are spread over the threads, the
// angle terms (:192-315) are enumerated per centre atom by one warp each: the warp compacts the
// centre's neighbour list (s >= eps2) in ascending index order, its lanes walk the i < k pairs.
// Contributions are accumulated with FP64 atomics into a shared-memory image of the Hessian when
// 8 n^2 bytes fit (n <= 156), else straight into global memory, then written out once, coalesced.
// The TR/ROT projection (a5) follows as a second kernel.  HBM-bound: algorithmic traffic is the
// 8 n^2-byte raw Hessian written once plus the projection's read + write.
#include "connectivity.cuh"

namespace mop {

constexpr int SW_THREADS = 512;
constexpr int SW_WARPS = SW_THREADS / 32;

struct SwartConst {
  double f = 0.12, tolth = 0.2, eps1 = 0.3 * 0.3, eps2 = 0.03310914970542981;  // wthr^2, wthr^2 / e
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

__device__ __forceinline__ double swart_screen(const double* xyz, const double* rad, int i, int j, double* dist) {
  double d = fmax(np_dist(xyz + 3 * i, xyz + 3 * j), 1e-8);
  const double cs = fmax(rad[i] + rad[j], 1e-8);
  *dist = d;
  return exp(1.0 - d / cs);
}

// upper triangle of H += h * u u^T over the atoms of one term (the kernel mirrors it at the end: the
// accumulation is atomic-throughput bound, the symmetric half would cost 1.8x the atomics)
__device__ __forceinline__ void add_outer(double* H, int n, const int* at, int nat, const double* u, double h) {
  for (int p = 0; p < 3 * nat; ++p) {
    const int gp = 3 * at[p / 3] + p % 3;
    const double hp = h * u[p];
    for (int q = 0; q < 3 * nat; ++q) {
      const int gq = 3 * at[q / 3] + q % 3;
      if (gq >= gp) atomicAdd(&H[(size_t)gp * n + gq], hp * u[q]);
    }
  }
}

__device__ void swart_bonds(int N, const double* xyz, const double* rad, double* H) {
  const int n = 3 * N;
  for (int e = threadIdx.x; e < N * N; e += SW_THREADS) {
    const int i = e / N, j = e - i * N;
    if (i >= j) continue;
    double d;
    const double s = swart_screen(xyz, rad, i, j, &d);
    double u[6];
    for (int c = 0; c < 3; ++c) {
      u[c] = (xyz[3 * i + c] - xyz[3 * j + c]) / d;
      u[3 + c] = -u[c];
    }
    const int at[2] = {i, j};
    add_outer(H, n, at, 2, u, 0.35 * (s * s * s));
  }
}

// one (i, j, k) angle term, i < k, centre j (swart.py:226-315)
__device__ void swart_angle(int N, const double* xyz, int i, int j, int k, double l1, double l2, double ss,
                            double* H) {
  const SwartConst C;
  double v1[3], v2[3], n1[3], n2[3];
  for (int c = 0; c < 3; ++c) {
    v1[c] = xyz[3 * i + c] - xyz[3 * j + c];
    v2[c] = xyz[3 * k + c] - xyz[3 * j + c];
    n1[c] = v1[c] / l1;
    n2[c] = v2[c] / l2;
  }
  double cs = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
  cs = fmin(fmax(cs, -1.0), 1.0);
  const double s2 = fmax(1e-12, 1.0 - cs * cs);
  const double sn = sqrt(s2);
  const double den = fmax(sn, 1e-6);
  double bn[9];
  for (int c = 0; c < 3; ++c) {
    bn[c] = (cs * n1[c] - n2[c]) / (l1 * den);
    bn[6 + c] = (cs * n2[c] - n1[c]) / (l2 * den);
    bn[3 + c] = -(bn[c] + bn[6 + c]);
  }
  const double w = C.f + (1.0 - C.f) * sn;
  const double hb = 0.075 * (ss * ss) * (w * w);
  const double th1 = cs > 1.0 - C.tolth ? 1.0 - cs : 1.0 + cs;
  const int at[3] = {i, j, k};
  const int n = 3 * N;
  if (!(th1 < C.tolth)) {
    add_outer(H, n, at, 3, bn, hb);
    return;
  }
  const double q = th1 / C.tolth;
  const double sl = (1.0 - q * q) * (1.0 - q * q);
  if (!(cs > 1.0 - C.tolth)) {
    double bs[9];
    for (int c = 0; c < 9; ++c) bs[c] = (1.0 - sl) * bn[c];
    add_outer(H, n, at, 3, bs, hb);
    return;
  }
  // linear-bend pair (swart.py:135-190)
  double vn[3];
  cross3(v1, v2, vn);
  double nvn = norm3(vn);
  if (nvn < 1e-12) {
    const double sc1 = v1[0] / (l1 * l1);
    double cand[3] = {1.0 - sc1 * v1[0], -sc1 * v1[1], -sc1 * v1[2]};
    double cn = norm3(cand);
    if (!(cn >= 1e-12)) {
      const double sc2 = v1[1] / (l1 * l1);
      cand[0] = -sc2 * v1[0]; cand[1] = 1.0 - sc2 * v1[1]; cand[2] = -sc2 * v1[2];
      cn = fmax(norm3(cand), 1e-12);
    }
    vn[0] = cand[0]; vn[1] = cand[1]; vn[2] = cand[2];
    nvn = cn;
  }
  nvn = fmax(nvn, 1e-12);
  double vd[3], vn2[3];
  for (int c = 0; c < 3; ++c) {
    vn[c] /= nvn;
    vd[c] = v1[c] - v2[c];
  }
  cross3(vd, vn, vn2);
  const double n2n = fmax(norm3(vn2), 1e-12);
  double bp[9], bc[9];
  for (int c = 0; c < 3; ++c) {
    const double t = vn2[c] / n2n;
    bp[c] = vn[c] / l1;
    bp[6 + c] = vn[c] / l2;
    bp[3 + c] = -bp[c] - bp[6 + c];
    const double l0 = t / l1, l6 = t / l2;
    bc[c] = sl * l0 + (1.0 - sl) * bn[c];
    bc[6 + c] = sl * l6 + (1.0 - sl) * bn[6 + c];
    bc[3 + c] = sl * (-l0 - l6) + (1.0 - sl) * bn[3 + c];
  }
  add_outer(H, n, at, 3, bp, hb);
  add_outer(H, n, at, 3, bc, hb);
}

__global__ void __launch_bounds__(SW_THREADS, 1)
k_swart(int N, int in_smem, const double* __restrict__ xyz_all, const double* __restrict__ rad_all, int rad_stride,
        double* __restrict__ H_all, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const SwartConst C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = 3 * N;
  double* xyz = sm;                       // 3N
  double* rad = xyz + 3 * N;              // N
  double* nbs = rad + N;                  // SW_WARPS x N neighbour screens
  double* nbd = nbs + SW_WARPS * N;       // SW_WARPS x N neighbour distances
  int* nbi = (int*)(nbd + SW_WARPS * N);  // SW_WARPS x N neighbour indices
  double* Hs = (double*)(nbi + SW_WARPS * N);  // 16 N ints: 8-byte aligned
  double* Hg = H_all + (size_t)b * n * n;
  double* H = in_smem ? Hs : Hg;
  __shared__ int s_bad;
  for (int i = tid; i < 3 * N; i += SW_THREADS) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  for (int i = tid; i < N; i += SW_THREADS) rad[i] = rad_all[(size_t)b * rad_stride + i];
  for (int e = tid; e < n * n; e += SW_THREADS) H[e] = 0.0;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  swart_bonds(N, xyz, rad, H);
  // angles: one warp per centre
  double* ws = nbs + wid * N;
  double* wd = nbd + wid * N;
  int* wi = nbi + wid * N;
  for (int j = wid; j < N; j += SW_WARPS) {
    int cnt = 0;
    for (int i0 = 0; i0 < N; i0 += 32) {
      const int i = i0 + lane;
      double d = 0.0, s = 0.0;
      if (i < N && i != j) s = swart_screen(xyz, rad, i, j, &d);
      const bool ok = i < N && i != j && s >= C.eps2;
      const unsigned m = __ballot_sync(MOP_FULL_MASK, ok);
      if (ok) {
        const int slot = cnt + __popc(m & ((1u << lane) - 1u));
        wi[slot] = i; ws[slot] = s; wd[slot] = d;
      }
      cnt += __popc(m);
    }
    __syncwarp();
    const int npair = cnt * (cnt - 1) / 2;
    for (int t = lane; t < npair; t += 32) {
      // t -> (a < c): row a holds pairs (a, a+1..cnt-1)
      int a = (int)((2.0 * cnt - 1.0 - sqrt((2.0 * cnt - 1.0) * (2.0 * cnt - 1.0) - 8.0 * t)) * 0.5);
      while (a > 0 && a * (2 * cnt - a - 1) / 2 > t) --a;
      while ((a + 1) * (2 * cnt - a - 2) / 2 <= t) ++a;
      const int c = a + 1 + (t - a * (2 * cnt - a - 1) / 2);
      const double ss = ws[a] * ws[c];
      if (ss < C.eps1 || !(wd[a] > 1e-8 && wd[c] > 1e-8)) continue;
      swart_angle(N, xyz, wi[a], j, wi[c], wd[a], wd[c], ss, H);
    }
    __syncwarp();
  }
  __syncthreads();
  // NaN / Inf fallback to the stretch-only Hessian (swart.py:340-350)
  int bad = 0;
  for (int e = tid; e < n * n; e += SW_THREADS) bad |= !isfinite(H[e]);
  if (bad) s_bad = 1;
  __syncthreads();
  if (s_bad) {
    for (int e = tid; e < n * n; e += SW_THREADS) H[e] = 0.0;
    __syncthreads();
    swart_bonds(N, xyz, rad, H);
    __syncthreads();
  }
  if (tid == 0 && status) status[b] = s_bad;
  // mirror the accumulated upper triangle and write out
  for (int e = tid; e < n * n; e += SW_THREADS) {
    const int r = e / n, c = e - r * n;
    Hg[e] = (c >= r) ? H[e] : H[(size_t)c * n + r];
  }
}


// ------------------------------------------------------------------------------------------------
// Gather formulation (natoms <= 100): no atomics.  Every 3 x 3 atom-pair block of the raw Hessian is
// owned by one thread (off-diagonal blocks) or one warp (diagonal blocks), which sums the stretch term
// and every bend term that touches both atoms, re-evaluating the bend's Wilson vectors where needed
// (a bend (i, j, k) feeds six blocks, so it is evaluated six times: ~1e7 flops per structure at
// N = 50 against ~6e5 FP64 shared-memory atomics, which cost ~30 cycles each in the scatter kernel).
struct SwartBend {
  int nvec;        // 1 or 2 Wilson vectors
  double hb;       // force constant
  double U[2][9];  // (i, j, k) components
};

// bend (i, j, k), i < k, centre j; l1 = |x_i - x_j|, l2 = |x_k - x_j|, ss = s_ij s_jk (swart.py:226-315)
__device__ __forceinline__ void swart_bend_eval(const double* xyz, int i, int j, int k, double l1, double l2, double ss,
                                                SwartBend& o) {
  const SwartConst C;
  // reciprocals once (FP64 division is ~20 instructions): differs from the reference's divisions by
  // an ulp, far below the 1e-10 parity bar
  const double il1 = 1.0 / l1, il2 = 1.0 / l2;
  double v1[3], v2[3], n1[3], n2[3];
  for (int c = 0; c < 3; ++c) {
    v1[c] = xyz[3 * i + c] - xyz[3 * j + c];
    v2[c] = xyz[3 * k + c] - xyz[3 * j + c];
    n1[c] = v1[c] * il1;
    n2[c] = v2[c] * il2;
  }
  double cs = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
  cs = fmin(fmax(cs, -1.0), 1.0);
  const double s2 = fmax(1e-12, 1.0 - cs * cs);
  const double sn = sqrt(s2);
  const double iden = 1.0 / fmax(sn, 1e-6);
  const double f1 = il1 * iden, f2 = il2 * iden;
  double bn[9];
  for (int c = 0; c < 3; ++c) {
    bn[c] = (cs * n1[c] - n2[c]) * f1;
    bn[6 + c] = (cs * n2[c] - n1[c]) * f2;
    bn[3 + c] = -(bn[c] + bn[6 + c]);
  }
  const double w = C.f + (1.0 - C.f) * sn;
  o.hb = 0.075 * (ss * ss) * (w * w);
  o.nvec = 1;
  const double th1 = cs > 1.0 - C.tolth ? 1.0 - cs : 1.0 + cs;
  if (!(th1 < C.tolth)) {
    for (int c = 0; c < 9; ++c) o.U[0][c] = bn[c];
    return;
  }
  const double q = th1 / C.tolth;
  const double sl = (1.0 - q * q) * (1.0 - q * q);
  if (!(cs > 1.0 - C.tolth)) {
    for (int c = 0; c < 9; ++c) o.U[0][c] = (1.0 - sl) * bn[c];
    return;
  }
  double vn[3];
  cross3(v1, v2, vn);
  double nvn = norm3(vn);
  if (nvn < 1e-12) {
    const double sc1 = v1[0] / (l1 * l1);
    double cand[3] = {1.0 - sc1 * v1[0], -sc1 * v1[1], -sc1 * v1[2]};
    double cn = norm3(cand);
    if (!(cn >= 1e-12)) {
      const double sc2 = v1[1] / (l1 * l1);
      cand[0] = -sc2 * v1[0]; cand[1] = 1.0 - sc2 * v1[1]; cand[2] = -sc2 * v1[2];
      cn = fmax(norm3(cand), 1e-12);
    }
    vn[0] = cand[0]; vn[1] = cand[1]; vn[2] = cand[2];
    nvn = cn;
  }
  nvn = fmax(nvn, 1e-12);
  double vd[3], vn2[3];
  for (int c = 0; c < 3; ++c) {
    vn[c] /= nvn;
    vd[c] = v1[c] - v2[c];
  }
  cross3(vd, vn, vn2);
  const double n2n = fmax(norm3(vn2), 1e-12);
  o.nvec = 2;
  for (int c = 0; c < 3; ++c) {
    const double t = vn2[c] / n2n;
    o.U[0][c] = vn[c] * il1;
    o.U[0][6 + c] = vn[c] * il2;
    o.U[0][3 + c] = -o.U[0][c] - o.U[0][6 + c];
    const double l0 = t * il1, l6 = t * il2;
    o.U[1][c] = sl * l0 + (1.0 - sl) * bn[c];
    o.U[1][6 + c] = sl * l6 + (1.0 - sl) * bn[6 + c];
    o.U[1][3 + c] = sl * (-l0 - l6) + (1.0 - sl) * bn[3 + c];
  }
}

// blk += hb sum_v U_v[sa .. sa+2] U_v[sb .. sb+2]^T
__device__ __forceinline__ void swart_acc(double* blk, const SwartBend& o, int sa, int sb) {
  for (int v = 0; v < o.nvec; ++v)
    for (int p = 0; p < 3; ++p) {
      const double hp = o.hb * o.U[v][sa + p];
      for (int q = 0; q < 3; ++q) blk[3 * p + q] = fma(hp, o.U[v][sb + q], blk[3 * p + q]);
    }
}

__global__ void __launch_bounds__(SW_THREADS, 1)
k_swart_gather(int N, const double* __restrict__ xyz_all, const double* __restrict__ rad_all, int rad_stride,
               double* __restrict__ H_all, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const SwartConst C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = 3 * N;
  double* xyz = sm;          // 3N
  double* rad = xyz + 3 * N; // N
  double* D = rad + N;       // N x N distances (clamped)
  double* Sc = D + N * N;    // N x N screening, 0 on the diagonal
  __shared__ int s_bad;
  double* H = H_all + (size_t)b * n * n;
  for (int i = tid; i < 3 * N; i += SW_THREADS) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  for (int i = tid; i < N; i += SW_THREADS) rad[i] = rad_all[(size_t)b * rad_stride + i];
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (int e = tid; e < N * N; e += SW_THREADS) {
    const int i = e / N, j = e - i * N;
    if (i > j) continue;
    double d = 1.0, s = 0.0;
    if (i != j) s = swart_screen(xyz, rad, i, j, &d);
    D[i * N + j] = d; D[j * N + i] = d;
    Sc[i * N + j] = s; Sc[j * N + i] = s;
  }
  __syncthreads();
  // a bend (i, j, k) exists iff both ends are neighbours of the centre and the product screen passes
  auto pass = [&](int j, int i, int k, double* ss) -> bool {
    const double si = Sc[j * N + i], sk = Sc[j * N + k];
    if (!(si >= C.eps2 && sk >= C.eps2)) return false;
    *ss = si * sk;
    return *ss >= C.eps1 && D[i * N + j] > 1e-8 && D[k * N + j] > 1e-8;
  };
  for (int round = 0; round < 2; ++round) {
    const bool bends = round == 0;
    // ---- off-diagonal blocks (a < b): one thread each ----
    for (int e = tid; e < N * N; e += SW_THREADS) {
      const int a = e / N, bb = e - a * N;
      if (a >= bb) continue;
      double blk[9];
      {
        const double d = D[a * N + bb], s = Sc[a * N + bb], h = -0.35 * (s * s * s);
        double ev[3];
        for (int c = 0; c < 3; ++c) ev[c] = (xyz[3 * a + c] - xyz[3 * bb + c]) / d;
        for (int p = 0; p < 3; ++p)
          for (int q = 0; q < 3; ++q) blk[3 * p + q] = h * ev[p] * ev[q];
      }
      if (bends) {
        SwartBend o;
        double ss;
        for (int c = 0; c < N; ++c) {
          if (c == a || c == bb) continue;
          if (pass(c, a, bb, &ss)) {  // centre c, ends a < b
            swart_bend_eval(xyz, a, c, bb, D[a * N + c], D[bb * N + c], ss, o);
            swart_acc(blk, o, 0, 6);
          }
          {  // centre a, ends b and c
            const int i = bb < c ? bb : c, k = bb < c ? c : bb;
            if (pass(a, i, k, &ss)) {
              swart_bend_eval(xyz, i, a, k, D[i * N + a], D[k * N + a], ss, o);
              swart_acc(blk, o, 3, bb == i ? 0 : 6);
            }
          }
          {  // centre b, ends a and c
            const int i = a < c ? a : c, k = a < c ? c : a;
            if (pass(bb, i, k, &ss)) {
              swart_bend_eval(xyz, i, bb, k, D[i * N + bb], D[k * N + bb], ss, o);
              swart_acc(blk, o, a == i ? 0 : 6, 3);
            }
          }
        }
      }
      int bad = 0;
      for (int p = 0; p < 3; ++p)
        for (int q = 0; q < 3; ++q) {
          const double x = blk[3 * p + q];
          bad |= !isfinite(x);
          H[(size_t)(3 * a + p) * n + 3 * bb + q] = x;
          H[(size_t)(3 * bb + q) * n + 3 * a + p] = x;
        }
      if (bad) s_bad = 1;
    }
    // ---- diagonal blocks: one warp each, lanes split the third-atom loop ----
    for (int a = wid; a < N; a += SW_WARPS) {
      double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = lane; c < N; c += 32) {
        if (c == a) continue;
        const double d = D[a * N + c], s = Sc[a * N + c], h = 0.35 * (s * s * s);
        double ev[3];
        for (int p = 0; p < 3; ++p) ev[p] = (xyz[3 * a + p] - xyz[3 * c + p]) / d;
        for (int p = 0; p < 3; ++p)
          for (int q = 0; q < 3; ++q) blk[3 * p + q] = fma(h * ev[p], ev[q], blk[3 * p + q]);
        if (!bends) continue;
        SwartBend o;
        double ss;
        for (int c2 = 0; c2 < N; ++c2) {
          if (c2 == a || c2 == c) continue;
          if (c2 > c && pass(a, c, c2, &ss)) {  // centre a, ends c < c2
            swart_bend_eval(xyz, c, a, c2, D[c * N + a], D[c2 * N + a], ss, o);
            swart_acc(blk, o, 3, 3);
          }
          {  // centre c, ends a and c2
            const int i = a < c2 ? a : c2, k = a < c2 ? c2 : a;
            if (pass(c, i, k, &ss)) {
              swart_bend_eval(xyz, i, c, k, D[i * N + c], D[k * N + c], ss, o);
              const int sa = a == i ? 0 : 6;
              swart_acc(blk, o, sa, sa);
            }
          }
        }
      }
      int bad = 0;
      for (int p = 0; p < 9; ++p) {
        blk[p] = warp_sum(blk[p]);
        bad |= !isfinite(blk[p]);
      }
      if (lane < 9) H[(size_t)(3 * a + lane / 3) * n + 3 * a + lane % 3] = blk[lane];
      if (bad && lane == 0) s_bad = 1;
    }
    __syncthreads();
    if (!s_bad || !bends) break;  // non-finite: redo with the stretch terms only (swart.py:340-350)
  }
  if (tid == 0 && status) status[b] = s_bad;
}

}  // namespace mop

int mop_launch_project_trrot(int B, int n, const double* H, const double* Hbias, const double* x,
                             const double* g, double* Hp_out, double* gp_out, int32_t* status, int grad_rule,
                             cudaStream_t stream);

static size_t swart_smem(int N, bool in_smem) {
  size_t bytes = sizeof(double) * (4 * (size_t)N + 2 * mop::SW_WARPS * (size_t)N) +
                 sizeof(int) * (mop::SW_WARPS * (size_t)N + 2);
  if (in_smem) bytes += sizeof(double) * 9 * (size_t)N * N;
  return bytes;
}

extern "C" size_t mop_swart_workspace_bytes(int B, int natoms) {
  if (B <= 0 || natoms <= 0) return 0;
  return (size_t)B * 9 * natoms * natoms * sizeof(double);
}

extern "C" int mop_swart_hessian(int B, int natoms, const double* xyz, const double* radii, int radii_stride,
                                 double* H_out, double* Hraw_out, int32_t* status, void* work, size_t work_bytes,
                                 void* stream_) {
  MOP_REQUIRE(B >= 0 && natoms > 0, "mop_swart_hessian: B >= 0 and natoms > 0 required");
  MOP_REQUIRE(xyz && radii && H_out, "mop_swart_hessian: xyz, radii, H_out required");
  MOP_REQUIRE(radii_stride == 0 || radii_stride == natoms, "mop_swart_hessian: radii_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;
  double* Hraw = Hraw_out;
  if (!Hraw) {
    if (!work || work_bytes < mop_swart_workspace_bytes(B, natoms)) {
      mop_set_error("mop_swart_hessian: workspace too small");
      return MOP_ERR_WORKSPACE;
    }
    Hraw = (double*)work;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (natoms <= 100) {  // gather kernel: no atomics
    const size_t smem = sizeof(double) * (4 * (size_t)natoms + 2 * (size_t)natoms * natoms);
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_swart_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_swart_gather<<<B, mop::SW_THREADS, smem, stream>>>(natoms, xyz, radii, radii_stride, Hraw, status);
    MOP_CHECK_CUDA(cudaGetLastError());
    return mop_launch_project_trrot(B, 3 * natoms, Hraw, nullptr, xyz, nullptr, H_out, nullptr, nullptr, 0, stream);
  }
  bool in_smem = swart_smem(natoms, true) <= 220 * 1024;
  const size_t smem = swart_smem(natoms, in_smem);
  if (smem > 220 * 1024) {
    mop_set_error("mop_swart_hessian: natoms = %d too large for the neighbour staging", natoms);
    return MOP_ERR_UNSUPPORTED;
  }
  MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_swart, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mop::k_swart<<<B, mop::SW_THREADS, smem, stream>>>(natoms, in_smem ? 1 : 0, xyz, radii, radii_stride, Hraw, status);
  MOP_CHECK_CUDA(cudaGetLastError());
  return mop_launch_project_trrot(B, 3 * natoms, Hraw, nullptr, xyz, nullptr, H_out, nullptr, nullptr, 0, stream);
}
