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
// Gather formulation (natoms <= 100): no atomics.  Every off-diagonal 3 x 3 atom-pair block (a < b) of
// the raw Hessian is owned by one WARP, which sums the stretch term and every bend that touches both
// atoms: centre c with ends (a, b), centre a with ends (b, c), centre b with ends (a, c) — 3 (N - 2)
// candidates, of which the screens pass about a fifth.  The lanes screen the candidates 32 at a time
// (two contiguous rows of the screen table) and compact the survivors into the warp's queue with
// ballots; then every lane evaluates one queued bend per round — regular, near-180 and near-0 degree
// (linear-bend pair) bends through ONE code path, U1 = c1 bn + c2 t / l, U2 = c3 vn / l — into a private
// 3 x 3 accumulator that is reduced once per pair.  History (1024 structures, N = 50): FP64 shared-memory
// atomics 13.8 ms; thread-per-block gather that evaluated inside the candidate loop 16.7 ms at 6 of 32
// lanes active (profiles/r2_ncu_producers_summary.csv); this kernel 1.1 ms.  The diagonal blocks follow
// from translational invariance — every Wilson vector of a stretch or bend sums to zero over its
// atoms, so block(a, a) = - sum_{b != a} block(a, b) — summed in fixed order from the rows the CTA
// just wrote: a bend is evaluated three times instead of six and the result is deterministic.
// Reciprocal distances are tabulated and sqrt / division go through MUFU seeds + Newton steps (~1 ulp,
// no IEEE slow path); the parity bar is 1e-10.
constexpr int SWP_THREADS = 256;
constexpr int SWP_WARPS = SWP_THREADS / 32;

// MUFU seed + three Newton steps: ~1 ulp for normal x (callers clamp x >= 1e-12)
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
  }
  return r;
}

// smem doubles: xyz 3N | rad N | ID N^2 (1 / clamped distance) | Sc N^2 (screen; negated where the distance
// is degenerate, 0 on the diagonal) | per warp: bend queue (3N screens + 3N codes)
__host__ __device__ inline size_t swp_smem_bytes(int N) {
  return sizeof(double) * (4 * (size_t)N + 2 * (size_t)N * N + (size_t)SWP_WARPS * 3 * N) +
         sizeof(int) * ((size_t)SWP_WARPS * 3 * N + 2);
}

__global__ void __launch_bounds__(SWP_THREADS, 3)
k_swart_pair(int N, const double* __restrict__ xyz_all, const double* __restrict__ rad_all, int rad_stride,
             double* H_all, int32_t* __restrict__ status) {
  extern __shared__ double sm[];
  const SwartConst C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = 3 * N;
  double* xyz = sm;           // 3N
  double* rad = xyz + 3 * N;  // N
  double* ID = rad + N;       // N x N
  double* Sc = ID + N * N;    // N x N
  double* qss = Sc + N * N + (size_t)wid * 3 * N;                               // this warp's queue: s_ij s_jk
  int* qcode = (int*)(Sc + N * N + (size_t)SWP_WARPS * 3 * N) + (size_t)wid * 3 * N;  // (type << 8) | c
  __shared__ int s_bad;
  double* H = H_all + (size_t)b * n * n;
  for (int i = tid; i < 3 * N; i += SWP_THREADS) xyz[i] = xyz_all[(size_t)b * 3 * N + i];
  for (int i = tid; i < N; i += SWP_THREADS) rad[i] = rad_all[(size_t)b * rad_stride + i];
  if (tid == 0) s_bad = 0;
  __syncthreads();
  const int npair = N * (N - 1) / 2;
  // pair t -> (a < bb): row a holds the pairs (a, a+1 .. N-1)
  auto pair_of = [&](int t, int* pa, int* pb) {
    int a = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * t)) * 0.5);
    while (a > 0 && a * (2 * N - a - 1) / 2 > t) --a;
    while ((a + 1) * (2 * N - a - 2) / 2 <= t) ++a;
    *pa = a;
    *pb = a + 1 + (t - a * (2 * N - a - 1) / 2);
  };
  for (int t = tid; t < npair; t += SWP_THREADS) {
    int i, j;
    pair_of(t, &i, &j);
    double d;
    double s = swart_screen(xyz, rad, i, j, &d);
    const double id = 1.0 / d;
    if (!(d > 1e-8)) s = -s;  // bends skip degenerate distances (swart.py:226: l > 1e-8), stretches use |s|
    ID[i * N + j] = id; ID[j * N + i] = id;
    Sc[i * N + j] = s; Sc[j * N + i] = s;
  }
  for (int i = tid; i < N; i += SWP_THREADS) {
    ID[i * N + i] = 1.0;
    Sc[i * N + i] = 0.0;
  }
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    const bool bends = round == 0;
    // ---- off-diagonal blocks: one warp per pair (a < bb), pairs dealt cyclically ----
    for (int t = wid; t < npair; t += SWP_WARPS) {
      int a, bb;
      pair_of(t, &a, &bb);
      const double* Sa = Sc + a * N;
      const double* Sb = Sc + bb * N;
      const double sab = Sa[bb];
      int qn = 0;
      if (bends) {
        // screen the 3 (N - 2) candidate bends, lanes over the third atom c:
        //   type 0: centre c, ends (a, bb)   type 1: centre a, ends (bb, c)   type 2: centre bb, ends (a, c)
        const bool ab_ok = sab >= C.eps2;
        for (int c0 = 0; c0 < N; c0 += 32) {
          const int c = c0 + lane;
          const bool in = c < N && c != a && c != bb;
          const double sa = in ? Sa[c] : 0.0, sb = in ? Sb[c] : 0.0;
          const bool na = sa >= C.eps2, nb = sb >= C.eps2;
          const double ss0 = sa * sb, ss1 = sab * sa, ss2 = sab * sb;
          const bool p0 = na && nb && ss0 >= C.eps1;
          const bool p1 = ab_ok && na && ss1 >= C.eps1;
          const bool p2 = ab_ok && nb && ss2 >= C.eps1;
          const unsigned lt = (1u << lane) - 1u;
          unsigned m = __ballot_sync(MOP_FULL_MASK, p0);
          if (p0) { const int q = qn + __popc(m & lt); qss[q] = ss0; qcode[q] = c; }
          qn += __popc(m);
          m = __ballot_sync(MOP_FULL_MASK, p1);
          if (p1) { const int q = qn + __popc(m & lt); qss[q] = ss1; qcode[q] = 256 | c; }
          qn += __popc(m);
          m = __ballot_sync(MOP_FULL_MASK, p2);
          if (p2) { const int q = qn + __popc(m & lt); qss[q] = ss2; qcode[q] = 512 | c; }
          qn += __popc(m);
        }
        __syncwarp();
      }
      double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int q = lane; q < qn; q += 32) {
        const int code = qcode[q], ty = code >> 8, c = code & 255;
        const double ss = qss[q];
        // centre j, ends (p, q2): the bend is symmetric in its ends, so the end that belongs to the block comes first
        const int j = ty == 0 ? c : (ty == 1 ? a : bb);
        const int p = ty == 1 ? bb : a;
        const int q2 = ty == 0 ? bb : c;
        const double il1 = ID[j * N + p], il2 = ID[j * N + q2];
        double n1[3], n2[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double xj = xyz[3 * j + k];
          n1[k] = (xyz[3 * p + k] - xj) * il1;
          n2[k] = (xyz[3 * q2 + k] - xj) * il2;
        }
        double cs = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
        cs = fmin(fmax(cs, -1.0), 1.0);
        const double s2 = fmax(1e-12, 1.0 - cs * cs);
        const double rs = fast_rsqrt(s2);
        const double sn = s2 * rs;
        const double iden = fmin(rs, 1e6);  // 1 / max(sn, 1e-6)
        const double w = C.f + (1.0 - C.f) * sn;
        const double hb = 0.075 * (ss * ss) * (w * w);
        // Wilson vectors as  U1 = c1 bn + c2 t / l,  U2 = c3 vn / l  (swart.py:226-315):
        //   regular bend: c1 = 1;  near 180 degrees: c1 = 1 - sl;  near 0 degrees (cos > 0.8): the linear-bend pair
        //   c1 = 1 - sl, c2 = sl with t = unit((v1 - v2) x vn), c3 = 1 with vn = unit(v1 x v2)
        double c1 = 1.0, c2 = 0.0, c3 = 0.0;
        double tv[3] = {0.0, 0.0, 0.0}, vn[3] = {0.0, 0.0, 0.0};
        const bool lin = cs > 1.0 - C.tolth;
        const double th1 = lin ? 1.0 - cs : 1.0 + cs;
        if (th1 < C.tolth) {
          const double qq = th1 / C.tolth;
          const double sl = (1.0 - qq * qq) * (1.0 - qq * qq);
          c1 = 1.0 - sl;
          if (lin) {
            c2 = sl;
            c3 = 1.0;
            // v1 x v2 = l1 l2 (n1 x n2); only the direction enters unless it is degenerate (< 1e-12)
            vn[0] = n1[1] * n2[2] - n1[2] * n2[1];
            vn[1] = n1[2] * n2[0] - n1[0] * n2[2];
            vn[2] = n1[0] * n2[1] - n1[1] * n2[0];
            double nv2 = vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2];
            const double lim = 1e-12 * il1 * il2;
            if (nv2 < lim * lim) {  // exactly collinear: perpendicular to the lower-index end (swart.py:150-165)
              const double* nf = p < q2 ? n1 : n2;
              vn[0] = 1.0 - nf[0] * nf[0]; vn[1] = -nf[0] * nf[1]; vn[2] = -nf[0] * nf[2];
              nv2 = vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2];
              if (!(nv2 >= 1e-24)) {
                vn[0] = -nf[1] * nf[0]; vn[1] = 1.0 - nf[1] * nf[1]; vn[2] = -nf[1] * nf[2];
                nv2 = fmax(vn[0] * vn[0] + vn[1] * vn[1] + vn[2] * vn[2], 1e-24);
              }
            }
            const double rv = fast_rsqrt(nv2);
            vn[0] *= rv; vn[1] *= rv; vn[2] *= rv;
            // v1 - v2 is proportional to n1 il2 - n2 il1
            const double vd[3] = {n1[0] * il2 - n2[0] * il1, n1[1] * il2 - n2[1] * il1, n1[2] * il2 - n2[2] * il1};
            tv[0] = vd[1] * vn[2] - vd[2] * vn[1];
            tv[1] = vd[2] * vn[0] - vd[0] * vn[2];
            tv[2] = vd[0] * vn[1] - vd[1] * vn[0];
            const double rt = fast_rsqrt(fmax(tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2], 1e-300));
            tv[0] *= rt; tv[1] *= rt; tv[2] *= rt;
          }
        }
        const double f1 = c1 * il1 * iden, f2 = c1 * il2 * iden, g1 = c2 * il1, g2 = c2 * il2, h1 = c3 * il1, h2 = c3 * il2;
        double ua[3], ub[3], wa[3], wb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double up = fma(g1, tv[k], (cs * n1[k] - n2[k]) * f1);
          const double uq = fma(g2, tv[k], (cs * n2[k] - n1[k]) * f2);
          const double uj = -(up + uq);
          ua[k] = ty == 1 ? uj : up;
          ub[k] = ty == 0 ? uq : (ty == 1 ? up : uj);
          const double wp = vn[k] * h1, wq = vn[k] * h2, wj = -(wp + wq);
          wa[k] = ty == 1 ? wj : wp;
          wb[k] = ty == 0 ? wq : (ty == 1 ? wp : wj);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double hp = hb * ua[k], hw = hb * wa[k];
#pragma unroll
          for (int m = 0; m < 3; ++m) blk[3 * k + m] = fma(hw, wb[m], fma(hp, ub[m], blk[3 * k + m]));
        }
      }
      if (qn > 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) blk[k] = warp_sum(blk[k]);
      }
      if (lane < 9) {
        const int pp = lane / 3, qq = lane - 3 * pp;
        const double sabs = fabs(sab), h = -0.35 * (sabs * sabs * sabs), id = ID[a * N + bb];
        const double ep = (xyz[3 * a + pp] - xyz[3 * bb + pp]) * id, eq = (xyz[3 * a + qq] - xyz[3 * bb + qq]) * id;
        double x = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (k == lane) x = blk[k];
        x = fma(h * ep, eq, x);
        if (!isfinite(x)) s_bad = 1;
        H[(size_t)(3 * a + pp) * n + 3 * bb + qq] = x;
        H[(size_t)(3 * bb + qq) * n + 3 * a + pp] = x;
      }
      __syncwarp();
    }
    __syncthreads();
    if (!s_bad || !bends) break;  // non-finite: redo with the stretch terms only (swart.py:340-350)
    __syncthreads();
  }
  // diagonal blocks from translational invariance, fixed summation order
  for (int e = tid; e < 3 * n; e += SWP_THREADS) {
    const int r = e / 3, q = e - 3 * r, a = r / 3;
    const double* row = H + (size_t)r * n + q;
    double acc = 0.0;
    for (int c = 0; c < N; ++c)
      if (c != a) acc += row[3 * c];
    H[(size_t)r * n + 3 * a + q] = -acc;
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
  MOP_REQUIRE(radii_stride == 0 || radii_stride == natoms, "mop_swart_hessian: radii_stride must be 0 or natoms");
  if (B == 0) return MOP_OK;   // (an empty batch has no buffers)
  MOP_REQUIRE(xyz && radii && H_out, "mop_swart_hessian: xyz, radii, H_out required");
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
    const size_t smem = mop::swp_smem_bytes(natoms);
    MOP_CHECK_CUDA(cudaFuncSetAttribute(mop::k_swart_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mop::k_swart_pair<<<B, mop::SWP_THREADS, smem, stream>>>(natoms, xyz, radii, radii_stride, Hraw, status);
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
