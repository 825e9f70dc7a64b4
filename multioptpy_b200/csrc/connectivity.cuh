// Bond / angle / dihedral connectivity tables, bit-exact with
// Utils/bond_connectivity.py:7-134 (the integer part of the parity contract).
//
// One CTA per structure.  The bond matrix is evaluated in parallel with exactly the
// reference's arithmetic (non-fused squares, correctly rounded sqrt, threshold
// 1.1 * (R_i + R_j) in double); the tables are enumerated in the reference's loop order
// by block-wide ORDERED compaction (ballot + prefix), so indices and order are identical.
#pragma once
#include "common.cuh"

namespace mop {

struct ConnTables {
  int* bonds;   // [capB][2]
  int* angles;  // [capA][3]
  int* dihs;    // [capD][4]
  int capB, capA, capD;
  int nb, na, nd;  // filled on return (same value in every thread)
  int overflow;
};

// Ordered append of at most one item per thread: returns the slot of this thread's item
// (or -1) given the running count *cnt (shared).  All threads of the CTA must call.
// wtot: shared int[33].
__device__ __forceinline__ int ordered_slot(bool flag, int* cnt, int* wtot) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const unsigned m = __ballot_sync(MOP_FULL_MASK, flag);
  const int pre = __popc(m & ((1u << lane) - 1u));
  if (lane == 0) wtot[w] = __popc(m);
  __syncthreads();
  int base = *cnt;
  for (int i = 0; i < w; ++i) base += wtot[i];
  int total = 0;
  for (int i = 0; i < nw; ++i) total += wtot[i];
  __syncthreads();
  if (threadIdx.x == 0) *cnt += total;
  __syncthreads();
  return flag ? base + pre : -1;
}

// distance with numpy.linalg.norm's arithmetic: sqrt((dx*dx + dy*dy) + dz*dz), no FMA
__device__ __forceinline__ double np_dist(const double* a, const double* b) {
  const double dx = __dsub_rn(a[0], b[0]), dy = __dsub_rn(a[1], b[1]), dz = __dsub_rn(a[2], b[2]);
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// bondm[i*N+j] = 1 iff dist(i,j) <= (R_i + R_j) * factor, 0 on the diagonal
// (bond_connect_matrix, bond_connectivity.py:13-41; distance row i is coord - coord[i]).
__device__ __forceinline__ void bond_matrix(int N, const double* xyz, const double* rad, double factor,
                                            unsigned char* bondm) {
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int i = e / N, j = e - i * N;
    unsigned char c = 0;
    if (i != j) {
      const double d = np_dist(xyz + 3 * j, xyz + 3 * i);
      const double thr = __dmul_rn(__dadd_rn(rad[j], rad[i]), factor);
      c = d <= thr;
    }
    bondm[e] = c;
  }
  __syncthreads();
}

// candidate dihedral for the ordered angle pair (i < j): the four condition blocks of
// dihedral_angle_connect_table (bond_connectivity.py:69-128), first hit wins.
__device__ __forceinline__ bool dihedral_candidate(const int* A, int i, int j, const unsigned char* bm,
                                                   int N, int out[4]) {
  const int a0 = A[3 * i], a1 = A[3 * i + 1], a2 = A[3 * i + 2];
  const int b0 = A[3 * j], b1 = A[3 * j + 1], b2 = A[3 * j + 2];
#define MOP_BOND(p, q) (bm[(p) * N + (q)] == 1)
#define MOP_SET(w, x, y, z) (out[0] = (w), out[1] = (x), out[2] = (y), out[3] = (z))
  if ((a1 == b1 && a2 == b2) || (a1 == b2 && a2 == b1)) {
    MOP_SET(a0, a1, a2, b0);
    if (MOP_BOND(out[2], out[3])) return true;
    MOP_SET(b0, a0, a1, a2);
    if (MOP_BOND(out[1], out[0])) return true;
  }
  if ((a1 == b1 && a0 == b0) || (a1 == b0 && a0 == b1)) {
    MOP_SET(b2, a0, a1, a2);
    if (MOP_BOND(out[1], out[0])) return true;
    MOP_SET(a0, a1, a2, b2);
    if (MOP_BOND(out[2], out[3])) return true;
  }
  if ((a1 == b0 && a2 == b1) || (a1 == b1 && a2 == b0)) {
    MOP_SET(a0, a1, a2, b2);
    if (MOP_BOND(out[2], out[3])) return true;
    MOP_SET(b2, a0, a1, a2);
    if (MOP_BOND(out[1], out[0])) return true;
  }
  if ((a0 == b1 && a1 == b2) || (a0 == b2 && a1 == b1)) {
    MOP_SET(b0, a0, a1, a2);
    if (MOP_BOND(out[1], out[0])) return true;
    MOP_SET(a0, a1, a2, b0);
    if (MOP_BOND(out[2], out[3])) return true;
  }
#undef MOP_BOND
#undef MOP_SET
  return false;
}

// Enumerate the three tables.  cnt: shared int[3]; wtot: shared int[33].
__device__ __forceinline__ void enumerate_tables(int N, const unsigned char* bm, ConnTables& T, int* cnt,
                                                 int* wtot) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 3) cnt[tid] = 0;
  __syncthreads();
  int overflow = 0;
  // bonds: (i, j), i <= j, row-major (bond_connect_table :43-54)
  for (int e0 = 0; e0 < N * N; e0 += nt) {
    const int e = e0 + tid;
    bool f = false;
    int i = 0, j = 0;
    if (e < N * N) {
      i = e / N;
      j = e - i * N;
      f = (i <= j) && bm[e] == 1;
    }
    const int s = ordered_slot(f, &cnt[0], wtot);
    if (s >= 0) {
      if (s < T.capB) {
        T.bonds[2 * s] = i;
        T.bonds[2 * s + 1] = j;
      } else {
        overflow = 1;
      }
    }
  }
  // angles: for i, for j (bonded to i), for n > j bonded to i and NOT to j -> [j, i, n] (:56-67)
  for (int i = 0; i < N; ++i) {
    for (int e0 = 0; e0 < N * N; e0 += nt) {
      const int e = e0 + tid;
      bool f = false;
      int j = 0, n = 0;
      if (e < N * N) {
        j = e / N;
        n = e - j * N;
        f = (n > j) && bm[i * N + j] == 1 && bm[i * N + n] == 1 && bm[j * N + n] == 0;
      }
      const int s = ordered_slot(f, &cnt[1], wtot);
      if (s >= 0) {
        if (s < T.capA) {
          T.angles[3 * s] = j;
          T.angles[3 * s + 1] = i;
          T.angles[3 * s + 2] = n;
        } else {
          overflow = 1;
        }
      }
    }
  }
  __syncthreads();
  const int na = min(cnt[1], T.capA);
  // dihedrals: ordered pairs of angles (i < j) (:69-128)
  for (int i = 0; i < na; ++i) {
    for (int j0 = i + 1; j0 < na; j0 += nt) {
      const int j = j0 + tid;
      int c[4] = {0, 0, 0, 0};
      bool f = false;
      if (j < na) f = dihedral_candidate(T.angles, i, j, bm, N, c);
      const int s = ordered_slot(f, &cnt[2], wtot);
      if (s >= 0) {
        if (s < T.capD) {
          T.dihs[4 * s] = c[0];
          T.dihs[4 * s + 1] = c[1];
          T.dihs[4 * s + 2] = c[2];
          T.dihs[4 * s + 3] = c[3];
        } else {
          overflow = 1;
        }
      }
    }
  }
  __syncthreads();
  overflow = __syncthreads_or(overflow);
  T.nb = min(cnt[0], T.capB);
  T.na = na;
  T.nd = min(cnt[2], T.capD);
  T.overflow = overflow || cnt[0] > T.capB || cnt[1] > T.capA || cnt[2] > T.capD;
  __syncthreads();
}

}  // namespace mop
