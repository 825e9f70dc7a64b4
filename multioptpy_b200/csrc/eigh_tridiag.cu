// Fast batched symmetric eigensolver (Householder tridiagonalisation + bisection +
// inverse iteration).  Placeholder until the kernel lands: reports "unsupported" so
// MOP_EIGH_AUTO selects the Jacobi path.
#include "common.cuh"

int mop_tridiag_supported(int n) { (void)n; return 0; }
size_t mop_tridiag_workspace_bytes(int B, int n) { (void)B; (void)n; return 0; }
int mop_launch_eigh_tridiag(int B, int n, const double* A, double* evals, double* evecs,
                            int32_t* status, void* work, size_t work_bytes, cudaStream_t stream) {
  (void)B; (void)n; (void)A; (void)evals; (void)evecs; (void)status; (void)work; (void)work_bytes; (void)stream;
  mop_set_error("tridiagonal eigensolver not built");
  return MOP_ERR_UNSUPPORTED;
}
