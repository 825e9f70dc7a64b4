/* Private symbols of libmop_b200.so: measurement probes (bench.py) and tuning / diagnostic hooks (tools/).  Not part
 * of the drop-in boundary (include/mop_b200.h); the hooks set process-wide state and must not be used concurrently
 * with product calls. */
#ifndef MOP_PRIVATE_H
#define MOP_PRIVATE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* FP64 FMA peak probe: one launch = blocks * 256 * iters * 64 * 2 flops; L2 flush / fill */
int mop_priv_bench_dfma(int blocks, int iters, double* out, void* stream);
int mop_priv_bench_fill(double* buf, size_t count, double value, void* stream);
/* unit probes: fast_rcp on an array; dependent-chain latencies (DFMA, DADD, DMUL, LDS, SHFL64, rcp, sqrt, div);
 * __syncthreads latency at a CTA size */
int mop_priv_fast_rcp(const double* x, double* out, size_t count, void* stream);
int mop_priv_latency(double* out, void* stream);
int mop_priv_barrier_latency(int threads, double* out, void* stream);
/* phase-cycle buffers ([B][16] int64 on the device, NULL = off) */
int mop_priv_spectrum_timing(void* buf);
int mop_priv_tridiag_blk_timing(void* buf);
int mop_priv_tridiag_cluster_timing(void* buf);
int mop_priv_large_timing(void* buf);   /* k_lg_trieig: [2][B][4] int64, second half = eigenvalues | vectors | cluster fix */
/* cluster tridiagonalisation: CTAs per matrix (1, 2, 4, 8; 0 = by batch size), lower-triangle symv on / off,
 * ablation mask (results invalid) */
int mop_priv_large_cluster(int cluster_ctas);
int mop_priv_tridiag_cluster_sym(int on);
int mop_priv_tridiag_cluster_ablate(int mask);
#ifdef __cplusplus
}
#endif
#endif
