/* neucam_b200_debug.h -- entry points that exist ONLY in the debug build of the library
 * (libneucam_b200_debug.so = the same sources compiled with -DNEUCAM_DEBUG; `python -m neucam_b200.build --debug`).
 * The product library (libneucam_b200.so, include/neucam_b200.h) contains none of this: no process-global debug
 * switches, no in-kernel timelines, no microbenchmark kernels.
 */
#ifndef NEUCAM_B200_DEBUG_H_
#define NEUCAM_B200_DEBUG_H_

#include "neucam_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Building-block self test (tests/test_umma_probe.py): one 128xN tcgen05 tile GEMM,
 * D[128,N] = A[128,K] * B[N,K]^T with fp32 accumulation.  flags: bit0 A MN-major (A given as [K,128]),
 * bit1 B MN-major (B given as [K,N]), bit2 A is bf16 (else fp16), bit3 B is bf16,
 * bit4 A tile staged by threads instead of TMA, bit5 A operand in tensor memory. */
int neucam_umma_probe(const void* a, const void* b, float* d, int N, int K, int flags, void* stream);

/* Tensor-pipe issue-rate microbenchmark (tools/umma_rate.py): `grid` CTAs in pairs; the leader of each pair
 * issues iters x 4 back-to-back MMAs (cta_group::2 M256xNxK16 if two_cta else cta_group::1 M128xNxK16) on
 * resident smem tiles; out_dev[2*cta] = issue cycles, out_dev[2*cta+1] = cycles until completion. */
int neucam_umma_rate(int N, int iters, int two_cta, int kblocks, int grid, long long* out_dev, void* stream);

/* CTA 0 of subsequent neucam_sine_mlp_fwd/bwd launches records globaltimer stamps of its warp roles into trace_dev
 * (int64[576], device); NULL switches it off.  See tools/chain_timeline.py. */
int neucam_debug_set_chain_trace(long long* trace_dev);
/* Experiments (outputs become garbage): bit0 = chain kernels issue no MMAs, bit1 = no weight TMA loads,
 * bit2 = no epilogue math. */
int neucam_debug_set_chain_flags(int flags);
int neucam_debug_set_relu_trace(void* trace);  /* CTA 0 event/clock trace of the next ReLU-chain forward launches */
int neucam_debug_set_relu_flags(int flags);    /* timing experiments on the ReLU chain (forward); results invalid */

#ifdef __cplusplus
}
#endif
#endif /* NEUCAM_B200_DEBUG_H_ */
