/* neucam_b200.h -- C ABI of libneucam_b200.so, the sm_100a implementation of NeuCam's training hot
 * path (reference: xhuangcv/neucam, training.py:95-271 over modules.py).
 *
 * The reference has no FFI of its own: its boundary is the Python nn.Module surface of modules.py.
 * Every entry point below therefore names the reference statements it replaces; the Python side
 * (neucam_b200/modules.py, neucam_b200/step.py) binds them with ctypes -- see INTEGRATION.md.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates,
 *     frees or retains them (PyTorch owns every buffer);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, >0 = cudaError_t, <0 = NEUCAM_ERR_*;
 *   - there is no CPU fallback: a non-sm_100 device or a missing driver entry point is an error.
 */
#ifndef NEUCAM_B200_H_
#define NEUCAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEUCAM_ABI_VERSION 1

#define NEUCAM_ERR_BAD_ARG (-1)
#define NEUCAM_ERR_NO_DRIVER (-2) /* cuTensorMapEncodeTiled entry point unavailable */
#define NEUCAM_ERR_TMAP (-3)      /* tensor-map encoding rejected */
#define NEUCAM_ERR_ARCH (-4)      /* device is not compute capability 10.x */

/* ABI version of this header; bumped on any signature change. */
int neucam_abi_version(void);

/* 0 if the current device can run the library (compute capability 10.x), NEUCAM_ERR_ARCH otherwise. */
int neucam_check_device(void);

/* Building-block self test (tests/test_umma_probe.py): one 128xN tcgen05 tile GEMM,
 * D[128,N] = A[128,K] * B[N,K]^T with fp32 accumulation.  flags: bit0 A MN-major (A given as [K,128]),
 * bit1 B MN-major (B given as [K,N]), bit2 A is bf16 (else fp16), bit3 B is bf16,
 * bit4 A tile staged by threads instead of TMA. */
int neucam_umma_probe(const void* a, const void* b, float* d, int N, int K, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUCAM_B200_H_ */
