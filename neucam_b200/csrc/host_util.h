// host_util.h -- host-side helpers shared by the C-ABI entry points: status codes, argument checks,
// and TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/neucam_b200.h"

namespace nchost {

#define NC_CHECK_ARG(cond)                                                                 \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      fprintf(stderr, "neucam_b200: bad argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      return NEUCAM_ERR_BAD_ARG;                                                           \
    }                                                                                      \
  } while (0)

#define NC_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fprintf(stderr, "neucam_b200: CUDA error %d (%s) at %s:%d\n", (int)_e,                 \
              cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
      return (int)_e;                                                                        \
    }                                                                                        \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor of 16-bit elements: `rows` x `cols`, row pitch `pitch_elems`.
// Box = box_cols (inner, must be 64 -> 128 B) x box_rows, SWIZZLE_128B.
inline int make_tmap_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return NEUCAM_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "neucam_b200: cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols);
    return NEUCAM_ERR_TMAP;
  }
  return 0;
}

// Small store tiles written by threads: box rows of 32 elements = 64 B, SWIZZLE_64B (the 16-byte piece index of a
// row is XOR-ed with (row >> 1) & 3; tile base 512-byte aligned).
inline int make_tmap_16b_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return NEUCAM_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "neucam_b200: cuTensorMapEncodeTiled (sw64) failed (%d) rows=%llu cols=%llu\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols);
    return NEUCAM_ERR_TMAP;
  }
  return 0;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace nchost
