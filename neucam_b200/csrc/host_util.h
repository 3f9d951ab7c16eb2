// host_util.h -- host-side helpers shared by the C-ABI entry points: status codes, argument checks,
// and TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

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

// Encoded tensor maps are cached by (device, base pointer, geometry): a training loop launches the same kernels on
// the same buffers every step, and cuTensorMapEncodeTiled costs microseconds of host time per map.  The cache is
// the library's only global state besides the per-device launch attributes.
struct TmapKey {
  const void* base;
  uint64_t rows, cols, pitch;
  uint32_t box_rows, box_cols, swizzle;
  int dev;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows &&
           box_cols == o.box_cols && swizzle == o.swizzle && dev == o.dev;
  }
};
struct TmapCache {
  std::mutex mu;
  std::vector<std::pair<TmapKey, CUtensorMap>> slots;   // small: linear probe, round-robin replacement
  size_t next = 0;
  static constexpr size_t CAP = 256;
};
inline TmapCache& tmap_cache() {
  static TmapCache c;
  return c;
}

inline int make_tmap_16b_any(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                             uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz) {
  int dev = 0;
  cudaGetDevice(&dev);
  const TmapKey key{base, rows, cols, pitch_elems, box_rows, box_cols, (uint32_t)swz, dev};
  TmapCache& c = tmap_cache();
  {
    std::lock_guard<std::mutex> g(c.mu);
    for (auto& e : c.slots)
      if (e.first == key) {
        *out = e.second;
        return 0;
      }
  }
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) return NEUCAM_ERR_NO_DRIVER;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "neucam_b200: cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu swizzle=%d\n", (int)r,
            (unsigned long long)rows, (unsigned long long)cols, (int)swz);
    return NEUCAM_ERR_TMAP;
  }
  std::lock_guard<std::mutex> g(c.mu);
  if (c.slots.size() < TmapCache::CAP) {
    c.slots.emplace_back(key, *out);
  } else {
    c.slots[c.next] = std::make_pair(key, *out);
    c.next = (c.next + 1) % TmapCache::CAP;
  }
  return 0;
}

// 2-D row-major tensor of 16-bit elements: `rows` x `cols`, row pitch `pitch_elems`.
// Box = box_cols (inner, must be 64 -> 128 B) x box_rows, SWIZZLE_128B.
inline int make_tmap_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_16b_any(out, base, rows, cols, pitch_elems, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Small store tiles written by threads: box rows of 32 elements = 64 B, SWIZZLE_64B (the 16-byte piece index of a
// row is XOR-ed with (row >> 1) & 3; tile base 512-byte aligned).
inline int make_tmap_16b_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_16b_any(out, base, rows, cols, pitch_elems, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_64B);
}

// ... and of 16 elements = 32 B, SWIZZLE_32B (piece index XOR-ed with (row >> 2) & 1; tile base 256-byte aligned).
inline int make_tmap_16b_sw32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_16b_any(out, base, rows, cols, pitch_elems, box_rows, box_cols, CU_TENSOR_MAP_SWIZZLE_32B);
}

constexpr int MAX_DEVICES = 64;

inline int sm_count() {
  static std::atomic<int> n[MAX_DEVICES];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= MAX_DEVICES) dev = 0;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: set it once per (kernel, device)
// -- a process that drives several GPUs launches on each of them.
inline int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  for (auto& e : done)
    if (e.first == func && e.second == dev) return 0;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    fprintf(stderr, "neucam_b200: cudaFuncSetAttribute(%d B) failed: %s\n", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  done.emplace_back(func, dev);
  return 0;
}

// Number of kernels this library has launched (all devices, all streams): bench.py reports it as `gpu_launches`.
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> c{0};
  return c;
}
#define NC_COUNT_LAUNCH(n) nchost::launch_counter().fetch_add((n), std::memory_order_relaxed)

// ---- fixed-order reductions ---------------------------------------------------------------------------------
// Every gradient / loss reduction over blocks goes through per-block partials in a caller-provided workspace and a
// second pass that sums them in block order (csrc/reduce.cu): two runs of the same binary on the same inputs are
// bit-identical (no float atomics anywhere in the library).
struct ReduceSeg {
  float* dst;   // destination tensor (ACCUMULATED into)
  int off;      // offset of its first element inside one partial record
  int n;        // elements
};
constexpr int MAX_REDUCE_SEGS = 16;
struct ReduceSegs {
  ReduceSeg seg[MAX_REDUCE_SEGS];
  int nseg;
};
// dst_s[i] += scale * sum_{p < nparts, in order} ws[p * stride + off_s + i]; scale = *scale_dev_inv ? 1 / that : 1.
int launch_reduce_partials(const float* ws, int nparts, int64_t stride, const ReduceSegs& segs,
                           const float* inv_scale_of, cudaStream_t stream);

// Workspace layouts (floats per partial record x records), shared by the kernels and neucam_reduce_ws_bytes.
constexpr int WS_SINE_BWD_REC = 3 * 512;                       // dW0 [512,2] | db0 [512] per CTA
constexpr int WS_WGRAD_REC = 256 * 256 + 256;                  // one 256 x 256 dW tile + 256 bias sums per (job, Q-slice)
constexpr int WS_WGRAD_WIDE_REC = 256 * 512 + 256;             // scene network: 256 x 512 per (job, Q-slice)
constexpr int WS_SKINNY_REC = 4 * 512;                         // <= 4 output rows x <= 512 columns per block
constexpr int WS_TM_GRADS = 3 * 3 * 128;                       // CRF: (u, a, v) x (r, g, b) x 128 hidden units
constexpr int WS_BLUR_PARAMS = 64 * 3 + 64 + 2 * (64 * 64 + 64) + 27 * 64 + 27;   // 10 331
inline int pad4(int n) { return (n + 3) & ~3; }
inline int64_t ws_bytes_sine_bwd() { return (int64_t)(sm_count() + 1) * WS_SINE_BWD_REC * 4; }
inline int64_t ws_bytes_wgrad() { return (int64_t)(sm_count() / 2 > 24 ? sm_count() / 2 : 24) * WS_WGRAD_WIDE_REC * 4; }
inline int64_t ws_bytes_skinny() { return (int64_t)(2 * sm_count()) * WS_SKINNY_REC * 4; }
inline int ws_rec_psf_crf(int N) { return WS_TM_GRADS + 8 + pad4(N); }   // + gd[3], sum sq, db4[3], pad | dexps[N]
inline int64_t ws_bytes_psf_crf(int64_t B, int N) { return ((B + 1 + 127) / 128) * (int64_t)ws_rec_psf_crf(N) * 4; }
inline int ws_rec_blur_bwd(int N) { return pad4(WS_BLUR_PARAMS) + pad4(N); }
inline int64_t ws_bytes_blur_bwd(int64_t B, int N) {   // one record per block: min(64-pixel groups, 3 x SMs)
  const int64_t groups = (B + 63) / 64;
  return (groups < 3 * sm_count() ? groups : (int64_t)(3 * sm_count())) * ws_rec_blur_bwd(N) * 4;
}
inline int64_t ws_bytes_rigidity(int64_t B) { return ((B + 255) / 256) * 4 * 4; }
inline int64_t ws_bytes_blend(int64_t KB) { return ((KB + 255) / 256) * 4 * 4; }
inline int64_t ws_bytes_head_prep() { return (int64_t)(2 * sm_count()) * 4 * 4; }   // one float4 record per block

#define NC_CHECK_WS(ws, ws_bytes, need)                                                                      \
  do {                                                                                                       \
    if ((ws) == nullptr || (int64_t)(ws_bytes) < (int64_t)(need)) {                                          \
      fprintf(stderr, "neucam_b200: reduction workspace too small: %lld B given, %lld B needed (%s:%d)\n",   \
              (long long)(ws_bytes), (long long)(need), __FILE__, __LINE__);                                 \
      return NEUCAM_ERR_WORKSPACE;                                                                           \
    }                                                                                                        \
  } while (0)

}  // namespace nchost
