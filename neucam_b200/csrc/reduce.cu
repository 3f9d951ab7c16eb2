// reduce.cu -- the second pass of every cross-block reduction in the library.
//
// No kernel of this library uses float atomics: a kernel that reduces over blocks (weight gradients, bias gradients,
// loss sums) writes one partial record per block into a caller-provided workspace, and the kernel below adds the
// records up IN BLOCK ORDER into the destination tensors.  Two runs of the same binary on the same inputs therefore
// produce bit-identical gradients, losses and parameters -- which is what makes "same PSNR after the same number of
// iterations" (BASELINE.json north_star) a testable statement for a chaotic training trajectory.
#include "host_util.h"

namespace {

// A block owns 32 consecutive elements of the record.  Its 16 warps split the nparts records between them (warp g sums
// records g, g + 16, g + 32, ... in that order, eight loads in flight) and the sixteen partial sums are then added in
// warp order by warp 0: a FIXED summation tree, so the result is bit-reproducible.  The kernel is latency-bound (a few
// MB spread over a few hundred blocks), so what matters is loads in flight: 16 x 8 per 32 outputs (the first version,
// one serial loop per output, took 24 us per call; 8 warps x 4 loads 10-20 us).
constexpr int RG = 16;   // record groups = warps per block
__global__ void __launch_bounds__(32 * RG)
reduce_partials_kernel(const float* __restrict__ ws, int nparts, long long stride, nchost::ReduceSegs segs, int total,
                       const float* __restrict__ inv_scale_of) {
  __shared__ float s_part[RG][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (i < total) {
    const float* p = ws + i;
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
    int q = g;
    for (; q + 7 * RG < nparts; q += 8 * RG) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcs(p + (long long)(q + k * RG) * stride);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += v[k];
    }
    {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (q + k * RG < nparts) ? __ldcs(p + (long long)(q + k * RG) * stride) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += v[k];
    }
    acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  s_part[g][lane] = acc;
  __syncthreads();
  if (g != 0 || i >= total) return;
  float* dst = nullptr;
#pragma unroll 1
  for (int s = 0; s < segs.nseg; ++s) {
    const int r = i - segs.seg[s].off;
    if (r >= 0 && r < segs.seg[s].n) dst = segs.seg[s].dst + r;
  }
  if (dst == nullptr) return;   // padding between segments
  float t = s_part[0][lane];
#pragma unroll
  for (int k = 1; k < RG; ++k) t += s_part[k][lane];
  if (inv_scale_of != nullptr) t *= 1.f / *inv_scale_of;
  *dst += t;
}

}  // namespace

namespace nchost {

int launch_reduce_partials(const float* ws, int nparts, int64_t stride, const ReduceSegs& segs,
                           const float* inv_scale_of, cudaStream_t stream) {
  int total = 0;
  for (int s = 0; s < segs.nseg; ++s)
    if (segs.seg[s].off + segs.seg[s].n > total) total = segs.seg[s].off + segs.seg[s].n;
  if (total == 0 || nparts <= 0) return 0;
  reduce_partials_kernel<<<(total + 31) / 32, 32 * RG, 0, stream>>>(ws, nparts, (long long)stride, segs, total,
                                                                    inv_scale_of);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

}  // namespace nchost

extern "C" int neucam_reduce_ws_bytes(int kind, int64_t n, int aux, int64_t* bytes_out) {
  NC_CHECK_ARG(bytes_out != nullptr && n >= 0 && aux >= 0);
  int64_t b = 0;
  switch (kind) {
    case NEUCAM_WS_SINE_BWD: b = nchost::ws_bytes_sine_bwd(); break;
    case NEUCAM_WS_WGRAD: b = nchost::ws_bytes_wgrad(); break;
    case NEUCAM_WS_HEAD_WGRAD: b = nchost::ws_bytes_skinny(); break;
    case NEUCAM_WS_PSF_CRF: b = nchost::ws_bytes_psf_crf(n, aux); break;
    case NEUCAM_WS_BLUR_BWD: b = nchost::ws_bytes_blur_bwd(n, aux); break;
    case NEUCAM_WS_RIGIDITY: b = nchost::ws_bytes_rigidity(n); break;
    case NEUCAM_WS_BLEND: b = nchost::ws_bytes_blend(n); break;
    case NEUCAM_WS_HEAD_PREP: b = nchost::ws_bytes_head_prep(); break;
    case NEUCAM_WS_TM: b = ((n + 127) / 128) * (int64_t)(nchost::WS_TM_GRADS + 4) * 4; break;
    default: NC_CHECK_ARG(!"unknown workspace kind");
  }
  *bytes_out = b;
  return 0;
}

extern "C" int neucam_launch_count(int64_t* count_out) {
  NC_CHECK_ARG(count_out != nullptr);
  *count_out = (int64_t)nchost::launch_counter().load();
  return 0;
}
