// step.cu -- neucam_step_fwd_bwd: the whole static training step (training.py:95-270 for the multi-focus /
// multi-exposure model set) enqueued from ONE C call.  It composes the library's own entry points in the order the
// reference's autograd graph dictates, all on the caller's stream.  (Running the blur-MLP backward on a second stream
// under the weight-gradient GEMM was measured and dropped: the SIMT warps delay the GEMM's single-thread issue paths by
// more than they hide -- 569 us together vs 455 + ~60 us in sequence, profiles/r2_findings.md.)
#include "host_util.h"
#include <cuda_fp16.h>

namespace {

// K = 1 (no blur model): the scene MLP is queried at the pixel itself, uv = coords[:, 1:3] (training.py:123-126)
__global__ void __launch_bounds__(256) coords_to_uv_kernel(const float* __restrict__ coords, int B,
                                                           float* __restrict__ uv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < B) *reinterpret_cast<float2*>(uv + 2 * p) = make_float2(coords[3 * p + 1], coords[3 * p + 2]);
}

#define NC_TRY(expr)         \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

}  // namespace

extern "C" int neucam_step_fwd_bwd(const neucam_step_args* a, void* stream, void* side_stream) {
  NC_CHECK_ARG(a != nullptr && a->struct_bytes == sizeof(neucam_step_args));
  NC_CHECK_ARG(a->B > 0 && (a->K == 9 || a->K == 1) && a->N > 0 && a->H > 1 && a->W > 1 && a->out_dim == 3);
  NC_CHECK_ARG(a->coords && a->coord_idx && a->img && a->w0 && a->b0 && a->b_hid && a->w4 && a->b4 && a->w_hid16 &&
               a->wT_hid16 && a->w4T16 && a->exps && a->g_exps && a->g_w0 && a->g_b0 && a->g_w4 && a->g_b4);
  NC_CHECK_ARG(a->uv && a->y && a->dy && a->duv && a->act16 && a->phase16 && a->dz16 && a->sums && a->absmax_bits &&
               a->scale && a->ws);
  const bool has_blur = a->K == 9;
  NC_CHECK_ARG(!has_blur || (a->focus && a->g_focus && a->kout && a->hid && a->dkw && a->ws_side));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t Q = (int64_t)a->K * a->B;
  const int phases = a->phases == 0 ? 3 : a->phases;
  NC_CHECK_ARG((phases & ~3) == 0);
  const bool whole = phases == 3;

  (void)side_stream;   // reserved: every kernel of the step runs on `stream` (see the header)
  if (phases & NEUCAM_STEP_TO_HIDDEN_WGRAD) {
  if (a->zero_ptr != nullptr && a->zero_bytes > 0) NC_CHECK_CUDA(cudaMemsetAsync(a->zero_ptr, 0, a->zero_bytes, st));
  NC_CHECK_CUDA(cudaMemsetAsync(a->sums, 0, 4 * sizeof(float), st));
  NC_CHECK_CUDA(cudaMemsetAsync(a->absmax_bits, 0, sizeof(uint32_t), st));

  // S2-S5: taps (training.py:99-126)
  if (has_blur) {
    NC_TRY(neucam_blur_fwd(a->coords, a->coord_idx, a->focus, a->blur_params, a->B, a->K, a->N, a->H, a->W,
                           a->offset_size, a->uv, a->kout, a->hid, stream));
  } else {
    coords_to_uv_kernel<<<(a->B + 255) / 256, 256, 0, st>>>(a->coords, a->B, a->uv);
    NC_CHECK_CUDA(cudaGetLastError());
    NC_COUNT_LAUNCH(1);
  }
  // S7: scene MLP on all K*B taps (training.py:128-130)
  NC_TRY(neucam_sine_mlp_fwd(a->uv, Q, a->w0, a->b0, a->w_hid16, a->b_hid, a->w4, a->b4, a->out_dim, a->y, a->act16,
                             a->phase16, stream));
  // S9-S12 forward and backward (training.py:157-188, 244-245)
  NC_TRY(neucam_psf_crf_loss(a->y, has_blur ? a->kout : nullptr, a->coord_idx, a->exps, nullptr, a->img, a->gamma_w,
                             a->mask_w, a->tm_params, a->g_tm, a->B, a->K, a->N, a->H, a->W, has_blur ? 1 : 0,
                             a->gamma_dev, a->fixed_value, a->ue_weight, a->global_count, a->dy,
                             has_blur ? a->dkw : nullptr, a->g_exps, a->g_b4, a->sums, a->absmax_bits, 16.0f, a->scale,
                             nullptr, a->tm_width, a->ws, a->ws_bytes, stream));
  // S14: scene MLP backward (training.py:270)
  NC_TRY(neucam_sine_mlp_bwd(a->uv, Q, a->w0, a->b0, a->wT_hid16, a->w4T16, a->out_dim, a->dy, a->phase16, a->scale,
                             a->dz16, a->duv, a->g_w0, a->g_b0, a->ws, a->ws_bytes, stream));
  // blur-MLP backward (needs d(uv) of the chain; training.py:99-126 backward).  A call that covers the whole step runs
  // it here; a data-parallel caller that splits the step gets it in the tails, under its all-reduce of the atlas bucket
  if (has_blur && whole)
    NC_TRY(neucam_blur_bwd(a->coords, a->coord_idx, a->focus, a->blur_params, a->g_blur, a->g_focus, a->B, a->K, a->N,
                           a->H, a->W, a->offset_size, a->kout, a->hid, a->dkw, a->duv, a->ws_side, a->ws_side_bytes,
                           stream));
  NC_TRY(neucam_sine_mlp_wgrad(a->dz16, a->act16, Q, a->scale, a->g_w[0], a->g_b[0], a->g_w[1], a->g_b[1], a->g_w[2],
                               a->g_b[2], a->ws, a->ws_bytes, stream));
  }
  if (!(phases & NEUCAM_STEP_TAILS)) return 0;
  if (has_blur && !whole)
    NC_TRY(neucam_blur_bwd(a->coords, a->coord_idx, a->focus, a->blur_params, a->g_blur, a->g_focus, a->B, a->K, a->N,
                           a->H, a->W, a->offset_size, a->kout, a->hid, a->dkw, a->duv, a->ws_side, a->ws_side_bytes,
                           stream));
  // a4 = sin(30 z_3) is regenerated from the stored phase of z_3 (third slice of phase16)
  const int64_t ph_stride = ((Q + 127) / 128) * 128 * 512;
  const uint16_t* phase3 = (const uint16_t*)a->phase16 + (size_t)2 * ph_stride;
  NC_TRY(neucam_sine_mlp_head_wgrad(phase3, a->dy, a->out_dim, Q, a->g_w4, a->ws, a->ws_bytes, stream));
  return 0;
}
