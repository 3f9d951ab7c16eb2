// tonemap.cu -- TonemapNet (modules.py:197-253, noise=None branch) as a standalone differentiable op: the camera
// response function  ldr_c = tanh(v_c . relu(u_c * lnx_c + a_c) + d_c),  lnx = x + ln2 * exposure,  per channel c.
//
// Inside the training step this math is fused into psf_crf_loss_kernel (pixel.cu); the standalone op serves the
// callers around the step that evaluate the CRF on its own: the CRF warm-up (utils.warm_tm, utils.py:265-282) and
// full-frame rendering (utils.write_video_summary / render_video.py).  Backward recomputes the pre-activations
// (2 x 128 FMAs per channel) instead of storing them; parameter gradients are reduced per block with
// thread = hidden unit loops and then over blocks in block order (reduce.cu) -- no atomics.
#include "host_util.h"

namespace {

constexpr int TMW = 128;
constexpr int PB = 128;
constexpr float LN2 = 0.6931471805599453f;

struct TmP {
  const float *u[3], *a[3], *v[3], *d[3];
};

__global__ void __launch_bounds__(PB)
tm_fwd_kernel(TmP tm, const float* __restrict__ x, const float* __restrict__ expo, int M, float* __restrict__ ldr) {
  __shared__ float s_u[3][TMW], s_a[3][TMW], s_v[3][TMW];
  const int tid = threadIdx.x;
  for (int c = 0; c < 3; ++c) {
    s_u[c][tid] = tm.u[c][tid];
    s_a[c][tid] = tm.a[c][tid];
    s_v[c][tid] = tm.v[c][tid];
  }
  __syncthreads();
  const int i = blockIdx.x * PB + tid;
  if (i >= M) return;
  const float e = LN2 * expo[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float lnx = x[3 * (size_t)i + c] + e;
    float o = tm.d[c][0];
#pragma unroll 8
    for (int k = 0; k < TMW; ++k) o = fmaf(s_v[c][k], fmaxf(fmaf(s_u[c][k], lnx, s_a[c][k]), 0.f), o);
    ldr[3 * (size_t)i + c] = tanhf(o);
  }
}

// record layout per block: [3 channels][gu 128 | ga 128 | gv 128] | gd[3] | pad
constexpr int REC = 3 * 3 * TMW + 4;

__global__ void __launch_bounds__(PB)
tm_bwd_kernel(TmP tm, const float* __restrict__ x, const float* __restrict__ expo, const float* __restrict__ dldr,
              int M, float* __restrict__ dx, float* __restrict__ dexpo, float* __restrict__ part) {
  __shared__ float s_u[3][TMW], s_a[3][TMW], s_v[3][TMW];
  __shared__ float s_x[3][PB + 1], s_g[3][PB + 1];
  __shared__ float s_red[PB / 32][4];
  const int tid = threadIdx.x;
  for (int c = 0; c < 3; ++c) {
    s_u[c][tid] = tm.u[c][tid];
    s_a[c][tid] = tm.a[c][tid];
    s_v[c][tid] = tm.v[c][tid];
  }
  __syncthreads();
  const int i = blockIdx.x * PB + tid;
  const bool valid = i < M;
  float g[3] = {0.f, 0.f, 0.f};
  float dsum = 0.f;
  const float e = valid ? LN2 * expo[i] : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float lnx = valid ? x[3 * (size_t)i + c] + e : 0.f;
    float o = tm.d[c][0], slope = 0.f;
#pragma unroll 8
    for (int k = 0; k < TMW; ++k) {
      const float pre = fmaf(s_u[c][k], lnx, s_a[c][k]);
      o = fmaf(s_v[c][k], fmaxf(pre, 0.f), o);
      slope += (pre > 0.f) ? s_v[c][k] * s_u[c][k] : 0.f;
    }
    const float t = tanhf(o);
    g[c] = valid ? dldr[3 * (size_t)i + c] * (1.f - t * t) : 0.f;
    const float dl = g[c] * slope;
    if (valid) dx[3 * (size_t)i + c] = dl;
    dsum += dl;
    s_x[c][tid] = lnx;
    s_g[c][tid] = g[c];
  }
  if (valid && dexpo != nullptr) dexpo[i] = LN2 * dsum;
  __syncthreads();
  float* rec = part + (size_t)blockIdx.x * REC;
  {
    int np = M - blockIdx.x * PB;
    if (np > PB) np = PB;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ui = s_u[c][tid], ai = s_a[c][tid], vi = s_v[c][tid];
      float gu = 0.f, ga = 0.f, gv = 0.f;
      for (int q = 0; q < np; ++q) {
        const float xx = s_x[c][q], gg = s_g[c][q];
        const float pre = fmaf(ui, xx, ai);
        if (pre > 0.f) {
          gv = fmaf(gg, pre, gv);
          const float t = gg * vi;
          gu = fmaf(t, xx, gu);
          ga += t;
        }
      }
      rec[c * 3 * TMW + tid] = gu;
      rec[c * 3 * TMW + TMW + tid] = ga;
      rec[c * 3 * TMW + 2 * TMW + tid] = gv;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float r = g[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if ((tid & 31) == 0) s_red[tid >> 5][c] = r;
  }
  __syncthreads();
  if (tid < 3) {
    float t = 0.f;
    for (int w = 0; w < PB / 32; ++w) t += s_red[w][tid];
    rec[3 * 3 * TMW + tid] = t;
  }
}

int fill(TmP* tm, const float* const* p) {
  for (int c = 0; c < 3; ++c) {
    tm->u[c] = p[4 * c + 0]; tm->a[c] = p[4 * c + 1]; tm->v[c] = p[4 * c + 2]; tm->d[c] = p[4 * c + 3];
    if (!(tm->u[c] && tm->a[c] && tm->v[c] && tm->d[c])) return 1;
  }
  return 0;
}

}  // namespace

extern "C" int neucam_tm_fwd(const float* const* tm_params, int tm_width, const float* x, const float* expo, int64_t M,
                             float* ldr, void* stream) {
  NC_CHECK_ARG(tm_params && x && expo && ldr && M > 0 && M < (1ll << 31) - PB && tm_width == TMW);
  TmP tm;
  NC_CHECK_ARG(fill(&tm, tm_params) == 0);
  tm_fwd_kernel<<<(unsigned)((M + PB - 1) / PB), PB, 0, (cudaStream_t)stream>>>(tm, x, expo, (int)M, ldr);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int neucam_tm_bwd(const float* const* tm_params, int tm_width, const float* x, const float* expo,
                             const float* dldr, int64_t M, float* dx, float* dexpo, float* const* tm_grads, void* ws,
                             int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(tm_params && tm_grads && x && expo && dldr && dx && M > 0 && M < (1ll << 31) - PB && tm_width == TMW);
  TmP tm;
  NC_CHECK_ARG(fill(&tm, tm_params) == 0);
  const int grid = (int)((M + PB - 1) / PB);
  NC_CHECK_WS(ws, ws_bytes, (int64_t)grid * REC * 4);
  tm_bwd_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(tm, x, expo, dldr, (int)M, dx, dexpo,
                                                       reinterpret_cast<float*>(ws));
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  int n = 0;
  for (int c = 0; c < 3; ++c) {
    NC_CHECK_ARG(tm_grads[4 * c] && tm_grads[4 * c + 1] && tm_grads[4 * c + 2] && tm_grads[4 * c + 3]);
    segs.seg[n++] = {tm_grads[4 * c + 0], c * 3 * TMW, TMW};
    segs.seg[n++] = {tm_grads[4 * c + 1], c * 3 * TMW + TMW, TMW};
    segs.seg[n++] = {tm_grads[4 * c + 2], c * 3 * TMW + 2 * TMW, TMW};
    segs.seg[n++] = {tm_grads[4 * c + 3], 3 * 3 * TMW + c, 1};
  }
  segs.nseg = n;
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), grid, REC, segs, nullptr,
                                        (cudaStream_t)stream);
}
