// adam.cu -- fused Adam over the flat fp32 parameter / gradient / moment buffers (torch.optim.Adam with
// the defaults the reference constructs it with, training.py:51: betas (0.9, 0.999), eps 1e-8, no weight
// decay, no amsgrad), plus the re-emission of the fp16 tensor-core operand copies of the atlas hidden
// weights (row-major [out,in] for the forward, transposed for the dgrad chain).
//
// HBM-bound: 28 B / parameter (read p, g, m, v; write p, m, v).
#include "host_util.h"
#include <cuda_fp16.h>

namespace {

// step bookkeeping lives on the device so that the whole training step can be replayed from a CUDA
// graph: state[0] = step count (as float, exact below 2^24), state[1] = 1 - beta1^t, state[2] = 1 - beta2^t
__global__ void adam_tick_kernel(float* state, float beta1, float beta2) {
  const double t = (double)state[0] + 1.0;
  state[0] = (float)t;
  state[1] = (float)(1.0 - pow((double)beta1, t));
  state[2] = (float)(1.0 - pow((double)beta2, t));
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, const float* __restrict__ state, float lr, float beta1, float beta2, float eps,
            float grad_scale) {
  const float bc1 = state[1], bc2 = state[2];
  const float step_size = lr / bc1;
  const float rsqrt_bc2 = 1.f / sqrtf(bc2);
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      ma[k] = ma[k] + (gk - ma[k]) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
      va[k] = va[k] * beta2 + (1.f - beta2) * gk * gk;          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
      const float denom = sqrtf(va[k]) * rsqrt_bc2 + eps;
      pa[k] = pa[k] - step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gk = g[i] * grad_scale;
    const float mk = m[i] + (gk - m[i]) * (1.f - beta1);
    const float vk = v[i] * beta2 + (1.f - beta2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    p[i] = p[i] - step_size * (mk / (sqrtf(vk) * rsqrt_bc2 + eps));
  }
}

// w: fp32 [512,512] row-major ([out,in]).  Writes fp16 copies: plain, and transposed * 30 (32x32 smem tiles).
__global__ void __launch_bounds__(256)
pack_hidden_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ w3,
                   __half* __restrict__ w16, __half* __restrict__ wT16) {
  __shared__ float tile[32][33];
  const float* w = blockIdx.z == 0 ? w1 : (blockIdx.z == 1 ? w2 : w3);
  __half* o = w16 + (size_t)blockIdx.z * 512 * 512;
  __half* oT = wT16 + (size_t)blockIdx.z * 512 * 512;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 8 * j;
    const float x = w[(size_t)r * 512 + c0 + tx];
    tile[ty + 8 * j][tx] = x;
    o[(size_t)r * 512 + c0 + tx] = __float2half_rn(x);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 8 * j;
    // the transposed copy feeds the dgrad chain, whose epilogue multiplies by 30 cos(.): the 30 is folded in here
    oT[(size_t)c * 512 + r0 + tx] = __float2half_rn(30.f * tile[tx][ty + 8 * j]);
  }
}

// head weights W4 [out_dim,512] -> fp16 [512,64]: row j = (W4[0][j], .., W4[out_dim-1][j], 0, ...): the K-major,
// zero-padded B operand of the backward's head dgrad (one K = 16 tensor-core step)
__global__ void pack_head_kernel(const float* __restrict__ w4, int out_dim, __half* __restrict__ w4t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 512 * 64
  if (i >= 512 * 64) return;
  const int j = i >> 6, c = i & 63;
  w4t[i] = __float2half_rn(c < out_dim ? w4[c * 512 + j] : 0.f);
}

}  // namespace

extern "C" int neucam_pack_head_weights(const float* w4, int out_dim, void* w4T16, void* stream) {
  NC_CHECK_ARG(w4 && w4T16 && out_dim >= 1 && out_dim <= 4);
  pack_head_kernel<<<128, 256, 0, (cudaStream_t)stream>>>(w4, out_dim, (__half*)w4T16);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int neucam_adam_tick(float* state, float beta1, float beta2, void* stream) {
  NC_CHECK_ARG(state != nullptr);
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, beta1, beta2);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int neucam_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* state,
                                float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  NC_CHECK_ARG(p && g && m && v && state && n > 0);
  NC_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)nchost::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (long long)n, state, lr, beta1, beta2, eps,
                                                            grad_scale);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int neucam_pack_hidden_weights(const float* w1, const float* w2, const float* w3, void* w16, void* wT16,
                                          void* stream) {
  NC_CHECK_ARG(w1 && w2 && w3 && w16 && wT16);
  pack_hidden_kernel<<<dim3(16, 16, 3), 256, 0, (cudaStream_t)stream>>>(w1, w2, w3, (__half*)w16, (__half*)wT16);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}
