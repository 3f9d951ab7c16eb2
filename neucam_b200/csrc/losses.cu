// losses.cu -- the regularisers of the layered configs as single-pass kernels (value AND gradient), and the
// device-side power-of-two loss scale the fp16 backward chains use.
//
//   rigidity loss   reference utils.get_rigidity_loss (utils.py:402-454, derivative_amount = 1), called at
//                   training.py:203-217: backward-difference Jacobian J = d(u,v)/d(x,y) of the uv mapping between a
//                   pixel and its upper / left neighbour, loss = |J^T J|_F + |(J^T J + 0.001 I)^-1|_F, in closed form
//                   for the 2x2 algebra with the analytic derivative.
#include "host_util.h"

namespace {

// thread = pixel.  uv_ref [B,2]; uv_moved [2,B,2] (pixel one row up, pixel one column left).
// NB the reference converts BOTH u differences with (H-1)/2 and BOTH v differences with (W-1)/2 (utils.py:425-428):
// ku, kv carry those factors divided by uv_mapping_scale.
__global__ void __launch_bounds__(256)
rigidity_kernel(const float* __restrict__ uv_ref, const float* __restrict__ uv_moved, int B, float ku, float kv,
                const float* __restrict__ coef, float* __restrict__ loss_sum, float* __restrict__ g_ref,
                float* __restrict__ g_moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float L = 0.f;
  if (i < B) {
    const float cf = *coef;   // weight * share / B (* mean mask weight), device scalar
    const float2 r = reinterpret_cast<const float2*>(uv_ref)[i];
    const float2 up = reinterpret_cast<const float2*>(uv_moved)[i];
    const float2 lf = reinterpret_cast<const float2*>(uv_moved)[(size_t)B + i];
    const float a = (r.x - lf.x) * ku, b = (r.x - up.x) * ku;   // du/dx, du/dy
    const float c = (r.y - lf.y) * kv, d = (r.y - up.y) * kv;   // dv/dx, dv/dy
    const float p = a * a + c * c, q = a * b + c * d, s = b * b + d * d;   // J^T J = [[p,q],[q,s]]
    const float pe = p + 0.001f, se = s + 0.001f;
    const float det = pe * se - q * q;
    const float n1 = sqrtf(p * p + 2.f * q * q + s * s);
    const float n2 = sqrtf(pe * pe + 2.f * q * q + se * se);
    const float idet = 1.f / det;
    L = n1 + n2 * idet;
    const float in1 = 1.f / fmaxf(n1, 1e-30f), in2 = 1.f / n2;
    const float Lp = p * in1 + pe * in2 * idet - n2 * se * idet * idet;
    const float Ls = s * in1 + se * in2 * idet - n2 * pe * idet * idet;
    const float Lq = 2.f * q * in1 + 2.f * q * in2 * idet + n2 * 2.f * q * idet * idet;
    const float La = 2.f * a * Lp + b * Lq, Lb = a * Lq + 2.f * b * Ls;
    const float Lc = 2.f * c * Lp + d * Lq, Ld = c * Lq + 2.f * d * Ls;
    reinterpret_cast<float2*>(g_ref)[i] = make_float2(cf * ku * (La + Lb), cf * kv * (Lc + Ld));
    reinterpret_cast<float2*>(g_moved)[i] = make_float2(-cf * ku * Lb, -cf * kv * Ld);
    reinterpret_cast<float2*>(g_moved)[(size_t)B + i] = make_float2(-cf * ku * La, -cf * kv * Lc);
    L *= cf;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
  __shared__ float s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = L;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(loss_sum, t);
  }
}

__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ bits) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));   // non-negative floats order as uints
}

// scale = 2^floor(log2(target / max|x|)), clamped to [2^-20, 2^64]; 1 when max is 0 or not finite
__global__ void scale_from_bits_kernel(const unsigned int* bits, float target, float* scale_out) {
  const float m = __uint_as_float(*bits);
  float s = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(target / m, &e);
    s = ldexpf(1.f, e - 1);
    s = fminf(fmaxf(s, 9.5367431640625e-07f), 1.8446744e19f);
  }
  *scale_out = s;
}

}  // namespace

// loss_sum[0] += coef * sum_i L_i;  g_ref [B,2] = coef * dL/d uv_ref;  g_moved [2,B,2] = coef * dL/d uv_moved.
// ku = (H-1)/2 / uv_mapping_scale, kv = (W-1)/2 / uv_mapping_scale; coef: device scalar.
extern "C" int neucam_rigidity_loss(const float* uv_ref, const float* uv_moved, int B, float ku, float kv,
                                    const float* coef_dev, float* loss_sum, float* g_ref, float* g_moved, void* stream) {
  NC_CHECK_ARG(uv_ref && uv_moved && coef_dev && loss_sum && g_ref && g_moved && B > 0);
  rigidity_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(uv_ref, uv_moved, B, ku, kv, coef_dev, loss_sum,
                                                                    g_ref, g_moved);
  NC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// scale_out[0] = power-of-two loss scale that brings max|x| into [target/2, target); scratch_bits: 4 bytes.
extern "C" int neucam_loss_scale(const float* x, int64_t n, float target, uint32_t* scratch_bits, float* scale_out,
                                 void* stream) {
  NC_CHECK_ARG(x && scratch_bits && scale_out && n > 0 && target > 0.f);
  NC_CHECK_CUDA(cudaMemsetAsync(scratch_bits, 0, 4, (cudaStream_t)stream));
  int grid = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (grid > 1184) grid = 1184;
  absmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, scratch_bits);
  NC_CHECK_CUDA(cudaGetLastError());
  scale_from_bits_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(scratch_bits, target, scale_out);
  NC_CHECK_CUDA(cudaGetLastError());
  return 0;
}
