// losses.cu -- the regularisers of the layered configs as single-pass kernels (value AND gradient), and the
// device-side power-of-two loss scale the fp16 backward chains use.
//
//   layer blend     (1 - alpha) back + alpha front with the sparsity / alpha regularisers (training.py:143-153,
//                   220-238), forward and backward kernels (second half of this file)
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
    loss_sum[4 * blockIdx.x] = t;   // this block's partial record; summed in block order by reduce.cu
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
                                    const float* coef_dev, float* loss_sum, float* g_ref, float* g_moved, void* ws,
                                    int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(uv_ref && uv_moved && coef_dev && loss_sum && g_ref && g_moved && B > 0);
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_rigidity(B));
  const int grid = (B + 255) / 256;
  rigidity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(uv_ref, uv_moved, B, ku, kv, coef_dev,
                                                           reinterpret_cast<float*>(ws), g_ref, g_moved);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  segs.nseg = 1;
  segs.seg[0] = {loss_sum, 0, 1};
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), grid, 4, segs, nullptr,
                                        (cudaStream_t)stream);
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
  NC_COUNT_LAUNCH(2);
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Front of a ReLU MLP's backward (relu_mlp.py): dpre = dy * d(out)/d(pre) from the forward output y (tanh: 1 - y^2,
// sigmoid: y (1 - y), none: 1), the output-bias gradient db_last[c] += sum_r dpre[r,c] and the power-of-two loss scale of
// max|dpre| in one pass over [Q, out_dim] -- it replaces three pointwise torch kernels, an abs-max pass and a column
// reduction (13-20 us each at Q = 55 k..166 k) per network call.
// ------------------------------------------------------------------------------------------------------------------
namespace {

template <int NC>
__global__ void __launch_bounds__(256)
head_prep_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long Q, int out_kind,
                 float* __restrict__ dpre, float* __restrict__ part, unsigned int* __restrict__ bits) {
  __shared__ float s_red[8][NC];
  float sum[NC], m = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) sum[c] = 0.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < Q; r += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float g = dy[r * NC + c];
      float d = g;
      if (out_kind != 0) {
        const float o = y[r * NC + c];
        d = out_kind == 1 ? g * (1.f - o * o) : g * o * (1.f - o);
        dpre[r * NC + c] = d;
      }
      sum[c] += d;
      m = fmaxf(m, fabsf(d));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
#pragma unroll
    for (int c = 0; c < NC; ++c) sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    if (m > 0.f) atomicMax(bits, __float_as_uint(m));   // non-negative floats order as uints; max is order-independent
#pragma unroll
    for (int c = 0; c < NC; ++c) s_red[w][c] = sum[c];
  }
  __syncthreads();
  if (threadIdx.x < NC) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_red[k][threadIdx.x];
    part[(size_t)blockIdx.x * 4 + threadIdx.x] = t;   // record of this block; reduce.cu adds them in block order
  }
}

}  // namespace

extern "C" int neucam_relu_mlp_head_prep(const float* dy, const float* y, int out_dim, int out_kind, int64_t Q,
                                         float target, float* dpre, float* db_last, uint32_t* scratch_bits,
                                         float* scale_out, void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(dy && db_last && scratch_bits && scale_out && Q > 0 && target > 0.f);
  NC_CHECK_ARG(out_dim >= 1 && out_dim <= 4 && out_kind >= 0 && out_kind <= 2);
  NC_CHECK_ARG(out_kind == 0 || (y && dpre));   // out_kind 0: dpre = dy, nothing is read from y or written to dpre
  int grid = (int)((Q + 255) / 256);
  if (grid > 2 * nchost::sm_count()) grid = 2 * nchost::sm_count();
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_head_prep());
  cudaStream_t st = (cudaStream_t)stream;
  NC_CHECK_CUDA(cudaMemsetAsync(scratch_bits, 0, 4, st));
  float* part = reinterpret_cast<float*>(ws);
  switch (out_dim) {
    case 1: head_prep_kernel<1><<<grid, 256, 0, st>>>(dy, y, Q, out_kind, dpre, part, scratch_bits); break;
    case 2: head_prep_kernel<2><<<grid, 256, 0, st>>>(dy, y, Q, out_kind, dpre, part, scratch_bits); break;
    case 3: head_prep_kernel<3><<<grid, 256, 0, st>>>(dy, y, Q, out_kind, dpre, part, scratch_bits); break;
    default: head_prep_kernel<4><<<grid, 256, 0, st>>>(dy, y, Q, out_kind, dpre, part, scratch_bits); break;
  }
  NC_CHECK_CUDA(cudaGetLastError());
  scale_from_bits_kernel<<<1, 1, 0, st>>>(scratch_bits, target, scale_out);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(2);
  nchost::ReduceSegs segs{};
  segs.nseg = 1;
  segs.seg[0] = {db_last, 0, out_dim};
  return nchost::launch_reduce_partials(part, grid, 4, segs, nullptr, st);
}

// ------------------------------------------------------------------------------------------------------------------
// layer blend + sparsity / alpha regularisers (reference training.py:143-153 and 220-234, alpha from the alpha MLP)
//   out      = (1 - a) * back + a * front                      per tap (k, b)
//   sparse1  = w1 * mean_{k,b} (1 - a_ref[b])^2 * sum_c exp(front[k,b,c])^2          (|(1-a) e^front|^2, 226-227)
//   sparse2  =      mean_b  -1 / (log(a_ref + 1e-24) + log(1 - a_ref + 1e-24))        (231; the reference multiplies
//                                                                                      by the boolean flag = 1)
//   alpha    = w3 * mean_b a_ref[b]                                                    (236-238)
// a_ref = alpha of the centre tap (training.py:153).  The caller folds the means and the data-parallel share into
// c1 = w1 * share / (K B), c2 = share / B, c3 = w3 * share / B (0 switches a term off).
// ------------------------------------------------------------------------------------------------------------------
namespace {

// thread = tap (k, b).  sums[0..2] += this block's share of sparse1, sparse2, alpha.
__global__ void __launch_bounds__(256)
blend_fwd_kernel(const float* __restrict__ back, const float* __restrict__ front, const float* __restrict__ alpha,
                 int K, int B, float c1, float c2, float c3, float* __restrict__ out, float* __restrict__ e2,
                 float* __restrict__ sums) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)K * B;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    const int k = (int)(i / B), b = (int)(i % B);
    const float a = alpha[i];
    const float a_ref = alpha[(long long)(K / 2) * B + b];
    const float* fb = back + 3 * i;
    const float* ff = front + 3 * i;
    float e = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out[3 * i + c] = (1.f - a) * fb[c] + a * ff[c];
      const float x = expf(ff[c]);
      e = fmaf(x, x, e);
    }
    e2[i] = e;
    s1 = c1 * (1.f - a_ref) * (1.f - a_ref) * e;
    if (k == K / 2) {
      s2 = c2 * (-1.f / (logf(a_ref + 1e-24f) + logf(1.f - a_ref + 1e-24f)));
      s3 = c3 * a_ref;
    }
  }
  __shared__ float s_part[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_part[0][threadIdx.x >> 5] = s1;
    s_part[1][threadIdx.x >> 5] = s2;
    s_part[2][threadIdx.x >> 5] = s3;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[threadIdx.x][w];
    sums[4 * blockIdx.x + threadIdx.x] = t;   // this block's partial record; summed in block order by reduce.cu
  }
}

// gradients of  <dout, out> + g * (sparse1 + sparse2 + alpha)  w.r.t. back, front and alpha
__global__ void __launch_bounds__(256)
blend_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ back, const float* __restrict__ front,
                 const float* __restrict__ alpha, const float* __restrict__ e2, const float* __restrict__ g_reg,
                 int K, int B, float c1, float c2, float c3, float* __restrict__ d_back, float* __restrict__ d_front,
                 float* __restrict__ d_alpha) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * B) return;
  const int k = (int)(i / B), b = (int)(i % B);
  const float g = *g_reg;
  const float a = alpha[i];
  const float a_ref = alpha[(long long)(K / 2) * B + b];
  const float w = g * c1 * (1.f - a_ref) * (1.f - a_ref);
  float da = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float dy = dout[3 * i + c], fb = back[3 * i + c], ff = front[3 * i + c];
    d_back[3 * i + c] = (1.f - a) * dy;
    const float x = expf(ff);
    d_front[3 * i + c] = a * dy + w * 2.f * x * x;
    da = fmaf(ff - fb, dy, da);
  }
  if (k == K / 2) {
    float esum = 0.f;
    for (int kk = 0; kk < K; ++kk) esum += e2[(long long)kk * B + b];
    const float l = logf(a_ref + 1e-24f) + logf(1.f - a_ref + 1e-24f);
    const float ds2 = (1.f / (a_ref + 1e-24f) - 1.f / (1.f - a_ref + 1e-24f)) / (l * l);
    da += g * (-2.f * c1 * (1.f - a_ref) * esum + c2 * ds2 + c3);
  }
  d_alpha[i] = da;
}

}  // namespace

// back, front, out [K,B,3]; alpha [K,B]; e2 [K,B] scratch kept for the backward; sums[3] += (sparse1, sparse2, alpha).
extern "C" int neucam_layer_blend_fwd(const float* back, const float* front, const float* alpha, int K, int B,
                                      float c1, float c2, float c3, float* out, float* e2, float* sums, void* ws,
                                      int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(back && front && alpha && out && e2 && sums && K >= 1 && B > 0);
  const long long n = (long long)K * B;
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_blend(n));
  const int grid = (int)((n + 255) / 256);
  blend_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(back, front, alpha, K, B, c1, c2, c3, out, e2,
                                                            reinterpret_cast<float*>(ws));
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  segs.nseg = 1;
  segs.seg[0] = {sums, 0, 3};
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), grid, 4, segs, nullptr,
                                        (cudaStream_t)stream);
}

// dout [K,B,3] = dL/d out; *g_reg (device scalar) = dL/d(sparse1 + sparse2 + alpha).  Writes d_back, d_front [K,B,3]
// and d_alpha [K,B].
extern "C" int neucam_layer_blend_bwd(const float* dout, const float* back, const float* front, const float* alpha,
                                      const float* e2, const float* g_reg, int K, int B, float c1, float c2, float c3,
                                      float* d_back, float* d_front, float* d_alpha, void* stream) {
  NC_CHECK_ARG(dout && back && front && alpha && e2 && g_reg && d_back && d_front && d_alpha && K >= 1 && B > 0);
  const long long n = (long long)K * B;
  blend_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dout, back, front, alpha, e2, g_reg, K,
                                                                                   B, c1, c2, c3, d_back, d_front,
                                                                                   d_alpha);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}
