// relu_aux.cu -- the two HBM-bound helpers around relu_chain.cu:
//   * input encoding: (t,y,x) -> the 64-feature fp16 operand tile [raw coords + NeRF positional encoding
//     (reference modules.py:567-580: per frequency i and coordinate j: sin(2^i pi c_j), cos(2^i pi c_j)) as fp16
//     "hi" values | their fp16 rounding residuals "lo" | zero padding], and its backward (d/d coords);
//   * weight packing: the fp32 master weights -> the fp16 hi / lo K-block "stages" the forward and backward chains
//     stream (one launch for the whole network instead of dozens of small tensor ops).
#include <cuda_fp16.h>

#include "host_util.h"

namespace {

constexpr int FEAT = 64;
constexpr int WID = 256;
constexpr float PI_F = 3.14159265358979323846f;

// thread = sample.  NF (number of frequencies) is a template parameter: with the encoded width E known at compile time
// the 64 output halves are assembled in registers (a run-time E indexes the arrays dynamically and sends them to local
// memory: 18 us per call at Q = 165 888 for a 21 MB write).
template <int NF>
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ x, long long Q, __half* __restrict__ feat) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  constexpr int E = 3 + 6 * NF;
  float v[E];
  const float c[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
  v[0] = c[0]; v[1] = c[1]; v[2] = c[2];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const float w = (float)((double)(1 << i) * 3.14159265358979323846);   // fp32(2^i pi), as torch multiplies
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float a = w * c[j];
      sincosf(a, &v[3 + 6 * i + 2 * j], &v[3 + 6 * i + 2 * j + 1]);
    }
  }
  __half out[FEAT];
#pragma unroll
  for (int j = 0; j < FEAT; ++j) out[j] = __float2half_rn(0.f);
#pragma unroll
  for (int j = 0; j < E; ++j) {
    const __half hi = __float2half_rn(v[j]);
    out[j] = hi;
    out[E + j] = __float2half_rn(v[j] - __half2float(hi));
  }
  uint4* dst = reinterpret_cast<uint4*>(feat + q * FEAT);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t w4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      w4[k] = (uint32_t)__half_as_ushort(out[8 * j + 2 * k]) | ((uint32_t)__half_as_ushort(out[8 * j + 2 * k + 1]) << 16);
    dst[j] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

// dx_j = dfeat[j] + sum_i 2^i pi (cos(a) dfeat[sin col] - sin(a) dfeat[cos col]);  only the hi columns carry gradient.
// NF is a template parameter so that the E = 3 + 6 NF gradient values of a row are requested up front (as 16-byte loads
// where the row allows) instead of one by one between the sincosf calls.
template <int NF>
__global__ void __launch_bounds__(256)
encode_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dfeat, long long Q, float* __restrict__ dx) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  constexpr int E = 3 + 6 * NF;
  constexpr int E4 = (E + 3) / 4;
  float g[4 * E4];
  const float4* g4 = reinterpret_cast<const float4*>(dfeat + q * FEAT);   // rows are 256-byte aligned
#pragma unroll
  for (int k = 0; k < E4; ++k) {
    const float4 t = g4[k];
    g[4 * k] = t.x; g[4 * k + 1] = t.y; g[4 * k + 2] = t.z; g[4 * k + 3] = t.w;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float c = x[3 * q + j];
    float acc = g[j];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const float w = (float)((double)(1 << i) * 3.14159265358979323846);
      const float a = w * c;
      float sn, cs;
      sincosf(a, &sn, &cs);
      acc = fmaf(w, cs * g[3 + 6 * i + 2 * j] - sn * g[3 + 6 * i + 2 * j + 1], acc);
    }
    dx[3 * q + j] = acc;
  }
}

struct StageDesc {
  const float* w;   // fp32 weight [rows, ld] (torch Linear layout: [out, in])
  int ld;           // its row pitch
  int col0;         // first source column (forward) / unused (backward)
  short kind;       // 0: forward K-block (cols col0 + kb*64 ..), 1: forward encoded-input block, 2: backward K-block
  short kb;         // K-block index (kinds 0, 2)
  short lo;         // 0: hi part, 1: lo part
  short E;          // encoded width (kind 1: columns [0,E) and, for the hi part, their copy at [E,2E))
  int in_rows;      // kind 2: number of valid input features (rows of the transposed weight)
};
constexpr int MAX_DESC = 160;
struct PackPlan {
  StageDesc d[MAX_DESC];
};

// 8 CTAs per stage (blockIdx.y = 32-row slice for the forward blocks / 8-column slice for the transposed ones):
// out[256 rows][64 cols] fp16
__global__ void __launch_bounds__(256)
pack_kernel(const __grid_constant__ PackPlan plan, int n_fwd, __half* __restrict__ w_stages,
            __half* __restrict__ wT_stages) {
  const StageDesc& d = plan.d[blockIdx.x];
  __half* out = (blockIdx.x < n_fwd ? w_stages + (size_t)blockIdx.x * WID * 64
                                    : wT_stages + (size_t)(blockIdx.x - n_fwd) * WID * 64);
  const int e0 = blockIdx.y * (WID * 64 / 8);
  for (int e = e0 + threadIdx.x; e < e0 + WID * 64 / 8; e += 256) {
    int n, k;
    float v = 0.f;
    if (d.kind == 2) {
      // transposed: row n = input feature, column k = output neuron kb*64 + k; the contiguous SOURCE direction is n
      // (W[j][n]), so consecutive threads walk n
      n = e % WID; k = e / WID;
      if (n < d.in_rows) v = d.w[(size_t)(d.kb * 64 + k) * d.ld + n];
    } else {
      n = e / 64; k = e % 64;
      if (d.kind == 0) {
        v = d.w[(size_t)n * d.ld + d.col0 + d.kb * 64 + k];
      } else if (k < d.E) {
        v = d.w[(size_t)n * d.ld + d.col0 + k];
      } else if (k < 2 * d.E && d.lo == 0) {
        v = d.w[(size_t)n * d.ld + d.col0 + k - d.E];
      }
    }
    const __half hi = __float2half_rn(v);
    out[(size_t)n * 64 + k] = d.lo ? __float2half_rn(v - __half2float(hi)) : hi;
  }
}

}  // namespace

// feat16 [Q,64] fp16 from coordinates x [Q,3] fp32; num_freq = 0: no positional encoding (E = 3), else E = 3 + 6 nf.
extern "C" int neucam_relu_mlp_encode(const float* x, int64_t Q, int num_freq, void* feat16, void* stream) {
  NC_CHECK_ARG(x && feat16 && Q > 0 && num_freq >= 0 && 2 * (3 + 6 * num_freq) <= FEAT);
  const unsigned grid = (unsigned)((Q + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (num_freq) {
    case 0: encode_kernel<0><<<grid, 256, 0, st>>>(x, Q, (__half*)feat16); break;
    case 1: encode_kernel<1><<<grid, 256, 0, st>>>(x, Q, (__half*)feat16); break;
    case 2: encode_kernel<2><<<grid, 256, 0, st>>>(x, Q, (__half*)feat16); break;
    case 3: encode_kernel<3><<<grid, 256, 0, st>>>(x, Q, (__half*)feat16); break;
    default: encode_kernel<4><<<grid, 256, 0, st>>>(x, Q, (__half*)feat16); break;   // 2 * (3 + 6 nf) <= 64: nf <= 4
  }
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

// dx [Q,3] from dfeat [Q,64] fp32 (the gradient w.r.t. the encoded input written by neucam_relu_mlp_bwd).
extern "C" int neucam_relu_mlp_encode_bwd(const float* x, const float* dfeat, int64_t Q, int num_freq, float* dx,
                                          void* stream) {
  NC_CHECK_ARG(x && dfeat && dx && Q > 0 && num_freq >= 0 && 2 * (3 + 6 * num_freq) <= FEAT);
  NC_CHECK_ARG(((uintptr_t)dfeat & 15) == 0);
  const unsigned grid = (unsigned)((Q + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (num_freq) {
    case 0: encode_bwd_kernel<0><<<grid, 256, 0, st>>>(x, dfeat, Q, dx); break;
    case 1: encode_bwd_kernel<1><<<grid, 256, 0, st>>>(x, dfeat, Q, dx); break;
    case 2: encode_bwd_kernel<2><<<grid, 256, 0, st>>>(x, dfeat, Q, dx); break;
    case 3: encode_bwd_kernel<3><<<grid, 256, 0, st>>>(x, dfeat, Q, dx); break;
    default: encode_bwd_kernel<4><<<grid, 256, 0, st>>>(x, dfeat, Q, dx); break;
  }
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

// Packs one ReLU MLP's weights.  w[l]: device pointers of the L Linear weights (fp32 [out,in_l], row pitch ld[l]);
// forward plan per stage s < S: st_layer[s], st_kb[s] (0..3 K-block of the 256 hidden inputs, 4 = the encoded-input
// block), st_lo[s] (0 hi / 1 lo part).  Backward stages (fixed order: layers H-1..1, then 0 if with_input_grad; per
// layer 4 K-blocks x (hi, lo)) go to wT_stages.  E = encoded width.
extern "C" int neucam_relu_mlp_pack(const float* const* w, const int* ld, int L, int E, int S, const uint8_t* st_layer,
                                    const uint8_t* st_kb, const uint8_t* st_lo, int with_input_grad, void* w_stages,
                                    void* wT_stages, void* stream) {
  NC_CHECK_ARG(w && ld && st_layer && st_kb && st_lo && w_stages && L >= 3 && L <= 9 && E >= 3 && 2 * E <= FEAT);
  const int H = L - 1;
  const int n_bwd = wT_stages ? 8 * ((H - 1) + (with_input_grad ? 1 : 0)) : 0;
  NC_CHECK_ARG(S >= 1 && S + n_bwd <= MAX_DESC);
  PackPlan plan;
  for (int s = 0; s < S; ++s) {
    const int l = st_layer[s];
    NC_CHECK_ARG(l < H && st_kb[s] <= 4 && st_lo[s] <= 1 && w[l]);
    StageDesc& d = plan.d[s];
    d.w = w[l]; d.ld = ld[l]; d.lo = st_lo[s]; d.E = (short)E; d.in_rows = 0;
    if (st_kb[s] == 4) { d.kind = 1; d.kb = 0; d.col0 = (l == 0) ? 0 : WID; }
    else { NC_CHECK_ARG(l > 0); d.kind = 0; d.kb = st_kb[s]; d.col0 = 0; }
  }
  int n = S;
  for (int g = 0; g * 8 < n_bwd; ++g) {
    const int l = (g < H - 1) ? (H - 1 - g) : 0;
    for (int s = 0; s < 8; ++s) {
      StageDesc& d = plan.d[n++];
      d.w = w[l]; d.ld = ld[l]; d.col0 = 0; d.kind = 2; d.kb = (short)(s >> 1); d.lo = (short)(s & 1); d.E = (short)E;
      d.in_rows = (l == 0) ? E : WID;
    }
  }
  pack_kernel<<<dim3(n, 8), 256, 0, (cudaStream_t)stream>>>(plan, S, (__half*)w_stages, (__half*)wT_stages);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}
