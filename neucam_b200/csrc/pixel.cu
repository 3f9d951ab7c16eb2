// pixel.cu -- the per-pixel stages of NeuCam's step that surround the scene MLP (training.py:95-271).
// They touch O(B) data (B = pixels per step) against the O(K*B*512) of the atlas kernels, so they are
// plain SIMT kernels designed for few launches, coalesced [feature][pixel] scratch layouts and
// block-level reduction of parameter gradients before any global atomic.
//
//   blur_fwd   S2-S5: neighbour patch (utils.py:587-612), focus gather (training.py:104-108), blur-kernel
//              SimpleMLP 3->64->64->64->27 with softmax|tanh head (modules.py:282-306), defocus offsets
//              and centre-tap reset (training.py:115-121)            -> uv [K,B,2], kout [27,B], hidden
//   psf_crf_loss  S9-S12 forward AND backward: PSF-weighted sum (157), exposure gather 3*tanh(e) (161-166,
//              modules.py:407-408), TonemapNet (modules.py:219-247), image_mse (+ random gamma, utils.py:615-619,
//              training.py:180-183; + mask weights after freeze), ue_loss (244-245); emits dy [K,B,3] for the
//              atlas backward, d(kernel weights), CRF / exposure gradients, loss sums and max|dy|.
//   blur_bwd   backward of blur_fwd given d(kernel weights) and d(uv): softmax/tanh/ReLU-MLP backward,
//              focus gradient scatter.
#include "host_util.h"

namespace {

constexpr int KT = 9;           // taps (patch_size) supported by the fused path
constexpr int BH = 64;          // blur MLP hidden width
constexpr int BO = 3 * KT;      // blur MLP outputs
constexpr int PB = 128;         // pixels per block
constexpr int TMW = 128;        // TonemapNet hidden width
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// blur MLP forward
// ------------------------------------------------------------------------------------------------
struct BlurParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4;   // layers.{0..3}.{weight,bias}, weight = [out,in]
};

// dense 64->NOUT layer for one pixel per thread; input column in smem [BH][PB], weights in smem as
// [in][out] (so that 4 outputs are one broadcast LDS.128); accumulators in registers.
template <int NOUT>
__device__ __forceinline__ void dense_from_smem(const float* s_in, const float* s_wt, const float* s_b,
                                                int tid, float (&acc)[NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = s_b[o];
#pragma unroll 2
  for (int i = 0; i < BH; ++i) {
    const float a = s_in[i * PB + tid];
    const float* wr = s_wt + i * NOUT;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(wr[o], a, acc[o]);
  }
}

__global__ void __launch_bounds__(PB)
blur_fwd_kernel(const float* __restrict__ coords, const long long* __restrict__ coord_idx,
                const float* __restrict__ focus, BlurParams bp, int B, int HW, float sy, float sx,
                float offset_size, float* __restrict__ uv, float* __restrict__ kout,
                float* __restrict__ hid) {
  extern __shared__ float sm[];
  float* s_w2t = sm;                  // [64][64]  (in, out)
  float* s_w3t = s_w2t + BH * BH;     // [64][64]
  float* s_w4t = s_w3t + BH * BH;     // [64][28]  (padded to 28 outputs)
  float* s_w1 = s_w4t + BH * 28;      // [64][4]   (w1[o][0..2], b1[o])
  float* s_b = s_w1 + BH * 4;         // b2[64], b3[64], b4[28]
  float* s_h = s_b + 64 + 64 + 28;    // [64][PB]
  const int tid = threadIdx.x;
  for (int i = tid; i < BH * BH; i += PB) {
    const int o = i / BH, k = i % BH;
    s_w2t[k * BH + o] = bp.w2[i];
    s_w3t[k * BH + o] = bp.w3[i];
  }
  for (int i = tid; i < BH * 28; i += PB) {
    const int k = i / 28, o = i % 28;
    s_w4t[i] = (o < BO) ? bp.w4[o * BH + k] : 0.f;
  }
  for (int o = tid; o < BH; o += PB) {
    s_w1[o * 4 + 0] = bp.w1[o * 3 + 0];
    s_w1[o * 4 + 1] = bp.w1[o * 3 + 1];
    s_w1[o * 4 + 2] = bp.w1[o * 3 + 2];
    s_w1[o * 4 + 3] = bp.b1[o];
    s_b[o] = bp.b2[o];
    s_b[64 + o] = bp.b3[o];
  }
  for (int o = tid; o < 28; o += PB) s_b[128 + o] = (o < BO) ? bp.b4[o] : 0.f;
  __syncthreads();

  const int p = blockIdx.x * PB + tid;
  const bool valid = p < B;
  const int pc = valid ? p : (B - 1);
  const float ct = coords[3 * pc], cy = coords[3 * pc + 1], cx = coords[3 * pc + 2];
  (void)ct;
  const float f = focus[(int)(coord_idx[pc] / HW)];

  // layer 1 (3 -> 64) + ReLU
#pragma unroll 8
  for (int o = 0; o < BH; ++o) {
    const float4 w = *reinterpret_cast<const float4*>(s_w1 + o * 4);
    const float h = fmaxf(fmaf(w.x, f, fmaf(w.y, cy, fmaf(w.z, cx, w.w))), 0.f);
    s_h[o * PB + tid] = h;
    if (valid) hid[(size_t)(0 * BH + o) * B + p] = h;
  }
  float acc[BH];
  dense_from_smem<BH>(s_h, s_w2t, s_b, tid, acc);
#pragma unroll
  for (int o = 0; o < BH; ++o) {
    const float h = fmaxf(acc[o], 0.f);
    s_h[o * PB + tid] = h;
    if (valid) hid[(size_t)(1 * BH + o) * B + p] = h;
  }
  dense_from_smem<BH>(s_h, s_w3t, s_b + 64, tid, acc);
#pragma unroll
  for (int o = 0; o < BH; ++o) {
    const float h = fmaxf(acc[o], 0.f);
    s_h[o * PB + tid] = h;
    if (valid) hid[(size_t)(2 * BH + o) * B + p] = h;
  }
  float out[28];
  dense_from_smem<28>(s_h, s_w4t, s_b + 128, tid, out);

  // head: softmax over the K kernel weights, tanh over the 2K offsets (modules.py:303-306)
  float mx = out[0];
#pragma unroll
  for (int k = 1; k < KT; ++k) mx = fmaxf(mx, out[k]);
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    out[k] = __expf(out[k] - mx);
    den += out[k];
  }
  const float rden = 1.f / den;
#pragma unroll
  for (int k = 0; k < KT; ++k) out[k] *= rden;
#pragma unroll
  for (int k = KT; k < BO; ++k) out[k] = tanhf(out[k]);
  if (!valid) return;
#pragma unroll
  for (int k = 0; k < BO; ++k) kout[(size_t)k * B + p] = out[k];

  // tap coordinates: neighbour grid + learned offsets, centre tap = the pixel itself
#pragma unroll
  for (int k = 0; k < KT; ++k) {
    const int dy = k / 3 - 1, dx = k % 3 - 1;
    float y = cy + dy * sy, x = cx + dx * sx;
    y += out[KT + k] * sy * offset_size;
    x += out[2 * KT + k] * sx * offset_size;
    if (k == KT / 2) {
      y = cy;
      x = cx;
    }
    *reinterpret_cast<float2*>(uv + 2 * ((size_t)k * B + p)) = make_float2(y, x);
  }
}

// ------------------------------------------------------------------------------------------------
// PSF sum + exposure + CRF + loss, forward and backward
// ------------------------------------------------------------------------------------------------
struct TmParams {
  const float *u[3], *a[3], *v[3], *d[3];   // layers_{r,g,b}.0.weight [128,1], .0.bias [128], .2.weight [1,128], .2.bias [1]
  float *gu[3], *ga[3], *gv[3], *gd[3];     // gradients (accumulated into)
};

struct MidParams {
  int B, K, HW, N;
  int use_gamma, use_mask, has_blur;
  const float* gamma;     // device scalar: the step's random gamma (host RNG stream, training.py:181)
  float fixed_value;      // white-balance target (training.py:244)
  float ue_weight;        // 1/world_size: pixel-independent term counted once across ranks
  float inv_count;        // 1 / (3 * B_global)
  const float* y;         // [K,B,3] atlas output (log radiance)
  const float* kout;      // [27,B]   (rows 0..K-1 = PSF weights)
  const long long* coord_idx;
  const float* exps;      // [N] raw exposure parameters e
  const float* fixed_exps;  // [B] per-pixel exposure when not learned (or null)
  const float* gt;        // [B,3]
  const float* gamma_w;   // [B] or null
  const float* mask_w;    // [B] or null
  float* dy;              // [K,B,3]
  float* dkw;             // [K,B]   d loss / d PSF weight
  float* sums;            // [4]: sum sq err (img_loss * 3B), ue_loss, unused, unused  (accumulated)
  float* part;            // per-block partial records (layout: REC_* below), summed in block order by reduce.cu
  int rec;                // floats per record
  unsigned int* dy_absmax_bits;   // atomicMax target (max is order-independent)
  float* ldr_out;         // optional [B,3] tone-mapped prediction (for summaries), or null
};

// partial record of one block: CRF gradients gu | ga | gv per channel, then the scalars, then the per-frame exposure
// gradients
constexpr int REC_TM = 0;                 // [3 channels][gu 128 | ga 128 | gv 128]
constexpr int REC_GD = 3 * 3 * TMW;       // gd[3]
constexpr int REC_SQ = REC_GD + 3;        // sum of squared errors
constexpr int REC_DB4 = REC_GD + 4;       // db4[3]
constexpr int REC_DEXP = REC_GD + 8;      // dexps[N]
static_assert(REC_GD == nchost::WS_TM_GRADS, "workspace layout");

__global__ void __launch_bounds__(PB)
psf_crf_loss_kernel(MidParams mp, TmParams tm) {
  __shared__ float s_u[3][TMW], s_a[3][TMW], s_v[3][TMW];
  __shared__ float s_x[3][PB + 1], s_g[3][PB + 1];
  __shared__ float s_de[PB];
  __shared__ int s_fr[PB];
  __shared__ float s_red[PB / 32][8];
  const int tid = threadIdx.x;
  for (int c = 0; c < 3; ++c) {
    s_u[c][tid] = tm.u[c][tid];
    s_a[c][tid] = tm.a[c][tid];
    s_v[c][tid] = tm.v[c][tid];
  }
  s_fr[tid] = -1;
  s_de[tid] = 0.f;
  float* rec = mp.part + (size_t)blockIdx.x * mp.rec;
  __syncthreads();

  const int B = mp.B, K = mp.K;
  const int p = blockIdx.x * PB + tid;
  // one extra virtual pixel (index B) carries the white-balance term tm(0, 0) (training.py:244-245)
  const bool is_pixel = p < B;
  const bool is_ue = (p == B);

  float x[3] = {0.f, 0.f, 0.f};
  float kw[KT];
  float dt = 0.f, dtanh = 0.f;
  int frame = 0;
  if (is_pixel) {
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        kw[k] = mp.has_blur ? mp.kout[(size_t)k * B + p] : 1.f;
        const float* yy = mp.y + 3 * ((size_t)k * B + p);
        x[0] = fmaf(kw[k], yy[0], x[0]);
        x[1] = fmaf(kw[k], yy[1], x[1]);
        x[2] = fmaf(kw[k], yy[2], x[2]);
      }
    }
    if (mp.exps != nullptr) {
      frame = (int)(mp.coord_idx[p] / mp.HW);
      const float th = tanhf(mp.exps[frame]);
      dt = 3.f * th;
      dtanh = 3.f * (1.f - th * th);
    } else {
      dt = mp.fixed_exps[p];
    }
  }
  float lnx[3], ldr[3], g[3], dlnx[3] = {0.f, 0.f, 0.f};
  float sq = 0.f, ue = 0.f, amax = 0.f;
  float gam = 1.f, lg = 1.f;
  if (mp.use_gamma) {
    gam = *mp.gamma;
    lg = 1.f / logf(gam + 1.f);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lnx[c] = x[c] + LN2 * dt;
    float o = tm.d[c][0];
#pragma unroll 8
    for (int i = 0; i < TMW; ++i) o = fmaf(s_v[c][i], fmaxf(fmaf(s_u[c][i], lnx[c], s_a[c][i]), 0.f), o);
    ldr[c] = tanhf(o);
    float dl = 0.f;   // d loss / d ldr
    if (is_pixel) {
      const float gt = mp.gt[3 * (size_t)p + c];
      float fake = ldr[c], real = gt, dfake = 1.f;
      if (mp.use_gamma) {
        const float gw = mp.gamma_w[p];
        const float af = fmaf(0.5f * ldr[c] + 0.5f, gam, 1.f);
        const float ar = fmaf(0.5f * gt + 0.5f, gam, 1.f);
        fake += (2.f * logf(af) * lg - 1.f) * gw;
        real += (2.f * logf(ar) * lg - 1.f) * gw;
        dfake = 1.f + gw * gam * lg / af;
      }
      const float diff = fake - real;
      const float mw = mp.use_mask ? mp.mask_w[p] : 1.f;
      sq += mw * diff * diff;
      dl = 2.f * mw * diff * mp.inv_count * dfake;
      if (mp.ldr_out) mp.ldr_out[3 * (size_t)p + c] = ldr[c];
    } else if (is_ue) {
      const float e = ldr[c] - mp.fixed_value;
      ue += fabsf(e) * (1.f / 3.f);
      dl = (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) * (1.f / 3.f) * mp.ue_weight;
    }
    g[c] = dl * (1.f - ldr[c] * ldr[c]);
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < TMW; ++i) {
      const float pre = fmaf(s_u[c][i], lnx[c], s_a[c][i]);
      acc += (pre > 0.f) ? s_v[c][i] * s_u[c][i] : 0.f;
    }
    dlnx[c] = g[c] * acc;
    s_x[c][tid] = lnx[c];
    s_g[c][tid] = g[c];
  }
  if (is_pixel) {
    // gradients towards the atlas output and the PSF weights (training.py:157)
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (k < K) {
        const float* yy = mp.y + 3 * ((size_t)k * B + p);
        float* dd = mp.dy + 3 * ((size_t)k * B + p);
        const float d0 = kw[k] * dlnx[0], d1 = kw[k] * dlnx[1], d2 = kw[k] * dlnx[2];
        dd[0] = d0; dd[1] = d1; dd[2] = d2;
        amax = fmaxf(amax, fmaxf(fabsf(d0), fmaxf(fabsf(d1), fabsf(d2))));
        if (mp.has_blur) mp.dkw[(size_t)k * B + p] = dlnx[0] * yy[0] + dlnx[1] * yy[1] + dlnx[2] * yy[2];
      }
    }
    if (mp.exps != nullptr) {
      s_fr[tid] = frame;
      s_de[tid] = LN2 * (dlnx[0] + dlnx[1] + dlnx[2]) * dtanh;
    }
  }
  __syncthreads();
  // exposure gradients of this block: thread = frame, pixels in order (no atomics)
  if (mp.exps != nullptr) {
    for (int f0 = 0; f0 < mp.N; f0 += PB) {
      const int f = f0 + tid;
      if (f < mp.N) {
        float t = 0.f;
        for (int q = 0; q < PB; ++q)
          if (s_fr[q] == f) t += s_de[q];
        rec[REC_DEXP + f] = t;
      }
    }
  }

  // CRF parameter gradients: thread = hidden unit, loop over the block's pixels (no cross-thread reduce)
  {
    const int i = tid;
    int np = B + 1 - blockIdx.x * PB;
    if (np > PB) np = PB;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float ui = s_u[c][i], ai = s_a[c][i], vi = s_v[c][i];
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
      rec[REC_TM + c * 3 * TMW + i] = gu;
      rec[REC_TM + c * 3 * TMW + TMW + i] = ga;
      rec[REC_TM + c * 3 * TMW + 2 * TMW + i] = gv;
    }
  }
  // block reductions: d bias of the CRF output layer, loss sums, max |dy|
  float r0 = warp_sum((is_pixel || is_ue) ? g[0] : 0.f);
  float r1 = warp_sum((is_pixel || is_ue) ? g[1] : 0.f);
  float r2 = warp_sum((is_pixel || is_ue) ? g[2] : 0.f);
  float r3 = warp_sum(sq);
  float ksum = 0.f;
  if (is_pixel) {
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K) ksum += kw[k];
  }
  float r4 = warp_sum(ksum * dlnx[0]), r5 = warp_sum(ksum * dlnx[1]), r6 = warp_sum(ksum * dlnx[2]);
  if ((tid & 31) == 0) {
    s_red[tid >> 5][0] = r0; s_red[tid >> 5][1] = r1; s_red[tid >> 5][2] = r2; s_red[tid >> 5][3] = r3;
    s_red[tid >> 5][4] = r4; s_red[tid >> 5][5] = r5; s_red[tid >> 5][6] = r6;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((tid & 31) == 0 && amax > 0.f) atomicMax(mp.dy_absmax_bits, __float_as_uint(amax));
  __syncthreads();
  if (tid < 7) {
    float t = 0.f;
    for (int w = 0; w < PB / 32; ++w) t += s_red[w][tid];
    rec[REC_GD + tid] = t;     // gd[0..2] | sum sq | db4[0..2]
  }
  if (is_ue) mp.sums[1] += ue;   // the one virtual pixel of the grid: a single writer
}

// scale = 2^floor(log2(target / max|dy|)): keeps the scaled gradients of the atlas backward in fp16 range
__global__ void loss_scale_kernel(const unsigned int* absmax_bits, float target, float* scale_out) {
  const float m = __uint_as_float(*absmax_bits);
  float s = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(target / m, &e);            // target/m = f * 2^e, f in [0.5,1)
    s = ldexpf(1.f, e - 1);
    s = fminf(fmaxf(s, 9.5367431640625e-07f), 1.8446744e19f);
  }
  *scale_out = s;
}

// ------------------------------------------------------------------------------------------------
// blur MLP backward
// ------------------------------------------------------------------------------------------------
struct BlurGrads {
  float *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4, *focus;   // accumulated into (by the reduction pass)
};
// offsets of the tensors inside one block's partial record
constexpr int BR_W1 = 0, BR_B1 = BR_W1 + BH * 3, BR_W2 = BR_B1 + BH, BR_B2 = BR_W2 + BH * BH, BR_W3 = BR_B2 + BH,
              BR_B3 = BR_W3 + BH * BH, BR_W4 = BR_B3 + BH, BR_B4 = BR_W4 + BO * BH, BR_END = BR_B4 + BO;
static_assert(BR_END == nchost::WS_BLUR_PARAMS, "workspace layout");
constexpr int BR_FOCUS = (BR_END + 3) & ~3;

// Tiles of the blur backward live in smem as [pixel][feature] with a 68-float row pitch: a pixel-thread
// writes/reads its own row with conflict-free 128-bit accesses, and the block-level outer products read
// 4-feature vectors of a common pixel row.
constexpr int TP = 68;

// gw[o][i] += sum_q dz[q][o] * h[q][i]   (NOUT x NIN <= 64 x 64, NOUT % 8 == 0 or handled by guards, NIN % 4 == 0)
// thread t owns the 8 x 4 block  o in [8*(t/16), +8), i in [4*(t%16), +4).
template <int NOUT, int NIN>
__device__ __forceinline__ void block_outer_accum(const float* s_dz, const float* s_h, int np, float* gw, int tid,
                                                  bool first) {
  const int o0 = (tid >> 4) * 8, i0 = (tid & 15) * 4;
  if (o0 >= NOUT || i0 >= NIN) return;
  float acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 2
  for (int q = 0; q < np; ++q) {
    const float4 d0 = *reinterpret_cast<const float4*>(s_dz + q * TP + o0);
    const float4 d1 = *reinterpret_cast<const float4*>(s_dz + q * TP + o0 + 4);
    const float4 hv = *reinterpret_cast<const float4*>(s_h + q * TP + i0);
    const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dd[a], hh[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (o0 + a < NOUT && i0 + b < NIN) {   // this block's partial record: the same thread owns an entry for every group
        float* g = gw + (o0 + a) * NIN + i0 + b;
        *g = first ? acc[a][b] : *g + acc[a][b];
      }
}

// gb[o] += sum_q dz[q][o]
template <int NOUT>
__device__ __forceinline__ void block_bias_accum(const float* s_dz, int np, float* gb, int tid, bool first) {
  if (tid < NOUT) {
    float t = 0.f;
    for (int q = 0; q < np; ++q) t += s_dz[q * TP + tid];
    gb[tid] = first ? t : gb[tid] + t;
  }
}

__global__ void __launch_bounds__(PB)
blur_bwd_kernel(const float* __restrict__ coords, const long long* __restrict__ coord_idx,
                const float* __restrict__ focus, BlurParams bp, float* __restrict__ part, int rec_floats, int B, int HW,
                int N, float sy, float sx, float offset_size, const float* __restrict__ kout,
                const float* __restrict__ hid, const float* __restrict__ dkw, const float* __restrict__ duv) {
  extern __shared__ __align__(16) float sm[];
  // The weight matrices are read straight from global memory: every thread of a warp reads the same row element at
  // the same time (one broadcast load through L1), and keeping the 39 KB out of shared memory brings a block down to
  // 71 KB, so that a block fits next to a CTA of the weight-gradient GEMM this kernel runs concurrently with.
  const float* __restrict__ s_w2 = bp.w2;    // [64][64] (out, in) as stored
  const float* __restrict__ s_w3 = bp.w3;    // [64][64]
  const float* __restrict__ s_w4 = bp.w4;    // [27][64]
  float* s_dz = sm;                    // [PB][TP]  current layer's d(pre-activation)
  float* s_h = s_dz + PB * TP;         // [PB][TP]  that layer's input activation
  float* s_red = s_h + PB * TP;        // [PB]
  int* s_fr = reinterpret_cast<int*>(s_red + PB);   // [PB]
  const int tid = threadIdx.x;
  // Persistent: ONE block per SM walks the 128-pixel groups, so that this kernel occupies 71 KB and 4 warps of every SM
  // and the weight-gradient GEMM's CTA (145 KB) fits beside it.  The block adds its groups into ONE partial record
  // (every entry is owned by one thread, groups in order): the second pass sums gridDim.x records, not B / 128.
  const int ngroups = (B + PB - 1) / PB;
  float* rec = part + (size_t)blockIdx.x * rec_floats;
#pragma unroll 1
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
  const bool first = grp == (int)blockIdx.x;

  const int p0 = grp * PB;
  const int p = p0 + tid;
  const bool valid = p < B;
  int np = B - p0;
  if (np > PB) np = PB;
  float* my_dz = s_dz + tid * TP;
  float* my_h = s_h + tid * TP;

  // ---- d(logits) of the head ----
  float dlog[28];
  dlog[27] = 0.f;
  if (valid) {
    float w[KT], gsum = 0.f;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      w[k] = kout[(size_t)k * B + p];
      dlog[k] = dkw[(size_t)k * B + p];
      gsum = fmaf(dlog[k], w[k], gsum);
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) dlog[k] = w[k] * (dlog[k] - gsum);   // softmax backward
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float2 d = *reinterpret_cast<const float2*>(duv + 2 * ((size_t)k * B + p));
      const float ty = kout[(size_t)(KT + k) * B + p], tx = kout[(size_t)(2 * KT + k) * B + p];
      const bool centre = (k == KT / 2);   // centre tap is overwritten with the raw coords (training.py:120)
      dlog[KT + k] = centre ? 0.f : d.x * sy * offset_size * (1.f - ty * ty);
      dlog[2 * KT + k] = centre ? 0.f : d.y * sx * offset_size * (1.f - tx * tx);
    }
  } else {
#pragma unroll
    for (int k = 0; k < BO; ++k) dlog[k] = 0.f;
  }

  // ---- layer 4 (64 -> 27): dW4, db4, dh3 ----
#pragma unroll
  for (int o = 0; o < 28; o += 4) *reinterpret_cast<float4*>(my_dz + o) = make_float4(dlog[o], dlog[o + 1], dlog[o + 2], dlog[o + 3]);
#pragma unroll
  for (int o = 28; o < 32; o += 4) *reinterpret_cast<float4*>(my_dz + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  float hcur[BH];   // post-ReLU activation of the layer whose pre-activation gradient is being formed
#pragma unroll
  for (int i = 0; i < BH; ++i) hcur[i] = valid ? hid[(size_t)(2 * BH + i) * B + p] : 0.f;
#pragma unroll
  for (int i = 0; i < BH; i += 4) *reinterpret_cast<float4*>(my_h + i) = make_float4(hcur[i], hcur[i + 1], hcur[i + 2], hcur[i + 3]);
  __syncthreads();
  block_outer_accum<BO, BH>(s_dz, s_h, np, rec + BR_W4, tid, first);
  block_bias_accum<BO>(s_dz, np, rec + BR_B4, tid, first);
  float dh[BH];
#pragma unroll
  for (int i = 0; i < BH; ++i) dh[i] = 0.f;
#pragma unroll
  for (int o = 0; o < BO; ++o) {
    const float d = dlog[o];
    const float* wr = s_w4 + o * BH;
#pragma unroll
    for (int i = 0; i < BH; ++i) dh[i] = fmaf(wr[i], d, dh[i]);
  }
  __syncthreads();

  // ---- layers 3 and 2 (64 -> 64) ----
#pragma unroll 1
  for (int l = 2; l >= 1; --l) {
    float dz[BH];
#pragma unroll
    for (int i = 0; i < BH; ++i) dz[i] = (hcur[i] > 0.f) ? dh[i] : 0.f;   // ReLU backward
#pragma unroll
    for (int i = 0; i < BH; i += 4) *reinterpret_cast<float4*>(my_dz + i) = make_float4(dz[i], dz[i + 1], dz[i + 2], dz[i + 3]);
    // input activation of layer l+1 (0-based hidden index l-1)
#pragma unroll
    for (int i = 0; i < BH; ++i) hcur[i] = valid ? hid[(size_t)((l - 1) * BH + i) * B + p] : 0.f;
#pragma unroll
    for (int i = 0; i < BH; i += 4) *reinterpret_cast<float4*>(my_h + i) = make_float4(hcur[i], hcur[i + 1], hcur[i + 2], hcur[i + 3]);
    __syncthreads();
    float* gw = rec + ((l == 2) ? BR_W3 : BR_W2);
    float* gb = rec + ((l == 2) ? BR_B3 : BR_B2);
    const float* sw = (l == 2) ? s_w3 : s_w2;
    block_outer_accum<BH, BH>(s_dz, s_h, np, gw, tid, first);
    block_bias_accum<BH>(s_dz, np, gb, tid, first);
#pragma unroll
    for (int i = 0; i < BH; ++i) dh[i] = 0.f;
#pragma unroll 4
    for (int o = 0; o < BH; ++o) {
      const float d = my_dz[o];
      const float* wr = sw + o * BH;
#pragma unroll
      for (int i = 0; i < BH; ++i) dh[i] = fmaf(wr[i], d, dh[i]);
    }
    __syncthreads();
  }

  // ---- layer 1 (3 -> 64): dW1, db1, d focus ----
  const int pc = valid ? p : (B - 1);
  const float in0 = focus[(int)(coord_idx[pc] / HW)], in1 = coords[3 * pc + 1], in2 = coords[3 * pc + 2];
  float dfocus = 0.f;
  {
    float dz[BH];
#pragma unroll
    for (int i = 0; i < BH; ++i) {
      dz[i] = (valid && hcur[i] > 0.f) ? dh[i] : 0.f;
      dfocus = fmaf(bp.w1[i * 3], dz[i], dfocus);
    }
#pragma unroll
    for (int i = 0; i < BH; i += 4) *reinterpret_cast<float4*>(my_dz + i) = make_float4(dz[i], dz[i + 1], dz[i + 2], dz[i + 3]);
  }
  *reinterpret_cast<float4*>(my_h) = make_float4(in0, in1, in2, 0.f);
  __syncthreads();
  block_outer_accum<BH, 3>(s_dz, s_h, np, rec + BR_W1, tid, first);
  block_bias_accum<BH>(s_dz, np, rec + BR_B1, tid, first);
  // focus gradient: per-frame gather over the block's pixels in order (training.py:104-108 backward; no atomics)
  s_red[tid] = dfocus;
  s_fr[tid] = valid ? (int)(coord_idx[p] / HW) : -1;
  __syncthreads();
  for (int f0 = 0; f0 < N; f0 += PB) {
    const int f = f0 + tid;
    if (f < N) {
      float t = 0.f;
      for (int q = 0; q < PB; ++q)
        if (s_fr[q] == f) t += s_red[q];
      rec[BR_FOCUS + f] = first ? t : rec[BR_FOCUS + f] + t;
    }
  }
  __syncthreads();   // the next group rewrites the shared tiles
  }
}


// ------------------------------------------------------------------------------------------------
// on-device batch sampler / gather (replaces Implicit3DWrapper.__getitem__, dataio.py:356-393, for a video
// that is resident in HBM): B indices drawn WITH replacement (Philox-4x32-10 keyed by seed, counter = step,
// pixel) or taken from `idx_in`; coords recomputed from the index exactly like get_mgrid (dataio.py:24-45:
// float64 i/(dim-1), -0.5, *2, then float32; t uses max(N-1,1)); data / per-pixel weights gathered.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__global__ void __launch_bounds__(256)
sample_batch_kernel(const float* __restrict__ data, const float* __restrict__ gamma_w_all,
                    const float* __restrict__ mask_w_all, const long long* __restrict__ idx_in, int N, int H, int W,
                    int B, unsigned long long seed, unsigned long long step, float* __restrict__ coords,
                    long long* __restrict__ coord_idx, float* __restrict__ img, float* __restrict__ gamma_w,
                    float* __restrict__ mask_w) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B) return;
  const unsigned long long total = (unsigned long long)N * H * W;
  unsigned long long idx;
  if (idx_in != nullptr) {
    idx = (unsigned long long)idx_in[p];
  } else {
    uint32_t c[4] = {(uint32_t)p, (uint32_t)step, (uint32_t)(step >> 32), 0x4E435553u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const unsigned long long r64 = ((unsigned long long)c[0] << 32) | c[1];
    idx = __umul64hi(r64, total);   // floor(r64 * total / 2^64): unbiased to 2^-64 * total
  }
  const int hw = H * W;
  const int f = (int)(idx / hw), rem = (int)(idx % hw), yy = rem / W, xx = rem % W;
  const int nt = N - 1 > 1 ? N - 1 : 1;
  coords[3 * p + 0] = (float)(((double)f / nt - 0.5) * 2.0);
  coords[3 * p + 1] = (float)(((double)yy / (H - 1) - 0.5) * 2.0);
  coords[3 * p + 2] = (float)(((double)xx / (W - 1) - 0.5) * 2.0);
  coord_idx[p] = (long long)idx;
  img[3 * p + 0] = data[3 * idx + 0];
  img[3 * p + 1] = data[3 * idx + 1];
  img[3 * p + 2] = data[3 * idx + 2];
  if (gamma_w != nullptr && gamma_w_all != nullptr) gamma_w[p] = gamma_w_all[idx];
  if (mask_w != nullptr && mask_w_all != nullptr) mask_w[p] = mask_w_all[idx];
}


// ------------------------------------------------------------------------------------------------
// value-only gradient regulariser of the CRF (training.py:248-252 with modules.py:249-251): the minimum over M
// random (radiance, exposure) points of d mean(ldr) / d lnx.  The reference computes it with autograd.grad and no
// create_graph, so it is a constant w.r.t. the parameters: it only shows up in the logged loss.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tm_min_slope_kernel(TmParams tm, const float* __restrict__ rad, const float* __restrict__ expo, int M,
                    unsigned int* __restrict__ min_bits /* order-preserving encoding of the float minimum */) {
  __shared__ float s_u[3][TMW], s_a[3][TMW], s_v[3][TMW];
  const int tid = threadIdx.x;
  if (tid < TMW)
    for (int c = 0; c < 3; ++c) {
      s_u[c][tid] = tm.u[c][tid];
      s_a[c][tid] = tm.a[c][tid];
      s_v[c][tid] = tm.v[c][tid];
    }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + tid;
  float mn = 3.4e38f;
  if (i < M) {
    const float e = LN2 * expo[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = rad[3 * i + c] + e;
      float o = tm.d[c][0], sl = 0.f;
      for (int k = 0; k < TMW; ++k) {
        const float pre = fmaf(s_u[c][k], x, s_a[c][k]);
        if (pre > 0.f) {
          o = fmaf(s_v[c][k], pre, o);
          sl = fmaf(s_v[c][k], s_u[c][k], sl);
        }
      }
      const float t = tanhf(o);
      mn = fminf(mn, (1.f - t * t) * sl / (3.f * (float)M));   // d mean(ldr) / d lnx[i,c]
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((tid & 31) == 0) {
    unsigned int b = __float_as_uint(mn);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // monotone float -> uint
    atomicMin(min_bits, b);
  }
}

}  // namespace

extern "C" int neucam_tm_min_slope(const float* const* tm_params, const float* rad, const float* expo, int M,
                                   uint32_t* min_bits, void* stream) {
  NC_CHECK_ARG(tm_params && rad && expo && min_bits && M > 0);
  TmParams tm{};
  for (int c = 0; c < 3; ++c) {
    tm.u[c] = tm_params[4 * c + 0]; tm.a[c] = tm_params[4 * c + 1];
    tm.v[c] = tm_params[4 * c + 2]; tm.d[c] = tm_params[4 * c + 3];
  }
  NC_CHECK_CUDA(cudaMemsetAsync(min_bits, 0xFF, 4, (cudaStream_t)stream));
  tm_min_slope_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(tm, rad, expo, M, min_bits);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

namespace {
}  // namespace

extern "C" int neucam_sample_batch(const float* data, const float* gamma_w_all, const float* mask_w_all,
                                   const int64_t* idx_in, int N, int H, int W, int B, uint64_t seed, uint64_t step,
                                   float* coords, int64_t* coord_idx, float* img, float* gamma_w, float* mask_w,
                                   void* stream) {
  NC_CHECK_ARG(data && coords && coord_idx && img && N > 0 && H > 1 && W > 1 && B > 0);
  sample_batch_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      data, gamma_w_all, mask_w_all, (const long long*)idx_in, N, H, W, B, (unsigned long long)seed,
      (unsigned long long)step, coords, (long long*)coord_idx, img, gamma_w, mask_w);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

namespace {
}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int neucam_blur_fwd(const float* coords, const int64_t* coord_idx, const float* focus,
                               const float* const* blur_params, int B, int K, int N, int H, int W,
                               float offset_size, float* uv, float* kout, float* hid, void* stream) {
  NC_CHECK_ARG(coords && coord_idx && focus && blur_params && uv && kout && hid);
  NC_CHECK_ARG(K == KT && B > 0 && N > 0 && H > 1 && W > 1);
  BlurParams bp{blur_params[0], blur_params[1], blur_params[2], blur_params[3],
                blur_params[4], blur_params[5], blur_params[6], blur_params[7]};
  const size_t smem = sizeof(float) * (2 * BH * BH + BH * 28 + BH * 4 + 64 + 64 + 28 + BH * PB);
  int rc = nchost::ensure_dyn_smem((const void*)blur_fwd_kernel, (int)smem);
  if (rc) return rc;
  const int grid = (B + PB - 1) / PB;
  blur_fwd_kernel<<<grid, PB, smem, (cudaStream_t)stream>>>(
      coords, (const long long*)coord_idx, focus, bp, B, H * W, 2.f / (H - 1), 2.f / (W - 1), offset_size, uv,
      kout, hid);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

extern "C" int neucam_blur_bwd(const float* coords, const int64_t* coord_idx, const float* focus,
                               const float* const* blur_params, float* const* blur_grads, float* dfocus, int B,
                               int K, int N, int H, int W, float offset_size, const float* kout,
                               const float* hid, const float* dkw, const float* duv, void* ws, int64_t ws_bytes,
                               void* stream) {
  NC_CHECK_ARG(coords && coord_idx && focus && blur_params && blur_grads && dfocus && kout && hid && dkw && duv);
  NC_CHECK_ARG(K == KT && B > 0 && N > 0 && H > 1 && W > 1);
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_blur_bwd(B, N));
  BlurParams bp{blur_params[0], blur_params[1], blur_params[2], blur_params[3],
                blur_params[4], blur_params[5], blur_params[6], blur_params[7]};
  const size_t smem = sizeof(float) * (2 * PB * TP + 2 * PB);
  int rc = nchost::ensure_dyn_smem((const void*)blur_bwd_kernel, (int)smem);
  if (rc) return rc;
  rc = nchost::prefer_max_smem_carveout((const void*)blur_bwd_kernel);   // shares SMs with the weight-gradient GEMM
  if (rc) return rc;
  const int grid = (B + PB - 1) / PB;      // pixel groups; one partial record per launched block
  const int rec = nchost::ws_rec_blur_bwd(N);
  const int launch_blocks = grid < nchost::sm_count() ? grid : nchost::sm_count();
  blur_bwd_kernel<<<launch_blocks, PB, smem, (cudaStream_t)stream>>>(
      coords, (const long long*)coord_idx, focus, bp, reinterpret_cast<float*>(ws), rec, B, H * W, N, 2.f / (H - 1),
      2.f / (W - 1), offset_size, kout, hid, dkw, duv);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  const int offs[8] = {BR_W1, BR_B1, BR_W2, BR_B2, BR_W3, BR_B3, BR_W4, BR_B4};
  const int lens[8] = {BH * 3, BH, BH * BH, BH, BH * BH, BH, BO * BH, BO};
  for (int i = 0; i < 8; ++i) segs.seg[i] = {blur_grads[i], offs[i], lens[i]};
  segs.seg[8] = {dfocus, BR_FOCUS, N};
  segs.nseg = 9;
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), launch_blocks, rec, segs, nullptr,
                                        (cudaStream_t)stream);
}

extern "C" int neucam_psf_crf_loss(const float* y, const float* kout, const int64_t* coord_idx,
                                   const float* exps, const float* fixed_exps, const float* gt,
                                   const float* gamma_w, const float* mask_w, const float* const* tm_params,
                                   float* const* tm_grads, int B, int K, int N, int H, int W, int has_blur,
                                   const float* gamma_dev, float fixed_value, float ue_weight,
                                   int64_t global_count, float* dy, float* dkw, float* dexps, float* db4, float* sums,
                                   uint32_t* dy_absmax_bits, float scale_target, float* scale_out,
                                   float* ldr_out, int tm_width, void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(y && coord_idx && gt && tm_params && tm_grads && dy && db4 && sums && dy_absmax_bits && scale_out);
  NC_CHECK_ARG(tm_width == TMW);   // the kernels are built for TonemapNet(W=128), the only width any config uses
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_psf_crf(B, N));
  const int use_gamma = gamma_dev != nullptr;
  NC_CHECK_ARG((exps != nullptr) != (fixed_exps != nullptr));
  NC_CHECK_ARG(exps == nullptr || dexps != nullptr);
  NC_CHECK_ARG(!has_blur || (kout && dkw));
  NC_CHECK_ARG(!use_gamma || gamma_w);
  NC_CHECK_ARG(K >= 1 && K <= KT && B > 0 && global_count > 0);
  MidParams mp{};
  mp.B = B; mp.K = K; mp.HW = H * W; mp.N = N;
  mp.use_gamma = use_gamma; mp.use_mask = mask_w != nullptr; mp.has_blur = has_blur;
  mp.gamma = gamma_dev; mp.fixed_value = fixed_value; mp.ue_weight = ue_weight;
  mp.part = reinterpret_cast<float*>(ws); mp.rec = nchost::ws_rec_psf_crf(N);
  mp.inv_count = 1.f / (float)global_count;
  mp.y = y; mp.kout = kout; mp.coord_idx = (const long long*)coord_idx; mp.exps = exps;
  mp.fixed_exps = fixed_exps; mp.gt = gt; mp.gamma_w = gamma_w; mp.mask_w = mask_w;
  mp.dy = dy; mp.dkw = dkw; mp.sums = sums; mp.dy_absmax_bits = dy_absmax_bits;
  mp.ldr_out = ldr_out;
  TmParams tm{};
  for (int c = 0; c < 3; ++c) {
    tm.u[c] = tm_params[4 * c + 0]; tm.a[c] = tm_params[4 * c + 1];
    tm.v[c] = tm_params[4 * c + 2]; tm.d[c] = tm_params[4 * c + 3];
    tm.gu[c] = tm_grads[4 * c + 0]; tm.ga[c] = tm_grads[4 * c + 1];
    tm.gv[c] = tm_grads[4 * c + 2]; tm.gd[c] = tm_grads[4 * c + 3];
  }
  const int grid = (B + 1 + PB - 1) / PB;   // +1: the white-balance virtual pixel
  psf_crf_loss_kernel<<<grid, PB, 0, (cudaStream_t)stream>>>(mp, tm);
  NC_CHECK_CUDA(cudaGetLastError());
  loss_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dy_absmax_bits, scale_target, scale_out);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(2);
  // fixed-order sum of the blocks' records into the (accumulated) gradient tensors and loss sums
  nchost::ReduceSegs segs{};
  int n = 0;
  for (int c = 0; c < 3; ++c) {
    segs.seg[n++] = {tm.gu[c], REC_TM + c * 3 * TMW, TMW};
    segs.seg[n++] = {tm.ga[c], REC_TM + c * 3 * TMW + TMW, TMW};
    segs.seg[n++] = {tm.gv[c], REC_TM + c * 3 * TMW + 2 * TMW, TMW};
    segs.seg[n++] = {tm.gd[c], REC_GD + c, 1};
  }
  segs.seg[n++] = {sums, REC_SQ, 1};
  segs.seg[n++] = {db4, REC_DB4, 3};
  if (exps != nullptr) segs.seg[n++] = {dexps, REC_DEXP, N};
  segs.nseg = n;
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), grid, mp.rec, segs, nullptr,
                                        (cudaStream_t)stream);
}
