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

// The blur MLP runs as register-tiled SIMT GEMMs over 64-pixel groups: 256 threads = 16 x 16 tiles of 4 x 4 outputs, so
// that one 16-byte shared-memory load of each operand feeds 16 FMAs (a thread-per-pixel layout needs a load per 4 FMAs
// and leaves an SM with 4-9 warps at B = 42 525 pixels).  Activation tiles live in shared memory as [feature 64][pixel
// 64] floats; the sixteen 16-byte pixel quads of a row are XOR-swizzled with the row's feature quad, which makes every
// access pattern below conflict-free: a warp reading one row's quads (GEMM over features), sixteen rows' quad g (GEMM
// over pixels) or 32 consecutive pixels of a row (coalesced copies to / from global memory).
constexpr int GP = 64;          // pixels per group
constexpr int BT = 256;         // threads per block
__device__ __forceinline__ int tile_off(int f, int g) { return f * GP + (((g ^ (f >> 2)) & 15) << 2); }
__device__ __forceinline__ int tile_at(int f, int px) { return tile_off(f, px >> 2) + (px & 3); }

// out[px][o] = bias[o] + sum_i in[i][px] * w[o][i] for px in [4 tx, +4), o in [4 ty, +4); w = [out][in] as stored (no
// transposition: the loop walks the inputs four at a time, four 16-byte loads of each operand per 64 FMAs).  The sum
// runs bias first, then i ascending (the order of a plain per-pixel loop).
__device__ __forceinline__ void fwd_tile(const float* s_in, const float* s_w, const float* s_bias, int tx, int ty,
                                         float (&acc)[4][4]) {
  const float4 bv = *reinterpret_cast<const float4*>(s_bias + 4 * ty);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    acc[a][0] = bv.x; acc[a][1] = bv.y; acc[a][2] = bv.z; acc[a][3] = bv.w;
  }
#pragma unroll 2
  for (int i0 = 0; i0 < BH; i0 += 4) {
    float4 w[4], h[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) w[b] = *reinterpret_cast<const float4*>(s_w + (4 * ty + b) * BH + i0);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = *reinterpret_cast<const float4*>(s_in + tile_off(i0 + j, tx));
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float ww[4] = {w[b].x, w[b].y, w[b].z, w[b].w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float t = acc[a][b];
        t = fmaf(ww[0], a == 0 ? h[0].x : a == 1 ? h[0].y : a == 2 ? h[0].z : h[0].w, t);
        t = fmaf(ww[1], a == 0 ? h[1].x : a == 1 ? h[1].y : a == 2 ? h[1].z : h[1].w, t);
        t = fmaf(ww[2], a == 0 ? h[2].x : a == 1 ? h[2].y : a == 2 ? h[2].z : h[2].w, t);
        t = fmaf(ww[3], a == 0 ? h[3].x : a == 1 ? h[3].y : a == 2 ? h[3].z : h[3].w, t);
        acc[a][b] = t;
      }
    }
  }
}

template <bool RELU>
__device__ __forceinline__ void store_tile(float* s_out, int tx, int fq, const float (&acc)[4][4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    float4 v = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
    if (RELU) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    *reinterpret_cast<float4*>(s_out + tile_off(4 * fq + b, tx)) = v;
  }
}

// rows [0, nrows) of a tile -> dst[f][p0 + px] (row pitch B), coalesced over pixels
__device__ __forceinline__ void tile_to_global(const float* s_t, float* __restrict__ dst, int nrows, int p0, int B, int tid) {
  for (int idx = tid; idx < nrows * GP; idx += BT) {
    const int f = idx >> 6, px = idx & 63;
    if (p0 + px < B) dst[(size_t)f * B + p0 + px] = s_t[tile_at(f, px)];
  }
}
__device__ __forceinline__ void tile_from_global(float* s_t, const float* __restrict__ src, int nrows, int p0, int B, int tid) {
  for (int idx = tid; idx < nrows * GP; idx += BT) {
    const int f = idx >> 6, px = idx & 63;
    s_t[tile_at(f, px)] = (p0 + px < B) ? src[(size_t)f * B + p0 + px] : 0.f;
  }
}

constexpr int BLUR_FWD_SMEM_FLOATS = 2 * BH * BH + BH * 28 + BH * 4 + 160 + 3 * GP + 2 * BH * GP;

__global__ void __launch_bounds__(BT, 3)
blur_fwd_kernel(const float* __restrict__ coords, const long long* __restrict__ coord_idx,
                const float* __restrict__ focus, BlurParams bp, int B, int HW, float sy, float sx,
                float offset_size, float* __restrict__ uv, float* __restrict__ kout,
                float* __restrict__ hid) {
  extern __shared__ __align__(16) float sm[];
  float* s_w2 = sm;                   // [64 out][64 in] as stored
  float* s_w3 = s_w2 + BH * BH;       // [64][64]
  float* s_w4 = s_w3 + BH * BH;       // [28][64]  (row 27 zero)
  float* s_w1 = s_w4 + BH * 28;       // [64][4]   (w1[o][0..2], b1[o])
  float* s_b = s_w1 + BH * 4;         // b2[64], b3[64], b4[28] (+4)
  float* s_x = s_b + 160;             // [3][64]: focus, y, x of the group's pixels
  float* s_t0 = s_x + 3 * GP;         // activation tiles (ping-pong)
  float* s_t1 = s_t0 + BH * GP;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.x * GP;
  for (int i = tid; i < BH * BH; i += BT) {
    s_w2[i] = __ldg(bp.w2 + i);
    s_w3[i] = __ldg(bp.w3 + i);
  }
  for (int i = tid; i < BH * 28; i += BT) s_w4[i] = (i < BO * BH) ? __ldg(bp.w4 + i) : 0.f;
  if (tid < BH) {
    const int o = tid;
    *reinterpret_cast<float4*>(s_w1 + o * 4) = make_float4(bp.w1[o * 3], bp.w1[o * 3 + 1], bp.w1[o * 3 + 2], bp.b1[o]);
    s_b[o] = bp.b2[o];
    s_b[64 + o] = bp.b3[o];
    if (o < 32) s_b[128 + o] = (o < BO) ? bp.b4[o] : 0.f;
  } else if (tid < BH + GP) {
    const int px = tid - BH;
    const int p = p0 + px, pc = p < B ? p : B - 1;
    s_x[px] = focus[(int)(coord_idx[pc] / HW)];
    s_x[GP + px] = coords[3 * pc + 1];
    s_x[2 * GP + px] = coords[3 * pc + 2];
  }
  __syncthreads();

  float acc[4][4];
  {
    // layer 1 (3 -> 64) + ReLU
    const float4 f4 = *reinterpret_cast<const float4*>(s_x + 4 * tx), y4 = *reinterpret_cast<const float4*>(s_x + GP + 4 * tx),
                 x4 = *reinterpret_cast<const float4*>(s_x + 2 * GP + 4 * tx);
    const float ff[4] = {f4.x, f4.y, f4.z, f4.w}, yy[4] = {y4.x, y4.y, y4.z, y4.w}, xx[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float4 w = *reinterpret_cast<const float4*>(s_w1 + (4 * ty + b) * 4);
#pragma unroll
      for (int a = 0; a < 4; ++a) acc[a][b] = fmaf(w.x, ff[a], fmaf(w.y, yy[a], fmaf(w.z, xx[a], w.w)));
    }
    store_tile<true>(s_t0, tx, ty, acc);
  }
  __syncthreads();
  tile_to_global(s_t0, hid, BH, p0, B, tid);
  fwd_tile(s_t0, s_w2, s_b, tx, ty, acc);
  store_tile<true>(s_t1, tx, ty, acc);
  __syncthreads();
  tile_to_global(s_t1, hid + (size_t)BH * B, BH, p0, B, tid);
  fwd_tile(s_t1, s_w3, s_b + 64, tx, ty, acc);
  store_tile<true>(s_t0, tx, ty, acc);
  __syncthreads();
  tile_to_global(s_t0, hid + (size_t)2 * BH * B, BH, p0, B, tid);
  if (ty < 7) {
    fwd_tile(s_t0, s_w4, s_b + 128, tx, ty, acc);
    store_tile<false>(s_t1, tx, ty, acc);
  }
  __syncthreads();

  // head: softmax over the K kernel weights, tanh over the 2K offsets (modules.py:303-306); warps 0-1 take the softmax of
  // their pixel, warps 2-7 six offsets each
  const int px = tid & 63, part = tid >> 6;
  const int p = p0 + px;
  const bool valid = p < B;
  if (part == 0) {
    float out[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) out[k] = s_t1[tile_at(k, px)];
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
    for (int k = 0; k < KT; ++k) {
      out[k] *= rden;
      if (valid) kout[(size_t)k * B + p] = out[k];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int o = KT + (part - 1) * 6 + j;
      const float v = tanhf(s_t1[tile_at(o, px)]);
      s_t1[tile_at(o, px)] = v;
      if (valid) kout[(size_t)o * B + p] = v;
    }
  }
  __syncthreads();
  if (!valid) return;
  // tap coordinates: neighbour grid + learned offsets, centre tap = the pixel itself
  const float cy = s_x[GP + px], cx = s_x[2 * GP + px];
  for (int k = part; k < KT; k += 4) {
    const int dy = k / 3 - 1, dx = k % 3 - 1;
    float y = cy + dy * sy, x = cx + dx * sx;
    y += s_t1[tile_at(KT + k, px)] * sy * offset_size;
    x += s_t1[tile_at(2 * KT + k, px)] * sx * offset_size;
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
  // per channel and hidden unit: (u, a, v, v * u) as ONE 16-byte shared-memory word: the CRF and its slope are evaluated
  // in a single pass over the 128 units (one broadcast LDS.128, FFMA, FMNMX, FFMA, FSETP, predicated FADD per unit)
  __shared__ float4 s_tm[3][TMW];
  __shared__ float s_x[3][PB + 1], s_g[3][PB + 1];
  __shared__ float s_de[PB];
  __shared__ int s_fr[PB];
  __shared__ float s_red[PB / 32][8];
  const int tid = threadIdx.x;
  for (int c = 0; c < 3; ++c) {
    const float u = tm.u[c][tid], v = tm.v[c][tid];
    s_tm[c][tid] = make_float4(u, tm.a[c][tid], v, v * u);
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
    float o = tm.d[c][0], acc = 0.f;   // CRF output and d(CRF)/d(lnx) = sum over the open units of v * u
#pragma unroll 8
    for (int i = 0; i < TMW; ++i) {
      const float4 t = s_tm[c][i];
      const float pre = fmaf(t.x, lnx[c], t.y);
      o = fmaf(t.z, fmaxf(pre, 0.f), o);
      acc += (pre > 0.f) ? t.w : 0.f;
    }
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
      const float ui = s_tm[c][i].x, ai = s_tm[c][i].y, vi = s_tm[c][i].z;
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

// gw[o][i] (+)= sum_px dz[o][px] * h[i][px], gb[o] (+)= sum_px dz[o][px] over the group's 64 pixels, pixels ascending.
// Thread (oq, iq) owns the 4 x 4 block o in [4 oq, +4), i in [4 iq, +4); NOQ output quads are live (7 for the head).
template <int NOQ>
__device__ __forceinline__ void wgrad_tile(const float* s_dz, const float* s_h, float* gw, float* gb, int nout, bool first,
                                           int tid) {
  const int oq = tid >> 4, iq = tid & 15;
  if (oq >= NOQ) return;
  float acc[4][4], bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 2
  for (int g = 0; g < GP / 4; ++g) {
    float4 d[4], h[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) d[a] = *reinterpret_cast<const float4*>(s_dz + tile_off(4 * oq + a, g));
#pragma unroll
    for (int b = 0; b < 4; ++b) h[b] = *reinterpret_cast<const float4*>(s_h + tile_off(4 * iq + b, g));
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float t = acc[a][b];
        t = fmaf(d[a].x, h[b].x, t);
        t = fmaf(d[a].y, h[b].y, t);
        t = fmaf(d[a].z, h[b].z, t);
        t = fmaf(d[a].w, h[b].w, t);
        acc[a][b] = t;
      }
      if (iq == 0) bs[a] = (((bs[a] + d[a].x) + d[a].y) + d[a].z) + d[a].w;
    }
  }
  // this block's partial record: every entry is owned by one thread, for every group
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = 4 * oq + a;
    if (o >= nout) continue;
    float4* g4 = reinterpret_cast<float4*>(gw + o * BH + 4 * iq);
    float4 v = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    if (!first) {
      const float4 old = *g4;
      v = make_float4(old.x + v.x, old.y + v.y, old.z + v.z, old.w + v.w);
    }
    *g4 = v;
    if (iq == 0) gb[o] = first ? bs[a] : gb[o] + bs[a];
  }
}

// dh[px][i] = sum_{o < NO} dz[o][px] * w[o][i] (w = [out][in] as stored, o ascending) for px in [4 tx, +4), i in [4 ti, +4)
template <int NO>
__device__ __forceinline__ void dgrad_tile(const float* s_dz, const float* s_w, int tx, int ti, float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
  for (int o = 0; o < NO; ++o) {
    const float4 d = *reinterpret_cast<const float4*>(s_dz + tile_off(o, tx));
    const float4 w = *reinterpret_cast<const float4*>(s_w + o * BH + 4 * ti);
    const float dd[4] = {d.x, d.y, d.z, d.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ww[b], dd[a], acc[a][b]);
  }
}

// ReLU backward of a dgrad tile against the layer's input activation (still in s_h): keeps the gradient where h > 0
__device__ __forceinline__ void relu_mask(const float* s_h, int tx, int ti, float (&acc)[4][4]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const float4 h = *reinterpret_cast<const float4*>(s_h + tile_off(4 * ti + b, tx));
    acc[0][b] = h.x > 0.f ? acc[0][b] : 0.f;
    acc[1][b] = h.y > 0.f ? acc[1][b] : 0.f;
    acc[2][b] = h.z > 0.f ? acc[2][b] : 0.f;
    acc[3][b] = h.w > 0.f ? acc[3][b] : 0.f;
  }
}

constexpr int BLUR_BWD_SMEM_FLOATS = 2 * BH * BH + 28 * BH + 2 * BH * GP + 3 * GP + 2 * GP;

// (three blocks of 8 warps per SM: the kernel is bound by the latency of its dependent phases, not by a pipe)
__global__ void __launch_bounds__(BT, 3)
blur_bwd_kernel(const float* __restrict__ coords, const long long* __restrict__ coord_idx,
                const float* __restrict__ focus, BlurParams bp, float* __restrict__ part, int rec_floats, int B, int HW,
                int N, float sy, float sx, float offset_size, const float* __restrict__ kout,
                const float* __restrict__ hid, const float* __restrict__ dkw, const float* __restrict__ duv) {
  extern __shared__ __align__(16) float sm[];
  float* s_w2 = sm;                    // [64 out][64 in] as stored
  float* s_w3 = s_w2 + BH * BH;        // [64][64]
  float* s_w4 = s_w3 + BH * BH;        // [28][64] (row 27 zero)
  float* s_dz = s_w4 + 28 * BH;        // tile: current layer's d(pre-activation) [out][pixel]
  float* s_h = s_dz + BH * GP;         // tile: that layer's input activation [in][pixel]
  float* s_x = s_h + BH * GP;          // [3][64]: focus, y, x of the group's pixels
  float* s_df = s_x + 3 * GP;          // [64] d loss / d focus per pixel
  int* s_fr = reinterpret_cast<int*>(s_df + GP);   // [64] frame of the pixel (-1 past the end)
  const int tid = threadIdx.x, tx = tid & 15, ti = tid >> 4;
  for (int i = tid; i < BH * BH; i += BT) {
    s_w2[i] = bp.w2[i];
    s_w3[i] = bp.w3[i];
  }
  for (int i = tid; i < 28 * BH; i += BT) s_w4[i] = (i < BO * BH) ? bp.w4[i] : 0.f;
  // Persistent: at most three blocks per SM (72 KB of shared memory each) walk the 64-pixel groups and each adds its
  // groups into ONE partial record, every entry owned by one thread, groups in order: the second pass sums gridDim.x
  // records.
  const int ngroups = (B + GP - 1) / GP;
  float* rec = part + (size_t)blockIdx.x * rec_floats;
#pragma unroll 1
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const bool first = grp == (int)blockIdx.x;
    const int p0 = grp * GP;
    // ---- d(logits) of the head: softmax backward (warps 0-1, one pixel per thread), tanh backward of the offsets ----
    {
      const int px = tid & 63, prt = tid >> 6;
      const int p = p0 + px;
      const bool valid = p < B;
      if (prt == 0) {
        float w[KT], d[KT], gsum = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          w[k] = valid ? kout[(size_t)k * B + p] : 0.f;
          d[k] = valid ? dkw[(size_t)k * B + p] : 0.f;
          gsum = fmaf(d[k], w[k], gsum);
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) s_dz[tile_at(k, px)] = w[k] * (d[k] - gsum);
        s_dz[tile_at(BO, px)] = 0.f;
        const int pc = valid ? p : (B - 1);
        const int fr = (int)(coord_idx[pc] / HW);
        s_x[px] = focus[fr];
        s_x[GP + px] = coords[3 * pc + 1];
        s_x[2 * GP + px] = coords[3 * pc + 2];
        s_fr[px] = valid ? fr : -1;
      } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const int o = KT + (prt - 1) * 6 + j;
          const int k = (o - KT) % KT, isx = o >= 2 * KT;
          float v = 0.f;
          if (valid && k != KT / 2) {   // the centre tap is overwritten with the raw coords (training.py:120)
            const float d = duv[2 * ((size_t)k * B + p) + isx];
            const float t = kout[(size_t)o * B + p];
            v = d * (isx ? sx : sy) * offset_size * (1.f - t * t);
          }
          s_dz[tile_at(o, px)] = v;
        }
      }
    }
    tile_from_global(s_h, hid + (size_t)2 * BH * B, BH, p0, B, tid);
    __syncthreads();

    float acc[4][4];
    // ---- layer 4 (64 -> 27): dW4, db4, dh3 ----
    wgrad_tile<7>(s_dz, s_h, rec + BR_W4, rec + BR_B4, BO, first, tid);
    dgrad_tile<BO>(s_dz, s_w4, tx, ti, acc);
    relu_mask(s_h, tx, ti, acc);
    __syncthreads();
    store_tile<false>(s_dz, tx, ti, acc);
    tile_from_global(s_h, hid + (size_t)BH * B, BH, p0, B, tid);
    __syncthreads();
    // ---- layers 3 and 2 (64 -> 64) ----
    wgrad_tile<16>(s_dz, s_h, rec + BR_W3, rec + BR_B3, BH, first, tid);
    dgrad_tile<BH>(s_dz, s_w3, tx, ti, acc);
    relu_mask(s_h, tx, ti, acc);
    __syncthreads();
    store_tile<false>(s_dz, tx, ti, acc);
    tile_from_global(s_h, hid, BH, p0, B, tid);
    __syncthreads();
    wgrad_tile<16>(s_dz, s_h, rec + BR_W2, rec + BR_B2, BH, first, tid);
    dgrad_tile<BH>(s_dz, s_w2, tx, ti, acc);
    relu_mask(s_h, tx, ti, acc);
    __syncthreads();
    store_tile<false>(s_dz, tx, ti, acc);
    __syncthreads();
    // ---- layer 1 (3 -> 64): dW1, db1, d focus ----
    {
      const int o = tid >> 2, c = tid & 3;
      float t = 0.f;
#pragma unroll 4
      for (int g = 0; g < GP / 4; ++g) {
        const float4 d = *reinterpret_cast<const float4*>(s_dz + tile_off(o, g));
        if (c < 3) {
          const float4 xv = *reinterpret_cast<const float4*>(s_x + c * GP + 4 * g);
          t = fmaf(d.x, xv.x, t);
          t = fmaf(d.y, xv.y, t);
          t = fmaf(d.z, xv.z, t);
          t = fmaf(d.w, xv.w, t);
        } else {
          t = (((t + d.x) + d.y) + d.z) + d.w;
        }
      }
      float* dst = (c < 3) ? (rec + BR_W1 + o * 3 + c) : (rec + BR_B1 + o);
      *dst = first ? t : *dst + t;
    }
    if (tid < GP) {
      float t = 0.f;
#pragma unroll 8
      for (int o = 0; o < BH; ++o) t = fmaf(__ldg(bp.w1 + o * 3), s_dz[tile_at(o, tid)], t);
      s_df[tid] = t;
    }
    __syncthreads();
    // focus gradient: per-frame gather over the group's pixels in order (training.py:104-108 backward; no atomics)
    for (int f = tid; f < N; f += BT) {
      float t = 0.f;
      for (int q = 0; q < GP; ++q)
        if (s_fr[q] == f) t += s_df[q];
      rec[BR_FOCUS + f] = first ? t : rec[BR_FOCUS + f] + t;
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
  const size_t smem = sizeof(float) * BLUR_FWD_SMEM_FLOATS;
  int rc = nchost::ensure_dyn_smem((const void*)blur_fwd_kernel, (int)smem);
  if (rc) return rc;
  const int grid = (B + GP - 1) / GP;
  blur_fwd_kernel<<<grid, BT, smem, (cudaStream_t)stream>>>(
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
  const size_t smem = sizeof(float) * BLUR_BWD_SMEM_FLOATS;
  int rc = nchost::ensure_dyn_smem((const void*)blur_bwd_kernel, (int)smem);
  if (rc) return rc;
  const int grid = (B + GP - 1) / GP;      // pixel groups; one partial record per launched block
  const int rec = nchost::ws_rec_blur_bwd(N);
  const int launch_blocks = grid < 3 * nchost::sm_count() ? grid : 3 * nchost::sm_count();
  blur_bwd_kernel<<<launch_blocks, BT, smem, (cudaStream_t)stream>>>(
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
