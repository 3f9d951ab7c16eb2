// sine_chain.cu -- the scene sine-MLP (SingleBVPNet 2 -> 512 -> 512 -> 512 -> 512 -> 3, modules.py:430-481)
// forward and input/activation backward as ONE persistent, warp-specialised kernel per direction.
//
// A CTA owns a 128-sample row tile.  The tile's activations live in shared memory as the K-major
// SWIZZLE_128B A operand (8 k-blocks of 128 rows x 64 fp16 = 128 KB) for the whole chain; the three
// 512x512 hidden layers run on tcgen05 tensor cores (fp16 operands, fp32 accumulators = all 512 TMEM
// columns).  Two CTAs form a pair (cluster of 2, tcgen05 cta_group::2): the leader's MMA thread issues
// M=256 x N=256 x K=16 instructions that multiply BOTH CTAs' operand tiles by one 256-row weight tile of
// which each CTA stages only half (128 rows, 16 KB per k-block) through its 4-stage TMA ring.  That halves
// the shared-memory operand bandwidth and the L2->SMEM weight traffic per SM, which is what bounds a
// single-CTA version (measured: 128x128x16 MMAs at 41 % of tensor peak, see profiles/).
//
//   forward  (FCBlock.forward / BatchLinear.forward / Sine.forward, modules.py:17-35,90-95):
//       a1 = sin(30 (uv W0^T + b0))            SIMT prologue (K = 2)
//       a_{l+1} = sin(30 (a_l W_l^T + b_l))    l = 1..3, tcgen05 + epilogue
//       y  = a4 W4^T + b4                      SIMT, fused in the last epilogue (N = 3)
//     side outputs for the backward: a1..a3 (TMA-stored straight from the operand tile) and the PHASE
//     of 30 z_l (l = 1..3) as 16-bit fixed point in revolutions, from which the backward regenerates
//     cos(30 z_l) with one MUFU (tile-blocked layout so the per-row threads store/load coalesced).  a4 is
//     not stored: the head weight gradient regenerates it as sin of the stored phase of z3 (wgrad.cu).
//   backward (autograd of the same, training.py:270):
//       dz3 = (dy W4) * 30 cos(30 z3)          tcgen05 (one K = 16 step on a zero-padded fp16 dy / W4^T pair) + epilogue
//       dz_{l-1} = (dz_l 30 W_l) * cos(30 z_{l-1})   l = 3..1, tcgen05 + epilogue (30 folded into the fp16 W^T)
//       duv = dz0 W0                           SIMT, fused in the last epilogue
//       dW0 = dz0^T uv, db0 = sum dz0          column pass over the dz0 tile while it is still in smem
//     dz1..dz3 are TMA-stored for the weight-gradient kernel (wgrad.cu).  All dz carry a power-of-two
//     loss scale read from device memory so that they sit in fp16's normal range.
//
// Overlap of tensor cores and epilogue inside one tile: a layer's MMAs are issued as two 256-column
// halves with their own TMEM-full barriers, so the epilogue of half 0 runs while half 1 is still on
// the tensor cores; the epilogue itself works in four 128-column chunks.  The next layer's operand overwrites the current one in place, k-block by k-block:
// k-block kb may only be overwritten once the LAST half's MMAs have consumed it (kfree[kb/2], a
// tcgen05.commit after k-blocks 1 and 3) and once its TMA store has drained (stored[kb]).  Until then the packed fp16
// results of chunks 0..2 are parked in the (already drained) accumulator columns with tcgen05.st.  The
// next layer's first chunk starts as soon as its first k-blocks land (a_ready[kb]).
//
// Warp roles (640 threads per CTA): warps 0..15 = prologue/epilogue, warp 16 = TMA weight producer, warp 17 =
// MMA issuer (leader CTA) / k-block-ready relay (peer CTA), warp 18 = TMEM allocator, warp 19 = TMA store
// issuer.  Epilogue: thread = row; warp e
// handles the 32 columns [128 c + 32 (e/4), +32) of every chunk c, i.e. half a row of one operand k-block).
// 16 epilogue warps = 4 per scheduler: the epilogue is latency-bound, not pipe-bound (profiles/).
#include "host_util.h"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int HID = 512;
constexpr int TILE_M = 128;
constexpr uint32_t STAGE_BYTES = 128 * 64 * 2;     // weight chunk: 128 output rows x 64 k
constexpr uint32_t A_BLOCK_BYTES = TILE_M * 128;   // one 64-column k-block of the operand tile
constexpr uint32_t A_BYTES = 8 * A_BLOCK_BYTES;    // 128 x 512 fp16
constexpr int NTHREADS = 640;
constexpr int FIRST_EPI_WARP = 0;
// Control warps take the HIGHEST warp ids: the warp scheduler arbitrates highest-id-first, and the single-lane
// producer / MMA-issuer / relay / store threads must never queue behind the 16 busy epilogue warps (measured:
// with control warps at ids 0..3 every mbarrier poll of the MMA thread took 100-900 ns).
constexpr int WARP_PRODUCER = 16, WARP_MMA = 17, WARP_ALLOC = 18, WARP_STORE = 19;
constexpr int PAIR = 2;                            // CTAs per cluster (tcgen05 cta_group::2)
constexpr uint16_t PAIR_MASK = 0x3;

// Shared-memory layout.  The weight ring is what bounds the tensor pipe inside a layer: with a stage round trip
// (commit -> empty -> producer -> TMA -> full -> issuer) of ~1.3 us and 260 ns of MMAs per stage, bytes in flight =
// ring bytes decide the MMA rate (in-kernel timeline, profiles/r2_findings.md).  The backward needs neither the head
// weights nor the hidden biases in shared memory, so it affords a fifth stage.
template <bool BWD>
struct Lay {
  static constexpr int NSTAGE = BWD ? 5 : 4;
  static constexpr uint32_t OFF_A = 0;
  static constexpr uint32_t OFF_W = OFF_A + A_BYTES;
  static constexpr uint32_t OFF_P0 = OFF_W + NSTAGE * STAGE_BYTES;    // float4[512]: 30*W0[j,0], 30*W0[j,1], 30*b0[j], 0
  static constexpr uint32_t OFF_W4 = OFF_P0 + 512 * 16;               // float4[512]: W4[0..3][j] (missing rows = 0)  (forward)
  static constexpr uint32_t OFF_B30 = OFF_W4 + (BWD ? 0 : 512 * 16);  // float[3][512]: 30*b_l                       (forward)
  static constexpr uint32_t OFF_XCH = OFF_B30 + (BWD ? 0 : 3 * 512 * 4);   // float4[4][128]: per-column-quarter partial sums
  static constexpr uint32_t OFF_UV = OFF_XCH + 4 * 128 * 16;          // float2[128]: the tile's input coordinates
  static constexpr uint32_t OFF_BAR = OFF_UV + 128 * 8;               // 62 mbarriers
  static constexpr uint32_t OFF_TMEM = OFF_BAR + 512;
  static constexpr uint32_t SMEM_USED = OFF_TMEM + 16;
  static constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;            // + alignment slack
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
};

// mbarrier slots (8 reserved per ring barrier kind)
constexpr uint32_t BAR_WFULL = 0, BAR_WEMPTY = 8, BAR_ACC = 16, BAR_KFREE = 20, BAR_AREADY = 28, BAR_STORED = 36,
                   BAR_PEER = 44,   // leader only: the peer CTA's operand k-block kb is ready
                   BAR_AOUT = 52,   // k-block kb of the tile's LAST output (a4 / dz0) is written (store warp only)
                   BAR_DYR = 60,    // backward: the tile's scaled output-gradient operand (dy, K = 16) is written
                   BAR_PEERDY = 61; // leader only: same, for the peer CTA
// a_ready / peer count the MMA-visible operand events only (prologue + first two layers = 3 per tile): every
// such event needs MMA progress on the previous one, so no waiter can be lapped by a full barrier phase.

constexpr float TWO_PI = 6.283185307179586f;
constexpr float PH_PER_RAD = 65536.0f / TWO_PI;   // phase counts per radian
constexpr float MAGIC = 12582912.0f;              // 1.5 * 2^23: float -> int rounding on the FMA pipe

struct ChainParams {
  int Q;
  int ntiles;
  int out_dim;           // 3 (or 4 for a foreground atlas without alpha MLP)
  int store_act;         // forward: write a1..a4 / phases for the backward
  const float* uv;       // [Q,2]
  const float* w0;       // [512,2]
  const float* b0;       // [512]
  const float* b_hid;    // [3,512]
  const float* w4;       // [out_dim,512]
  const float* b4;       // [out_dim]
  uint16_t* phase;       // [3][ntiles*128*512] tile-blocked 16-bit phases of 30 z_l, l = 1..3
  float* y;              // forward out [Q,out_dim]
  const float* dy;       // backward in [Q,out_dim]
  float* duv;            // backward out [Q,2]
  float* part;           // backward out: per-CTA partial record [gridDim.x][dW0 512x2 | db0 512] (reduce.cu sums them)
  const float* scale;    // backward: device scalar, power-of-two loss scale
#ifdef NEUCAM_DEBUG
  long long* trace;      // optional debug timeline of CTA 0 (clock64 stamps), or null
  int dbg;               // debug experiments: bit0 = do not issue MMAs, bit1 = do not issue weight loads
#endif
};

// The in-kernel timeline and the "no MMA / no TMA / no epilogue math" experiments exist only in the debug build
// (-DNEUCAM_DEBUG -> libneucam_b200_debug.so); the product library compiles them out.
#ifdef NEUCAM_DEBUG
// trace layout: [role 0=mma,1=epi half0,2=epi half1 | +3 for the peer CTA][tile 0..1][g 0..2][step 0..7][4 stamps]
// (globaltimer so that the two SMs of the pair share a time base)
__device__ __forceinline__ void trace_stamp(long long* tr, int role, int tile_i, int g, int step, int k) {
  if (tr != nullptr && blockIdx.x < 2 && tile_i < 2 && g >= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tr[((((role + 3 * (int)blockIdx.x) * 2 + tile_i) * 3 + g) * 8 + step) * 4 + k] = (long long)t;
  }
}
#define CHAIN_DBG(bit) ((p.dbg & (bit)) != 0)
#else
#define trace_stamp(...) ((void)0)
#define CHAIN_DBG(bit) false
#endif

struct ChainMaps {
  CUtensorMap w;         // stacked hidden weights [3*512, 512] fp16 (forward: W_l; backward: W_l^T), box 64 x 128
  CUtensorMap act[4];    // forward: a1..a3, (unused) ; backward: dz3, dz2, dz1, (unused)   ([Q,512] fp16, box 64x128)
  CUtensorMap w4t;       // backward: head weights transposed + zero-padded, fp16 [512, 64], box 64 x 128
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// saturating pack (finite max instead of inf) for the scaled gradients
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
// cos of a 16-bit phase (revolutions * 65536): exact int->float through the mantissa, one FFMA, one MUFU
__device__ __forceinline__ float cos_of_phase(uint32_t p16) {
  const float f = __uint_as_float(0x4B000000u | p16);   // 2^23 + p
  return __cosf(fmaf(f, TWO_PI / 65536.0f, -8388608.0f * (TWO_PI / 65536.0f)));
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// one 128-byte row (64 fp16 = 32 packed words) of k-block `kb` of the operand tile
__device__ __forceinline__ void write_a_row(uint32_t a_base, uint32_t kb, uint32_t row, const uint32_t (&w)[32]) {
  const uint32_t blk = a_base + kb * A_BLOCK_BYTES;
#pragma unroll
  for (uint32_t q = 0; q < 8; ++q)
    st_shared_v4(blk + sw128_offset(row, q), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// half of a 128-byte row (32 fp16 = 16 packed words): 16-byte chunks qb..qb+3 of k-block `kb`
__device__ __forceinline__ void write_a_half_row(uint32_t a_base, uint32_t kb, uint32_t row, uint32_t qb,
                                                 const uint32_t (&w)[16]) {
  const uint32_t blk = a_base + kb * A_BLOCK_BYTES;
#pragma unroll
  for (uint32_t q = 0; q < 4; ++q)
    st_shared_v4(blk + sw128_offset(row, qb + q), w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
}

// ---- per-chunk epilogue math ----------------------------------------------------------------------------------
// One thread turns its 32 accumulator columns of one row into 16 packed fp16 pairs of the next operand, 16 columns
// (one tcgen05.ld) at a time so that the second load is in flight under the first half's math.  The layer-
// dependent variants are TEMPLATES selected by a warp-uniform branch around the whole chunk: written as conditions
// inside one unrolled loop, the compiler if-converts them and every element pays for both sides (measured in SASS:
// 2 MUFU + 5 FFMA + 4 FMUL per element in the backward instead of 1 + 1 + 2).

// forward: a = sin(30 z + 30 b); STORE: also the 16-bit phase of the angle (for the backward's cos); HEAD: accumulate
// this thread's share of the 512 -> out_dim output layer.
template <bool STORE, bool HEAD>
__device__ __forceinline__ void fwd_half_math(const uint32_t (&r)[16], uint32_t* __restrict__ w, const float* __restrict__ b30c,
                                              const float4* __restrict__ w4c, uint16_t* __restrict__ ph_dst, bool ph_ok,
                                              float (&acc)[4]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t pw[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int i = q * 8 + jj * 2;
      const float t0 = fmaf(__uint_as_float(r[i]), 30.f, b30c[i]), t1 = fmaf(__uint_as_float(r[i + 1]), 30.f, b30c[i + 1]);
      const float h0 = __sinf(t0), h1 = __sinf(t1);
      w[q * 4 + jj] = pack_h2(h0, h1);
      if (STORE) {
        // phase in revolutions * 65536, rounded on the FMA pipe (1.5 * 2^23 magic number): one FFMA from the angle
        // the sine already needed
        const uint32_t p0 = __float_as_uint(fmaf(t0, PH_PER_RAD, MAGIC));
        const uint32_t p1 = __float_as_uint(fmaf(t1, PH_PER_RAD, MAGIC));
        pw[jj] = __byte_perm(p0, p1, 0x5410);
      }
      if (HEAD) {
        const float4 wa = w4c[i], wb = w4c[i + 1];
        acc[0] = fmaf(h0, wa.x, fmaf(h1, wb.x, acc[0]));
        acc[1] = fmaf(h0, wa.y, fmaf(h1, wb.y, acc[1]));
        acc[2] = fmaf(h0, wa.z, fmaf(h1, wb.z, acc[2]));
        acc[3] = fmaf(h0, wa.w, fmaf(h1, wb.w, acc[3]));
      }
    }
    if (STORE && ph_ok)
      *reinterpret_cast<uint4*>(ph_dst + (size_t)q * (TILE_M * 8)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
  }
}

// backward: dz = (dh W^T-accumulator) * cos(angle) (the operand W^T carries the factor 30).  !LAST: the angle comes from
// the stored 16-bit phase; LAST (layer 0): it is recomputed from the input coordinates, and the thread accumulates its
// share of the coordinate gradient duv = dz0 W0.
template <bool LAST>
__device__ __forceinline__ void bwd_half_math(const uint32_t (&r)[16], uint32_t* __restrict__ w, const uint4* __restrict__ pv,
                                              const float4* __restrict__ p0c, float u, float v, float (&acc)[4]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t pw[4] = {pv[q].x, pv[q].y, pv[q].z, pv[q].w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int i = q * 8 + jj * 2;
      float z0, z1;
      if (!LAST) {
        z0 = __uint_as_float(r[i]) * cos_of_phase(pw[jj] & 0xFFFFu);
        z1 = __uint_as_float(r[i + 1]) * cos_of_phase(pw[jj] >> 16);
      } else {
        const float4 q0 = p0c[i], q1 = p0c[i + 1];
        z0 = __uint_as_float(r[i]) * __cosf(fmaf(u, q0.x, fmaf(v, q0.y, q0.z)));
        z1 = __uint_as_float(r[i + 1]) * __cosf(fmaf(u, q1.x, fmaf(v, q1.y, q1.z)));
        acc[0] = fmaf(z0, q0.x, fmaf(z1, q1.x, acc[0]));
        acc[1] = fmaf(z0, q0.y, fmaf(z1, q1.y, acc[1]));
      }
      w[q * 4 + jj] = pack_h2_sat(z0, z1);
    }
  }
}

// BWD: backward chain; STORE (forward only): write a1..a4 and the phases for a later backward (training) or not (inference)
template <bool BWD, bool STORE>
__global__ void __cluster_dims__(PAIR, 1, 1) __launch_bounds__(NTHREADS, 1)
sine_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using L = Lay<BWD>;
  constexpr int NSTAGE = L::NSTAGE;
  constexpr uint32_t OFF_A = L::OFF_A, OFF_W = L::OFF_W, OFF_P0 = L::OFF_P0, OFF_W4 = L::OFF_W4, OFF_B30 = L::OFF_B30,
                     OFF_XCH = L::OFF_XCH, OFF_UV = L::OFF_UV, OFF_BAR = L::OFF_BAR, OFF_TMEM = L::OFF_TMEM;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);

  float4* s_p0 = reinterpret_cast<float4*>(sm + OFF_P0);
  float4* s_w4 = reinterpret_cast<float4*>(sm + OFF_W4);
  float* s_b30 = reinterpret_cast<float*>(sm + OFF_B30);
  float4* s_xch = reinterpret_cast<float4*>(sm + OFF_XCH);
  float2* s_uv = reinterpret_cast<float2*>(sm + OFF_UV);
  const uint32_t bars = base + OFF_BAR;
  auto bar = [bars](uint32_t slot) { return bars + 8u * slot; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup -------------------------------------------------------------------------
  for (int j = tid; j < HID; j += NTHREADS) {
    s_p0[j] = make_float4(30.f * p.w0[2 * j], 30.f * p.w0[2 * j + 1], 30.f * p.b0[j], 0.f);
    if (!BWD) {
      float w4v[4];
      for (int c = 0; c < 4; ++c) w4v[c] = (c < p.out_dim) ? p.w4[c * HID + j] : 0.f;
      s_w4[j] = make_float4(w4v[0], w4v[1], w4v[2], w4v[3]);
      for (int l = 0; l < 3; ++l) {
        const float b30 = 30.f * p.b_hid[l * HID + j];
        s_b30[l * HID + j] = b30;
      }
    }
  }
  if (tid == 0) {
    for (uint32_t s = 0; s < NSTAGE; ++s) {
      mbar_init(bar(BAR_WFULL + s), 1);
      mbar_init(bar(BAR_WEMPTY + s), 1);
    }
    for (uint32_t c = 0; c < 4; ++c) mbar_init(bar(BAR_ACC + c), 1);
    for (uint32_t k = 0; k < 8; ++k) {
      mbar_init(bar(BAR_KFREE + k), 1);
      mbar_init(bar(BAR_AREADY + k), 256);
      mbar_init(bar(BAR_STORED + k), 1);
      mbar_init(bar(BAR_PEER + k), 1);
      mbar_init(bar(BAR_AOUT + k), 256);
    }
    mbar_init(bar(BAR_DYR), 128);
    mbar_init(bar(BAR_PEERDY), 1);
    mbar_fence_init();
    tma_prefetch_desc(&maps.w);
  }
  if (warp == WARP_ALLOC) {
    tmem_alloc_2cta(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / commit / TMA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const bool do_store_any = BWD || STORE;
  const uint32_t crank = cluster_ctarank();
  const bool leader = (crank == 0);
  // a pair walks tiles (2j, 2j+1) + it * gridDim.x together; it runs while its leader's tile exists.  When
  // only the peer's tile is past the end it runs a phantom tile (all rows invalid, no global traffic).
  const int lead0 = (int)(blockIdx.x & ~1u);

  if (warp == WARP_PRODUCER) {
    // ===================== TMA weight producer =====================
    // (the whole warp runs the loop and the waits; only the issue instructions are predicated to lane 0, so the
    //  warp never sits diverged against lanes parked at the final barrier)
    {
      uint32_t s = 0, ph = 0;
      // the leader's w_full barriers as shared::cluster addresses (mapa is linear inside a CTA's window: stage s is
      // at + 8 s; an indexed local array here would put a local-memory load on the producer's critical path)
      const uint32_t full_bar0 = mapa_shared(bar(BAR_WFULL), 0);
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x) {
        if (BWD) {
          // head dgrad dh4 = dy W4 runs on the tensor cores too: one K = 16 step per 256-column half
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar(BAR_WEMPTY + s), ph ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(bar(BAR_WFULL + s), 2 * STAGE_BYTES);
              tma_load_2d_2cta(base + OFF_W + s * STAGE_BYTES, &maps.w4t, full_bar0 + 8u * s, 0, h * 256 + (int)crank * 128);
            }
            __syncwarp();
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
        for (int g = 0; g < 3; ++g) {
          const int wl = BWD ? (2 - g) : g;
          for (int h = 0; h < 2; ++h) {
            for (int kb = 0; kb < 8; ++kb) {
              mbar_wait(bar(BAR_WEMPTY + s), ph ^ 1);
              if (elect_one()) {
                if (CHAIN_DBG(2)) {
                  if (leader) mbar_arrive(bar(BAR_WFULL + s));
                } else {
                  // the leader's barrier collects both halves of the 256-row weight tile
                  if (leader) mbar_arrive_expect_tx(bar(BAR_WFULL + s), 2 * STAGE_BYTES);
                  tma_load_2d_2cta(base + OFF_W + s * STAGE_BYTES, &maps.w, full_bar0 + 8u * s, kb * 64,
                                   wl * HID + h * 256 + (int)crank * 128);
                }
              }
              __syncwarp();
              if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer (leader CTA of the pair) =====================
    if (leader) {
      const uint32_t idesc = make_idesc_f16(256, 256, FMT_F16, FMT_F16, MAJOR_K, MAJOR_K);
      uint32_t s = 0, ph = 0, ev = 0;   // ev = operand-write event counter (a_ready parity)
#ifdef NEUCAM_DEBUG
      long long* trm = (lane == 0) ? p.trace : nullptr;
#endif
      int tile_i = 0;
      uint32_t dyp = 0;   // parity of the dy-operand barriers (one phase per tile)
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x, ++tile_i) {
        if (BWD) {
          mbar_wait(bar(BAR_DYR), dyp);
          mbar_wait(bar(BAR_PEERDY), dyp);
          dyp ^= 1;
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar(BAR_WFULL + s), ph);
            tc_fence_after_sync();
            if (elect_one()) {
              // A = the dy operand in k-block slot 7 (columns 0..15), B = this half of W4^T: a single K = 16 MMA
              umma_f16_ss_2cta(tmem_base + h * 256, make_sdesc_sw128(base + OFF_A + 7 * A_BLOCK_BYTES, 0, 1024),
                               make_sdesc_sw128(base + OFF_W + s * STAGE_BYTES, 0, 1024), idesc, 0u);
              umma_commit_2cta(bar(BAR_WEMPTY + s), PAIR_MASK);
              if (h == 1) {
                umma_commit_2cta(bar(BAR_KFREE + 0), PAIR_MASK);
                umma_commit_2cta(bar(BAR_KFREE + 1), PAIR_MASK);
              }
              umma_commit_2cta(bar(BAR_ACC + 2 * h), PAIR_MASK);
              umma_commit_2cta(bar(BAR_ACC + 2 * h + 1), PAIR_MASK);
            }
            __syncwarp();
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
          }
        }
        for (int g = 0; g < 3; ++g) {
          for (int h = 0; h < 2; ++h) {
            trace_stamp(trm, 0, tile_i, g, 2 * h, 0);
            for (int kb = 0; kb < 8; ++kb) {
              if (h == 0) {
                // the first half chases the epilogues of the previous layer k-block by k-block; accumulator
                // half 0 (chunks 0,1 = k-blocks 0..3) must be drained in BOTH CTAs before it is overwritten
                if (kb == 0) {
                  for (int k = 0; k < 4; ++k) {
                    mbar_wait(bar(BAR_AREADY + k), ev & 1);
                    trace_stamp(trm, 0, tile_i, g, 4, k);
                    mbar_wait(bar(BAR_PEER + k), ev & 1);
                    trace_stamp(trm, 0, tile_i, g, 5, k);
                  }
                } else if (kb >= 4) {
                  mbar_wait(bar(BAR_AREADY + kb), ev & 1);
                  mbar_wait(bar(BAR_PEER + kb), ev & 1);
                }
                if (kb == 0) trace_stamp(trm, 0, tile_i, g, 0, 1);
                if (kb == 7) trace_stamp(trm, 0, tile_i, g, 0, 2);
              }
              mbar_wait(bar(BAR_WFULL + s), ph);
              tc_fence_after_sync();
              if (elect_one()) {
                const uint32_t a_addr = base + OFF_A + kb * A_BLOCK_BYTES;
                const uint32_t b_addr = base + OFF_W + s * STAGE_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                  if (CHAIN_DBG(1)) break;
                  umma_f16_ss_2cta(tmem_base + h * 256, make_sdesc_sw128(a_addr + k16 * 32, 0, 1024),
                                   make_sdesc_sw128(b_addr + k16 * 32, 0, 1024), idesc, (kb | k16) != 0 ? 1u : 0u);
                }
                umma_commit_2cta(bar(BAR_WEMPTY + s), PAIR_MASK);
                // operand k-blocks (2c, 2c+1) of chunk c = 0,1 are dead once the last half has consumed them
                if (h == 1 && (kb == 1 || kb == 3)) umma_commit_2cta(bar(BAR_KFREE + (kb >> 1)), PAIR_MASK);
              }
              __syncwarp();
              if (++s == NSTAGE) { s = 0; ph ^= 1; }
            }
            if (elect_one()) {
              umma_commit_2cta(bar(BAR_ACC + 2 * h), PAIR_MASK);
              umma_commit_2cta(bar(BAR_ACC + 2 * h + 1), PAIR_MASK);
            }
            __syncwarp();
            trace_stamp(trm, 0, tile_i, g, 2 * h, 3);
          }
          ++ev;
        }
      }
    } else {
      // peer CTA: this warp has no MMAs to issue; it relays "operand k-block ready" to the leader
      uint32_t ev = 0;
      const uint32_t peer_bar0 = mapa_shared(bar(BAR_PEER), 0);
#ifdef NEUCAM_DEBUG
      long long* trm = (lane == 0) ? p.trace : nullptr;
#endif
      int tile_i = 0;
      const uint32_t peer_dy = mapa_shared(bar(BAR_PEERDY), 0);
      uint32_t dyp = 0;
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x, ++tile_i) {
        if (BWD) {
          mbar_wait(bar(BAR_DYR), dyp);
          dyp ^= 1;
          if (elect_one()) mbar_arrive_cluster(peer_dy);
          __syncwarp();
        }
        for (int e3 = 0; e3 < 3; ++e3) {
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(bar(BAR_AREADY + kb), ev & 1);
            if (elect_one()) mbar_arrive_cluster(peer_bar0 + 8u * kb);
            __syncwarp();
            // debug timeline: relay times of the event consumed by layer e3 (stored under "MMA role" of the peer)
            trace_stamp(trm, 0, tile_i, e3, 4 + (kb >> 2), kb & 3);
          }
          ++ev;
        }
      }
    }
  } else if (warp == WARP_STORE) {
    // ===================== TMA store issuer =====================
    {
      uint32_t ev = 0, tcount = 0;
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x, ++tcount) {
        const int tile = lt + (int)crank;
        const bool tile_ok = tile < p.ntiles;
        // forward: a1 (prologue), a2, a3 -- the last activation a4 is never materialised: its only reader, the head
        // weight gradient, regenerates it from the stored phase of z3.  backward: dz3, dz2, dz1, and the dz0 event that
        // only hands the k-blocks back.
        for (int e4 = 0; e4 < (BWD ? 4 : 3); ++e4) {
          const bool do_store = tile_ok && (BWD ? (e4 < 3) : STORE) && !CHAIN_DBG(16);
          for (int kb = 0; kb < 8; ++kb) {
            if (e4 < 3) mbar_wait(bar(BAR_AREADY + kb), ev & 1);
            else        mbar_wait(bar(BAR_AOUT + kb), tcount & 1);
            if (elect_one()) {
              if (do_store) {
                tma_store_2d(&maps.act[e4], base + OFF_A + kb * A_BLOCK_BYTES, kb * 64, tile * TILE_M);
                tma_store_commit();
                if (kb > 0) tma_store_wait_read1();
              }
              if (kb > 0) mbar_arrive(bar(BAR_STORED + kb - 1));
            }
            __syncwarp();
          }
          if (elect_one()) {
            if (do_store) tma_store_wait_read0();
            mbar_arrive(bar(BAR_STORED + 7));
          }
          __syncwarp();
          if (e4 < 3) ++ev;
        }
      }
      if (do_store_any && elect_one()) tma_store_wait0();
    }
  } else if (warp < 16) {
    // ===================== prologue / epilogue warps =====================
    // 16 warps: warp e owns rows 32*(e%4).. (its TMEM lane quadrant) and, of every 128-column chunk c, the 32
    // columns [128 c + 32 (e/4), +32) = one half of the 128-byte row of k-block 2c + (e/8).
    const uint32_t e = warp - FIRST_EPI_WARP;
    const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may access
    const uint32_t sub = e >> 2;             // which 32 columns of every 128-column chunk
    const uint32_t row = quad * 32 + lane;   // row inside the tile
    const uint32_t qb = (sub & 1) * 4;       // first 16-byte chunk of this thread's half row
    const uint32_t a_base = base + OFF_A;
    const uint32_t t_row = tmem_base + ((quad * 32u) << 16);
    uint32_t lp = 0;                         // layer parity (acc_full / kfree phases)
    uint32_t ev = 0;                         // operand-write event counter (stored parity)
    float gscale = 1.f;
    if (BWD) gscale = *p.scale;
    const uint32_t et = tid - FIRST_EPI_WARP * 32;          // 0..511: column owned in the dW0 column pass
    float gw0[3] = {0.f, 0.f, 0.f};                         // layer-0 weight-gradient partials (backward)
    const size_t ph_layer_stride = (size_t)p.ntiles * TILE_M * HID;
#ifdef NEUCAM_DEBUG
    long long* tr = (quad == 0 && lane == 0 && (sub & 1) == 0) ? p.trace : nullptr;
    const int role = 1 + (int)(sub >> 1);
#endif
    int tile_i = 0;

    for (int lt = lead0; lt < p.ntiles; lt += gridDim.x, ++tile_i) {
      const int tile = lt + (int)crank;
      const bool tile_ok = tile < p.ntiles;       // false: phantom tile of a half-empty pair
      const long long grow = (long long)tile * TILE_M + row;
      const bool valid = tile_ok && grow < p.Q;
      float u = 0.f, v = 0.f;
      if (valid) {
        const float2 t = *reinterpret_cast<const float2*>(p.uv + 2 * grow);
        u = t.x;
        v = t.y;
      }
      const size_t ph_row_off = ((size_t)tile * 64 * TILE_M + row) * 8;   // + chunk8 * 128 * 8

      // ---------------- prologue ----------------
      // forward: first operand tile a1 = sin(30 (uv W0^T + b0)) (SIMT, K = 2)
      // backward: only the K = 16 operand of the head dgrad (dy scaled by 30 * loss scale, fp16); dz3 is then
      //           produced by the regular chunked epilogue of a tensor-core "layer -1"
      trace_stamp(tr, role, tile_i, 0, 7, 0);
      if (!BWD) {
#pragma unroll 1
        for (uint32_t c = 0; c < 4; ++c) {
          const uint32_t col = c * 128 + sub * 32;
          const uint32_t kb = col >> 6;
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 q0 = s_p0[col + 2 * i], q1 = s_p0[col + 2 * i + 1];
            w[i] = pack_h2(__sinf(fmaf(u, q0.x, fmaf(v, q0.y, q0.z))), __sinf(fmaf(u, q1.x, fmaf(v, q1.y, q1.z))));
          }
          if (ev > 0) mbar_wait(bar(BAR_STORED + kb), (ev - 1) & 1);
          write_a_half_row(a_base, kb, row, qb, w);
          fence_proxy_async_smem();
          mbar_arrive(bar(BAR_AREADY + kb));
        }
        ++ev;
      } else if (sub == 0) {
        s_uv[row] = make_float2(u, v);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        if (valid) {
          const float* dyr = p.dy + (size_t)grow * p.out_dim;
          const float sc = gscale * 30.f;
          d0 = dyr[0] * sc;
          d1 = dyr[1] * sc;
          d2 = dyr[2] * sc;
          if (p.out_dim > 3) d3 = dyr[3] * sc;
        }
        // k-block slot 7 is free here: the previous tile's column pass over dz0 ended with epi_bar_sync
        const uint32_t blk7 = a_base + 7 * A_BLOCK_BYTES;
        st_shared_v4(blk7 + sw128_offset(row, 0), pack_h2_sat(d0, d1), pack_h2_sat(d2, d3), 0u, 0u);
        st_shared_v4(blk7 + sw128_offset(row, 1), 0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        mbar_arrive(bar(BAR_DYR));
      }
      trace_stamp(tr, role, tile_i, 0, 7, 1);

      // ---------------- tensor-core layers (backward: g = -1 is the head dgrad) ----------------
#pragma unroll 1
      for (int g = BWD ? -1 : 0; g < 3; ++g) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};   // head (fwd) / duv (bwd) partials
        const float* b30 = s_b30 + (g < 0 ? 0 : g) * HID;
        uint16_t* ph_out = p.phase + (size_t)(g < 0 ? 0 : g) * ph_layer_stride + ph_row_off;
        const uint16_t* ph_in = p.phase + (size_t)(g < 2 ? 1 - g : 0) * ph_layer_stride + ph_row_off;

        // order: A(0) A(1) B(0) B(1) A(2) A(3)   (A = drain + math, B = parked result -> operand).
        // Chunks 0,1 (accumulator half 0) are drained and parked while half 1 is still on the tensor cores; their
        // k-blocks 0..3 free up early in half 1 (kfree), so B(0), B(1) also run under it and the next layer's
        // first half can start the moment half 1 retires.  Chunks 2,3 then write their k-blocks directly.
#pragma unroll 1
        // forward, last layer: nothing consumes a4 on the chip (no next GEMM) or off it (see the store warp): the four A
        // steps only feed the head accumulators (and the phase stream) -- no parking, no B steps, no operand writes
        const bool last_fwd = !BWD && g == 2;
        for (int step = 0; step < 6; ++step) {
          const bool is_b = (step == 2 || step == 3);
          if (last_fwd && is_b) continue;
          const uint32_t c = is_b ? (step - 2) : (step < 2 ? step : step - 2);
          const bool park = (step < 2);
          const uint32_t col = c * 128 + sub * 32;
          const uint32_t kb = col >> 6;
          uint32_t w[16];
          trace_stamp(tr, role, tile_i, g, step, 0);
          if (!is_b) {
            uint4 pv[4];
            if (BWD && g < 2) {
              // phases do not depend on the MMA: fetch them while waiting for the accumulator
              const uint16_t* src = ph_in + (size_t)(col >> 3) * (TILE_M * 8);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                pv[q] = (tile_ok && !CHAIN_DBG(8)) ? *reinterpret_cast<const uint4*>(src + (size_t)q * (TILE_M * 8)) : make_uint4(0, 0, 0, 0);
            }
            mbar_wait(bar(BAR_ACC + c), lp);
            tc_fence_after_sync();
            trace_stamp(tr, role, tile_i, g, step, 1);
            // the 32 columns are drained as two 16-column loads: the second is in flight while the first is worked on
            uint32_t ra[16], rb[16];
            tmem_ld_32x32b_x16(t_row + col, ra);
            tmem_ld_wait();
            tmem_ld_32x32b_x16(t_row + col + 16, rb);
            if (CHAIN_DBG(4)) {
              // debug: no epilogue math (isolates tensor-pipe rate from SIMT/LDS contention)
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                w[i] = ra[2 * i];
                w[8 + i] = rb[2 * i];
              }
            } else if (!BWD) {
              uint16_t* ph_dst = ph_out + (size_t)(col >> 3) * (TILE_M * 8);
              if (g == 2) {
                fwd_half_math<STORE, true>(ra, w, b30 + col, s_w4 + col, ph_dst, tile_ok, acc);
                tmem_ld_wait();
                fwd_half_math<STORE, true>(rb, w + 8, b30 + col + 16, s_w4 + col + 16, ph_dst + 2 * (TILE_M * 8), tile_ok, acc);
              } else {
                fwd_half_math<STORE, false>(ra, w, b30 + col, s_w4 + col, ph_dst, tile_ok, acc);
                tmem_ld_wait();
                fwd_half_math<STORE, false>(rb, w + 8, b30 + col + 16, s_w4 + col + 16, ph_dst + 2 * (TILE_M * 8), tile_ok, acc);
              }
            } else {
              if (g < 2) {
                bwd_half_math<false>(ra, w, pv, s_p0 + col, u, v, acc);
                tmem_ld_wait();
                bwd_half_math<false>(rb, w + 8, pv + 2, s_p0 + col + 16, u, v, acc);
              } else {
                bwd_half_math<true>(ra, w, pv, s_p0 + col, u, v, acc);
                tmem_ld_wait();
                bwd_half_math<true>(rb, w + 8, pv + 2, s_p0 + col + 16, u, v, acc);
              }
            }
            trace_stamp(tr, role, tile_i, g, step, 2);
            if (last_fwd) continue;
            if (park) {
              // park the packed result in the accumulator columns this thread has just drained
              tmem_st_32x32b_x16(t_row + col, w);
              tmem_st_wait();
              trace_stamp(tr, role, tile_i, g, step, 3);
              continue;
            }
          } else {
            tmem_ld_32x32b_x16(t_row + col, w);
            tmem_ld_wait();
            mbar_wait(bar(BAR_KFREE + c), lp);
            trace_stamp(tr, role, tile_i, g, step, 1);
          }
          // operand write of this thread's half row of k-block kb (B steps and chunks 2,3)
          mbar_wait(bar(BAR_STORED + kb), (ev - 1) & 1);
          if (is_b) trace_stamp(tr, role, tile_i, g, step, 2);
          write_a_half_row(a_base, kb, row, qb, w);
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(bar((g < 2 ? BAR_AREADY : BAR_AOUT) + kb));
          trace_stamp(tr, role, tile_i, g, step, 3);
        }
        lp ^= 1;
        if (!last_fwd) ++ev;

        if (g == 2) {
          // combine the 4 column quarters of every row: head output (forward) / coordinate gradient (backward)
          s_xch[sub * TILE_M + row] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          epi_bar_sync();
          if (sub == 0) {
            float4 o = s_xch[row];
#pragma unroll
            for (int q = 1; q < 4; ++q) {   // fixed order: results are bit-reproducible run to run
              const float4 t = s_xch[q * TILE_M + row];
              o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            if (valid) {
              if (!BWD) {
                float* yr = p.y + (size_t)grow * p.out_dim;
                yr[0] = o.x + p.b4[0];
                yr[1] = o.y + p.b4[1];
                yr[2] = o.z + p.b4[2];
                if (p.out_dim > 3) yr[3] = o.w + p.b4[3];
              } else {
                // s_p0 holds 30*W0 and dz carries the loss scale: undo both
                const float inv = 1.f / (30.f * gscale);
                *reinterpret_cast<float2*>(p.duv + 2 * grow) = make_float2(o.x * inv, o.y * inv);
              }
            }
          }
          if (BWD) {
            // dW0[j,:] += sum_r dz0[r,j] * uv[r,:], db0[j] += sum_r dz0[r,j]: column pass over the dz0 tile
            // that now sits in the operand buffer (never stored to HBM).  epi_bar_sync above ordered all writes.
            const uint32_t cc = et;
            const uint32_t col_addr = a_base + (cc >> 6) * A_BLOCK_BYTES + (cc & 7) * 2;
            const uint32_t chunk = (cc & 63) >> 3;
#pragma unroll 8
            for (uint32_t r2 = 0; r2 < TILE_M; ++r2) {
              unsigned short hv;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(col_addr + sw128_offset(r2, chunk)));
              const float z = __half2float(__ushort_as_half(hv));
              const float2 t = s_uv[r2];
              gw0[0] = fmaf(z, t.x, gw0[0]);
              gw0[1] = fmaf(z, t.y, gw0[1]);
              gw0[2] += z;
            }
          }
          epi_bar_sync();   // the next tile's prologue overwrites the operand tile, s_uv and re-uses s_xch
        }
      }
    }
    if (BWD) {
      // this CTA's share of dW0 / db0 (still carrying the loss scale): one record per CTA, summed in CTA order by
      // reduce_partials_kernel -- no float atomics, bit-reproducible
      float* rec = p.part + (size_t)blockIdx.x * (3 * HID);
      *reinterpret_cast<float2*>(rec + 2 * et) = make_float2(gw0[0], gw0[1]);
      rec[2 * HID + et] = gw0[2];
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // neither CTA exits (or frees TMEM) while the pair's MMAs / remote arrives are in flight
  if (warp == WARP_ALLOC) tmem_dealloc_2cta(tmem_base, 512);
}

int build_maps(ChainMaps* m, const void* w16, void* const act[4], int64_t Q, bool need_act,
               const void* w4t16 = nullptr) {
  int rc = nchost::make_tmap_16b(&m->w, w16, 3 * HID, HID, HID, 128, 64);
  if (rc) return rc;
  if (w4t16 != nullptr) {
    rc = nchost::make_tmap_16b(&m->w4t, w4t16, HID, 64, 64, 128, 64);
    if (rc) return rc;
  } else {
    m->w4t = m->w;   // never dereferenced (forward)
  }
  for (int i = 0; i < 4; ++i) {
    if (!need_act) {
      m->act[i] = m->w;   // never dereferenced
      continue;
    }
    rc = nchost::make_tmap_16b(&m->act[i], act[i], (uint64_t)Q, HID, HID, TILE_M, 64);
    if (rc) return rc;
  }
  return 0;
}

#ifdef NEUCAM_DEBUG
long long* g_trace = nullptr;
int g_dbg = 0;
#endif

int chain_grid(int ntiles) {
  int grid = nchost::sm_count();
  if (grid > ntiles) grid = ntiles;
  return (grid + PAIR - 1) / PAIR * PAIR;   // whole pairs; an odd tile count gives one pair a phantom tile
}

template <bool BWD, bool STORE>
int launch(const ChainMaps& maps, const ChainParams& p, cudaStream_t stream) {
  constexpr uint32_t SMEM_BYTES = Lay<BWD>::SMEM_BYTES;
  int rc = nchost::ensure_dyn_smem((const void*)sine_chain_kernel<BWD, STORE>, (int)SMEM_BYTES);
  if (rc) return rc;
  sine_chain_kernel<BWD, STORE><<<chain_grid(p.ntiles), NTHREADS, SMEM_BYTES, stream>>>(maps, p);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

}  // namespace

#ifdef NEUCAM_DEBUG
// Debug: when non-null, CTA 0 of the next chain launches writes clock64 stamps of its MMA / epilogue roles
// into trace[6*2*3*8*4] (tools/chain_timeline.py decodes it).
extern "C" int neucam_debug_set_chain_trace(long long* trace_dev) {
  g_trace = trace_dev;
  return 0;
}
// Debug experiments on the chain kernels (results are garbage): bit0 = no MMA issue, bit1 = no weight loads.
extern "C" int neucam_debug_set_chain_flags(int flags) {
  g_dbg = flags;
  return 0;
}
#endif

extern "C" int neucam_sine_mlp_fwd(const float* uv, int64_t Q, const float* w0, const float* b0,
                                   const void* w_hid16, const float* b_hid, const float* w4,
                                   const float* b4, int out_dim, float* y, void* act16, void* phase16,
                                   void* stream) {
  NC_CHECK_ARG(uv && w0 && b0 && w_hid16 && b_hid && w4 && b4 && y);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - TILE_M);
  NC_CHECK_ARG(out_dim == 3 || out_dim == 4);
  NC_CHECK_ARG((act16 == nullptr) == (phase16 == nullptr));
  ChainParams p{};
  p.Q = (int)Q;
  p.ntiles = (int)((Q + TILE_M - 1) / TILE_M);
  p.out_dim = out_dim;
  p.store_act = act16 != nullptr;
  p.uv = uv; p.w0 = w0; p.b0 = b0; p.b_hid = b_hid; p.w4 = w4; p.b4 = b4;
  p.phase = reinterpret_cast<uint16_t*>(phase16);
  p.y = y;
#ifdef NEUCAM_DEBUG
  p.trace = g_trace;
  p.dbg = g_dbg;
#endif
  ChainMaps maps;
  void* act[4];
  for (int i = 0; i < 4; ++i) act[i] = act16 ? (void*)((__half*)act16 + (size_t)(i < 3 ? i : 2) * Q * HID) : nullptr;
  int rc = build_maps(&maps, w_hid16, act, Q, p.store_act);
  if (rc) return rc;
  return p.store_act ? launch<false, true>(maps, p, (cudaStream_t)stream)
                     : launch<false, false>(maps, p, (cudaStream_t)stream);
}

extern "C" int neucam_sine_mlp_bwd(const float* uv, int64_t Q, const float* w0, const float* b0,
                                   const void* wT_hid16, const void* w4T16, int out_dim, const float* dy,
                                   const void* phase16, const float* scale_dev, void* dz16, float* duv,
                                   float* dw0, float* db0, void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(uv && w0 && b0 && wT_hid16 && w4T16 && dy && phase16 && scale_dev && dz16 && duv && dw0 && db0);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - TILE_M);
  NC_CHECK_ARG(out_dim == 3 || out_dim == 4);
  NC_CHECK_WS(ws, ws_bytes, nchost::ws_bytes_sine_bwd());
  ChainParams p{};
  p.Q = (int)Q;
  p.ntiles = (int)((Q + TILE_M - 1) / TILE_M);
  p.out_dim = out_dim;
  p.store_act = 1;
  p.uv = uv; p.w0 = w0; p.b0 = b0; p.w4 = nullptr;
  p.phase = reinterpret_cast<uint16_t*>(const_cast<void*>(phase16));
  p.dy = dy; p.duv = duv; p.scale = scale_dev; p.part = reinterpret_cast<float*>(ws);
#ifdef NEUCAM_DEBUG
  p.trace = g_trace;
  p.dbg = g_dbg;
#endif
  ChainMaps maps;
  // store order of the backward chain: dz3 (prologue), dz2, dz1 -> dz16 = [3,Q,512] holding dz_1..dz_3
  // (dz0 never leaves the SM: its weight gradient is reduced in-kernel)
  void* act[4];
  for (int i = 0; i < 3; ++i) act[i] = (void*)((__half*)dz16 + (size_t)(2 - i) * Q * HID);
  act[3] = act[2];
  int rc = build_maps(&maps, wT_hid16, act, Q, true, w4T16);
  if (rc) return rc;
  rc = launch<true, true>(maps, p, (cudaStream_t)stream);
  if (rc) return rc;
  // dW0 += sum over CTAs (in CTA order) of their records / loss scale; same for db0
  nchost::ReduceSegs segs{};
  segs.nseg = 2;
  segs.seg[0] = {dw0, 0, 2 * HID};
  segs.seg[1] = {db0, 2 * HID, HID};
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), chain_grid(p.ntiles), 3 * HID, segs,
                                        scale_dev, (cudaStream_t)stream);
}
