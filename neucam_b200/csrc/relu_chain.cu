// relu_chain.cu -- the 256-wide ReLU MLPs of the layered configs (uv-mapping and alpha networks:
// reference modules.py:255-306 `SimpleMLP`, called at training.py:131-137,148-150 and utils.py:418-419),
// forward and input/activation gradient, as ONE persistent tcgen05 kernel per direction.
//
//   forward   feat[Q,64] --W0--> relu --W1--> relu ... --W_{H-1}--> relu --w_last (SIMT)--> tanh|sigmoid
//   backward  dpre[Q,out] --w_last (SIMT)--> mask --W_{H-1}^T--> mask ... --W_1^T--> mask [--W_0^T--> dfeat]
//
// A CTA owns a 128-sample tile for the whole chain; two CTAs form a pair (cluster of 2, tcgen05 cta_group::2).  The
// activation tile never touches shared memory: it lives in TENSOR MEMORY as the A operand of tcgen05.mma (lane = sample
// row, two fp16 per 32-bit column: 128 columns for the hi parts, 128 for the lo parts, next to the 256 accumulator
// columns = all 512 columns of the SM).  Every layer is a sequence of weight "stages" of 256 x 64 fp16 of which each CTA
// of the pair stages HALF (128 rows, 16 KB) through its 8-deep TMA ring; the leader's MMA thread issues M256 x N256 x K16
// MMAs over both CTAs' operand tiles (a single-CTA M128 version spent 183 cycles per MMA against 131 at the tensor rate:
// the full 8 KB B tile per MMA plus the ring refill saturate one SM's shared memory).  16 epilogue warps drain the
// accumulator, apply bias + ReLU (forward) or the forward's gate bits (backward), write the next operand tile back into
// tensor memory with tcgen05.st and TMA-store it to HBM for the weight-gradient GEMMs through 1 KB staging slots per
// warp (32 rows x 16 columns, SWIZZLE_32B; direct st.global from the row-per-thread layout cost 32 L1 tag requests per
// instruction and 8 k cycles per GEMM).  The encoded input (raw coordinates + NeRF positional encoding, zero-padded to
// 64 features) is a 16 KB shared-memory operand tile loaded by TMA, so the first layer and the skip connection are
// tensor-core GEMMs too.
//
// Epilogue / MMA overlap.  All 512 TMEM columns hold ONE tile, so a second tile cannot be in flight; instead the NEXT
// GEMM of the same tile chases the epilogue k-block by k-block.  Thread (lane quadrant, sub) owns 16 columns of EVERY
// 64-column k-block; an epilogue first pulls its 64 accumulator values into registers (ACCFREE: the accumulator may be
// overwritten), then produces k-block 0, hands it over (AREADY[0]), k-block 1, ... -- the next GEMM's MMAs for k-block kb
// run while k-blocks kb+1.. are still being produced, and the copies for HBM are staged and stored after the last
// hand-off.  The peer CTA's MMA warp has no MMAs to issue; it relays its CTA's ACCFREE / AREADY / feat-tile events to the
// leader.  Measured (Q = 165 888, 4-layer network): forward 178 -> 152 us, backward 166 -> 133 us against the version
// whose whole epilogue and whole GEMM took turns (one CTA per tile: 178 -> 156 / 166 -> 140; pairs: the rest).
//
// Precision.  A ReLU network's gradient is discontinuous in its pre-activations: with plain fp16 operands the
// forward's ~4e-4 relative error in z flips ~3e-4 of the ReLU gates against the fp32 reference, which alone puts
// ~2 % error on every gradient (measured).  The FORWARD therefore runs split-fp16: activations and weights are
// carried as hi + lo fp16 pairs and each product is three MMAs (hi*hi + lo*hi + hi*lo, fp32 accumulation), i.e.
// ~2^-21 relative.  The backward is smooth in its operands, but the rigidity loss (utils.py:402-454) differences
// the network at neighbouring pixels, so per-sample rounding of dz / h (random, ~3e-4) is amplified ~30x in the
// summed weight gradients (measured 1.7e-2 on a first-layer bias, 77x amplification): dz, the transposed weights
// and the saved activations are carried as hi + lo pairs as well (three MMAs per product), with a power-of-two
// loss scale on dz.
//
// Warp roles: 0..15 epilogue (warp e: TMEM lane quadrant e & 3 = 32 sample rows, sub = e >> 2 = which 16 columns of every
// k-block), 16 TMA producer, 17 MMA issuer (leader) / event relay (peer) + TMEM owner (control warps have the highest
// ids).  What bounds it now: the hand-off latency at the start of every GEMM (accumulator complete -> 64 columns per
// thread to registers -> first k-block written and published, through the peer relay -> first MMA), ~1.5 us of a
// ~5.8 us layer; the MMAs themselves are ~2.6 us of it.
#include "host_util.h"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int WID = 256;
constexpr int TILE_M = 128;
constexpr int NSTAGE = 8;                        // weight stages in flight (8 x 16 KB per CTA of the pair)
constexpr int MAX_GEMM = 7;
constexpr int MAX_STAGES = 64;
constexpr int FEAT = 64;
constexpr uint32_t KB_BYTES = TILE_M * 128;      // one k-block of the operand tile: 128 rows x 64 fp16
constexpr uint32_t STAGE_BYTES = (WID / 2) * 128;   // this CTA's half of a weight stage: 128 of its 256 rows x 64 fp16
constexpr uint16_t PAIR_MASK = 0x3;
constexpr uint32_t OFF_F = 0;                                    // encoded-input operand tile (forward)
constexpr uint32_t OFF_W = OFF_F + KB_BYTES;
constexpr uint32_t OFF_STAGE = OFF_W + NSTAGE * STAGE_BYTES;     // per epilogue warp: 32 rows x 64 B hi | lo
constexpr uint32_t OFF_BIAS = OFF_STAGE + 16 * 4096;             // float [MAX_GEMM][256] (forward)
// tensor-memory columns: accumulator | hi operand tile | lo operand tile
constexpr uint32_t TM_ACC = 0, TM_AHI = 256, TM_ALO = 384;
constexpr uint32_t OFF_HEAD = OFF_BIAS + MAX_GEMM * WID * 4;     // float [4][256] last-layer weights
constexpr uint32_t OFF_XCH = OFF_HEAD + 4 * WID * 4;              // float4 [3][128]: output-layer partial sums
constexpr uint32_t OFF_BAR = OFF_XCH + 3 * TILE_M * 16;
constexpr uint32_t OFF_TMEM = OFF_BAR + 256;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
constexpr int NEPI_WARPS = 16;                    // 4 per scheduler: the epilogue is SIMT-latency bound with fewer
constexpr int NTHREADS = (NEPI_WARPS + 2) * 32;
constexpr int WARP_PRODUCER = NEPI_WARPS, WARP_MMA = NEPI_WARPS + 1;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// ACC: a GEMM's accumulator is complete (MMA -> epilogue).  ACCFREE: all 512 epilogue threads hold their accumulator
// columns in registers (epilogue -> MMA: the next GEMM may overwrite the accumulator).  AREADY[kb]: k-block kb (64 hidden
// units) of the next operand tile is written (every epilogue thread owns 16 columns of every k-block).
// Two CTAs form a pair (cluster of 2, tcgen05 cta_group::2): the leader's MMA thread issues M256 x N256 x K16 MMAs over
// both CTAs' 128-row tiles (each operand tile in its own CTA's tensor memory) against a weight stage of which each CTA
// stages half; the peer CTA's MMA warp relays its epilogue's ACCFREE / AREADY / FFULL events to the leader (PEER_*).
enum { BAR_WFULL = 0, BAR_WEMPTY = 8, BAR_ACC = 16, BAR_ACCFREE = 17, BAR_FFULL = 18, BAR_FEMPTY = 19, BAR_AREADY = 20,
       BAR_PEER_ACCFREE = 24, BAR_PEER_FFULL = 25, BAR_PEER_AREADY = 26 };

struct ReluMaps {
  CUtensorMap feat;            // [Q,64] fp16, box 64 x 128 (forward)
  CUtensorMap w;               // weight stages [S*256, 64] fp16, box 64 x 256
  CUtensorMap out_hi[MAX_GEMM];  // forward h_l / backward dz_l: [Q,256] fp16, SWIZZLE_64B box 32 x 32 (hi parts)
  CUtensorMap out_lo[MAX_GEMM];  // the same tensors' lo parts (fp16 rounding residuals)
};

struct ReluParams {
  int Q, ntiles, n_gemm, H;            // H = number of saved activation layers (= forward GEMMs)
  int stage_begin[MAX_GEMM + 1];
  uint8_t a_src[MAX_STAGES];           // per stage: 0..3 = k-block of the operand tile, 4 = encoded-input tile,
  uint8_t a_src2[MAX_STAGES];          //   5..8 = k-block of the lo operand tile; a_src2 = 255: single product
  int feat_last_gemm;                  // forward: last GEMM of a tile that reads the encoded-input tile
  const float* bias;                   // forward: [n_gemm][256]
  const float* w_last;                 // [out_dim][256] fp32
  const float* b_last;                 // [out_dim]
  int out_dim, out_kind;               // 0 none, 1 tanh, 2 sigmoid
  int store;                           // TMA-store the operand tiles (activations / dz) to HBM
  int store_lo;                        // ... their lo parts too
  int lo_operand;                      // backward: carry dz as hi + lo (forward always does for the activations)
  float* y;                            // forward [Q,out_dim]
  const float* dpre;                   // backward [Q,out_dim]: dL/d(pre-activation of the last layer)
  uint32_t* gates_out;                 // forward [H][Q][8]: one bit per hidden unit, set where h > 0 (or null)
  const uint32_t* gates;               // backward: the same bits (ReLU masks)
  const float* scale;                  // backward: device scalar loss scale
  float* dfeat;                        // backward [Q,64] fp32 or null (the last GEMM then computes it)
  long long* trace;                    // diagnostics: CTA 0 writes (event id, clock64) pairs per role, or null
  int dbg;                             // timing experiments (results are garbage): 1 no operand write-back,
                                       // 2 no MMA issue, 4 no weight loads, 8 no accumulator drain
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// One warp's 32 rows x 16 columns of one k-block (hi and lo) -> a 1 KB staging slot each -> TMA stores.  The caller
// has made sure that earlier stores from the slot have finished READING it.
__device__ __forceinline__ void store_kb_tma(const CUtensorMap* m_hi, const CUtensorMap* m_lo, uint32_t slot,
                                             uint32_t lane, int col0, int row0, const uint32_t (&w)[8],
                                             const uint32_t (&wl)[8], bool with_lo) {
  const uint32_t hi = slot + lane * 32u, lo = hi + 1024u;
  const uint32_t sw = (lane >> 2) & 1u;   // SWIZZLE_32B: 16-byte piece index ^ address bit 7 = (row >> 2) & 1
  st_shared_v4(hi + (sw << 4), w[0], w[1], w[2], w[3]);
  st_shared_v4(hi + ((sw ^ 1u) << 4), w[4], w[5], w[6], w[7]);
  if (with_lo) {
    st_shared_v4(lo + (sw << 4), wl[0], wl[1], wl[2], wl[3]);
    st_shared_v4(lo + ((sw ^ 1u) << 4), wl[4], wl[5], wl[6], wl[7]);
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one()) {
    tma_store_2d(m_hi, slot, col0, row0);
    if (with_lo) tma_store_2d(m_lo, slot + 1024u, col0, row0);
    tma_store_commit();
  }
}

// trace[role][0] = number of events; events at trace[role][1 + 2 i] = id, [2 + 2 i] = clock64.  256 events per role.
#ifndef NEUCAM_DEBUG
#define stamp(...) ((void)0)
#define RELU_DBG(bit) false
#else
#define RELU_DBG(bit) ((p.dbg & (bit)) != 0)
__device__ __forceinline__ void stamp(long long* trace, int role, int id) {
  if (trace == nullptr || blockIdx.x != 0 || (threadIdx.x & 31) != 0) return;
  long long* t = trace + role * 520;
  const long long n = t[0];
  if (n < 256) {
    t[1 + 2 * n] = id;
    t[2 + 2 * n] = clock64();
    t[0] = n + 1;
  }
}
#endif

template <bool BWD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
relu_chain_kernel(const __grid_constant__ ReluMaps maps, const __grid_constant__ ReluParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  float* s_bias = reinterpret_cast<float*>(sm + OFF_BIAS);
  float* s_head = reinterpret_cast<float*>(sm + OFF_HEAD);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);
  const uint32_t bar0 = base + OFF_BAR;
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (!BWD)
    for (int i = tid; i < p.n_gemm * WID; i += NTHREADS) s_bias[i] = p.bias[i];
  for (int i = tid; i < 4 * WID; i += NTHREADS) s_head[i] = (i < p.out_dim * WID) ? p.w_last[i] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar(BAR_WFULL + s), 1);
      mbar_init(bar(BAR_WEMPTY + s), 1);
    }
    mbar_init(bar(BAR_ACC), 1);
    mbar_init(bar(BAR_ACCFREE), NEPI_WARPS * 32);
    for (int k = 0; k < 4; ++k) mbar_init(bar(BAR_AREADY + k), NEPI_WARPS * 32);
    mbar_init(bar(BAR_FFULL), 1);
    mbar_init(bar(BAR_FEMPTY), 1);
    mbar_init(bar(BAR_PEER_ACCFREE), 1);
    mbar_init(bar(BAR_PEER_FFULL), 1);
    for (int k = 0; k < 4; ++k) mbar_init(bar(BAR_PEER_AREADY + k), 1);
    mbar_fence_init();
  }
  if (warp == WARP_MMA) {
    tmem_alloc_2cta(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before any remote arrive / commit / TMA signal
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int S = p.stage_begin[p.n_gemm];
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  // a pair walks tiles (2j, 2j+1) + it * gridDim.x together; it runs while its leader's tile exists.  When only the
  // peer's tile is past the end it runs a phantom tile (all rows invalid: TMA loads zero-fill, stores are clipped).
  const int lead0 = (int)(blockIdx.x & ~1u);

  if (warp == WARP_PRODUCER) {
    // ---------------- TMA producer: encoded-input tile (forward) + the weight stages of every GEMM ----------------
    uint32_t slot = 0, ph = 0, fph = 0;
    const uint32_t full0 = mapa_shared(bar(BAR_WFULL), 0);   // the leader's w_full barriers (shared::cluster addresses)
    for (int lt = lead0; lt < p.ntiles; lt += gridDim.x) {
      const int tile = lt + (int)crank;
      if (!BWD) {
        mbar_wait(bar(BAR_FEMPTY), fph ^ 1);
        fph ^= 1;
        if (elect_one()) {
          mbar_arrive_expect_tx(bar(BAR_FFULL), KB_BYTES);
          tma_load_2d(base + OFF_F, &maps.feat, bar(BAR_FFULL), 0, tile * TILE_M);
        }
        __syncwarp();
      }
      for (int s = 0; s < S; ++s) {
        mbar_wait_backoff(bar(BAR_WEMPTY + slot), ph ^ 1, 32);
        stamp(p.trace, 0, 100 + s);
        if (elect_one()) {
          if (RELU_DBG(4)) {
            if (leader) mbar_arrive(bar(BAR_WFULL + slot));
          } else {
            // the leader's barrier collects both halves of the 256-row weight stage
            if (leader) mbar_arrive_expect_tx(bar(BAR_WFULL + slot), 2 * STAGE_BYTES);
            tma_load_2d_2cta(base + OFF_W + slot * STAGE_BYTES, &maps.w, full0 + 8u * slot, 0,
                             s * WID + (int)crank * (WID / 2));
          }
        }
        __syncwarp();
        if (++slot == NSTAGE) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == WARP_MMA) {
    // ---------------- MMA issuer (leader CTA) / event relay (peer CTA) ----------------
    if (leader) {
      const uint32_t idesc = make_idesc_f16(2 * TILE_M, WID, FMT_F16, FMT_F16, MAJOR_K, MAJOR_K);
      uint32_t slot = 0, ph = 0, aph = 0, fph = 0, free_ph = 0;
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x) {
        for (int g = 0; g < p.n_gemm; ++g) {
          stamp(p.trace, 1, 10 * g + 0);
          // the previous GEMM's accumulator is in the epilogue threads' registers (both CTAs): it may be overwritten.
          // The operand k-blocks are waited for one by one below, so this GEMM's first MMAs run under the epilogue
          // that is still producing the later k-blocks.
          mbar_wait_backoff(bar(BAR_ACCFREE), free_ph, 32);
          mbar_wait(bar(BAR_PEER_ACCFREE), free_ph);
          free_ph ^= 1;
          stamp(p.trace, 1, 10 * g + 1);
          const bool reads_tile = BWD || g > 0;      // the forward's first GEMM reads the encoded-input tile only
          uint32_t have = reads_tile ? 0u : 0xFu;    // operand k-blocks already waited for
          if (!BWD && g == 0) {
            mbar_wait(bar(BAR_FFULL), fph);
            mbar_wait(bar(BAR_PEER_FFULL), fph);
            fph ^= 1;
          }
          tc_fence_after_sync();
          for (int s = p.stage_begin[g]; s < p.stage_begin[g + 1]; ++s) {
            const uint32_t srcs[2] = {p.a_src[s], p.a_src2[s]};
            if (srcs[0] < 4u && !(have & (1u << srcs[0]))) {
              mbar_wait(bar(BAR_AREADY + srcs[0]), aph);
              mbar_wait(bar(BAR_PEER_AREADY + srcs[0]), aph);
              have |= 1u << srcs[0];
              tc_fence_after_sync();
            }
            mbar_wait(bar(BAR_WFULL + slot), ph);
            tc_fence_after_sync();
            if (elect_one()) {
              const uint32_t b_addr = base + OFF_W + slot * STAGE_BYTES;
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const uint32_t src = srcs[t];
                if (src == 255u) continue;
                // operand k-block: 64 K-elements = 32 tensor-memory columns; one K16 step = 8 columns
                const uint32_t a_tm = tmem_base + (src < 4 ? TM_AHI + src * 32 : TM_ALO + (src - 5) * 32);
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                  if (RELU_DBG(2)) continue;
                  const uint32_t acc = (s > p.stage_begin[g] || t > 0 || k16 > 0) ? 1u : 0u;
                  const uint64_t bd = make_sdesc_sw128(b_addr + k16 * 32, 0, 1024);
                  if (src == 4)
                    umma_f16_ss_2cta(tmem_base + TM_ACC, make_sdesc_sw128(base + OFF_F + k16 * 32, 0, 1024), bd, idesc, acc);
                  else
                    umma_f16_ts_2cta(tmem_base + TM_ACC, a_tm + k16 * 8, bd, idesc, acc);
                }
              }
              umma_commit_2cta(bar(BAR_WEMPTY + slot), PAIR_MASK);
            }
            __syncwarp();
            if (++slot == NSTAGE) { slot = 0; ph ^= 1; }
          }
          if (reads_tile) {
            // a plan that skips a k-block still consumes its phase: the four barriers advance together
            for (uint32_t k = 0; k < 4; ++k)
              if (!(have & (1u << k))) {
                mbar_wait(bar(BAR_AREADY + k), aph);
                mbar_wait(bar(BAR_PEER_AREADY + k), aph);
              }
            aph ^= 1;
          }
          if (elect_one()) {
            if (!BWD && g == p.feat_last_gemm) umma_commit_2cta(bar(BAR_FEMPTY), PAIR_MASK);
            umma_commit_2cta(bar(BAR_ACC), PAIR_MASK);
          }
          __syncwarp();
          stamp(p.trace, 1, 10 * g + 2);
        }
      }
    } else {
      // peer CTA: no MMAs to issue; this warp hands its CTA's epilogue / TMA events to the leader's MMA thread, in
      // the order the leader waits for them
      const uint32_t r_free = mapa_shared(bar(BAR_PEER_ACCFREE), 0), r_ffull = mapa_shared(bar(BAR_PEER_FFULL), 0),
                     r_ready0 = mapa_shared(bar(BAR_PEER_AREADY), 0);
      uint32_t aph = 0, fph = 0, free_ph = 0;
      for (int lt = lead0; lt < p.ntiles; lt += gridDim.x) {
        for (int g = 0; g < p.n_gemm; ++g) {
          mbar_wait_backoff(bar(BAR_ACCFREE), free_ph, 32);
          free_ph ^= 1;
          tc_fence_after_sync();
          tc_fence_before_sync();
          if (elect_one()) mbar_arrive_cluster(r_free);
          __syncwarp();
          if (!BWD && g == 0) {
            mbar_wait(bar(BAR_FFULL), fph);
            fph ^= 1;
            if (elect_one()) mbar_arrive_cluster(r_ffull);
            __syncwarp();
          }
          if (BWD || g > 0) {
            for (uint32_t k = 0; k < 4; ++k) {
              mbar_wait(bar(BAR_AREADY + k), aph);
              tc_fence_after_sync();
              tc_fence_before_sync();
              if (elect_one()) mbar_arrive_cluster(r_ready0 + 8u * k);
              __syncwarp();
            }
            aph ^= 1;
          }
        }
      }
    }
  } else {
    // ---------------- epilogue warps: thread = (sample row = TMEM lane, 16 columns of EVERY k-block) ----------------
    // Thread (quad, sub) owns row quad * 32 + lane and, of each 64-column k-block kb, the 16 columns [64 kb + 16 sub,
    // +16).  An epilogue first pulls its 64 accumulator values into registers (-> ACCFREE), then turns them into the
    // next operand tile k-block by k-block (-> AREADY[kb]): the next GEMM's MMAs for k-block 0 run while k-blocks 1..3
    // are still being produced, instead of a whole epilogue and a whole GEMM taking turns.
    const uint32_t quad = warp & 3u, sub = warp >> 2;
    const uint32_t row = quad * 32u + lane;
    const uint32_t t_row = tmem_base + ((quad * 32u) << 16);
    float4* s_xch = reinterpret_cast<float4*>(sm + OFF_XCH);
    const uint32_t stage = base + OFF_STAGE + warp * 4096u;
    const bool with_lo = p.store_lo != 0;
    uint32_t acc_ph = 0;
    mbar_arrive(bar(BAR_ACCFREE));   // the first GEMM of the first tile finds the accumulator free
    float scale = 1.f, inv_scale = 1.f;
    if (BWD) {
      scale = *p.scale;
      inv_scale = 1.f / scale;
    }
    // one k-block share of this thread: 16 values -> hi / lo fp16 pairs -> operand tile (tensor memory).  The copies for
    // HBM (the weight-gradient GEMMs' operands) are kept in registers and stored AFTER all four k-blocks have been handed
    // to the MMA warp: staging, proxy fence and TMA issue stay off the path the next GEMM waits on.
    // (Deferring the wait for k-block kb's tensor-memory stores until after the math of kb + 1 was measured: it hides
    //  the store latency but delays the first hand-off, which is what the next GEMM waits for: forward 152 -> 160 us.)
    auto pack_kb = [&](uint32_t kb, const float (&v)[16], bool to_tmem, bool sat, uint32_t (&w)[8], uint32_t (&wl)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        w[j] = sat ? pack_h2_sat(v[2 * j], v[2 * j + 1]) : pack_h2(v[2 * j], v[2 * j + 1]);
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        wl[j] = sat ? pack_h2_sat(v[2 * j] - hi.x, v[2 * j + 1] - hi.y) : pack_h2(v[2 * j] - hi.x, v[2 * j + 1] - hi.y);
      }
      if (to_tmem) {
        if (!RELU_DBG(1)) {
          tmem_st_32x32b_x8(t_row + TM_AHI + kb * 32 + sub * 8, w);
          if (!BWD || p.lo_operand) tmem_st_32x32b_x8(t_row + TM_ALO + kb * 32 + sub * 8, wl);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        mbar_arrive(bar(BAR_AREADY + kb));   // k-block kb is in tensor memory: the MMA warp may use it
      }
    };
    // the warp's 32 rows x 4 x 16 columns (hi[, lo]) -> staging -> TMA stores; hi only: four 1 KB slots, one round;
    // hi + lo: 2 KB per k-block, two rounds of two k-blocks
    auto store_tile = [&](const uint32_t (&W)[4][8], const uint32_t (&WL)[4][8], int layer, int tile) {
      if (!p.store) return;
      const int row0 = tile * TILE_M + (int)quad * 32;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 0 || with_lo) {
          if (elect_one()) tma_store_wait_read0();   // the slots' previous stores have been read out
          __syncwarp();
        }
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
          const int kb = 2 * half + k2;
          const uint32_t slot = stage + (with_lo ? (uint32_t)k2 * 2048u : (uint32_t)kb * 1024u);
          store_kb_tma(&maps.out_hi[layer], &maps.out_lo[layer], slot, lane, kb * 64 + (int)sub * 16, row0, W[kb], WL[kb],
                       with_lo);
        }
      }
    };
    for (int lt = lead0; lt < p.ntiles; lt += gridDim.x) {
      const int tile = lt + (int)crank;
      const long long grow = (long long)tile * TILE_M + row;
      const bool valid = tile < p.ntiles && grow < p.Q;
      if (BWD) {
        // prologue: dz_{H-1} = mask(h_{H-1}) * (dpre . w_last), scaled, as the first operand tile
        float dp[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid)
          for (int o = 0; o < p.out_dim; ++o) dp[o] = p.dpre[grow * p.out_dim + o] * scale;
        const uint16_t* grow_gates = reinterpret_cast<const uint16_t*>(p.gates + ((size_t)(p.H - 1) * p.Q + grow) * 8);
        uint32_t W[4][8], WL[4][8];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t gw = valid ? grow_gates[4 * kb + sub] : 0u;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = kb * 64 + sub * 16 + j;
            float d = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) d = fmaf(dp[o], s_head[o * WID + col], d);
            v[j] = ((gw >> j) & 1u) ? d : 0.f;       // ReLU closed -> 0
          }
          pack_kb((uint32_t)kb, v, true, true, W[kb], WL[kb]);
        }
        store_tile(W, WL, p.H - 1, tile);
      }
      for (int g = 0; g < p.n_gemm; ++g) {
        const bool last = (g == p.n_gemm - 1);
        const bool dfeat_gemm = BWD && last && p.dfeat != nullptr;
        const int layer = BWD ? (p.H - 2 - g) : g;        // layer whose operand tile this epilogue produces
        const bool produces = BWD ? !last : !last;        // a later GEMM of this tile reads the tile written here
        uint32_t gw[4] = {0u, 0u, 0u, 0u};
        if (BWD && !dfeat_gemm && valid) {   // fetched before the accumulator wait: the latency hides under the MMAs
          const uint16_t* gg = reinterpret_cast<const uint16_t*>(p.gates + ((size_t)layer * p.Q + grow) * 8);
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) gw[kb] = gg[4 * kb + sub];
        }
        if (warp == 0) stamp(p.trace, 2, 10 * g + 0);
        mbar_wait_backoff(bar(BAR_ACC), acc_ph, 64);
        acc_ph ^= 1;
        if (warp == 0) stamp(p.trace, 2, 10 * g + 1);
        tc_fence_after_sync();
        if (dfeat_gemm) {
          // input gradient: first 64 accumulator columns (the sub == 0 warps), unscaled, fp32
          for (uint32_t c = 0; c < 2 && sub == 0; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + TM_ACC + c * 32, r);
            tmem_ld_wait();
            if (valid) {
              float4* dst = reinterpret_cast<float4*>(p.dfeat + grow * FEAT + c * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(__uint_as_float(r[4 * q]) * inv_scale, __uint_as_float(r[4 * q + 1]) * inv_scale,
                                     __uint_as_float(r[4 * q + 2]) * inv_scale, __uint_as_float(r[4 * q + 3]) * inv_scale);
            }
          }
          tc_fence_before_sync();
          mbar_arrive(bar(BAR_ACCFREE));
          continue;   // last GEMM of the tile
        }
        // the thread's 4 x 16 accumulator columns -> registers; from here on the accumulator is free
        uint32_t r[4][16];
        if (!RELU_DBG(8)) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) tmem_ld_32x32b_x16(t_row + TM_ACC + kb * 64 + sub * 16, r[kb]);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int j = 0; j < 16; ++j) r[kb][j] = 0x3f000000u + kb + j;
        }
        tc_fence_before_sync();
        mbar_arrive(bar(BAR_ACCFREE));
        float yacc[4] = {0.f, 0.f, 0.f, 0.f};
        uint16_t* gate_dst = (!BWD && p.gates_out != nullptr && valid)
                                 ? reinterpret_cast<uint16_t*>(p.gates_out + ((size_t)g * p.Q + grow) * 8) : nullptr;
        uint32_t W[4][8], WL[4][8];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          float v[16];
          if (!BWD) {
            const float* bz = s_bias + g * WID + kb * 64 + sub * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(r[kb][j]) + bz[j], 0.f);
            if (last) {
              const float* hw = s_head + kb * 64 + sub * 16;
#pragma unroll
              for (int j = 0; j < 16; ++j)
#pragma unroll
                for (int o = 0; o < 4; ++o) yacc[o] = fmaf(v[j], hw[o * WID + j], yacc[o]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = ((gw[kb] >> j) & 1u) ? __uint_as_float(r[kb][j]) : 0.f;
          }
          pack_kb((uint32_t)kb, v, produces, BWD, W[kb], WL[kb]);
          if (!BWD && gate_dst != nullptr) {
            // gate bit = "the STORED fp16 activation is positive"
            uint32_t gbits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              gbits |= ((W[kb][j] & 0x7FFFu) != 0u ? 1u : 0u) << (2 * j);
              gbits |= ((W[kb][j] & 0x7FFF0000u) != 0u ? 1u : 0u) << (2 * j + 1);
            }
            gate_dst[4 * kb + sub] = (uint16_t)gbits;
          }
        }
        store_tile(W, WL, layer, tile);
        if (warp == 0) stamp(p.trace, 2, 10 * g + 2);

        if (!BWD && last) {
          // output layer: combine the four column quarters of every row in a fixed order (bit-reproducible)
          if (sub > 0) s_xch[(sub - 1) * TILE_M + row] = make_float4(yacc[0], yacc[1], yacc[2], yacc[3]);
          epi_bar_sync();
          if (sub == 0 && valid) {
            float4 o4 = make_float4(yacc[0], yacc[1], yacc[2], yacc[3]);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              const float4 t = s_xch[q * TILE_M + row];
              o4.x += t.x; o4.y += t.y; o4.z += t.z; o4.w += t.w;
            }
            const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (o < p.out_dim) {
                float v = ov[o] + p.b_last[o];
                if (p.out_kind == 1) v = tanhf(v);
                else if (p.out_kind == 2) v = 1.f / (1.f + expf(-v));
                p.y[grow * p.out_dim + o] = v;
              }
            }
          }
          epi_bar_sync();   // s_xch is free again before the next tile's last GEMM
        }
      }
    }
    if (elect_one()) tma_store_wait0();
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // neither CTA exits (or frees tensor memory) while the pair's MMAs / remote arrives are in flight
  if (warp == WARP_MMA) tmem_dealloc_2cta(tmem_base, 512);
}

#ifdef NEUCAM_DEBUG
int g_relu_dbg = 0;
long long* g_relu_trace = nullptr;
#endif

// whole CTA pairs; an odd tile count gives one pair a phantom tile
int pair_grid(int ntiles) {
  int grid = nchost::sm_count();
  if (grid > ntiles) grid = ntiles;
  return (grid + 1) / 2 * 2;
}

int check_stage_plan(int n_gemm, const int* stage_begin) {
  if (n_gemm < 1 || n_gemm > MAX_GEMM) return 1;
  if (stage_begin[0] != 0) return 1;
  for (int g = 0; g < n_gemm; ++g)
    if (stage_begin[g + 1] <= stage_begin[g]) return 1;
  return stage_begin[n_gemm] > MAX_STAGES;
}

}  // namespace

#ifdef NEUCAM_DEBUG
// Timing experiments on the ReLU chain (results are garbage): see ReluParams::dbg.
extern "C" int neucam_debug_set_relu_flags(int flags) {
  g_relu_dbg = flags;
  return 0;
}
// Diagnostics: CTA 0 of the next forward launches records (event, clock64) pairs into trace [3][520] int64 (zeroed by
// the caller): role 0 producer (100 + stage), 1 MMA warp (10 g + {0 wait operand, 1 start, 2 committed}), 2 epilogue
// warp 0 (10 g + {0 wait accumulator, 1 start, 2 done}).  NULL switches it off.
extern "C" int neucam_debug_set_relu_trace(void* trace) {
  g_relu_trace = (long long*)trace;
  return 0;
}
#endif

// forward of one ReLU MLP.  feat16 [Q,64] fp16 (encoded input, zero padded); w_stages [S*256,64] fp16 (per GEMM:
// its 256 x 64 K-blocks in order); stage_begin[n_gemm+1]; a_src[S] (0..3 operand k-block, 4 encoded input) and
// a_src2[S] (5..8: the stage ALSO multiplies that k-block of the lo operand tile; 255: it does not);
// bias [n_gemm,256]; w_last [out_dim,256], b_last [out_dim] fp32; out_kind 0 none / 1 tanh / 2 sigmoid;
// y [Q,out_dim] fp32; act16 / act_lo16 [n_gemm,Q,256] fp16 (saved activations, hi and lo parts) and gates
// [n_gemm,Q,8] uint32 (bit u of a row = hidden unit u is open), or all three NULL.
extern "C" int neucam_relu_mlp_fwd(const void* feat16, int64_t Q, const void* w_stages, int n_gemm,
                                   const int* stage_begin, const uint8_t* a_src, const uint8_t* a_src2,
                                   const float* bias,
                                   const float* w_last, const float* b_last, int out_dim, int out_kind, float* y,
                                   void* act16, void* act_lo16, void* gates, void* stream) {
  NC_CHECK_ARG(feat16 && w_stages && stage_begin && a_src && a_src2 && bias && w_last && b_last && y);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - 128 && out_dim >= 1 && out_dim <= 4 && out_kind >= 0 && out_kind <= 2);
  NC_CHECK_ARG(check_stage_plan(n_gemm, stage_begin) == 0);
  NC_CHECK_ARG((act16 == nullptr) == (gates == nullptr) && (act_lo16 == nullptr || act16 != nullptr));
  ReluMaps maps;
  ReluParams p{};
  const int S = stage_begin[n_gemm];
  int rc = nchost::make_tmap_16b(&maps.feat, feat16, (uint64_t)Q, FEAT, FEAT, TILE_M, 64);
  if (rc) return rc;
  rc = nchost::make_tmap_16b(&maps.w, w_stages, (uint64_t)S * WID, 64, 64, WID / 2, 64);   // a CTA's half of a stage
  if (rc) return rc;
  p.feat_last_gemm = 0;
  for (int g = 0; g < n_gemm; ++g) {
    if (act16) {
      rc = nchost::make_tmap_16b_sw32(&maps.out_hi[g], (const __half*)act16 + (size_t)g * Q * WID, (uint64_t)Q, WID, WID, 32, 16);
      if (rc) return rc;
      if (act_lo16) {
        rc = nchost::make_tmap_16b_sw32(&maps.out_lo[g], (const __half*)act_lo16 + (size_t)g * Q * WID, (uint64_t)Q, WID, WID, 32, 16);
        if (rc) return rc;
      } else {
        maps.out_lo[g] = maps.out_hi[g];
      }
    } else {
      maps.out_hi[g] = maps.feat;
      maps.out_lo[g] = maps.feat;
    }
    p.stage_begin[g] = stage_begin[g];
    for (int s = stage_begin[g]; s < stage_begin[g + 1]; ++s) {
      NC_CHECK_ARG(a_src[s] <= 4 && (a_src2[s] == 255 || (a_src2[s] >= 5 && a_src2[s] <= 8)));
      NC_CHECK_ARG(g > 0 || (a_src[s] == 4 && a_src2[s] == 255));   // the first GEMM reads the encoded input only
      p.a_src[s] = a_src[s];
      p.a_src2[s] = a_src2[s];
      if (a_src[s] == 4) p.feat_last_gemm = g;
    }
  }
  p.stage_begin[n_gemm] = S;
  p.Q = (int)Q;
  p.ntiles = (int)((Q + TILE_M - 1) / TILE_M);
  p.n_gemm = n_gemm;
  p.H = n_gemm;
  p.bias = bias;
  p.w_last = w_last;
  p.b_last = b_last;
  p.out_dim = out_dim;
  p.out_kind = out_kind;
  p.store = act16 != nullptr;
  p.store_lo = act_lo16 != nullptr;
  p.lo_operand = 1;
  p.gates_out = (uint32_t*)gates;
  p.y = y;
#ifdef NEUCAM_DEBUG
  p.dbg = g_relu_dbg;
  p.trace = g_relu_trace;
#endif
  rc = nchost::ensure_dyn_smem((const void*)relu_chain_kernel<false>, (int)SMEM_BYTES);
  if (rc) return rc;
  const int grid = pair_grid(p.ntiles);
  relu_chain_kernel<false><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(maps, p);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}

// backward of one ReLU MLP w.r.t. its activations (and, when dfeat is given, its encoded input).
// dpre [Q,out_dim] fp32 = dL/d(last layer's pre-activation); gates [H,Q,8] uint32 saved by the forward;
// wT_stages: per backward GEMM (layer H-1 down to 1, then layer 0 when dfeat != NULL) the four 256 x 64 K-blocks of
// the TRANSPOSED weight (rows = input feature, zero padded to 256), each as a hi stage followed by a lo stage;
// scale_dev: power-of-two loss scale applied to dz;
// dz16 [H,Q,256] fp16 out (scaled); dz_lo16: its lo parts, or NULL = carry dz in single fp16 (per-sample rounding is
// random and averages out over the samples of a weight gradient; the WEIGHT stages stay split); dfeat [Q,64] fp32 or NULL.
extern "C" int neucam_relu_mlp_bwd(const float* dpre, int64_t Q, const void* wT_stages, int H, const float* w_last,
                                   int out_dim, const void* gates, const float* scale_dev, void* dz16,
                                   void* dz_lo16, float* dfeat, void* stream) {
  NC_CHECK_ARG(dpre && wT_stages && w_last && gates && scale_dev && dz16);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - 128 && out_dim >= 1 && out_dim <= 4 && H >= 1 && H <= MAX_GEMM);
  NC_CHECK_ARG(H - 1 + (dfeat ? 1 : 0) <= MAX_GEMM);
  const int n_gemm = (H - 1) + (dfeat ? 1 : 0);
  ReluMaps maps;
  ReluParams p{};
  int rc = nchost::make_tmap_16b(&maps.feat, wT_stages, 8 * WID, 64, 64, TILE_M, 64);   // unused in this direction
  if (rc) return rc;
  rc = nchost::make_tmap_16b(&maps.w, wT_stages, (uint64_t)(n_gemm > 0 ? n_gemm : 1) * 8 * WID, 64, 64, WID / 2, 64);
  if (rc) return rc;
  // per K-block: a hi stage (x dz_hi and dz_lo) and a lo stage (x dz_hi)
  for (int g = 0; g <= n_gemm; ++g) p.stage_begin[g] = 8 * g;
  for (int s = 0; s < 8 * n_gemm; ++s) {
    const int kb = (s >> 1) & 3;
    p.a_src[s] = (uint8_t)kb;
    // with dz_lo16: dz is carried as hi + lo and the hi weight stage multiplies both; without: only the transposed
    // weights are split (their rounding is common to all samples and would not average out)
    p.a_src2[s] = ((s & 1) || dz_lo16 == nullptr) ? (uint8_t)255 : (uint8_t)(5 + kb);
  }
  p.Q = (int)Q;
  p.ntiles = (int)((Q + TILE_M - 1) / TILE_M);
  p.n_gemm = n_gemm;
  p.H = H;
  p.w_last = w_last;
  p.out_dim = out_dim;
  for (int l = 0; l < MAX_GEMM; ++l) {
    if (l < H) {
      rc = nchost::make_tmap_16b_sw32(&maps.out_hi[l], (const __half*)dz16 + (size_t)l * Q * WID, (uint64_t)Q, WID, WID, 32, 16);
      if (rc) return rc;
      if (dz_lo16) {
        rc = nchost::make_tmap_16b_sw32(&maps.out_lo[l], (const __half*)dz_lo16 + (size_t)l * Q * WID, (uint64_t)Q, WID, WID, 32, 16);
        if (rc) return rc;
      } else {
        maps.out_lo[l] = maps.out_hi[l];
      }
    } else {
      maps.out_hi[l] = maps.w;
      maps.out_lo[l] = maps.w;
    }
  }
  p.store = 1;
  p.store_lo = dz_lo16 != nullptr;
  p.lo_operand = dz_lo16 != nullptr;
  p.dpre = dpre;
  p.gates = (const uint32_t*)gates;
  p.scale = scale_dev;
  p.dfeat = dfeat;
  rc = nchost::ensure_dyn_smem((const void*)relu_chain_kernel<true>, (int)SMEM_BYTES);
  if (rc) return rc;
  const int grid = pair_grid(p.ntiles);
  relu_chain_kernel<true><<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(maps, p);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  return 0;
}
