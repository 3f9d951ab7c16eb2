// wgrad.cu -- weight gradients of the scene sine-MLP (the MmBackward / sum nodes of autograd for
// BatchLinear, modules.py:24-25, reduced over all Q samples of the step).
//
//   hidden layers l = 1..3:  dW_l[j,k] = sum_r dz_l[r,j] * a_l[r,k],   db_l[j] = sum_r dz_l[r,j]
//     a split-K tcgen05 GEMM on CTA pairs (cta_group::2): both operands are read as MN-major tiles straight
//     out of the row-major [Q,512] fp16 buffers the chain kernels TMA-stored (no transposes anywhere); each
//     pair owns one 256 x 256 output tile of one layer for one slice of the Q axis and writes it as a partial record
//     into the reduction workspace; wgrad_reduce_kernel then adds the slices up in slice order (no float atomics:
//     bit-reproducible).  The bias gradient rides along as a 16-column MMA against an all-ones tile.
//   head:  dW4[c,k] = sum_r dy[r,c] * a4[r,k]  (c < 4) -- an HBM-bound SIMT column reduction, per-block partials; a4 is
//     regenerated from the stored phase of z3 (the forward chain never writes it).
//
// dz carries the power-of-two loss scale of the backward chain; it is divided out here.
#include "host_util.h"
#include "tc05.cuh"

using namespace tc05;

namespace {

constexpr int HID = 512;
constexpr int NSTAGE = 6;
constexpr uint32_t ATOM_BYTES = 64 * 128;    // 64 k-rows x 64 MN elements (128 B)
constexpr uint32_t A_STAGE = 2 * ATOM_BYTES; // 64 k-rows x 128 m  (this CTA's half of the pair's 256 m)
constexpr uint32_t B_STAGE = 2 * ATOM_BYTES; // 64 k-rows x 128 n  (this CTA's half of the pair's 256 n)
constexpr uint32_t STAGE_BYTES = A_STAGE + B_STAGE;
constexpr uint32_t OFF_ONES = NSTAGE * STAGE_BYTES;   // 8 KB of fp16 1.0
constexpr uint32_t OFF_BAR = OFF_ONES + ATOM_BYTES;
constexpr uint32_t OFF_TMEM = OFF_BAR + 128;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16 + 1024;
constexpr int NTHREADS = 192;
constexpr uint16_t PAIR_MASK = 0x3;

// One 256 x 256 output tile of one weight gradient: dw[m0 + i, n0 + j] += sum_r dz[r, m0 + i] * act[r, n0 + j].
// Columns of `act` beyond its extent are zero-filled by TMA, so narrow right-hand sides (the 64-feature encoded
// input of the ReLU MLPs) use the same tile shape and write back only `n_valid` columns.
struct WgradJob {
  CUtensorMap dz;       // [Q, *] fp16 row-major, box 64 x 64
  CUtensorMap act;      // [Q, *] fp16 row-major, box 64 x 64
  float* dw;            // fp32, row pitch ldw (accumulated into)
  float* db;            // fp32 [*] or null: db[m0 + i] += sum_r dz[r, m0 + i]
  int m0, n0, ldw, n_valid;
};
constexpr int MAX_JOBS = 24;
struct WgradJobs {
  WgradJob job[MAX_JOBS];
};

struct WgradParams {
  int Q;
  int nkb;              // ceil(Q / 64)
  int kb_per_split;
  int splits;
  const float* scale;   // device scalar: loss scale carried by dz
  float* part;          // workspace: [job][split][256 x 256 tile | 256 bias sums] partial records (still loss-scaled)
};
constexpr int REC = nchost::WS_WGRAD_REC;

// A CTA pair (cluster of 2, cta_group::2) owns one 256 x 256 tile of one layer's dW for one slice of the Q
// axis: each CTA stages its 128 dz-columns (A, M side) and its 128 activation-columns (B, N side) per 64-row
// k-block; the leader issues M256 x N256 x K16 MMAs; each CTA's TMEM holds its own 128 output rows.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
hidden_wgrad_kernel(const __grid_constant__ WgradJobs jobs, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const uint32_t bar_full = base + OFF_BAR, bar_empty = bar_full + 8 * NSTAGE, bar_acc = bar_empty + 8 * NSTAGE;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int split = blockIdx.x >> 1;
  const WgradJob& job = jobs.job[blockIdx.y];
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.nkb) kb1 = p.nkb;
  const int nk = kb1 - kb0;   // may be <= 0 for trailing splits (identical in both CTAs of a pair)
  const bool with_bias = job.db != nullptr;
  const int m0 = job.m0 + (int)crank * 128;   // this CTA's dz columns  = dW rows
  const int n0 = job.n0 + (int)crank * 128;   // this CTA's act columns = dW columns it STAGES (not the ones it owns)

  // all-ones tile for the bias-gradient MMA
  for (uint32_t i = tid; i < ATOM_BYTES / 4; i += NTHREADS)
    reinterpret_cast<uint32_t*>(sm + OFF_ONES)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc_2cta(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (nk > 0) {
    if (warp == 4) {
      // whole warp runs the loop / waits; lane 0 issues (no permanently diverged control warp)
      const CUtensorMap* ta = &job.dz;
      const CUtensorMap* tb = &job.act;
      uint32_t s = 0, ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one()) {
          const uint32_t full_leader = mapa_shared(bar_full + 8 * s, 0);
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);
          const uint32_t a_st = base + s * STAGE_BYTES, b_st = a_st + A_STAGE;
          for (int a = 0; a < 2; ++a) tma_load_2d_2cta(a_st + a * ATOM_BYTES, ta, full_leader, m0 + a * 64, kb * 64);
          for (int a = 0; a < 2; ++a) tma_load_2d_2cta(b_st + a * ATOM_BYTES, tb, full_leader, n0 + a * 64, kb * 64);
        }
        __syncwarp();
        if (++s == NSTAGE) { s = 0; ph ^= 1; }
      }
    } else if (warp == 5) {
      if (leader) {
        const uint32_t idesc = make_idesc_f16(256, 256, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
        const uint32_t idesc_b = make_idesc_f16(256, 16, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
        const uint32_t ones = base + OFF_ONES;
        uint32_t s = 0, ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_st = base + s * STAGE_BYTES, b_st = a_st + A_STAGE;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t ad = make_sdesc_sw128(a_st + k16 * 2048, ATOM_BYTES, 1024);
              const uint32_t accum = (kb > kb0 || k16 > 0) ? 1u : 0u;
              umma_f16_ss_2cta(tmem_base, ad, make_sdesc_sw128(b_st + k16 * 2048, ATOM_BYTES, 1024), idesc, accum);
              if (with_bias)
                umma_f16_ss_2cta(tmem_base + 256, ad, make_sdesc_sw128(ones + k16 * 2048, ATOM_BYTES, 1024), idesc_b, accum);
            }
            umma_commit_2cta(bar_empty + 8 * s, PAIR_MASK);
          }
          __syncwarp();
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(bar_acc, PAIR_MASK);
        __syncwarp();
      }
    } else {
      // epilogue warps 0..3 (control warps 4,5 have the highest ids = scheduling priority): TMEM lane quadrant = warp; this CTA owns dW rows m0.., all 256 columns of the tile
      const uint32_t quad = warp & 3;
      const uint32_t row = quad * 32 + lane;
      mbar_wait(bar_acc, 0);
      tc_fence_after_sync();
      // partial record of this (job, Q-slice): tile row = dW row (m0 - job.m0 + row), 256 columns; then 256 bias sums
      float* rec = p.part + ((size_t)blockIdx.y * p.splits + split) * REC;
      float* out = rec + (size_t)(crank * 128 + row) * 256;
      const uint32_t t_row = tmem_base + ((quad * 32u) << 16);
      const int nchunk = job.n_valid >> 5;
      for (int i = 0; i < nchunk; ++i) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + i * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(out + i * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      if (with_bias) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + 256, r);
        tmem_ld_wait();
        rec[256 * 256 + crank * 128 + row] = __uint_as_float(r[0]);
      }
    }
  } else if (warp < 4) {
    // an empty trailing Q-slice still owns a record: zero it so that the reduction reads defined values
    float* rec = p.part + ((size_t)blockIdx.y * p.splits + split) * REC;
    const uint32_t row = (warp & 3) * 32 + lane;
    float* out = rec + (size_t)(crank * 128 + row) * 256;
    for (int j = 0; j < job.n_valid; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (with_bias) rec[256 * 256 + crank * 128 + row] = 0.f;
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc_2cta(tmem_base, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// wide_wgrad_kernel: the 512 x 512 weight gradients of the scene network, 256 (rows) x 512 (columns) per CTA pair.
//
// hidden_wgrad_kernel covers a 512 x 512 gradient with four 256 x 256 tile jobs; each operand half is then read by two
// jobs and shared only through L2: ncu measured 3.25 GB of DRAM reads against 2.36 GB of operands (profiles/
// r2_ncu_static_summary.md) on a kernel that runs at 76 % of HBM peak -- the excess bytes are time.  Here a pair keeps
// ALL 512 columns of its 256 rows in tensor memory (2 x 256 accumulator columns per CTA = the whole TMEM), so a dz half is
// read exactly once and only the activations are shared between the two row-tile jobs of a layer.  With no accumulator
// column left for the bias MMA, db = column sums of dz is taken from the A tiles in shared memory by the four otherwise
// idle epilogue warps (thread = column, k-rows in order: fixed summation order).
//   stage (per CTA) = 64 k-rows x [128 dz columns | 256 activation columns] = 48 KB, 4 stages.
//   barriers: full[s] (leader: both CTAs' TMA bytes), mma_done[s] (tcgen05.commit, both CTAs), empty[s] (the CTA's four
//   bias warps have read the A tile): the producer refills a stage only after the MMAs AND the bias pass are done with it.
// ------------------------------------------------------------------------------------------------------------------
constexpr int WNSTAGE = 4;
constexpr uint32_t WA_STAGE = 2 * ATOM_BYTES;             // 64 k x 128 dz columns
constexpr uint32_t WB_STAGE = 4 * ATOM_BYTES;             // 64 k x 256 activation columns (this CTA's halves of both N = 256 MMAs)
constexpr uint32_t WSTAGE_BYTES = WA_STAGE + WB_STAGE;
constexpr uint32_t WOFF_BAR = WNSTAGE * WSTAGE_BYTES;
constexpr uint32_t WOFF_TMEM = WOFF_BAR + 128;
constexpr uint32_t WSMEM_BYTES = WOFF_TMEM + 16 + 1024;
constexpr int WREC = 256 * 512 + 256;

struct WideJob {
  CUtensorMap dz;       // [Q, 512] fp16 row-major, box 64 x 64
  CUtensorMap act;      // [Q, 512] fp16 row-major, box 64 x 64
  int m0;               // first of this job's 256 dW rows (= dz columns)
};
struct WideJobs {
  WideJob job[6];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
wide_wgrad_kernel(const __grid_constant__ WideJobs jobs, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const uint32_t bar_full = base + WOFF_BAR, bar_done = bar_full + 8 * WNSTAGE, bar_empty = bar_done + 8 * WNSTAGE,
                 bar_acc = bar_empty + 8 * WNSTAGE;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + WOFF_TMEM);

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int split = blockIdx.x >> 1;
  const WideJob& job = jobs.job[blockIdx.y];
  const int kb0 = split * p.kb_per_split;
  int kb1 = kb0 + p.kb_per_split;
  if (kb1 > p.nkb) kb1 = p.nkb;
  const int nk = kb1 - kb0;
  const int m0 = job.m0 + (int)crank * 128;   // this CTA's dz columns = the dW rows it owns

  if (tid == 0) {
    for (int s = 0; s < WNSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_done + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 4);
    }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc_2cta(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  float* rec = p.part + ((size_t)blockIdx.y * p.splits + split) * WREC;

  if (nk > 0) {
    if (warp == 4) {
      // ---- TMA producer ----
      uint32_t s = 0, ph = 0;
      const uint32_t full_leader0 = mapa_shared(bar_full, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * WSTAGE_BYTES);
          const uint32_t a_st = base + s * WSTAGE_BYTES, b_st = a_st + WA_STAGE;
          const uint32_t fb = full_leader0 + 8 * s;
          for (int a = 0; a < 2; ++a) tma_load_2d_2cta(a_st + a * ATOM_BYTES, &job.dz, fb, m0 + a * 64, kb * 64);
          // MMA h covers activation columns [256 h, 256 h + 256): this CTA stages its half [256 h + 128 crank, +128)
          for (int h = 0; h < 2; ++h)
            for (int a = 0; a < 2; ++a)
              tma_load_2d_2cta(b_st + (2 * h + a) * ATOM_BYTES, &job.act, fb, 256 * h + 128 * (int)crank + a * 64,
                               kb * 64);
        }
        __syncwarp();
        if (++s == WNSTAGE) { s = 0; ph ^= 1; }
      }
    } else if (warp == 5) {
      // ---- MMA issuer (leader) ----
      if (leader) {
        const uint32_t idesc = make_idesc_f16(256, 256, FMT_F16, FMT_F16, MAJOR_MN, MAJOR_MN);
        uint32_t s = 0, ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t a_st = base + s * WSTAGE_BYTES, b_st = a_st + WA_STAGE;
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t ad = make_sdesc_sw128(a_st + k16 * 2048, ATOM_BYTES, 1024);
              const uint32_t accum = (kb > kb0 || k16 > 0) ? 1u : 0u;
#pragma unroll
              for (int h = 0; h < 2; ++h)
                umma_f16_ss_2cta(tmem_base + 256 * h, ad,
                                 make_sdesc_sw128(b_st + 2 * h * ATOM_BYTES + k16 * 2048, ATOM_BYTES, 1024), idesc, accum);
            }
            umma_commit_2cta(bar_done + 8 * s, PAIR_MASK);
          }
          __syncwarp();
          if (++s == WNSTAGE) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(bar_acc, PAIR_MASK);
        __syncwarp();
      }
    } else {
      // ---- warps 0..3: bias column sums during the main loop, then the accumulator drain ----
      // thread t owns dz column m0 + t: element (k, t) of the A stage sits in atom t / 64 at k * 128 + ((t % 64) / 8 ^ k % 8) * 16
      const uint32_t col_off = (tid >> 6) * ATOM_BYTES + (tid & 7) * 2;
      const uint32_t chunk = (tid & 63) >> 3;
      float bsum = 0.f;
      uint32_t s = 0, ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_backoff(bar_done + 8 * s, ph, 64);     // the MMAs have consumed the stage: its data is (still) valid
        const uint32_t a_st = base + s * WSTAGE_BYTES + col_off;
#pragma unroll 8
        for (uint32_t k = 0; k < 64; ++k) {
          unsigned short hv;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(a_st + k * 128u + ((chunk ^ (k & 7u)) << 4)));
          bsum += __half2float(__ushort_as_half(hv));
        }
        __syncwarp();
        if (elect_one()) mbar_arrive(bar_empty + 8 * s);
        if (++s == WNSTAGE) { s = 0; ph ^= 1; }
      }
      const uint32_t quad = warp & 3;
      const uint32_t row = quad * 32 + lane;
      mbar_wait(bar_acc, 0);
      tc_fence_after_sync();
      float* out = rec + (size_t)(crank * 128 + row) * 512;
      const uint32_t t_row = tmem_base + ((quad * 32u) << 16);
      for (int i = 0; i < 16; ++i) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + i * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(out + i * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      rec[256 * 512 + crank * 128 + tid] = bsum;
    }
  } else if (warp < 4) {
    // an empty trailing Q-slice still owns a record: zero it so that the reduction reads defined values
    float* out = rec + (size_t)(crank * 128 + tid) * 512;
    for (int j = 0; j < 512; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    rec[256 * 512 + crank * 128 + tid] = 0.f;
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc_2cta(tmem_base, 512);
}

// Second pass of the split-K weight gradients: thread = (element (i, j) of the 256 x 256 tile space or one of the 256
// bias sums, job): dw[m0 + i, n0 + j] += (sum over the job's Q-slices, in slice order) / scale.  Jobs that target the
// SAME tile of the same tensor (the optional hi / lo operand splits of the ReLU MLPs) are chained on the host and
// walked in job order by the thread of the chain's first job: no atomics, no races, a fixed summation order.
struct ReduceJob {
  float* dw;
  float* db;
  int m0, n0, ldw, n_valid;
  int next;      // next job of the same destination tile, or -1
  int is_head;   // first job of its chain
};
struct ReduceJobs {
  ReduceJob job[MAX_JOBS];
};
// tile_w = columns of a record's tile (256: hidden_wgrad_kernel, 512: wide_wgrad_kernel); record = 256 x tile_w + 256 bias sums
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, ReduceJobs jobs, int splits, int tile_w,
                    const float* __restrict__ scale) {
  const int REC = 256 * tile_w + 256;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. REC-1
  if (e >= REC || !jobs.job[blockIdx.y].is_head) return;
  const float inv = 1.f / *scale;
  const bool bias = e >= 256 * tile_w;
  const int i = bias ? (e - 256 * tile_w) : (e / tile_w), j = e % tile_w;
#pragma unroll 1
  for (int jb = blockIdx.y; jb >= 0; jb = jobs.job[jb].next) {
    const ReduceJob& J = jobs.job[jb];
    if (bias ? (J.db == nullptr) : (j >= J.n_valid)) continue;
    const float* src = part + (size_t)jb * splits * REC + e;
    float a0 = 0.f, a1 = 0.f;
    int s = 0;
    for (; s + 2 <= splits; s += 2) {
      a0 += src[(size_t)s * REC];
      a1 += src[(size_t)(s + 1) * REC];
    }
    if (s < splits) a0 += src[(size_t)s * REC];
    const float t = (a0 + a1) * inv;
    if (bias) J.db[J.m0 + i] += t;
    else J.dw[(size_t)(J.m0 + i) * J.ldw + J.n0 + j] += t;
  }
}

// part[block][c][k] = sum over the block's rows of s[r][c] * x[r][k]   s: fp32 [Q,nc]; nc <= 4; x = fp16 [Q,WIDTH]
// row-major activations (the ReLU MLPs' last hidden layer).  HBM-bound: 512 threads = row lanes x column groups of 8
// (one 16-byte load each), 4 rows in flight per thread.  NC = number of output rows (1 alpha, 2 uv): with the
// accumulators sized for it (instead of always 4) a thread fits 64 registers and TWO blocks share an SM: 64 KB of loads
// in flight per SM instead of 32.
template <int WIDTH, int NC>
__global__ void __launch_bounds__(512, NC <= 2 ? 2 : 1)
skinny_wgrad_kernel(const __half* __restrict__ x, const float* __restrict__ s, int Q,
                    int rows_per_block, float* __restrict__ part) {
  constexpr int nc = NC;
  constexpr int UNR = 4;                     // rows in flight per thread
  constexpr int HID = WIDTH;                 // row length of x
  constexpr int CG = WIDTH / 8;              // column groups of 8 (one 16-byte load each)
  constexpr int RL = 512 / CG;               // row lanes
  extern __shared__ float s_acc_raw[];   // [RL][NC][WIDTH] floats (dynamic: above the 48 KB static limit for NC > 2)
  float (*s_acc)[NC][HID] = reinterpret_cast<float (*)[NC][HID]>(s_acc_raw);
  const int r0 = blockIdx.x * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > Q) r1 = Q;
  // consecutive threads = consecutive column groups of one row
  const int cg = threadIdx.x % CG;
  const int rl = threadIdx.x / CG;
  float acc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  for (int rb = r0 + rl; rb < r1; rb += UNR * RL) {
    uint4 v[UNR];
    float sv[UNR][NC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int r = rb + RL * u;
      if (r < r1) {
        const size_t off = (size_t)r * HID + cg * 8;
        v[u] = *reinterpret_cast<const uint4*>(x + off);
#pragma unroll
        for (int c = 0; c < NC; ++c) sv[u][c] = __ldg(s + (size_t)r * nc + c);
      } else {
        v[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int c = 0; c < NC; ++c) sv[u][c] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[jj]));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[c][2 * jj] = fmaf(sv[u][c], f.x, acc[c][2 * jj]);
          acc[c][2 * jj + 1] = fmaf(sv[u][c], f.y, acc[c][2 * jj + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_acc[rl][c][cg * 8 + j] = acc[c][j];
  __syncthreads();
  for (int e = threadIdx.x; e < nc * HID; e += 512) {
    const int c = e / HID, k = e % HID;
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < RL; ++l) t += s_acc[l][c][k];
    part[(size_t)blockIdx.x * nchost::WS_SKINNY_REC + e] = t;   // reduced over blocks, in block order, by reduce.cu
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Scene-network head: dW4[c,k] = sum_r dy[r,c] * sin(phase3[r,k]) as a STREAMING kernel.  The tile-blocked phase layout
// [tile][column group of 8][row 128][8] makes the 8 column groups (= 64 columns, an "octet") of one 128-row tile a
// contiguous 16 KB run: a block owns one octet of a range of tiles and pulls those runs (plus the tile's 1.5-2 KB of dy)
// through a ring of cp.async.bulk stages, so the bytes in flight per SM are the ring (4 x 18 KB x 2 blocks), not what
// the register file of the math threads can hold -- the per-thread-load version of this kernel sat at 2.8 TB/s because
// every iteration exposed a full HBM latency.  512 threads: thread = (column group of the octet, row lane of 64), two
// 16-byte items per stage and a fixed set of nc x 8 accumulators for the whole range.  The block's sums (butterfly over
// the warp, then the two warps of a column group) go to record `range` of the workspace; reduce.cu adds the records in
// range order.  HBM-bound: one LOP + FFMA + MUFU + nc FFMA per element.
// ------------------------------------------------------------------------------------------------------------------
constexpr int HS_STAGES = 4;
constexpr uint32_t HS_PH_BYTES = 8 * 128 * 16;            // one octet of one tile
constexpr uint32_t HS_DY_BYTES = 128 * 4 * 4;             // dy of one tile (nc <= 4)
constexpr uint32_t HS_STAGE_BYTES = HS_PH_BYTES + HS_DY_BYTES;
constexpr uint32_t HS_SMEM_BYTES = HS_STAGES * HS_STAGE_BYTES + 64 + 128;   // + barriers + alignment slack

template <int NC>
__global__ void __launch_bounds__(512, 2)
head_wgrad_stream_kernel(const uint16_t* __restrict__ phase, const float* __restrict__ dy, int Q, int ntiles,
                         int tiles_per_range, float* __restrict__ part) {
  extern __shared__ __align__(128) uint8_t hs_raw[];
  const uint32_t raw = smem_u32(hs_raw);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* sm = hs_raw + (base - raw);
  const uint32_t bar0 = base + HS_STAGES * HS_STAGE_BYTES;
  const int tid = threadIdx.x, lane = tid & 31;
  const int octet = blockIdx.x & 7, range = blockIdx.x >> 3;
  const int t0 = range * tiles_per_range;
  int t1 = t0 + tiles_per_range;
  if (t1 > ntiles) t1 = ntiles;
  const int n = t1 - t0;
  const int cgl = tid >> 6, rl = tid & 63;

  if (tid == 0) {
    for (int s = 0; s < HS_STAGES; ++s) mbar_init(bar0 + 8u * s, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // a tile whose rows run past Q has no whole dy run to copy: its dy is written by the threads (zeros past Q)
  auto issue = [&](int i) {
    const int tile = t0 + i;
    const uint32_t st = base + (uint32_t)(i % HS_STAGES) * HS_STAGE_BYTES, bar = bar0 + 8u * (uint32_t)(i % HS_STAGES);
    const bool whole = (long long)(tile + 1) * 128 <= Q;
    mbar_arrive_expect_tx(bar, HS_PH_BYTES + (whole ? (uint32_t)(128 * NC * 4) : 0u));
    bulk_load_1d(st, phase + ((size_t)tile * 64 + (size_t)octet * 8) * (128 * 8), HS_PH_BYTES, bar);
    if (whole) bulk_load_1d(st + HS_PH_BYTES, dy + (size_t)tile * 128 * NC, 128 * NC * 4, bar);
  };
  if (tid == 0)
    for (int i = 0; i < HS_STAGES && i < n; ++i) issue(i);

  float acc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;

  for (int i = 0; i < n; ++i) {
    const int s = i % HS_STAGES;
    const uint8_t* st = sm + (size_t)s * HS_STAGE_BYTES;
    float* s_dy = reinterpret_cast<float*>(const_cast<uint8_t*>(st) + HS_PH_BYTES);
    const int tile = t0 + i;
    if ((long long)(tile + 1) * 128 > Q) {
      for (int e = tid; e < 128 * NC; e += 512) {
        const long long r = (long long)tile * 128 + e / NC;
        s_dy[e] = (r < Q) ? dy[r * NC + e % NC] : 0.f;
      }
      __syncthreads();
    }
    mbar_wait(bar0 + 8u * s, (uint32_t)(i / HS_STAGES) & 1u);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = rl + 64 * h;
      const uint4 v = *reinterpret_cast<const uint4*>(st + ((size_t)cgl * 128 + row) * 16);
      float d[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) d[c] = s_dy[row * NC + c];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        // sin of a 16-bit phase (revolutions * 65536): exact int -> float through the mantissa, one FFMA, one MUFU
        constexpr float K = 6.283185307179586f / 65536.0f;
        const float f0 = __sinf(fmaf(__uint_as_float(0x4B000000u | (w[jj] & 0xFFFFu)), K, -8388608.0f * K));
        const float f1 = __sinf(fmaf(__uint_as_float(0x4B000000u | (w[jj] >> 16)), K, -8388608.0f * K));
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          acc[c][2 * jj] = fmaf(d[c], f0, acc[c][2 * jj]);
          acc[c][2 * jj + 1] = fmaf(d[c], f1, acc[c][2 * jj + 1]);
        }
      }
    }
    __syncthreads();   // every thread is done with stage s: refill it
    if (tid == 0 && i + HS_STAGES < n) issue(i + HS_STAGES);
  }

  // rows -> one sum per (c, column): butterfly over the warp's 32 row lanes (a fixed tree), then the column group's two
  // warps through shared memory (the ring is idle now)
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = acc[c][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      acc[c][j] = t;
    }
  float* s_red = reinterpret_cast<float*>(sm);   // [16 warps][NC * 8]
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) s_red[(tid >> 5) * (NC * 8) + c * 8 + j] = acc[c][j];
  }
  __syncthreads();
  if (tid < 8 * NC * 8) {
    const int g = tid / (NC * 8), e = tid % (NC * 8), c = e / 8, j = e % 8;
    const float t = s_red[(2 * g) * (NC * 8) + e] + s_red[(2 * g + 1) * (NC * 8) + e];
    part[(size_t)range * nchost::WS_SKINNY_REC + c * HID + octet * 64 + g * 8 + j] = t;
  }
}

}  // namespace

// shared launch: n_jobs pair tiles x Q-slices filling the machine's CTA pairs, then the fixed-order slice reduction
static int launch_wgrad_jobs(WgradJobs& jobs, int n_jobs, int64_t Q, const float* scale_dev, void* ws,
                             int64_t ws_bytes, void* stream) {
  WgradParams p{};
  p.Q = (int)Q;
  p.nkb = (int)((Q + 63) / 64);
  int splits = (nchost::sm_count() / 2) / n_jobs;
  if (splits < 1) splits = 1;
  if (splits > p.nkb) splits = p.nkb;
  p.kb_per_split = (p.nkb + splits - 1) / splits;
  p.splits = splits;
  p.scale = scale_dev;
  p.part = reinterpret_cast<float*>(ws);
  NC_CHECK_WS(ws, ws_bytes, (int64_t)n_jobs * splits * REC * 4);
  int rc = nchost::ensure_dyn_smem((const void*)hidden_wgrad_kernel, (int)SMEM_BYTES);
  if (rc) return rc;
  dim3 grid(2 * splits, n_jobs, 1);
  hidden_wgrad_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(jobs, p);
  NC_CHECK_CUDA(cudaGetLastError());
  ReduceJobs rj;
  for (int i = 0; i < n_jobs; ++i) {
    const WgradJob& j = jobs.job[i];
    rj.job[i] = ReduceJob{j.dw, j.db, j.m0, j.n0, j.ldw, j.n_valid, -1, 1};
  }
  for (int i = 0; i < n_jobs; ++i)        // chain the jobs that write the same destination tile, in job order
    for (int k = i + 1; k < n_jobs; ++k)
      if (rj.job[k].dw == rj.job[i].dw && rj.job[k].m0 == rj.job[i].m0 && rj.job[k].n0 == rj.job[i].n0) {
        rj.job[i].next = k;
        rj.job[k].is_head = 0;
        break;
      }
  wgrad_reduce_kernel<<<dim3((REC + 255) / 256, n_jobs), 256, 0, (cudaStream_t)stream>>>(p.part, rj, splits, 256,
                                                                                         scale_dev);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(2);
  return 0;
}

static int fill_job(WgradJob& j, const void* dz, int dz_cols, const void* act, int act_cols, int64_t Q, float* dw,
                    float* db, int m0, int n0, int ldw, int n_valid) {
  int rc = nchost::make_tmap_16b(&j.dz, dz, (uint64_t)Q, dz_cols, dz_cols, 64, 64);
  if (rc) return rc;
  rc = nchost::make_tmap_16b(&j.act, act, (uint64_t)Q, act_cols, act_cols, 64, 64);
  if (rc) return rc;
  j.dw = dw; j.db = db; j.m0 = m0; j.n0 = n0; j.ldw = ldw; j.n_valid = n_valid;
  return 0;
}

extern "C" int neucam_sine_mlp_wgrad(const void* dz16, const void* act16, int64_t Q, const float* scale_dev,
                                     float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3,
                                     void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(dz16 && act16 && scale_dev && dw1 && db1 && dw2 && db2 && dw3 && db3);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - 128);
  float* dw[3] = {dw1, dw2, dw3};
  float* db[3] = {db1, db2, db3};
  WideJobs jobs;
  ReduceJobs rj;
  int n = 0;
  for (int l = 0; l < 3; ++l) {
    // dz16 = [3,Q,512] holding dz_1..dz_3 ; act16 = [3,Q,512] holding a_1..a_3
    const __half* dz = (const __half*)dz16 + (size_t)l * Q * HID;
    const __half* a = (const __half*)act16 + (size_t)l * Q * HID;
    for (int mt = 0; mt < 2; ++mt) {
      int rc = nchost::make_tmap_16b(&jobs.job[n].dz, dz, (uint64_t)Q, HID, HID, 64, 64);
      if (rc) return rc;
      rc = nchost::make_tmap_16b(&jobs.job[n].act, a, (uint64_t)Q, HID, HID, 64, 64);
      if (rc) return rc;
      jobs.job[n].m0 = mt * 256;
      rj.job[n] = ReduceJob{dw[l], db[l], mt * 256, 0, HID, 512, -1, 1};
      ++n;
    }
  }
  WgradParams p{};
  p.Q = (int)Q;
  p.nkb = (int)((Q + 63) / 64);
  int splits = (nchost::sm_count() / 2) / n;
  if (splits < 1) splits = 1;
  if (splits > p.nkb) splits = p.nkb;
  p.kb_per_split = (p.nkb + splits - 1) / splits;
  p.splits = splits;
  p.scale = scale_dev;
  p.part = reinterpret_cast<float*>(ws);
  NC_CHECK_WS(ws, ws_bytes, (int64_t)n * splits * WREC * 4);
  int rc = nchost::ensure_dyn_smem((const void*)wide_wgrad_kernel, (int)WSMEM_BYTES);
  if (rc) return rc;
  wide_wgrad_kernel<<<dim3(2 * splits, n, 1), NTHREADS, WSMEM_BYTES, (cudaStream_t)stream>>>(jobs, p);
  NC_CHECK_CUDA(cudaGetLastError());
  wgrad_reduce_kernel<<<dim3((WREC + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(p.part, rj, splits, 512, scale_dev);
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(2);
  return 0;
}

// Weight gradients of one 256-wide ReLU MLP (reference modules.py:255-306) from the buffers the chain kernels stored.
//   dz16 / dz_lo16 [H,Q,256], act16 / act_lo16 [H,Q,256] (hi and lo fp16 parts), feat16 [Q,64] fp16.
//   Layer 0: dw_first [256,64] (encoded-input layout) and db[0]; layers l = 1..H-1: dw[l] with row pitch ldw[l]
//   (its first 256 columns are written) and db[l]; a layer with dw_skip[l] != NULL also gets that [256,64] =
//   dz_l^T feat.  With the lo tensors every product is split: (dz_hi + dz_lo)^T h_hi + dz_hi^T h_lo; dz_lo16 / act_lo16
//   may be NULL (single fp16 operands: their per-sample rounding averages out over Q).  All accumulated into.
extern "C" int neucam_relu_mlp_wgrad(const void* dz16, const void* dz_lo16, const void* act16, const void* act_lo16,
                                     const void* feat16, int64_t Q, int H, const float* scale_dev, float* dw_first,
                                     float* const* dw, const int* ldw, float* const* db, float* const* dw_skip,
                                     void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(dz16 && act16 && feat16 && scale_dev && dw_first && dw && ldw && db && dw_skip);
  NC_CHECK_ARG(Q > 0 && Q < (1ll << 31) - 128 && H >= 1 && H <= 8);
  constexpr int W = 256;
  WgradJobs jobs;
  int n = 0;
  auto add = [&](const void* dz, const void* act, int act_cols, float* w, float* b, int pitch, int n_valid) -> int {
    if (n >= MAX_JOBS) return NEUCAM_ERR_BAD_ARG;
    return fill_job(jobs.job[n++], dz, W, act, act_cols, Q, w, b, 0, 0, pitch, n_valid);
  };
  for (int l = 0; l < H; ++l) {
    const __half* dzh = (const __half*)dz16 + (size_t)l * Q * W;
    const __half* dzl = dz_lo16 ? (const __half*)dz_lo16 + (size_t)l * Q * W : nullptr;
    int rc = 0;
    if (l == 0) {
      NC_CHECK_ARG(db[0]);
      rc = add(dzh, feat16, 64, dw_first, db[0], 64, 64);
      if (!rc && dzl) rc = add(dzl, feat16, 64, dw_first, db[0], 64, 64);
    } else {
      NC_CHECK_ARG(dw[l] && db[l] && ldw[l] >= W);
      const __half* ah = (const __half*)act16 + (size_t)(l - 1) * Q * W;
      const __half* al = act_lo16 ? (const __half*)act_lo16 + (size_t)(l - 1) * Q * W : nullptr;
      rc = add(dzh, ah, W, dw[l], db[l], ldw[l], 256);
      if (!rc && dzl) rc = add(dzl, ah, W, dw[l], db[l], ldw[l], 256);
      if (!rc && al) rc = add(dzh, al, W, dw[l], nullptr, ldw[l], 256);
      if (!rc && dw_skip[l]) {
        rc = add(dzh, feat16, 64, dw_skip[l], nullptr, 64, 64);
        if (!rc && dzl) rc = add(dzl, feat16, 64, dw_skip[l], nullptr, 64, 64);
      }
    }
    if (rc) return rc;
  }
  return launch_wgrad_jobs(jobs, n, Q, scale_dev, ws, ws_bytes, stream);
}

template <int WIDTH, int NC>
static int launch_skinny_nc(const void* x16, const float* s, int64_t Q, float* out, void* ws, int64_t ws_bytes,
                            void* stream) {
  const int blocks = nchost::sm_count() * 2;
  int rows = (int)((Q + blocks - 1) / blocks);
  if (rows < 32) rows = 32;
  const int grid = (int)((Q + rows - 1) / rows);
  NC_CHECK_WS(ws, ws_bytes, (int64_t)grid * nchost::WS_SKINNY_REC * 4);
  constexpr int SMEM = (512 / (WIDTH / 8)) * NC * WIDTH * 4;
  int rc = nchost::ensure_dyn_smem((const void*)skinny_wgrad_kernel<WIDTH, NC>, SMEM);
  if (rc) return rc;
  skinny_wgrad_kernel<WIDTH, NC><<<grid, 512, SMEM, (cudaStream_t)stream>>>((const __half*)x16, s, (int)Q, rows,
                                                                           reinterpret_cast<float*>(ws));
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  segs.nseg = 1;
  segs.seg[0] = {out, 0, NC * WIDTH};
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), grid, nchost::WS_SKINNY_REC, segs, nullptr,
                                        (cudaStream_t)stream);
}
template <int WIDTH>
static int launch_skinny(const void* x16, const float* s, int nc, int64_t Q, float* out, void* ws, int64_t ws_bytes,
                         void* stream) {
  switch (nc) {
    case 1: return launch_skinny_nc<WIDTH, 1>(x16, s, Q, out, ws, ws_bytes, stream);
    case 2: return launch_skinny_nc<WIDTH, 2>(x16, s, Q, out, ws, ws_bytes, stream);
    case 3: return launch_skinny_nc<WIDTH, 3>(x16, s, Q, out, ws, ws_bytes, stream);
    default: return launch_skinny_nc<WIDTH, 4>(x16, s, Q, out, ws, ws_bytes, stream);
  }
}

extern "C" int neucam_sine_mlp_head_wgrad(const void* phase3_16, const float* dy, int out_dim, int64_t Q,
                                          float* dw4, void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(phase3_16 && dy && dw4);
  NC_CHECK_ARG((out_dim == 3 || out_dim == 4) && Q > 0);
  NC_CHECK_ARG(((uintptr_t)phase3_16 & 15) == 0 && ((uintptr_t)dy & 15) == 0);   // cp.async.bulk sources
  const int ntiles = (int)((Q + 127) / 128);
  int ranges = nchost::sm_count() * 2 / 8;     // 8 octet blocks per range, two blocks per SM
  if (ranges < 1) ranges = 1;
  if (ranges > ntiles) ranges = ntiles;
  const int tiles_per_range = (ntiles + ranges - 1) / ranges;
  ranges = (ntiles + tiles_per_range - 1) / tiles_per_range;
  NC_CHECK_WS(ws, ws_bytes, (int64_t)ranges * nchost::WS_SKINNY_REC * 4);
  const void* fn = out_dim == 3 ? (const void*)head_wgrad_stream_kernel<3> : (const void*)head_wgrad_stream_kernel<4>;
  int rc = nchost::ensure_dyn_smem(fn, (int)HS_SMEM_BYTES);
  if (rc) return rc;
  if (out_dim == 3)
    head_wgrad_stream_kernel<3><<<ranges * 8, 512, HS_SMEM_BYTES, (cudaStream_t)stream>>>(
        (const uint16_t*)phase3_16, dy, (int)Q, ntiles, tiles_per_range, reinterpret_cast<float*>(ws));
  else
    head_wgrad_stream_kernel<4><<<ranges * 8, 512, HS_SMEM_BYTES, (cudaStream_t)stream>>>(
        (const uint16_t*)phase3_16, dy, (int)Q, ntiles, tiles_per_range, reinterpret_cast<float*>(ws));
  NC_CHECK_CUDA(cudaGetLastError());
  NC_COUNT_LAUNCH(1);
  nchost::ReduceSegs segs{};
  segs.nseg = 1;
  segs.seg[0] = {dw4, 0, out_dim * HID};
  return nchost::launch_reduce_partials(reinterpret_cast<const float*>(ws), ranges, nchost::WS_SKINNY_REC, segs, nullptr,
                                        (cudaStream_t)stream);
}

// dw_last[c,k] += sum_r dpre[r,c] * (h[r,k] + h_lo[r,k]) for the last layer of a 256-wide ReLU MLP
// (h, h_lo fp16 [Q,256]; c < 4).
extern "C" int neucam_relu_mlp_head_wgrad(const void* h16, const void* h_lo16, const float* dpre, int out_dim,
                                          int64_t Q, float* dw_last, void* ws, int64_t ws_bytes, void* stream) {
  NC_CHECK_ARG(h16 && dpre && dw_last);
  NC_CHECK_ARG(out_dim >= 1 && out_dim <= 4 && Q > 0);
  int rc = launch_skinny<256>(h16, dpre, out_dim, Q, dw_last, ws, ws_bytes, stream);
  if (rc || h_lo16 == nullptr) return rc;
  return launch_skinny<256>(h_lo16, dpre, out_dim, Q, dw_last, ws, ws_bytes, stream);   // stream order: ws is free again
}
