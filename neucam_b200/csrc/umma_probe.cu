// umma_probe.cu -- single-tile tcgen05 GEMM used as the unit test of the building blocks every
// tensor-core kernel in this library relies on: TMA SWIZZLE_128B loads, K-major and MN-major
// shared-memory matrix descriptors, thread-written swizzled operand tiles, fp16/bf16 operand
// formats, tcgen05.mma issue/commit and the 32x32b TMEM read-back.
//
//   D[128, N] (fp32) = A[128, K] * B[N, K]^T      N % 16 == 0, N <= 256, K % 64 == 0, K <= 256
//
// flags: bit0 A is MN-major (global [K,128]); bit1 B is MN-major (global [K,N], N % 64 == 0);
//        bit2 A is bf16; bit3 B is bf16; bit4 A tile is written by threads (K-major only);
//        bit5 A operand lives in TENSOR MEMORY (each thread tcgen05.st's its own row; K-major only).
// Built only into the debug library (-DNEUCAM_DEBUG -> libneucam_b200_debug.so; declared in
// include/neucam_b200_debug.h): building-block self tests and microbenchmarks, not part of the product ABI.
#ifdef NEUCAM_DEBUG
#include "host_util.h"
#include "tc05.cuh"

using namespace tc05;

namespace {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const uint16_t* __restrict__ a_gmem, float* __restrict__ d_gmem, int N, int K,
                  int flags) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_slot;

  // 1024-byte aligned carve-up (SWIZZLE_128B atoms repeat every 1024 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = smem_base;
  const uint32_t a_bytes = 128u * (uint32_t)K * 2u;
  const uint32_t b_smem = a_smem + a_bytes;
  const uint32_t b_bytes = (uint32_t)N * (uint32_t)K * 2u;
  uint8_t* a_ptr = smem_raw + (a_smem - smem_u32(smem_raw));

  const bool a_mn = flags & 1, b_mn = flags & 2;
  const bool a_bf16 = flags & 4, b_bf16 = flags & 8;
  const bool a_manual = flags & 16;
  const bool a_tmem = flags & 32;

  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_d = tmem_base_slot;
  const uint32_t tmem_a = tmem_d + 256;   // A operand region (K / 2 columns)

  if (a_tmem) {
    // thread = row = TMEM lane: K fp16 values = K/2 packed 32-bit columns, 16 columns per store
    const uint32_t row = tid;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = *reinterpret_cast<const uint32_t*>(a_gmem + (size_t)row * K + 2 * (c0 + j));
      tmem_st_32x32b_x16(tmem_a + ((warp * 32u) << 16) + (uint32_t)c0, w);
    }
    tmem_st_wait();
    tc_fence_before_sync();
  }

  if (a_manual) {
    // every thread copies its own row of A (K-major): 16-byte chunks, XOR-swizzled inside 128 B
    const uint32_t row = tid;
    for (int kb = 0; kb < K / 64; ++kb) {
      for (uint32_t c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(a_gmem + (size_t)row * K + kb * 64 + c * 8);
        *reinterpret_cast<uint4*>(a_ptr + kb * (128 * 128) + sw128_offset(row, c)) = v;
      }
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (tid == 0) {
    const uint32_t tx = b_bytes + ((a_manual || a_tmem) ? 0u : a_bytes);
    mbar_arrive_expect_tx(bar_tma, tx);
    if (!a_manual && !a_tmem) {
      if (a_mn) {
        for (int a = 0; a < 2; ++a) tma_load_2d(a_smem + a * (K * 128), &tmap_a, bar_tma, a * 64, 0);
      } else {
        for (int kb = 0; kb < K / 64; ++kb)
          tma_load_2d(a_smem + kb * (128 * 128), &tmap_a, bar_tma, kb * 64, 0);
      }
    }
    if (b_mn) {
      for (int a = 0; a < N / 64; ++a) tma_load_2d(b_smem + a * (K * 128), &tmap_b, bar_tma, a * 64, 0);
    } else {
      for (int kb = 0; kb < K / 64; ++kb)
        tma_load_2d(b_smem + kb * (N * 128), &tmap_b, bar_tma, kb * 64, 0);
    }
    mbar_wait(bar_tma, 0);
    tc_fence_after_sync();

    const uint32_t idesc = make_idesc_f16(128, (uint32_t)N, a_bf16 ? FMT_BF16 : FMT_F16,
                                          b_bf16 ? FMT_BF16 : FMT_F16, a_mn ? MAJOR_MN : MAJOR_K,
                                          b_mn ? MAJOR_MN : MAJOR_K);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint64_t ad, bd;
      if (a_mn) ad = make_sdesc_sw128(a_smem + k16 * 2048, (uint32_t)K * 128u, 1024);
      else      ad = make_sdesc_sw128(a_smem + (k16 >> 2) * (128 * 128) + (k16 & 3) * 32, 0, 1024);
      if (b_mn) bd = make_sdesc_sw128(b_smem + k16 * 2048, (uint32_t)K * 128u, 1024);
      else      bd = make_sdesc_sw128(b_smem + (k16 >> 2) * (N * 128) + (k16 & 3) * 32, 0, 1024);
      if (a_tmem) umma_f16_ts(tmem_d, tmem_a + (uint32_t)k16 * 8u, bd, idesc, k16 > 0 ? 1u : 0u);
      else        umma_f16_ss(tmem_d, ad, bd, idesc, k16 > 0 ? 1u : 0u);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();

  mbar_wait(bar_mma, 0);
  tc_fence_after_sync();

  // epilogue: thread = row; 32 columns per tcgen05.ld
  const uint32_t row = tid;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem_d + ((warp * 32u) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) d_gmem[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 512);
}

}  // namespace

extern "C" int neucam_umma_probe(const void* a, const void* b, float* d, int N, int K, int flags,
                                 void* stream) {
  NC_CHECK_ARG(a != nullptr && b != nullptr && d != nullptr);
  NC_CHECK_ARG(N >= 16 && N <= 256 && N % 16 == 0);
  NC_CHECK_ARG(K >= 64 && K <= 256 && K % 64 == 0);
  const bool a_mn = flags & 1, b_mn = flags & 2, a_manual = flags & 16;
  NC_CHECK_ARG(!(a_mn && (a_manual || (flags & 32))));
  NC_CHECK_ARG(!b_mn || N % 64 == 0);

  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = nchost::make_tmap_16b(&ta, a, K, 128, 128, K, 64);
  else      rc = nchost::make_tmap_16b(&ta, a, 128, K, K, 128, 64);
  if (rc) return rc;
  if (b_mn) rc = nchost::make_tmap_16b(&tb, b, K, N, N, K, 64);
  else      rc = nchost::make_tmap_16b(&tb, b, N, K, K, N, 64);
  if (rc) return rc;

  const size_t smem = 1024 + (size_t)128 * K * 2 + (size_t)N * K * 2;
  NC_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ta, tb, (const uint16_t*)a, d, N, K, flags);
  NC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark: one CTA pair, the leader issues `iters` back-to-back
// tcgen05.mma.cta_group::2 M256 x N x K16 (or cta_group::1 M128 x N) on resident smem tiles and reports
// clock64 cycles from first issue to completion.  Tells what the tensor pipe can sustain for a shape.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_rate_kernel(int N, int iters, int two_cta, int kblocks, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar_dummy[2];
  const int overhead = (kblocks >> 8) & 3;   // bits 8,9: per-iteration control overhead to emulate (1 = commit, 2 = wait+fence)
  const bool a_in_tmem = (kblocks >> 10) & 1;   // bit 10: A operand from tensor memory (cta_group::1 only)
  kblocks &= 0xff;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (192u * 1024u) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3C003C00u;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar_dummy[0]), 1u << 20);   // commit target that never completes a phase
    mbar_init(smem_u32(&bar_dummy[1]), 1);          // fresh barrier: waiting on parity 1 returns at once
    mbar_fence_init();
  }
  if (warp == 0) {
    if (two_cta) { tmem_alloc_2cta(smem_u32(&slot), 512); tmem_relinquish_2cta(); }
    else         { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  const bool leader = cluster_ctarank() == 0;
  if (tid == 0 && (leader || !two_cta)) {
    const uint32_t idesc = make_idesc_f16(two_cta ? 256 : 128, (uint32_t)N, FMT_F16, FMT_F16, MAJOR_K, MAJOR_K);
    const uint32_t a0 = base, b0 = base + 128 * 1024;   // A: 8 k-blocks of 16 KB; B: up to 4 stages of 16 KB
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (overhead & 2) {
        mbar_wait(smem_u32(&bar_dummy[1]), 1);
        tc_fence_after_sync();
      }
      const uint32_t a_addr = a0 + (it % kblocks) * 16384;
      const uint32_t b_addr = b0 + (it & 1) * 32768;   // a full N=256 tile is 32 KB (1-CTA) / 16 KB (2-CTA half)
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        if (two_cta) umma_f16_ss_2cta(tm, make_sdesc_sw128(a_addr + k16 * 32, 0, 1024), make_sdesc_sw128(b_addr + k16 * 32, 0, 1024), idesc, 1u);
        else if (a_in_tmem) umma_f16_ts(tm, tm + 256 + (it % kblocks) * 32 + k16 * 8, make_sdesc_sw128(b_addr + k16 * 32, 0, 1024), idesc, 1u);
        else         umma_f16_ss(tm, make_sdesc_sw128(a_addr + k16 * 32, 0, 1024), make_sdesc_sw128(b_addr + k16 * 32, 0, 1024), idesc, 1u);
      }
      if (overhead & 1) {
        if (two_cta) umma_commit_2cta(smem_u32(&bar_dummy[0]), 0x3);
        else         umma_commit(smem_u32(&bar_dummy[0]));
      }
    }
    const long long t1 = clock64();
    if (two_cta) umma_commit_2cta(smem_u32(&bar), 0x1);
    else         umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2 + 0] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    if (two_cta) tmem_dealloc_2cta(tm, 512);
    else         tmem_dealloc(tm, 512);
  }
}
}  // namespace

extern "C" int neucam_umma_rate(int N, int iters, int two_cta, int kblocks, int grid, long long* out_dev,
                                void* stream) {
  NC_CHECK_ARG(N >= 16 && N <= 256 && iters > 0 && (kblocks & 0xff) >= 1 && (kblocks & 0xff) <= 8 && grid >= 2 && grid % 2 == 0);
  const size_t smem = 193 * 1024 + 1024;
  NC_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(N, iters, two_cta, kblocks, out_dev);
  NC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // NEUCAM_DEBUG
