// CTA-pair (cta_group::2) variant of the implicit-GEMM convolution.
//
// Two CTAs of a cluster (the two SMs of a TPC) compute a 256-pixel x BN-cout tile together:
//   * each CTA TMA-loads ITS OWN 128-pixel activation tile and HALF of the weight tile (BN/2 rows),
//   * the leader CTA issues tcgen05.mma.cta_group::2 (M = 256): each SM multiplies its 128 rows with
//     the full BN-row weight tile, whose halves it reads from both CTAs' shared memory,
//   * each CTA's TMEM holds the accumulator of its own 128 rows and each CTA runs its own epilogue.
// Per SM and per 64-channel block this moves 16 KB + BN*64 B through TMA and shared memory instead of
// 16 KB + BN*128 B: the single-CTA kernel is shared-memory-bandwidth bound (TMA writes + UMMA operand reads),
// worst for Cout = 128 layers (microbench: 797 TFLOP/s vs 1 244 for BN = 256, cuBLAS sustained 1 392).
//
// Barrier protocol (every barrier exists at the same offset in both CTAs):
//   full[s]   leader's copy only: 2 arrivals (leader expect_tx + peer remote arrive) + bytes of both CTAs' TMA
//   empty[s]  own copy: 1 arrival from the leader's tcgen05.commit multicast
//   tfull[a]  own copy: 1 arrival from the leader's tcgen05.commit multicast
//   tempty[a] leader's copy only: 16 arrivals (8 epilogue warps x 2 CTAs)
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_sm100_common.cuh"

namespace advs {

template <int BN>
struct ConvCfg2 {
  static constexpr uint32_t b_bytes = (BN / 2) * 128;          // this CTA's half of the weight tile
  static constexpr uint32_t stage_bytes = kABytes + b_bytes;
  static constexpr int stages = (BN == 256) ? 6 : 8;
  static constexpr uint32_t bar_bytes = 512;
  static constexpr uint32_t stat_bytes = 2 * 4 * BN * 8;
  static constexpr uint32_t smem_bytes = stages * stage_bytes + bar_bytes + stat_bytes + 1024;
  static constexpr uint32_t tmem_cols = 2 * BN;
};

template <int BN, bool A0F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
k_conv_sm100_2cta(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  using Cfg = ConvCfg2<BN>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* stat_smem = reinterpret_cast<float2*>(smem + STAGES * Cfg::stage_bytes + Cfg::bar_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_pairs = (a.m_tiles + 1) >> 1;
  const int total_items = m_pairs * a.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 6; ++i) tma_prefetch_desc(&maps.a[i]);
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&maps.b[i]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<Cfg::tmem_cols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();      // everything above overlapped the previous kernel's tail; from here on its results are visible

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = pair; item < total_items; item += num_pairs) {
        const int m_pair = item / a.n_tiles, n_tile = item - m_pair * a.n_tiles;
        const int m_tile = m_pair * 2 + (int)rank;     // may be == m_tiles for an odd tail: fully out of bounds -> zeros
        const int w0 = (m_tile % a.tiles_w) * a.tw;
        const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.th;
        const int n0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.tn;
        const int nb = n_tile * BN + (int)rank * (BN / 2);
        for (int s = 0; s < a.nseg; ++s) {
          const int taps = a.taps[s];
          for (int tap = 0; tap < taps; ++tap) {
            const CUtensorMap* amap;
            int cw = w0, ch = h0;
            if (taps == 9) {
              const int dy = tap / 3, dx = tap - dy * 3;
              if (s == 0 && a.stride == 2) {
                amap = &maps.a[(dy != 1 ? 2 : 0) + (dx != 1 ? 1 : 0)];
                ch += (dy == 0) ? -1 : 0;
                cw += (dx == 0) ? -1 : 0;
              } else {
                amap = &maps.a[s == 0 ? 0 : 3 + s];
                ch += dy - 1;
                cw += dx - 1;
              }
            } else if (taps == 4) {   // 2x2 phase of an upsample-conv on the low-res input
              amap = &maps.a[0];
              ch += (tap >> 1) - 1 + a.up_a;
              cw += (tap & 1) - 1 + a.up_b;
            } else {
              amap = &maps.a[s == 0 ? 0 : 3 + s];
            }
            for (int cb = 0; cb < a.cblks[s]; ++cb) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::stage_bytes;
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (a.a_bytes + Cfg::b_bytes));
              tma_load_4d_2cta(sa, amap, &full[stage], cb * 64, cw, ch, n0);
              tma_load_3d_2cta(sa + kABytes, &maps.b[s], &full[stage], cb * 64, tap, nb);
              if (!leader) mbar_arrive_leader(&full[stage]);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (leader) {   // the whole warp walks the schedule; one elected lane issues (warp-uniform control flow)
      // segment 0 in fp16 or bf16 (compile-time), shortcut segments always bf16: the descriptors stay immediates
      constexpr uint32_t idesc0 = umma_idesc_f16kind(256, BN, A0F16, A0F16);
      constexpr uint32_t idesc1 = umma_idesc_f16kind(256, BN, false, false);
      const int kb_seg0 = a.taps[0] * a.cblks[0];
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = pair; item < total_items; item += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        // compile-time instruction descriptors (see conv_sm100_halo.cu): K blocks of segment 0, then the shortcut segments
        auto k_blocks = [&](const int kb_lo, const int kb_hi, auto idesc_c) {
          constexpr uint32_t idesc = decltype(idesc_c)::value;
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::stage_bytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty[stage], 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        };
        if constexpr (A0F16) {
          k_blocks(0, kb_seg0, std::integral_constant<uint32_t, idesc0>{});
          k_blocks(kb_seg0, a.total_kb, std::integral_constant<uint32_t, idesc1>{});
        } else {
          k_blocks(0, a.total_kb, std::integral_constant<uint32_t, idesc0>{});
        }
        if (elect_one()) umma_commit_2cta(&tfull[acc], 3);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..9, both CTAs, each on its own 128 rows) =================
    const int q = warp & 3;
    const int col_half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int rows_valid = a.tw * a.th * a.tn;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = pair; item < total_items; item += num_pairs) {
      const int m_pair = item / a.n_tiles, n_tile = item - m_pair * a.n_tiles;
      const int m_tile = m_pair * 2 + (int)rank;
      const int w0 = (m_tile % a.tiles_w) * a.tw;
      const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.th;
      const int n0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.tn;
      const int dn = row / (a.th * a.tw);
      const int rem = row - dn * (a.th * a.tw);
      const int dh = rem / a.tw, dw = rem - dh * a.tw;
      const int b = n0 + dn;
      const bool valid = row < rows_valid && b < a.B && m_tile < a.m_tiles;
      const int t = a.up ? (2 * (h0 + dh) + a.up_a) * (2 * a.W) + 2 * (w0 + dw) + a.up_b : (h0 + dh) * a.W + (w0 + dw);
      const size_t m = (size_t)b * a.epi.HW + t;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool want_stats = a.stats != nullptr;
      const bool wide = a.epi.y_lo != nullptr;
#pragma unroll 1
      for (int chunk = col_half * (BN / 64); chunk < (col_half + 1) * (BN / 64); ++chunk) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + chunk * 32, r);
        tmem_wait_ld();
        const int n = n_tile * BN + chunk * 32;
        if (n >= a.epi.cout_valid) break;
        float v[32];
        if (valid) {
          epilogue_compute32(a.epi, r, v, m, b, n);
          if (wide) epilogue_write32<true>(a.epi, v, m, b, t, n);
          else epilogue_write32<false>(a.epi, v, m, b, t, n);
        }
        if (want_stats) {
          // GroupNorm statistics of the tensor just written (taken before the bf16 rounding: the rounding
          // error is zero-mean and ~1e-6 of the variance)
          stats_stage(stat_smem, a, BN, acc, q, chunk, lane, warp_stats32(v, valid, lane, a.stats_gran));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
      if (want_stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (m_tile < a.m_tiles) stats_flush(stat_smem, a, BN, acc, n_tile, m_tile, (warp - 2) * 32 + lane);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) tmem_dealloc_2cta<Cfg::tmem_cols>(tmem_base);
}

uint32_t conv_2cta_smem_bytes(int bn) { return bn == 256 ? ConvCfg2<256>::smem_bytes : ConvCfg2<128>::smem_bytes; }

int launch_conv_2cta(const ConvPlan* plan, cudaStream_t st) {
  if (first_use_on_device(kOnceConv2Cta)) {
    cudaError_t e1 = cudaFuncSetAttribute(k_conv_sm100_2cta<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfg2<128>::smem_bytes);
    cudaError_t e2 = cudaFuncSetAttribute(k_conv_sm100_2cta<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfg2<256>::smem_bytes);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_conv_sm100_2cta<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfg2<128>::smem_bytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_conv_sm100_2cta<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfg2<256>::smem_bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_error("conv_sm100_launch(2cta): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      forget_first_use(kOnceConv2Cta);
      return ADVS_ERR_CUDA;
    }
  }
  const bool f16 = plan->args.operand_f16 != 0;
  if (plan->bn == 256) {
    if (f16) launch_pdl(k_conv_sm100_2cta<256, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100_2cta<256, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  } else {
    if (f16) launch_pdl(k_conv_sm100_2cta<128, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100_2cta<128, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  }
  ADVS_CHECK_LAUNCH("conv_sm100_launch(2cta)");
  return ADVS_OK;
}

}  // namespace advs
