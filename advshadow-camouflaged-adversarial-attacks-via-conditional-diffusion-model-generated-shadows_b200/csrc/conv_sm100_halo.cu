// CTA-pair implicit-GEMM convolution with HALO REUSE, for rows of >= 128 pixels (W % 128 == 0, stride 1).
//
// conv_sm100_2cta.cu re-loads the activation tile once per tap (9x for a 3x3 conv) from L2.  Here each
// (segment, 64-channel block) loads ONE halo box {64 ch, 128 + kx - 1 pixels, ky rows} and every tap reads
// it through a UMMA descriptor shifted by whole 128-byte rows -- legal because the 128-byte swizzle is a
// function of the absolute shared-memory address (verified on hardware by advs_selftest_umma_row_shift).
// L2 -> SM traffic per 64-channel block of a 3x3 conv drops from 9 x 16 KB to 49 KB; the weight tiles
// stream through their own ring.  Microbench (128->128 3x3 @256^2, batch 64) motivating it: the layer was
// L2-bandwidth bound at 795 TFLOP/s while BN = 256 layers reach 1 400.
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_sm100_common.cuh"

namespace advs {

#ifdef ADVS_ATTN_TRACE   // debug build (make trace): clock64 stamps of CTA 0's first epilogue warp and of the MMA warp
__device__ long long g_conv_trace[64 * 16];
#define CTR(slot) do { if (blockIdx.x == 0 && lane == 0 && tix < 64) g_conv_trace[tix * 16 + (slot)] = clock64(); } while (0)
#else
#define CTR(slot) do { } while (0)
#endif

template <int BN>
struct ConvCfgH {
  static constexpr uint32_t b_bytes = (BN / 2) * 128;          // this CTA's half of one tap's weight tile
  static constexpr uint32_t a_buf_bytes = 49 * 1024;           // >= 3 rows x 130 pixels x 128 B, 1 KB aligned
  static constexpr int a_bufs = 2;                             // (3 buffers + a 5-deep weight ring measured slower)
  // BN = 128: an MMA per tap is only 4 x 64 tensor cycles, about what one mbarrier wait + commit costs the
  // issuing warp, so three taps (one kernel row) share a ring stage and a barrier round trip
  static constexpr int tps = (BN == 256) ? 1 : 3;              // taps per weight-ring stage
  static constexpr uint32_t stage_bytes = tps * b_bytes;
  static constexpr int stages = (BN == 256) ? 6 : 4;           // weight-tile ring (BN = 128: 3 / 4 / 5 stages measured
                                                               // 6 200 / 5 400 / 5 400 cycles per tile)
  static constexpr uint32_t off_b = a_bufs * a_buf_bytes;
  static constexpr uint32_t off_bar = off_b + stages * stage_bytes;
  static constexpr uint32_t bar_bytes = 256;                   // 2 * stages + 2 * a_bufs + 4 mbarriers + the TMEM slot
  static constexpr uint32_t stat_bytes = 2 * 4 * BN * 8;
  // + slack for aligning the dynamic shared-memory base up to 1 KB (the kernel traps if that is ever not enough)
  static constexpr uint32_t used_bytes = off_bar + bar_bytes + stat_bytes;
  static constexpr uint32_t smem_bytes = used_bytes + 1024;
  static_assert(smem_bytes <= 232448, "halo kernel: shared memory over the 227 KB limit");
  static_assert((2 * stages + 2 * a_bufs + 4) * 8 + 8 <= bar_bytes, "halo kernel: barrier area too small");
  static constexpr uint32_t tmem_cols = 2 * BN;
};

template <int BN, bool A0F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
k_conv_sm100_2cta_halo(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  using Cfg = ConvCfgH<BN>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if ((uint32_t)(smem - smem_raw) + Cfg::used_bytes > Cfg::smem_bytes) __trap();   // base less aligned than assumed
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::off_bar);   // weight ring
  uint64_t* empty = full + STAGES;
  uint64_t* a_full = empty + STAGES;                                   // activation (halo) buffers
  uint64_t* a_empty = a_full + Cfg::a_bufs;
  uint64_t* tfull = a_empty + Cfg::a_bufs;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* stat_smem = reinterpret_cast<float2*>(smem + Cfg::off_bar + Cfg::bar_bytes);

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
    for (int s = 0; s < Cfg::a_bufs; ++s) {
      mbar_init(&a_full[s], 2);
      mbar_init(&a_empty[s], 1);
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
      int stage = 0, abuf = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = pair; item < total_items; item += num_pairs) {
        const int m_pair = item / a.n_tiles, n_tile = item - m_pair * a.n_tiles;
        const int m_tile = m_pair * 2 + (int)rank;     // may be == m_tiles for an odd tail: fully out of bounds -> zeros
        const int w0 = (m_tile % a.tiles_w) * a.tw;
        const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.th;
        const int n0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.tn;
        const int nb = n_tile * BN + (int)rank * (BN / 2);
        for (int s = 0; s < a.nseg; ++s) {
          const int taps = a.taps[s];
          // one box per (segment, channel block) serves every tap: origin = top-left of the halo
          int cw = w0, ch = h0;
          if (taps == 9) { cw -= 1; ch -= 1; }
          else if (taps == 4) { cw += a.up_b - 1; ch += a.up_a - 1; }
          const CUtensorMap* amap = &maps.a[s == 0 ? 0 : 3 + s];
          for (int cb = 0; cb < a.cblks[s]; ++cb) {
            mbar_wait(&a_empty[abuf], aphase ^ 1);
            if (leader) mbar_arrive_expect_tx(&a_full[abuf], 2 * a.a_bytes_seg[s]);
            tma_load_4d_2cta(smem + abuf * Cfg::a_buf_bytes, amap, &a_full[abuf], cb * 64, cw, ch, n0);
            if (!leader) mbar_arrive_leader(&a_full[abuf]);
            if (++abuf == Cfg::a_bufs) { abuf = 0; aphase ^= 1; }
            for (int tap0 = 0; tap0 < taps; tap0 += Cfg::tps) {
              const int ng = min(Cfg::tps, taps - tap0);
              mbar_wait(&empty[stage], phase ^ 1);
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * ng * Cfg::b_bytes);
              for (int g = 0; g < ng; ++g)
                tma_load_3d_2cta(smem + Cfg::off_b + stage * Cfg::stage_bytes + g * Cfg::b_bytes, &maps.b[s], &full[stage],
                                 cb * 64, tap0 + g, nb);
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
      int stage = 0, abuf = 0;
      uint32_t phase = 0, aphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tix = 0;
      for (int item = pair; item < total_items; item += num_pairs, ++tix) {
        CTR(12);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        CTR(13);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t first = 1;
        // the instruction descriptor must stay an immediate / uniform register: a run-time select puts it into a
        // vector register and costs an extra R2UR in front of every MMA group (this kernel is issue-bound at BN = 128)
        auto segment = [&](const int s, auto idesc_c) {
          constexpr uint32_t idesc = decltype(idesc_c)::value;
          const int taps = a.taps[s];
          for (int cb = 0; cb < a.cblks[s]; ++cb) {
            mbar_wait(&a_full[abuf], aphase);
            const uint32_t a_base = smem_u32(smem + abuf * Cfg::a_buf_bytes);
            for (int tap0 = 0; tap0 < taps; tap0 += Cfg::tps) {
              const int ng = min(Cfg::tps, taps - tap0);
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint32_t b_base = smem_u32(smem + Cfg::off_b + stage * Cfg::stage_bytes);
              if (elect_one()) {
#pragma unroll
                for (int g = 0; g < Cfg::tps; ++g) {
                  if (g < ng) {
                    // the tap is a whole-row shift inside the halo box; SWIZZLE_128B is a function of the absolute
                    // shared-memory address, so a row-shifted descriptor reads what TMA wrote (selftest_sm100.cu)
                    const int tap = tap0 + g;
                    const int row_off = taps == 9 ? (tap / 3) * 130 + tap % 3 : (taps == 4 ? (tap >> 1) * 129 + (tap & 1) : 0);
                    const uint64_t adesc = umma_desc_k_sw128(a_base + (uint32_t)row_off * 128u);
                    const uint64_t bdesc = umma_desc_k_sw128(b_base + (uint32_t)g * Cfg::b_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_bf16_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                     (first && g == 0 && k == 0) ? 0u : 1u);
                  }
                }
                umma_commit_2cta(&empty[stage], 3);
              }
              __syncwarp();
              first = 0;
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit_2cta(&a_empty[abuf], 3);
            __syncwarp();
            if (++abuf == Cfg::a_bufs) { abuf = 0; aphase ^= 1; }
          }
        };
        segment(0, std::integral_constant<uint32_t, idesc0>{});
        for (int s = 1; s < a.nseg; ++s) segment(s, std::integral_constant<uint32_t, idesc1>{});
        if (elect_one()) umma_commit_2cta(&tfull[acc], 3);
        __syncwarp();
        CTR(14);
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
    int tix = warp == 2 ? 0 : 64;
    for (int item = pair; item < total_items; item += num_pairs, ++tix) {
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
      CTR(0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      CTR(1);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool want_stats = a.stats != nullptr;
      const bool wide = a.epi.y_lo != nullptr;
#pragma unroll 1
      for (int chunk = col_half * (BN / 64); chunk < (col_half + 1) * (BN / 64); ++chunk) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + chunk * 32, r);
        tmem_wait_ld();
        if (chunk == 0) CTR(2);
        const int n = n_tile * BN + chunk * 32;
        if (n >= a.epi.cout_valid) break;
        float v[32];
        if (valid) {
          epilogue_compute32(a.epi, r, v, m, b, n);
          if (chunk == 0) CTR(3);
          if (wide) epilogue_write32<true>(a.epi, v, m, b, t, n);
          else epilogue_write32<false>(a.epi, v, m, b, t, n);
          if (chunk == 0) CTR(4);
        }
        if (want_stats) {
          // GroupNorm statistics of the tensor just written (taken before the bf16 rounding: the rounding
          // error is zero-mean and ~1e-6 of the variance)
          stats_stage(stat_smem, a, BN, acc, q, chunk, lane, warp_stats32(v, valid, lane, a.stats_gran));
        }
        if (chunk == 0) CTR(5);
      }
      CTR(6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[acc]);
      CTR(7);
      if (want_stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        CTR(8);
        if (m_tile < a.m_tiles) stats_flush(stat_smem, a, BN, acc, n_tile, m_tile, (warp - 2) * 32 + lane);
      }
      CTR(9);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) tmem_dealloc_2cta<Cfg::tmem_cols>(tmem_base);
}

#ifdef ADVS_ATTN_TRACE
}  // namespace advs
extern "C" int advs_debug_conv_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, advs::g_conv_trace, sizeof(long long) * 64 * 16) == cudaSuccess ? 0 : -2;
}
namespace advs {
#endif

uint32_t conv_2cta_halo_smem_bytes(int bn) { return bn == 256 ? ConvCfgH<256>::smem_bytes : ConvCfgH<128>::smem_bytes; }

int launch_conv_2cta_halo(const ConvPlan* plan, cudaStream_t st) {
  if (first_use_on_device(kOnceConvHalo)) {
    cudaError_t e1 = cudaFuncSetAttribute(k_conv_sm100_2cta_halo<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfgH<128>::smem_bytes);
    cudaError_t e2 = cudaFuncSetAttribute(k_conv_sm100_2cta_halo<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfgH<256>::smem_bytes);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_conv_sm100_2cta_halo<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfgH<128>::smem_bytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_conv_sm100_2cta_halo<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfgH<256>::smem_bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      set_error("conv_sm100_launch(halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      forget_first_use(kOnceConvHalo);
      return ADVS_ERR_CUDA;
    }
  }
  const bool f16 = plan->args.operand_f16 != 0;
  if (plan->bn == 256) {
    if (f16) launch_pdl(k_conv_sm100_2cta_halo<256, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100_2cta_halo<256, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  } else {
    if (f16) launch_pdl(k_conv_sm100_2cta_halo<128, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100_2cta_halo<128, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  }
  ADVS_CHECK_LAUNCH("conv_sm100_launch(halo)");
  return ADVS_OK;
}

}  // namespace advs
