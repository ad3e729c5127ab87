// K1/K2/K3: convolution as a persistent, warp-specialised implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels, BN couts] (fp32, TMEM) += A[128 pixels, 64 ch] (bf16, smem) * W[BN couts, 64 ch]^T
//
// * activations are NHWC bf16; one TMA 4-D box {64 ch, tw, th, tn} per (tap, channel block) lands a
//   128-row K-major SWIZZLE_128B tile -- the 3x3 taps are just shifted box coordinates and the
//   zero padding is TMA's out-of-bounds fill, so no im2col buffer ever exists;
// * stride-2 convolutions read four "parity" views of the input (even/odd rows x even/odd columns);
// * ResidualBlock's 1x1 shortcut conv (and the two halves of a virtual concat) are extra K-segments
//   accumulated into the same TMEM tile;
// * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2-9 = epilogue (TMEM -> registers ->
//   bias / time-embedding / residual -> bf16 -> global, + optional per-channel GroupNorm partial sums).  TMEM holds two accumulator tiles so the
//   epilogue of tile i overlaps the main loop of tile i+1.
//
// Reference ops replaced: nn.Conv2d at dm1:73, 86, 90, 114-115, 134, 148 (+ the adds at dm1:101,103,127).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "sm100.cuh"
#include "conv_sm100_common.cuh"

namespace advs {

template <int BN, bool A0F16>
__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_sm100(const __grid_constant__ ConvMaps maps, const ConvArgs a) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::stage_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* stat_smem = reinterpret_cast<float2*>(smem + STAGES * Cfg::stage_bytes + Cfg::bar_bytes);  // [2][4][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 6; ++i) tma_prefetch_desc(&maps.a[i]);
    for (int i = 0; i < 3; ++i) tma_prefetch_desc(&maps.b[i]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::tmem_cols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();      // everything above overlapped the previous kernel's tail; from here on its results are visible

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int w0 = (m_tile % a.tiles_w) * a.tw;
        const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.th;
        const int n0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.tn;
        const int nb = n_tile * BN;
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
              mbar_arrive_expect_tx(&full[stage], a.a_bytes + Cfg::b_bytes);
              tma_load_4d(sa, amap, &full[stage], cb * 64, cw, ch, n0);
              tma_load_3d(sa + kABytes, &maps.b[s], &full[stage], cb * 64, tap, nb);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp walks the schedule, one elected lane issues) =================
    {
      // segment 0 in fp16 or bf16 (compile-time), shortcut segments always bf16: the descriptors stay immediates
      constexpr uint32_t idesc0 = umma_idesc_f16kind(128, BN, A0F16, A0F16);
      constexpr uint32_t idesc1 = umma_idesc_f16kind(128, BN, false, false);
      const int kb_seg0 = a.taps[0] * a.cblks[0];
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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
            for (int k = 0; k < 4; ++k)  // 4 x (K = 16 bf16 = 32 B) per 64-channel block
              umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
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
        if (elect_one()) umma_commit(&tfull[acc]);
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    // two warps per TMEM lane quarter: warps 2-5 take the first half of the tile's columns, 6-9 the second
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int col_half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int rows_valid = a.tw * a.th * a.tn;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const int w0 = (m_tile % a.tiles_w) * a.tw;
      const int h0 = ((m_tile / a.tiles_w) % a.tiles_h) * a.th;
      const int n0 = (m_tile / (a.tiles_w * a.tiles_h)) * a.tn;
      const int dn = row / (a.th * a.tw);
      const int rem = row - dn * (a.th * a.tw);
      const int dh = rem / a.tw, dw = rem - dh * a.tw;
      const int b = n0 + dn;
      const bool valid = row < rows_valid && b < a.B;
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
        if (n >= a.epi.cout_valid) break;   // zero-padded output channels (uniform across the warp)
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
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (want_stats) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
        stats_flush(stat_smem, a, BN, acc, n_tile, m_tile, (warp - 2) * 32 + lane);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::tmem_cols>(tmem_base);
}

uint32_t conv_2cta_smem_bytes(int bn);
int launch_conv_2cta(const ConvPlan* plan, cudaStream_t st);
uint32_t conv_2cta_halo_smem_bytes(int bn);
int launch_conv_2cta_halo(const ConvPlan* plan, cudaStream_t st);

// ---- host ------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int encode_bf16_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* who) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("%s: cuTensorMapEncodeTiled not available (no CUDA driver?)", who);
    return ADVS_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", who,
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return ADVS_ERR_CUDA;
  }
  return ADVS_OK;
}

static int largest_divisor_leq(int n, int cap) {
  int best = 1;
  for (int d = 1; d <= cap && d <= n; ++d)
    if (n % d == 0) best = d;
  return best;
}

static int num_sms() { return current_device_sms(); }

}  // namespace advs

using namespace advs;

extern "C" {

int advs_conv_sm100_stats_parts(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  int tw = largest_divisor_leq(W, 128);
  int th = largest_divisor_leq(H, 128 / tw);
  // an image smaller than one tile: tiles span images when B > 1.  The answer must not depend on B (the statistics
  // path decides the fp32 summation order, and results are bit-reproducible across batch sizes), so B = 1 says 0 too
  if (th == H && tw == W && 128 / (tw * th) > 1) return 0;
  return (W / tw) * (H / th);
}

int advs_conv_sm100_plan(const advs_conv_params* p, void* plan_host) {
  int rc = validate_conv(p, "conv_sm100_plan");
  if (rc) return rc;
  ADVS_CHECK_ARG(plan_host && ((uintptr_t)plan_host % 64) == 0, "conv_sm100_plan: plan buffer must be 64-byte aligned");
  ADVS_CHECK_ARG(p->dtype == ADVS_BF16, "conv_sm100_plan: bf16 only");
  ADVS_CHECK_ARG(p->Cout % 64 == 0, "conv_sm100_plan: Cout must be a multiple of 64");
  for (int s = 0; s < p->nseg; ++s)
    ADVS_CHECK_ARG(p->seg[s].C % 64 == 0, "conv_sm100_plan: segment channels must be multiples of 64");
  if (p->out_mode == 1) {
    int dh = p->Cout / (3 * p->heads);
    ADVS_CHECK_ARG(dh % 32 == 0 || dh == 16, "conv_sm100_plan: head dim must be 16 or a multiple of 32");
    ADVS_CHECK_ARG(p->qkv_dh_pad == 0 || (p->qkv_dh_pad >= dh && p->qkv_dh_pad % 8 == 0),
                   "conv_sm100_plan: qkv_dh_pad must be a multiple of 8 and >= the head dim");
  }
  ADVS_CHECK_ARG(!p->bias || ((uintptr_t)p->bias % 16) == 0, "conv_sm100_plan: bias must be 16-byte aligned");
  ADVS_CHECK_ARG(!p->temb || (((uintptr_t)p->temb % 16) == 0 && p->temb_stride % 4 == 0),
                 "conv_sm100_plan: temb must be 16-byte aligned with stride%%4==0");

  ConvPlan* plan = reinterpret_cast<ConvPlan*>(plan_host);
  memset(plan, 0, sizeof(ConvPlan));
  ConvArgs& a = plan->args;
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cout = p->Cout;
  a.stride = p->stride; a.nseg = p->nseg;
  a.tw = largest_divisor_leq(p->W, 128);
  a.th = largest_divisor_leq(p->H, 128 / a.tw);
  a.tn = 1;
  if (a.th == p->H && a.tw == p->W) {
    a.tn = 128 / (a.tw * a.th);
    if (a.tn > p->B) a.tn = p->B;
    if (a.tn < 1) a.tn = 1;
  }
  a.tiles_w = p->W / a.tw;
  a.tiles_h = p->H / a.th;
  a.tiles_n = (p->B + a.tn - 1) / a.tn;
  a.m_tiles = a.tiles_w * a.tiles_h * a.tiles_n;
  a.a_bytes = (uint32_t)(a.tw * a.th * a.tn) * 128u;
  a.total_kb = 0;
  for (int s = 0; s < 3; ++s) {
    a.taps[s] = s < p->nseg ? p->seg[s].taps : 0;
    a.cblks[s] = s < p->nseg ? p->seg[s].C / 64 : 0;
    a.total_kb += a.taps[s] * a.cblks[s];
  }
  a.epi = make_epilogue(*p);
  a.up = p->up_phase != 0;
  a.up_a = a.up ? ((p->up_phase - 1) >> 1) : 0;
  a.up_b = a.up ? ((p->up_phase - 1) & 1) : 0;
  if (a.up) a.epi.HW = 4 * p->H * p->W;          // output pixels per image
  a.stats_tpi = a.tiles_w * a.tiles_h;
  a.stats_rpi = a.up ? 4 * a.stats_tpi : a.stats_tpi;
  a.stats_off = a.up ? (p->up_phase - 1) * a.stats_tpi : 0;
  a.stats = p->stats_partial;
  a.stats_gran = p->stats_gran == 4 ? 4 : 1;
  a.operand_f16 = p->operand_f16;
  // tcgen05.mma kind::f16 with different A and B formats is an illegal instruction on sm_100 (tools/gpu/probe_mixed_mma.py),
  // and the shortcut segments read the unbounded residual stream: only "segment 0 entirely fp16" exists
  ADVS_CHECK_ARG(p->operand_f16 == 0 || p->operand_f16 == 5,
                 "conv_sm100_plan: operand_f16 must be 0 or 5 (segment 0: activations AND weights in fp16)");
  if (a.stats) {
    ADVS_CHECK_ARG(p->stats_gran == 0 || p->stats_gran == 1 || p->stats_gran == 4, "conv_sm100_plan: stats_gran must be 0, 1 or 4");
    ADVS_CHECK_ARG(p->out_mode == 0, "conv_sm100_plan: stats_partial needs out_mode 0");
    ADVS_CHECK_ARG(a.tn == 1, "conv_sm100_plan: stats_partial needs tiles that do not span images (H*W >= 128)");
  }

  int bn = 128;
  if (p->Cout % 256 == 0 && (long long)a.m_tiles * (p->Cout / 256) >= num_sms()) bn = 256;
  plan->bn = bn;
  a.n_tiles = (p->Cout + bn - 1) / bn;
  // CTA-pair kernel unless disabled (ADVS_CONV_2CTA=0) or there is only one pixel tile
  static const int env_2cta = [] { const char* e = getenv("ADVS_CONV_2CTA"); return e ? atoi(e) : 1; }();
  const int two_cta = (env_2cta && a.m_tiles >= 2) ? 1 : 0;
  plan->two_cta = two_cta;
  // halo reuse: whole 128-pixel row segments, stride 1, a kernel wider than 1x1
  static const int env_halo = [] { const char* e = getenv("ADVS_CONV_HALO"); return e ? atoi(e) : 1; }();
  const int halo = (two_cta && env_halo && p->stride == 1 && a.tw == 128 && a.th == 1 && a.tn == 1 &&
                    (p->seg[0].taps == 9 || p->seg[0].taps == 4)) ? 1 : 0;
  plan->halo = halo;
  for (int s = 0; s < 3; ++s) {
    const int t = a.taps[s];
    a.a_bytes_seg[s] = (uint32_t)(t == 9 ? 130 * 3 : (t == 4 ? 129 * 2 : 128)) * 128u;
  }

  // ---- TMA descriptors ----
  uint32_t abox[4] = {64u, (uint32_t)a.tw, (uint32_t)a.th, (uint32_t)a.tn};
  for (int s = 0; s < p->nseg; ++s) {
    if (halo) {
      const int t = p->seg[s].taps;
      abox[1] = t == 9 ? 130u : (t == 4 ? 129u : 128u);
      abox[2] = t == 9 ? 3u : (t == 4 ? 2u : 1u);
    }
    const int C = p->seg[s].C;
    const char* base = reinterpret_cast<const char*>(p->seg[s].x);
    if (s == 0 && p->stride == 2) {
      const int Hin = 2 * p->H, Win = 2 * p->W;
      for (int hp = 0; hp < 2; ++hp)
        for (int wp = 0; wp < 2; ++wp) {
          uint64_t dims[4] = {(uint64_t)C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
          uint64_t str[4] = {2, (uint64_t)2 * C * 2, (uint64_t)2 * Win * C * 2, (uint64_t)Hin * Win * C * 2};
          const char* b2 = base + ((size_t)hp * Win + wp) * C * 2;
          rc = encode_bf16_map(&plan->maps.a[hp * 2 + wp], b2, 4, dims, str, abox, "conv_sm100_plan(A/stride2)");
          if (rc) return rc;
        }
    } else {
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->B};
      uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)p->W * C * 2, (uint64_t)p->H * p->W * C * 2};
      rc = encode_bf16_map(&plan->maps.a[s == 0 ? 0 : 3 + s], base, 4, dims, str, abox, "conv_sm100_plan(A)");
      if (rc) return rc;
    }
    uint64_t wd[3] = {(uint64_t)C, (uint64_t)p->seg[s].taps, (uint64_t)p->Cout};
    uint64_t ws[3] = {2, (uint64_t)C * 2, (uint64_t)p->seg[s].taps * C * 2};
    const uint32_t wbox[3] = {64u, 1u, (uint32_t)(two_cta ? bn / 2 : bn)};   // CTA pair: each CTA loads half the rows
    rc = encode_bf16_map(&plan->maps.b[s], p->seg[s].w, 3, wd, ws, wbox, "conv_sm100_plan(W)");
    if (rc) return rc;
  }
  // unused descriptor slots: replicate a valid one so prefetch.tensormap never sees garbage
  for (int i = 0; i < 6; ++i) {
    bool used = (i == 0) || (p->stride == 2 && i < 4) || (i >= 4 && (i - 3) < p->nseg);
    if (!used) plan->maps.a[i] = plan->maps.a[0];
  }
  for (int i = p->nseg; i < 3; ++i) plan->maps.b[i] = plan->maps.b[0];

  const int total_tiles = a.m_tiles * a.n_tiles;
  if (two_cta) {
    const int items = ((a.m_tiles + 1) / 2) * a.n_tiles, max_pairs = num_sms() / 2;
    plan->grid = 2 * (items < max_pairs ? items : max_pairs);
    plan->smem_bytes = halo ? conv_2cta_halo_smem_bytes(bn) : conv_2cta_smem_bytes(bn);
  } else {
    plan->grid = total_tiles < num_sms() ? total_tiles : num_sms();
    plan->smem_bytes = bn == 256 ? ConvCfg<256>::smem_bytes : ConvCfg<128>::smem_bytes;
  }
  plan->magic = 0xC0A7B200u;
  return ADVS_OK;
}

int advs_conv_sm100_launch(const void* plan_host, void* stream) {
  const ConvPlan* plan = reinterpret_cast<const ConvPlan*>(plan_host);
  ADVS_CHECK_ARG(plan && plan->magic == 0xC0A7B200u, "conv_sm100_launch: not a plan");
  if (plan->halo) return launch_conv_2cta_halo(plan, (cudaStream_t)stream);
  if (plan->two_cta) return launch_conv_2cta(plan, (cudaStream_t)stream);
  if (first_use_on_device(kOnceConv1Cta)) {
    cudaError_t e1 = cudaFuncSetAttribute(k_conv_sm100<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfg<128>::smem_bytes);
    cudaError_t e2 = cudaFuncSetAttribute(k_conv_sm100<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          ConvCfg<256>::smem_bytes);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(k_conv_sm100<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfg<128>::smem_bytes);
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_conv_sm100<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     ConvCfg<256>::smem_bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
      forget_first_use(kOnceConv1Cta);
      set_error("conv_sm100_launch: cudaFuncSetAttribute failed: %s",
                cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      return ADVS_ERR_CUDA;
    }
  }
  const bool f16 = plan->args.operand_f16 != 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->bn == 256) {
    if (f16) launch_pdl(k_conv_sm100<256, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100<256, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  } else {
    if (f16) launch_pdl(k_conv_sm100<128, true>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
    else launch_pdl(k_conv_sm100<128, false>, dim3(plan->grid), dim3(kConvThreads), plan->smem_bytes, st, plan->maps, plan->args);
  }
  ADVS_CHECK_LAUNCH("conv_sm100_launch");
  return ADVS_OK;
}

}  // extern "C"
