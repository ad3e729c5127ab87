// Shared pieces of the tcgen05 implicit-GEMM convolution kernels (1-CTA and 2-CTA variants).
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace advs {

using namespace sm100;

struct ConvMaps {
  CUtensorMap a[6];  // [0..3] segment 0 (stride 1: only [0]; stride 2: parity hp*2+wp), [4],[5] segments 1,2
  CUtensorMap b[3];
};

struct ConvArgs {
  int B, H, W, Cout;
  int tw, th, tn;
  int tiles_w, tiles_h, tiles_n, m_tiles, n_tiles;
  int stride, nseg;
  int taps[3], cblks[3];
  int total_kb;
  uint32_t a_bytes;
  uint32_t a_bytes_seg[3];   // halo kernel: bytes of the activation box of each K-segment
  float* stats;   // optional per-tile channel sums / sums of squares (GroupNorm fusion)
  int stats_tpi, stats_rpi, stats_off;   // row = (m_tile / tpi) * rpi + off + m_tile % tpi
  int stats_gran;                        // 1: one {sum, sumsq} per channel; 4: per 4 consecutive channels
  int up_a, up_b, up;                    // upsample phase: output pixel (2h+a, 2w+b); up = 0/1
  int operand_f16;                       // advs_conv_params.operand_f16: 0, or 5 = segment 0 (activations and weights) in fp16
  EpilogueParams epi;
};

struct ConvPlan {
  ConvMaps maps;
  ConvArgs args;
  int bn;
  int grid;
  uint32_t smem_bytes;
  uint32_t magic;
  int two_cta;   // 1: CTA-pair kernel (cta_group::2), grid is a multiple of 2
  int halo;      // 1: CTA-pair kernel with one activation halo box per channel block (conv_sm100_halo.cu)
};
static_assert(sizeof(ConvPlan) <= ADVS_CONV_PLAN_BYTES, "ConvPlan does not fit ADVS_CONV_PLAN_BYTES");

constexpr int kConvThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr uint32_t kABytes = 128 * 128;  // 128 rows x 64 bf16

template <int BN>
struct ConvCfg {
  static constexpr uint32_t b_bytes = BN * 128;
  static constexpr uint32_t stage_bytes = kABytes + b_bytes;
  static constexpr int stages = (BN == 256) ? 4 : 6;
  static constexpr uint32_t bar_bytes = 256;
  static constexpr uint32_t stat_bytes = 2 * 4 * BN * 8;   // [2 accumulators][4 warps][BN] float2
  static constexpr uint32_t smem_bytes = stages * stage_bytes + bar_bytes + stat_bytes + 1024;  // + alignment slack
  static constexpr uint32_t tmem_cols = 2 * BN;
};

// 256-bit global accesses (sm_100): a thread's 32 output channels are two 32-byte pieces instead of four
// 16-byte ones.  Every thread of the warp is on a different pixel row, so each access instruction touches 32
// cache lines whatever its width -- halving the instruction count halves the LSU time of the epilogue.
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint32_t* w) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}

// 32 consecutive output channels [n, n+32) of one pixel row: v = acc + bias + temb + residual
__device__ __forceinline__ void epilogue_compute32(const EpilogueParams& e, const uint32_t* acc, float* v, size_t m,
                                                   int b, int n) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 t4 = __ldg(bp + j);
      v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
    }
  }
  if (e.temb) {
    const float4* tp = reinterpret_cast<const float4*>(e.temb + (size_t)b * e.temb_stride + n);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 t4 = __ldg(tp + j);
      v[4 * j] += t4.x; v[4 * j + 1] += t4.y; v[4 * j + 2] += t4.z; v[4 * j + 3] += t4.w;
    }
  }
  if (e.out_mode == 0 && e.residual) {
    const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(e.residual) + m * e.Cout + n;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t r8[8];
      ld_global_nc_256(rp + 16 * j, r8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(r8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[16 * j + 2 * i] += f.x;
        v[16 * j + 2 * i + 1] += f.y;
      }
    }
  }
}

// WIDE: also store the int8 mantissa extension (e.y_lo != nullptr).  A template parameter, chosen by a warp-uniform
// branch at the call site: with a run-time `if (e.y_lo)` inside the unrolled loops ptxas predicated the encoder
// into every instantiation and the plain epilogue lost 25 % (profiles/microbench_r02.md).
template <bool WIDE>
__device__ __forceinline__ void epilogue_write32(const EpilogueParams& e, const float* v, size_t m, int b, int t, int n) {
  if (e.out_mode == 0) {
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(e.y) + m * e.Cout + n;
    if constexpr (WIDE) {
      uint32_t lo8[8];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t u[16], o[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) u[i] = wide_round_bits(v[16 * j + i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = wide_hi2(u[2 * i], u[2 * i + 1]);
        st_global_256(yp + 16 * j, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) lo8[4 * j + i] = wide_lo4(u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
      }
      st_global_256(e.y_lo + m * e.Cout + n, lo8);   // 32 channels x int8
    } else {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = pack_bf16x2(v[16 * j + 2 * i], v[16 * j + 2 * i + 1]);
        st_global_256(yp + 16 * j, o);
      }
    }
  } else if (e.out_mode == 2) {
    float* dst = reinterpret_cast<float*>(e.y) + ((size_t)b * e.cout_valid + n) * e.HW + t;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < e.cout_valid) dst[(size_t)j * e.HW] = v[j];
  } else if (e.dh >= 32) {
    // 32 consecutive output channels = one piece of a head's [q | k | v] block
    const int head = n / (3 * e.dh);
    const int r = n - head * 3 * e.dh;
    const int which = r / e.dh;
    const int d0 = r - which * e.dh;
    const size_t bh = (size_t)b * e.heads + head;
    if (which < 2) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(which == 0 ? e.q : e.k) + (bh * e.HW + t) * e.dh_pad + d0;
      uint4* yp = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        o.x = pack_bf16x2(v[8 * j] * e.qk_scale, v[8 * j + 1] * e.qk_scale);
        o.y = pack_bf16x2(v[8 * j + 2] * e.qk_scale, v[8 * j + 3] * e.qk_scale);
        o.z = pack_bf16x2(v[8 * j + 4] * e.qk_scale, v[8 * j + 5] * e.qk_scale);
        o.w = pack_bf16x2(v[8 * j + 6] * e.qk_scale, v[8 * j + 7] * e.qk_scale);
        yp[j] = o;
      }
    } else {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.vt) + (bh * e.dh_pad + d0) * e.HW + t;
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[(size_t)j * e.HW] = __float2bfloat16_rn(v[j]);
    }
  } else {
    // head dim 16 (IDDM): the 32 channels are two 16-wide pieces of (possibly different) q / k / v blocks
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int nn = n + sub * 16;
      const int head = nn / (3 * e.dh);
      const int r = nn - head * 3 * e.dh;
      const int which = r / e.dh;
      const int d0 = r - which * e.dh;
      const size_t bh = (size_t)b * e.heads + head;
      const float* vv = v + sub * 16;
      if (which < 2) {
        uint4* yp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(which == 0 ? e.q : e.k) + (bh * e.HW + t) * e.dh_pad + d0);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          yp[j] = make_uint4(pack_bf16x2(vv[8 * j] * e.qk_scale, vv[8 * j + 1] * e.qk_scale),
                             pack_bf16x2(vv[8 * j + 2] * e.qk_scale, vv[8 * j + 3] * e.qk_scale),
                             pack_bf16x2(vv[8 * j + 4] * e.qk_scale, vv[8 * j + 5] * e.qk_scale),
                             pack_bf16x2(vv[8 * j + 6] * e.qk_scale, vv[8 * j + 7] * e.qk_scale));
      } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.vt) + (bh * e.dh_pad + d0) * e.HW + t;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(size_t)j * e.HW] = __float2bfloat16_rn(vv[j]);
      }
    }
  }
}

// Column sums over the 32 rows held by a warp: v[i] of lane l = value (row l, column i).  After the five
// exchange rounds lane l holds the sum of column l (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_column_sums(float* v, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? v[i + s] : v[i];
      const float send = upper ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// GroupNorm statistics of the 32 rows x 32 columns a warp holds (v[i] of lane l = row l, column i; rows that are
// not valid count as zero).  gran = 1: lane l returns {sum, sumsq} of column l (31 shuffles per quantity).
// gran = 4: the four columns of a channel quad are added in the thread first, then 8 values go through the
// exchange: 9 shuffles per quantity; lanes 0..7 return quad (lane) -- every GroupNorm on the path has a multiple
// of four channels per group, and the partial rows shrink fourfold.
__device__ __forceinline__ float2 warp_stats32(float* v, bool valid, int lane, int gran) {
  if (gran == 4) {
    float g[8], gq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float a0 = valid ? v[4 * k] : 0.f, a1 = valid ? v[4 * k + 1] : 0.f;
      const float a2 = valid ? v[4 * k + 2] : 0.f, a3 = valid ? v[4 * k + 3] : 0.f;
      g[k] = (a0 + a1) + (a2 + a3);
      gq[k] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3)));
    }
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
      const bool upper = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float keep = upper ? g[i + s] : g[i], send = upper ? g[i] : g[i + s];
        g[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        const float keepq = upper ? gq[i + s] : gq[i], sendq = upper ? gq[i] : gq[i + s];
        gq[i] = keepq + __shfl_xor_sync(0xffffffffu, sendq, s);
      }
    }
    float cs = g[0], cq = gq[0];
    cs += __shfl_xor_sync(0xffffffffu, cs, 8);
    cq += __shfl_xor_sync(0xffffffffu, cq, 8);
    cs += __shfl_xor_sync(0xffffffffu, cs, 16);
    cq += __shfl_xor_sync(0xffffffffu, cq, 16);
    return make_float2(cs, cq);
  }
  float sq[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = valid ? v[j] : 0.f;
    v[j] = x;
    sq[j] = x * x;
  }
  const float cs = warp_column_sums(v, lane);
  const float cq = warp_column_sums(sq, lane);
  return make_float2(cs, cq);
}

// stat_smem: [2 accumulators][4 row quarters][bn / gran] float2.  One warp's 32-column chunk:
__device__ __forceinline__ void stats_stage(float2* stat_smem, const ConvArgs& a, int bn, int acc, int q, int chunk, int lane,
                                            float2 st) {
  const int per = 32 / a.stats_gran;
  if (lane < per) stat_smem[(acc * 4 + q) * (bn / a.stats_gran) + chunk * per + lane] = st;
}
// the 256 epilogue threads add the four row quarters of a finished tile and write its partial row
__device__ __forceinline__ void stats_flush(const float2* stat_smem, const ConvArgs& a, int bn, int acc, int n_tile, int m_tile,
                                            int e) {
  const int slots = bn / a.stats_gran, total = a.Cout / a.stats_gran;
  const size_t prow = (size_t)((m_tile / a.stats_tpi) * a.stats_rpi + a.stats_off + m_tile % a.stats_tpi);
  for (int c = e; c < slots; c += 256) {
    const int n = n_tile * slots + c;
    if (n < total) {
      const float2 t0 = stat_smem[(acc * 4 + 0) * slots + c], t1 = stat_smem[(acc * 4 + 1) * slots + c];
      const float2 t2 = stat_smem[(acc * 4 + 2) * slots + c], t3 = stat_smem[(acc * 4 + 3) * slots + c];
      reinterpret_cast<float2*>(a.stats)[prow * total + n] =
          make_float2((t0.x + t1.x) + (t2.x + t3.x), (t0.y + t1.y) + (t2.y + t3.y));
    }
  }
}

}  // namespace advs
