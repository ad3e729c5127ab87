// K4: fused self-attention softmax(q^T k) v on the 5th-gen tensor cores (flash-style, no T x T matrix).
//   reference: AttentionBlock.forward dm1:120-125 (two einsums around an fp32 softmax that
//   materialises [B*heads, T, T]).
//
// One CTA per (image*head, 128-query tile); keys/values stream through in blocks of 128:
//   warp 0   : TMA producer (Q once; K_j, V^T_j per block)
//   warp 1   : tcgen05.mma issuer   S_j = Q K_j^T  (TMEM, double buffered),  O += P_j V_j  (TMEM)
//   warps 2-9: online softmax, one query row per thread PAIR (two warps split each block's keys): S row TMEM -> registers, running max /
//              sum in fp32 (base-2 domain), P -> bf16 -> swizzled smem for the PV MMA; the O
//              accumulator is rescaled in TMEM only when the running max grew by more than 2^8
//              (lazy rescale); final O / l -> bf16 NHWC.
// q and k arrive pre-scaled by dh^-1/4 each (dm1:121-122), so the softmax scale is 1.
#include <string.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advs {

using namespace sm100;

struct AttnMaps {
  CUtensorMap q, k, vt;
};

struct AttnArgs {
  int B, heads, T, dh;
  __nv_bfloat16* o;
};

struct AttnPlan {
  AttnMaps maps;
  AttnArgs args;
  uint32_t smem_bytes;
  uint32_t magic;
};
static_assert(sizeof(AttnPlan) <= ADVS_ATTN_PLAN_BYTES, "AttnPlan does not fit ADVS_ATTN_PLAN_BYTES");

constexpr int kAttnThreads = 320;   // TMA warp + MMA warp + 8 softmax warps
constexpr int kBQ = 128;   // queries per CTA
constexpr int kBK = 128;   // keys per block
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLazyThreshold = 8.0f;  // rescale O only if the row max grew by > 2^8
// (ex2.approx.ftz.bf16x2 was tried to halve the MUFU load: ptxas splits it into two MUFU.EX2.BF16 ops on
//  sm_100a, so it buys nothing and costs accuracy -- scores keep the fp32 ex2.)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef ADVS_ATTN_TRACE
__device__ long long g_attn_trace[64 * 16];
#define TR(slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < 64) g_attn_trace[j * 16 + (slot)] = clock64(); } while (0)
#else
#define TR(slot) do { } while (0)
#endif

template <int DH>
struct AttnCfg {
  static constexpr int kv_stages = (DH == 256) ? 1 : 2;
  static constexpr uint32_t q_bytes = kBQ * DH * 2;
  static constexpr uint32_t k_bytes = kBK * DH * 2;
  static constexpr uint32_t v_bytes = DH * kBK * 2;
  static constexpr uint32_t p_bytes = kBQ * kBK * 2;
  static constexpr uint32_t off_q = 0;
  static constexpr uint32_t off_k = off_q + q_bytes;
  static constexpr uint32_t off_v = off_k + kv_stages * k_bytes;
  static constexpr uint32_t off_p = off_v + kv_stages * v_bytes;
  static constexpr uint32_t off_bar = off_p + p_bytes;
  static constexpr uint32_t smem_bytes = off_bar + 128 + 2048 + 896;   // barriers + max/sum exchange + alignment slack (base is 128-B aligned)
  static constexpr uint32_t tmem_cols = 512;
  static constexpr uint32_t o_col = 256;
};

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_sm100(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  using Cfg = AttnCfg<DH>;
  constexpr int KVS = Cfg::kv_stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::off_bar);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // KVS
  uint64_t* k_empty = bars + 3;       // KVS
  uint64_t* v_full = bars + 5;        // KVS
  uint64_t* v_empty = bars + 7;       // KVS
  uint64_t* s_full = bars + 9;        // 2
  uint64_t* p_full = bars + 11;       // 1
  uint64_t* o_done = bars + 12;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int nblk = a.T / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.vt);
    mbar_init(q_full, 1);
    for (int s = 0; s < KVS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 8);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::tmem_cols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::q_bytes);
      for (int sl = 0; sl < DH / 64; ++sl)
        tma_load_2d(smem + Cfg::off_q + sl * (kBQ * 128), &maps.q, q_full, sl * 64, bh * a.T + q_tile * kBQ);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::k_bytes);
        for (int sl = 0; sl < DH / 64; ++sl)
          tma_load_2d(smem + Cfg::off_k + st * Cfg::k_bytes + sl * (kBK * 128), &maps.k, &k_full[st], sl * 64,
                      bh * a.T + j * kBK);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], Cfg::v_bytes);
        for (int sl = 0; sl < kBK / 64; ++sl)
          tma_load_2d(smem + Cfg::off_v + st * Cfg::v_bytes + sl * (DH * 128), &maps.vt, &v_full[st],
                      j * kBK + sl * 64, bh * DH);
        if (++st == KVS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp walks the schedule, one elected lane issues) =================
    {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kBK);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH);
      const uint32_t q_addr = smem_u32(smem + Cfg::off_q);
      const uint32_t p_addr = smem_u32(smem + Cfg::off_p);
      // S_j = Q K_j^T, then signal the softmax warps and release the K stage (elected lane only)
      auto issue_s = [&](int j, int st) {
        const uint32_t k_addr = smem_u32(smem + Cfg::off_k + st * Cfg::k_bytes);
        const uint32_t d = tmem_base + (uint32_t)((j & 1) * kBK);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            const uint32_t off = (uint32_t)(k >> 2) * (128 * 128) + (uint32_t)(k & 3) * 32;
            umma_bf16(d, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off), idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[j & 1]);
          umma_commit(&k_empty[st]);
        }
        __syncwarp();
      };
      int st = 0;       // stage of block j (the PV side)
      uint32_t ph = 0;
      int st_s = 0;     // stage of the S being issued
      uint32_t ph_s = 0;
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      if (++st_s == KVS) { st_s = 0; ph_s ^= 1; }
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          TR(11);
          mbar_wait(&k_full[st_s], ph_s);
          tc_fence_after();
          TR(12);
          issue_s(j + 1, st_s);
          if (++st_s == KVS) { st_s = 0; ph_s ^= 1; }
        }
        TR(8);
        mbar_wait(p_full, (uint32_t)(j & 1));
        TR(9);
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        TR(10);
        const uint32_t v_addr = smem_u32(smem + Cfg::off_v + st * Cfg::v_bytes);
        const uint32_t d = tmem_base + Cfg::o_col;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint32_t offp = (uint32_t)(k >> 2) * (kBQ * 128) + (uint32_t)(k & 3) * 32;
            const uint32_t offv = (uint32_t)(k >> 2) * (DH * 128) + (uint32_t)(k & 3) * 32;
            umma_bf16(d, umma_desc_k_sw128(p_addr + offp), umma_desc_k_sw128(v_addr + offv), idesc_o,
                      (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(o_done);
          umma_commit(&v_empty[st]);
        }
        __syncwarp();
        if (++st == KVS) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ================= softmax / correction / output (warps 2..9) =================
    // Two warps per TMEM lane quarter: warps 2-5 own keys [0,64) of each block, warps 6-9 keys [64,128).
    // The softmax is MUFU-bound (one ex2 per score); two warps per scheduler let one warp's TMEM loads,
    // smem stores and barrier waits hide behind the other's exponentials.  The row maximum is exchanged
    // through shared memory once per block; the row sums are combined once at the end.
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint8_t* p_smem = smem + Cfg::off_p;
    float* xch = reinterpret_cast<float*>(smem + Cfg::off_bar + 128);   // [2 parities][2 halves][128 rows]
    constexpr int HB = kBK / 2;   // keys per warp
    constexpr int HD = DH / 2;    // O columns per warp (rescale / output)
    float m_used = -INFINITY;     // base-2 running max actually subtracted (identical in both partners)
    float l = 0.f;                // this half's share of the row sum
    for (int j = 0; j < nblk; ++j) {
      if (warp == 2) TR(0);
      mbar_wait(&s_full[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      if (warp == 2) TR(1);
      float s[HB];
#pragma unroll
      for (int c = 0; c < HB / 32; ++c)
        tmem_ld_32x32b_x32(lane_addr + (uint32_t)((j & 1) * kBK + half * HB + c * 32), reinterpret_cast<uint32_t*>(s) + c * 32);
      tmem_wait_ld();
      if (warp == 2) TR(2);
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < HB; ++i) mx = fmaxf(mx, s[i]);
      // exchange the half-row maxima with the partner warp (same rows, other 64 keys)
      xch[((j & 1) * 2 + half) * 128 + row] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * 128 + row]) * kLog2e;
      if (warp == 2) TR(3);
      float alpha = 1.f;
      bool grow = mx > m_used + kLazyThreshold;
      if (grow) {
        alpha = exp2f(m_used - mx);  // 0 on the first block (m_used = -inf)
        m_used = mx;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < HB; ++i) {
        s[i] = fast_exp2(fmaf(s[i], kLog2e, -m_used));
        sum += s[i];
      }
      l = fmaf(l, alpha, sum);
      if (warp == 2) TR(4);
      // previous PV must be done before P is overwritten / O is rescaled
      if (j > 0) {
        mbar_wait(o_done, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        if (warp == 2) TR(5);
        if (__any_sync(0xffffffffu, grow)) {   // both partners take the same decision (same max); each rescales half of O
#pragma unroll 1
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(lane_addr + Cfg::o_col + half * HD + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x32(lane_addr + Cfg::o_col + half * HD + c * 32, r);
          }
          tmem_wait_st();
        }
      }
      // P -> bf16 -> smem, K-major SWIZZLE_128B: slab `half` of [128 rows][64 keys]
#pragma unroll
      for (int c8 = 0; c8 < HB / 8; ++c8) {
        uint4 v;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(s[c8 * 8 + 0], s[c8 * 8 + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(s[c8 * 8 + 2], s[c8 * 8 + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(s[c8 * 8 + 4], s[c8 * 8 + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(s[c8 * 8 + 6], s[c8 * 8 + 7]);
        v.x = *reinterpret_cast<uint32_t*>(&h0);
        v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2);
        v.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(p_smem + half * (kBQ * 128) + row * 128 + ((c8 ^ (row & 7)) << 4)) = v;
      }
      if (warp == 2) TR(6);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 2) TR(7);
    }
    // ---- output: O / l (row sum = both halves) ----
    asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");   // partner has finished reading the last max
    xch[half * 128 + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
    l += xch[(half ^ 1) * 128 + row];
    mbar_wait(o_done, (uint32_t)((nblk - 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    const int b = bh / a.heads, head = bh - b * a.heads;
    __nv_bfloat16* orow = a.o + ((size_t)b * a.T + q_tile * kBQ + row) * ((size_t)a.heads * DH) + (size_t)head * DH + half * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_addr + Cfg::o_col + half * HD + c * 32, r);
      tmem_wait_ld();
      uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 v;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
        v.x = *reinterpret_cast<uint32_t*>(&h0);
        v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2);
        v.w = *reinterpret_cast<uint32_t*>(&h3);
        dst[i] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::tmem_cols>(tmem_base);
}

template <int DH>
static int attn_launch(const AttnPlan* plan, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(k_attention_sm100<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         AttnCfg<DH>::smem_bytes);
    if (e != cudaSuccess) {
      set_error("attention_sm100_launch: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ADVS_ERR_CUDA;
    }
    attr_done = true;
  }
  dim3 grid(plan->args.T / kBQ, plan->args.B * plan->args.heads);
  k_attention_sm100<DH><<<grid, kAttnThreads, AttnCfg<DH>::smem_bytes, st>>>(plan->maps, plan->args);
  ADVS_CHECK_LAUNCH("attention_sm100_launch");
  return ADVS_OK;
}

}  // namespace advs

using namespace advs;

#ifdef ADVS_ATTN_TRACE
extern "C" int advs_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(long long) * 64 * 16) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" {

int advs_attention_sm100_plan(const void* q, const void* k, const void* vt, void* o, int B, int heads, int T, int dh,
                              void* plan_host) {
  ADVS_CHECK_ARG(q && k && vt && o && plan_host, "attention_sm100_plan: null pointer");
  ADVS_CHECK_ARG(((uintptr_t)plan_host % 64) == 0, "attention_sm100_plan: plan buffer must be 64-byte aligned");
  ADVS_CHECK_ARG(B > 0 && heads > 0 && T > 0 && T % 128 == 0, "attention_sm100_plan: T must be a positive multiple of 128");
  ADVS_CHECK_ARG(dh == 64 || dh == 128 || dh == 256, "attention_sm100_plan: dh must be 64, 128 or 256");
  ADVS_CHECK_ARG((long long)B * heads <= 65535, "attention_sm100_plan: B*heads must be <= 65535");
  AttnPlan* plan = reinterpret_cast<AttnPlan*>(plan_host);
  memset(plan, 0, sizeof(AttnPlan));
  const uint64_t rows = (uint64_t)B * heads * T;
  {
    uint64_t dims[2] = {(uint64_t)dh, rows};
    uint64_t str[2] = {2, (uint64_t)dh * 2};
    uint32_t box[2] = {64u, 128u};
    int rc = encode_bf16_map(&plan->maps.q, q, 2, dims, str, box, "attention_sm100_plan(Q)");
    if (rc) return rc;
    rc = encode_bf16_map(&plan->maps.k, k, 2, dims, str, box, "attention_sm100_plan(K)");
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)T, (uint64_t)B * heads * dh};
    uint64_t str[2] = {2, (uint64_t)T * 2};
    uint32_t box[2] = {64u, (uint32_t)dh};
    int rc = encode_bf16_map(&plan->maps.vt, vt, 2, dims, str, box, "attention_sm100_plan(VT)");
    if (rc) return rc;
  }
  plan->args.B = B;
  plan->args.heads = heads;
  plan->args.T = T;
  plan->args.dh = dh;
  plan->args.o = reinterpret_cast<__nv_bfloat16*>(o);
  plan->magic = 0xA77EB200u;
  return ADVS_OK;
}

int advs_attention_sm100_launch(const void* plan_host, void* stream) {
  const AttnPlan* plan = reinterpret_cast<const AttnPlan*>(plan_host);
  ADVS_CHECK_ARG(plan && plan->magic == 0xA77EB200u, "attention_sm100_launch: not a plan");
  switch (plan->args.dh) {
    case 64: return attn_launch<64>(plan, (cudaStream_t)stream);
    case 128: return attn_launch<128>(plan, (cudaStream_t)stream);
    case 256: return attn_launch<256>(plan, (cudaStream_t)stream);
  }
  ADVS_CHECK_ARG(false, "attention_sm100_launch: bad dh");
}

}  // extern "C"
