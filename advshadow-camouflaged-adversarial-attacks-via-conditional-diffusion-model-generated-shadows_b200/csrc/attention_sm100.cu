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

// Three warpgroups: {TMA warp, MMA warp, two idle warps} + 8 softmax warps.  Registers are allocated to warps
// in groups of four, so 12 warps start at 168 registers each; the first group hands most of its share to
// the softmax groups (setmaxnreg), which keep two blocks of scores in registers.
constexpr int kAttnThreads = 384;
constexpr int kAttnFirstSoftmaxWarp = 4;
constexpr int kBQ = 128;   // queries per CTA
constexpr int kBK = 128;   // keys per block
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLazyThreshold = 8.0f;  // rescale O only if the row max grew by > 2^8
// (ex2.approx.ftz.bf16x2 was tried to halve the MUFU load: ptxas splits it into two MUFU.EX2.BF16 ops on
//  sm_100a, so it buys nothing and costs accuracy -- scores keep the fp32 ex2.)
// volatile: the softmax loop places its exponentials by hand between barrier waits (see below)
__device__ __forceinline__ float ex2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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
  // dh <= 128: P lives in tensor memory (two tiles of 64 columns next to S0 | S1 | O) and is the PV MMA's A
  // operand from there -- no shared-memory round trip, and PV only streams V.  dh = 256 fills TMEM with
  // S and O, so its P goes through one swizzled shared-memory tile.
  static constexpr bool p_in_tmem = DH <= 128;
  static constexpr int p_bufs = p_in_tmem ? 2 : 1;
  static constexpr uint32_t off_bar = off_p + (p_in_tmem ? 0 : p_bytes);
  static constexpr uint32_t p_col = 384;
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
  uint64_t* o_done = bars + 12;       // 2: PV of even / odd blocks (a waiter may then lag two blocks behind)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

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
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::tmem_cols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kAttnFirstSoftmaxWarp) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::q_bytes);
      for (int sl = 0; sl < DH / 64; ++sl)
        tma_load_2d(smem + Cfg::off_q + sl * (kBQ * 128), &maps.q, q_full, sl * 64, bh * a.T + q_tile * kBQ);
      // keys: stage st is free again as soon as S_j has drained
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], Cfg::k_bytes);
        for (int sl = 0; sl < DH / 64; ++sl)
          tma_load_2d(smem + Cfg::off_k + st * Cfg::k_bytes + sl * (kBK * 128), &maps.k, &k_full[st], sl * 64,
                      bh * a.T + j * kBK);
        if (++st == KVS) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ================= TMA producer for V^T (its own warp: a V stage is released by PV_j, much later than the
    // K stage of the same block, and one thread issuing both in order would hold the next S behind it) ====
    if (lane == 0) {
      tma_prefetch_desc(&maps.vt);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
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
      // PV of block j, then release the V stage and tell the softmax warps P / O may be touched again
      auto issue_pv = [&](int j, int st) {
        const uint32_t v_addr = smem_u32(smem + Cfg::off_v + st * Cfg::v_bytes);
        const uint32_t d = tmem_base + Cfg::o_col;
        const uint32_t p_tmem = tmem_base + Cfg::p_col + (uint32_t)(j & 1) * (kBK / 2);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint32_t offp = (uint32_t)(k >> 2) * (kBQ * 128) + (uint32_t)(k & 3) * 32;
            const uint32_t offv = (uint32_t)(k >> 2) * (DH * 128) + (uint32_t)(k & 3) * 32;
            if (Cfg::p_in_tmem)
              umma_bf16_ts(d, p_tmem + (uint32_t)k * 8, umma_desc_k_sw128(v_addr + offv), idesc_o, (j | k) != 0 ? 1u : 0u);
            else
              umma_bf16(d, umma_desc_k_sw128(p_addr + offp), umma_desc_k_sw128(v_addr + offv), idesc_o,
                        (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&o_done[j & 1]);
          umma_commit(&v_empty[st]);
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
      if (KVS >= 2) {
        // Two key stages: S runs TWO blocks ahead.  When the softmax warps hand over P_j they have long since
        // read S_{j+1} into registers, so S_{j+2} is issued first and PV_j second: S_{j+2} is complete by the
        // middle of the exponentials of block j+1, whose shadow then hides the max / exchange of block j+2.
        if (nblk > 1) {
          mbar_wait(&k_full[st_s], ph_s);
          tc_fence_after();
          issue_s(1, st_s);
          if (++st_s == KVS) { st_s = 0; ph_s ^= 1; }
        }
        for (int j = 0; j < nblk; ++j) {
          TR(8);
          mbar_wait(p_full, (uint32_t)(j & 1));
          TR(9);
          if (j + 2 < nblk) {
            mbar_wait(&k_full[st_s], ph_s);
            tc_fence_after();
            TR(11);
            issue_s(j + 2, st_s);
            TR(12);
            if (++st_s == KVS) { st_s = 0; ph_s ^= 1; }
          }
          mbar_wait(&v_full[st], ph);
          tc_fence_after();
          TR(10);
          issue_pv(j, st);
          TR(13);
          if (++st == KVS) { st = 0; ph ^= 1; }
        }
      } else {
        // one key stage (dh = 256): S one block ahead
        for (int j = 0; j < nblk; ++j) {
          if (j + 1 < nblk) {
            mbar_wait(&k_full[st_s], ph_s);
            tc_fence_after();
            issue_s(j + 1, st_s);
            if (++st_s == KVS) { st_s = 0; ph_s ^= 1; }
          }
          mbar_wait(p_full, (uint32_t)(j & 1));
          mbar_wait(&v_full[st], ph);
          tc_fence_after();
          issue_pv(j, st);
          if (++st == KVS) { st = 0; ph ^= 1; }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ================= softmax / correction / output (warps 4..11) =================
    // Two warps per TMEM lane quarter: warps 4-7 own keys [0,64) of each block, warps 8-11 keys [64,128).
    // The softmax is MUFU-bound (one ex2 per score); two warps per scheduler let one warp's TMEM loads,
    // smem stores and barrier waits hide behind the other's exponentials.  The row maximum is exchanged
    // through shared memory once per block; the row sums are combined once at the end.
    const int qd = warp & 3;
    const int half = (warp - kAttnFirstSoftmaxWarp) >> 2;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint8_t* p_smem = smem + Cfg::off_p;
    float* xch = reinterpret_cast<float*>(smem + Cfg::off_bar + 128);   // [2 parities][2 halves][128 rows]
    constexpr int HB = kBK / 2;   // keys per warp
    constexpr int HD = DH / 2;    // O columns per warp (rescale / output)
    float m_used = -INFINITY;     // base-2 running max actually subtracted (identical in both partners)
    float l = 0.f;                // this half's share of the row sum
    // The loop is software-pipelined by hand around the MUFU pipe (16 ex2 / clk / SM is the softmax bound; the
    // two warps of a row pair share a scheduler and run in lockstep, so nothing else hides their latencies):
    // the 64 exponentials of block j are issued in four groups, and between the groups the warp fetches S_{j+1}
    // from TMEM, takes its row maximum and exchanges it with the partner.  All pieces are volatile asm so
    // ptxas keeps this order.
    float sa[HB], sb[HB];
    auto ld_scores = [&](int j, float* dst) {
#pragma unroll
      for (int c = 0; c < HB / 32; ++c)
        tmem_ld_32x32b_x32(lane_addr + (uint32_t)((j & 1) * kBK + half * HB + c * 32), reinterpret_cast<uint32_t*>(dst) + c * 32);
    };
    auto half_max = [&](const float* v) {
      float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
      for (int i = 4; i < HB; i += 4) {
        m0 = fmaxf(m0, v[i]); m1 = fmaxf(m1, v[i + 1]); m2 = fmaxf(m2, v[i + 2]); m3 = fmaxf(m3, v[i + 3]);
      }
      return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    };
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
    ld_scores(0, sa);
    tmem_wait_ld();
    float mx = half_max(sa);
    xch[half * 128 + row] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + row]) * kLog2e;
    auto block = [&](const int j, float* __restrict__ s, float* __restrict__ sn) {
      const bool more = j + 1 < nblk;
      constexpr int PB = Cfg::p_bufs;
      if (warp == kAttnFirstSoftmaxWarp) TR(0);
      float alpha = 1.f;
      const bool grow = mx > m_used + kLazyThreshold;
      if (grow) {
        alpha = exp2f(m_used - mx);  // 0 on the first block (m_used = -inf)
        m_used = mx;
      }
      // P -> bf16.  Tensor-memory P: this thread's row, 32-bit column = two keys; the pieces go out as soon
      // as they are packed (the tile's previous reader is PV_{j-2}, long done).  Shared-memory P (dh = 256):
      // K-major SWIZZLE_128B slab `half` of [128 rows][64 keys], written after PV_{j-1} has drained.
      uint8_t* p_row = p_smem + half * (kBQ * 128) + row * 128;
      const uint32_t p_taddr = lane_addr + Cfg::p_col + (uint32_t)((j & 1) * (kBK / 2) + half * (HB / 2));
      auto p_store = [&](int c8, const uint32_t* w) {   // c8: group of 8 keys = 4 packed words
        *reinterpret_cast<uint4*>(p_row + ((c8 ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      };
      auto p_store16 = [&](int g, const uint32_t* w) {  // g: group of 16 keys = 8 packed words
        if (PB == 2) tmem_st_32x32b_x8(p_taddr + (uint32_t)g * 8, w);
      };
      if (PB == 2 && j >= 2) mbar_wait(&o_done[j & 1], (uint32_t)(((j - 2) >> 1) & 1));
      // Four groups of 16 exponentials.  The two warps of a row pair share a scheduler and run in lockstep, so
      // the MUFU pipe is only kept busy if every exponential is followed by its share of the other work
      // (row sum and bf16 packing of the previous group, the row maximum of the next block): written
      // interleaved here, one slice per ex2.
      constexpr int G = HB / 4;
      uint32_t pk[HB / 2];
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < G; ++i) s[i] = ex2_pinned(fmaf(s[i], kLog2e, -m_used));
#pragma unroll
      for (int i = 0; i < G; ++i) {
        s[G + i] = ex2_pinned(fmaf(s[G + i], kLog2e, -m_used));
        if (i & 1) { (i & 2 ? sum3 : sum1) += s[i]; pk[i >> 1] = pack_bf16x2(s[i - 1], s[i]); } else (i & 2 ? sum2 : sum0) += s[i];
      }
      p_store16(0, pk);
      if (warp == kAttnFirstSoftmaxWarp) TR(1);
      if (more) {      // S_{j+1} was issued when P_{j-1} was handed over: complete about now
        mbar_wait(&s_full[(j + 1) & 1], (uint32_t)(((j + 1) >> 1) & 1));
        tc_fence_after();
        ld_scores(j + 1, sn);
      }
      if (warp == kAttnFirstSoftmaxWarp) TR(2);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        s[2 * G + i] = ex2_pinned(fmaf(s[2 * G + i], kLog2e, -m_used));
        if (i & 1) { (i & 2 ? sum3 : sum1) += s[G + i]; pk[(G + i) >> 1] = pack_bf16x2(s[G + i - 1], s[G + i]); } else (i & 2 ? sum2 : sum0) += s[G + i];
      }
      p_store16(1, pk + 8);
      if (more) tmem_wait_ld();
      if (warp == kAttnFirstSoftmaxWarp) TR(3);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        s[3 * G + i] = ex2_pinned(fmaf(s[3 * G + i], kLog2e, -m_used));
        if (i & 1) { (i & 2 ? sum3 : sum1) += s[2 * G + i]; pk[(2 * G + i) >> 1] = pack_bf16x2(s[2 * G + i - 1], s[2 * G + i]); } else (i & 2 ? sum2 : sum0) += s[2 * G + i];
        if (more) {
          m0 = fmaxf(m0, sn[4 * i]); m1 = fmaxf(m1, sn[4 * i + 1]); m2 = fmaxf(m2, sn[4 * i + 2]); m3 = fmaxf(m3, sn[4 * i + 3]);
        }
      }
      float mxn = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      if (more) xch[(((j + 1) & 1) * 2 + half) * 128 + row] = mxn;
      p_store16(2, pk + 16);
#pragma unroll
      for (int i = 0; i < G; ++i) {
        if (i & 1) { (i & 2 ? sum3 : sum1) += s[3 * G + i]; pk[(3 * G + i) >> 1] = pack_bf16x2(s[3 * G + i - 1], s[3 * G + i]); } else (i & 2 ? sum2 : sum0) += s[3 * G + i];
      }
      p_store16(3, pk + 24);
      if (more) {
        asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
        mxn = fmaxf(mxn, xch[(((j + 1) & 1) * 2 + (half ^ 1)) * 128 + row]) * kLog2e;
      }
      l = fmaf(l, alpha, (sum0 + sum1) + (sum2 + sum3));
      if (warp == kAttnFirstSoftmaxWarp) TR(4);
      // PV_{j-1} must be done before O is rescaled (and, with one P tile, before P is overwritten)
      if (j > 0 && (PB == 1 || __any_sync(0xffffffffu, grow))) {
        mbar_wait(&o_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));
        tc_fence_after();
        if (warp == kAttnFirstSoftmaxWarp) TR(5);
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
      if (PB == 1) {
#pragma unroll
        for (int c8 = 0; c8 < HB / 8; ++c8) p_store(c8, pk + 4 * c8);
      }
      if (warp == kAttnFirstSoftmaxWarp) TR(6);
      if (PB == 2) tmem_wait_st();
      else fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == kAttnFirstSoftmaxWarp) TR(7);
      mx = mxn;
    };
    // two blocks per trip so the score registers of block j+1 become "current" without being copied
    for (int j = 0; j < nblk; j += 2) {
      block(j, sa, sb);
      if (j + 1 < nblk) block(j + 1, sb, sa);
    }
    // ---- output: O / l (row sum = both halves) ----
    asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");   // partner has finished reading the last max
    xch[half * 128 + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(2 + qd) : "memory");
    l += xch[(half ^ 1) * 128 + row];
    mbar_wait(&o_done[(nblk - 1) & 1], (uint32_t)(((nblk - 1) >> 1) & 1));
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
