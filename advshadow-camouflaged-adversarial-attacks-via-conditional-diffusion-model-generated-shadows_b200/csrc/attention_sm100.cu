// K4: fused self-attention softmax(q^T k) v on the 5th-gen tensor cores (flash-style, no T x T matrix).
//   reference: AttentionBlock.forward dm1:120-125 (two einsums around an fp32 softmax that
//   materialises [B*heads, T, T]).
//
// One CTA per (image*head, 128-query tile); keys/values stream through in blocks of 128:
//   warp 0     : TMA producer for Q (once) and the K blocks
//   warp 1     : tcgen05.mma issuer for S_j = Q K_j^T  (TMEM, two buffers)
//   warp 3     : tcgen05.mma issuer for O += P_j V_j   (TMEM)
//   warp 2     : TMA producer for the V^T blocks (a V stage is released much later than the K stage of the
//                same block; one thread issuing both in order would hold the next S behind it)
//   warps 4-7  : online softmax of the EVEN key blocks, one query row per thread
//   warps 8-11 : online softmax of the ODD key blocks
// The softmax is bound by the MUFU pipe (one ex2 per score, 16 / clk / SM).  A warp and its twin of the other
// set share a scheduler and the same 32 rows but run independently: while one issues its 128 exponentials
// the other can wait for its scores, take the row maximum, pack and hand P over, so the MUFU pipe does
// not sit idle behind that latency.  (Forcing strict alternation with a token between the twins measured the
// same.)  What couples the two is one float per row, the running maximum
// m (base 2) that the shared O accumulator is scaled by: the warp of block j reads the m left by block j-1,
// raises it only if the row maximum grew by more than 2^8 (lazy rescale: then it also rescales O in TMEM
// after PV_{j-1} has drained), and publishes it for block j+1.  Each warp sums its own blocks' row sums and
// rescales them whenever it sees m has moved; they are added at the end.
// dh <= 128: P_j is written to tensor memory (64 columns, one tile per set) and is the A operand of the PV
// MMA from there, so PV only streams V from shared memory.  dh = 256 fills TMEM with S and O; its P goes
// through one swizzled shared-memory tile.
// q and k arrive pre-scaled by dh^-1/4 each (dm1:121-122), so the softmax scale is 1.
// T need not be a multiple of 128: the last key block is partial.  Its out-of-range score columns are set to -inf
// before the row maximum (they hold the next head's keys, or TMA zero fill past the end of the tensor), the V^T box
// is zero-filled by TMA past column T (so 0 * garbage never occurs), and query rows past T are computed but not
// stored.  Head dims below 64 (IDDM: 16 and 32) run as dh = 64 with zero-padded q / k / v^T; only the first
// dh_valid output columns of each head are stored, packed.
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
  int dh_valid;   // head dim of the problem; q / k / v^T are zero-padded to dh (= the kernel's DH) when smaller
  __nv_bfloat16* o;
};

struct AttnPlan {
  AttnMaps maps;
  AttnArgs args;
  uint32_t smem_bytes;
  uint32_t magic;
};
static_assert(sizeof(AttnPlan) <= ADVS_ATTN_PLAN_BYTES, "AttnPlan does not fit ADVS_ATTN_PLAN_BYTES");

// Three warpgroups: {K producer, S issuer, V producer, PV issuer} + two softmax sets of four warps.  Registers are
// allocated to warps in groups of four, so 12 warps start at 168 each; the first group hands most of its
// share to the softmax groups (setmaxnreg), which hold a whole block of scores per thread.
constexpr int kAttnThreads = 384;
constexpr int kAttnFirstSoftmaxWarp = 4;
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
  static constexpr bool p_in_tmem = DH <= 128;
  static constexpr uint32_t q_bytes = kBQ * DH * 2;
  static constexpr uint32_t k_bytes = kBK * DH * 2;
  static constexpr uint32_t v_bytes = DH * kBK * 2;
  static constexpr uint32_t p_bytes = kBQ * kBK * 2;
  static constexpr uint32_t off_q = 0;
  static constexpr uint32_t off_k = off_q + q_bytes;
  static constexpr uint32_t off_v = off_k + kv_stages * k_bytes;
  static constexpr uint32_t off_p = off_v + kv_stages * v_bytes;
  static constexpr uint32_t off_bar = off_p + (p_in_tmem ? 0 : p_bytes);
  static constexpr uint32_t bar_bytes = 256;                                   // 25 mbarriers + the TMEM slot
  static constexpr uint32_t xch_bytes = 3 * 128 * 4;                           // m[128] + row sums of both sets
  static constexpr uint32_t smem_bytes = off_bar + bar_bytes + xch_bytes + 1024;   // + alignment slack
  static constexpr uint32_t tmem_cols = 512;
  static constexpr uint32_t o_col = 256;     // S0 | S1 | O (dh columns) | P0 | P1 (64 columns each, dh <= 128)
  static constexpr uint32_t p_col = 384;
};

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_sm100(const __grid_constant__ AttnMaps maps, const AttnArgs a) {
  using Cfg = AttnCfg<DH>;
  constexpr int KVS = Cfg::kv_stages;
  constexpr bool PT = Cfg::p_in_tmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::off_bar);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // KVS
  uint64_t* k_empty = bars + 3;       // KVS
  uint64_t* v_full = bars + 5;        // KVS
  uint64_t* v_empty = bars + 7;       // KVS
  uint64_t* s_full = bars + 9;        // 2: S_j complete in buffer j & 1
  uint64_t* s_free = bars + 11;       // 2: the softmax warps hold S_j in registers, the buffer may take S_{j+2}
  uint64_t* p_full = bars + 13;       // 2: P_j written (and O rescaled if needed)
  uint64_t* o_done = bars + 15;       // 2: PV_j complete
  uint64_t* m_posted = bars + 17;     // [4 row quarters][2]: block j's warp has published the running max
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
  float* m_sh = reinterpret_cast<float*>(smem + Cfg::off_bar + Cfg::bar_bytes);   // [128] running max (base 2)
  float* l_sh = m_sh + 128;                                                       // [2 sets][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int nblk = (a.T + kBK - 1) / kBK;
  const int tail = a.T - (nblk - 1) * kBK;       // valid keys in the last block (kBK when T is a multiple of it)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    mbar_init(q_full, 1);
    for (int s = 0; s < KVS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 4);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_done[s], 1);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&m_posted[s], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::tmem_cols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp < kAttnFirstSoftmaxWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ================= TMA producer: Q, then the K blocks =================
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
          if (++st == KVS) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 2) {
      // ================= TMA producer: the V^T blocks =================
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
      // ================= MMA issuer for S_j = Q K_j^T (whole warp walks the schedule, one elected lane issues) ====
      // tcgen05.mma issue blocks while the tensor pipe's queue is full, so an issuing warp spends its time
      // either waiting for a barrier or inside the issue; with S and PV on two warps one's barrier waits
      // hide behind the other's MMAs.  The two streams are ordered only through the softmax warps' barriers.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kBK);
      const uint32_t q_addr = smem_u32(smem + Cfg::off_q);
      int ks = 0;
      uint32_t kph = 0;
      mbar_wait(q_full, 0);
      for (int j = 0; j < nblk; ++j) {
        TR(8);
        mbar_wait(&k_full[ks], kph);
        if (j >= 2) mbar_wait(&s_free[j & 1], (uint32_t)(((j - 2) >> 1) & 1));   // block j-2's scores are in registers
        tc_fence_after();
        TR(9);
        const uint32_t k_addr = smem_u32(smem + Cfg::off_k + ks * Cfg::k_bytes);
        const uint32_t d = tmem_base + (uint32_t)((j & 1) * kBK);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k) {
            const uint32_t off = (uint32_t)(k >> 2) * (128 * 128) + (uint32_t)(k & 3) * 32;
            umma_bf16(d, umma_desc_k_sw128(q_addr + off), umma_desc_k_sw128(k_addr + off), idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit(&s_full[j & 1]);
          umma_commit(&k_empty[ks]);
        }
        __syncwarp();
        TR(10);
        if (++ks == KVS) { ks = 0; kph ^= 1; }
      }
    } else {
      // ================= MMA issuer for O += P_j V_j (warp 3) =================
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, DH);
      const uint32_t p_addr = smem_u32(smem + Cfg::off_p);
      int vs = 0;
      uint32_t vph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&v_full[vs], vph);
        mbar_wait(&p_full[j & 1], (uint32_t)((j >> 1) & 1));
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + Cfg::off_v + vs * Cfg::v_bytes);
        const uint32_t d = tmem_base + Cfg::o_col;
        const uint32_t p_tmem = tmem_base + Cfg::p_col + (uint32_t)(j & 1) * (kBK / 2);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint32_t offp = (uint32_t)(k >> 2) * (kBQ * 128) + (uint32_t)(k & 3) * 32;
            const uint32_t offv = (uint32_t)(k >> 2) * (DH * 128) + (uint32_t)(k & 3) * 32;
            if (PT)
              umma_bf16_ts(d, p_tmem + (uint32_t)k * 8, umma_desc_k_sw128(v_addr + offv), idesc_o, (j | k) != 0 ? 1u : 0u);
            else
              umma_bf16(d, umma_desc_k_sw128(p_addr + offp), umma_desc_k_sw128(v_addr + offv), idesc_o,
                        (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&o_done[j & 1]);   // the softmax warps may touch P / O again
          umma_commit(&v_empty[vs]);
        }
        __syncwarp();
        if (++vs == KVS) { vs = 0; vph ^= 1; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ================= softmax / correction / output (warps 4..11) =================
    const int qd = warp & 3;                                   // TMEM lane quarter = 32 query rows
    const int set = (warp - kAttnFirstSoftmaxWarp) >> 2;       // 0: even key blocks, 1: odd key blocks
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    uint8_t* p_row = smem + Cfg::off_p + row * 128;
    float m_last = -INFINITY;     // the running max this warp's row sum is currently scaled by
    float l = 0.f;                // row sum over this set's blocks
    for (int j = set; j < nblk; j += 2) {
      if (warp == kAttnFirstSoftmaxWarp) TR(0);
      mbar_wait(&s_full[set], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      float s[kBK];
#pragma unroll
      for (int c = 0; c < kBK / 32; ++c)
        tmem_ld_32x32b_x32(lane_addr + (uint32_t)(set * kBK + c * 32), reinterpret_cast<uint32_t*>(s) + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[set]);      // S_{j+2} may overwrite the buffer
      if (j == nblk - 1 && tail < kBK) {             // partial last key block (warp-uniform branch)
#pragma unroll
        for (int i = 0; i < kBK; ++i)
          if (i >= tail) s[i] = -INFINITY;
      }
      if (warp == kAttnFirstSoftmaxWarp) TR(1);
      float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
      for (int i = 4; i < kBK; i += 4) {
        m0 = fmaxf(m0, s[i]); m1 = fmaxf(m1, s[i + 1]); m2 = fmaxf(m2, s[i + 2]); m3 = fmaxf(m3, s[i + 3]);
      }
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * kLog2e;
      // running max: take over what block j-1 left, raise it if this block needs it, publish for block j+1
      float m_used = -INFINITY;
      if (j > 0) {
        mbar_wait(&m_posted[qd * 2 + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
        m_used = *reinterpret_cast<volatile float*>(&m_sh[row]);
      }
      if (m_used != m_last) l *= exp2f(m_last - m_used);   // the other set raised it (or first own block: l = 0)
      float alpha = 1.f;
      const bool grow = mx > m_used + kLazyThreshold;
      if (grow) {
        alpha = exp2f(m_used - mx);   // 0 on the first block (m_used = -inf)
        m_used = mx;
        l *= alpha;
      }
      m_last = m_used;
      *reinterpret_cast<volatile float*>(&m_sh[row]) = m_used;
      __syncwarp();
      if (lane == 0) mbar_arrive(&m_posted[qd * 2 + (j & 1)]);
      if (warp == kAttnFirstSoftmaxWarp) TR(2);
      // the P tile's previous reader is PV_{j-2}
      if (PT && j >= 2) mbar_wait(&o_done[set], (uint32_t)(((j - 2) >> 1) & 1));
      const uint32_t p_taddr = lane_addr + Cfg::p_col + (uint32_t)(set * (kBK / 2));
      uint32_t pk[PT ? 8 : kBK / 2];
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int g = 0; g < kBK / 16; ++g) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s[g * 16 + i] = fast_exp2(fmaf(s[g * 16 + i], kLog2e, -m_used));
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          sum0 += s[g * 16 + i]; sum1 += s[g * 16 + i + 1]; sum2 += s[g * 16 + i + 2]; sum3 += s[g * 16 + i + 3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[(PT ? 0 : g * 8) + i] = pack_bf16x2(s[g * 16 + 2 * i], s[g * 16 + 2 * i + 1]);
        if (PT) tmem_st_32x32b_x8(p_taddr + (uint32_t)g * 8, pk);   // this row, keys 16g .. 16g+15
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      if (warp == kAttnFirstSoftmaxWarp) TR(3);
      // PV_{j-1} must have drained before O is rescaled (and, with the shared-memory tile, before P is overwritten)
      const bool any_grow = __any_sync(0xffffffffu, grow);
      if (j > 0 && (!PT || any_grow)) {
        mbar_wait(&o_done[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1));
        tc_fence_after();
        if (any_grow) {
#pragma unroll 1
          for (int c = 0; c < DH / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(lane_addr + Cfg::o_col + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_32x32b_x32(lane_addr + Cfg::o_col + c * 32, r);
          }
        }
      }
      if (warp == kAttnFirstSoftmaxWarp) TR(4);
      if (PT) {
        tmem_wait_st();
      } else {
        // K-major SWIZZLE_128B, two slabs of [128 rows][64 keys]
#pragma unroll
        for (int c8 = 0; c8 < kBK / 8; ++c8)
          *reinterpret_cast<uint4*>(p_row + (c8 >> 3) * (kBQ * 128) + (((c8 & 7) ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
        tmem_wait_st();
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[set]);
      if (warp == kAttnFirstSoftmaxWarp) TR(5);
    }
    // ---- output: O / l.  Both sets have published their last max; bring both row sums to it and add ----
    if (nblk > 0) {
      const int jl = nblk - 1;
      mbar_wait(&m_posted[qd * 2 + (jl & 1)], (uint32_t)((jl >> 1) & 1));
      const float m_final = *reinterpret_cast<volatile float*>(&m_sh[row]);
      if (m_final != m_last) l *= exp2f(m_last - m_final);
    }
    l_sh[set * 128 + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(10 + qd) : "memory");
    l += l_sh[(set ^ 1) * 128 + row];
    mbar_wait(&o_done[(nblk - 1) & 1], (uint32_t)(((nblk - 1) >> 1) & 1));
    tc_fence_after();
    const float inv = 1.f / l;
    constexpr int HD = DH / 2;    // O columns per warp: the two sets split the columns of their 32 rows
    const int b = bh / a.heads, head = bh - b * a.heads;
    const int dv = a.dh_valid;
    const bool row_ok = q_tile * kBQ + row < a.T;
    __nv_bfloat16* orow = a.o + ((size_t)b * a.T + q_tile * kBQ + row) * ((size_t)a.heads * dv) + (size_t)head * dv + set * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      if (set * HD + c * 32 >= dv) break;            // padded head dim: nothing to store (warp-uniform)
      uint32_t r[32];
      tmem_ld_32x32b_x32(lane_addr + Cfg::o_col + set * HD + c * 32, r);
      tmem_wait_ld();
      uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (row_ok && set * HD + c * 32 + 8 * i < dv)
        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv),
                            pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv),
                            pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv),
                            pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::tmem_cols>(tmem_base);
}

template <int DH>
static int attn_launch(const AttnPlan* plan, cudaStream_t st) {
  constexpr int slot = DH == 64 ? kOnceAttn64 : (DH == 128 ? kOnceAttn128 : kOnceAttn256);
  if (first_use_on_device(slot)) {
    cudaError_t e = cudaFuncSetAttribute(k_attention_sm100<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         AttnCfg<DH>::smem_bytes);
    if (e != cudaSuccess) {
      forget_first_use(slot);
      set_error("attention_sm100_launch: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ADVS_ERR_CUDA;
    }
  }
  dim3 grid((plan->args.T + kBQ - 1) / kBQ, plan->args.B * plan->args.heads);
  launch_pdl(k_attention_sm100<DH>, grid, dim3(kAttnThreads), AttnCfg<DH>::smem_bytes, st, plan->maps, plan->args);
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

int advs_attention_sm100_plan_ex(const void* q, const void* k, const void* vt, void* o, int B, int heads, int T, int dh,
                                 int dh_valid, void* plan_host) {
  ADVS_CHECK_ARG(q && k && vt && o && plan_host, "attention_sm100_plan: null pointer");
  ADVS_CHECK_ARG(((uintptr_t)plan_host % 64) == 0, "attention_sm100_plan: plan buffer must be 64-byte aligned");
  ADVS_CHECK_ARG(B > 0 && heads > 0 && T > 0 && T % 8 == 0,
                 "attention_sm100_plan: T must be a positive multiple of 8 (TMA row pitch of v^T: 16 bytes)");
  ADVS_CHECK_ARG(dh == 64 || dh == 128 || dh == 256, "attention_sm100_plan: dh must be 64, 128 or 256");
  ADVS_CHECK_ARG(dh_valid > 0 && dh_valid <= dh && dh_valid % 8 == 0 && (dh_valid == dh || dh == 64),
                 "attention_sm100_plan: dh_valid must be a multiple of 8, <= dh, and < dh only with dh = 64");
  ADVS_CHECK_ARG(((uintptr_t)o % 16) == 0 && (heads * dh_valid) % 8 == 0, "attention_sm100_plan: output rows must be 16-byte aligned");
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
  plan->args.dh_valid = dh_valid;
  plan->args.o = reinterpret_cast<__nv_bfloat16*>(o);
  plan->magic = 0xA77EB200u;
  return ADVS_OK;
}

int advs_attention_sm100_plan(const void* q, const void* k, const void* vt, void* o, int B, int heads, int T, int dh,
                              void* plan_host) {
  return advs_attention_sm100_plan_ex(q, k, vt, o, B, heads, T, dh, dh, plan_host);
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
