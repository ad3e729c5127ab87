// K5: GroupNorm(G, C) [+ SiLU] over NHWC activations, optionally over the virtual concat of two
// sources (reference: norm_layer dm1:62-63 used at dm1:71-72, 83-84, 113, 240-241; the concat is
// torch.cat([h, hs.pop()], 1) at dm1:265).
//
// Three launches, all bandwidth-bound and deterministic (no float atomics):
//   1. k_gn_partial : per (image, pixel-chunk, channel) sum / sum-of-squares        reads x once
//   2. k_gn_finalize: per (image, group) mean/rstd in fp64 -> per-channel scale/shift (tiny)
//   3. k_gn_apply   : y = act(x * scale + shift)                                    reads x, writes y
// Algorithmic bytes: 2 reads + 1 write per element (6 B/elem bf16, 12 B/elem fp32).
#include "common.cuh"

namespace advs {

struct GnGeom {
  int chunks;          // pixel chunks per image
  int pix_per_chunk;
};

static GnGeom gn_geom(int B, int HW) {
  // aim for >= 16 CTAs per SM over the whole launch, chunks of at least 64 pixels
  int want = (148 * 16 + B - 1) / B;
  int max_chunks = (HW + 63) / 64;
  int chunks = want < max_chunks ? want : max_chunks;
  if (chunks < 1) chunks = 1;
  GnGeom g;
  g.pix_per_chunk = (HW + chunks - 1) / chunks;
  g.chunks = (HW + g.pix_per_chunk - 1) / g.pix_per_chunk;
  return g;
}

// Geometry of the statistics pass.  It depends on the image size ONLY: the fp32 partial sums are then formed in the
// same order whatever the batch size, so an image's result does not depend on which batch (or sub-batch of a
// multi-stream sampler) it is processed in -- trajectories are bit-reproducible across batch sizes.
static GnGeom gn_geom_stats(int HW) {
  int chunks = (HW + 1023) / 1024;          // ~1024 pixels per chunk, at most 64 chunks per image
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  GnGeom g;
  g.pix_per_chunk = (HW + chunks - 1) / chunks;
  g.chunks = (HW + g.pix_per_chunk - 1) / g.pix_per_chunk;
  return g;
}

// part layout: [B][chunks][Ctot][2]
template <typename T>
__global__ void k_gn_partial(const T* __restrict__ x, int C, int c_off, int Ctot, int HW, int pix_per_chunk,
                             float* __restrict__ part) {
  extern __shared__ float red[];  // [ppi][C]
  pdl_sync();
  const int cv = C / 8;
  const int ppi = blockDim.x / cv;  // pixels per iteration (>= 1 by host check)
  const int pl = threadIdx.x / cv;
  const int j = threadIdx.x % cv;
  const bool active = pl < ppi;
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int p0 = chunk * pix_per_chunk;
  int p1 = p0 + pix_per_chunk;
  if (p1 > HW) p1 = HW;

  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  if (active) {
    const T* base = x + ((size_t)b * HW) * C + j * 8;
    int p = p0 + pl;
    for (; p + 3 * ppi < p1; p += 4 * ppi) {   // 4 independent 16-byte loads in flight per thread
      Vec8<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u].load(base + (size_t)(p + u * ppi) * C);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        v[u].to_float(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s[i] += f[i];
          ss[i] = fmaf(f[i], f[i], ss[i]);
        }
      }
    }
    for (; p < p1; p += ppi) {
      Vec8<T> v;
      v.load(base + (size_t)p * C);
      float f[8];
      v.to_float(f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        ss[i] = fmaf(f[i], f[i], ss[i]);
      }
    }
  }
  float* out = part + (((size_t)b * gridDim.x + chunk) * Ctot + c_off) * 2;
  // pass 1: sums
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * C + j * 8 + i] = s[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < ppi; ++q) t += red[q * C + c];
    out[c * 2] = t;
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pl * C + j * 8 + i] = ss[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < ppi; ++q) t += red[q * C + c];
    out[c * 2 + 1] = t;
  }
}

// one 128-thread CTA per (image, group); the group's channels may straddle the two concatenated sources.
// part_s layout: [B][parts_s][C_s][2].  Threads sweep (part, channel-in-group) pairs with the channel
// fastest, so each warp reads contiguous float2 runs; fp64 accumulation, fixed reduction order.
__global__ void __launch_bounds__(128) k_gn_finalize(const float* __restrict__ part0, int c0, int parts0, int gran0,
                                                     const float* __restrict__ part1, int c1, int parts1, int gran1, int groups,
                                                     int HW, float eps, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ scale_shift) {
  __shared__ double red[2][4];
  pdl_sync();
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int Ctot = c0 + c1;
  const int cpg = Ctot / groups;
  const int glo = g * cpg, ghi = glo + cpg;
  double S = 0.0, SS = 0.0;
#pragma unroll
  for (int src = 0; src < 2; ++src) {
    const float* base = src == 0 ? part0 : part1;
    const int Cs = src == 0 ? c0 : c1, parts = src == 0 ? parts0 : parts1, off = src == 0 ? 0 : c0;
    const int gran = src == 0 ? gran0 : gran1;      // channels per partial slot (1 or 4; slots never straddle a group)
    int lo = glo > off ? glo : off;
    int hi = ghi < off + Cs ? ghi : off + Cs;
    const int n = (hi - lo) / gran;
    if (hi <= lo || parts <= 0) continue;
    const int C = Cs / gran;
    const float* p = base + ((size_t)b * parts * C + (lo - off) / gran) * 2;
    for (int it = threadIdx.x; it < parts * n; it += 128) {
      const int pi = it / n, cc = it - pi * n;
      const float2 v = *reinterpret_cast<const float2*>(p + ((size_t)pi * C + cc) * 2);
      S += (double)v.x;
      SS += (double)v.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, o);
    SS += __shfl_xor_sync(0xffffffffu, SS, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = S;
    red[1][threadIdx.x >> 5] = SS;
  }
  __syncthreads();
  S = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
  SS = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  const double n = (double)HW * cpg;
  const double mean = S / n;
  double var = SS / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float meanf = (float)mean;
  for (int c = glo + threadIdx.x; c < ghi; c += 128) {
    const float sc = rstd * gamma[c];
    const float sh = beta[c] - meanf * sc;
    *reinterpret_cast<float2*>(scale_shift + ((size_t)b * Ctot + c) * 2) = make_float2(sc, sh);
  }
}

// thread = one 8-channel vector position, looping over the pixels of its chunk: the 16 scale/shift
// floats stay in registers, so the stream is exactly one 16-byte load + one 16-byte store per vector.
template <typename T, bool SILU, bool WIDE = false, bool OUT_F16 = false>
__global__ void k_gn_apply(const T* __restrict__ x0, int c0, const T* __restrict__ x1, int c1, int HW,
                           int pix_per_chunk, const float* __restrict__ scale_shift, T* __restrict__ y,
                           const uint8_t* __restrict__ lo0 = nullptr, const uint8_t* __restrict__ lo1 = nullptr) {
  pdl_sync();
  const int Ctot = c0 + c1;
  const int cv = Ctot / 8;
  const int ppi = blockDim.x / cv;
  const int pl = threadIdx.x / cv;
  const int j = threadIdx.x % cv;
  if (pl >= ppi) return;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_chunk;
  int p1 = p0 + pix_per_chunk;
  if (p1 > HW) p1 = HW;
  const int c = j * 8;
  float sc[8], sh[8];
  {
    const float4* ab = reinterpret_cast<const float4*>(scale_shift + ((size_t)b * Ctot + c) * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = __ldg(ab + i);
      sc[2 * i] = t.x; sh[2 * i] = t.y; sc[2 * i + 1] = t.z; sh[2 * i + 1] = t.w;
    }
  }
  const bool first = c < c0;
  const T* src = first ? x0 + ((size_t)b * HW) * c0 + c : x1 + ((size_t)b * HW) * c1 + (c - c0);
  const int cs = first ? c0 : c1;
  T* dst = y + ((size_t)b * HW) * Ctot + c;
  // "wide" sources carry 8 more mantissa bits per element in a companion int8 tensor (common.cuh)
  const uint8_t* lsrc = nullptr;
  if constexpr (WIDE) {
    const uint8_t* l = first ? lo0 : lo1;
    if (l) lsrc = first ? l + ((size_t)b * HW) * c0 + c : l + ((size_t)b * HW) * c1 + (c - c0);
  }
  auto load_lo = [&](int p) {
    uint2 r = make_uint2(0u, 0u);
    if constexpr (WIDE) {
      if (lsrc)
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(lsrc + (size_t)p * cs));
    }
    return r;
  };
  auto body = [&](const Vec8<T>& vin, uint2 lo, int p) {
    float f[8];
    if constexpr (WIDE && sizeof(T) == 2) wide_decode8(vin.v, lo, f);
    else vin.to_float(f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = fmaf(f[i], sc[i], sh[i]);
      if (SILU) {
        if constexpr (sizeof(T) == 4) a = silu_acc(a);
        else a = silu_f(a);
      }
      f[i] = a;
    }
    Vec8<T> vo;
    if constexpr (OUT_F16 && sizeof(T) == 2) {   // fp16 GEMM operand: same 16 bits, three more of them mantissa
      __half2* h = reinterpret_cast<__half2*>(&vo.v);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    } else {
      vo.from_float(f);
    }
    vo.store_stream(dst + (size_t)p * Ctot);
  };
  int p = p0 + pl;
  for (; p + 7 * ppi < p1; p += 8 * ppi) {   // 8 independent 16-byte loads in flight per thread
    Vec8<T> v[8];
    uint2 l[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v[u].load_stream(src + (size_t)(p + u * ppi) * cs);
      l[u] = load_lo(p + u * ppi);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) body(v[u], l[u], p + u * ppi);
  }
  for (; p < p1; p += ppi) {
    Vec8<T> v;
    v.load_stream(src + (size_t)p * cs);
    body(v, load_lo(p), p);
  }
}

template <typename T>
static int gn_partial_impl(const void* x, int C, int B, int HW, float* part, cudaStream_t st) {
  GnGeom g = gn_geom_stats(HW);
  int cv = C / 8;
  int threads = cv <= 256 ? 256 : 1024;
  int ppi = threads / cv;
  size_t smem = (size_t)ppi * C * sizeof(float);
  dim3 grid(g.chunks, B);
  launch_pdl(k_gn_partial<T>, grid, dim3(threads), smem, st, (const T*)x, C, 0, C, HW, g.pix_per_chunk, part);
  ADVS_CHECK_LAUNCH("groupnorm_partial");
  return ADVS_OK;
}

static int gn_finalize_impl(const float* part0, int c0, int parts0, int gran0, const float* part1, int c1, int parts1,
                            int gran1, int B, int HW, int groups, float eps, const float* gamma, const float* beta,
                            float* scale_shift, cudaStream_t st) {
  launch_pdl(k_gn_finalize, dim3(B * groups), dim3(128), 0, st, part0, c0, parts0, gran0, part1, c1, parts1, gran1, groups, HW, eps,
             gamma, beta, scale_shift);
  ADVS_CHECK_LAUNCH("groupnorm_finalize");
  return ADVS_OK;
}

template <typename T>
static int gn_stats_impl(const void* x0, int c0, const void* x1, int c1, int B, int HW, int groups, float eps,
                         const float* gamma, const float* beta, float* scale_shift, void* ws, cudaStream_t st) {
  GnGeom g = gn_geom_stats(HW);
  float* p0 = (float*)ws;
  float* p1 = p0 + (size_t)B * g.chunks * c0 * 2;
  int rc = gn_partial_impl<T>(x0, c0, B, HW, p0, st);
  if (rc) return rc;
  if (c1) {
    rc = gn_partial_impl<T>(x1, c1, B, HW, p1, st);
    if (rc) return rc;
  }
  return gn_finalize_impl(p0, c0, g.chunks, 1, c1 ? p1 : nullptr, c1, c1 ? g.chunks : 0, 1, B, HW, groups, eps, gamma, beta,
                          scale_shift, st);
}

template <typename T>
static int gn_apply_impl(const void* x0, int c0, const void* x1, int c1, int B, int HW, const float* ss, int silu,
                         void* y, cudaStream_t st) {
  GnGeom g = gn_geom(B, HW);
  int cv = (c0 + c1) / 8;
  int threads = cv <= 256 ? 256 : 1024;
  dim3 grid(g.chunks, B);
  if (silu)
    launch_pdl(k_gn_apply<T, true, false, false>, grid, dim3(threads), 0, st, (const T*)x0, c0, (const T*)x1, c1, HW, g.pix_per_chunk,
               ss, (T*)y, (const uint8_t*)nullptr, (const uint8_t*)nullptr);
  else
    launch_pdl(k_gn_apply<T, false, false, false>, grid, dim3(threads), 0, st, (const T*)x0, c0, (const T*)x1, c1, HW, g.pix_per_chunk,
               ss, (T*)y, (const uint8_t*)nullptr, (const uint8_t*)nullptr);
  ADVS_CHECK_LAUNCH("groupnorm_apply");
  return ADVS_OK;
}

static int gn_apply_wide_impl(const void* x0, const void* lo0, int c0, const void* x1, const void* lo1, int c1, int B, int HW,
                              const float* ss, int silu, void* y, bool out_f16, cudaStream_t st) {
  using T = __nv_bfloat16;
  GnGeom g = gn_geom(B, HW);
  int cv = (c0 + c1) / 8;
  int threads = cv <= 256 ? 256 : 1024;
  dim3 grid(g.chunks, B);
  const bool wide = lo0 || lo1;
#define ADVS_GN(S, W_, F) launch_pdl(k_gn_apply<T, S, W_, F>, grid, dim3(threads), 0, st, (const T*)x0, c0, (const T*)x1, c1, HW, \
                                     g.pix_per_chunk, ss, (T*)y, (const uint8_t*)lo0, (const uint8_t*)lo1)
  if (silu) {
    if (wide) { if (out_f16) ADVS_GN(true, true, true); else ADVS_GN(true, true, false); }
    else { if (out_f16) ADVS_GN(true, false, true); else ADVS_GN(true, false, false); }
  } else {
    if (wide) { if (out_f16) ADVS_GN(false, true, true); else ADVS_GN(false, true, false); }
    else { if (out_f16) ADVS_GN(false, false, true); else ADVS_GN(false, false, false); }
  }
#undef ADVS_GN
  ADVS_CHECK_LAUNCH("groupnorm_apply_wide");
  return ADVS_OK;
}

}  // namespace advs

using namespace advs;

extern "C" {

size_t advs_groupnorm_workspace_bytes(int B, int HW, int C) {
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  GnGeom g = gn_geom_stats(HW);
  return (size_t)B * g.chunks * C * 2 * sizeof(float);
}

int advs_groupnorm_stats(const void* x0, int c0, const void* x1, int c1, int B, int HW, int groups, float eps,
                         const float* gamma, const float* beta, float* scale_shift, void* workspace,
                         size_t workspace_bytes, int dtype, void* stream) {
  ADVS_CHECK_ARG(x0 && c0 > 0 && B > 0 && HW > 0 && groups > 0, "groupnorm_stats: bad args");
  if (!x1) c1 = 0;
  ADVS_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm_stats: channel counts must be multiples of 8");
  ADVS_CHECK_ARG((c0 + c1) % groups == 0, "groupnorm_stats: channels not divisible by groups");
  ADVS_CHECK_ARG(c0 / 8 <= 1024 && c1 / 8 <= 1024, "groupnorm_stats: at most 8192 channels per source");
  ADVS_CHECK_ARG(gamma && beta && scale_shift && workspace, "groupnorm_stats: null pointer");
  ADVS_CHECK_ARG(workspace_bytes >= advs_groupnorm_workspace_bytes(B, HW, c0 + c1), "groupnorm_stats: workspace too small");
  if (dtype == ADVS_F32)
    return gn_stats_impl<float>(x0, c0, x1, c1, B, HW, groups, eps, gamma, beta, scale_shift, workspace, (cudaStream_t)stream);
  if (dtype == ADVS_BF16)
    return gn_stats_impl<__nv_bfloat16>(x0, c0, x1, c1, B, HW, groups, eps, gamma, beta, scale_shift, workspace, (cudaStream_t)stream);
  ADVS_CHECK_ARG(false, "groupnorm_stats: bad dtype");
}

int advs_groupnorm_partial_parts(int B, int HW) {
  if (B <= 0 || HW <= 0) return 0;
  return gn_geom_stats(HW).chunks;
}

int advs_groupnorm_partial(const void* x, int C, int B, int HW, float* part, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && part && C > 0 && C % 8 == 0 && C / 8 <= 1024 && B > 0 && HW > 0, "groupnorm_partial: bad args");
  if (dtype == ADVS_F32) return gn_partial_impl<float>(x, C, B, HW, part, (cudaStream_t)stream);
  if (dtype == ADVS_BF16) return gn_partial_impl<__nv_bfloat16>(x, C, B, HW, part, (cudaStream_t)stream);
  ADVS_CHECK_ARG(false, "groupnorm_partial: bad dtype");
}

int advs_groupnorm_finalize_ex(const float* part0, int c0, int parts0, int gran0, const float* part1, int c1, int parts1,
                               int gran1, int B, int HW, int groups, float eps, const float* gamma, const float* beta,
                               float* scale_shift, void* stream) {
  ADVS_CHECK_ARG(part0 && c0 > 0 && parts0 > 0 && B > 0 && HW > 0 && groups > 0, "groupnorm_finalize: bad args");
  if (!part1) { c1 = 0; parts1 = 0; gran1 = 1; }
  ADVS_CHECK_ARG(c1 == 0 || parts1 > 0, "groupnorm_finalize: second source has no partial rows");
  ADVS_CHECK_ARG((c0 + c1) % groups == 0, "groupnorm_finalize: channels not divisible by groups");
  ADVS_CHECK_ARG(gamma && beta && scale_shift, "groupnorm_finalize: null pointer");
  ADVS_CHECK_ARG((gran0 == 1 || gran0 == 4) && (gran1 == 1 || gran1 == 4), "groupnorm_finalize: gran must be 1 or 4");
  const int cpg = (c0 + c1) / groups;
  ADVS_CHECK_ARG((gran0 == 1 && gran1 == 1) || (cpg % 4 == 0 && c0 % 4 == 0 && c1 % 4 == 0),
                 "groupnorm_finalize: 4-channel partial slots need channels per group and both sources to be multiples of 4");
  return gn_finalize_impl(part0, c0, parts0, gran0, part1, c1, parts1, gran1, B, HW, groups, eps, gamma, beta,
                          scale_shift, (cudaStream_t)stream);
}

int advs_groupnorm_finalize(const float* part0, int c0, int parts0, const float* part1, int c1, int parts1, int B,
                            int HW, int groups, float eps, const float* gamma, const float* beta, float* scale_shift,
                            void* stream) {
  return advs_groupnorm_finalize_ex(part0, c0, parts0, 1, part1, c1, parts1, 1, B, HW, groups, eps, gamma, beta,
                                    scale_shift, stream);
}

int advs_groupnorm_apply(const void* x0, int c0, const void* x1, int c1, int B, int HW, const float* scale_shift,
                         int silu, void* y, int dtype, void* stream) {
  ADVS_CHECK_ARG(x0 && c0 > 0 && B > 0 && HW > 0 && scale_shift && y, "groupnorm_apply: bad args");
  if (!x1) c1 = 0;
  ADVS_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm_apply: channel counts must be multiples of 8");
  ADVS_CHECK_ARG((c0 + c1) / 8 <= 1024, "groupnorm_apply: at most 8192 channels");
  if (dtype == ADVS_F32) return gn_apply_impl<float>(x0, c0, x1, c1, B, HW, scale_shift, silu, y, (cudaStream_t)stream);
  if (dtype == ADVS_BF16) return gn_apply_impl<__nv_bfloat16>(x0, c0, x1, c1, B, HW, scale_shift, silu, y, (cudaStream_t)stream);
  ADVS_CHECK_ARG(false, "groupnorm_apply: bad dtype");
}

int advs_groupnorm_apply_wide(const void* x0, const void* lo0, int c0, const void* x1, const void* lo1, int c1, int B,
                              int HW, const float* scale_shift, int silu, void* y, int y_dtype, void* stream) {
  ADVS_CHECK_ARG(y_dtype == ADVS_BF16 || y_dtype == ADVS_F16, "groupnorm_apply_wide: y_dtype must be ADVS_BF16 or ADVS_F16");
  ADVS_CHECK_ARG(x0 && c0 > 0 && B > 0 && HW > 0 && scale_shift && y, "groupnorm_apply_wide: bad args");
  if (!x1) { c1 = 0; lo1 = nullptr; }
  ADVS_CHECK_ARG(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm_apply_wide: channel counts must be multiples of 8");
  ADVS_CHECK_ARG((c0 + c1) / 8 <= 1024, "groupnorm_apply_wide: at most 8192 channels");
  ADVS_CHECK_ARG(((uintptr_t)lo0 | (uintptr_t)lo1) % 8 == 0, "groupnorm_apply_wide: lo pointers must be 8-byte aligned");
  return gn_apply_wide_impl(x0, lo0, c0, x1, lo1, c1, B, HW, scale_shift, silu, y, y_dtype == ADVS_F16, (cudaStream_t)stream);
}

}  // extern "C"
