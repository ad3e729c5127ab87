// Bandwidth-bound kernels: timestep embedding, small linears, weight packing, stem/head
// convolutions, nearest upsample, DDIM/DDPM updates, per-step row select.
#include "common.cuh"

namespace advs {

// ---------------------------------------------------------------------------------------
// K6: sinusoidal embedding (reference dm1:16-33): [cos | sin], divisor `half`.
__global__ void k_timestep_embedding(const int64_t* __restrict__ t, int nt, int half,
                                     const float* __restrict__ freqs, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * half) return;
  int r = i / half, j = i % half;
  float a = __fmul_rn((float)t[r], freqs[j]);
  out[(size_t)r * 2 * half + j] = cosf(a);
  out[(size_t)r * 2 * half + half + j] = sinf(a);
}

// one warp per output element; rows are tiny (time-embedding MLPs)
__global__ void k_linear_f32(const float* __restrict__ x, const float* __restrict__ w,
                             const float* __restrict__ b, float* __restrict__ y, int rows, int in_f,
                             int out_f, int silu_in, int silu_out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= rows * out_f) return;
  int r = warp / out_f, o = warp % out_f;
  const float* xr = x + (size_t)r * in_f;
  const float* wr = w + (size_t)o * in_f;
  float acc = 0.f;
  for (int i = lane; i < in_f; i += 32) {
    float xv = xr[i];
    if (silu_in) xv = silu_acc(xv);
    acc = fmaf(xv, wr[i], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc + (b ? b[o] : 0.f);
    if (silu_out) v = silu_acc(v);
    y[(size_t)r * out_f + o] = v;
  }
}

// OIHW fp32 -> [O][kh*kw][I] T
template <typename T>
__global__ void k_pack_conv_weight(const float* __restrict__ w, T* __restrict__ dst, int O, int I, int taps) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)O * I * taps;
  if (i >= n) return;
  int c = (int)(i % I);
  int tap = (int)((i / I) % taps);
  int o = (int)(i / ((size_t)I * taps));
  dst[i] = from_f<T>(w[((size_t)o * I + c) * taps + tap]);
}

// OIHW 3x3 fp32 -> [phase][O][4][I] T (nearest-2x upsample folded into the kernel, see header)
// Stem (Cin = 3) on the tensor cores: the 3x3 neighbourhood of the fp32 NCHW input becomes one 64-entry bf16 row
// per pixel, so the stem is a 1x1 convolution with 64 input "channels" for the implicit-GEMM kernel.
// entry = part * 9*Cin + tap * Cin + ci; part 0 = bf16(x), part 1 = bf16(x - bf16(x)) when it fits (the fp32
// sampler state then reaches the bf16 MMA to ~2^-17 instead of 2^-9), the rest is zero.  Zero padding of the
// convolution = zeros here.
// One thread per pixel gathers its 9*Cin inputs (lanes = consecutive pixels of a row: coalesced), the block's 256
// rows of 128 bytes go through shared memory (144-byte pitch: conflict-free 16-byte writes) and leave as contiguous
// 16-byte pieces, so both sides touch whole cache lines.
// CIN > 0: compile-time channel count; CIN = 0: run-time Cin (<= 7).
constexpr int kStemPixPerBlock = 256;
constexpr int kStemPitch = 144;
// F16: the rows are fp16 pairs instead of bf16 pairs (the sampler state is bounded: |x_t| stays within a few units),
// to go with fp16 stem weights -- tcgen05 wants both operands of an MMA in the same 16-bit format.
template <bool F16>
__device__ __forceinline__ uint32_t stem_pack2(float a, float b) {
  if constexpr (F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    return pack_bf16x2(a, b);
  }
}
template <bool F16>
__device__ __forceinline__ float stem_round(float v) {
  if constexpr (F16) return __half2float(__float2half_rn(v));
  else return __bfloat162float(__float2bfloat16_rn(v));
}
template <int CIN, bool F16 = false>
__global__ void __launch_bounds__(kStemPixPerBlock)
k_stem_im2col(const float* __restrict__ x, __nv_bfloat16* __restrict__ col, int B, int H, int W, int cin_rt, int parts) {
  __shared__ __align__(16) uint8_t rows[kStemPixPerBlock * kStemPitch];
  pdl_sync();
  const int Cin = CIN > 0 ? CIN : cin_rt;
  const size_t npix = (size_t)B * H * W;
  const size_t pix0 = (size_t)blockIdx.x * kStemPixPerBlock;
  const size_t pix = pix0 + threadIdx.x;
  __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(rows + threadIdx.x * kStemPitch);
  if (pix < npix) {
    const int w0 = (int)(pix % W), h0 = (int)((pix / W) % H), b = (int)(pix / ((size_t)W * H));
    const int k9 = 9 * Cin;
    if constexpr (CIN > 0) {
      // all indices are compile-time: the row is built in registers and written with eight 16-byte stores
      float v[9 * CIN];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h0 + tap / 3 - 1, ww = w0 + tap % 3 - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
          v[tap * CIN + ci] = in ? __ldg(x + (((size_t)b * CIN + ci) * H + hh) * W + ww) : 0.f;
      }
      float e[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) e[i] = 0.f;
#pragma unroll
      for (int i = 0; i < 9 * CIN; ++i) {
        e[i] = v[i];
        if (18 * CIN <= 64) e[9 * CIN + i] = v[i] - stem_round<F16>(v[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        reinterpret_cast<uint4*>(row)[i] = make_uint4(stem_pack2<F16>(e[8 * i], e[8 * i + 1]), stem_pack2<F16>(e[8 * i + 2], e[8 * i + 3]),
                                                      stem_pack2<F16>(e[8 * i + 4], e[8 * i + 5]), stem_pack2<F16>(e[8 * i + 6], e[8 * i + 7]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(row)[i] = make_uint4(0u, 0u, 0u, 0u);
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h0 + tap / 3 - 1, ww = w0 + tap % 3 - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
        for (int ci = 0; ci < Cin; ++ci) {
          const float val = in ? __ldg(x + (((size_t)b * Cin + ci) * H + hh) * W + ww) : 0.f;
          const float hif = stem_round<F16>(val);
          uint16_t* r16 = reinterpret_cast<uint16_t*>(row);
          r16[tap * Cin + ci] = (uint16_t)(stem_pack2<F16>(hif, 0.f) & 0xFFFFu);
          if (parts > 1) r16[k9 + tap * Cin + ci] = (uint16_t)(stem_pack2<F16>(val - hif, 0.f) & 0xFFFFu);
        }
      }
    }
  }
  __syncthreads();
  // 256 rows x 8 pieces of 16 bytes, contiguous in global memory
  uint4* dst = reinterpret_cast<uint4*>(col + pix0 * 64);
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int g = it * kStemPixPerBlock + threadIdx.x;
    const int r = g >> 3, c = g & 7;
    if (pix0 + r < npix) dst[g] = *reinterpret_cast<const uint4*>(rows + r * kStemPitch + c * 16);
  }
}

// matching weight rows [O][64]: both parts carry bf16(w[o][ci][tap])
template <typename T>
__global__ void k_pack_stem_weight(const float* __restrict__ w, T* __restrict__ dst, int O, int I, int parts) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)O * 64) return;
  const int o = (int)(idx >> 6), slot = (int)(idx & 63);
  const int k9 = 9 * I;
  const int part = slot / k9, r = slot - part * k9;
  const int tap = r / I, ci = r - tap * I;
  dst[idx] = from_f<T>(part < parts ? w[((size_t)o * I + ci) * 9 + tap] : 0.f);
}

template <typename T>
__global__ void k_pack_upconv_weight(const float* __restrict__ w, T* __restrict__ dst, int O, int I) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)4 * O * 4 * I;
  if (i >= n) return;
  int c = (int)(i % I);
  int tap = (int)((i / I) % 4);
  int o = (int)((i / ((size_t)4 * I)) % O);
  int phase = (int)(i / ((size_t)4 * I * O));
  const int a = phase >> 1, b = phase & 1, p = tap >> 1, q = tap & 1;
  // rows of the 3x3 kernel that land on low-res row offset p for output-row parity a
  const int r0 = (a == 0) ? (p == 0 ? 0 : 1) : (p == 0 ? 0 : 2);
  const int r1 = (a == 0) ? (p == 0 ? 0 : 2) : (p == 0 ? 1 : 2);
  const int c0 = (b == 0) ? (q == 0 ? 0 : 1) : (q == 0 ? 0 : 2);
  const int c1 = (b == 0) ? (q == 0 ? 0 : 2) : (q == 0 ? 1 : 2);
  const float* wp = w + ((size_t)o * I + c) * 9;
  float acc = 0.f;
  for (int dy = r0; dy <= r1; ++dy)
    for (int dx = c0; dx <= c1; ++dx) acc += wp[dy * 3 + dx];
  dst[i] = from_f<T>(acc);
}

// ---------------------------------------------------------------------------------------
// stem: x NCHW fp32 (Cin small) -> y NHWC T.  One thread = one pixel x CPT output channels:
// the 9*Cin inputs sit in registers (coalesced NCHW loads: consecutive lanes = consecutive pixels),
// weights are read from smem as warp-wide broadcasts ([k][cout] layout, float4 per load).
template <typename T, int CPT, int KMAX>
__global__ void __launch_bounds__(128) k_conv3x3_stem(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, T* __restrict__ y, int B, int H,
                                                      int W, int Cin, int Cout) {
  extern __shared__ float sw[];  // [9*Cin][Cout]
  const int K = 9 * Cin;
  for (int j = threadIdx.x; j < Cout * K; j += blockDim.x) {  // conflict-free smem stores
    int k = j / Cout, o = j - k * Cout;
    sw[j] = w[o * K + k];
  }
  __syncthreads();
  const size_t total = (size_t)B * H * W;
  const int cbase = blockIdx.y * CPT;
  for (int it = 0; it < 8; ++it) {
  size_t p = ((size_t)blockIdx.x * 8 + it) * blockDim.x + threadIdx.x;
  if (p >= total) return;
  int wo = (int)(p % W);
  int ho = (int)((p / W) % H);
  int b = (int)(p / ((size_t)W * H));
  float xin[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) xin[k] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    int hi = ho + tap / 3 - 1, wi = wo + tap % 3 - 1;
    bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
#pragma unroll
    for (int c = 0; c < KMAX / 9; ++c)
      if (c < Cin && ok) xin[tap * (KMAX / 9) + c] = x[(((size_t)b * Cin + c) * H + hi) * W + wi];
  }
  float acc[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) acc[i] = bias ? bias[cbase + i] : 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
    for (int c = 0; c < KMAX / 9; ++c) {
      if (c < Cin) {
        const float xv = xin[tap * (KMAX / 9) + c];
        const float4* wr = reinterpret_cast<const float4*>(sw + (tap * Cin + c) * Cout + cbase);
#pragma unroll
        for (int i = 0; i < CPT / 4; ++i) {
          float4 w4 = wr[i];
          acc[4 * i] = fmaf(xv, w4.x, acc[4 * i]);
          acc[4 * i + 1] = fmaf(xv, w4.y, acc[4 * i + 1]);
          acc[4 * i + 2] = fmaf(xv, w4.z, acc[4 * i + 2]);
          acc[4 * i + 3] = fmaf(xv, w4.w, acc[4 * i + 3]);
        }
      }
    }
  }
  T* yp = y + p * Cout + cbase;
#pragma unroll
  for (int i = 0; i < CPT / 8; ++i) {
    Vec8<T> v;
    v.from_float(acc + 8 * i);
    v.store(yp + 8 * i);
  }
  }
}

// head: x NHWC T -> y NCHW fp32 (Cout small, <= 4). one warp per pixel, lanes split channels.
template <typename T, int COUT_MAX>
__global__ void k_conv3x3_head(const T* __restrict__ x, const float* __restrict__ w,
                               const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W,
                               int Cin, int Cout) {
  extern __shared__ float sw[];  // [Cout][9][Cin]
  for (int i = threadIdx.x; i < Cout * 9 * Cin; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  int lane = threadIdx.x & 31;
  size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  size_t total = (size_t)B * H * W;
  if (warp >= total) return;
  int wo = (int)(warp % W);
  int ho = (int)((warp / W) % H);
  int b = (int)(warp / ((size_t)W * H));
  float acc[COUT_MAX];
#pragma unroll
  for (int o = 0; o < COUT_MAX; ++o) acc[o] = 0.f;
  for (int tap = 0; tap < 9; ++tap) {
    int hi = ho + tap / 3 - 1, wi = wo + tap % 3 - 1;
    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
    const T* xp = x + (((size_t)b * H + hi) * W + wi) * Cin;
    for (int c = lane * 4; c < Cin; c += 128) {
      float xv[4];
      if constexpr (sizeof(T) == 4) {
        float4 t = *reinterpret_cast<const float4*>(xp + c);
        xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      } else {
        uint2 t = *reinterpret_cast<const uint2*>(xp + c);
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
        float2 d = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
        xv[0] = a.x; xv[1] = a.y; xv[2] = d.x; xv[3] = d.y;
      }
#pragma unroll
      for (int o = 0; o < COUT_MAX; ++o) {
        if (o < Cout) {
          const float* wp = sw + ((size_t)o * 9 + tap) * Cin + c;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[o] = fmaf(xv[i], wp[i], acc[o]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COUT_MAX; ++o) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    for (int o = 0; o < Cout; ++o)
      y[(((size_t)b * Cout + o) * H + ho) * W + wo] = acc[o] + (bias ? bias[o] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void k_upsample2x(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  // y [B,2H,2W,C]; one thread per 8 channels of an output pixel
  const int cv = C / 8;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * 4 * H * W * cv;
  if (idx >= total) return;
  int j = (int)(idx % cv);
  size_t p = idx / cv;
  int wo = (int)(p % (2 * W));
  int ho = (int)((p / (2 * W)) % (2 * H));
  int b = (int)(p / ((size_t)4 * W * H));
  Vec8<T> v;
  v.load(x + ((((size_t)b * H + (ho >> 1)) * W + (wo >> 1)) * C + j * 8));
  v.store(y + p * C + j * 8);
}

// ---------------------------------------------------------------------------------------
// K8. DDIM update in the reference's exact fp32 operation order (dm1:457-472).
// x and out may alias (the samplers update the state in place): no __restrict__ on either.
__global__ void k_ddim_step(const float* x, const float* __restrict__ eps,
                            const float* __restrict__ noise, float* out, size_t n,
                            const float* __restrict__ coef, const int32_t* __restrict__ step_dev, int clip) {
  pdl_sync();
  const float* c = coef + 8 * (size_t)(*step_dev);
  const float s1 = c[0], sa = c[1], sp = c[2], cdir = c[3], sigma = c[4];
  size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float xv[4], ev[4], zv[4] = {0.f, 0.f, 0.f, 0.f}, ov[4];
    if (i + 4 <= n) {
      float4 t = *reinterpret_cast<const float4*>(x + i);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      t = *reinterpret_cast<const float4*>(eps + i);
      ev[0] = t.x; ev[1] = t.y; ev[2] = t.z; ev[3] = t.w;
      if (noise) {
        t = *reinterpret_cast<const float4*>(noise + i);
        zv[0] = t.x; zv[1] = t.y; zv[2] = t.z; zv[3] = t.w;
      }
    } else {
      for (int k = 0; k < 4; ++k) {
        bool ok = i + k < n;
        xv[k] = ok ? x[i + k] : 0.f;
        ev[k] = ok ? eps[i + k] : 0.f;
        zv[k] = (ok && noise) ? noise[i + k] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0 = __fdiv_rn(__fsub_rn(xv[k], __fmul_rn(s1, ev[k])), sa);
      if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      float v = __fadd_rn(__fmul_rn(sp, x0), __fmul_rn(cdir, ev[k]));
      ov[k] = __fadd_rn(v, __fmul_rn(sigma, zv[k]));
    }
    if (i + 4 <= n) {
      *reinterpret_cast<float4*>(out + i) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    } else {
      for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = ov[k];
    }
  }
}

// DDPM ancestral step (dm1:356-395): x0 = c0*x - c1*eps; clamp; mean = c2*x0 + c3*x; + c4*z
__global__ void k_ddpm_step(const float* x, const float* __restrict__ eps,
                            const float* __restrict__ noise, float* out, size_t n,
                            const float* __restrict__ coef, const int32_t* __restrict__ step_dev, int clip) {
  pdl_sync();
  const float* c = coef + 8 * (size_t)(*step_dev);
  const float c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float xv = x[i], ev = eps[i], zv = noise ? noise[i] : 0.f;
    float x0 = __fsub_rn(__fmul_rn(c0, xv), __fmul_rn(c1, ev));
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    float mean = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, xv));
    out[i] = __fadd_rn(mean, __fmul_rn(c4, zv));
  }
}

__global__ void k_advance_step(int32_t* step_dev, int advance) {
  pdl_sync();
  *step_dev += advance;
}

__global__ void k_select_row(const float* __restrict__ table, int row_floats,
                             const int32_t* __restrict__ step_dev, float* __restrict__ dst, int reps) {
  pdl_sync();
  const float* src = table + (size_t)(*step_dev) * row_floats;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < row_floats; i += gridDim.x * blockDim.x) {
    float v = src[i];
    for (int r = 0; r < reps; ++r) dst[(size_t)r * row_floats + i] = v;
  }
}

}  // namespace advs

using namespace advs;

extern "C" {

int advs_timestep_embedding(const int64_t* t, int nt, const float* freqs, int half, float* out, void* stream) {
  ADVS_CHECK_ARG(t && out && freqs && nt > 0 && half >= 1, "timestep_embedding: bad args");
  int n = nt * half;
  k_timestep_embedding<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, nt, half, freqs, out);
  ADVS_CHECK_LAUNCH("timestep_embedding");
  return ADVS_OK;
}

int advs_linear_f32(const float* x, const float* w, const float* b, float* y, int rows, int in_f, int out_f,
                    int silu_in, int silu_out, void* stream) {
  ADVS_CHECK_ARG(x && w && y && rows > 0 && in_f > 0 && out_f > 0, "linear_f32: bad args");
  size_t warps = (size_t)rows * out_f;
  size_t blocks = (warps * 32 + 255) / 256;
  k_linear_f32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, w, b, y, rows, in_f, out_f, silu_in, silu_out);
  ADVS_CHECK_LAUNCH("linear_f32");
  return ADVS_OK;
}

int advs_pack_conv_weight(const float* w, void* dst, int O, int I, int kh, int kw, int dtype, void* stream) {
  ADVS_CHECK_ARG(w && dst && O > 0 && I > 0 && kh > 0 && kw > 0, "pack_conv_weight: bad args");
  size_t n = (size_t)O * I * kh * kw;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == ADVS_F32)
    k_pack_conv_weight<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (float*)dst, O, I, kh * kw);
  else if (dtype == ADVS_BF16)
    k_pack_conv_weight<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)dst, O, I, kh * kw);
  else if (dtype == ADVS_F16)
    k_pack_conv_weight<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)dst, O, I, kh * kw);
  else
    ADVS_CHECK_ARG(false, "pack_conv_weight: bad dtype");
  ADVS_CHECK_LAUNCH("pack_conv_weight");
  return ADVS_OK;
}

int advs_pack_upconv_weight(const float* w, void* dst, int O, int I, int dtype, void* stream) {
  ADVS_CHECK_ARG(w && dst && O > 0 && I > 0, "pack_upconv_weight: bad args");
  size_t n = (size_t)16 * O * I;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == ADVS_F32)
    k_pack_upconv_weight<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (float*)dst, O, I);
  else if (dtype == ADVS_BF16)
    k_pack_upconv_weight<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)dst, O, I);
  else if (dtype == ADVS_F16)
    k_pack_upconv_weight<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)dst, O, I);
  else
    ADVS_CHECK_ARG(false, "pack_upconv_weight: bad dtype");
  ADVS_CHECK_LAUNCH("pack_upconv_weight");
  return ADVS_OK;
}

int advs_stem_im2col_ex(const float* x, void* col, int B, int H, int W, int Cin, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && col && B > 0 && H > 0 && W > 0 && Cin > 0 && 9 * Cin <= 64, "stem_im2col: bad args (needs 9*Cin <= 64)");
  ADVS_CHECK_ARG(dtype == ADVS_BF16 || dtype == ADVS_F16, "stem_im2col: dtype must be ADVS_BF16 or ADVS_F16");
  const size_t npix = (size_t)B * H * W;
  const unsigned blocks = (unsigned)((npix + kStemPixPerBlock - 1) / kStemPixPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* c = (__nv_bfloat16*)col;
  const int parts = 18 * Cin <= 64 ? 2 : 1;
  if (Cin == 3) {
    if (dtype == ADVS_F16) launch_pdl(k_stem_im2col<3, true>, dim3(blocks), dim3(kStemPixPerBlock), 0, st, x, c, B, H, W, Cin, 2);
    else launch_pdl(k_stem_im2col<3, false>, dim3(blocks), dim3(kStemPixPerBlock), 0, st, x, c, B, H, W, Cin, 2);
  } else {
    if (dtype == ADVS_F16) launch_pdl(k_stem_im2col<0, true>, dim3(blocks), dim3(kStemPixPerBlock), 0, st, x, c, B, H, W, Cin, parts);
    else launch_pdl(k_stem_im2col<0, false>, dim3(blocks), dim3(kStemPixPerBlock), 0, st, x, c, B, H, W, Cin, parts);
  }
  ADVS_CHECK_LAUNCH("stem_im2col");
  return ADVS_OK;
}

int advs_stem_im2col(const float* x, void* col, int B, int H, int W, int Cin, void* stream) {
  return advs_stem_im2col_ex(x, col, B, H, W, Cin, ADVS_BF16, stream);
}

int advs_pack_stem_weight_ex(const float* w, void* dst, int O, int I, int dtype, void* stream) {
  ADVS_CHECK_ARG(w && dst && O > 0 && I > 0 && 9 * I <= 64, "pack_stem_weight: bad args (needs 9*I <= 64)");
  ADVS_CHECK_ARG(dtype == ADVS_BF16 || dtype == ADVS_F16, "pack_stem_weight: dtype must be ADVS_BF16 or ADVS_F16");
  const size_t n = (size_t)O * 64;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const int parts = 18 * I <= 64 ? 2 : 1;
  if (dtype == ADVS_F16)
    k_pack_stem_weight<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__half*)dst, O, I, parts);
  else
    k_pack_stem_weight<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)dst, O, I, parts);
  ADVS_CHECK_LAUNCH("pack_stem_weight");
  return ADVS_OK;
}

int advs_pack_stem_weight(const float* w, void* dst, int O, int I, void* stream) {
  return advs_pack_stem_weight_ex(w, dst, O, I, ADVS_BF16, stream);
}

int advs_conv3x3_stem(const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int Cin,
                      int Cout, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && w && y && B > 0 && H > 0 && W > 0, "conv3x3_stem: bad args");
  ADVS_CHECK_ARG(Cout % 32 == 0 && Cin >= 1 && Cin <= 4, "conv3x3_stem: needs Cout%%32==0 and Cin<=4");
  size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
  ADVS_CHECK_ARG(smem <= 48 * 1024, "conv3x3_stem: weights exceed 48 KB of shared memory");
  size_t total = (size_t)B * H * W;
  dim3 grid((unsigned)((total + 1023) / 1024), Cout / 32);
  if (dtype == ADVS_F32)
    k_conv3x3_stem<float, 32, 36><<<grid, 128, smem, (cudaStream_t)stream>>>(x, w, bias, (float*)y, B, H, W, Cin, Cout);
  else
    k_conv3x3_stem<__nv_bfloat16, 32, 36><<<grid, 128, smem, (cudaStream_t)stream>>>(x, w, bias, (__nv_bfloat16*)y, B, H, W, Cin, Cout);
  ADVS_CHECK_LAUNCH("conv3x3_stem");
  return ADVS_OK;
}

int advs_conv3x3_head(const void* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                      int Cout, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && w && y && B > 0 && H > 0 && W > 0, "conv3x3_head: bad args");
  ADVS_CHECK_ARG(Cout >= 1 && Cout <= 4 && Cin % 4 == 0, "conv3x3_head: needs Cout<=4 and Cin%%4==0");
  size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
  ADVS_CHECK_ARG(smem <= 48 * 1024, "conv3x3_head: weights exceed 48 KB of shared memory");
  size_t total_threads = (size_t)B * H * W * 32;
  unsigned blocks = (unsigned)((total_threads + 255) / 256);
  if (dtype == ADVS_F32)
    k_conv3x3_head<float, 4><<<blocks, 256, smem, (cudaStream_t)stream>>>((const float*)x, w, bias, y, B, H, W, Cin, Cout);
  else
    k_conv3x3_head<__nv_bfloat16, 4><<<blocks, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, w, bias, y, B, H, W, Cin, Cout);
  ADVS_CHECK_LAUNCH("conv3x3_head");
  return ADVS_OK;
}

int advs_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample_nearest2x: bad args (C%%8)");
  size_t total = (size_t)B * 4 * H * W * (C / 8);
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == ADVS_F32)
    k_upsample2x<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, B, H, W, C);
  else
    k_upsample2x<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C);
  ADVS_CHECK_LAUNCH("upsample_nearest2x");
  return ADVS_OK;
}

static unsigned ew_blocks(size_t n, int per_thread) {
  size_t b = (n / per_thread + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (unsigned)b;
}

int advs_ddim_step(const float* x, const float* eps, const float* noise, float* out, size_t n, const float* coef,
                   int32_t* step_dev, int advance, int clip, void* stream) {
  ADVS_CHECK_ARG(x && eps && out && coef && step_dev && n > 0, "ddim_step: bad args");
  ADVS_CHECK_ARG(((uintptr_t)x | (uintptr_t)eps | (uintptr_t)out | (uintptr_t)noise) % 16 == 0,
                 "ddim_step: pointers must be 16-byte aligned");
  launch_pdl(k_ddim_step, dim3(ew_blocks(n, 4)), dim3(256), 0, (cudaStream_t)stream, x, eps, noise, out, n, coef, (const int32_t*)step_dev, clip);
  ADVS_CHECK_LAUNCH("ddim_step");
  if (advance) {
    launch_pdl(k_advance_step, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev, advance);
    ADVS_CHECK_LAUNCH("ddim_step/advance");
  }
  return ADVS_OK;
}

int advs_ddpm_step(const float* x, const float* eps, const float* noise, float* out, size_t n, const float* coef,
                   int32_t* step_dev, int advance, int clip, void* stream) {
  ADVS_CHECK_ARG(x && eps && out && coef && step_dev && n > 0, "ddpm_step: bad args");
  launch_pdl(k_ddpm_step, dim3(ew_blocks(n, 1)), dim3(256), 0, (cudaStream_t)stream, x, eps, noise, out, n, coef, (const int32_t*)step_dev, clip);
  ADVS_CHECK_LAUNCH("ddpm_step");
  if (advance) {
    launch_pdl(k_advance_step, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev, advance);
    ADVS_CHECK_LAUNCH("ddpm_step/advance");
  }
  return ADVS_OK;
}

int advs_select_row(const float* table, int row_floats, const int32_t* step_dev, float* dst, int reps, void* stream) {
  ADVS_CHECK_ARG(table && step_dev && dst && row_floats > 0 && reps > 0, "select_row: bad args");
  int blocks = (row_floats + 255) / 256;
  if (blocks > 148) blocks = 148;
  launch_pdl(k_select_row, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, table, row_floats, step_dev, dst, reps);
  ADVS_CHECK_LAUNCH("select_row");
  return ADVS_OK;
}

}  // extern "C"
