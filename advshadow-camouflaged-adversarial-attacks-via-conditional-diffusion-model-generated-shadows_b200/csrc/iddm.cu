// Bandwidth kernels of the IDDM class-conditional UNet + CFG DDIM sampler:
//   reference: model/networks/unet.py:17-128, model/modules/{conv,block,attention}.py, model/networks/base.py:56-68,
//   model/samples/ddim.py:48-100.
// Convolutions / token linears / attention reuse the implicit-GEMM and attention kernels; this file adds MaxPool2d(2),
// bilinear x2 (align_corners=True) written into a concat slice, LayerNorm, a GroupNorm-apply with residual / embedding /
// activation options, activations, the sin|cos position encoding + label embedding, CFG lerp and the uint8 cast.
#include "common.cuh"

namespace advs {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <typename T>
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return sizeof(T) == 4 ? silu_acc(v) : silu_f(v);
  if (act == 2) return gelu_erf(v);
  return v;
}

// ---- MaxPool2d(2): y[b,h,w,:] = max over the 2x2 window (block.py:25) ----
template <typename T>
__global__ void k_maxpool2x2(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  const int cv = C / 8;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * (H / 2) * (W / 2) * cv;
  if (idx >= total) return;
  int j = (int)(idx % cv);
  size_t p = idx / cv;
  int wo = (int)(p % (W / 2)), ho = (int)((p / (W / 2)) % (H / 2)), b = (int)(p / ((size_t)(W / 2) * (H / 2)));
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      Vec8<T> v;
      v.load(x + ((((size_t)b * H + 2 * ho + dy) * W + 2 * wo + dx) * C + j * 8));
      float f[8];
      v.to_float(f);
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
    }
  Vec8<T> o;
  o.from_float(m);
  o.store(y + p * C + j * 8);
}

// ---- nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) (block.py:56), written into channels
//      [coff, coff+C) of a [B,2H,2W,cstride] tensor (the torch.cat([skip_x, x]) of block.py:73 never copies twice) ----
template <typename T>
__global__ void k_upsample_bilinear2x(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int cstride,
                                      int coff) {
  const int cv = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)B * Ho * Wo * cv;
  if (idx >= total) return;
  int j = (int)(idx % cv);
  size_t p = idx / cv;
  int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho), b = (int)(p / ((size_t)Wo * Ho));
  // torch: src = dst * (in - 1) / (out - 1), computed in fp32
  const float sh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sh * ho, fx = sw * wo;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
  const T* base = x + (size_t)b * H * W * C + j * 8;
  Vec8<T> v00, v01, v10, v11;
  v00.load(base + ((size_t)y0 * W + x0) * C);
  v01.load(base + ((size_t)y0 * W + x1) * C);
  v10.load(base + ((size_t)y1 * W + x0) * C);
  v11.load(base + ((size_t)y1 * W + x1) * C);
  float a[8], bb[8], c[8], d[8], o[8];
  v00.to_float(a); v01.to_float(bb); v10.to_float(c); v11.to_float(d);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = hy * (hx * a[i] + lx * bb[i]) + ly * (hx * c[i] + lx * d[i]);
  Vec8<T> ov;
  ov.from_float(o);
  ov.store(y + p * cstride + coff + j * 8);
}

template <typename T>
__global__ void k_copy_channels(const T* __restrict__ x, T* __restrict__ y, size_t npix, int C, int cstride, int coff) {
  const int cv = C / 8;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * cv) return;
  int j = (int)(idx % cv);
  size_t p = idx / cv;
  Vec8<T> v;
  v.load(x + p * C + j * 8);
  v.store(y + p * cstride + coff + j * 8);
}

// ---- nn.LayerNorm([C]) over the channel axis of each token (attention.py:29,31); one warp per token ----
template <typename T, bool OUT_F16 = false>
__global__ void k_layernorm(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                            T* __restrict__ y, size_t rows, int C, float eps) {
  size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f(xr[c]);
  s = warp_sum(s);
  const float mean = s / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    float d = to_f(xr[c]) - mean;
    q = fmaf(d, d, q);
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / C + eps);
  T* yr = y + row * C;
  for (int c = lane; c < C; c += 32) {
    const float v = (to_f(xr[c]) - mean) * rstd * g[c] + bta[c];
    if constexpr (OUT_F16 && sizeof(T) == 2) reinterpret_cast<__half*>(yr)[c] = __float2half_rn(v);   // bounded: fp16 GEMM operand
    else yr[c] = from_f<T>(v);
  }
}

// ---- GroupNorm apply with the DoubleConv / Down/UpBlock epilogues folded in:
//      y = act( x*scale + shift (+ residual) ) (+ emb[b, c])     (conv.py:40-66, block.py:44-46) ----
template <typename T, bool OUT_F16 = false>
__global__ void k_gn_apply_ex(const T* __restrict__ x, int HW, int C, size_t npix, const float* __restrict__ scale_shift,
                              const T* __restrict__ residual, const float* __restrict__ emb, int emb_stride, int act,
                              T* __restrict__ y, const uint8_t* __restrict__ x_lo = nullptr) {
  const int cv = C / 8;
  size_t total = npix * cv;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int j = (int)(idx % cv);
    size_t p = idx / cv;
    int b = (int)(p / HW);
    int c = j * 8;
    Vec8<T> v;
    v.load(x + p * C + c);
    float f[8], r[8];
    if constexpr (sizeof(T) == 2) {
      if (x_lo) wide_decode8(v.v, *reinterpret_cast<const uint2*>(x_lo + p * C + c), f);   // bf16 + int8 extension
      else v.to_float(f);
    } else {
      v.to_float(f);
    }
    if (residual) {
      Vec8<T> rv;
      rv.load(residual + p * C + c);
      rv.to_float(r);
    }
    const float* ab = scale_shift + ((size_t)b * C + c) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = fmaf(f[i], ab[2 * i], ab[2 * i + 1]);
      if (residual) a += r[i];
      a = act_apply<T>(a, act);
      if (emb) a += emb[(size_t)b * emb_stride + c + i];
      f[i] = a;
    }
    if constexpr (OUT_F16 && sizeof(T) == 2) {
      __half2* h = reinterpret_cast<__half2*>(&v.v);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    } else {
      v.from_float(f);
    }
    v.store(y + p * C + c);
  }
}

template <typename T>
__global__ void k_act(const T* __restrict__ x, T* __restrict__ y, size_t n, int act) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  Vec8<T> v;
  v.load(x + i);
  float f[8];
  v.to_float(f);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = act_apply<T>(f[k], act);
  v.from_float(f);
  v.store(y + i);
}

// ---- BaseNet.pos_encoding (base.py:56-68): [sin(t f_j) | cos(t f_j)], then time += label_emb(y) (unet.py:106-107) ----
__global__ void k_pos_encoding(const int64_t* __restrict__ t, const float* __restrict__ inv_freq, int half,
                               const int64_t* __restrict__ labels, const float* __restrict__ label_emb, int nt,
                               float* __restrict__ out, int n_labeled) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * half) return;
  int r = i / half, j = i % half;
  float a = __fmul_rn((float)t[r], inv_freq[j]);
  float s = sinf(a), c = cosf(a);
  if (labels && r < n_labeled) {
    const float* e = label_emb + (size_t)labels[r] * 2 * half;
    s = __fadd_rn(s, e[j]);
    c = __fadd_rn(c, e[half + j]);
  }
  out[(size_t)r * 2 * half + j] = s;
  out[(size_t)r * 2 * half + half + j] = c;
}

// ---- torch.lerp(uncond, cond, w) (ddim.py:89), ATen's two-branch formula ----
__global__ void k_cfg_lerp(const float* __restrict__ uncond, const float* __restrict__ cond, float w,
                           float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = uncond[i], e = cond[i];
  const float d = __fsub_rn(e, s);
  // ATen CPU kernel: |w| < 0.5 -> fma(w, end - start, start), else end - (end - start) * (1 - w)
  out[i] = fabsf(w) < 0.5f ? __fmaf_rn(w, d, s) : __fsub_rn(e, __fmul_rn(d, __fsub_rn(1.f, w)));
}

// ---- x = (x + 1) * 0.5; (x * 255).type(uint8) WITHOUT clamp (ddim.py:97-99): C-style truncation then wrap ----
__global__ void k_to_uint8(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __fmul_rn(__fmul_rn(__fadd_rn(x[i], 1.f), 0.5f), 255.f);
  out[i] = (uint8_t)(long long)v;
}

static unsigned blocks_for(size_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

}  // namespace advs

using namespace advs;

#define DISPATCH_T(dtype, CALL_F32, CALL_BF16, who)        \
  if ((dtype) == ADVS_F32) { CALL_F32; }                   \
  else if ((dtype) == ADVS_BF16) { CALL_BF16; }            \
  else ADVS_CHECK_ARG(false, who ": bad dtype");

extern "C" {

int advs_maxpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2x2: bad args");
  size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (k_maxpool2x2<float><<<blocks_for(total), 256, 0, st>>>((const float*)x, (float*)y, B, H, W, C)),
             (k_maxpool2x2<__nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C)),
             "maxpool2x2");
  ADVS_CHECK_LAUNCH("maxpool2x2");
  return ADVS_OK;
}

int advs_upsample_bilinear2x(const void* x, void* y, int B, int H, int W, int C, int y_cstride, int y_coff, int dtype,
                             void* stream) {
  ADVS_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C % 8 == 0 && y_cstride % 8 == 0 && y_coff % 8 == 0 &&
                     y_coff + C <= y_cstride, "upsample_bilinear2x: bad args");
  size_t total = (size_t)B * 4 * H * W * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype,
             (k_upsample_bilinear2x<float><<<blocks_for(total), 256, 0, st>>>((const float*)x, (float*)y, B, H, W, C, y_cstride, y_coff)),
             (k_upsample_bilinear2x<__nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C, y_cstride, y_coff)),
             "upsample_bilinear2x");
  ADVS_CHECK_LAUNCH("upsample_bilinear2x");
  return ADVS_OK;
}

int advs_copy_channels(const void* x, void* y, size_t npix, int C, int y_cstride, int y_coff, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && y && npix > 0 && C % 8 == 0 && y_cstride % 8 == 0 && y_coff % 8 == 0 && y_coff + C <= y_cstride,
                 "copy_channels: bad args");
  size_t total = npix * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (k_copy_channels<float><<<blocks_for(total), 256, 0, st>>>((const float*)x, (float*)y, npix, C, y_cstride, y_coff)),
             (k_copy_channels<__nv_bfloat16><<<blocks_for(total), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, npix, C, y_cstride, y_coff)),
             "copy_channels");
  ADVS_CHECK_LAUNCH("copy_channels");
  return ADVS_OK;
}

int advs_layernorm_f16out(const void* x, const float* gamma, const float* beta, void* y, size_t rows, int C, float eps,
                          void* stream) {
  ADVS_CHECK_ARG(x && y && gamma && beta && rows > 0 && C > 0, "layernorm: bad args");
  k_layernorm<__nv_bfloat16, true><<<blocks_for(rows * 32), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, rows, C, eps);
  ADVS_CHECK_LAUNCH("layernorm_f16out");
  return ADVS_OK;
}

int advs_layernorm(const void* x, const float* gamma, const float* beta, void* y, size_t rows, int C, float eps, int dtype,
                   void* stream) {
  ADVS_CHECK_ARG(x && y && gamma && beta && rows > 0 && C > 0, "layernorm: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = blocks_for(rows * 32);
  DISPATCH_T(dtype, (k_layernorm<float><<<blocks, 256, 0, st>>>((const float*)x, gamma, beta, (float*)y, rows, C, eps)),
             (k_layernorm<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, rows, C, eps)),
             "layernorm");
  ADVS_CHECK_LAUNCH("layernorm");
  return ADVS_OK;
}

int advs_groupnorm_apply_ex(const void* x, int B, int HW, int C, const float* scale_shift, const void* residual,
                            const float* emb, int emb_stride, int act, void* y, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && y && scale_shift && B > 0 && HW > 0 && C % 8 == 0 && act >= 0 && act <= 2, "groupnorm_apply_ex: bad args");
  size_t npix = (size_t)B * HW;
  size_t total = npix * (C / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype,
             (k_gn_apply_ex<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, HW, C, npix, scale_shift, (const float*)residual, emb, emb_stride, act, (float*)y)),
             (k_gn_apply_ex<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, HW, C, npix, scale_shift, (const __nv_bfloat16*)residual, emb, emb_stride, act, (__nv_bfloat16*)y)),
             "groupnorm_apply_ex");
  ADVS_CHECK_LAUNCH("groupnorm_apply_ex");
  return ADVS_OK;
}

int advs_groupnorm_apply_ex16(const void* x, const void* x_lo, int B, int HW, int C, const float* scale_shift,
                              const void* residual, const float* emb, int emb_stride, int act, void* y, int y_dtype,
                              void* stream) {
  ADVS_CHECK_ARG(x && y && scale_shift && B > 0 && HW > 0 && C % 8 == 0 && act >= 0 && act <= 2, "groupnorm_apply_ex16: bad args");
  ADVS_CHECK_ARG(y_dtype == ADVS_BF16 || y_dtype == ADVS_F16, "groupnorm_apply_ex16: y_dtype must be ADVS_BF16 or ADVS_F16");
  ADVS_CHECK_ARG(((uintptr_t)x_lo % 8) == 0, "groupnorm_apply_ex16: x_lo must be 8-byte aligned");
  size_t npix = (size_t)B * HW;
  size_t total = npix * (C / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  using T = __nv_bfloat16;
  if (y_dtype == ADVS_F16)
    k_gn_apply_ex<T, true><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, HW, C, npix, scale_shift, (const T*)residual, emb,
                                                             emb_stride, act, (T*)y, (const uint8_t*)x_lo);
  else
    k_gn_apply_ex<T, false><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, HW, C, npix, scale_shift, (const T*)residual, emb,
                                                              emb_stride, act, (T*)y, (const uint8_t*)x_lo);
  ADVS_CHECK_LAUNCH("groupnorm_apply_ex16");
  return ADVS_OK;
}

int advs_activation(const void* x, void* y, size_t n, int act, int dtype, void* stream) {
  ADVS_CHECK_ARG(x && y && n > 0 && n % 8 == 0 && act >= 0 && act <= 2, "activation: bad args (n%%8)");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (k_act<float><<<blocks_for(n / 8), 256, 0, st>>>((const float*)x, (float*)y, n, act)),
             (k_act<__nv_bfloat16><<<blocks_for(n / 8), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, act)),
             "activation");
  ADVS_CHECK_LAUNCH("activation");
  return ADVS_OK;
}

int advs_pos_encoding_ex(const int64_t* t, int nt, const float* inv_freq, int half, const int64_t* labels,
                         const float* label_emb, int n_labeled, float* out, void* stream) {
  ADVS_CHECK_ARG(t && inv_freq && out && nt > 0 && half > 0 && (!labels || label_emb) && n_labeled >= 0, "pos_encoding: bad args");
  k_pos_encoding<<<blocks_for((size_t)nt * half, 128), 128, 0, (cudaStream_t)stream>>>(t, inv_freq, half, labels, label_emb, nt, out,
                                                                                      n_labeled);
  ADVS_CHECK_LAUNCH("pos_encoding");
  return ADVS_OK;
}

int advs_pos_encoding(const int64_t* t, int nt, const float* inv_freq, int half, const int64_t* labels,
                      const float* label_emb, float* out, void* stream) {
  return advs_pos_encoding_ex(t, nt, inv_freq, half, labels, label_emb, nt, out, stream);
}

int advs_cfg_lerp(const float* uncond, const float* cond, float w, float* out, size_t n, void* stream) {
  ADVS_CHECK_ARG(uncond && cond && out && n > 0, "cfg_lerp: bad args");
  k_cfg_lerp<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(uncond, cond, w, out, n);
  ADVS_CHECK_LAUNCH("cfg_lerp");
  return ADVS_OK;
}

int advs_to_uint8(const float* x, uint8_t* out, size_t n, void* stream) {
  ADVS_CHECK_ARG(x && out && n > 0, "to_uint8: bad args");
  k_to_uint8<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(x, out, n);
  ADVS_CHECK_LAUNCH("to_uint8");
  return ADVS_OK;
}

}  // extern "C"
