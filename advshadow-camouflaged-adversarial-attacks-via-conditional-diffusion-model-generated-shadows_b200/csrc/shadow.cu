// K9: shadow mask construction, 5x5 Gaussian blur, masked compositing, attack-success flags.
//   reference: create_shadow_mask dm2:552-570 (= ts:156-174), apply_gaussian_blur ts:147-153,
//   apply_shadow dm2:615-654 / ts:224-266, compute_asr ASR_fast.py:101-126.
// All arithmetic replays the reference's fp32 operation order with explicit round-to-nearest
// intrinsics (no FMA contraction) so that masks and composites are bit-exact.
#include "common.cuh"

namespace advs {

__device__ __forceinline__ float disk_mask_value(int w, int h, float c0, float c1, float r) {
  // dist = sqrt((X - c[0])**2 + (Y - c[1])**2);  mask = (dist <= r).float()      (dm2:567-569)
  float dx = __fsub_rn((float)w, c0);
  float dy = __fsub_rn((float)h, c1);
  float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  return d <= r ? 1.f : 0.f;
}

__global__ void k_disk_mask(const float* __restrict__ centers, const float* __restrict__ radii, int B, int H, int W,
                            float* __restrict__ mask) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)B * H * W;
  if (i >= n) return;
  int w = (int)(i % W), h = (int)((i / W) % H), b = (int)(i / ((size_t)W * H));
  mask[i] = disk_mask_value(w, h, centers[2 * b], centers[2 * b + 1], radii[b]);
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// cv2.GaussianBlur(src,(5,5),0): fixed table [1,4,6,4,1]/16, rows then columns, BORDER_REFLECT_101
__global__ void k_blur5(const float* __restrict__ src, float* __restrict__ dst, int B, int H, int W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)B * H * W;
  if (i >= n) return;
  const float kw[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  int w = (int)(i % W), h = (int)((i / W) % H), b = (int)(i / ((size_t)W * H));
  const float* s = src + (size_t)b * H * W;
  float acc = 0.f;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    int hy = reflect101(h + dy - 2, H);
    float row = 0.f;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      int wx = reflect101(w + dx - 2, W);
      row = __fadd_rn(row, __fmul_rn(kw[dx], s[(size_t)hy * W + wx]));
    }
    acc = __fadd_rn(acc, __fmul_rn(kw[dy], row));
  }
  dst[i] = acc;
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__global__ void k_shadow_composite(const float* __restrict__ img, const float* __restrict__ smask,
                                   const float* __restrict__ fmask, int Cm, const float* __restrict__ adv,
                                   float one_minus_intensity, float* __restrict__ shadowed_out,
                                   float* __restrict__ out, int B, int C, int H, int W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t hw = (size_t)H * W;
  size_t n = (size_t)B * C * hw;
  if (i >= n) return;
  size_t pix = i % hw;
  int c = (int)((i / hw) % C);
  int b = (int)(i / (hw * C));
  float fm = fmask[((size_t)b * Cm + (Cm == 1 ? 0 : c)) * hw + pix];
  float m = __fmul_rn(smask[(size_t)b * hw + pix], fm);          // combined_mask      (dm2:642)
  float x = img[i];
  float keep = __fmul_rn(x, __fsub_rn(1.f, m));                   // image*(1-m)
  float shadowed = __fadd_rn(keep, __fmul_rn(m, __fmul_rn(x, one_minus_intensity)));  // (dm2:645)
  if (shadowed_out) shadowed_out[i] = shadowed;
  if (out) {
    float a = adv ? adv[i] : shadowed;
    out[i] = clamp01(__fadd_rn(keep, __fmul_rn(a, m)));          // (dm2:650-653)
  }
}

// Fused tail of the shadow sampler (one pass over the images instead of update + mask + [blur] + composite):
//   x'  = DDIM update of the last step (dm1:457-472; skipped when eps == nullptr: x is then already final)
//   m   = disk(centre, radius) [5x5 Gaussian-blurred: ts:147-153, 244-247 / dt:851-854] * feature_mask
//   out = clamp(img*(1-m) + clip(x',0,1)*m, 0, 1)                                        (dm2:650-653)
// One thread per VEC consecutive pixels of a row and all channels: the mask is built once per pixel, the images move
// as 16-byte vectors, indices are 32-bit.  Every value is formed in the reference's fp32 operation order (the
// blurred {0,1} mask is a sum of dyadic rationals, exact in any order), so the result is bit-identical to the
// separate kernels above.  x and x_out may alias.
template <int VEC, bool BLUR>
__global__ void k_ddim_final_composite(const float* x, const float* __restrict__ eps, float* x_out,
                                       const float* __restrict__ coef, const int32_t* __restrict__ step_dev, int clip,
                                       const float* __restrict__ img, const float* __restrict__ centers,
                                       const float* __restrict__ radii, const float* __restrict__ fmask, int Cm,
                                       float* __restrict__ out, int C, int H, int W) {
  pdl_sync();
  const int b = blockIdx.y;
  const int wq = W / VEC;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * wq) return;
  const int h = q / wq, w0 = (q - h * wq) * VEC;
  const float c0 = centers[2 * b], c1 = centers[2 * b + 1], r = radii[b];
  float sm[VEC];
  if (BLUR) {
    const float kw[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    int wx[VEC + 4];
#pragma unroll
    for (int i = 0; i < VEC + 4; ++i) wx[i] = reflect101(w0 + i - 2, W);
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const int hy = reflect101(h + dy - 2, H);
      float d[VEC + 4];
#pragma unroll
      for (int i = 0; i < VEC + 4; ++i) d[i] = disk_mask_value(wx[i], hy, c0, c1, r);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float row = 0.f;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) row = __fadd_rn(row, __fmul_rn(kw[dx], d[k + dx]));
        acc[k] = __fadd_rn(acc[k], __fmul_rn(kw[dy], row));
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) sm[k] = acc[k];
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) sm[k] = disk_mask_value(w0 + k, h, c0, c1, r);
  }
  float s1 = 0.f, sa = 1.f, sp = 1.f, cdir = 0.f;
  if (eps) {
    const float* cf = coef + 8 * (size_t)(*step_dev);
    s1 = cf[0]; sa = cf[1]; sp = cf[2]; cdir = cf[3];
  }
  const size_t hw = (size_t)H * W;
  const size_t pix = (size_t)h * W + w0;
  float fm[VEC];
  for (int c = 0; c < C; ++c) {
    const size_t i = ((size_t)b * C + c) * hw + pix;
    float xv[VEC], ev[VEC], iv[VEC], ov[VEC];
    if (c == 0 || Cm != 1) {
      const float* fp = fmask + ((size_t)b * Cm + (Cm == 1 ? 0 : c)) * hw + pix;
      if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(fp); fm[0] = t.x; fm[1] = t.y; fm[2] = t.z; fm[3] = t.w; }
      else fm[0] = fp[0];
    }
    if (VEC == 4) {
      float4 t = *reinterpret_cast<const float4*>(x + i);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
      t = *reinterpret_cast<const float4*>(img + i);
      iv[0] = t.x; iv[1] = t.y; iv[2] = t.z; iv[3] = t.w;
      if (eps) { t = *reinterpret_cast<const float4*>(eps + i); ev[0] = t.x; ev[1] = t.y; ev[2] = t.z; ev[3] = t.w; }
    } else {
      xv[0] = x[i]; iv[0] = img[i];
      if (eps) ev[0] = eps[i];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float xf = xv[k];
      if (eps) {
        float x0 = __fdiv_rn(__fsub_rn(xv[k], __fmul_rn(s1, ev[k])), sa);
        if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        xf = __fadd_rn(__fmul_rn(sp, x0), __fmul_rn(cdir, ev[k]));   // sigma = 0 on this path (eta = 0)
        xv[k] = xf;
      }
      const float m = __fmul_rn(sm[k], fm[k]);
      const float keep = __fmul_rn(iv[k], __fsub_rn(1.f, m));
      ov[k] = clamp01(__fadd_rn(keep, __fmul_rn(clamp01(xf), m)));
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(out + i) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      if (eps) *reinterpret_cast<float4*>(x_out + i) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    } else {
      out[i] = ov[0];
      if (eps) x_out[i] = xv[0];
    }
  }
}

__global__ void k_advance_step_s(int32_t* step_dev, int advance) {
  pdl_sync();
  *step_dev += advance;
}

static int launch_final_composite(const float* x, const float* eps, float* x_out, const float* coef, int32_t* step_dev,
                                  int advance, int clip, const float* img, const float* centers, const float* radii,
                                  const float* fmask, int Cm, int blur, float* out, int B, int C, int H, int W,
                                  cudaStream_t st, const char* who) {
  const bool vec = W % 4 == 0 && (((uintptr_t)x | (uintptr_t)eps | (uintptr_t)x_out | (uintptr_t)img | (uintptr_t)fmask |
                                   (uintptr_t)out) % 16) == 0;
  const int quads = H * (vec ? W / 4 : W);
  dim3 grid((quads + 127) / 128, B);
#define ADVS_FC(V, BL) launch_pdl(k_ddim_final_composite<V, BL>, grid, dim3(128), 0, st, x, eps, x_out, coef, (const int32_t*)step_dev, clip, \
                                  img, centers, radii, fmask, Cm, out, C, H, W)
  if (vec) { if (blur) ADVS_FC(4, true); else ADVS_FC(4, false); }
  else { if (blur) ADVS_FC(1, true); else ADVS_FC(1, false); }
#undef ADVS_FC
  ADVS_CHECK_LAUNCH(who);
  if (eps && advance) {
    launch_pdl(k_advance_step_s, dim3(1), dim3(1), 0, st, step_dev, advance);
    ADVS_CHECK_LAUNCH(who);
  }
  return ADVS_OK;
}

__global__ void k_success_flags(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B,
                                int classes, uint8_t* __restrict__ flags, unsigned long long* __restrict__ counts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int ok = 0;
  if (b < B) {
    const float* l = logits + (size_t)b * classes;
    int best = 0;
    float bv = l[0];
    for (int c = 1; c < classes; ++c) {
      float v = l[c];
      if (v > bv || (v != v && bv == bv)) { bv = v; best = c; }   // torch.max: the first NaN wins and sticks
    }
    ok = ((int64_t)best != labels[b]) ? 1 : 0;
    flags[b] = (uint8_t)ok;
  }
  unsigned ballot = __ballot_sync(0xffffffffu, ok);
  unsigned active = __ballot_sync(0xffffffffu, b < B);
  if ((threadIdx.x & 31) == 0 && active) {
    atomicAdd(&counts[0], (unsigned long long)__popc(ballot));
    atomicAdd(&counts[1], (unsigned long long)__popc(active));
  }
}

}  // namespace advs

using namespace advs;

extern "C" {

int advs_shadow_disk_mask(const float* centers, const float* radii, int B, int H, int W, float* mask, void* stream) {
  ADVS_CHECK_ARG(centers && radii && mask && B > 0 && H > 0 && W > 0, "shadow_disk_mask: bad args");
  size_t n = (size_t)B * H * W;
  k_disk_mask<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(centers, radii, B, H, W, mask);
  ADVS_CHECK_LAUNCH("shadow_disk_mask");
  return ADVS_OK;
}

int advs_gaussian_blur5(const float* src, float* dst, int B, int H, int W, void* stream) {
  ADVS_CHECK_ARG(src && dst && src != dst && B > 0 && H > 0 && W > 0, "gaussian_blur5: bad args (no in-place)");
  size_t n = (size_t)B * H * W;
  k_blur5<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, B, H, W);
  ADVS_CHECK_LAUNCH("gaussian_blur5");
  return ADVS_OK;
}

int advs_shadow_composite(const float* img, const float* shadow_mask, const float* feature_mask, int Cm,
                          const float* adv, float one_minus_intensity, float* shadowed_out, float* out, int B, int C,
                          int H, int W, void* stream) {
  ADVS_CHECK_ARG(img && shadow_mask && feature_mask && (shadowed_out || out), "shadow_composite: null pointer");
  ADVS_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && (Cm == 1 || Cm == C), "shadow_composite: bad shape (Cm must be 1 or C)");
  size_t n = (size_t)B * C * H * W;
  k_shadow_composite<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      img, shadow_mask, feature_mask, Cm, adv, one_minus_intensity, shadowed_out, out, B, C, H, W);
  ADVS_CHECK_LAUNCH("shadow_composite");
  return ADVS_OK;
}

int advs_shadow_composite_generated_ex(const float* img, const float* x_final, const float* centers, const float* radii,
                                       const float* feature_mask, int Cm, int blur, float* out, int B, int C, int H, int W,
                                       void* stream) {
  ADVS_CHECK_ARG(img && x_final && centers && radii && feature_mask && out, "shadow_composite_generated: null pointer");
  ADVS_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0 && (Cm == 1 || Cm == C), "shadow_composite_generated: bad shape");
  return launch_final_composite(x_final, nullptr, nullptr, nullptr, nullptr, 0, 0, img, centers, radii, feature_mask, Cm, blur,
                                out, B, C, H, W, (cudaStream_t)stream, "shadow_composite_generated");
}

int advs_shadow_composite_generated(const float* img, const float* x_final, const float* centers, const float* radii,
                                    const float* feature_mask, int Cm, float* out, int B, int C, int H, int W,
                                    void* stream) {
  return advs_shadow_composite_generated_ex(img, x_final, centers, radii, feature_mask, Cm, 0, out, B, C, H, W, stream);
}

int advs_ddim_step_composite(const float* x, const float* eps, float* x_out, const float* coef, int32_t* step_dev,
                             int advance, int clip_denoised, const float* img, const float* centers, const float* radii,
                             const float* feature_mask, int Cm, int blur, float* out, int B, int C, int H, int W,
                             void* stream) {
  ADVS_CHECK_ARG(x && eps && x_out && coef && step_dev, "ddim_step_composite: null sampler pointer");
  ADVS_CHECK_ARG(img && centers && radii && feature_mask && out, "ddim_step_composite: null composite pointer");
  ADVS_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0 && (Cm == 1 || Cm == C), "ddim_step_composite: bad shape");
  return launch_final_composite(x, eps, x_out, coef, step_dev, advance, clip_denoised, img, centers, radii, feature_mask, Cm,
                                blur, out, B, C, H, W, (cudaStream_t)stream, "ddim_step_composite");
}

int advs_success_flags(const float* logits, const int64_t* labels, int B, int classes, uint8_t* flags,
                       int64_t* counts, void* stream) {
  ADVS_CHECK_ARG(logits && labels && flags && counts && B > 0 && classes > 0, "success_flags: bad args");
  k_success_flags<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, labels, B, classes, flags,
                                                                     (unsigned long long*)counts);
  ADVS_CHECK_LAUNCH("success_flags");
  return ADVS_OK;
}

}  // extern "C"
