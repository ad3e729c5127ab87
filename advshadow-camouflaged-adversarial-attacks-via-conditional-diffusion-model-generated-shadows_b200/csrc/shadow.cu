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

__global__ void k_shadow_composite_generated(const float* __restrict__ img, const float* __restrict__ xfin,
                                             const float* __restrict__ centers, const float* __restrict__ radii,
                                             const float* __restrict__ fmask, int Cm, float* __restrict__ out, int B,
                                             int C, int H, int W) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t hw = (size_t)H * W;
  size_t n = (size_t)B * C * hw;
  if (i >= n) return;
  size_t pix = i % hw;
  int w = (int)(pix % W), h = (int)(pix / W);
  int c = (int)((i / hw) % C);
  int b = (int)(i / (hw * C));
  float sm = disk_mask_value(w, h, centers[2 * b], centers[2 * b + 1], radii[b]);
  float fm = fmask[((size_t)b * Cm + (Cm == 1 ? 0 : c)) * hw + pix];
  float m = __fmul_rn(sm, fm);
  float x = img[i];
  float keep = __fmul_rn(x, __fsub_rn(1.f, m));
  float g = clamp01(xfin[i]);  // np.clip(generated, 0, 1), main.py:135
  out[i] = clamp01(__fadd_rn(keep, __fmul_rn(g, m)));
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
      if (v > bv) { bv = v; best = c; }
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

int advs_shadow_composite_generated(const float* img, const float* x_final, const float* centers, const float* radii,
                                    const float* feature_mask, int Cm, float* out, int B, int C, int H, int W,
                                    void* stream) {
  ADVS_CHECK_ARG(img && x_final && centers && radii && feature_mask && out, "shadow_composite_generated: null pointer");
  ADVS_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && (Cm == 1 || Cm == C), "shadow_composite_generated: bad shape");
  size_t n = (size_t)B * C * H * W;
  k_shadow_composite_generated<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      img, x_final, centers, radii, feature_mask, Cm, out, B, C, H, W);
  ADVS_CHECK_LAUNCH("shadow_composite_generated");
  return ADVS_OK;
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
