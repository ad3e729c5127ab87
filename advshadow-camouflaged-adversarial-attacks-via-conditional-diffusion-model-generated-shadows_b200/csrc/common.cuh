// Shared helpers for the advshadow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "advshadow_b200.h"

namespace advs {

void set_error(const char* fmt, ...);

// Kernel attributes (cudaFuncSetAttribute) and SM counts are per DEVICE, and one process may drive several GPUs
// (engines are cached per device index): `first_use_on_device(slot)` is true exactly once per (slot, current device).
enum OnceSlot { kOnceConv1Cta = 0, kOnceConv2Cta, kOnceConvHalo, kOnceAttn64, kOnceAttn128, kOnceAttn256, kOnceSelftest,
                kOnceSlots };
bool first_use_on_device(int slot);
void forget_first_use(int slot);   // the attribute call failed: try again on the next launch
int current_device_sms();          // multiProcessorCount of the current device (148 if the query fails)

#define ADVS_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      advs::set_error(__VA_ARGS__);    \
      return ADVS_ERR_ARG;             \
    }                                  \
  } while (0)

#define ADVS_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      advs::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return ADVS_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// A UNet step is ~335 dependent launches on one stream (inside a CUDA graph).  With the programmatic-stream-
// serialization attribute the NEXT kernel's CTAs may be scheduled -- and run their prologue: barrier init, TMEM
// allocation, descriptor prefetch, index arithmetic -- while the current kernel is still running; they block in
// griddepcontrol.wait until it has completed and its writes are visible.  Every kernel launched through
// launch_pdl() therefore executes pdl_wait() in every thread before its first global-memory access that could
// depend on an earlier kernel, and pdl_launch_dependents() right after (one grid of look-ahead).  The gain is at
// small batches, where a step is latency-bound (batch 1, 256x256: 4.2 ms per step for 335 kernels).
// Off by default: advs_set_pdl(1) (ShadowSampler does that while it captures the graph of a batch <= 2 engine in a
// single-process run) or ADVS_PDL=1 turn it on; launches are plain otherwise.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_wait(); pdl_launch_dependents(); }

// ---- storage-type helpers -------------------------------------------------------------
template <typename T> struct Vec8;  // 8 consecutive elements
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
  __device__ __forceinline__ void load_stream(const float* p) { load(p); }
  __device__ __forceinline__ void store_stream(float* p) const { store(p); }
  __device__ __forceinline__ void to_float(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ void from_float(const float* f) {
    a = make_float4(f[0], f[1], f[2], f[3]);
    b = make_float4(f[4], f[5], f[6], f[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  // streaming variants for read-once / write-once tensors (GroupNorm apply): bypass L1, evict-first in L2
  __device__ __forceinline__ void load_stream(const __nv_bfloat16* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  }
  __device__ __forceinline__ void store_stream(__nv_bfloat16* p) const {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
  __device__ __forceinline__ void to_float(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void from_float(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// bf16-mode SiLU: x*sigmoid(x) = h + h*tanh(h), h = x/2 -- one MUFU op (tanh.approx, rel. err 2^-11,
// below bf16 resolution) instead of ex2 + rcp; GN+SiLU is MUFU-bound otherwise.
__device__ __forceinline__ float silu_f(float x) {
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// accurate variant (fp32 parity mode): expf instead of the fast intrinsic
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- fused conv epilogue shared by the SIMT and the tcgen05 kernels ---------------------
struct EpilogueParams {
  const float* bias;
  const float* temb;
  int temb_stride;
  int out_mode;
  const void* residual;
  void* y;
  void* q;
  void* k;
  void* vt;
  int heads;
  int dh;
  int dh_pad;      // storage head dim of q / k / vt (>= dh)
  float qk_scale;
  int HW;    // pixels per image (T for attention)
  int Cout;
  int cout_valid;  // out_mode 2: real output channels
  uint8_t* y_lo;   // out_mode 0, optional: 8 more mantissa bits per element of y ("wide" pre-norm storage)
};

inline EpilogueParams make_epilogue(const advs_conv_params& p) {
  EpilogueParams e;
  e.bias = p.bias;
  e.temb = p.temb;
  e.temb_stride = p.temb_stride;
  e.out_mode = p.out_mode;
  e.residual = p.residual;
  e.y = p.y;
  e.q = p.q;
  e.k = p.k;
  e.vt = p.vt;
  e.heads = p.heads > 0 ? p.heads : 1;
  e.dh = p.out_mode == 1 ? p.Cout / (3 * e.heads) : 0;
  e.dh_pad = p.qkv_dh_pad > e.dh ? p.qkv_dh_pad : e.dh;
  e.qk_scale = p.qk_scale;
  e.HW = p.H * p.W;
  e.Cout = p.Cout;
  e.cout_valid = (p.out_mode == 2) ? p.cout_valid : p.Cout;
  e.y_lo = p.out_mode == 0 ? reinterpret_cast<uint8_t*>(p.y_lo) : nullptr;
  return e;
}

int validate_conv(const advs_conv_params* p, const char* who);

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- "wide" pre-norm storage: bf16 + 8 mantissa-extension bits -----------------------------------
// A tensor that is only ever read by a normalisation loses accuracy twice in bf16: once when it is stored and
// once more when the normalised value is rounded for the GEMM.  Such tensors keep a companion int8 tensor with
// the next 8 mantissa bits: v ~= bits(hi) + (lo << 8) as an integer add on the fp32 bit pattern -- IEEE bit
// patterns are monotone in the magnitude, so the add carries correctly across binades and needs no exponent
// arithmetic.  Encoding is integer-only and cheaper than the plain bf16 conversion it replaces:
//   u  = bits(v) + 0x8000          hi = u >> 16   (bf16 rounded to nearest, ties away from zero)
//   lo = int8((u >> 8) & 0xFF) - 128 = byte 1 of u, sign bit flipped      (floor of the signed remainder / 256)
// hi is also a GEMM operand and the residual for every other consumer; 0 <= |v| - |decoded| < 2^-15 |v| (2^-16 on average).
__device__ __forceinline__ uint32_t wide_round_bits(float v) { return __float_as_uint(v) + 0x8000u; }
// two rounded bit patterns -> packed bf16x2 (their upper halves)
__device__ __forceinline__ uint32_t wide_hi2(uint32_t u0, uint32_t u1) { return __byte_perm(u0, u1, 0x7632); }
// four rounded bit patterns -> four int8 extensions in one word
__device__ __forceinline__ uint32_t wide_lo4(uint32_t u0, uint32_t u1, uint32_t u2, uint32_t u3) {
  return __byte_perm(__byte_perm(u0, u1, 0x0051), __byte_perm(u2, u3, 0x0051), 0x5410) ^ 0x80808080u;
}
// element k (0..3) of a packed int8 word as (int)lo << 8
// (one PRMT: result bytes = {0, byte K, sign(byte K), sign(byte K)}; the sign-replication bit of the selector
// nibbles is a feature of the PTX instruction -- __byte_perm() only honours the low three bits of each nibble)
template <int K>
__device__ __forceinline__ int wide_lo_shifted(uint32_t lo4) {
  constexpr uint32_t sel = K == 0 ? 0x8804u : (K == 1 ? 0x9914u : (K == 2 ? 0xAA24u : 0xBB34u));
  int r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo4), "r"(0u), "r"(sel));
  return r;
}
// 8 bf16 (uint4) + 8 int8 (uint2) -> 8 floats
__device__ __forceinline__ void wide_decode8(const uint4& hi, const uint2& lo, float* f) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  f[0] = __uint_as_float((h[0] << 16) + wide_lo_shifted<0>(lo.x));
  f[1] = __uint_as_float((h[0] & 0xFFFF0000u) + wide_lo_shifted<1>(lo.x));
  f[2] = __uint_as_float((h[1] << 16) + wide_lo_shifted<2>(lo.x));
  f[3] = __uint_as_float((h[1] & 0xFFFF0000u) + wide_lo_shifted<3>(lo.x));
  f[4] = __uint_as_float((h[2] << 16) + wide_lo_shifted<0>(lo.y));
  f[5] = __uint_as_float((h[2] & 0xFFFF0000u) + wide_lo_shifted<1>(lo.y));
  f[6] = __uint_as_float((h[3] << 16) + wide_lo_shifted<2>(lo.y));
  f[7] = __uint_as_float((h[3] & 0xFFFF0000u) + wide_lo_shifted<3>(lo.y));
}

}  // namespace advs
