// Fused convolution epilogue shared by the SIMT and tcgen05 implicit-GEMM kernels.
//   v += bias[n] + temb[b, n] + residual[m, n]  ->  NHWC store, or the attention q/k/v^T split.
#pragma once
#include "common.cuh"

namespace advs {

// scalar path: one accumulator element (pixel m of image b at in-image index t, channel n)
template <typename T>
__device__ __forceinline__ void epilogue_store1(const EpilogueParams& e, size_t m, int b, int t, int n, float v) {
  if (e.bias) v += e.bias[n];
  if (e.temb) v += e.temb[(size_t)b * e.temb_stride + n];
  if (e.out_mode == 0) {
    if (e.residual) v += to_f(reinterpret_cast<const T*>(e.residual)[m * e.Cout + n]);
    if constexpr (sizeof(T) == 2) {
      if (e.y_lo) {   // "wide" pre-norm storage, see common.cuh
        const uint32_t u = wide_round_bits(v);
        reinterpret_cast<uint16_t*>(e.y)[m * e.Cout + n] = (uint16_t)(u >> 16);
        e.y_lo[m * e.Cout + n] = (uint8_t)(((u >> 8) & 0xFFu) ^ 0x80u);
        return;
      }
    }
    reinterpret_cast<T*>(e.y)[m * e.Cout + n] = from_f<T>(v);
  } else if (e.out_mode == 2) {
    if (n < e.cout_valid) reinterpret_cast<float*>(e.y)[((size_t)b * e.cout_valid + n) * e.HW + t] = v;
  } else {
    int head = n / (3 * e.dh);
    int r = n - head * 3 * e.dh;
    int which = r / e.dh;
    int d = r - which * e.dh;
    size_t bh = (size_t)b * e.heads + head;
    if (which == 0) reinterpret_cast<T*>(e.q)[(bh * e.HW + t) * e.dh_pad + d] = from_f<T>(v * e.qk_scale);
    else if (which == 1) reinterpret_cast<T*>(e.k)[(bh * e.HW + t) * e.dh_pad + d] = from_f<T>(v * e.qk_scale);
    else reinterpret_cast<T*>(e.vt)[(bh * e.dh_pad + d) * e.HW + t] = from_f<T>(v);
  }
}

}  // namespace advs
