// Generic SIMT implicit-GEMM convolution (fp32 accumulate, FFMA) + batched NT GEMM + row softmax.
// This is the fp32 "<=1e-4" mode of the sampler and the bring-up/validation path for the bf16
// tcgen05 kernels; it is NOT the throughput path.
//   reference ops replaced: nn.Conv2d at dm1:73,86,90,114-115,134,148; einsum/softmax dm1:122-124.
#include "common.cuh"
#include "epilogue.cuh"

namespace advs {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtConvArgs {
  int B, H, W, Cout, stride, nseg;
  const void* x[3];
  const void* w[3];
  int C[3];
  int taps[3];
  EpilogueParams epi;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float* f);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* f) {
  float4 t = *reinterpret_cast<const float4*>(p);
  f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

template <typename T>
__global__ void __launch_bounds__(256) k_conv_simt(SimtConvArgs a) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const size_t M = (size_t)a.B * a.H * a.W;
  const size_t m0 = (size_t)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;

  // loader roles
  const int lrow = tid / 4;  // pixel (A) or cout (B) within the tile
  const int kq = tid % 4;    // channel quad within the K step
  const size_t lm = m0 + lrow;
  const bool lm_ok = lm < M;
  int lb = 0, lho = 0, lwo = 0;
  if (lm_ok) {
    lwo = (int)(lm % a.W);
    lho = (int)((lm / a.W) % a.H);
    lb = (int)(lm / ((size_t)a.W * a.H));
  }
  const int ln = n0 + lrow;
  const bool ln_ok = ln < a.Cout;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int s = 0; s < a.nseg; ++s) {
    const T* xs = reinterpret_cast<const T*>(a.x[s]);
    const T* ws = reinterpret_cast<const T*>(a.w[s]);
    const int C = a.C[s], taps = a.taps[s];
    const int stride = (s == 0) ? a.stride : 1;
    const int Hin = a.H * stride, Win = a.W * stride;
    for (int tap = 0; tap < taps; ++tap) {
      int hi = lho * stride, wi = lwo * stride;
      if (taps == 9) { hi += tap / 3 - 1; wi += tap % 3 - 1; }
      const bool pix_ok = lm_ok && hi >= 0 && hi < Hin && wi >= 0 && wi < Win;
      const T* xp = xs + (((size_t)lb * Hin + hi) * Win + wi) * C;
      const T* wp = ws + ((size_t)ln * taps + tap) * C;
      for (int c0 = 0; c0 < C; c0 += SBK) {
        const int c = c0 + kq * 4;
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (pix_ok && c < C) load4<T>(xp + c, av);
        if (ln_ok && c < C) load4<T>(wp + c, bv);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          As[kq * 4 + i][lrow] = av[i];
          Bs[kq * 4 + i][lrow] = bv[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
          float4 ar = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          float4 br = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          float aa[4] = {ar.x, ar.y, ar.z, ar.w}, bb[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
      }
    }
  }
  const int HW = a.H * a.W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    size_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int b = (int)(m / HW), t = (int)(m % HW);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < a.Cout) epilogue_store1<T>(a.epi, m, b, t, n, acc[i][j]);
    }
  }
}

// ---- batched C[bh] = A[bh] * B[bh]^T (both row-major, K contiguous), fp32 accumulate ------------
struct GemmNTArgs {
  const void* A; size_t sA; int lda;   // [batch][M][K]
  const void* Bm; size_t sB; int ldb;  // [batch][N][K]
  void* C; int ldc;                    // element offset of batch bh: (bh/inner)*sC_outer + (bh%inner)*sC_inner
  size_t sC_outer, sC_inner; int inner;
  int M, N, K;
};

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) k_gemm_nt(GemmNTArgs g) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Bs[SBK][SBN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int bh = blockIdx.z;
  const TA* A = reinterpret_cast<const TA*>(g.A) + (size_t)bh * g.sA;
  const TB* Bp = reinterpret_cast<const TB*>(g.Bm) + (size_t)bh * g.sB;
  TC* C = reinterpret_cast<TC*>(g.C) + (size_t)(bh / g.inner) * g.sC_outer + (size_t)(bh % g.inner) * g.sC_inner;
  const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
  const int lrow = tid / 4, kq = tid % 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < g.K; k0 += SBK) {
    const int k = k0 + kq * 4;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + lrow < g.M && k < g.K) load4<TA>(A + (size_t)(m0 + lrow) * g.lda + k, av);
    if (n0 + lrow < g.N && k < g.K) load4<TB>(Bp + (size_t)(n0 + lrow) * g.ldb + k, bv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[kq * 4 + i][lrow] = av[i];
      Bs[kq * 4 + i][lrow] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      float4 ar = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 br = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float aa[4] = {ar.x, ar.y, ar.z, ar.w}, bb[4] = {br.x, br.y, br.z, br.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < g.N) C[(size_t)m * g.ldc + n] = from_f<TC>(acc[i][j]);
    }
  }
}

// in-place softmax over the last axis; one warp per row (dm1:123)
__global__ void k_softmax_rows(float* __restrict__ s, size_t rows, int T) {
  size_t row = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* r = s + row * T;
  float mx = -INFINITY;
  for (int i = lane; i < T; i += 32) mx = fmaxf(mx, r[i]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < T; i += 32) {
    float e = expf(r[i] - mx);
    r[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  for (int i = lane; i < T; i += 32) r[i] = r[i] / sum;
}

template <typename T>
static int attention_simt_impl(const void* q, const void* k, const void* vt, void* o, int B, int heads, int Tn, int dh,
                               float* scores, cudaStream_t st) {
  const int BH = B * heads;
  ADVS_CHECK_ARG(BH <= 65535, "attention_simt: B*heads must be <= 65535");
  {  // S = Q K^T
    GemmNTArgs g;
    g.A = q; g.sA = (size_t)Tn * dh; g.lda = dh;
    g.Bm = k; g.sB = (size_t)Tn * dh; g.ldb = dh;
    g.C = scores; g.ldc = Tn; g.sC_outer = (size_t)Tn * Tn; g.sC_inner = 0; g.inner = 1;
    g.M = Tn; g.N = Tn; g.K = dh;
    dim3 grid((Tn + SBM - 1) / SBM, (Tn + SBN - 1) / SBN, BH);
    k_gemm_nt<T, T, float><<<grid, 256, 0, st>>>(g);
    ADVS_CHECK_LAUNCH("attention_simt/qk");
  }
  {
    size_t rows = (size_t)BH * Tn;
    size_t blocks = (rows * 32 + 255) / 256;
    k_softmax_rows<<<(unsigned)blocks, 256, 0, st>>>(scores, rows, Tn);
    ADVS_CHECK_LAUNCH("attention_simt/softmax");
  }
  {  // O[b, t, head*dh + d] = sum_s P[bh, t, s] * VT[bh, d, s]
    GemmNTArgs g;
    g.A = scores; g.sA = (size_t)Tn * Tn; g.lda = Tn;
    g.Bm = vt; g.sB = (size_t)dh * Tn; g.ldb = Tn;
    g.C = o; g.ldc = heads * dh; g.sC_outer = (size_t)Tn * heads * dh; g.sC_inner = dh; g.inner = heads;
    g.M = Tn; g.N = dh; g.K = Tn;
    dim3 grid((Tn + SBM - 1) / SBM, (dh + SBN - 1) / SBN, BH);
    k_gemm_nt<float, T, T><<<grid, 256, 0, st>>>(g);
    ADVS_CHECK_LAUNCH("attention_simt/pv");
  }
  return ADVS_OK;
}

}  // namespace advs

using namespace advs;

extern "C" {

int advs_conv_simt(const advs_conv_params* p, void* stream) {
  int rc = validate_conv(p, "conv_simt");
  if (rc) return rc;
  ADVS_CHECK_ARG(p->stats_partial == nullptr, "conv_simt: stats_partial is only produced by the sm100 kernel");
  ADVS_CHECK_ARG(p->up_phase == 0, "conv_simt: upsample phases are an sm100-kernel feature (use upsample + 3x3 conv)");
  SimtConvArgs a;
  a.B = p->B; a.H = p->H; a.W = p->W; a.Cout = p->Cout; a.stride = p->stride; a.nseg = p->nseg;
  for (int s = 0; s < 3; ++s) {
    a.x[s] = s < p->nseg ? p->seg[s].x : nullptr;
    a.w[s] = s < p->nseg ? p->seg[s].w : nullptr;
    a.C[s] = s < p->nseg ? p->seg[s].C : 0;
    a.taps[s] = s < p->nseg ? p->seg[s].taps : 0;
  }
  a.epi = make_epilogue(*p);
  size_t M = (size_t)p->B * p->H * p->W;
  dim3 grid((unsigned)((M + SBM - 1) / SBM), (p->Cout + SBN - 1) / SBN);
  ADVS_CHECK_ARG(grid.y <= 65535, "conv_simt: Cout too large");
  if (p->dtype == ADVS_F32) k_conv_simt<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else k_conv_simt<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  ADVS_CHECK_LAUNCH("conv_simt");
  return ADVS_OK;
}

size_t advs_attention_simt_workspace_bytes(int B, int heads, int T) {
  if (B <= 0 || heads <= 0 || T <= 0) return 0;
  return (size_t)B * heads * T * T * sizeof(float);
}

int advs_attention_simt(const void* q, const void* k, const void* vt, void* o, int B, int heads, int T, int dh,
                        void* workspace, size_t workspace_bytes, int dtype, void* stream) {
  ADVS_CHECK_ARG(q && k && vt && o && workspace, "attention_simt: null pointer");
  ADVS_CHECK_ARG(B > 0 && heads > 0 && T > 0 && dh > 0 && dh % 4 == 0 && T % 4 == 0, "attention_simt: bad shape (T, dh multiples of 4)");
  ADVS_CHECK_ARG(workspace_bytes >= advs_attention_simt_workspace_bytes(B, heads, T), "attention_simt: workspace too small");
  if (dtype == ADVS_F32) return attention_simt_impl<float>(q, k, vt, o, B, heads, T, dh, (float*)workspace, (cudaStream_t)stream);
  if (dtype == ADVS_BF16) return attention_simt_impl<__nv_bfloat16>(q, k, vt, o, B, heads, T, dh, (float*)workspace, (cudaStream_t)stream);
  ADVS_CHECK_ARG(false, "attention_simt: bad dtype");
}

}  // extern "C"
