// Hardware self-test: does a K-major SWIZZLE_128B UMMA descriptor whose start address is shifted by a
// whole number of 128-byte rows (not a multiple of 8 rows) read the rows a TMA box wrote?  This decides
// whether one halo tile in shared memory can serve all 3x3 taps of the implicit-GEMM conv.
//   A: 272 rows x 64 bf16 written with the address-based 128B swizzle (exactly TMA's pattern),
//      A[r][c] = r + c/64.  B: 64x64 identity.  D[i][j] = A[r0 + i][j] if the shifted read is right.
#include "common.cuh"
#include "sm100.cuh"

namespace advs {
using namespace sm100;

__global__ void __launch_bounds__(128, 1) k_selftest_umma_row_shift(int r0, int base_offset_mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);                 // 272 rows * 128 B
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + 272 * 128);    // 64 rows * 128 B (offset 34 KB, 1 KB aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 272 * 128 + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 272 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    int chunk = (c / 8) ^ (r & 7);
    A[r * 64 + chunk * 8 + (c % 8)] = __float2bfloat16_rn((float)(r % 256) + (float)c / 64.f);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    int r = i / 64, c = i % 64;
    int chunk = (c / 8) ^ (r & 7);
    Bm[r * 64 + chunk * 8 + (c % 8)] = __float2bfloat16_rn(r == c ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(A) + (uint32_t)r0 * 128u;
    uint64_t adesc = umma_desc_k_sw128(a_addr);
    if (base_offset_mode) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(Bm));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem_base);
}

}  // namespace advs

extern "C" int advs_selftest_umma_row_shift(int r0, int base_offset_mode, float* out, void* stream) {
  ADVS_CHECK_ARG(out && r0 >= 0 && r0 <= 140, "selftest_umma_row_shift: bad args");
  const int smem = 272 * 128 + 64 * 128 + 64 + 1024;
  if (advs::first_use_on_device(advs::kOnceSelftest))
    cudaFuncSetAttribute(advs::k_selftest_umma_row_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  advs::k_selftest_umma_row_shift<<<1, 128, smem, (cudaStream_t)stream>>>(r0, base_offset_mode, out);
  ADVS_CHECK_LAUNCH("selftest_umma_row_shift");
  return ADVS_OK;
}
