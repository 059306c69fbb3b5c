// Bring-up experiment (not part of the product ABI): how does tcgen05.mma address a K-major SWIZZLE_128B operand
// whose start is NOT 1024-byte aligned (shifted by whole 128-byte rows) and whose 8-row groups are not 1024 bytes
// apart?  Result decides whether conv taps can be served from one shared halo tile (see DESIGN.md "next").
#include "../internal.h"
#include "../ptx.cuh"

namespace pb {

__global__ void __launch_bounds__(128, 1)
debug_umma_shift_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        float* __restrict__ D, int shift_rows, int sbo_bytes, int base_off_mode, int row_bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_s = smem;               // 192 rows x row_bytes
  uint8_t* b_s = smem + 192 * 128;   // 64 rows x row_bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 64 * 128);
  uint64_t* bar2 = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 192 * row_bytes + 64 * row_bytes);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(a_s)), "l"(reinterpret_cast<uint64_t>(&tmap_a)), "r"(smem_u32(bar)), "r"(0), "r"(0)
        : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(b_s)), "l"(reinterpret_cast<uint64_t>(&tmap_b)), "r"(smem_u32(bar)), "r"(0), "r"(0)
        : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(a_s) + (uint32_t)shift_rows * (uint32_t)row_bytes;
    uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
                     (1ull << 46) | ((row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull)) << 61);
    if (base_off_mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = umma_desc_kmajor(smem_u32(b_s), (uint32_t)row_bytes);
    const uint32_t idesc = umma_idesc_f16(64, 0);
    for (int k = 0; k < row_bytes / 32; ++k) umma_f16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k != 0);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 64);
}

}  // namespace pb

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int pb_debug_umma_shift(const void* A /*half [192][64]*/, const void* B /*half [64][64]*/,
                                   float* D /*[128][64]*/, int shift_rows, int sbo_bytes, int base_off_mode,
                                   int row_bytes, void* stream) {
  using namespace pb;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  PB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn2 enc = reinterpret_cast<EncodeTiledFn2>(p);
  CUtensorMap ma, mb;
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint32_t es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, 192};
    cuuint64_t str[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, 192};
    PB_CHECK(enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(A), dims, str, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode A failed");
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, 64};
    cuuint64_t str[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)row_bytes / 2, 64};
    PB_CHECK(enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(B), dims, str, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "encode B failed");
  }
  const size_t smem = 192 * 128 + 64 * 128 + 64 + 1024;
  PB_CUDA(cudaFuncSetAttribute(debug_umma_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  debug_umma_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(ma, mb, D, shift_rows, sbo_bytes,
                                                                              base_off_mode, row_bytes);
  PB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// Micro-benchmark: cycles per UMMA (M=128, K=16, N given) when A comes from shared memory (SS) versus when A is first
// copied smem -> TMEM with tcgen05.cp.128x256b and the MMA reads it from TMEM (TS).  Decides whether the conv kernels
// should stage A through TMEM.  One CTA per SM, `iters` back-to-back k-steps on fixed (garbage) operands.
// ------------------------------------------------------------------------------------------------------------
namespace pb {
__global__ void __launch_bounds__(128, 1) debug_umma_rate_kernel(long long* out, int N, int mode, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem);               // 16 KB: 128 rows x 128 B
    const uint32_t b_addr = smem_u32(smem + 16 * 1024);   // 32 KB: 256 rows x 128 B
    const uint64_t adesc = umma_desc_kmajor(a_addr, 128);
    const uint64_t bdesc = umma_desc_kmajor(b_addr, 128);
    const uint32_t idesc = umma_idesc_f16(N, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int k = it & 3;
      if (mode == 0) {
        umma_f16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
      } else {
        const uint32_t a_t = tmem + 256u + (uint32_t)((it & 7) * 8);  // 8 rotating A buffers of 8 columns
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(a_t), "l"(adesc + (uint64_t)(2 * k)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
            "r"(a_t), "l"(bdesc + (uint64_t)(2 * k)), "r"(idesc), "r"(1u)
            : "memory");
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}
}  // namespace pb

extern "C" int pb_debug_umma_rate(long long* out, int N, int mode, int iters, void* stream) {
  using namespace pb;
  const size_t smem = 64 * 1024 + 64 + 1024;
  PB_CUDA(cudaFuncSetAttribute(debug_umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  debug_umma_rate_kernel<<<num_sms(), 128, smem, static_cast<cudaStream_t>(stream)>>>(out, N, mode, iters);
  PB_CUDA(cudaGetLastError());
  return 0;
}
