// Bring-up experiment (not part of the product ABI): what does one 16-channel epilogue chunk cost per warp, and how
// does it scale with the number of epilogue warps?  Each warp repeats `iters` times: tcgen05.ld 16 columns -> (bias
// + activation) -> fp16 pack -> one 32-byte store per lane (pixel stride 128 bytes), in selectable parts.
//   parts bit 0: tcgen05.ld + wait     bit 1: bias (smem) + pack + store     bit 2: SiLU     bit 3: one-reciprocal-per-four SiLU (silu4)
//   bit 4: issue the next tcgen05.ld before processing the current chunk (software pipeline)
#include "../conv_common.cuh"
#include "../internal.h"
#include "../ptx.cuh"

namespace pb {

__global__ void __launch_bounds__(512, 1)
debug_epi_kernel(long long* __restrict__ cycles, __half* __restrict__ out, int parts, int iters) {
  __shared__ uint32_t tptr;
  __shared__ __align__(16) float sbias[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) sbias[threadIdx.x] = 0.01f * threadIdx.x;
  if (warp == 0) {
    tmem_alloc(&tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t_addr = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  __half* op = out + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;  // one 128-byte pixel row per thread
  uint32_t ra[16], rb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) ra[i] = rb[i] = __float_as_uint(0.25f * (float)(i + lane));
  const bool ld = parts & 1, st = parts & 2, silu = parts & 4, pairs = parts & 8, pipe = parts & 16;
  __syncthreads();
  const long long t0 = clock64();
  if (ld && pipe) tmem_ld16(t_addr, ra);
  for (int it = 0; it < iters; ++it) {
    const int c = it & 3;
#define PB_DBG_STAGE(cur, nxt)                                                                   \
  {                                                                                              \
    if (ld) {                                                                                    \
      if (pipe) {                                                                                \
        tmem_ld_wait16(cur);                                                                     \
        tmem_ld16(t_addr + (uint32_t)(((it + 1) & 31) * 16), nxt);                               \
      } else {                                                                                   \
        tmem_ld16(t_addr + (uint32_t)((it & 31) * 16), cur);                                     \
        tmem_ld_wait16(cur);                                                                     \
      }                                                                                          \
    }                                                                                            \
    float v[16];                                                                                 \
    _Pragma("unroll") for (int q = 0; q < 4; ++q) {                                              \
      const float4 b = *reinterpret_cast<const float4*>(sbias + c * 16 + 4 * q);                 \
      v[4 * q + 0] = __uint_as_float(cur[4 * q + 0]) + b.x;                                      \
      v[4 * q + 1] = __uint_as_float(cur[4 * q + 1]) + b.y;                                      \
      v[4 * q + 2] = __uint_as_float(cur[4 * q + 2]) + b.z;                                      \
      v[4 * q + 3] = __uint_as_float(cur[4 * q + 3]) + b.w;                                      \
    }                                                                                            \
    if (silu) {                                                                                  \
      if (pairs) {                                                                               \
        _Pragma("unroll") for (int i = 0; i < 4; ++i) silu4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);             \
      } else {                                                                                   \
        _Pragma("unroll") for (int i = 0; i < 16; ++i) v[i] = __fdividef(v[i], 1.f + __expf(-v[i])); \
      }                                                                                          \
    }                                                                                            \
    uint4 pk[2];                                                                                 \
    __half2* h2 = reinterpret_cast<__half2*>(pk);                                                \
    _Pragma("unroll") for (int j = 0; j < 8; ++j) h2[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]); \
    if (st) st_global_256(op + c * 16, pk[0], pk[1]);                                            \
    else if (pk[0].x == 0x12345678u) op[0] = __float2half(1.f); /* keep the math alive */        \
  }
    if (it & 1) PB_DBG_STAGE(rb, ra) else PB_DBG_STAGE(ra, rb)
#undef PB_DBG_STAGE
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tptr, 512);
  }
}

}  // namespace pb

extern "C" int pb_debug_epi_bench(long long* cycles /*[grid][16]*/, void* out /*half [grid*threads*64]*/, int nwarps,
                                  int parts, int iters, void* stream) {
  using namespace pb;
  debug_epi_kernel<<<num_sms(), nwarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      cycles, reinterpret_cast<__half*>(out), parts, iters);
  PB_CUDA(cudaGetLastError());
  return 0;
}
