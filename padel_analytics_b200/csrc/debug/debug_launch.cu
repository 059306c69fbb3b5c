// Bring-up experiment (debug library only): what does the GPU front end need per kernel launch, as a function of the
// launch's shape?  Chains of n empty kernels; each stamps %globaltimer at entry of CTA 0.  Variants:
//   0: 148 x 480 threads, no dynamic smem, one int parameter
//   1: + 120 KB dynamic shared memory
//   2: + a 640-byte struct parameter
//   3: + two CUtensorMap __grid_constant__ parameters (what the conv kernels take)
//   4: variant 3 launched with the programmatic-stream-serialization attribute
#include "../internal.h"
#include "../ptx.cuh"

namespace pb {
struct BigParams {
  int v[160];
};

__global__ void __launch_bounds__(480, 1) dbg_launch_small(long long* out, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) out[i] = (long long)globaltimer_ns();
}
__global__ void __launch_bounds__(480, 1) dbg_launch_big(long long* out, int i, const __grid_constant__ BigParams p) {
  if (blockIdx.x == 0 && threadIdx.x == 0) out[i] = (long long)globaltimer_ns() + (p.v[3] & 0);
}
__global__ void __launch_bounds__(480, 1)
dbg_launch_tmap(const __grid_constant__ CUtensorMap a, const __grid_constant__ CUtensorMap b, long long* out, int i,
                const __grid_constant__ BigParams p) {
  griddep_launch_dependents();
  griddep_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0)
    out[i] = (long long)globaltimer_ns() + (p.v[3] & 0) + (reinterpret_cast<const int*>(&a)[0] & 0) +
             (reinterpret_cast<const int*>(&b)[0] & 0);
}
// variants 5-8: what a conv kernel does around its body, piece by piece
//   5: tcgen05.alloc 512 columns + relinquish + dealloc       6: + mbarrier inits and fence
//   7: + one TMA box load (4 KB) completed on an mbarrier      8: + one UMMA (M128 N64 K16) committed to an mbarrier
__global__ void __launch_bounds__(480, 1)
dbg_launch_pieces(const __grid_constant__ CUtensorMap tm, long long* out, int i, int level, void* gstore) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tbase;
  __shared__ __align__(8) uint64_t bars[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[i] = (long long)globaltimer_ns();
  if (level >= 6 && threadIdx.x == 32) {
    for (int k = 0; k < 4; ++k) mbar_init(&bars[k], k == 2 ? 32 : 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tbase, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint8_t* buf = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
  if (level >= 7 && warp == 0 && lane == 0) {
    mbar_arrive_expect_tx(&bars[0], 4096);
    tma_load_3d(buf, &tm, &bars[0], 0, 0, 0);
    mbar_wait(&bars[0], 0);
  }
  __syncthreads();
  if (level >= 8 && warp == 1) {
    const uint32_t lead = elect_one();
    const uint64_t ad = umma_desc_kmajor(smem_u32(buf), 32);
    umma_f16_p(tbase, ad, ad, umma_idesc_f16(64, 0), 0u, lead);
    umma_commit_p(&bars[1], lead);
    if (lane == 0) mbar_wait(&bars[1], 0);
    tc_fence_after();
  }
  if (level >= 9) {  // epilogue-like global stores: 32 bytes per thread, 128-byte stride
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, i, 7);
    char* gp = reinterpret_cast<char*>(gstore) + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(gp), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
  }
  if (level >= 10 && warp >= 7) {  // 32 more TMA boxes (128 KB) per CTA, as a deep-K tile would pull
    if (lane == 0 && warp == 7) {
      for (int k = 0; k < 32; ++k) {
        mbar_arrive_expect_tx(&bars[2], 4096);
        tma_load_3d(buf + 4096 * (k % 16), &tm, &bars[2], 0, 0, k);
      }
      mbar_wait(&bars[2], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}
}  // namespace pb

extern "C" int pb_debug_launch_chain(long long* out, int variant, int n, void* stream) {
  using namespace pb;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  BigParams p{};
  CUtensorMap ta{}, tb{};
  const size_t smem = variant >= 1 ? 120 * 1024 : 0;
  PB_CUDA((cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(dbg_launch_small), smem));
  PB_CUDA((cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(dbg_launch_big), smem));
  PB_CUDA((cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(dbg_launch_tmap), smem));
  if (variant >= 5) {
    static CUtensorMap tm;
    static void* gbuf = nullptr;
    if (gbuf == nullptr) {
      PB_CUDA(cudaMalloc(&gbuf, 1 << 20));
      PB_CUDA(cudaMemset(gbuf, 0, 1 << 20));
      void* fp = nullptr;
      cudaDriverEntryPointQueryResult q;
      PB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
      auto encode = reinterpret_cast<EncodeTiledFn>(fp);
      cuuint64_t dims[3] = {16, 128, 64};
      cuuint64_t strides[2] = {32, 32 * 128};
      cuuint32_t box[3] = {16, 128, 1}, estr[3] = {1, 1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, gbuf, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      PB_CHECK(r == CUDA_SUCCESS, "debug: encode failed %d", (int)r);
    }
    PB_CUDA((cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(dbg_launch_pieces), 120 * 1024));
    static void* gstore = nullptr;
    if (gstore == nullptr) PB_CUDA(cudaMalloc(&gstore, (size_t)148 * 480 * 128));
    // variants 11 / 12: variant 10 on 37 / 74 CTAs (is the TMA pull bound per SM or by aggregate L2 bandwidth?)
    const int grid = variant == 11 ? 37 : (variant == 12 ? 74 : 148);
    const int level = variant > 10 ? 10 : variant;
    for (int i = 0; i < n; ++i) dbg_launch_pieces<<<grid, 480, 120 * 1024, s>>>(tm, out, i, level, gstore);
    PB_CUDA(cudaGetLastError());
    return 0;
  }
  for (int i = 0; i < n; ++i) {
    if (variant <= 1) {
      dbg_launch_small<<<148, 480, smem, s>>>(out, i);
    } else if (variant == 2) {
      dbg_launch_big<<<148, 480, smem, s>>>(out, i, p);
    } else if (variant == 3) {
      dbg_launch_tmap<<<148, 480, smem, s>>>(ta, tb, out, i, p);
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(480);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      PB_CUDA(cudaLaunchKernelEx(&cfg, dbg_launch_tmap, ta, tb, out, i, p));
    }
  }
  PB_CUDA(cudaGetLastError());
  return 0;
}
