// Per-pixel temporal median of T uint8 frames: the TrackNet background of
// /root/reference/trackers/ball_tracker/iterable.py:58-81
//     median = np.median(np.array(frames_rgb), 0)   ->   median.astype("uint8")
// np.median of an even count averages the two middle order statistics in float64 and the uint8 cast truncates, i.e.
//     out = (s[(T-1)/2] + s[T/2]) >> 1        (both indices coincide for odd T).
//
// HBM-bound selection kernel, no sort and no per-pixel histogram memory: each thread owns four consecutive byte
// positions (one 32-bit word per frame, so a warp reads 128 contiguous bytes of every frame) and finds the order
// statistics by bisection on the VALUE: eight passes over its T words, each pass counting, with the per-byte SIMD
// compare __vcmpleu4, how many samples are <= the current mid-points.  The per-byte counters are packed 4 x 8 bit and
// spilled to 32-bit counters every 255 frames.  Both order statistics are searched in the same passes.
// Traffic: at most 8 reads of the frame stack (fewer when a block's T x 1 KB column stays in L2); 400 frames of
// 1080p = 2.5 GB -> a few milliseconds, against seconds for np.median on the host.
#include "internal.h"

namespace pb {

__device__ __forceinline__ uint32_t load_word(const uint8_t* p, long long pos, long long n) {
  if (pos + 4 <= n) return *reinterpret_cast<const uint32_t*>(p + pos);
  uint32_t w = 0;  // ragged tail: byte loads, missing bytes read as 0 (never stored)
  for (int b = 0; b < 4; ++b)
    if (pos + b < n) w |= (uint32_t)p[pos + b] << (8 * b);
  return w;
}

__global__ void __launch_bounds__(256) median_u8_kernel(const uint8_t* __restrict__ frames, int T, long long n,
                                                        uint8_t* __restrict__ out, int swap_rb) {
  const int k1 = (T - 1) / 2, k2 = T / 2;
  const long long nwords = (n + 3) / 4;
  for (long long wi = blockIdx.x * (long long)blockDim.x + threadIdx.x; wi < nwords;
       wi += (long long)gridDim.x * blockDim.x) {
    const long long pos = wi * 4;
    // per byte lane: smallest v with count(x <= v) >= k + 1 is s[k]; search range [lo, hi] packed 4 x 8 bit
    uint32_t lo1 = 0u, hi1 = 0xFFFFFFFFu, lo2 = 0u, hi2 = 0xFFFFFFFFu;
    for (int pass = 0; pass < 8; ++pass) {
      const uint32_t mid1 = __vhaddu4(lo1, hi1), mid2 = __vhaddu4(lo2, hi2);  // floor((lo + hi) / 2) per byte
      uint32_t c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
      for (int t0 = 0; t0 < T; t0 += 255) {
        const int t1 = t0 + 255 < T ? t0 + 255 : T;
        uint32_t p1 = 0, p2 = 0;
        for (int t = t0; t < t1; ++t) {
          const uint32_t x = load_word(frames + (long long)t * n, pos, n);
          p1 += __vcmpleu4(x, mid1) & 0x01010101u;
          p2 += __vcmpleu4(x, mid2) & 0x01010101u;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          c1[b] += (p1 >> (8 * b)) & 0xFFu;
          c2[b] += (p2 >> (8 * b)) & 0xFFu;
        }
      }
      uint32_t m1 = 0, m2 = 0;  // 0xFF in the lanes whose count reached k + 1 (answer <= mid)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        m1 |= (c1[b] >= (uint32_t)(k1 + 1) ? 0xFFu : 0u) << (8 * b);
        m2 |= (c2[b] >= (uint32_t)(k2 + 1) ? 0xFFu : 0u) << (8 * b);
      }
      // answer <= mid: hi = mid; else lo = mid + 1 (per-byte add: a lane at 255 must not carry into its neighbour)
      hi1 = (hi1 & ~m1) | (mid1 & m1);
      lo1 = (lo1 & m1) | (__vadd4(mid1, 0x01010101u) & ~m1);
      hi2 = (hi2 & ~m2) | (mid2 & m2);
      lo2 = (lo2 & m2) | (__vadd4(mid2, 0x01010101u) & ~m2);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long p = pos + b;
      if (p >= n) break;
      const uint32_t a = (lo1 >> (8 * b)) & 0xFFu, c = (lo2 >> (8 * b)) & 0xFFu;
      long long q = p;
      if (swap_rb) {  // BGR frames -> RGB median (iterable.py:63): channel c of a pixel goes to 2 - c
        const long long px = p / 3;
        q = px * 3 + (2 - (p - px * 3));
      }
      out[q] = (uint8_t)((a + c) >> 1);
    }
  }
}

}  // namespace pb

extern "C" int pb_median_u8(const uint8_t* frames, int T, long long frame_bytes, uint8_t* out, int swap_rb,
                            void* stream) {
  using namespace pb;
  PB_CHECK(frames && out, "median: null pointer");
  PB_CHECK(T >= 1 && frame_bytes >= 1, "median: T = %d, frame_bytes = %lld", T, frame_bytes);
  PB_CHECK((reinterpret_cast<uintptr_t>(frames) & 3) == 0 && (frame_bytes % 4 == 0 || T == 1),
           "median: frames must be 4-byte aligned with frame_bytes %% 4 == 0");
  PB_CHECK(!swap_rb || frame_bytes % 3 == 0, "median: swap_rb needs 3-channel pixels");
  const long long nwords = (frame_bytes + 3) / 4;
  long long blocks = (nwords + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  median_u8_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(frames, T, frame_bytes, out, swap_rb);
  PB_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
