// InpaintNet (1-D conv U-Net that repairs gaps in the ball trajectory) as ONE kernel: one CTA per trajectory window,
// all activations in shared memory, fp32 math.  Network: /root/reference/trackers/ball_tracker/models.py:101-130
//   x = cat(coor (L,2), mask (L,1)) -> (3,L)
//   down_1 3->32, down_2 32->64, down_3 64->128, buttleneck 128->256->256,
//   up_1 cat(256,128)->128, up_2 cat(128,64)->64, up_3 cat(64,32)->32   (Conv1d k3 'same' + LeakyReLU(0.01))
//   predictor 32->2 (Conv1d k3 'same') + sigmoid -> (L,2)
// Called from ball_tracker.py:573-576 in the reference (which hard-codes .cuda() there).
#include "internal.h"

namespace pb {

constexpr int kInpMaxL = 32;

struct InpLayer {
  int cin, cout, w_off, b_off;  // offsets (floats) into the packed weight blob: w [cout][cin][3], b [cout]
};
struct InpParams {
  InpLayer layer[9];
  int L;
};

// out[co][l] = act(b[co] + sum_{ci,k} w[co][ci][k] * in[ci][l+k-1]); `in` may be the concatenation of two buffers
__device__ void inp_conv(const float* __restrict__ blob, const InpLayer ly, const float* in0, int c0, const float* in1,
                         float* out, int L, int act) {
  for (int co = threadIdx.x; co < ly.cout; co += blockDim.x) {
    float acc[kInpMaxL];
    const float bias = blob[ly.b_off + co];
#pragma unroll
    for (int l = 0; l < kInpMaxL; ++l) acc[l] = bias;
    const float* w = blob + ly.w_off + (size_t)co * ly.cin * 3;
    for (int ci = 0; ci < ly.cin; ++ci) {
      const float* src = ci < c0 ? in0 + ci * L : in1 + (ci - c0) * L;
      const float w0 = w[ci * 3], w1 = w[ci * 3 + 1], w2 = w[ci * 3 + 2];
#pragma unroll
      for (int l = 0; l < kInpMaxL; ++l) {
        if (l < L) {
          const float xm = l > 0 ? src[l - 1] : 0.f;
          const float xc = src[l];
          const float xp = l + 1 < L ? src[l + 1] : 0.f;
          acc[l] = fmaf(w2, xp, fmaf(w1, xc, fmaf(w0, xm, acc[l])));
        }
      }
    }
#pragma unroll
    for (int l = 0; l < kInpMaxL; ++l) {
      if (l < L) {
        float v = acc[l];
        if (act == 0) v = v > 0.f ? v : 0.01f * v;  // LeakyReLU default slope
        else v = 1.f / (1.f + expf(-v));
        out[co * L + l] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) inpaintnet_kernel(const float* __restrict__ coor, const float* __restrict__ mask,
                                                         const float* __restrict__ blob, InpParams p,
                                                         float* __restrict__ out) {
  extern __shared__ float sm[];
  const int L = p.L;
  float* x0 = sm;             // 3 x L
  float* x1 = x0 + 3 * L;     // 32
  float* x2 = x1 + 32 * L;    // 64
  float* x3 = x2 + 64 * L;    // 128
  float* ba = x3 + 128 * L;   // 256
  float* bb = ba + 256 * L;   // 256
  float* u1 = bb + 256 * L;   // 128
  float* u2 = u1 + 128 * L;   // 64
  float* u3 = u2 + 64 * L;    // 32
  float* y = u3 + 32 * L;     // 2
  const size_t n = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    x0[0 * L + i] = coor[(n * L + i) * 2 + 0];
    x0[1 * L + i] = coor[(n * L + i) * 2 + 1];
    x0[2 * L + i] = mask[n * L + i];
  }
  __syncthreads();
  inp_conv(blob, p.layer[0], x0, 3, nullptr, x1, L, 0);
  __syncthreads();
  inp_conv(blob, p.layer[1], x1, 32, nullptr, x2, L, 0);
  __syncthreads();
  inp_conv(blob, p.layer[2], x2, 64, nullptr, x3, L, 0);
  __syncthreads();
  inp_conv(blob, p.layer[3], x3, 128, nullptr, ba, L, 0);
  __syncthreads();
  inp_conv(blob, p.layer[4], ba, 256, nullptr, bb, L, 0);
  __syncthreads();
  inp_conv(blob, p.layer[5], bb, 256, x3, u1, L, 0);  // cat([x, x3])
  __syncthreads();
  inp_conv(blob, p.layer[6], u1, 128, x2, u2, L, 0);  // cat([x, x2])
  __syncthreads();
  inp_conv(blob, p.layer[7], u2, 64, x1, u3, L, 0);   // cat([x, x1])
  __syncthreads();
  inp_conv(blob, p.layer[8], u3, 32, nullptr, y, L, 1);
  __syncthreads();
  for (int i = threadIdx.x; i < L * 2; i += blockDim.x) out[n * L * 2 + i] = y[(i & 1) * L + (i >> 1)];
}

}  // namespace pb

using namespace pb;

extern "C" int pb_inpaintnet_forward(const float* coor, const float* mask, int N, int L, const float* weights,
                                     float* out, void* stream) {
  PB_CHECK(coor && mask && weights && out, "inpaintnet: null pointer");
  PB_CHECK(L >= 1 && L <= kInpMaxL, "inpaintnet: sequence length %d not in [1, %d]", L, kInpMaxL);
  if (N <= 0) return 0;
  static const int chans[9][2] = {{3, 32}, {32, 64}, {64, 128}, {128, 256}, {256, 256},
                                  {384, 128}, {192, 64}, {96, 32}, {32, 2}};
  InpParams p;
  p.L = L;
  int off = 0;
  for (int i = 0; i < 9; ++i) {
    p.layer[i].cin = chans[i][0];
    p.layer[i].cout = chans[i][1];
    p.layer[i].w_off = off;
    off += chans[i][0] * chans[i][1] * 3;
    p.layer[i].b_off = off;
    off += chans[i][1];
  }
  const size_t smem = (size_t)(3 + 32 + 64 + 128 + 256 + 256 + 128 + 64 + 32 + 2) * L * sizeof(float);
  PB_CUDA((cudaError_t)ensure_dynamic_smem(reinterpret_cast<const void*>(inpaintnet_kernel), smem));
  inpaintnet_kernel<<<N, 256, smem, static_cast<cudaStream_t>(stream)>>>(coor, mask, weights, p, out);
  PB_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
