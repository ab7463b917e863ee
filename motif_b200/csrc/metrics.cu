// Output path of the evaluation loop on the device (second half): crop of the decoded frames to the ground-truth
// size, L1 loss, BT.601 luma and per-frame MSE of test.py:187-235 in ONE pass over the two tensors (the reference runs ~12 eager
// elementwise / reduction kernels with full-size temporaries, then reads single numbers back).  HBM-bound: 24 bytes per pixel.
#include "common.cuh"

namespace motif {

// one CTA per (frame, slab of rows); thread <-> pixels of the slab, coalesced along x
__global__ void __launch_bounds__(256) frame_metrics_kernel(const float* __restrict__ fake, const float* __restrict__ real, double* __restrict__ out, int hp,
                                                            int wp, int h, int w, int rows_per_cta) {
  const int f = blockIdx.y;
  const float* fk = fake + (size_t)f * 3 * hp * wp;
  const float* rl = real + (size_t)f * 3 * h * w;
  const int y0 = blockIdx.x * rows_per_cta, y1 = min(y0 + rows_per_cta, h);
  float s_abs = 0.f, s_sq = 0.f;
  for (int i = threadIdx.x; i < (y1 - y0) * w; i += blockDim.x) {
    const int y = y0 + i / w, x = i - (i / w) * w;
    float yr, yf;
    {
      // test.py:212-217: *255, luma, /255 + 16, /255 -- the same fp32 operations in the same order
      const size_t pf = (size_t)y * wp + x, pr = (size_t)y * w + x;
      const float r0 = rl[pr], r1 = rl[pr + (size_t)h * w], r2 = rl[pr + 2 * (size_t)h * w];
      const float f0 = fk[pf], f1 = fk[pf + (size_t)hp * wp], f2 = fk[pf + 2 * (size_t)hp * wp];
      s_abs += fabsf(__fsub_rn(r0, f0)) + fabsf(__fsub_rn(r1, f1)) + fabsf(__fsub_rn(r2, f2));
      const float a0 = __fmul_rn(r0, 255.f), a1 = __fmul_rn(r1, 255.f), a2 = __fmul_rn(r2, 255.f);
      const float b0 = __fmul_rn(f0, 255.f), b1 = __fmul_rn(f1, 255.f), b2 = __fmul_rn(f2, 255.f);
      yr = __fadd_rn(__fadd_rn(__fmul_rn(a0, 65.481f), __fmul_rn(a1, 128.553f)), __fmul_rn(a2, 24.966f));
      yf = __fadd_rn(__fadd_rn(__fmul_rn(b0, 65.481f), __fmul_rn(b1, 128.553f)), __fmul_rn(b2, 24.966f));
      yr = __fdiv_rn(__fadd_rn(__fdiv_rn(yr, 255.f), 16.f), 255.f);
      yf = __fdiv_rn(__fadd_rn(__fdiv_rn(yf, 255.f), 16.f), 255.f);
    }
    const float dlt = __fsub_rn(yr, yf);
    s_sq = fmaf(dlt, dlt, s_sq);
  }
  __shared__ double red[2][8];
  double a = s_abs, b = s_sq;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = a, red[1][threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) a += red[0][i], b += red[1][i];
    atomicAdd(out + 2 * f, a);
    atomicAdd(out + 2 * f + 1, b);
  }
}

}  // namespace motif

using namespace motif;

extern "C" int motif_frame_metrics(const float* fake, const float* real, double* out, int n_frames, int hp, int wp, int h, int w, void* stream) {
  MOTIF_REQUIRE(fake && real && out, "frame_metrics: null pointer");
  MOTIF_REQUIRE(n_frames > 0 && h > 0 && w > 0 && hp >= h && wp >= w && n_frames <= 65535, "frame_metrics: bad size");
  cudaStream_t st = (cudaStream_t)stream;
  MOTIF_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 2 * n_frames, st));
  const int rows_per_cta = ceil_div(h, ceil_div(148 * 4, n_frames) < h ? ceil_div(148 * 4, n_frames) : h);
  frame_metrics_kernel<<<dim3(ceil_div(h, rows_per_cta), n_frames), 256, 0, st>>>(fake, real, out, hp, wp, h, w, rows_per_cta);
  MOTIF_LAUNCHED("frame_metrics_kernel");
  return 0;
}
