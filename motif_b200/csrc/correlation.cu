// PWC-Net 9x9 cost volume for sm_100a.  Semantics: OpticalFlow/correlation.py:17-112, 294-348 (reference
// paths): out[b, 9*(dy+4)+(dx+4), y, x] = (1/C) sum_c first[b,c,y,x] * second[b,c,y+dy,x+dx], zeros outside.
//
// The reference spends one warp per output pixel, re-reads `second` from global memory for each of the 81
// displacements and first copies both inputs into padded NHWC buffers.  Here a CTA owns a 32x8 output
// tile: per chunk of 8 channels it stages the `first` tile and the (32+8)x(8+8) halo window of `second`
// straight from NCHW into shared memory (zero padding applied while staging, no rearranged copies), and
// each thread keeps a 4-pixel x 9-dx x 3-dy register tile (108 accumulators) fed by float4 shared loads,
// so every staged value is reused 81 times from registers/shared memory.
//
// The coarse PWC-Net levels are tiny images with many channels (196 x 12 x 20: two tiles), far too few CTAs for 148
// SMs.  There the channel range is split over a thread-block CLUSTER (up to 8 CTAs per output tile): every CTA
// accumulates its slice, the partial register tiles are parked in shared memory and the cluster's rank 0 adds them in
// rank order through distributed shared memory -- deterministic, no atomics, no workspace, output written once.
#include <cooperative_groups.h>

#include "common.cuh"

namespace motif {

constexpr int kTX = 32;   // output tile width
constexpr int kTY = 8;    // output tile height
constexpr int kCC = 8;    // channels staged per step (s1 8 KB + s2 20 KB of static shared memory)
constexpr int kWinW = kTX + 8;
constexpr int kWinH = kTY + 8;
constexpr int kCorrThreads = 8 * kTY * 3;  // 8 x-groups of 4 pixels, kTY rows, 3 dy-groups of 3

constexpr int kAccPerThread = 3 * 9 * 4;
constexpr int kCorrPartBytes = kAccPerThread * kCorrThreads * (int)sizeof(float);  // parked partial tile of one CTA

// grid.z = batch * ksplit; the ksplit CTAs of one output tile form a cluster (1, 1, ksplit)
__global__ void __launch_bounds__(kCorrThreads) corr_kernel(const float* __restrict__ first, const float* __restrict__ second,
                                                            float* __restrict__ out, int c, int h, int w, int ksplit) {
  __shared__ __align__(16) float s1[kCC][kTY][kTX];
  __shared__ __align__(16) float s2[kCC][kWinH][kWinW];
  extern __shared__ __align__(16) float part[];  // [kAccPerThread][kCorrThreads] when ksplit > 1

  const int b = blockIdx.z / ksplit, krank = blockIdx.z % ksplit;
  const int n_chunks = (c + kCC - 1) / kCC;
  const int chunk0 = (int)((long long)n_chunks * krank / ksplit), chunk1 = (int)((long long)n_chunks * (krank + 1) / ksplit);
  const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
  const int tid = threadIdx.x;
  const int xg = tid & 7;          // which 4-pixel group
  const int ty = (tid >> 3) % kTY;  // output row in the tile
  const int dg = tid / (8 * kTY);   // dy group: dy index 3*dg .. 3*dg+2
  const size_t hw = (size_t)h * w;
  const float* f1 = first + (size_t)b * c * hw;
  const float* f2 = second + (size_t)b * c * hw;

  float acc[3][9][4];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int d = 0; d < 9; ++d)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[r][d][p] = 0.0f;

  for (int c0 = chunk0 * kCC; c0 < chunk1 * kCC; c0 += kCC) {
    const int cc = min(kCC, c - c0);
    __syncthreads();
    for (int i = tid; i < kCC * kTY * kTX; i += kCorrThreads) {
      const int ch = i / (kTY * kTX), rem = i % (kTY * kTX);
      const int yy = y0 + rem / kTX, xx = x0 + rem % kTX;
      float v = 0.0f;
      if (ch < cc && yy < h && xx < w) v = __ldg(f1 + (size_t)(c0 + ch) * hw + (size_t)yy * w + xx);
      s1[ch][rem / kTX][rem % kTX] = v;
    }
    for (int i = tid; i < kCC * kWinH * kWinW; i += kCorrThreads) {
      const int ch = i / (kWinH * kWinW), rem = i % (kWinH * kWinW);
      const int yy = y0 - 4 + rem / kWinW, xx = x0 - 4 + rem % kWinW;
      float v = 0.0f;
      if (ch < cc && yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(f2 + (size_t)(c0 + ch) * hw + (size_t)yy * w + xx);
      s2[ch][rem / kWinW][rem % kWinW] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int ch = 0; ch < kCC; ++ch) {
      const float4 a = *reinterpret_cast<const float4*>(&s1[ch][ty][4 * xg]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* row = &s2[ch][ty + 3 * dg + r][4 * xg];
        const float4 w0 = *reinterpret_cast<const float4*>(row);
        const float4 w1 = *reinterpret_cast<const float4*>(row + 4);
        const float4 w2 = *reinterpret_cast<const float4*>(row + 8);
        const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int d = 0; d < 9; ++d)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[r][d][p] = fmaf(av[p], win[p + d], acc[r][d][p]);
      }
    }
  }

  if (ksplit > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (krank != 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int d = 0; d < 9; ++d)
#pragma unroll
          for (int p = 0; p < 4; ++p) part[((r * 9 + d) * 4 + p) * kCorrThreads + tid] = acc[r][d][p];
    }
    cluster.sync();
    if (krank == 0) {
      for (int k = 1; k < ksplit; ++k) {
        const float* remote = cluster.map_shared_rank(part, k);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int d = 0; d < 9; ++d)
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[r][d][p] += remote[((r * 9 + d) * 4 + p) * kCorrThreads + tid];
      }
    }
    cluster.sync();  // the parked tiles stay alive until rank 0 has read them
    if (krank != 0) return;
  }
  const int y = y0 + ty;
  if (y >= h) return;
  const float denom = (float)c;
  float* ob = out + (size_t)b * 81 * hw + (size_t)y * w;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      const int tc = 9 * (3 * dg + r) + d;
      float* o = ob + (size_t)tc * hw + x0 + 4 * xg;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (x0 + 4 * xg + p < w) o[p] = acc[r][d][p] / denom;  // correlation.py:108: total_sum / (float)sumelems
    }
}

}  // namespace motif

using namespace motif;

extern "C" int motif_corr_fwd(const float* first, const float* second, float* out, int b, int c, int h, int w, void* stream) {
  MOTIF_REQUIRE(first && second && out, "corr: null pointer");
  MOTIF_REQUIRE(b > 0 && c > 0 && h > 0 && w > 0, "corr: non-positive size b=%d c=%d h=%d w=%d", b, c, h, w);
  const int tiles = ceil_div(w, kTX) * ceil_div(h, kTY) * b;
  // split the channels over a cluster while the grid would leave most of the 148 SMs idle (2 CTAs fit per SM)
  int ksplit = 1;
  while (ksplit < 8 && tiles * ksplit * 2 <= 296 && ceil_div(c, kCC) >= ksplit * 4) ksplit *= 2;
  MOTIF_REQUIRE((long long)b * ksplit <= 65535, "corr: batch too large");
  static bool attr_done = false;
  if (!attr_done) {
    MOTIF_CUDA(cudaFuncSetAttribute(corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrPartBytes));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(w, kTX), ceil_div(h, kTY), b * ksplit);
  cfg.blockDim = dim3(kCorrThreads);
  cfg.dynamicSmemBytes = ksplit > 1 ? kCorrPartBytes : 0;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = ksplit;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    ProfScope prof("corr_kernel", (cudaStream_t)stream);
    MOTIF_CUDA(cudaLaunchKernelEx(&cfg, corr_kernel, first, second, out, c, h, w, ksplit));
    MOTIF_LAUNCHED("corr_kernel");
  }
  return 0;
}
