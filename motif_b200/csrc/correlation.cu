// PWC-Net 9x9 cost volume for sm_100a.  Semantics: OpticalFlow/correlation.py:17-112, 294-348 (reference
// paths): out[b, 9*(dy+4)+(dx+4), y, x] = (1/C) sum_c first[b,c,y,x] * second[b,c,y+dy,x+dx], zeros outside.
//
// The reference spends one warp per output pixel, re-reads `second` from global memory for each of the 81
// displacements and first copies both inputs into padded NHWC buffers.  Here a CTA owns a 32x4 output
// tile: per chunk of 8 channels it stages the `first` tile and the (32+8)x(4+8) halo window of `second`
// straight from NCHW into shared memory (zero padding applied while staging, no rearranged copies), and
// each thread keeps a 4-pixel x 9-dx register tile of one displacement row fed by float4 shared loads, so every staged
// value is reused 81 times from registers/shared memory; chunks arrive through cp.async, one in flight behind the one
// being multiplied.
//
// The coarse PWC-Net levels are tiny images with many channels (196 x 12 x 20: three tiles), far too few CTAs for 148
// SMs.  There the channel range is split over a thread-block CLUSTER (up to 8 CTAs per output tile): every CTA
// accumulates its slice, the partial register tiles are parked in shared memory and the cluster's rank 0 adds them in
// rank order through distributed shared memory -- deterministic, no atomics, no workspace, output written once.
#include <cooperative_groups.h>

#include "common.cuh"

namespace motif {

constexpr int kTX = 32;   // output tile width
constexpr int kTY = 4;    // output tile height
constexpr int kCC = 8;    // channels staged per step (4 KB of `first` + 15 KB of `second`, two such stages)
constexpr int kWinW = kTX + 8;
constexpr int kWinH = kTY + 8;
// thread = (4-pixel group xg, output row ty, displacement row dy): 4 x 9 accumulators.  288 threads and ~64 registers
// give three CTAs = 27 warps per SM; the first version kept 3 dy rows per thread (108 accumulators, 168 registers, 12
// warps per SM) and spent 8 cycles per instruction and warp waiting on its own dependent chains.
constexpr int kCorrThreads = 8 * kTY * 9;
constexpr int kAccPerThread = 9 * 4;
constexpr int kCorrPartBytes = kAccPerThread * kCorrThreads * (int)sizeof(float);  // parked partial tile of one CTA

constexpr int kS1 = kCC * kTY * kTX, kS2 = kCC * kWinH * kWinW;  // floats of one staged chunk: `first` tile, `second` window
constexpr int kStageFloats = kS1 + kS2;
constexpr int kCorrStageBytes = 2 * kStageFloats * (int)sizeof(float);  // two chunks: one being multiplied, one in flight
constexpr int kCorrSmemBytes = kCorrStageBytes > kCorrPartBytes ? kCorrStageBytes : kCorrPartBytes;  // the parked tile reuses the stages

__device__ __forceinline__ void corr_cp16(float* dst, const float* src, bool inside) {
  const int sz = inside ? 16 : 0;  // src-size 0: zero fill (the reference's zero padding, correlation.py:32-40)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void corr_cp4(float* dst, const float* src, bool inside) {
  const int sz = inside ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}

// What one thread copies of every staged chunk: one 4-pixel group of one plane position (the same for all kCC channels,
// so the address arithmetic is done once per kernel, not once per element).  Threads 0..159 own the 16 x 10 groups of the
// `second` window, threads 160..223 the 8 x 8 groups of the `first` tile.
struct CorrCopy {
  int dst;          // float offset inside a stage (channel 0)
  int plane;        // floats between channels inside the stage
  long long src;    // float offset inside one input channel plane
  unsigned inside;  // VEC: bit 0; scalar: bits 0..3, one per pixel
  int which;        // 0: nothing, 1: first, 2: second
};
__device__ __forceinline__ CorrCopy corr_copy_setup(int tid, int x0, int y0, int h, int w) {
  CorrCopy cp;
  cp.which = 0, cp.dst = 0, cp.plane = 0, cp.src = 0, cp.inside = 0;
  int yy, xx;
  if (tid < kWinH * (kWinW / 4)) {
    const int row = tid / (kWinW / 4), g = tid % (kWinW / 4);
    yy = y0 - 4 + row, xx = x0 - 4 + 4 * g;
    cp.which = 2, cp.dst = kS1 + row * kWinW + 4 * g, cp.plane = kWinH * kWinW;
  } else if (tid < kWinH * (kWinW / 4) + kTY * (kTX / 4)) {
    const int t = tid - kWinH * (kWinW / 4), row = t / (kTX / 4), g = t % (kTX / 4);
    yy = y0 + row, xx = x0 + 4 * g;
    cp.which = 1, cp.dst = row * kTX + 4 * g, cp.plane = kTY * kTX;
  } else {
    return cp;
  }
  const bool yin = yy >= 0 && yy < h;
#pragma unroll
  for (int p = 0; p < 4; ++p) cp.inside |= (yin && xx + p >= 0 && xx + p < w) ? (1u << p) : 0u;
  cp.src = (long long)yy * w + xx;
  return cp;
}
template <bool VEC>
__device__ __forceinline__ void corr_stage(float* __restrict__ st, const float* __restrict__ f1, const float* __restrict__ f2, int c0, int cc,
                                           size_t hw, const CorrCopy& cp) {
  if (cp.which != 0) {
    const float* base = cp.which == 1 ? f1 : f2;
    const float* src = base + (size_t)c0 * hw + cp.src;
    float* dst = st + cp.dst;
#pragma unroll
    for (int ch = 0; ch < kCC; ++ch) {
      const bool chin = ch < cc;
      if (VEC) {
        const bool in = chin && (cp.inside & 1u);
        corr_cp16(dst, in ? src : base, in);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const bool in = chin && ((cp.inside >> p) & 1u);
          corr_cp4(dst + p, in ? src + p : base, in);
        }
      }
      src += hw;
      dst += cp.plane;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// grid.z = batch * ksplit; the ksplit CTAs of one output tile form a cluster (1, 1, ksplit)
template <bool VEC>
__global__ void __launch_bounds__(kCorrThreads, 3) corr_kernel(const float* __restrict__ first, const float* __restrict__ second,
                                                               float* __restrict__ out, int c, int h, int w, int ksplit) {
  extern __shared__ __align__(16) float dyn[];  // [2][kS1 + kS2] staged chunks; afterwards [kAccPerThread][kCorrThreads] when ksplit > 1
  float* part = dyn;

  const int b = blockIdx.z / ksplit, krank = blockIdx.z % ksplit;
  const int n_chunks = (c + kCC - 1) / kCC;
  const int chunk0 = n_chunks * krank / ksplit, chunk1 = n_chunks * (krank + 1) / ksplit;
  const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
  const int tid = threadIdx.x;
  const int xg = tid & 7;          // which 4-pixel group
  const int ty = (tid >> 3) % kTY;  // output row in the tile
  const int dy = tid / (8 * kTY);   // displacement row 0..8 (dy - 4 of the reference)
  const size_t hw = (size_t)h * w;
  const float* f1 = first + (size_t)b * c * hw;
  const float* f2 = second + (size_t)b * c * hw;
  const CorrCopy cp = corr_copy_setup(tid, x0, y0, h, w);

  float acc[9][4];
#pragma unroll
  for (int d = 0; d < 9; ++d)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[d][p] = 0.0f;

  if (chunk0 < chunk1) corr_stage<VEC>(dyn, f1, f2, chunk0 * kCC, min(kCC, c - chunk0 * kCC), hw, cp);
  for (int q = chunk0; q < chunk1; ++q) {
    const float* st = dyn + ((q - chunk0) & 1) * kStageFloats;
    if (q + 1 < chunk1) {  // the other stage was last read in iteration q - 1, which every thread left through the barrier below
      corr_stage<VEC>(dyn + ((q + 1 - chunk0) & 1) * kStageFloats, f1, f2, (q + 1) * kCC, min(kCC, c - (q + 1) * kCC), hw, cp);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // chunk q of every thread has landed
    const float* a_p = st + ty * kTX + 4 * xg;
    const float* w_p = st + kS1 + (ty + dy) * kWinW + 4 * xg;
#pragma unroll
    for (int ch = 0; ch < kCC; ++ch) {
      const float4 a = *reinterpret_cast<const float4*>(a_p + ch * (kTY * kTX));
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float* row = w_p + ch * (kWinH * kWinW);
      const float4 w0 = *reinterpret_cast<const float4*>(row);
      const float4 w1 = *reinterpret_cast<const float4*>(row + 4);
      const float4 w2 = *reinterpret_cast<const float4*>(row + 8);
      const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
      for (int d = 0; d < 9; ++d)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[d][p] = fmaf(av[p], win[p + d], acc[d][p]);
    }
    __syncthreads();  // stage (q & 1) may be refilled by the next iteration's request
  }

  if (ksplit > 1) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (krank != 0) {
#pragma unroll
      for (int d = 0; d < 9; ++d)
#pragma unroll
        for (int p = 0; p < 4; ++p) part[(d * 4 + p) * kCorrThreads + tid] = acc[d][p];
    }
    cluster.sync();
    if (krank == 0) {
      // rank order per accumulator (deterministic); the loads of all ranks are independent and travel together
      const float* remote[8];
#pragma unroll
      for (int k = 1; k < 8; ++k) remote[k] = cluster.map_shared_rank(part, k < ksplit ? k : 0) + tid;
#pragma unroll
      for (int d = 0; d < 9; ++d) {
        float v[8][4];
#pragma unroll
        for (int k = 1; k < 8; ++k)
#pragma unroll
          for (int p = 0; p < 4; ++p) v[k][p] = k < ksplit ? remote[k][(d * 4 + p) * kCorrThreads] : 0.0f;
#pragma unroll
        for (int k = 1; k < 8; ++k)
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (k < ksplit) acc[d][p] += v[k][p];
      }
    }
    cluster.sync();  // the parked tiles stay alive until rank 0 has read them
    if (krank != 0) return;
  }
  const int y = y0 + ty;
  if (y >= h) return;
  const float denom = (float)c;
  float* ob = out + (size_t)b * 81 * hw + (size_t)y * w + x0 + 4 * xg;
  const bool vec_out = VEC && x0 + 4 * xg + 3 < w;  // VEC: w % 4 == 0 and 16-byte aligned planes (the launcher checks `out` too)
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    float* o = ob + (size_t)(9 * dy + d) * hw;
    // correlation.py:108: total_sum / (float)sumelems
    const float4 v = make_float4(acc[d][0] / denom, acc[d][1] / denom, acc[d][2] / denom, acc[d][3] / denom);
    if (vec_out) {
      *reinterpret_cast<float4*>(o) = v;
    } else {
      if (x0 + 4 * xg + 0 < w) o[0] = v.x;
      if (x0 + 4 * xg + 1 < w) o[1] = v.y;
      if (x0 + 4 * xg + 2 < w) o[2] = v.z;
      if (x0 + 4 * xg + 3 < w) o[3] = v.w;
    }
  }
}

}  // namespace motif

using namespace motif;

extern "C" int motif_corr_fwd(const float* first, const float* second, float* out, int b, int c, int h, int w, void* stream) {
  MOTIF_REQUIRE(first && second && out, "corr: null pointer");
  MOTIF_REQUIRE(b > 0 && c > 0 && h > 0 && w > 0, "corr: non-positive size b=%d c=%d h=%d w=%d", b, c, h, w);
  const int tiles = ceil_div(w, kTX) * ceil_div(h, kTY) * b;
  // split the channels over a cluster while the grid would leave most of the 148 SMs idle
  int ksplit = 1;
  while (ksplit < 8 && tiles * ksplit * 2 <= 148 * 3 && ceil_div(c, kCC) >= ksplit * 2 * 4) ksplit *= 2;  // three CTAs fit per SM; >= 4 chunks each
  MOTIF_REQUIRE((long long)b * ksplit <= 65535, "corr: batch too large");
  static bool attr_done_dev[64] = {false};
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    MOTIF_CUDA(cudaFuncSetAttribute(corr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmemBytes));
    MOTIF_CUDA(cudaFuncSetAttribute(corr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmemBytes));
    attr_done = true;
  }
  // 16-byte staging needs rows that start 16-byte aligned and groups of four pixels that are inside or outside as a whole
  const bool vec = (w % 4 == 0) && (((uintptr_t)first | (uintptr_t)second | (uintptr_t)out) & 15) == 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(w, kTX), ceil_div(h, kTY), b * ksplit);
  cfg.blockDim = dim3(kCorrThreads);
  cfg.dynamicSmemBytes = ksplit > 1 ? kCorrSmemBytes : kCorrStageBytes;
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
    if (vec) MOTIF_CUDA(cudaLaunchKernelEx(&cfg, corr_kernel<true>, first, second, out, c, h, w, ksplit));
    else MOTIF_CUDA(cudaLaunchKernelEx(&cfg, corr_kernel<false>, first, second, out, c, h, w, ksplit));
    MOTIF_LAUNCHED("corr_kernel");
  }
  return 0;
}
