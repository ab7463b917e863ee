// RAFT correlation lookup of one pyramid level.  The shipped model runs RAFT-small with
// alternate_corr=True (models/modules/Ours.py:417-430): AlternateCorrBlock.__call__ (models/core/corr.py:69-87) hands one
// level at a time to alt_cuda_corr.forward, a CUDA module the reference ships only as a binary.  The quantity is defined
// in-repo by CorrBlock (corr.py:8-56): the all-pairs correlation volume of the level sampled bilinearly
// (utils.py:57-70: grid_sample align_corners=True, zero padding) in a (2r+1)^2 window whose FIRST index moves x.
//
// One warp per query pixel.  The window's samples share one fractional offset, so the warp needs the (2r+2)^2 dot
// products with the integer neighbourhood of the query's target (lane <-> neighbourhood position, the query's feature row
// broadcast from shared memory) and then blends four of them per output.  Nothing of the H W x H2 W2 volume exists.
#include <stdlib.h>

#include "common.cuh"

namespace motif {

constexpr int kLookWarps = 8;
constexpr int kLookMaxC = 512;   // feature channels staged per query (RAFT: 128 small, 256 full)
constexpr int kLookMaxN = 10;    // 2 r + 2 for r <= 4

__global__ void __launch_bounds__(kLookWarps * 32) raft_corr_lookup_kernel(const float* __restrict__ fmap1, const float* __restrict__ fmap2,
                                                                           const float* __restrict__ coords, float* __restrict__ out,
                                                                           int B, int H, int W, int H2, int W2, int C, int r) {
  __shared__ __align__(16) float s_f1[kLookWarps][kLookMaxC];
  __shared__ float s_dot[kLookWarps][kLookMaxN * kLookMaxN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kLookWarps + warp;
  const long long hw = (long long)H * W;
  if (q >= (long long)B * hw) return;  // whole warps only; no block-wide barrier below
  const int b = (int)(q / hw);
  const int n = 2 * r + 2, nw = 2 * r + 1;
  const float* f1 = fmap1 + q * C;
  for (int k = lane; k < C; k += 32) s_f1[warp][k] = __ldg(f1 + k);
  float cx = __ldg(coords + 2 * q), cy = __ldg(coords + 2 * q + 1);
  const bool finite = isfinite(cx) && isfinite(cy);
  cx = finite ? fminf(fmaxf(cx, -1.0e6f), 1.0e6f) : -1.0e6f;  // far outside: every sample is zero padding
  cy = finite ? fminf(fmaxf(cy, -1.0e6f), 1.0e6f) : -1.0e6f;
  const float fx = floorf(cx), fy = floorf(cy);
  const float tx = cx - fx, ty = cy - fy;
  const int x0 = (int)fx - r, y0 = (int)fy - r;
  __syncwarp();
  const bool vec = (C & 3) == 0;
  for (int pos = lane; pos < n * n; pos += 32) {
    const int i = pos / n, j = pos - i * n;  // i moves x, j moves y
    const int xx = x0 + i, yy = y0 + j;
    float dot = 0.0f;
    if (xx >= 0 && xx < W2 && yy >= 0 && yy < H2) {
      const float* f2 = fmap2 + (((long long)b * H2 + yy) * W2 + xx) * C;
      if (vec) {
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        for (int k = 0; k < C; k += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(f2 + k));
          const float4 a = *reinterpret_cast<const float4*>(&s_f1[warp][k]);
          d0 = fmaf(a.x, v.x, d0), d1 = fmaf(a.y, v.y, d1), d2 = fmaf(a.z, v.z, d2), d3 = fmaf(a.w, v.w, d3);
        }
        dot = (d0 + d1) + (d2 + d3);
      } else {
        for (int k = 0; k < C; ++k) dot = fmaf(s_f1[warp][k], __ldg(f2 + k), dot);
      }
    }
    s_dot[warp][pos] = dot;
  }
  __syncwarp();
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
  const int y = (int)((q - (long long)b * hw) / W), x = (int)(q - (long long)b * hw - (long long)y * W);
  for (int o = lane; o < nw * nw; o += 32) {
    const int a = o / nw, c = o - a * nw;
    const float* d = &s_dot[warp][a * n + c];
    const float v = w00 * d[0] + w10 * d[n] + w01 * d[1] + w11 * d[n + 1];
    out[(((long long)b * nw * nw + o) * H + y) * W + x] = finite ? v : 0.0f;
  }
}

// Fast path for r <= 3 (RAFT-small: r = 3, 64 neighbourhood positions) and C a multiple of 128: lane <-> 4 channels
// (x C / 128), so the warp reads each 512-byte feature row of the neighbourhood with ONE coalesced LDG.128 instead of
// 32 strided ones (the generic kernel above is bound by its 32 wavefronts per load), keeps the 64 partial dot products in
// registers and finishes them with a transposing butterfly: 62 shuffles for all 64 sums instead of 5 per sum.
template <int CV>  // CV = C / 128
__global__ void __launch_bounds__(kLookWarps * 32, 3) raft_corr_lookup64_kernel(const float* __restrict__ fmap1, const float* __restrict__ fmap2,
                                                                             const float* __restrict__ coords, float* __restrict__ out,
                                                                             int B, int H, int W, int H2, int W2, int r) {
  __shared__ float s_dot[kLookWarps][64];
  constexpr int C = 128 * CV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q = (long long)blockIdx.x * kLookWarps + warp;
  const long long hw = (long long)H * W;
  if (q >= (long long)B * hw) return;
  const int b = (int)(q / hw);
  const int n = 2 * r + 2, nw = 2 * r + 1;
  float4 a[CV];
#pragma unroll
  for (int v = 0; v < CV; ++v) a[v] = __ldg(reinterpret_cast<const float4*>(fmap1 + q * C + 128 * v) + lane);
  float cx = __ldg(coords + 2 * q), cy = __ldg(coords + 2 * q + 1);
  const bool finite = isfinite(cx) && isfinite(cy);
  cx = finite ? fminf(fmaxf(cx, -1.0e6f), 1.0e6f) : -1.0e6f;
  cy = finite ? fminf(fmaxf(cy, -1.0e6f), 1.0e6f) : -1.0e6f;
  const float fx = floorf(cx), fy = floorf(cy);
  const float tx = cx - fx, ty = cy - fy;
  const int x0 = (int)fx - r, y0 = (int)fy - r;
  const float* f2b = fmap2 + (long long)b * H2 * W2 * C + 4 * lane;
  // Two passes of 32 neighbourhood positions each: 32 partial sums in registers instead of 64 lets three CTAs share an SM instead
  // of two (the kernel is latency-bound: every position is an independent 512-byte row load), and the transposing butterfly of a
  // pass (31 shuffles) leaves lane l with the finished sum of position 32 * pass + l.
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float part[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const int pos = 32 * pass + p;
      const int i = pos >> 3, j = pos & 7;  // slot (i, j) of an 8 x 8 grid; only i, j < n are used
      const int xx = x0 + i, yy = y0 + j;
      float d = 0.0f;
      if (i < n && j < n && xx >= 0 && xx < W2 && yy >= 0 && yy < H2) {
        const float* row = f2b + ((long long)yy * W2 + xx) * C;
#pragma unroll
        for (int v = 0; v < CV; ++v) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(row + 128 * v));
          d = fmaf(a[v].x, w.x, fmaf(a[v].y, w.y, fmaf(a[v].z, w.z, fmaf(a[v].w, w.w, d))));
        }
      }
      part[p] = d;
    }
    // transposing butterfly: after the step with offset o a lane keeps the half of its values selected by (lane & o)
#pragma unroll
    for (int half = 16, o = 16; half >= 1; half >>= 1, o >>= 1) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = up ? part[i] : part[i + half];
        const float keep = up ? part[i + half] : part[i];
        part[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    s_dot[warp][32 * pass + lane] = part[0];
  }
  __syncwarp();
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
  const int y = (int)((q - (long long)b * hw) / W), x = (int)(q - (long long)b * hw - (long long)y * W);
  for (int o = lane; o < nw * nw; o += 32) {
    const int ai = o / nw, c = o - ai * nw;
    const float* d = &s_dot[warp][ai * 8 + c];
    const float v = w00 * d[0] + w10 * d[8] + w01 * d[1] + w11 * d[9];
    out[(((long long)b * nw * nw + o) * H + y) * W + x] = finite ? v : 0.0f;
  }
}

}  // namespace motif

using namespace motif;

extern "C" int motif_raft_corr_lookup(const float* fmap1, const float* fmap2, const float* coords, float* out, int B, int H, int W, int H2,
                                      int W2, int C, int r, void* stream) {
  MOTIF_REQUIRE(fmap1 && fmap2 && coords && out, "raft_corr_lookup: null pointer");
  MOTIF_REQUIRE(B > 0 && H > 0 && W > 0 && H2 > 0 && W2 > 0, "raft_corr_lookup: non-positive size");
  MOTIF_REQUIRE(C > 0 && C <= kLookMaxC, "raft_corr_lookup: C=%d outside [1, %d]", C, kLookMaxC);
  MOTIF_REQUIRE(r >= 0 && 2 * r + 2 <= kLookMaxN, "raft_corr_lookup: radius %d outside [0, %d]", r, (kLookMaxN - 2) / 2);
  MOTIF_REQUIRE((C & 3) != 0 || ((((uintptr_t)fmap1 | (uintptr_t)fmap2) & 15) == 0), "raft_corr_lookup: feature maps must be 16-byte aligned");
  const long long queries = (long long)B * H * W;
  MOTIF_REQUIRE(queries < (1LL << 31), "raft_corr_lookup: too many queries");
  ProfScope prof("raft_corr_lookup_kernel", (cudaStream_t)stream);
  static const bool generic_only = getenv("MOTIF_RAFT_GENERIC") != nullptr;
  if (!generic_only && r <= 3 && (C == 128 || C == 256) && ((((uintptr_t)fmap1 | (uintptr_t)fmap2) & 15) == 0)) {
    const int grid = ceil_div(queries, kLookWarps);
    if (C == 128) raft_corr_lookup64_kernel<1><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, B, H, W, H2, W2, r);
    else raft_corr_lookup64_kernel<2><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, B, H, W, H2, W2, r);
    MOTIF_LAUNCHED("raft_corr_lookup_kernel");
    return 0;
  }
  raft_corr_lookup_kernel<<<ceil_div(queries, kLookWarps), kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, B, H, W, H2, W2, C, r);
  MOTIF_LAUNCHED("raft_corr_lookup_kernel");
  return 0;
}
