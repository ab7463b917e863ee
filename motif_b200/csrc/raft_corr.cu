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
// blockIdx.y = pyramid level: the whole AlternateCorrBlock.__call__ (corr.py:69-87) in one launch -- level l samples fmap2 pooled
// l times at coords / 2^l and writes channels [l (2r+1)^2, (l+1) (2r+1)^2) of the stacked output; with `inv_norm` the result is
// divided by sqrt(C) as corr.py:87 does.  The one-level entry point passes a single level, scale 1 and no normalisation.
struct LookupLevels {
  const float* fmap2[4];
  int h2[4], w2[4];
  int n_levels;
  float norm;  // 0: none; else 1 / sqrt(C) (torch divides a CUDA tensor by a CPU scalar tensor as a * (1 / b), BinaryDivTrueKernel.cu)
};
template <int CV>  // CV = C / 128
__global__ void __launch_bounds__(kLookWarps * 32, 3) raft_corr_lookup64_kernel(const float* __restrict__ fmap1, LookupLevels lv,
                                                                             const float* __restrict__ coords, float* __restrict__ out,
                                                                             int B, int H, int W, int r) {
  __shared__ float s_dot[kLookWarps][64];
  constexpr int C = 128 * CV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int level = blockIdx.y;
  const float* __restrict__ fmap2 = level == 0 ? lv.fmap2[0] : level == 1 ? lv.fmap2[1] : level == 2 ? lv.fmap2[2] : lv.fmap2[3];
  const int H2 = level == 0 ? lv.h2[0] : level == 1 ? lv.h2[1] : level == 2 ? lv.h2[2] : lv.h2[3];
  const int W2 = level == 0 ? lv.w2[0] : level == 1 ? lv.w2[1] : level == 2 ? lv.w2[2] : lv.w2[3];
  const float cscale = 1.0f / (float)(1 << level);  // coords / 2**level (corr.py:80): exact
  const long long q = (long long)blockIdx.x * kLookWarps + warp;
  const long long hw = (long long)H * W;
  if (q >= (long long)B * hw) return;
  const int b = (int)(q / hw);
  const int n = 2 * r + 2, nw = 2 * r + 1;
  float4 a[CV];
#pragma unroll
  for (int v = 0; v < CV; ++v) a[v] = __ldg(reinterpret_cast<const float4*>(fmap1 + q * C + 128 * v) + lane);
  float cx = __ldg(coords + 2 * q) * cscale, cy = __ldg(coords + 2 * q + 1) * cscale;
  const bool finite = isfinite(cx) && isfinite(cy);
  cx = finite ? fminf(fmaxf(cx, -1.0e6f), 1.0e6f) : -1.0e6f;
  cy = finite ? fminf(fmaxf(cy, -1.0e6f), 1.0e6f) : -1.0e6f;
  const float fx = floorf(cx), fy = floorf(cy);
  const float tx = cx - fx, ty = cy - fy;
  const int x0 = (int)fx - r, y0 = (int)fy - r;
  const float* f2b = fmap2 + (long long)b * H2 * W2 * C + 4 * lane;
  // The kernel issued ~1800 instructions per query, most of them per-position bounds tests and address arithmetic: validity of the
  // 8 x 8 slots is now two bit masks, the address one row offset (per slot row j) plus one column offset (per slot column i), both
  // clamped into the map so that every load is legal, and the sum of a slot outside the map is zeroed by one select afterwards.
  unsigned vx = 0u, vy = 0u;
  int coff[8], roff[8];  // element offsets inside this batch item's map (H2 * W2 * C < 2^31, checked by the launcher)
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int xx = x0 + t, yy = y0 + t;
    vx |= (t < n && xx >= 0 && xx < W2) ? (1u << t) : 0u;
    vy |= (t < n && yy >= 0 && yy < H2) ? (1u << t) : 0u;
    coff[t] = min(max(xx, 0), W2 - 1) * C;
    roff[t] = min(max(yy, 0), H2 - 1) * W2 * C;
  }
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    float part[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      // slot (i, j) = (4 * pass + (p >> 3), p & 7) of the 8 x 8 grid (position 32 * pass + p)
      const float* row = f2b + (roff[p & 7] + (pass == 0 ? coff[p >> 3] : coff[4 + (p >> 3)]));
      float d = 0.0f;
#pragma unroll
      for (int v = 0; v < CV; ++v) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + 128 * v));
        d = fmaf(a[v].x, w.x, fmaf(a[v].y, w.y, fmaf(a[v].z, w.z, fmaf(a[v].w, w.w, d))));
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
    {  // lane l holds position 32 * pass + l = slot (4 * pass + (l >> 3), l & 7)
      const bool ok = ((vx >> (4 * pass + (lane >> 3))) & (vy >> (lane & 7)) & 1u) != 0u;
      s_dot[warp][32 * pass + lane] = ok ? part[0] : 0.0f;
    }
  }
  __syncwarp();
  const float w00 = (1.0f - tx) * (1.0f - ty), w10 = tx * (1.0f - ty), w01 = (1.0f - tx) * ty, w11 = tx * ty;
  const int y = (int)((q - (long long)b * hw) / W), x = (int)(q - (long long)b * hw - (long long)y * W);
  for (int o = lane; o < nw * nw; o += 32) {
    const int ai = o / nw, c = o - ai * nw;
    const float* d = &s_dot[warp][ai * 8 + c];
    float v = w00 * d[0] + w10 * d[8] + w01 * d[1] + w11 * d[9];
    if (lv.norm != 0.0f) v = __fmul_rn(v, lv.norm);
    out[(((long long)b * lv.n_levels * nw * nw + (long long)level * nw * nw + o) * H + y) * W + x] = finite ? v : 0.0f;
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
  MOTIF_REQUIRE((long long)H2 * W2 * C < (1LL << 31), "raft_corr_lookup: feature map too large");
  ProfScope prof("raft_corr_lookup_kernel", (cudaStream_t)stream);
  static const bool generic_only = getenv("MOTIF_RAFT_GENERIC") != nullptr;
  if (!generic_only && r <= 3 && (C == 128 || C == 256) && ((((uintptr_t)fmap1 | (uintptr_t)fmap2) & 15) == 0)) {
    const int grid = ceil_div(queries, kLookWarps);
    LookupLevels lv{};
    lv.fmap2[0] = fmap2, lv.h2[0] = H2, lv.w2[0] = W2, lv.n_levels = 1, lv.norm = 0.0f;
    if (C == 128) raft_corr_lookup64_kernel<1><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, lv, coords, out, B, H, W, r);
    else raft_corr_lookup64_kernel<2><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, lv, coords, out, B, H, W, r);
    MOTIF_LAUNCHED("raft_corr_lookup_kernel");
    return 0;
  }
  raft_corr_lookup_kernel<<<ceil_div(queries, kLookWarps), kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, fmap2, coords, out, B, H, W, H2, W2, C, r);
  MOTIF_LAUNCHED("raft_corr_lookup_kernel");
  return 0;
}

// All pyramid levels of AlternateCorrBlock.__call__ (models/core/corr.py:69-87) in ONE launch: fmap2_levels[l] is fmap2 pooled l times
// ([B, h2[l], w2[l], C] channels-last), coords [B, H, W, 2] are the level-0 coordinates (the kernel divides by 2^l), out is the stacked
// [B, n_levels * (2r+1)^2, H, W] tensor, already divided by sqrt(C) when `normalize` is set (corr.py:87).  r <= 3, C = 128 or 256.
extern "C" int motif_raft_corr_lookup_pyramid(const float* fmap1, const float* const* fmap2_levels, const int* h2, const int* w2, int n_levels,
                                              const float* coords, float* out, int B, int H, int W, int C, int r, int normalize, void* stream) {
  MOTIF_REQUIRE(fmap1 && fmap2_levels && h2 && w2 && coords && out, "raft_corr_lookup_pyramid: null pointer");
  MOTIF_REQUIRE(n_levels >= 1 && n_levels <= 4, "raft_corr_lookup_pyramid: %d levels outside [1, 4]", n_levels);
  MOTIF_REQUIRE(B > 0 && H > 0 && W > 0, "raft_corr_lookup_pyramid: non-positive size");
  MOTIF_REQUIRE(r >= 0 && r <= 3 && (C == 128 || C == 256), "raft_corr_lookup_pyramid: r=%d, C=%d (supported: r <= 3, C = 128 / 256)", r, C);
  const long long queries = (long long)B * H * W;
  MOTIF_REQUIRE(queries < (1LL << 31), "raft_corr_lookup_pyramid: too many queries");
  LookupLevels lv{};
  uintptr_t align = (uintptr_t)fmap1;
  for (int l = 0; l < n_levels; ++l) {
    MOTIF_REQUIRE(fmap2_levels[l] != nullptr && h2[l] > 0 && w2[l] > 0, "raft_corr_lookup_pyramid: level %d is empty", l);
    MOTIF_REQUIRE((long long)h2[l] * w2[l] * C < (1LL << 31), "raft_corr_lookup_pyramid: feature map too large");
    lv.fmap2[l] = fmap2_levels[l], lv.h2[l] = h2[l], lv.w2[l] = w2[l];
    align |= (uintptr_t)fmap2_levels[l];
  }
  MOTIF_REQUIRE((align & 15) == 0, "raft_corr_lookup_pyramid: feature maps must be 16-byte aligned");
  lv.n_levels = n_levels;
  lv.norm = normalize ? 1.0f / sqrtf((float)C) : 0.0f;
  ProfScope prof("raft_corr_lookup_kernel", (cudaStream_t)stream);
  dim3 grid(ceil_div(queries, kLookWarps), n_levels);
  if (C == 128) raft_corr_lookup64_kernel<1><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, lv, coords, out, B, H, W, r);
  else raft_corr_lookup64_kernel<2><<<grid, kLookWarps * 32, 0, (cudaStream_t)stream>>>(fmap1, lv, coords, out, B, H, W, r);
  MOTIF_LAUNCHED("raft_corr_lookup_kernel");
  return 0;
}
