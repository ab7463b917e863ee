// Modulated deformable convolution (DCNv2) forward for the model's configuration (64 -> 64 channels, 8 deformable groups,
// 3x3 / stride 1 / padding 1: every call site of Ours.py:53-172) as an implicit GEMM on tcgen05.
// Replaces modulated_deformable_im2col_gpu_kernel + SGEMM of the reference's extension (src/cuda/dcn_v2_im2col_cuda.cu:125-195,
// src/cuda/dcn_v2_cuda.cu:126-152).  The [576, H*W] column matrix never exists in global memory:
//
//   out[p][co] = sum_k col[p][k] * weight[co][k],  M = pixels (128 per tile), N = 64, K = 576
//
// One persistent CTA per SM.  The whole weight matrix stays resident in shared memory as nine K-major 64 x 64 fp16 hi/lo
// blocks (144 KB, 128-byte swizzle); sixteen producer warps sample the displaced taps straight into the A operand of the
// MMA (a two-stage ring of 128 x 64 hi/lo blocks), one warp issues tcgen05.mma kind::f16 (three products per block:
// hi*hi + lo*hi + hi*lo, fp32 accumulation in tensor memory, the arithmetic of the decoder), four warps drain the
// double-buffered accumulator into the NCHW output.
//
// K is permuted (the weight image is built with the same permutation, so the sum is the reference's with another
// association): k'' = (g * 9 + tap) * 8 + c for channel 8 g + c.  One (pixel, group, tap) work item then owns eight
// consecutive K values = one 16-byte chunk of the pixel's swizzled row: the bilinear setup (offsets, mask, corner indices
// and weights, dcn_v2_im2col_cuda.cu:25-55, 163-186) is computed once per item and a block of 64 K values is eight items per
// pixel in (group, tap) order, which keeps the nine taps of a group's channels close together in time (L1 reuse).
//
// fp16 range: weights and columns are scaled by powers of two (exact) from max|weight| and the bound max|in| * max|mask|
// of a small statistics pre-pass, so that the largest operand lands near 2^13 and the lo pieces stay normal.
#include <cuda_fp16.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "tc_common.cuh"

namespace motif {
namespace dcn_tc {

using namespace tc;

constexpr int kThreads = 768;   // warp 0: MMA issuer, warp 1: TMEM owner, warps 4-7: epilogue, warps 8-23: producers
constexpr int kEpiWarp0 = 4;
constexpr int kProdWarp0 = 8, kProdWarps = 16;
constexpr int kTilePx = 128;
constexpr int kBlocksK = 9;                // 576 / 64
constexpr int kAHalf = kTilePx * 128;      // one fp16 128 x 64 block (hi or lo)
constexpr int kABytes = 2 * kAHalf;
constexpr int kBHalf = 64 * 128;           // one fp16 64 x 64 block
constexpr int kBBytes = 2 * kBHalf;
constexpr int kStages = 2;

struct Smem {
  unsigned char b[kBlocksK][kBBytes];  // must stay first (1024-byte aligned swizzle atoms)
  unsigned char a[kStages][kABytes];
  float bias[64];
  uint64_t full[kStages], empty[kStages], d_full[2], d_empty[2];
  uint32_t tmem_base;
  int tiles_done;  // tiles of this CTA whose accumulator has been drained (paces the L2 prefetch warp)
};

// max |x| of the three operand tensors as uint bit patterns (non-negative floats order like their bits; a NaN wins)
__global__ void __launch_bounds__(256) dcn_stats_kernel(const float* __restrict__ in, size_t n_in, const float* __restrict__ mask, size_t n_mask,
                                                        const float* __restrict__ w, size_t n_w, unsigned int* __restrict__ stats) {
  __shared__ unsigned int red[3][8];
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int m[3] = {0u, 0u, 0u};
  const float* ptr[3] = {in, mask, w};
  const size_t cnt[3] = {n_in, n_mask, n_w};
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const float* p = ptr[t];
    const size_t n = cnt[t];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      for (size_t i = i0; i < n / 4; i += stride) {
        const float4 v = __ldg(p4 + i);
        m[t] = max(max(m[t], __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)), max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
      }
      for (size_t i = (n / 4) * 4 + i0; i < n; i += stride) m[t] = max(m[t], __float_as_uint(fabsf(__ldg(p + i))));
    } else {
      for (size_t i = i0; i < n; i += stride) m[t] = max(m[t], __float_as_uint(fabsf(__ldg(p + i))));
    }
    m[t] = __reduce_max_sync(0xffffffffu, m[t]);
    if ((threadIdx.x & 31) == 0) red[t][threadIdx.x >> 5] = m[t];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned int v = 0u;
    for (int k = 0; k < 8; ++k) v = max(v, red[threadIdx.x][k]);
    if (v != 0u) atomicMax(stats + threadIdx.x, v);
  }
}

// 2^(13 - floor(log2 x)) for the bit pattern of a positive finite x (1 for zero / non-finite; exponent clamped to +-60)
__device__ __forceinline__ float pow2_scale(unsigned int bits) {
  const int e = (int)((bits >> 23) & 0xffu);
  if (e == 0 || e == 255) return 1.0f;
  int s = 13 - (e - 127);
  s = s < -60 ? -60 : (s > 60 ? 60 : s);
  return __uint_as_float((unsigned int)(s + 127) << 23);
}

// fp32 pair -> fp16 hi pair + fp16 lo pair (lo = fp16 of the exact fp32 remainder)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(p));
  return p != 0;
}

__global__ void __launch_bounds__(kThreads, 1) dcn_v2_tc_kernel(const float* __restrict__ in, const float* __restrict__ offset, const float* __restrict__ mask,
                                                               const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out,
                                                               const unsigned int* __restrict__ stats, int B, int H, int W) {
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = H * W;
  const long long total = (long long)B * hw;
  const int n_tiles = (int)((total + kTilePx - 1) / kTilePx);
  const int n_iters = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const float s_w = pow2_scale(stats[2]);
  const float s_a = pow2_scale(__float_as_uint(__uint_as_float(stats[0]) * fmaxf(__uint_as_float(stats[1]), 1.0f)));

  // ---- weight image: b[blk][n = co][k'' % 64], hi then lo, 128-byte swizzle ----
  for (int e = threadIdx.x; e < 64 * 576; e += kThreads) {
    const int co = e / 576, kk = e - co * 576;
    const int ci = kk / 9, tap = kk - ci * 9;
    const int k2 = ((ci >> 3) * 9 + tap) * 8 + (ci & 7);
    const float v = __ldg(weight + e) * s_w;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    unsigned char* blk = sm.b[k2 >> 6];
    const uint32_t off = sw128_offset_h(co, k2 & 63);
    *reinterpret_cast<__half*>(blk + off) = hi;
    *reinterpret_cast<__half*>(blk + kBHalf + off) = lo;
  }
  if (threadIdx.x < 64) sm.bias[threadIdx.x] = bias != nullptr ? bias[threadIdx.x] : 0.0f;
  if (threadIdx.x == 0) {
    sm.tiles_done = 0;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], kProdWarps);
      mbar_init(&sm.empty[s], 1);
    }
    for (int d = 0; d < 2; ++d) {
      mbar_init(&sm.d_full[d], 1);
      mbar_init(&sm.d_empty[d], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(&sm.tmem_base);
  fence_proxy_async_smem();  // the weight image is read by the async proxy (tcgen05.mma)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ---- MMA issuer: the whole warp walks the loop, one elected lane issues ----
    constexpr uint32_t idesc = idesc_f16(128, 64);
    const uint64_t b_desc0 = smem_desc_sw128(smem_u32(&sm.b[0][0]));
    const uint64_t a_desc0 = smem_desc_sw128(smem_u32(&sm.a[0][0]));
    uint32_t cnt = 0;
    for (int it = 0; it < n_iters; ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&sm.d_empty[buf], ((it >> 1) & 1) ^ 1);
      const uint32_t dcol = tmem_base + 64 * buf;
#pragma unroll 1
      for (int blk = 0; blk < kBlocksK; ++blk, ++cnt) {
        const uint32_t stage = cnt & 1;
        mbar_wait(&sm.full[stage], (cnt >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ahi = a_desc0 + (uint64_t)(stage * (kABytes >> 4)), alo = ahi + (kAHalf >> 4);
          const uint64_t bhi = b_desc0 + (uint64_t)(blk * (kBBytes >> 4)), blo = bhi + (kBHalf >> 4);
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint64_t a = (term == 1) ? alo : ahi;
            const uint64_t b = (term == 2) ? blo : bhi;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_f16_ss(dcol, a + 2 * ks, b + 2 * ks, idesc, (blk | term | ks) != 0);
          }
          mma_commit(&sm.empty[stage]);
          if (blk == kBlocksK - 1) mma_commit(&sm.d_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    // ---- epilogue: thread <-> pixel (TMEM lane), 64 output channels ----
    const int quad = warp & 3;
    const float inv = 1.0f / (s_a * s_w);
    for (int it = 0; it < n_iters; ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(&sm.d_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld64(tmem_base + ((uint32_t)(quad * 32) << 16) + 64 * buf, r);
      tc_fence_before();
      mbar_arrive(&sm.d_empty[buf]);
      if (quad == 0 && lane == 0) *(volatile int*)&sm.tiles_done = it + 1;
      const long long p = (long long)(blockIdx.x + it * gridDim.x) * kTilePx + quad * 32 + lane;
      if (p < total) {
        const int b = (int)(p / hw), s = (int)(p - (long long)b * hw);
        float* o = out + (size_t)b * 64 * hw + s;
#pragma unroll
        for (int co = 0; co < 64; ++co) o[(size_t)co * hw] = fmaf(__uint_as_float(r[co]), inv, sm.bias[co]);
      }
    }
  } else if (warp >= kProdWarp0) {
    // ---- producers: lane <-> pixel (32 consecutive pixels of the tile), the warp's two 16-byte chunks of every block ----
    // The offsets and the mask of a work item are read one block ahead (three registers per item): they come from HBM
    // (50 MB per call, read once) and a dependent round trip per item was the critical path of the first version.
    const int pw = warp - kProdWarp0;
    const int row = (pw & 3) * 32 + lane;
    const int c0 = pw >> 2;  // the warp's chunks of a block: c0 and c0 + 4
    const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    uint32_t cnt = 0;
    struct Pre {
      float off_h, off_w, m;
    };
    auto fetch = [&](int b, int s, int blk, int rep, bool live) {
      Pre r{0.f, 0.f, 0.f};
      if (live) {
        const int j = 8 * blk + c0 + 4 * rep, g = j / 9, tap = j - 9 * g;
        const float* op = offset + ((size_t)(b * 8 + g) * 18 + 2 * tap) * hw + s;
        r.off_h = __ldg(op), r.off_w = __ldg(op + hw);
        r.m = __ldg(mask + ((size_t)(b * 8 + g) * 9 + tap) * hw + s);
      }
      return r;
    };
    Pre nxt[2];
    {
      const long long p = (long long)blockIdx.x * kTilePx + row;
      const bool live = p < total;
      const int b = live ? (int)(p / hw) : 0, s = live ? (int)(p - (long long)b * hw) : 0;
      nxt[0] = fetch(b, s, 0, 0, live), nxt[1] = fetch(b, s, 0, 1, live);
    }
    for (int it = 0; it < n_iters; ++it) {
      const long long p = (long long)(blockIdx.x + it * gridDim.x) * kTilePx + row;
      const bool live = p < total;
      const int b = live ? (int)(p / hw) : 0, s = live ? (int)(p - (long long)b * hw) : 0;
      const int y = s / W, x = s - y * W;
      const float* in_b = in + (size_t)b * 64 * hw;
      // the pixel of this thread in the CTA's next tile (for the prefetch of its first block)
      const long long pn = p + (long long)gridDim.x * kTilePx;
      const bool live_n = (it + 1 < n_iters) && pn < total;
      const int bn = live_n ? (int)(pn / hw) : 0, sn = live_n ? (int)(pn - (long long)bn * hw) : 0;
#pragma unroll 1
      for (int blk = 0; blk < kBlocksK; ++blk, ++cnt) {
        const Pre cur[2] = {nxt[0], nxt[1]};
        if (blk + 1 < kBlocksK) {
          nxt[0] = fetch(b, s, blk + 1, 0, live), nxt[1] = fetch(b, s, blk + 1, 1, live);
        } else {
          nxt[0] = fetch(bn, sn, 0, 0, live_n), nxt[1] = fetch(bn, sn, 0, 1, live_n);
        }
        const uint32_t stage = cnt & 1;
        mbar_wait(&sm.empty[stage], ((cnt >> 1) & 1) ^ 1);
        unsigned char* a_hi = sm.a[stage] + row_off;
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int c = c0 + 4 * rep;  // 16-byte chunk of the row
          const int j = 8 * blk + c, g = j / 9, tap = j - 9 * g;
          // bilinear setup (dcn_v2_im2col_cuda.cu:163-186, 25-55); corners that contribute nothing read element 0 with weight 0
          int id[4] = {0, 0, 0, 0};
          float cw[4] = {0.f, 0.f, 0.f, 0.f};
          if (live) {
            const float m = cur[rep].m * s_a;
            const float h_im = (float)(y - 1 + tap / 3) + cur[rep].off_h, w_im = (float)(x - 1 + tap % 3) + cur[rep].off_w;
            if (h_im > -1.0f && w_im > -1.0f && h_im < (float)H && w_im < (float)W) {
              const float hf = floorf(h_im), wf = floorf(w_im);
              const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
              const float lh = h_im - hf, lw = w_im - wf, hh = 1.0f - lh, hw_ = 1.0f - lw;
              if (h_low >= 0 && w_low >= 0) id[0] = h_low * W + w_low, cw[0] = hh * hw_ * m;
              if (h_low >= 0 && w_high <= W - 1) id[1] = h_low * W + w_high, cw[1] = hh * lw * m;
              if (h_high <= H - 1 && w_low >= 0) id[2] = h_high * W + w_low, cw[2] = lh * hw_ * m;
              if (h_high <= H - 1 && w_high <= W - 1) id[3] = h_high * W + w_high, cw[3] = lh * lw * m;
            }
          }
          const float* plane = in_b + (size_t)(8 * g) * hw;
          float v[8];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const float* pl = plane + (size_t)ch * hw;
            const float v0 = __ldg(pl + id[0]), v1 = __ldg(pl + id[1]), v2 = __ldg(pl + id[2]), v3 = __ldg(pl + id[3]);
            v[ch] = fmaf(v3, cw[3], fmaf(v2, cw[2], fmaf(v1, cw[1], v0 * cw[0])));
          }
          uint4 hi, lo;
          split_pair(v[0], v[1], hi.x, lo.x);
          split_pair(v[2], v[3], hi.y, lo.y);
          split_pair(v[4], v[5], hi.z, lo.z);
          split_pair(v[6], v[7], hi.w, lo.w);
          const uint32_t off = (uint32_t)((c ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          *reinterpret_cast<uint4*>(a_hi + kAHalf + off) = lo;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.full[stage]);
      }
    }
  } else if (warp == 2) {
    // ---- L2 prefetch of the next tile's offsets and mask (216 channel rows of 512 B, read once from HBM) ----
    for (int it = 1; it < n_iters; ++it) {
      // one tile ahead of the producers is enough (a monotonic counter, polled: no phase to miss)
      while (*(volatile int*)&sm.tiles_done < it - 2) __nanosleep(500);
      const long long p0 = (long long)(blockIdx.x + it * gridDim.x) * kTilePx;
      for (int i = lane; i < 216 * 4; i += 32) {
        const int r = i >> 2;
        const long long p = p0 + (i & 3) * 32;
        if (p >= total) continue;
        const int b = (int)(p / hw), s = (int)(p - (long long)b * hw);
        const float* a = r < 144 ? offset + ((size_t)b * 144 + r) * hw + s : mask + ((size_t)b * 72 + (r - 144)) * hw + s;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// statistics slots: a ring, so that calls in flight on several streams of one device do not share a slot
__device__ unsigned int g_dcn_stats[64][4];
static std::atomic<unsigned int> g_stats_next{0};

}  // namespace dcn_tc

bool dcn_v2_tc_applicable(int Cin, int Cout, int dg) {
  static const bool off = getenv("MOTIF_DCN_SIMT") != nullptr && atoi(getenv("MOTIF_DCN_SIMT")) != 0;
  return !off && Cin == 64 && Cout == 64 && dg == 8;
}

int dcn_v2_tc_fwd(const float* in, const float* offset, const float* mask, const float* weight, const float* bias, float* out, int B, int H, int W,
                  cudaStream_t st) {
  using namespace dcn_tc;
  const int smem = (int)sizeof(Smem) + 1024;
  static bool attr_done_dev[64] = {false};
  static int n_sm_dev[64];
  const int slot = current_device_slot();
  if (!attr_done_dev[slot]) {
    MOTIF_CUDA(cudaFuncSetAttribute(dcn_v2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int dev = 0;
    MOTIF_CUDA(cudaGetDevice(&dev));
    MOTIF_CUDA(cudaDeviceGetAttribute(&n_sm_dev[slot], cudaDevAttrMultiProcessorCount, dev));
    attr_done_dev[slot] = true;
  }
  unsigned int* stats_base = nullptr;
  MOTIF_CUDA(cudaGetSymbolAddress((void**)&stats_base, g_dcn_stats));
  unsigned int* stats = stats_base + 4 * (g_stats_next.fetch_add(1, std::memory_order_relaxed) & 63u);
  MOTIF_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned int), st));
  const size_t hw = (size_t)H * W;
  const int n_sm = n_sm_dev[slot];
  dcn_stats_kernel<<<n_sm * 4, 256, 0, st>>>(in, (size_t)B * 64 * hw, mask, (size_t)B * 72 * hw, weight, (size_t)64 * 576, stats);
  MOTIF_LAUNCHED("dcn_stats_kernel");
  const int n_tiles = ceil_div((long long)B * hw, kTilePx);
  ProfScope prof("dcn_v2_tc_kernel", st);
  dcn_v2_tc_kernel<<<n_tiles < n_sm ? n_tiles : n_sm, kThreads, smem, st>>>(in, offset, mask, weight, bias, out, stats, B, H, W);
  MOTIF_LAUNCHED("dcn_v2_tc_kernel");
  return 0;
}

}  // namespace motif
