// Forward splatting for sm_100a: FunctionSoftsplat (summation/average/linear/softmax) plus the max and
// count reliability variants.  Semantics: models/softsplat_cp.py:12-52, 320-347;
// softsplat_max_cp.py:12-58, 254; softsplat_count_cp.py:14-52, 163-165 (reference paths).
//
// Two algorithms for the sum splat:
//  * destination-centric ("bin then gather"): a first pass bins, per destination pixel, the (source,
//    weight) pairs whose 2x2 footprint covers it (int atomics only, one pass over the flow, amortised
//    over all C channels); a second pass has one thread per DESTINATION pixel accumulate its <= K
//    contributions for every channel with coalesced reads of the NCHW planes and one coalesced store.
//    No float atomics, output written exactly once, deterministic: contributions are added in source
//    raster order, which is the order a single thread executing the reference kernel adds them.
//    Destinations with more than K contributions get the surplus through the atomic kernel below.
//  * reference-order scatter with red.global.add.f32: the reference algorithm with the flow/weights
//    hoisted out of the channel loop; used for the overflow and as an on-device cross-check.
#include <stdlib.h>

#include "common.cuh"

namespace motif {

constexpr int kBinSlots = 8;  // contributions kept per destination pixel before overflowing

struct SplatWorkspace {
  int* count;      // [n*hw]   contributions per destination pixel
  int* ent_src;    // [K][n*hw] source pixel (y*w+x) of slot k
  float* ent_w;    // [K][n*hw] bilinear weight of slot k
  unsigned char* ovf_mask;  // [n*hw] per SOURCE pixel: bit c set = corner c did not get a slot
  int* ovf_total;  // [1]   number of source pixels with a non-zero mask
  int* ovf_list;   // [n*hw] those source pixels (global index b*hw + s), in no particular order
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static size_t workspace_layout(int n, int h, int w, SplatWorkspace* ws, char* base) {
  const size_t p = (size_t)n * h * w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align256(bytes);
    return ptr;
  };
  // count | ovf_mask | ovf_total first and back to back: one memset arms a call
  char* a = take(p * sizeof(int));
  char* d = take(p);
  char* e = take(256);
  char* b = take(p * sizeof(int) * kBinSlots);
  char* c = take(p * sizeof(float) * kBinSlots);
  char* f = take(p * sizeof(int));
  if (ws) {
    ws->ovf_list = (int*)f;
    ws->count = (int*)a;
    ws->ent_src = (int*)b;
    ws->ent_w = (float*)c;
    ws->ovf_mask = (unsigned char*)d;
    ws->ovf_total = (int*)e;
  }
  return off;
}

// more than 1/64 of the source pixels overflowed a destination list
__host__ __device__ inline bool surplus_is_dense(int n_list, long long n_src) { return (long long)n_list * 64 > n_src; }

template <int MODE>
__device__ __forceinline__ float metric_scale(const float* metric, size_t idx) {
  if (MODE == MOTIF_SPLAT_LINEAR) return metric[idx];
  if (MODE == MOTIF_SPLAT_SOFTMAX) return expf(metric[idx]);
  return 1.0f;
}

// ------------------------------------------------------------------------------------------------
// Reference-order scatter.  One thread per source pixel, channel loop inside.
// only_overflow: contribute only the corners flagged in ovf_mask (surplus of the binning pass).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void splat_scatter_dense(const float* __restrict__ in, const float* __restrict__ flow,
                                                    const float* __restrict__ metric, float* __restrict__ out,
                                                    int n, int c, int h, int w, const unsigned char* __restrict__ ovf_mask) {
  const int hw = h * w;
  const int c_out = (MODE == MOTIF_SPLAT_SUMMATION) ? c : c + 1;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)n * hw;
       p += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(p / hw), s = (int)(p % hw);
    unsigned mask = 0xF;
    if (ovf_mask != nullptr) {
      mask = ovf_mask[p];
      if (mask == 0) continue;
    }
    const int y = s / w, x = s % w;
    const Footprint f = footprint(x, y, flow[((size_t)b * 2 + 0) * hw + s], flow[((size_t)b * 2 + 1) * hw + s]);
    if (!f.finite) continue;
    int dst[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cx, cy;
      ok[k] = corner_inside(f, k, w, h, cx, cy) && ((mask >> k) & 1);
      dst[k] = cy * w + cx;
    }
    if (!(ok[0] | ok[1] | ok[2] | ok[3])) continue;
    const float m = metric_scale<MODE>(metric, p);
    const float* src = in + (size_t)b * c * hw + s;
    float* o = out + (size_t)b * c_out * hw;
#pragma unroll 8
    for (int ch = 0; ch < c; ++ch) {
      float v = src[(size_t)ch * hw];
      if (MODE >= MOTIF_SPLAT_LINEAR) v = __fmul_rn(v, m);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) red_add_f32(o + (size_t)ch * hw + dst[k], __fmul_rn(v, f.w[k]));
    }
    if (MODE != MOTIF_SPLAT_SUMMATION) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) red_add_f32(o + (size_t)c * hw + dst[k], __fmul_rn(m, f.w[k]));
    }
  }
}

// the reference algorithm as an operator of its own (motif_splat_fwd_atomic)
template <int MODE>
__global__ void __launch_bounds__(256) splat_scatter_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                                            const float* __restrict__ metric, float* __restrict__ out,
                                                            int n, int c, int h, int w) {
  splat_scatter_dense<MODE>(in, flow, metric, out, n, c, h, w, nullptr);
}

// Surplus of the binning pass: one WARP per listed source pixel, lanes over channels (the list is sparse, so a thread
// per source would leave most of a warp idle and serialise 130 channels of atomics behind one thread).
template <int MODE>
__global__ void __launch_bounds__(256) splat_scatter_list_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                                                 const float* __restrict__ metric, float* __restrict__ out,
                                                                 int n, int c, int h, int w, SplatWorkspace ws) {
  const int hw = h * w;
  const int c_out = (MODE == MOTIF_SPLAT_SUMMATION) ? c : c + 1;
  const int n_list = *ws.ovf_total;
  if (n_list == 0) return;
  if (surplus_is_dense(n_list, (long long)n * hw)) {  // many overflowed sources: thread per source (neighbours coalesce), same launch
    splat_scatter_dense<MODE>(in, flow, metric, out, n, c, h, w, ws.ovf_mask);
    return;
  }
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_list; i += warps) {
    const int p = ws.ovf_list[i];
    const unsigned mask = ws.ovf_mask[p];
    const int b = p / hw, s = p - b * hw;
    const int y = s / w, x = s - y * w;
    const Footprint f = footprint(x, y, flow[((size_t)b * 2 + 0) * hw + s], flow[((size_t)b * 2 + 1) * hw + s]);
    if (!f.finite) continue;
    int dst[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cx, cy;
      ok[k] = corner_inside(f, k, w, h, cx, cy) && ((mask >> k) & 1);
      dst[k] = cy * w + cx;
    }
    const float m = metric_scale<MODE>(metric, p);
    const float* src = in + (size_t)b * c * hw + s;
    float* o = out + (size_t)b * c_out * hw;
    for (int ch = lane; ch < c_out; ch += 32) {
      float v = m;  // the normaliser channel (ch == c) splats the metric scale itself
      if (ch < c) {
        v = src[(size_t)ch * hw];
        if (MODE >= MOTIF_SPLAT_LINEAR) v = __fmul_rn(v, m);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) red_add_f32(o + (size_t)ch * hw + dst[k], __fmul_rn(v, f.w[k]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Destination-centric pass 1: bin (source, weight) per destination pixel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) splat_bin_kernel(const float* __restrict__ flow, SplatWorkspace ws, int n, int h, int w) {
  const int hw = h * w;
  const size_t total = (size_t)n * hw;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)total;
       p += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(p / hw), s = (int)(p % hw);
    const int y = s / w, x = s % w;
    const Footprint f = footprint(x, y, flow[((size_t)b * 2 + 0) * hw + s], flow[((size_t)b * 2 + 1) * hw + s]);
    if (!f.finite) continue;
    unsigned overflow = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cx, cy;
      if (!corner_inside(f, k, w, h, cx, cy)) continue;
      const size_t d = (size_t)b * hw + (size_t)cy * w + cx;
      const int slot = atomicAdd(ws.count + d, 1);
      if (slot < kBinSlots) {
        ws.ent_src[(size_t)slot * total + d] = s;
        ws.ent_w[(size_t)slot * total + d] = f.w[k];
      } else {
        overflow |= 1u << k;
      }
    }
    if (overflow) {
      ws.ovf_mask[p] = (unsigned char)overflow;
      ws.ovf_list[atomicAdd(ws.ovf_total, 1)] = (int)p;
    }
  }
}

__device__ __forceinline__ void cswap(int& sa, float& wa, int& sb, float& wb) {
  const bool sw = sa > sb;
  const int ts = sw ? sb : sa;
  const float tw = sw ? wb : wa;
  sb = sw ? sa : sb;
  wb = sw ? wa : wb;
  sa = ts;
  wa = tw;
}

// ------------------------------------------------------------------------------------------------
// Destination-centric pass 2: one thread per destination pixel, all channels.
// The channel loop is specialised on the largest contribution count of the warp (K = 4, 6 or 8 slots), so a
// warp whose destinations all have <= 4 contributions (the common case for smooth flows) executes 4 gathers
// and 4 multiply-adds per pixel-channel and nothing else.
// ------------------------------------------------------------------------------------------------
template <int MODE, int K>
__device__ __forceinline__ void gather_channels(const float* __restrict__ plane, float* __restrict__ optr, int c, size_t hw, int cnt,
                                                const unsigned (&src)[kBinSlots], const float (&wt)[kBinSlots], const float (&m)[kBinSlots],
                                                bool store) {
#pragma unroll 4
  for (int ch = 0; ch < c; ++ch) {
    float v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __ldg(plane + src[k]);  // dead slots read element 0 (weight 0, not accumulated)
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float t = v[k];
      if (MODE >= MOTIF_SPLAT_LINEAR) t = __fmul_rn(t, m[k]);
      t = __fmul_rn(t, wt[k]);
      if (k < cnt) acc = __fadd_rn(acc, t);
    }
    if (store) *optr = acc;
    plane += hw;
    optr += hw;
  }
}

// Slots of one destination pixel: sorted by source index (Batcher odd-even merge network for 8 keys) so that the
// accumulation order is source raster order whatever order the binning atomics handed the slots out in.
template <int MODE>
__device__ __forceinline__ int load_slots(const SplatWorkspace& ws, const float* __restrict__ metric, size_t total, size_t gd, int b, int hw,
                                          bool live, unsigned (&src)[kBinSlots], float (&wt)[kBinSlots], float (&m)[kBinSlots]) {
  // every slot is read in the same round trip as the count (slots past the count hold stale values: masked below)
  const int cnt_raw = ws.count[gd];
  int srci[kBinSlots];
#pragma unroll
  for (int k = 0; k < kBinSlots; ++k) {
    srci[k] = ws.ent_src[(size_t)k * total + gd];
    wt[k] = ws.ent_w[(size_t)k * total + gd];
  }
  const int cnt = live ? min(cnt_raw, kBinSlots) : 0;
#pragma unroll
  for (int k = 0; k < kBinSlots; ++k) {
    const bool on = k < cnt;
    srci[k] = on ? srci[k] : 0x7fffffff;
    wt[k] = on ? wt[k] : 0.0f;
  }
#define CS(a, b) cswap(srci[a], wt[a], srci[b], wt[b])
  CS(0, 1); CS(2, 3); CS(4, 5); CS(6, 7);
  CS(0, 2); CS(1, 3); CS(4, 6); CS(5, 7);
  CS(1, 2); CS(5, 6);
  CS(0, 4); CS(1, 5); CS(2, 6); CS(3, 7);
  CS(2, 4); CS(3, 5);
  CS(1, 2); CS(3, 4); CS(5, 6);
#undef CS
#pragma unroll
  for (int k = 0; k < kBinSlots; ++k) {
    src[k] = (k < cnt) ? (unsigned)srci[k] : 0u;
    m[k] = 1.0f;
    if (MODE >= MOTIF_SPLAT_LINEAR && k < cnt) m[k] = metric_scale<MODE>(metric, (size_t)b * hw + src[k]);
  }
  return cnt;
}

template <int MODE>
__global__ void __launch_bounds__(256) splat_gather_kernel(const float* __restrict__ in, const float* __restrict__ metric,
                                                           float* __restrict__ out, SplatWorkspace ws, int n, int c, int h, int w) {
  const int hw = h * w;
  const size_t total = (size_t)n * hw;
  const int c_out = (MODE == MOTIF_SPLAT_SUMMATION) ? c : c + 1;
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = d < hw;
  const size_t gd = (size_t)b * hw + (live ? d : 0);
  unsigned src[kBinSlots];
  float wt[kBinSlots], m[kBinSlots];
  const int cnt = load_slots<MODE>(ws, metric, total, gd, b, hw, live, src, wt, m);
  const int wmax = __reduce_max_sync(0xffffffffu, cnt);
  const float* plane = in + (size_t)b * c * hw;
  float* optr = out + (size_t)b * c_out * hw + (live ? d : 0);
  if (wmax <= 4) gather_channels<MODE, 4>(plane, optr, c, hw, cnt, src, wt, m, live);
  else if (wmax <= 6) gather_channels<MODE, 6>(plane, optr, c, hw, cnt, src, wt, m, live);
  else gather_channels<MODE, 8>(plane, optr, c, hw, cnt, src, wt, m, live);
  if (MODE != MOTIF_SPLAT_SUMMATION && live) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < kBinSlots; ++k)
      if (k < cnt) acc = __fadd_rn(acc, __fmul_rn(m[k], wt[k]));
    optr[(size_t)c * hw] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Destination-centric pass 2, tiled: one CTA = a 32 x 8 tile of destination pixels, one thread per destination.
// The sources a tile gathers from lie in a small window of every channel plane (the tile moved by the local flow,
// plus its spread); the window of kCH channels at a time is staged into shared memory with 16-byte cp.async copies
// (each input element crosses L2 -> SM once per tile instead of once per destination that uses it -- four on
// average -- and as full sectors instead of 4-byte gathers), double-buffered against the accumulation, which
// then gathers with LDS and runs two channels per instruction (FMUL2 / FADD2; each lane rounds like the scalar
// op, order unchanged: still bit-exact against the reference kernel in raster order).  Tiles whose window does
// not fit (strongly diverging flow) and planes whose rows are not 16-byte aligned take the direct path above.
// ------------------------------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 8;    // destination tile
// Staging buffer: kBufFloats floats of dynamic shared memory per CTA (two CTAs per SM), cut into stages of kCH channel
// planes.  A plane holds the tile's actual source window (rw x rh floats, rw a multiple of 4, at most 256 16-byte
// columns so that every thread copies one) plus a zero word -- NOT a worst-case 64 x 16 rectangle: a smooth flow needs
// ~40 x 11, so the same memory holds seven chunks instead of three and five to six of them are in flight.  (With three
// stages a chunk took one memory round trip, ~1.5 us: two chunks in flight cannot cover the latency.)
#ifndef MOTIF_SPLAT_BUF_KB
#define MOTIF_SPLAT_BUF_KB 98
#endif
constexpr int kBufFloats = MOTIF_SPLAT_BUF_KB * 256;
constexpr int kMaxCols16 = 256;     // 16-byte columns of a window: one per thread
constexpr int kTiledSmem = kBufFloats * (int)sizeof(float);
constexpr int kPlaneLarge = kMaxCols16 * 4 + 4, kDepthLarge = kBufFloats / (8 * kPlaneLarge);  // any window: 3 stages in 98 KB
constexpr int kPlaneSmall = 512 + 4, kDepthSmall = kBufFloats / (8 * kPlaneSmall);              // windows <= 512 floats: 6 stages
static_assert(kDepthLarge >= 3 && kDepthSmall >= 3, "at least three stages (one barrier per chunk)");
constexpr int kCH = 8;              // channels per stage

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// channels [0, nch) of one staged chunk, K slots per destination, two channels per instruction.  Slots past the
// list length point at the zero word that follows every staged plane and carry weight 0: they add +0 (exact) without a
// predicate, and can never pick up a non-finite input the way a real window element could.
template <int MODE, int K, bool FULL, int PLANE>
__device__ __forceinline__ void tile_channels(const float* __restrict__ stage, float* __restrict__ optr, int nch, size_t hw,
                                              const int (&off)[kBinSlots], const float (&wt)[kBinSlots], const float (&m)[kBinSlots], bool store) {
#pragma unroll
  for (int cp = 0; cp < kCH / 2; ++cp) {
    if (!FULL && 2 * cp >= nch) break;
    const float* p0 = stage + (2 * cp) * PLANE;
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      f32x2 t = pack2(p0[off[k]], p0[off[k] + PLANE]);
      if (MODE >= MOTIF_SPLAT_LINEAR) t = fmul2(t, pack2(m[k], m[k]));
      t = fmul2(t, pack2(wt[k], wt[k]));
      // scalar adds: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding instead of the reference's two)
      float t0, t1;
      unpack2(t, t0, t1);
      a0 = __fadd_rn(a0, t0);
      a1 = __fadd_rn(a1, t1);
    }
    if (store) {
      optr[0] = a0;
      if (FULL || 2 * cp + 1 < nch) optr[hw] = a1;
    }
    optr += 2 * hw;
  }
}

// The chunk pipeline of one tile with D stages: chunk q + D - 1 is requested when chunk q is consumed; one barrier per
// chunk (it publishes chunk q and retires chunk q - 1, whose stage is the one refilled).
template <int MODE, int D, int PLANE>
__device__ __forceinline__ void tile_pipeline(float* __restrict__ buf, int c, size_t hw, const float* __restrict__ csrc, int cdst, bool copier,
                                              float* __restrict__ optr, int wmax, const int (&off)[kBinSlots], const float (&wt)[kBinSlots],
                                              const float (&m)[kBinSlots], bool live) {
  const int n_chunks = (c + kCH - 1) / kCH;
  constexpr int chunk = kCH * PLANE;
  static_assert(D * chunk <= kBufFloats, "stages do not fit the staging buffer");
  auto stage_in = [&](int q) {
    if (copier) {
      const int c0 = q * kCH, nch = min(kCH, c - c0);
      float* dst = buf + (q % D) * chunk + cdst;
      const float* sp = csrc + (size_t)c0 * hw;
      for (int ch = 0; ch < nch; ++ch) cp_async16(dst + ch * PLANE, sp + (size_t)ch * hw);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int q = 0; q < D - 1; ++q) {
    if (q < n_chunks) stage_in(q);
    else cp_async_commit();
  }
  for (int q = 0; q < n_chunks; ++q) {
    cp_async_wait<D - 2>();  // chunk q has landed (this thread's copies; the barrier publishes everyone's)
    __syncthreads();         // ... and every thread is done with chunk q - 1, whose stage is refilled next
    if (q + D - 1 < n_chunks) stage_in(q + D - 1);
    else cp_async_commit();
    const int c0 = q * kCH, nch = min(kCH, c - c0);
    float* o = optr + (size_t)c0 * hw;
    const float* st = buf + (q % D) * chunk;
    if (nch == kCH) {
      if (wmax <= 4) tile_channels<MODE, 4, true, PLANE>(st, o, nch, hw, off, wt, m, live);
      else if (wmax <= 6) tile_channels<MODE, 6, true, PLANE>(st, o, nch, hw, off, wt, m, live);
      else tile_channels<MODE, 8, true, PLANE>(st, o, nch, hw, off, wt, m, live);
    } else {
      tile_channels<MODE, 8, false, PLANE>(st, o, nch, hw, off, wt, m, live);
    }
  }
}

#ifndef MOTIF_SPLAT_CTAS
#define MOTIF_SPLAT_CTAS 2
#endif
template <int MODE>
__global__ void __launch_bounds__(256, MOTIF_SPLAT_CTAS) splat_gather_tiled_kernel(const float* __restrict__ in, const float* __restrict__ metric,
                                                                    float* __restrict__ out, SplatWorkspace ws, int n, int c, int h, int w) {
  extern __shared__ __align__(16) float buf[];  // [stages][kCH][plane], plane = window + zero word
  __shared__ int s_box[4];  // min x, min y, max x, max y of the sources this tile gathers from
  const int hw = h * w;
  const size_t total = (size_t)n * hw;
  const int c_out = (MODE == MOTIF_SPLAT_SUMMATION) ? c : c + 1;
  const int b = blockIdx.y;
  const int tiles_x = (w + kTW - 1) / kTW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = ((int)blockIdx.x % tiles_x) * kTW + tx, y = ((int)blockIdx.x / tiles_x) * kTH + ty;
  const bool live = (x < w) & (y < h);
  const int d = live ? y * w + x : 0;
  const size_t gd = (size_t)b * hw + d;
  if (threadIdx.x == 0) s_box[0] = s_box[1] = 0x7fffffff, s_box[2] = s_box[3] = -1;
  // (an L2 guess prefetch of the metric plane and the first channel planes around the tile, issued here before the slot lists are
  //  read, does not shorten the three-round-trip prologue measurably: gather 0.285 ms either way)
  unsigned src[kBinSlots];
  float wt[kBinSlots], m[kBinSlots];
  const int cnt = load_slots<MODE>(ws, metric, total, gd, b, hw, live, src, wt, m);
  int sx[kBinSlots], sy[kBinSlots];
  int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -1, mxy = -1;
#pragma unroll
  for (int k = 0; k < kBinSlots; ++k) {
    sy[k] = (int)(src[k] / (unsigned)w);
    sx[k] = (int)(src[k] - (unsigned)sy[k] * (unsigned)w);
    if (k < cnt) mnx = min(mnx, sx[k]), mny = min(mny, sy[k]), mxx = max(mxx, sx[k]), mxy = max(mxy, sy[k]);
  }
  mnx = __reduce_min_sync(0xffffffffu, mnx), mny = __reduce_min_sync(0xffffffffu, mny);
  mxx = __reduce_max_sync(0xffffffffu, mxx), mxy = __reduce_max_sync(0xffffffffu, mxy);
  const int wmax = __reduce_max_sync(0xffffffffu, cnt);
  __syncthreads();
  if (tx == 0 && mxx >= 0) atomicMin(&s_box[0], mnx), atomicMin(&s_box[1], mny), atomicMax(&s_box[2], mxx), atomicMax(&s_box[3], mxy);
  __syncthreads();
  const int x0a = s_box[0] & ~3, y0 = s_box[1];
  const int rw4 = (((s_box[2] + 4) & ~3) - x0a) >> 2, rh = s_box[3] - y0 + 1;  // window: rw4 16-byte columns x rh rows
  const int rw = 4 * rw4;
  const float* plane_in = in + (size_t)b * c * hw;
  float* optr = out + (size_t)b * c_out * hw + d;
  const bool empty = s_box[2] < 0;
  const int n16 = empty ? 0 : rw4 * rh;  // 16-byte columns of the window: one per thread
  if (n16 > kMaxCols16) {  // window too large for the stage: direct gathers (block-uniform branch)
    if (wmax <= 4) gather_channels<MODE, 4>(plane_in, optr, c, hw, cnt, src, wt, m, live);
    else if (wmax <= 6) gather_channels<MODE, 6>(plane_in, optr, c, hw, cnt, src, wt, m, live);
    else gather_channels<MODE, 8>(plane_in, optr, c, hw, cnt, src, wt, m, live);
  } else {
    // plane capacity (compile-time, so that the channel planes sit at immediate offsets): small windows get the deep pipeline
    const bool small = n16 * 4 <= kPlaneSmall - 4;
    const int zero_at = small ? kPlaneSmall - 4 : kPlaneLarge - 4;
    int off[kBinSlots];
#pragma unroll
    for (int k = 0; k < kBinSlots; ++k) {
      const bool on = k < cnt;
      off[k] = on ? (sy[k] - y0) * rw + (sx[k] - x0a) : zero_at;  // dead slots read the zero word behind the window
      wt[k] = on ? wt[k] : 0.0f;
      m[k] = on ? m[k] : 1.0f;
    }
    // the zero word of every plane the pipeline will use (cp.async never writes it; published by the first barrier)
    if ((int)threadIdx.x < (small ? kDepthSmall : kDepthLarge) * kCH) buf[threadIdx.x * (small ? kPlaneSmall : kPlaneLarge) + zero_at] = 0.0f;
    // this thread's 16-byte column of the window (same for every channel)
    const bool copier = (int)threadIdx.x < n16;
    const int crow = copier ? (int)threadIdx.x / rw4 : 0, ccol = copier ? (int)threadIdx.x - crow * rw4 : 0;
    const float* csrc = plane_in + (size_t)(y0 + crow) * w + x0a + 4 * ccol;
    const int cdst = crow * rw + 4 * ccol;
    if (small) tile_pipeline<MODE, kDepthSmall, kPlaneSmall>(buf, c, hw, csrc, cdst, copier, optr, wmax, off, wt, m, live);
    else tile_pipeline<MODE, kDepthLarge, kPlaneLarge>(buf, c, hw, csrc, cdst, copier, optr, wmax, off, wt, m, live);
  }
  if (MODE != MOTIF_SPLAT_SUMMATION && live) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < kBinSlots; ++k)
      if (k < cnt) acc = __fadd_rn(acc, __fmul_rn(m[k], wt[k]));
    optr[(size_t)c * hw] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Max and count variants (single-channel in the shipped pipeline): reference-order atomics.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, float v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void __launch_bounds__(256) splat_max_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                                                        float* __restrict__ out, int n, int c, int h, int w) {
  const int hw = h * w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)n * hw;
       p += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(p / hw), s = (int)(p % hw);
    const Footprint f = footprint(s % w, s / w, flow[((size_t)b * 2 + 0) * hw + s], flow[((size_t)b * 2 + 1) * hw + s]);
    if (!f.finite) continue;
    for (int ch = 0; ch < c; ++ch) {
      const float v = in[((size_t)b * c + ch) * hw + s];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int cx, cy;
        if (!corner_inside(f, k, w, h, cx, cy)) continue;
        const float cand = __fmul_rn(v, f.w[k]);
        // atomicMaxFloat (softsplat_max_cp.py:13-18) on a cell that starts at 1.0: only a candidate
        // that is >= 0 can win (int compare); negative or NaN candidates never change the cell.
        if (cand >= 0.0f) red_max_nonneg(out + ((size_t)b * c + ch) * hw + (size_t)cy * w + cx, cand);
      }
    }
  }
}

__global__ void __launch_bounds__(256) splat_count_kernel(const float* __restrict__ flow, float* __restrict__ out, int n, int h, int w) {
  const int hw = h * w;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)n * hw;
       p += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(p / hw), s = (int)(p % hw);
    // the reference count kernel has no isfinite assert; a non-finite position is out of the image anyway
    const Footprint f = footprint(s % w, s / w, flow[((size_t)b * 2 + 0) * hw + s], flow[((size_t)b * 2 + 1) * hw + s]);
    if (!f.finite) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int cx, cy;
      if (corner_inside(f, k, w, h, cx, cy)) red_add_f32(out + (size_t)b * hw + (size_t)cy * w + cx, 1.0f);
    }
  }
}

static int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * 16;  // 148 SMs x resident CTAs; kernels are grid-stride
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <int MODE>
static int launch_scatter(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h, int w, cudaStream_t st) {
  ProfScope prof("splat_scatter_kernel", st);
  splat_scatter_kernel<MODE><<<grid_for((long long)n * h * w, 256), 256, 0, st>>>(in, flow, metric, out, n, c, h, w);
  MOTIF_LAUNCHED("splat_scatter_kernel");
  return 0;
}

// Surplus of the binning pass (destinations with more than kBinSlots contributions): ONE launch, which returns at once
// when nothing overflowed, takes a warp per listed source when few did and a thread per source when many did.
template <int MODE>
static int launch_surplus(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h, int w, const SplatWorkspace& ws,
                          cudaStream_t st) {
  ProfScope prof("splat_scatter_kernel", st);
  splat_scatter_list_kernel<MODE><<<148 * 8, 256, 0, st>>>(in, flow, metric, out, n, c, h, w, ws);
  MOTIF_LAUNCHED("splat_scatter_kernel");
  return 0;
}

template <int MODE>
static int launch_gather(const float* in, const float* metric, float* out, const SplatWorkspace& ws, int n, int c, int h, int w,
                         cudaStream_t st) {
  // the tiled kernel stages 16-byte columns: every plane row must start 16-byte aligned
  static const bool force_direct = getenv("MOTIF_SPLAT_DIRECT") != nullptr;
  const bool tiled = !force_direct && (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  ProfScope prof("splat_gather_kernel", st);
  if (tiled) {
    dim3 grid(ceil_div(w, kTW) * ceil_div(h, kTH), n);
    static bool attr_done_dev[64] = {false};  // per MODE instantiation
  bool& attr_done = attr_done_dev[current_device_slot()];
    if (!attr_done) {
      MOTIF_CUDA(cudaFuncSetAttribute(splat_gather_tiled_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTiledSmem));
      attr_done = true;
    }
    splat_gather_tiled_kernel<MODE><<<grid, 256, kTiledSmem, st>>>(in, metric, out, ws, n, c, h, w);
  } else {
    dim3 grid(ceil_div((long long)h * w, 256), n);
    splat_gather_kernel<MODE><<<grid, 256, 0, st>>>(in, metric, out, ws, n, c, h, w);
  }
  MOTIF_LAUNCHED("splat_gather_kernel");
  return 0;
}

static int check_splat_args(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h, int w, int mode) {
  MOTIF_REQUIRE(in && flow && out, "splat: null pointer");
  MOTIF_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "splat: non-positive size n=%d c=%d h=%d w=%d", n, c, h, w);
  MOTIF_REQUIRE(mode >= MOTIF_SPLAT_SUMMATION && mode <= MOTIF_SPLAT_SOFTMAX, "splat: unknown mode %d", mode);
  MOTIF_REQUIRE(mode < MOTIF_SPLAT_LINEAR || metric != nullptr, "splat: mode %d needs a metric", mode);
  MOTIF_REQUIRE((long long)n * (c + 1) * h * w < (1LL << 40), "splat: tensor too large");
  MOTIF_REQUIRE((long long)h * w < (1LL << 31), "splat: image too large");
  return 0;
}

}  // namespace motif

using namespace motif;

extern "C" size_t motif_splat_workspace_bytes(int n, int h, int w) {
  if (n <= 0 || h <= 0 || w <= 0) return 0;
  return workspace_layout(n, h, w, nullptr, nullptr);
}

extern "C" int motif_splat_fwd_atomic(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h,
                                      int w, int mode, void* stream) {
  if (int rc = check_splat_args(in, flow, metric, out, n, c, h, w, mode)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int c_out = mode == MOTIF_SPLAT_SUMMATION ? c : c + 1;
  MOTIF_CUDA(cudaMemsetAsync(out, 0, (size_t)n * c_out * h * w * sizeof(float), st));
  switch (mode) {
    case MOTIF_SPLAT_SUMMATION: return launch_scatter<MOTIF_SPLAT_SUMMATION>(in, flow, metric, out, n, c, h, w, st);
    case MOTIF_SPLAT_AVERAGE: return launch_scatter<MOTIF_SPLAT_AVERAGE>(in, flow, metric, out, n, c, h, w, st);
    case MOTIF_SPLAT_LINEAR: return launch_scatter<MOTIF_SPLAT_LINEAR>(in, flow, metric, out, n, c, h, w, st);
    default: return launch_scatter<MOTIF_SPLAT_SOFTMAX>(in, flow, metric, out, n, c, h, w, st);
  }
}

extern "C" int motif_splat_fwd(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h, int w,
                               int mode, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_splat_args(in, flow, metric, out, n, c, h, w, mode)) return rc;
  MOTIF_REQUIRE(workspace != nullptr, "splat: null workspace");
  if (workspace_bytes < motif_splat_workspace_bytes(n, h, w))
    return fail(MOTIF_E_WORKSPACE, "splat: workspace %zu < %zu bytes", workspace_bytes, motif_splat_workspace_bytes(n, h, w));
  cudaStream_t st = (cudaStream_t)stream;
  SplatWorkspace ws;
  workspace_layout(n, h, w, &ws, (char*)workspace);
  const size_t p = (size_t)n * h * w;
  MOTIF_CUDA(cudaMemsetAsync(ws.count, 0, (size_t)((char*)ws.ovf_total - (char*)ws.count) + sizeof(int), st));  // count, ovf_mask, ovf_total
  {
    ProfScope prof("splat_bin_kernel", st);
    splat_bin_kernel<<<grid_for((long long)p, 256), 256, 0, st>>>(flow, ws, n, h, w);
    MOTIF_LAUNCHED("splat_bin_kernel");
  }
  int rc;
  switch (mode) {
    case MOTIF_SPLAT_SUMMATION:
      rc = launch_gather<MOTIF_SPLAT_SUMMATION>(in, metric, out, ws, n, c, h, w, st);
      if (!rc) rc = launch_surplus<MOTIF_SPLAT_SUMMATION>(in, flow, metric, out, n, c, h, w, ws, st);
      break;
    case MOTIF_SPLAT_AVERAGE:
      rc = launch_gather<MOTIF_SPLAT_AVERAGE>(in, metric, out, ws, n, c, h, w, st);
      if (!rc) rc = launch_surplus<MOTIF_SPLAT_AVERAGE>(in, flow, metric, out, n, c, h, w, ws, st);
      break;
    case MOTIF_SPLAT_LINEAR:
      rc = launch_gather<MOTIF_SPLAT_LINEAR>(in, metric, out, ws, n, c, h, w, st);
      if (!rc) rc = launch_surplus<MOTIF_SPLAT_LINEAR>(in, flow, metric, out, n, c, h, w, ws, st);
      break;
    default:
      rc = launch_gather<MOTIF_SPLAT_SOFTMAX>(in, metric, out, ws, n, c, h, w, st);
      if (!rc) rc = launch_surplus<MOTIF_SPLAT_SOFTMAX>(in, flow, metric, out, n, c, h, w, ws, st);
      break;
  }
  return rc;
}

extern "C" int motif_splat_max_fwd(const float* in, const float* flow, float* out, int n, int c, int h, int w, void* stream) {
  if (int rc = check_splat_args(in, flow, nullptr, out, n, c, h, w, MOTIF_SPLAT_SUMMATION)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)n * c * h * w;
  fill_kernel<<<grid_for((long long)total, 256), 256, 0, st>>>(out, 1.0f, total);
  MOTIF_LAUNCHED("fill_kernel");
  splat_max_kernel<<<grid_for((long long)n * h * w, 256), 256, 0, st>>>(in, flow, out, n, c, h, w);
  MOTIF_LAUNCHED("splat_max_kernel");
  return 0;
}

extern "C" int motif_splat_count_fwd(const float* flow, float* out, int n, int h, int w, void* stream) {
  MOTIF_REQUIRE(flow && out, "splat_count: null pointer");
  MOTIF_REQUIRE(n > 0 && h > 0 && w > 0, "splat_count: non-positive size");
  cudaStream_t st = (cudaStream_t)stream;
  MOTIF_CUDA(cudaMemsetAsync(out, 0, (size_t)n * h * w * sizeof(float), st));
  splat_count_kernel<<<grid_for((long long)n * h * w, 256), 256, 0, st>>>(flow, out, n, h, w);
  MOTIF_LAUNCHED("splat_count_kernel");
  return 0;
}
