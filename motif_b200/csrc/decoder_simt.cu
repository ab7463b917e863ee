// Space-time local implicit decoder, exact-fp32 CUDA-core pipeline (Ours.py:659-858 of the reference).
//
// Three kernels per decode, all one-thread-per-HR-pixel with the SIREN activations of a pixel kept in a
// private column of shared memory and the weights read through warp-uniform (broadcast) read-only loads:
//   imnet_kernel       once per clip and reference frame: gather nearest latent -> imnet -> imf[2B][qs][64]
//   flow_splat_kernel  per timestamp: gather -> flow_imnet -> (dx,dy,z) -> forward-splat BOTH references of the
//                      131-channel softmax splat, the max splat and the count splat into one pixel-major
//                      accumulator with vectorised red.global.add.v4.f32 (a warp cooperates on one source pixel
//                      at a time: 32 lanes x float4 = the 128 feature channels of one corner in one instruction)
//   synth_kernel       per timestamp: normalise/blend (Ours.py:810-836), append extra/residual/t, synth_net, clamp;
//                      re-arms the accumulators for the next timestamp while it still has them in cache.
// The 130-channel splat input, the 198-channel synth input and every 256-wide hidden activation of the
// reference never exist in HBM.  This path computes in true fp32 (FFMA + sinf) and is the numerical
// yardstick for the tcgen05 path in decoder_tc.cu.
#include "decoder_common.cuh"

namespace motif {

constexpr int kThreads = 256;  // threads per CTA == HR pixels per CTA

// act[k][tid]: column `tid` is private to the thread, so no barriers are needed between layers.
struct ActBuf {
  float v[64][kThreads];
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// acc[jj] = sum_k W[(j0+jj)*ldw + k] * in[k], 16 outputs x 64 inputs; W rows are warp-uniform addresses.
__device__ __forceinline__ void dense16(const float* __restrict__ W, int ldw, int j0, const float (&in)[64], float (&acc)[16]) {
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.0f;
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const float4 w = ldg4(W + (size_t)(j0 + jj) * ldw + 4 * k4);
      acc[jj] = fmaf(w.x, in[4 * k4 + 0], acc[jj]);
      acc[jj] = fmaf(w.y, in[4 * k4 + 1], acc[jj]);
      acc[jj] = fmaf(w.z, in[4 * k4 + 2], acc[jj]);
      acc[jj] = fmaf(w.w, in[4 * k4 + 3], acc[jj]);
    }
  }
}

__device__ __forceinline__ float siren_act(float pre) { return sinf(30.0f * pre); }  // SIREN.py:45

__device__ __forceinline__ void load_col(const ActBuf& a, float (&h)[64]) {
#pragma unroll
  for (int k = 0; k < 64; ++k) h[k] = a.v[k][threadIdx.x];
}

__device__ __forceinline__ void load_row64(const float* __restrict__ row, float (&h)[64]) {
#pragma unroll
  for (int k4 = 0; k4 < 16; ++k4) {
    const float4 v = ldg4(row + 4 * k4);
    h[4 * k4 + 0] = v.x;
    h[4 * k4 + 1] = v.y;
    h[4 * k4 + 2] = v.z;
    h[4 * k4 + 3] = v.w;
  }
}

// 64 -> 64 sine layer with plain bias: out column <- sin(30 * (W in + b))
__device__ __forceinline__ void sine_layer64(const float* __restrict__ W, const float* __restrict__ bias, const float (&in)[64], ActBuf& out) {
  for (int j0 = 0; j0 < 64; j0 += 16) {
    float acc[16];
    dense16(W, 64, j0, in, acc);
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) out.v[j0 + jj][threadIdx.x] = siren_act(acc[jj] + __ldg(bias + j0 + jj));
  }
}

// ---------------------------------------------------------------------------------------------------
// imnet: 66 -> 64 -> 64 -> 256 -> 64  (Ours.py:471, 737)
// ---------------------------------------------------------------------------------------------------
template <bool ENS>
__global__ void __launch_bounds__(kThreads, 1) imnet_kernel(motif_geom_t g, const float* __restrict__ feat, const float* __restrict__ wp,
                                                           float* __restrict__ imf, float* __restrict__ imf_low) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ActBuf& X = *reinterpret_cast<ActBuf*>(smem_raw);
  ActBuf& Y = *reinterpret_cast<ActBuf*>(smem_raw + sizeof(ActBuf));
  const int qs = g.HH * g.WW;
  const int q = blockIdx.x * kThreads + threadIdx.x;
  if (q >= qs) return;
  const int rb = blockIdx.y;  // r*B + b
  const int qy = q / g.WW, qx = q % g.WW;
  float ew[4] = {1.0f, 0.0f, 0.0f, 0.0f};
  if (ENS) ensemble_weights(qy, qx, g, ew);
  float4* dst = reinterpret_cast<float4*>(imf + ((size_t)rb * qs + q) * 64);
  float4* dst_low = ENS ? reinterpret_cast<float4*>(imf_low + ((size_t)rb * qs + q) * 64) : nullptr;
#pragma unroll 1
  for (int k = 0; k < (ENS ? 4 : 1); ++k) {
    const Query qu = ENS ? ensemble_query(qy, qx, g, k) : make_query(qy, qx, g);
    const float wk = pick4(ew, k);
    float h[64];
    load_row64(feat + ((size_t)rb * g.H * g.W + (size_t)qu.iy * g.W + qu.ix) * 64, h);
    if (ENS) {  // q_feat_low: ret = ret + feat * (area / tot_area), Ours.py:762 (product and sum rounded separately)
#pragma unroll
      for (int i4 = 0; i4 < 16; ++i4) {
        const float4 p = k == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : dst_low[i4];
        dst_low[i4] = make_float4(__fadd_rn(p.x, __fmul_rn(h[4 * i4], wk)), __fadd_rn(p.y, __fmul_rn(h[4 * i4 + 1], wk)),
                                  __fadd_rn(p.z, __fmul_rn(h[4 * i4 + 2], wk)), __fadd_rn(p.w, __fmul_rn(h[4 * i4 + 3], wk)));
      }
    }
    // layer 0: 64 gathered features on the FMA tile, rel_y / rel_x as rank-1 terms
    for (int j0 = 0; j0 < 64; j0 += 16) {
      float acc[16];
      dense16(wp + WeightPack::i_a0, 64, j0, h, acc);
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const float4 e = ldg4(wp + WeightPack::i_e0 + 4 * (j0 + jj));
        X.v[j0 + jj][threadIdx.x] = siren_act(fmaf(e.z, qu.rel_x, fmaf(e.y, qu.rel_y, acc[jj])) + e.x);
      }
    }
    load_col(X, h);
    sine_layer64(wp + WeightPack::i_a1, wp + WeightPack::i_b1, h, Y);
    // layers 2+3 fused: 64 hidden units at a time -> sine -> accumulate the 256 -> 64 output layer
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = __ldg(wp + WeightPack::i_b3 + i);
    for (int c = 0; c < 4; ++c) {
      load_col(Y, h);
      for (int j0 = 0; j0 < 64; j0 += 16) {
        float acc[16];
        dense16(wp + WeightPack::i_a2, 64, 64 * c + j0, h, acc);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) X.v[j0 + jj][threadIdx.x] = siren_act(acc[jj] + __ldg(wp + WeightPack::i_b2 + 64 * c + j0 + jj));
      }
      load_col(X, h);
      for (int i0 = 0; i0 < 64; i0 += 16) {
        float acc[16];
        dense16(wp + WeightPack::i_a3 + 64 * c, 256, i0, h, acc);
        // o[] must be indexed statically: unrolled select over the four 16-wide output groups
#pragma unroll
        for (int grp = 0; grp < 4; ++grp)
          if (i0 == 16 * grp) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) o[16 * grp + jj] += acc[jj];
          }
      }
    }
    if (!ENS) {
#pragma unroll
      for (int i4 = 0; i4 < 16; ++i4) dst[i4] = make_float4(o[4 * i4], o[4 * i4 + 1], o[4 * i4 + 2], o[4 * i4 + 3]);
    } else {
#pragma unroll
      for (int i4 = 0; i4 < 16; ++i4) {
        const float4 p = k == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : dst[i4];
        dst[i4] = make_float4(__fadd_rn(p.x, __fmul_rn(o[4 * i4], wk)), __fadd_rn(p.y, __fmul_rn(o[4 * i4 + 1], wk)),
                              __fadd_rn(p.z, __fmul_rn(o[4 * i4 + 2], wk)), __fadd_rn(p.w, __fmul_rn(o[4 * i4 + 3], wk)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// flow_imnet (67 -> 64 -> 64 -> 256 -> 3, Ours.py:470, 736) + the three forward splats of both references
// (Ours.py:777-806; softsplat_cp.py:320-347 softmax mode, softsplat_max_cp.py, softsplat_count_cp.py).
// ---------------------------------------------------------------------------------------------------
template <bool ENS>
__global__ void __launch_bounds__(kThreads, 1) flow_splat_kernel(motif_geom_t g, int B, int N, int n, float t, float alpha,
                                                                const float* __restrict__ feat, const float* __restrict__ flow_feat,
                                                                const float* __restrict__ imf, const float* __restrict__ wp,
                                                                DecodeScratch sc, float* __restrict__ flow_out, int b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ActBuf& X = *reinterpret_cast<ActBuf*>(smem_raw);
  ActBuf& Y = *reinterpret_cast<ActBuf*>(smem_raw + sizeof(ActBuf));
  const int qs = g.HH * g.WW;
  const int q = blockIdx.x * kThreads + threadIdx.x;
  const bool live = q < qs;
  const int qy = live ? q / g.WW : 0, qx = live ? q % g.WW : 0;
  const Query qu0 = make_query(qy, qx, g);
  const size_t lr = (size_t)qu0.iy * g.W + qu0.ix;
  const int lane = threadIdx.x & 31;
  float ew[4] = {1.0f, 0.0f, 0.0f, 0.0f};
  if (ENS) ensemble_weights(qy, qx, g, ew);

  for (int r = 0; r < 2; ++r) {
    const int rb = r * B + b;
    float bdx = 0.f, bdy = 0.f, bz = 0.f;  // blended prediction (Ours.py:758-764); the single one without the ensemble
#pragma unroll 1
    for (int k = 0; k < (ENS ? 4 : 1); ++k) {
    const Query qu = ENS ? ensemble_query(qy, qx, g, k) : qu0;
    const size_t lr = (size_t)qu.iy * g.W + qu.ix;
    float dx = 0.f, dy = 0.f, zraw = 0.f;
    if (live) {
      float h[64];
      load_row64(flow_feat + ((size_t)rb * g.H * g.W + lr) * 64, h);
      for (int j0 = 0; j0 < 64; j0 += 16) {
        float acc[16];
        dense16(wp + WeightPack::f_a0, 64, j0, h, acc);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float4 e = ldg4(wp + WeightPack::f_e0 + 4 * (j0 + jj));
          X.v[j0 + jj][threadIdx.x] = siren_act(fmaf(e.w, qu.rel_x, fmaf(e.z, qu.rel_y, fmaf(e.y, t, acc[jj]))) + e.x);
        }
      }
      load_col(X, h);
      sine_layer64(wp + WeightPack::f_a1, wp + WeightPack::f_b1, h, Y);
      load_col(Y, h);
      dx = __ldg(wp + WeightPack::f_b3 + 0);
      dy = __ldg(wp + WeightPack::f_b3 + 1);
      zraw = __ldg(wp + WeightPack::f_b3 + 2);
      for (int j0 = 0; j0 < 256; j0 += 16) {
        float acc[16];
        dense16(wp + WeightPack::f_a2, 64, j0, h, acc);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float s = siren_act(acc[jj] + __ldg(wp + WeightPack::f_b2 + j0 + jj));
          dx = fmaf(s, __ldg(wp + WeightPack::f_a3 + 0 * 256 + j0 + jj), dx);
          dy = fmaf(s, __ldg(wp + WeightPack::f_a3 + 1 * 256 + j0 + jj), dy);
          zraw = fmaf(s, __ldg(wp + WeightPack::f_a3 + 2 * 256 + j0 + jj), zraw);
        }
      }
    }
    if (ENS) {
      const float wk = pick4(ew, k);
      bdx = __fadd_rn(bdx, __fmul_rn(dx, wk)), bdy = __fadd_rn(bdy, __fmul_rn(dy, wk)), bz = __fadd_rn(bz, __fmul_rn(zraw, wk));
    } else {
      bdx = dx, bdy = dy, bz = zraw;
    }
    }
    const float dx = bdx, dy = bdy, zraw = bz;
    // Ours.py:794: flow = raw * 20. * (HH / H);  z = relu(raw_z) * alpha
    const float fx = __fmul_rn(__fmul_rn(dx, 20.0f), g.flow_scale);
    const float fy = __fmul_rn(__fmul_rn(dy, 20.0f), g.flow_scale);
    const float z = __fmul_rn(fmaxf(zraw, 0.0f), alpha);
    const float e = expf(z);
    if (live && flow_out != nullptr) {  // Ours.py:858: flow / 20.0 / (HH / H)
      float* fo = flow_out + ((size_t)(rb * N + n) * 2) * qs + q;
      fo[0] = __fdiv_rn(__fdiv_rn(fx, 20.0f), g.flow_scale);
      fo[qs] = __fdiv_rn(__fdiv_rn(fy, 20.0f), g.flow_scale);
    }
    Footprint f = footprint(qx, qy, fx, fy);
    if (!live) f.finite = false;

    // warp-cooperative scatter: lanes 0-15 carry imnet(q) (64 ch), lanes 16-31 the nearest latent (64 ch)
    for (int p = 0; p < 32; ++p) {
      if (!__shfl_sync(0xffffffffu, (int)f.finite, p)) continue;
      const int sx0 = __shfl_sync(0xffffffffu, f.x0, p), sy0 = __shfl_sync(0xffffffffu, f.y0, p);
      const float se = __shfl_sync(0xffffffffu, e, p);
      const int sq = __shfl_sync(0xffffffffu, q, p);
      const size_t slr = __shfl_sync(0xffffffffu, (unsigned long long)lr, p);
      const float sdx = __shfl_sync(0xffffffffu, dx, p), sdy = __shfl_sync(0xffffffffu, dy, p);
      float w4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) w4[k] = __shfl_sync(0xffffffffu, f.w[k], p);
      const float* srow = lane < 16 ? imf + ((size_t)rb * qs + sq) * 64 + 4 * lane
                          : (ENS ? sc.imf_low + ((size_t)rb * qs + sq) * 64 + 4 * (lane - 16)
                                 : feat + ((size_t)rb * g.H * g.W + slr) * 64 + 4 * (lane - 16));
      float4 v = ldg4(srow);
      // softsplat_cp.py:332: tenInput * tenMetric.exp() is rounded before the kernel multiplies by the weight
      v.x = __fmul_rn(v.x, se);
      v.y = __fmul_rn(v.y, se);
      v.z = __fmul_rn(v.z, se);
      v.w = __fmul_rn(v.w, se);
      const float edx = __fmul_rn(sdx, se), edy = __fmul_rn(sdy, se);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cx = sx0 + (k & 1), cy = sy0 + (k >> 1);
        if ((cx < 0) | (cx >= g.WW) | (cy < 0) | (cy >= g.HH)) continue;
        const size_t d = (size_t)b * qs + (size_t)cy * g.WW + cx;
        const float wk = w4[k];
        red_add_v4(sc.acc_main + d * 128 + 4 * lane, __fmul_rn(v.x, wk), __fmul_rn(v.y, wk), __fmul_rn(v.z, wk), __fmul_rn(v.w, wk));
        if (lane == 0) red_add_v4(sc.acc_side + d * 4, __fmul_rn(edx, wk), __fmul_rn(edy, wk), __fmul_rn(se, wk), 1.0f);
        if (lane == 1) red_max_nonneg(sc.acc_max + d, __fmul_rn(se, wk));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// blend (Ours.py:810-836) + synth_net (198 -> 64 -> 64 -> 64 -> 256 -> 3, Ours.py:487-491, 839-858) + clamp
// ---------------------------------------------------------------------------------------------------
template <bool ENS>
__global__ void __launch_bounds__(kThreads, 1) synth_kernel(motif_geom_t g, int B, int N, int n, float t, const float* __restrict__ residual,
                                                           const float* __restrict__ wp, DecodeScratch sc, float* __restrict__ rgb,
                                                           float* __restrict__ dbg_in, int b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ActBuf& X = *reinterpret_cast<ActBuf*>(smem_raw);
  ActBuf& Y = *reinterpret_cast<ActBuf*>(smem_raw + sizeof(ActBuf));
  const int qs = g.HH * g.WW;
  const int q = blockIdx.x * kThreads + threadIdx.x;
  if (q >= qs) return;
  const Query qu = make_query(q / g.WW, q % g.WW, g);
  const size_t d = (size_t)b * qs + q;
  const int bn = b * N + n;

  float4* side_p = reinterpret_cast<float4*>(sc.acc_side + d * 4);
  const float4 side = *side_p;
  const float zmax = sc.acc_max[d];
  *side_p = make_float4(0.f, 0.f, 0.f, 0.f);
  sc.acc_max[d] = 1.0f;
  // Ours.py:813-814: warped_z[warped_z == 0] = 1; output /= warped_z
  const float wz = side.z == 0.0f ? 1.0f : side.z;
  const float cnt = side.w;
  // Ours.py:826-829: count_ (0 -> 1), warped_z_ (== 1.0 -> 0)
  const float cnt_ = cnt == 0.0f ? 1.0f : cnt;
  const float wz_ = wz == 1.0f ? 0.0f : wz;
  float ex[8];
  ex[0] = 1.0f;                         // bias
  ex[1] = __fdiv_rn(side.x, wz);        // dx'
  ex[2] = __fdiv_rn(side.y, wz);        // dy'
  ex[3] = zmax;                         // Ours.py:834 extra = [z_max, count / 16, warped_z_ / count_]
  ex[4] = __fdiv_rn(cnt, 16.0f);
  ex[5] = __fdiv_rn(wz_, cnt_);
  ex[6] = t;
  ex[7] = 0.0f;
  float* dbg = dbg_in ? dbg_in + (size_t)bn * 198 * qs + q : nullptr;
  if (dbg) {
    dbg[(size_t)64 * qs] = ex[1];
    dbg[(size_t)65 * qs] = ex[2];
    dbg[(size_t)130 * qs] = ex[3];
    dbg[(size_t)131 * qs] = ex[4];
    dbg[(size_t)132 * qs] = ex[5];
    dbg[(size_t)197 * qs] = t;
  }

  // layer 0 as three 64-wide K blocks accumulated in X, then the rank-1 extras and the sine
  float h[64];
  float4* main_p = reinterpret_cast<float4*>(sc.acc_main + d * 128);
  for (int kb = 0; kb < 3; ++kb) {
    if (kb < 2) {
#pragma unroll
      for (int k4 = 0; k4 < 16; ++k4) {
        const float4 v = main_p[16 * kb + k4];
        main_p[16 * kb + k4] = make_float4(0.f, 0.f, 0.f, 0.f);
        h[4 * k4 + 0] = __fdiv_rn(v.x, wz);
        h[4 * k4 + 1] = __fdiv_rn(v.y, wz);
        h[4 * k4 + 2] = __fdiv_rn(v.z, wz);
        h[4 * k4 + 3] = __fdiv_rn(v.w, wz);
      }
    } else if (!ENS) {
      load_row64(residual + ((size_t)b * g.H * g.W + (size_t)qu.iy * g.W + qu.ix) * 64, h);
    } else {  // q_residual blended over the four latents (Ours.py:762)
      float ew[4];
      ensemble_weights(q / g.WW, q % g.WW, g, ew);
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const Query qk = ensemble_query(q / g.WW, q % g.WW, g, k);
        const float wk = pick4(ew, k);
        const float* row = residual + ((size_t)b * g.H * g.W + (size_t)qk.iy * g.W + qk.ix) * 64;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          const float4 v = ldg4(row + 4 * k4);
          h[4 * k4 + 0] = __fadd_rn(k == 0 ? 0.f : h[4 * k4 + 0], __fmul_rn(v.x, wk));
          h[4 * k4 + 1] = __fadd_rn(k == 0 ? 0.f : h[4 * k4 + 1], __fmul_rn(v.y, wk));
          h[4 * k4 + 2] = __fadd_rn(k == 0 ? 0.f : h[4 * k4 + 2], __fmul_rn(v.z, wk));
          h[4 * k4 + 3] = __fadd_rn(k == 0 ? 0.f : h[4 * k4 + 3], __fmul_rn(v.w, wk));
        }
      }
    }
    if (dbg) {
      const int c0 = kb == 0 ? 0 : (kb == 1 ? 66 : 133);
#pragma unroll
      for (int k = 0; k < 64; ++k) dbg[(size_t)(c0 + k) * qs] = h[k];
    }
    const int off = kb == 0 ? WeightPack::s_a0a : (kb == 1 ? WeightPack::s_a0b : WeightPack::s_a0c);
    for (int j0 = 0; j0 < 64; j0 += 16) {
      float acc[16];
      dense16(wp + off, 64, j0, h, acc);
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        if (kb == 0) X.v[j0 + jj][threadIdx.x] = acc[jj];
        else X.v[j0 + jj][threadIdx.x] += acc[jj];
      }
    }
  }
  for (int j = 0; j < 64; ++j) {
    const float4 e0 = ldg4(wp + WeightPack::s_e0 + 8 * j), e1 = ldg4(wp + WeightPack::s_e0 + 8 * j + 4);
    float pre = X.v[j][threadIdx.x];
    pre = fmaf(e0.y, ex[1], pre);
    pre = fmaf(e0.z, ex[2], pre);
    pre = fmaf(e0.w, ex[3], pre);
    pre = fmaf(e1.x, ex[4], pre);
    pre = fmaf(e1.y, ex[5], pre);
    pre = fmaf(e1.z, ex[6], pre);
    Y.v[j][threadIdx.x] = siren_act(pre + e0.x);
  }
  load_col(Y, h);
  sine_layer64(wp + WeightPack::s_a1, wp + WeightPack::s_b1, h, X);
  load_col(X, h);
  sine_layer64(wp + WeightPack::s_a2, wp + WeightPack::s_b2, h, Y);
  load_col(Y, h);
  float o0 = __ldg(wp + WeightPack::s_b4 + 0), o1 = __ldg(wp + WeightPack::s_b4 + 1), o2 = __ldg(wp + WeightPack::s_b4 + 2);
  for (int j0 = 0; j0 < 256; j0 += 16) {
    float acc[16];
    dense16(wp + WeightPack::s_a3, 64, j0, h, acc);
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const float s = siren_act(acc[jj] + __ldg(wp + WeightPack::s_b3 + j0 + jj));
      o0 = fmaf(s, __ldg(wp + WeightPack::s_a4 + 0 * 256 + j0 + jj), o0);
      o1 = fmaf(s, __ldg(wp + WeightPack::s_a4 + 1 * 256 + j0 + jj), o1);
      o2 = fmaf(s, __ldg(wp + WeightPack::s_a4 + 2 * 256 + j0 + jj), o2);
    }
  }
  // output [N, B, 3, HH, WW] clamped (Ours.py:853-858)
  float* out = rgb + ((size_t)(n * B + b) * 3) * qs + q;
  out[0] = fminf(fmaxf(o0, 0.0f), 1.0f);
  out[(size_t)qs] = fminf(fmaxf(o1, 0.0f), 1.0f);
  out[(size_t)2 * qs] = fminf(fmaxf(o2, 0.0f), 1.0f);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
__global__ void geometry_kernel(motif_geom_t g, int32_t* iy, int32_t* ix, float* coord, float* rel) {
  const int qs = g.HH * g.WW;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= qs) return;
  const Query qu = make_query(q / g.WW, q % g.WW, g);
  if (iy) iy[q] = qu.iy;
  if (ix) ix[q] = qu.ix;
  if (coord) {
    coord[2 * q] = qu.cy;
    coord[2 * q + 1] = qu.cx;
  }
  if (rel) {
    rel[2 * q] = qu.rel_y;
    rel[2 * q + 1] = qu.rel_x;
  }
}

// NCHW [rows][C][hw] -> [rows][hw][C] through a 32x32 shared tile
// pixels [p_begin, p_end) of every plane (the LR rows a destination row band needs; everything by default)
__global__ void pack_latents_kernel(const float* __restrict__ in, float* __restrict__ out, int c, int hw, int p_begin, int p_end) {
  __shared__ float tile[32][33];
  const int row = blockIdx.z;
  const int p0 = p_begin + blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (size_t)row * c * hw;
  float* dst = out + (size_t)row * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int ch = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (ch < c && p < p_end) ? src[(size_t)ch * hw + p] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, ch = c0 + threadIdx.x;
    if (p < p_end && ch < c) dst[(size_t)p * c + ch] = tile[threadIdx.x][i];
  }
}

struct PackJob {
  const float* src;  // [rows][ld]
  int dst_off, rows, cols, ld, col0, dst_ld;
};
constexpr int kMaxJobs = 48;
struct PackJobs {
  PackJob j[kMaxJobs];
  int n;
};

__global__ void pack_weights_kernel(PackJobs jobs, float* __restrict__ wp) {
  const PackJob jb = jobs.j[blockIdx.x];
  for (int i = threadIdx.x; i < jb.rows * jb.cols; i += blockDim.x) {
    const int r = i / jb.cols, cidx = i % jb.cols;
    wp[jb.dst_off + r * jb.dst_ld + cidx] = jb.src[(size_t)r * jb.ld + jb.col0 + cidx];
  }
}

int pack_weights(const motif_decode_t* a, float* wpack, cudaStream_t st) {
  using P = WeightPack;
  const motif_siren_t &F = a->flow_imnet, &I = a->imnet, &S = a->synth_net;
  MOTIF_REQUIRE(F.n_layers == 4 && I.n_layers == 4 && S.n_layers == 5, "decode: unexpected SIREN depth (%d, %d, %d)", F.n_layers, I.n_layers, S.n_layers);
  for (int l = 0; l < 5; ++l) {
    if (l < 4) MOTIF_REQUIRE(F.weight[l] && F.bias[l] && I.weight[l] && I.bias[l], "decode: null weight pointer (layer %d)", l);
    MOTIF_REQUIRE(S.weight[l] && S.bias[l], "decode: null synth_net weight pointer (layer %d)", l);
  }
  PackJobs jobs;
  jobs.n = 0;
  auto add = [&](const float* src, int dst_off, int rows, int cols, int ld, int col0, int dst_ld) {
    jobs.j[jobs.n++] = PackJob{src, dst_off, rows, cols, ld, col0, dst_ld};
  };
  // flow_imnet: weight[0] is [64][67]
  add(F.weight[0], P::f_a0, 64, 64, 67, 0, 64);
  add(F.bias[0], P::f_e0 + 0, 64, 1, 1, 0, 4);
  add(F.weight[0], P::f_e0 + 1, 64, 3, 67, 64, 4);
  add(F.weight[1], P::f_a1, 64, 64, 64, 0, 64);
  add(F.bias[1], P::f_b1, 1, 64, 64, 0, 64);
  add(F.weight[2], P::f_a2, 256, 64, 64, 0, 64);
  add(F.bias[2], P::f_b2, 1, 256, 256, 0, 256);
  add(F.weight[3], P::f_a3, 3, 256, 256, 0, 256);
  add(F.bias[3], P::f_b3, 1, 3, 3, 0, 4);
  // imnet: weight[0] is [64][66]
  add(I.weight[0], P::i_a0, 64, 64, 66, 0, 64);
  add(I.bias[0], P::i_e0 + 0, 64, 1, 1, 0, 4);
  add(I.weight[0], P::i_e0 + 1, 64, 2, 66, 64, 4);
  add(I.weight[1], P::i_a1, 64, 64, 64, 0, 64);
  add(I.bias[1], P::i_b1, 1, 64, 64, 0, 64);
  add(I.weight[2], P::i_a2, 256, 64, 64, 0, 64);
  add(I.bias[2], P::i_b2, 1, 256, 256, 0, 256);
  add(I.weight[3], P::i_a3, 64, 256, 256, 0, 256);
  add(I.bias[3], P::i_b3, 1, 64, 64, 0, 64);
  // synth_net: weight[0] is [64][198]
  add(S.weight[0], P::s_a0a, 64, 64, 198, 0, 64);
  add(S.weight[0], P::s_a0b, 64, 64, 198, 66, 64);
  add(S.weight[0], P::s_a0c, 64, 64, 198, 133, 64);
  add(S.bias[0], P::s_e0 + 0, 64, 1, 1, 0, 8);
  add(S.weight[0], P::s_e0 + 1, 64, 2, 198, 64, 8);   // dx', dy'
  add(S.weight[0], P::s_e0 + 3, 64, 3, 198, 130, 8);  // zmax, cnt/16, wz/cnt
  add(S.weight[0], P::s_e0 + 6, 64, 1, 198, 197, 8);  // t
  add(S.weight[1], P::s_a1, 64, 64, 64, 0, 64);
  add(S.bias[1], P::s_b1, 1, 64, 64, 0, 64);
  add(S.weight[2], P::s_a2, 64, 64, 64, 0, 64);
  add(S.bias[2], P::s_b2, 1, 64, 64, 0, 64);
  add(S.weight[3], P::s_a3, 256, 64, 64, 0, 64);
  add(S.bias[3], P::s_b3, 1, 256, 256, 0, 256);
  add(S.weight[4], P::s_a4, 3, 256, 256, 0, 256);
  add(S.bias[4], P::s_b4, 1, 3, 3, 0, 4);
  MOTIF_CUDA(cudaMemsetAsync(wpack, 0, sizeof(float) * P::total, st));
  pack_weights_kernel<<<jobs.n, 256, 0, st>>>(jobs, wpack);
  MOTIF_LAUNCHED("pack_weights_kernel");
  return 0;
}

int decode_layout(int B, int N, int H, int W, int HH, int WW, DecodeScratch* s, char* base, size_t* bytes) {
  (void)N; (void)H; (void)W;
  const size_t qs = (size_t)HH * WW;
  size_t off = 0;
  auto take = [&](size_t nbytes) {
    char* p = base ? base + off : nullptr;
    off += (nbytes + 255) & ~size_t(255);
    return (float*)p;
  };
  float* imf = take(sizeof(float) * 2 * B * qs * 64);
  float* iml = take(sizeof(float) * 2 * B * qs * 64);
  float* am = take(sizeof(float) * B * qs * 128);
  float* as = take(sizeof(float) * B * qs * 4);
  float* ax = take(sizeof(float) * B * qs);
  float* wp = take(sizeof(float) * WeightPack::total);
  float* wi = take(tc_image_bytes());
  if (s) *s = DecodeScratch{imf, iml, am, as, ax, wp, wi};
  if (bytes) *bytes = off;
  return 0;
}

__global__ void fill_ones_kernel(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1.0f;
}

int check_decode(const motif_decode_t* a) {
  MOTIF_REQUIRE(a != nullptr, "decode: null args");
  const motif_geom_t& g = a->geom;
  MOTIF_REQUIRE(g.B > 0 && g.N > 0 && g.H > 0 && g.W > 0 && g.HH > 0 && g.WW > 0, "decode: non-positive size");
  MOTIF_REQUIRE(g.seq_hh && g.seq_ww && g.seq_h && g.seq_w, "decode: null coordinate sequence");
  MOTIF_REQUIRE(a->feat && a->flow_feat && a->residual && a->target_t && a->rgb, "decode: null tensor pointer");
  MOTIF_REQUIRE(a->workspace != nullptr, "decode: null workspace");
  MOTIF_REQUIRE(a->n_begin >= 0 && a->n_end <= g.N && a->n_begin <= a->n_end, "decode: bad timestamp range [%d,%d)", a->n_begin, a->n_end);
  MOTIF_REQUIRE((long long)g.HH * g.WW < (1LL << 30), "decode: HR image too large");
  return 0;
}

int decode_simt(const motif_decode_t* a, cudaStream_t st) {
  const motif_geom_t& g = a->geom;
  DecodeScratch sc;
  size_t need = 0;
  decode_layout(g.B, g.N, g.H, g.W, g.HH, g.WW, &sc, (char*)a->workspace, &need);
  if (a->workspace_bytes < need) return fail(MOTIF_E_WORKSPACE, "decode: workspace %zu < %zu bytes", a->workspace_bytes, need);
  if (int rc = pack_weights(a, sc.wpack, st)) return rc;
  const int qs = g.HH * g.WW;
  const size_t smem = 2 * sizeof(ActBuf);
  static bool attr_done_dev[64] = {false};
  bool& attr_done = attr_done_dev[current_device_slot()];
  if (!attr_done) {
    MOTIF_CUDA(cudaFuncSetAttribute(imnet_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTIF_CUDA(cudaFuncSetAttribute(flow_splat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTIF_CUDA(cudaFuncSetAttribute(synth_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTIF_CUDA(cudaFuncSetAttribute(imnet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTIF_CUDA(cudaFuncSetAttribute(flow_splat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOTIF_CUDA(cudaFuncSetAttribute(synth_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int tiles = ceil_div(qs, kThreads);
  const bool ens = a->local_ensemble != 0;
  if (a->n_begin == a->n_end) return 0;
  {
    ProfScope prof("imnet_kernel", st);
    if (ens) imnet_kernel<true><<<dim3(tiles, 2 * g.B), kThreads, smem, st>>>(g, a->feat, sc.wpack, sc.imf, sc.imf_low);
    else imnet_kernel<false><<<dim3(tiles, 2 * g.B), kThreads, smem, st>>>(g, a->feat, sc.wpack, sc.imf, sc.imf_low);
    MOTIF_LAUNCHED("imnet_kernel");
  }
  MOTIF_CUDA(cudaMemsetAsync(sc.acc_main, 0, sizeof(float) * (size_t)g.B * qs * 128, st));
  MOTIF_CUDA(cudaMemsetAsync(sc.acc_side, 0, sizeof(float) * (size_t)g.B * qs * 4, st));
  fill_ones_kernel<<<148 * 8, 256, 0, st>>>(sc.acc_max, (size_t)g.B * qs);
  MOTIF_LAUNCHED("fill_ones_kernel");
  for (int b = 0; b < g.B; ++b)
    for (int n = a->n_begin; n < a->n_end; ++n) {
      const float t = a->target_t[b * g.N + n];
      {
        ProfScope prof("flow_splat_kernel", st);
        if (ens) flow_splat_kernel<true><<<tiles, kThreads, smem, st>>>(g, g.B, g.N, n, t, a->alpha, a->feat, a->flow_feat, sc.imf, sc.wpack, sc, a->flow_out, b);
        else flow_splat_kernel<false><<<tiles, kThreads, smem, st>>>(g, g.B, g.N, n, t, a->alpha, a->feat, a->flow_feat, sc.imf, sc.wpack, sc, a->flow_out, b);
        MOTIF_LAUNCHED("flow_splat_kernel");
      }
      {
        ProfScope prof("synth_kernel", st);
        if (ens) synth_kernel<true><<<tiles, kThreads, smem, st>>>(g, g.B, g.N, n, t, a->residual, sc.wpack, sc, a->rgb, a->dbg_synth_in, b);
        else synth_kernel<false><<<tiles, kThreads, smem, st>>>(g, g.B, g.N, n, t, a->residual, sc.wpack, sc, a->rgb, a->dbg_synth_in, b);
        MOTIF_LAUNCHED("synth_kernel");
      }
    }
  return 0;
}

}  // namespace motif

using namespace motif;

extern "C" int motif_query_geometry(const motif_geom_t* g, int32_t* iy, int32_t* ix, float* coord, float* rel, void* stream) {
  MOTIF_REQUIRE(g && g->seq_hh && g->seq_ww && g->seq_h && g->seq_w, "query_geometry: null pointer");
  MOTIF_REQUIRE(g->H > 0 && g->W > 0 && g->HH > 0 && g->WW > 0, "query_geometry: non-positive size");
  const int qs = g->HH * g->WW;
  geometry_kernel<<<ceil_div(qs, 256), 256, 0, (cudaStream_t)stream>>>(*g, iy, ix, coord, rel);
  MOTIF_LAUNCHED("geometry_kernel");
  return 0;
}

extern "C" int motif_pack_latents(const float* nchw, float* packed, int rows, int channels, int hw, void* stream) {
  MOTIF_REQUIRE(nchw && packed, "pack_latents: null pointer");
  MOTIF_REQUIRE(rows > 0 && channels > 0 && hw > 0 && rows <= 65535, "pack_latents: bad size");
  dim3 grid(ceil_div(hw, 32), ceil_div(channels, 32), rows);
  pack_latents_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(nchw, packed, channels, hw, 0, hw);
  MOTIF_LAUNCHED("pack_latents_kernel");
  return 0;
}

extern "C" int motif_pack_latents_range(const float* nchw, float* packed, int rows, int channels, int hw, int p_begin, int p_end, void* stream) {
  MOTIF_REQUIRE(nchw && packed, "pack_latents: null pointer");
  MOTIF_REQUIRE(rows > 0 && channels > 0 && hw > 0 && rows <= 65535 && p_begin >= 0 && p_begin <= p_end && p_end <= hw, "pack_latents: bad size or range");
  if (p_begin == p_end) return 0;
  dim3 grid(ceil_div(p_end - p_begin, 32), ceil_div(channels, 32), rows);
  pack_latents_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(nchw, packed, channels, hw, p_begin, p_end);
  MOTIF_LAUNCHED("pack_latents_kernel");
  return 0;
}

extern "C" size_t motif_decode_workspace_bytes(int B, int N, int H, int W, int HH, int WW) {
  if (B <= 0 || N <= 0 || H <= 0 || W <= 0 || HH <= 0 || WW <= 0) return 0;
  size_t bytes = 0;
  decode_layout(B, N, H, W, HH, WW, nullptr, nullptr, &bytes);
  const size_t b16 = decode_f16_workspace_bytes(B, N, H, W, HH, WW);  // one workspace serves every precision
  return bytes > b16 ? bytes : b16;
}
