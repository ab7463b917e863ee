// Pieces shared by the decoder kernels: query geometry (bit-exact with the reference), the HBM layouts of
// the decoder scratch buffers and the packed weight image.
#pragma once
#include "common.cuh"

namespace motif {

// -----------------------------------------------------------------------------------------------------
// Query geometry of one HR pixel: Ours.py:667-689 (coord + eps, clamp), ATen grid_sample nearest with
// align_corners=False (Ours.py:704: index = nearbyint(((c + 1) * size - 1) / 2)), Ours.py:720-722 (rel).
// The four 1-D sequences are the ones make_coord builds on the host (passed in, not recomputed), every
// other operation is an explicitly rounded fp32 op so that nvcc cannot contract what the reference did
// not; the un-normalisation uses the same fused multiply-add the CUDA build of ATen compiles to.
// -----------------------------------------------------------------------------------------------------
struct Query {
  int iy, ix;          // nearest LR latent
  float cy, cx;        // shifted + clamped coordinate (what grid_sample saw)
  float rel_y, rel_x;  // (hr_coord - q_coord) * (H, W)
};

__device__ __forceinline__ int nearest_index(float c, int size) {
  const float u = fmaf(__fadd_rn(c, 1.0f), (float)size, -1.0f) * 0.5f;
  int i = __float2int_rn(u);  // nearbyint: ties to even
  return min(max(i, 0), size - 1);
}

__device__ __forceinline__ Query make_query(int qy, int qx, const motif_geom_t& g) {
  const float lo = (float)(-1 + 1e-6), hi = (float)(1 - 1e-6);
  const float hy = __ldg(g.seq_hh + qy), hx = __ldg(g.seq_ww + qx);
  Query q;
  q.cy = fminf(fmaxf(__fadd_rn(hy, 1e-6f), lo), hi);
  q.cx = fminf(fmaxf(__fadd_rn(hx, 1e-6f), lo), hi);
  q.iy = nearest_index(q.cy, g.H);
  q.ix = nearest_index(q.cx, g.W);
  q.rel_y = __fmul_rn(__fsub_rn(hy, __ldg(g.seq_h + q.iy)), (float)g.H);
  q.rel_x = __fmul_rn(__fsub_rn(hx, __ldg(g.seq_w + q.ix)), (float)g.W);
  return q;
}

// Local ensemble (LunaTokis.local_ensemble = True; False as shipped, Ours.py:453): the query is evaluated at the four
// latents reached by shifting the coordinate by (v0 / H, v1 / W), v in {-1, +1}^2 (Ours.py:660-663, 686-689), and
// the four predictions are blended with the area weights of the DIAGONALLY opposite latent (Ours.py:754-764).
// The shift `v * r + 1e-6` is formed in python double (r = 2 / size / 2) and cast to fp32 by the in-place add.
__device__ __forceinline__ Query make_query_shifted(int qy, int qx, const motif_geom_t& g, int v0, int v1) {
  const float lo = (float)(-1 + 1e-6), hi = (float)(1 - 1e-6);
  const float hy = __ldg(g.seq_hh + qy), hx = __ldg(g.seq_ww + qx);
  const float s0 = (float)__dadd_rn(__dmul_rn((double)v0, __ddiv_rn(__ddiv_rn(2.0, (double)g.H), 2.0)), 1e-6);
  const float s1 = (float)__dadd_rn(__dmul_rn((double)v1, __ddiv_rn(__ddiv_rn(2.0, (double)g.W), 2.0)), 1e-6);
  Query q;
  q.cy = fminf(fmaxf(__fadd_rn(hy, s0), lo), hi);
  q.cx = fminf(fmaxf(__fadd_rn(hx, s1), lo), hi);
  q.iy = nearest_index(q.cy, g.H);
  q.ix = nearest_index(q.cx, g.W);
  q.rel_y = __fmul_rn(__fsub_rn(hy, __ldg(g.seq_h + q.iy)), (float)g.H);
  q.rel_x = __fmul_rn(__fsub_rn(hx, __ldg(g.seq_w + q.ix)), (float)g.W);
  return q;
}
// shift k of the reference's loop nest `for vx in [-1, 1]: for vy in [-1, 1]` (vx moves coordinate 0 = y)
__device__ __forceinline__ Query ensemble_query(int qy, int qx, const motif_geom_t& g, int k) {
  return make_query_shifted(qy, qx, g, (k >> 1) ? 1 : -1, (k & 1) ? 1 : -1);
}
// w[k] = area[3 - k] / (area[0] + area[1] + area[2] + area[3]),  area = |rel_y * rel_x| + 1e-9
__device__ __forceinline__ void ensemble_weights(int qy, int qx, const motif_geom_t& g, float (&w)[4]) {
  float area[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const Query q = ensemble_query(qy, qx, g, k);
    area[k] = __fadd_rn(fabsf(__fmul_rn(q.rel_y, q.rel_x)), 1e-9f);
  }
  const float tot = __fadd_rn(__fadd_rn(__fadd_rn(area[0], area[1]), area[2]), area[3]);
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = __fdiv_rn(area[3 - k], tot);
}
__device__ __forceinline__ float pick4(const float (&w)[4], int k) { return k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3])); }

// -----------------------------------------------------------------------------------------------------
// Scratch layout (all fp32, pixel-major so that one pixel's channels are contiguous):
//   imf      [2B][qs][64]   imnet output per reference frame (clip-invariant)
//   imf_low  [2B][qs][64]   local ensemble only: the blended nearest latents (q_feat_low, Ours.py:723, 762)
//   acc_main [B][qs][128]   sum-splat accumulator of one timestamp: imnet' (64) | nearest feat' (64),
//                           both references accumulate into the same cell (Ours.py:811 sums them anyway)
//   acc_side [B][qs][4]     (sum e*dx, sum e*dy, sum e [the normaliser], count)
//   acc_max  [B][qs]        max splat of exp(z), starts at 1.0
//   wpack                   packed weights (see WeightPack)
// -----------------------------------------------------------------------------------------------------
struct DecodeScratch {
  float* imf;
  float* imf_low;
  float* acc_main;
  float* acc_side;
  float* acc_max;
  float* wpack;
  float* wimg;  // tensor-core weight block images (tc_pack.cuh)
};

// Offsets (in floats) into the packed weight image.  K is split into 64-wide blocks that multiply
// gathered/blended feature vectors, and a few "extra" columns (t, rel, dx', dy', zmax, ...) that are applied
// as rank-1 updates together with the bias.  Every [out][64] block is row-major with 64 contiguous inputs.
struct WeightPack {
  // flow_imnet 67 -> 64 -> 64 -> 256 -> 3       input cols: [flow_feat 0..63 | t 64 | rel_y 65 | rel_x 66]
  static constexpr int f_a0 = 0;                       // [64][64]
  static constexpr int f_e0 = f_a0 + 64 * 64;          // [64][4]  bias, w_t, w_rely, w_relx
  static constexpr int f_a1 = f_e0 + 64 * 4;           // [64][64]
  static constexpr int f_b1 = f_a1 + 64 * 64;          // [64]
  static constexpr int f_a2 = f_b1 + 64;               // [256][64]
  static constexpr int f_b2 = f_a2 + 256 * 64;         // [256]
  static constexpr int f_a3 = f_b2 + 256;              // [3][256]
  static constexpr int f_b3 = f_a3 + 3 * 256;          // [4]
  // imnet 66 -> 64 -> 64 -> 256 -> 64           input cols: [feat 0..63 | rel_y 64 | rel_x 65]
  static constexpr int i_a0 = f_b3 + 4;                // [64][64]
  static constexpr int i_e0 = i_a0 + 64 * 64;          // [64][4]  bias, w_rely, w_relx, 0
  static constexpr int i_a1 = i_e0 + 64 * 4;
  static constexpr int i_b1 = i_a1 + 64 * 64;
  static constexpr int i_a2 = i_b1 + 64;               // [256][64]
  static constexpr int i_b2 = i_a2 + 256 * 64;
  static constexpr int i_a3 = i_b2 + 256;              // [64][256]
  static constexpr int i_b3 = i_a3 + 64 * 256;         // [64]
  // synth_net 198 -> 64 -> 64 -> 64 -> 256 -> 3  input cols (Ours.py:788-791, 834, 839-844):
  //   [imnet' 0..63 | dx' 64 | dy' 65 | feat' 66..129 | zmax 130 | cnt/16 131 | wz/cnt 132 | residual 133..196 | t 197]
  static constexpr int s_a0a = i_b3 + 64;              // [64][64] cols 0..63
  static constexpr int s_a0b = s_a0a + 64 * 64;        // [64][64] cols 66..129
  static constexpr int s_a0c = s_a0b + 64 * 64;        // [64][64] cols 133..196
  static constexpr int s_e0 = s_a0c + 64 * 64;         // [64][8]  bias, w_dx, w_dy, w_zmax, w_cnt, w_wz, w_t, 0
  static constexpr int s_a1 = s_e0 + 64 * 8;
  static constexpr int s_b1 = s_a1 + 64 * 64;
  static constexpr int s_a2 = s_b1 + 64;
  static constexpr int s_b2 = s_a2 + 64 * 64;
  static constexpr int s_a3 = s_b2 + 64;               // [256][64]
  static constexpr int s_b3 = s_a3 + 256 * 64;
  static constexpr int s_a4 = s_b3 + 256;              // [3][256]
  static constexpr int s_b4 = s_a4 + 3 * 256;          // [4]
  static constexpr int total = s_b4 + 4;
};

int pack_weights(const motif_decode_t* a, float* wpack, cudaStream_t st);
int decode_layout(int B, int N, int H, int W, int HH, int WW, DecodeScratch* s, char* base, size_t* bytes);
int decode_simt(const motif_decode_t* a, cudaStream_t st);
int decode_tc(const motif_decode_t* a, cudaStream_t st);
int decode_f16(const motif_decode_t* a, cudaStream_t st);
size_t decode_f16_workspace_bytes(int B, int N, int H, int W, int HH, int WW);
int check_decode(const motif_decode_t* a);
size_t tc_image_bytes();
int tc_set_trace(long long* buf, int capacity);
int f16_set_trace(long long* buf, int capacity);
int f16_set_wait_debug(unsigned int* mapped);

}  // namespace motif
