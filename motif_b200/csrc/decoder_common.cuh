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

// -----------------------------------------------------------------------------------------------------
// Scratch layout (all fp32, pixel-major so that one pixel's channels are contiguous):
//   imf      [2B][qs][64]   imnet output per reference frame (clip-invariant)
//   acc_main [B][qs][128]   sum-splat accumulator of one timestamp: imnet' (64) | nearest feat' (64),
//                           both references accumulate into the same cell (Ours.py:811 sums them anyway)
//   acc_side [B][qs][4]     (sum e*dx, sum e*dy, sum e [the normaliser], count)
//   acc_max  [B][qs]        max splat of exp(z), starts at 1.0
//   wpack                   packed weights (see WeightPack)
// -----------------------------------------------------------------------------------------------------
struct DecodeScratch {
  float* imf;
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

}  // namespace motif
