// Modulated deformable convolution (DCNv2) forward, fused.  Replaces the reference's extension
// models/modules/DCNv2 (dcn_v2.py:13-47 -> src/cuda/dcn_v2_im2col_cuda.cu:25-55, 125-195 + SGEMM in dcn_v2_cuda.cu), which
// needs THC and cannot be built on torch 2.x: PCD_Align / the ConvLSTM alignment of the encoder call it twelve times per
// frame pair (Ours.py:53-172) with 3x3 kernels, stride 1, padding 1, 64 -> 64 channels, 8 deformable groups.
//
// The reference materialises the [C_in * 9, H * W] column matrix (132 MB per call at Adobe LR size) and multiplies it with
// cuBLAS.  Here a CTA owns 32 output pixels x 64 output channels and walks the deformable groups: per group it samples
// the group's channels at the 9 displaced taps into shared memory (bilinear setup -- offsets, mask, four corner indices
// and weights -- computed once per (tap, pixel) and shared by the group's channels), stages the matching 64 x (cpg * 9)
// slice of the weight matrix, and accumulates a 2-channel x 4-pixel register tile per thread.  No column matrix exists.
#include "common.cuh"

namespace motif {

constexpr int kDcnPx = 32;        // output pixels per CTA (consecutive in the flattened image)
constexpr int kDcnCo = 64;        // output channels per CTA
constexpr int kDcnThreads = 256;  // (kDcnCo / 2) channel pairs x (kDcnPx / 4) pixel quads
constexpr int kDcnMaxCpg = 8;     // channels per deformable group staged at once (the model: 64 channels / 8 groups)
constexpr int kDcnKMax = kDcnMaxCpg * 9;

__global__ void __launch_bounds__(kDcnThreads, 4) dcn_v2_fwd_kernel(const float* __restrict__ in, const float* __restrict__ offset,
                                                                 const float* __restrict__ mask, const float* __restrict__ weight,
                                                                 const float* __restrict__ bias, float* __restrict__ out, int B, int Cin,
                                                                 int Cout, int H, int W, int dg) {
  __shared__ __align__(16) float s_col[kDcnKMax][kDcnPx + 4];  // [k = ch * 9 + tap][pixel]
  __shared__ __align__(16) float s_w[kDcnKMax][kDcnCo + 2];    // [k][output channel of this CTA]; even pitch: float2 reads stay aligned
  __shared__ int s_idx[4][9 * kDcnPx];                         // corner offsets inside a channel plane (-1: contributes nothing)
  __shared__ float s_cw[4][9 * kDcnPx];                        // corner weights x mask
  const int hw = H * W;
  const int tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * kDcnPx;  // first pixel of the tile in [0, B * hw)
  const int co0 = blockIdx.y * kDcnCo;
  const int cpg = Cin / dg;
  const int K = cpg * 9;
  const bool wvec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(weight) & 15) == 0;  // then (Cin * 9) % 4 == 0 as well
  const int o2 = tid >> 3, pq = tid & 7;  // output channels co0 + 2 o2 (+1), pixels 4 pq .. 4 pq + 3
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  for (int g = 0; g < dg; ++g) {
    __syncthreads();  // the previous group's product is done with s_col / s_w / the setup tables
    // ---- bilinear setup per (tap, pixel): dcn_v2_im2col_cuda.cu:163-186 and :25-55 ----
    for (int i = tid; i < 9 * kDcnPx; i += kDcnThreads) {
      const int tap = i / kDcnPx, px = i - tap * kDcnPx;
      const long long p = p0 + px;
      int idx[4] = {-1, -1, -1, -1};
      float cw[4] = {0.f, 0.f, 0.f, 0.f};
      if (p < (long long)B * hw) {
        const int b = (int)(p / hw), s = (int)(p - (long long)b * hw);
        const int y = s / W, x = s - y * W;
        const float* op = offset + ((size_t)(b * dg + g) * 18 + 2 * tap) * hw + s;
        const float off_h = __ldg(op), off_w = __ldg(op + hw);
        const float m = __ldg(mask + ((size_t)(b * dg + g) * 9 + tap) * hw + s);
        const float h_im = (float)(y - 1 + tap / 3) + off_h, w_im = (float)(x - 1 + tap % 3) + off_w;
        if (h_im > -1.0f && w_im > -1.0f && h_im < (float)H && w_im < (float)W) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h_low = (int)hf, w_low = (int)wf, h_high = h_low + 1, w_high = w_low + 1;
          const float lh = h_im - hf, lw = w_im - wf, hh = 1.0f - lh, hw_ = 1.0f - lw;
          if (h_low >= 0 && w_low >= 0) idx[0] = h_low * W + w_low, cw[0] = hh * hw_ * m;
          if (h_low >= 0 && w_high <= W - 1) idx[1] = h_low * W + w_high, cw[1] = hh * lw * m;
          if (h_high <= H - 1 && w_low >= 0) idx[2] = h_high * W + w_low, cw[2] = lh * hw_ * m;
          if (h_high <= H - 1 && w_high <= W - 1) idx[3] = h_high * W + w_high, cw[3] = lh * lw * m;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) s_idx[c][i] = idx[c], s_cw[c][i] = cw[c];
    }
    // ---- weight slice of this group: weight[o][g * cpg + ch][tap] is K contiguous floats per output channel ----
    // (K = cpg * 9 contiguous floats per output channel: coalesced 16-byte loads when K % 4 == 0 and the rows are aligned --
    // a thread-per-(k, o) loop with o fastest touched one 32-byte sector per element and cost more than the convolution)
    if (wvec) {
      const int k4n = K >> 2;
#pragma unroll
      for (int u = 0; u < (kDcnCo * kDcnKMax / 4 + kDcnThreads - 1) / kDcnThreads; ++u) {
        const int i = tid + u * kDcnThreads;
        if (i >= kDcnCo * k4n) break;
        const int o = i / k4n, k4 = i - o * k4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (co0 + o < Cout) v = __ldg(reinterpret_cast<const float4*>(weight + ((size_t)(co0 + o) * Cin + (size_t)g * cpg) * 9) + k4);
        s_w[4 * k4][o] = v.x, s_w[4 * k4 + 1][o] = v.y, s_w[4 * k4 + 2][o] = v.z, s_w[4 * k4 + 3][o] = v.w;
      }
    } else {
      for (int i = tid; i < kDcnCo * K; i += kDcnThreads) {
        const int o = i / K, k = i - o * K;
        s_w[k][o] = (co0 + o < Cout) ? __ldg(weight + ((size_t)(co0 + o) * Cin + (size_t)g * cpg) * 9 + k) : 0.0f;
      }
    }
    __syncthreads();
    // ---- sample the group's channels: col[ch * 9 + tap][px] = sum_corners in[b, g * cpg + ch, idx] * (corner weight * mask) ----
    // (rolled: unrolling it fully needs 110 registers and loses more to occupancy -- 0.53 ms -- than the batched loads gain)
    for (int i = tid; i < cpg * 9 * kDcnPx; i += kDcnThreads) {
      const int ch = i / (9 * kDcnPx), r = i - ch * (9 * kDcnPx);  // r = tap * kDcnPx + px
      const int tap = r / kDcnPx, px = r - tap * kDcnPx;
      const long long p = p0 + px;
      float v = 0.0f;
      if (p < (long long)B * hw) {
        const int b = (int)(p / hw);
        const float* plane = in + ((size_t)b * Cin + (size_t)g * cpg + ch) * hw;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int id = s_idx[c][r];
          if (id >= 0) v = fmaf(__ldg(plane + id), s_cw[c][r], v);
        }
      }
      s_col[ch * 9 + tap][px] = v;
    }
    __syncthreads();
    // ---- product: 2 output channels x 4 pixels per thread ----
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 c4 = *reinterpret_cast<const float4*>(&s_col[k][4 * pq]);
      const float2 w2 = *reinterpret_cast<const float2*>(&s_w[k][2 * o2]);
      acc[0][0] = fmaf(w2.x, c4.x, acc[0][0]), acc[0][1] = fmaf(w2.x, c4.y, acc[0][1]);
      acc[0][2] = fmaf(w2.x, c4.z, acc[0][2]), acc[0][3] = fmaf(w2.x, c4.w, acc[0][3]);
      acc[1][0] = fmaf(w2.y, c4.x, acc[1][0]), acc[1][1] = fmaf(w2.y, c4.y, acc[1][1]);
      acc[1][2] = fmaf(w2.y, c4.z, acc[1][2]), acc[1][3] = fmaf(w2.y, c4.w, acc[1][3]);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int o = co0 + 2 * o2 + u;
    if (o >= Cout) continue;
    const float bv = bias != nullptr ? __ldg(bias + o) : 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long p = p0 + 4 * pq + q;
      if (p >= (long long)B * hw) continue;
      const int b = (int)(p / hw), s = (int)(p - (long long)b * hw);
      out[((size_t)b * Cout + o) * hw + s] = acc[u][q] + bv;
    }
  }
}

// tcgen05 implicit GEMM for the model's configuration (dcn_v2_tc.cu)
bool dcn_v2_tc_applicable(int Cin, int Cout, int dg);
int dcn_v2_tc_fwd(const float* in, const float* offset, const float* mask, const float* weight, const float* bias, float* out, int B, int H, int W,
                  cudaStream_t st);

}  // namespace motif

using namespace motif;

extern "C" int motif_dcn_v2_fwd(const float* in, const float* offset, const float* mask, const float* weight, const float* bias, float* out,
                                int B, int Cin, int Cout, int H, int W, int deformable_groups, void* stream) {
  MOTIF_REQUIRE(in && offset && mask && weight && out, "dcn_v2: null pointer");
  MOTIF_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && deformable_groups > 0, "dcn_v2: non-positive size");
  MOTIF_REQUIRE(Cin % deformable_groups == 0, "dcn_v2: C_in=%d not divisible by deformable_groups=%d", Cin, deformable_groups);
  MOTIF_REQUIRE(Cin / deformable_groups <= kDcnMaxCpg, "dcn_v2: more than %d channels per deformable group", kDcnMaxCpg);
  MOTIF_REQUIRE((long long)B * H * W < (1LL << 31) && (long long)H * W < (1LL << 30), "dcn_v2: image too large");
  if (dcn_v2_tc_applicable(Cin, Cout, deformable_groups)) return dcn_v2_tc_fwd(in, offset, mask, weight, bias, out, B, H, W, (cudaStream_t)stream);
  dim3 grid(ceil_div((long long)B * H * W, kDcnPx), ceil_div(Cout, kDcnCo));
  ProfScope prof("dcn_v2_fwd_kernel", (cudaStream_t)stream);
  dcn_v2_fwd_kernel<<<grid, kDcnThreads, 0, (cudaStream_t)stream>>>(in, offset, mask, weight, bias, out, B, Cin, Cout, H, W, deformable_groups);
  MOTIF_LAUNCHED("dcn_v2_fwd_kernel");
  return 0;
}
