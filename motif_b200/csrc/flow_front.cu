// Reliability maps and flow-encoder input of LunaTokis.forward in ONE kernel (the step immediately
// upstream of the hot path): models/modules/Ours.py:562-578 (psi_photo, psi_flow, psi_var; BackWarp :892-923; the 3x3
// gaussian conv3d with reflect padding) and :613-637 (the concatenation flow_process consumes; trans=False,
// input_Z=True as shipped).  The reference spends ~40 small LR kernels and a dozen temporaries here; this is LR-sized
// work (4 B H W threads), so the win is launch latency, not bandwidth.
//
// One thread per (frame pair p = 2 r + j, clip b, LR pixel): two bilinear back-warps (grid_sample align_corners=True,
// border padding, coordinates normalised by w -- not w - 1 -- as BackWarp does), three L1 means, the windowed flow
// variance, and the seven output channels of block j of reference r:  [flow / 20 (2) | psi_photo, psi_flow / 10,
// psi_var | duration_0 / 8, duration_1 / 8].
#include "common.cuh"

namespace motif {

// grid_sample(bilinear, align_corners=True, padding_mode='border') source position for destination index `i` moved by
// `d`: BackWarp normalises with (i + d) / n * 2 - 1 (Ours.py:911-912), ATen un-normalises with ((g + 1) / 2) * (n - 1)
// and clips to [0, n - 1].
__device__ __forceinline__ float backwarp_pos(int i, float d, int n) {
  const float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn((float)i, d), (float)n), 2.0f), 1.0f);
  const float u = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), (float)(n - 1));
  return fminf((float)(n - 1), fmaxf(u, 0.0f));
}

struct Bilinear {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1); zero where the corner is outside
};

__device__ __forceinline__ Bilinear bilinear_at(float sx, float sy, int w, int h) {
  Bilinear b;
  const float fx = floorf(sx), fy = floorf(sy);
  b.x0 = (int)fx, b.y0 = (int)fy, b.x1 = b.x0 + 1, b.y1 = b.y0 + 1;
  const float tx = __fsub_rn(sx, fx), ty = __fsub_rn(sy, fy);
  const float ux = __fsub_rn(1.0f, tx), uy = __fsub_rn(1.0f, ty);
  const bool x1in = b.x1 < w, y1in = b.y1 < h;  // x0 / y0 are inside after the clip
  b.w00 = __fmul_rn(ux, uy);
  b.w01 = x1in ? __fmul_rn(tx, uy) : 0.0f;
  b.w10 = y1in ? __fmul_rn(ux, ty) : 0.0f;
  b.w11 = (x1in && y1in) ? __fmul_rn(tx, ty) : 0.0f;
  b.x1 = x1in ? b.x1 : b.x0;
  b.y1 = y1in ? b.y1 : b.y0;
  return b;
}

__device__ __forceinline__ float sample(const float* __restrict__ plane, const Bilinear& b, int w) {
  const float v00 = __ldg(plane + (size_t)b.y0 * w + b.x0), v01 = __ldg(plane + (size_t)b.y0 * w + b.x1);
  const float v10 = __ldg(plane + (size_t)b.y1 * w + b.x0), v11 = __ldg(plane + (size_t)b.y1 * w + b.x1);
  return fmaf(v11, b.w11, fmaf(v10, b.w10, fmaf(v01, b.w01, __fmul_rn(v00, b.w00))));
}

__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256) flow_front_kernel(const float* __restrict__ fr0, const float* __restrict__ fr1,
                                                         const float* __restrict__ flow, const float* __restrict__ gf,
                                                         float* __restrict__ out, int B, int H, int W) {
  const int hw = H * W;
  const long long total = 4LL * B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % hw);
    const int pb = (int)(i / hw);  // p * B + b
    const int p = pb / B, b = pb - p * B;
    const int y = s / W, x = s - y * W;
    const float* fp = flow + (size_t)pb * 2 * hw;
    const float u = __ldg(fp + s), v = __ldg(fp + hw + s);
    const Bilinear bl = bilinear_at(backwarp_pos(x, u, W), backwarp_pos(y, v, H), W, H);

    // psi_photo (Ours.py:567-568): |A - bwarp(B, flow)| averaged over RGB; A = [fr0, fr0, fr1, fr1][p], B = [fr0, fr1, fr0, fr1][p]
    const float* a_img = ((p >> 1) ? fr1 : fr0) + (size_t)b * 3 * hw;
    const float* b_img = ((p & 1) ? fr1 : fr0) + (size_t)b * 3 * hw;
    float photo = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) photo = __fadd_rn(photo, fabsf(__fsub_rn(__ldg(a_img + (size_t)c * hw + s), sample(b_img + (size_t)c * hw, bl, W))));
    photo = __fdiv_rn(photo, 3.0f);

    // psi_flow (Ours.py:570-576): |flow_p - bwarp(-flow_p', flow_p)|, p' = p with the two frames exchanged
    const int pq = ((p & 1) << 1) | (p >> 1);
    const float* fq = flow + (size_t)(pq * B + b) * 2 * hw;
    const float d0 = fabsf(__fsub_rn(u, -sample(fq, bl, W))), d1 = fabsf(__fsub_rn(v, -sample(fq + hw, bl, W)));
    const float pflow = __fdiv_rn(__fdiv_rn(__fadd_rn(d0, d1), 2.0f), 10.0f);

    // psi_var (Ours.py:577-582): sqrt(max(G * f^2 - (G * f)^2, 1e-9)) averaged over the two flow components
    float m[2] = {0.0f, 0.0f}, q[2] = {0.0f, 0.0f};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = reflect(y + dy, H);
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = reflect(x + dx, W);
        const float g = __ldg(gf + (dy + 1) * 3 + (dx + 1));
        const float f0 = __ldg(fp + (size_t)yy * W + xx), f1 = __ldg(fp + hw + (size_t)yy * W + xx);
        m[0] = fmaf(g, f0, m[0]);
        m[1] = fmaf(g, f1, m[1]);
        q[0] = fmaf(g, __fmul_rn(f0, f0), q[0]);
        q[1] = fmaf(g, __fmul_rn(f1, f1), q[1]);
      }
    }
    const float s0 = sqrtf(fmaxf(__fsub_rn(q[0], __fmul_rn(m[0], m[0])), 1e-9f));
    const float s1 = sqrtf(fmaxf(__fsub_rn(q[1], __fmul_rn(m[1], m[1])), 1e-9f));
    const float pvar = __fdiv_rn(__fadd_rn(s0, s1), 2.0f);

    // Ours.py:613-631: block j = p & 1 of reference r = p >> 1; durations [[0,0],[0,8],[8,0],[8,8]].reshape(2,4) / 8
    const int r = p >> 1, j = p & 1;
    float* o = out + ((size_t)(r * B + b) * 14 + 7 * j) * hw + s;
    o[0] = __fdiv_rn(u, 20.0f);
    o[hw] = __fdiv_rn(v, 20.0f);
    o[2 * (size_t)hw] = photo;
    o[3 * (size_t)hw] = pflow;
    o[4 * (size_t)hw] = pvar;
    o[5 * (size_t)hw] = (float)r;                 // r = 0: (0, 0 | 0, 1)   r = 1: (1, 0 | 1, 1)
    o[6 * (size_t)hw] = (float)j;
  }
}

}  // namespace motif

using namespace motif;

extern "C" int motif_flow_front(const float* fr0, const float* fr1, const float* flow, const float* g_filter, float* out, int B, int H, int W,
                                void* stream) {
  MOTIF_REQUIRE(fr0 && fr1 && flow && g_filter && out, "flow_front: null pointer");
  MOTIF_REQUIRE(B > 0 && H >= 2 && W >= 2, "flow_front: bad size B=%d H=%d W=%d (reflect padding needs H, W >= 2)", B, H, W);
  MOTIF_REQUIRE(4LL * B * H * W < (1LL << 31), "flow_front: too large");
  const long long total = 4LL * B * H * W;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  ProfScope prof("flow_front_kernel", (cudaStream_t)stream);
  flow_front_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(fr0, fr1, flow, g_filter, out, B, H, W);
  MOTIF_LAUNCHED("flow_front_kernel");
  return 0;
}
