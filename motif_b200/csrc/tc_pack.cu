#include "tc_pack.cuh"

namespace motif {

__global__ void pack_block_kernel(const float* __restrict__ w, int ldw, int n0, int k0, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    const float v = (k0 + k < ldw) ? w[(size_t)(n0 + n) * ldw + k0 + k] : 0.0f;
    const float hi = tc::tf32_rna(v);
    const float lo = tc::tf32_rna(v - hi);
    const uint32_t off = tc::sw128_offset(64, n, k) >> 2;
    dst[off] = hi;
    dst[(tc::kBlockHalfBytes >> 2) + off] = lo;
  }
}

}  // namespace motif
