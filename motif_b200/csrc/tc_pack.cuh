// Weight "block images": a 64(out) x 64(in) sub-matrix of a layer, split into TF32 hi and lo parts
// (round-to-nearest both), each stored exactly as the tcgen05 B operand wants it in shared memory
// (K-major, 128-byte swizzle, two 32-wide K halves), so that one 1-D bulk copy brings a block in.
#pragma once
#include "tc_common.cuh"

namespace motif {
namespace tc {
constexpr int kBlockHalfBytes = 64 * 64 * 4;          // one of hi / lo
constexpr int kBlockImageBytes = 2 * kBlockHalfBytes;  // hi image followed by lo image
}  // namespace tc

// dst[image] <- W[n0 .. n0+64)[k0 .. k0+64) of a row-major [rows][ldw] matrix; out-of-range columns read as 0.
__global__ void pack_block_kernel(const float* __restrict__ w, int ldw, int n0, int k0, float* __restrict__ dst);

}  // namespace motif
