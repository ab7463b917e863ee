// Self-test of the tcgen05 building blocks on one 128 x 64 x 64 tile: D = X * W^T with X staged into TMEM
// (hi/lo split), W as a pre-packed swizzled shared-memory image brought in by one bulk copy, 3xTF32 (or
// 1xTF32) tcgen05.mma with the accumulator in TMEM, read back with tcgen05.ld.  Used by tests/test_tc_gpu.py
// to validate descriptors, the swizzle and the error-compensated product against an fp64 matmul.
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_pack.cuh"

namespace motif {

using namespace tc;

__global__ void __launch_bounds__(192, 1) tc_selftest_kernel(const float* __restrict__ x, const float* __restrict__ wimg, float* __restrict__ d,
                                                            int terms) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sB = reinterpret_cast<float*>(smem);  // 32 KB: hi image then lo image
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBlockImageBytes);
  uint64_t* w_full = bars + 0;
  uint64_t* a_ready = bars + 1;
  uint64_t* d_ready = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_ready, 128);
    mbar_init(d_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t colAhi = 0, colAlo = 64, colD = 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, kBlockImageBytes);
      bulk_g2s(sB, wimg, kBlockImageBytes, w_full);
      mbar_wait(w_full, 0);
      mbar_wait(a_ready, 0);
      tc_fence_after();
      const uint32_t idesc = idesc_tf32(128, 64);
      const uint32_t bhi = smem_u32(sB), blo = smem_u32(sB) + kBlockImageBytes / 2;
      bool acc = false;
      for (int term = 0; term < terms; ++term) {
        // term 0: A_hi*B_hi, term 1: A_lo*B_hi, term 2: A_hi*B_lo
        const uint32_t acol = (term == 1) ? colAlo : colAhi;
        const uint32_t bbase = (term == 2) ? blo : bhi;
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bdesc = smem_desc_sw128(bbase + (ks >> 2) * 8192 + (ks & 3) * 32);
          mma_tf32_ts(tmem + colD, tmem + acol + ks * 8, bdesc, idesc, acc);
          acc = true;
        }
      }
      mma_commit(d_ready);
    }
  } else if (warp >= 2) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = x[row * 64 + c0 + j];
        const float h = tf32_rna(v);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(tf32_rna(v - h));
      }
      tmem_st16(lane_addr + colAhi + c0, hi);
      tmem_st16(lane_addr + colAlo + c0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(a_ready);
    mbar_wait(d_ready, 0);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(lane_addr + colD + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) d[row * 64 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// MMA issue-rate probe: `reps` back-to-back tcgen05.mma (M = 128, N = n, K = 8, A from TMEM or smem) on garbage
// operands; out[0] = cycles from first issue to commit completion, out[1] = cycles spent issuing.
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(long long* out, int n, int reps, int a_in_tmem, int n_acc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = idesc_tf32(128, n);
    const uint32_t b0 = smem_u32(smem), a0 = smem_u32(smem) + 65536;
    uint64_t bdesc[4], adesc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bdesc[j] = smem_desc_sw128(b0 + j * 32);
      adesc[j] = smem_desc_sw128(a0 + j * 32);
    }
    const uint32_t d0 = tmem + 256, d1 = tmem + 256 + ((n_acc > 1) ? n : 0);
    const long long t0 = clock64();
    // 8 MMAs per trip with loop-invariant descriptors; accumulators alternate when n_acc == 2
    for (int i = 0; i < reps; i += 8) {
      const uint32_t accf = i > 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dcol = (j & 1) ? d1 : d0;
        const uint32_t acc = (j < 2) ? accf : 1u;
        if (a_in_tmem) {
          mma_tf32_ts(dcol, tmem + j * 8, bdesc[j & 3], idesc, acc);
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(dcol),
              "l"(adesc[j & 3]), "l"(bdesc[j & 3]), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
    }
    const long long t1 = clock64();
    mma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0;
    out[1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

}  // namespace motif

using namespace motif;

extern "C" int motif_tc_mma_rate(long long* out, int n, int reps, int a_in_tmem, int n_acc, void* stream) {
  MOTIF_REQUIRE(out && n >= 16 && n <= 256 && n % 16 == 0 && reps > 0 && n_acc >= 1 && n_acc <= 2 && n_acc * n <= 256 && reps % 8 == 0, "tc_mma_rate: bad argument");
  const int smem = 2 * 65536 + 1024;
  MOTIF_CUDA(cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out, n, reps, a_in_tmem, n_acc);
  MOTIF_LAUNCHED("tc_mma_rate_kernel");
  return 0;
}

// x [128][64], w [64][64] ([out][in], row-major), d [128][64] = x * w^T; scratch >= 32 KB device memory.
extern "C" int motif_tc_selftest(const float* x, const float* w, float* d, float* scratch, int terms, void* stream) {
  MOTIF_REQUIRE(x && w && d && scratch, "tc_selftest: null pointer");
  MOTIF_REQUIRE(terms == 1 || terms == 3, "tc_selftest: terms must be 1 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  pack_block_kernel<<<1, 256, 0, st>>>(w, 64, 0, 0, scratch);
  MOTIF_LAUNCHED("pack_block_kernel");
  const int smem = tc::kBlockImageBytes + 1024;
  MOTIF_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_selftest_kernel<<<1, 192, smem, st>>>(x, scratch, d, terms);
  MOTIF_LAUNCHED("tc_selftest_kernel");
  return 0;
}
