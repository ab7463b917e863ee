// Stand-alone tcgen05.mma issue/throughput probe (sm_100a).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_mma tools/probe_mma.cu
// For every (kind, M, N, A source, accumulators, grid) it issues `reps` back-to-back MMAs from one thread per
// CTA on zero operands and prints cycles per MMA (issue-only and until the commit lands) next to the ideal
// tensor-pipe time  M_eff * N * K / (FMA per clock)  with M_eff = max(M, 128).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../motif_b200/csrc/tc_common.cuh"

using namespace motif::tc;

__host__ __device__ constexpr uint32_t idesc_make(int kind_f16, int m, int n) {
  // kind::tf32: A/B format 2; kind::f16: A/B format 0 (f16); D = f32
  return (1u << 4) | ((kind_f16 ? 0u : 2u) << 7) | ((kind_f16 ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int KIND_F16>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND_F16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <int KIND_F16>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND_F16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

template <int KIND_F16, int A_TMEM>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int m, int n, int reps, int n_acc) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    const uint32_t idesc = idesc_make(KIND_F16, m, n);
    const uint32_t b0 = smem_u32(smem), a0 = smem_u32(smem) + 65536;
    uint64_t bdesc[4], adesc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bdesc[j] = smem_desc_sw128(b0 + j * 32);
      adesc[j] = smem_desc_sw128(a0 + j * 32);
    }
    const uint32_t d0 = tmem + 256, d1 = tmem + 256 + ((n_acc > 1) ? n : 0);
    const long long t0 = clock64();
    for (int i = 0; i < reps; i += 8) {
      const uint32_t accf = i > 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t dcol = (j & 1) ? d1 : d0;
        const uint32_t acc = (j < 2) ? accf : 1u;
        if (A_TMEM)
          mma_ts<KIND_F16>(dcol, tmem + j * 8, bdesc[j & 3], idesc, acc);
        else
          mma_ss<KIND_F16>(dcol, adesc[j & 3], bdesc[j & 3], idesc, acc);
      }
    }
    const long long t1 = clock64();
    mma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t2 - t0;
    out[2 * blockIdx.x + 1] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

template <int KIND_F16, int A_TMEM>
void run(long long* d_out, int m, int n, int n_acc, int grid, int reps) {
  const int smem = 2 * 65536 + 2048;
  CK(cudaFuncSetAttribute(rate_kernel<KIND_F16, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long h[2 * 148];
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<KIND_F16, A_TMEM><<<grid, 128, smem>>>(d_out, m, n, reps, n_acc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaMemcpy(h, d_out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
  double tot = 0, iss = 0, mx = 0;
  for (int i = 0; i < grid; ++i) {
    tot += h[2 * i];
    iss += h[2 * i + 1];
    if (h[2 * i] > mx) mx = h[2 * i];
  }
  tot /= grid;
  iss /= grid;
  const int k = KIND_F16 ? 16 : 8;
  const double fma_per_clk = KIND_F16 ? 4096.0 : 2048.0;
  const double ideal = (double)(m < 128 ? 128 : m) * n * k / fma_per_clk;
  printf("%-4s M=%3d N=%3d A=%-4s acc=%d grid=%3d : %7.1f cyc/mma (max CTA %7.1f)  issue %6.1f  ideal %6.1f  util %5.1f%%\n",
         KIND_F16 ? "f16" : "tf32", m, n, A_TMEM ? "tmem" : "smem", n_acc, grid, tot / reps, mx / reps, iss / reps, ideal, 100.0 * ideal * reps / tot);
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 2 * 148));
  const int reps = 512;
  const int ns[] = {32, 64, 128, 256};
  for (int grid : {1, 148}) {
    for (int n : ns) {
      for (int n_acc : {1, 2}) {
        if (n_acc * n > 256) continue;
        run<0, 1>(d_out, 128, n, n_acc, grid, reps);
        run<0, 0>(d_out, 128, n, n_acc, grid, reps);
        run<1, 1>(d_out, 128, n, n_acc, grid, reps);
        run<1, 0>(d_out, 128, n, n_acc, grid, reps);
      }
    }
    for (int n : {64, 256}) {
      run<1, 1>(d_out, 64, n, 1, grid, reps);
      run<1, 0>(d_out, 64, n, 1, grid, reps);
    }
  }
  return 0;
}
