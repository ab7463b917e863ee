// sm_100a building blocks for the tensor-core decoder: mbarrier, 1-D bulk async copy, TMEM allocation,
// tcgen05.ld/st/mma/commit wrappers and the shared-memory operand descriptors.  Hand-written inline PTX;
// bit layouts follow the PTX ISA "tcgen05" chapter (instruction descriptor, matrix descriptor).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace motif {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or a hardware time limit
// elapses, instead of burning issue slots that the working warps of the same scheduler need.  (The hint is an immediate: values
// from 10 ns to 100 us measured alike, and a constant-bank operand cost one load per failed try.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 100000;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking test: has the phase with the given parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spins until the phase with the given parity has completed.  A bounded spin turns a protocol bug into a
// trap (reported as a CUDA error) instead of a hung GPU.  Debug aid: when a host-mapped buffer has been installed
// (motif_tc_set_wait_debug), a thread whose wait expires first records (barrier address, parity, block, thread) there --
// host memory survives the trap -- and keeps waiting a little longer so that the other stuck threads can record too.
// The retry loop is the hot part (a failed try is ~12 % of the instructions the MLP kernels issue): one counter update and
// one branch per failed try; everything else lives in the out-of-line expiry path.
static __device__ unsigned int* g_wait_dbg = nullptr;  // per translation unit
static __device__ __noinline__ void mbar_wait_expired(uint64_t* bar, uint32_t parity) {
  if (g_wait_dbg != nullptr) {
    const unsigned int i = atomicAdd(g_wait_dbg, 1u);
    if (i < 255u) {
      g_wait_dbg[4 + 4 * i] = smem_u32(bar);
      g_wait_dbg[5 + 4 * i] = parity;
      g_wait_dbg[6 + 4 * i] = blockIdx.x;
      g_wait_dbg[7 + 4 * i] = threadIdx.x;
    }
    __threadfence_system();
  }
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 16); ++it)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
// first-generation wait loop (constant-bank hint, two counter tests per failed try): kept selectable, see mbar_wait_sel
static __constant__ uint32_t c_mbar_hint = 100000u;
__device__ __forceinline__ bool mbar_try_wait_c(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(c_mbar_hint)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_v1(uint64_t* bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait_c(bar, parity); ++it) {
    if (it == (1u << 20)) mbar_wait_expired(bar, parity);
    if (it > (1u << 20) + (1u << 16)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t left = 1u << 20;
#pragma unroll 1
  do {
    if (--left == 0u) {
      mbar_wait_expired(bar, parity);
      return;
    }
#if defined(MOTIF_WAIT_BACKOFF) && MOTIF_WAIT_BACKOFF > 0
    __nanosleep(MOTIF_WAIT_BACKOFF);
#endif
  } while (!mbar_try_wait(bar, parity));
}

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier (no tensor map needed) --------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------------------
// Executed by ONE full warp.  The allocated base address (lane 0, first column) is written to *slot (smem).
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 16 consecutive columns: thread i of the warp <-> TMEM lane (quadrant base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Issue (do not wait for) a 16-column load into r[0..15]; pair with tmem_wait_ld() before r is read.
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// All 64 columns of an accumulator block: four loads in flight, one wait.
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  tmem_ld16p(taddr, r);
  tmem_ld16p(taddr + 16, r + 16);
  tmem_ld16p(taddr + 32, r + 32);
  tmem_ld16p(taddr + 48, r + 48);
  tmem_wait_ld();
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- operand descriptors ----------------------------------------------------------------------------------
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major:
//   [4,6) D format = 1 (f32)   [7,10) A format = 2 (tf32)   [10,13) B format = 2 (tf32)
//   [15] A major = 0 (K)  [16] B major = 0 (K)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Shared-memory matrix descriptor for a K-major operand stored with the 128-byte swizzle:
// rows of 128 bytes (32 tf32 along K), 8 rows per 1024-byte swizzle atom, atoms of consecutive 8-row
// groups 1024 bytes apart (stride byte offset).  [0,14) start >> 4, [16,30) leading byte offset >> 4
// (unused for swizzled K-major, 1), [32,46) stride byte offset >> 4, [46,48) version = 1, [61,64) layout = 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Byte offset of element (row, k) of a [rows x 64] tf32 K-major operand block stored as two 32-wide
// K halves, each `rows` x 128 B with the 128-byte swizzle (16-byte chunk index XOR row%8).
__host__ __device__ constexpr uint32_t sw128_offset(int rows, int row, int k) {
  return (uint32_t)((k >> 5) * rows * 128 + (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4);
}

// D[tmem] (+)= A[tmem] * B[smem]^T, M = 128, K = 8 per instruction, one elected thread issues.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// Arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- kind::f16 (fp16 operands, fp32 accumulate; K = 16 per instruction) ----------------------------------------
// Same descriptor fields as idesc_tf32 with A/B format 0 (f16).
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A holds two fp16 per 32-bit column (even k in the low half), 8 columns per K = 16.
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major with the 128-byte swizzle.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// Byte offset of element (row, k) of a [rows x 64] fp16 K-major operand block: one 128-byte row per matrix row
// (all 64 k), 8-row swizzle atoms of 1024 bytes, 16-byte chunk index XOR row%8.
__host__ __device__ constexpr uint32_t sw128_offset_h(int row, int k) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
// Makes shared-memory writes of the generic proxy (st.shared) visible to the async proxy (tcgen05.mma operand reads).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// L2 prefetch of a contiguous global range (bytes: multiple of 16).
__device__ __forceinline__ void prefetch_l2(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

// Round to TF32 (10-bit mantissa), ties away from zero, with two integer ops (finite inputs).
__device__ __forceinline__ float tf32_round(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

}  // namespace tc
}  // namespace motif
