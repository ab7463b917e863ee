// Host prelude that lets the reference's scalar CUDA-C kernel strings compile with g++.
// TEST INFRASTRUCTURE ONLY (oracle/build_ref.py).  Contains no reference code: it only
// supplies the CUDA built-ins those kernels use (blockIdx/threadIdx, atomics on one
// thread at a time, bit casts, __syncthreads through cooperative fibers).
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ucontext.h>

struct motif_dim3 { int x, y, z; };
static motif_dim3 blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, gridDim = {1, 1, 1};

#define __global__
#define __device__
#define __forceinline__ inline
using std::floor;
using std::isfinite;

// One logical thread runs at a time, so the atomics are plain read-modify-writes.
static inline float atomicAdd(float* addr, float v) { float o = *addr; *addr = o + v; return o; }
static inline int atomicMax(int* addr, int v) { int o = *addr; if (v > o) *addr = v; return o; }
static inline unsigned int atomicMin(unsigned int* addr, unsigned int v) { unsigned int o = *addr; if (v < o) *addr = v; return o; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned int __float_as_uint(float f) { unsigned int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned int i) { float f; std::memcpy(&f, &i, 4); return f; }

// dynamic shared memory of the (single) running block
static char motif_dyn_smem[1 << 18];

// ---- cooperative fibers: a block's threads run round-robin, switching at __syncthreads ----
#define MOTIF_MAX_THREADS 64
static ucontext_t motif_main_ctx, motif_ctx[MOTIF_MAX_THREADS];
static char* motif_stack[MOTIF_MAX_THREADS];
static int motif_done[MOTIF_MAX_THREADS];
static int motif_cur = -1;
static void (*motif_body)() = nullptr;

static inline void __syncthreads() {
  if (motif_cur >= 0) swapcontext(&motif_ctx[motif_cur], &motif_main_ctx);
}
static void motif_trampoline() {
  motif_body();
  motif_done[motif_cur] = 1;
}
// Run `body` as `nthreads` fibers of one block; threadIdx.x is set before every resume.
static void motif_run_block(int nthreads, void (*body)()) {
  motif_body = body;
  for (int t = 0; t < nthreads; ++t) {
    if (!motif_stack[t]) motif_stack[t] = (char*)std::malloc(1 << 17);
    getcontext(&motif_ctx[t]);
    motif_ctx[t].uc_stack.ss_sp = motif_stack[t];
    motif_ctx[t].uc_stack.ss_size = 1 << 17;
    motif_ctx[t].uc_link = &motif_main_ctx;
    makecontext(&motif_ctx[t], motif_trampoline, 0);
    motif_done[t] = 0;
  }
  int alive = nthreads;
  while (alive > 0) {
    alive = 0;
    for (int t = 0; t < nthreads; ++t) {
      if (motif_done[t]) continue;
      motif_cur = t;
      threadIdx.x = t;
      swapcontext(&motif_main_ctx, &motif_ctx[t]);
      if (!motif_done[t]) ++alive;
    }
  }
  motif_cur = -1;
}
