// Shared host/device helpers of libmotif_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/motif_b200.h"

namespace motif {

extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
}

#define MOTIF_REQUIRE(cond, ...)                                      \
  do {                                                                \
    if (!(cond)) return ::motif::fail(MOTIF_E_BADARG, __VA_ARGS__);   \
  } while (0)

#define MOTIF_CUDA(expr)                                              \
  do {                                                                \
    int _rc = ::motif::check_cuda((expr), #expr);                     \
    if (_rc) return _rc;                                              \
  } while (0)

// Count a launch and surface launch-configuration errors immediately.
#define MOTIF_LAUNCHED(name)                                          \
  do {                                                                \
    ::motif::g_launches.fetch_add(1, std::memory_order_relaxed);      \
    int _rc = ::motif::check_cuda(cudaGetLastError(), name);          \
    if (_rc) return _rc;                                              \
  } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Function attributes (dynamic shared memory size, carve-out) belong to the DEVICE the kernel is loaded on: a process
// that drives several GPUs (the reference wraps the model in DataParallel, VideoSR_base_model.py:36) must set them once
// per device, not once per process.  Returns the slot of the current device in a caller-owned `done[64]` table.
inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}

// Optional per-kernel timing (motif_prof_*): CUDA events recorded on the launch stream around a kernel.
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(const char* name, cudaStream_t stream);
  ~ProfScope();
};

// ---------------------------------------------------------------------------------------
// Bilinear forward-splat footprint (models/softsplat_cp.py:23-38; identical in the max and
// count files).  All fp32, no contraction possible (differences then one product).
// ---------------------------------------------------------------------------------------
struct Footprint {
  int x0, y0;          // north-west corner
  float w[4];          // NW, NE, SW, SE
  bool finite;
};

__device__ __forceinline__ Footprint footprint(int x, int y, float flow_x, float flow_y) {
  Footprint f;
  const float fx = __fadd_rn((float)x, flow_x);
  const float fy = __fadd_rn((float)y, flow_y);
  f.finite = isfinite(fx) && isfinite(fy) && fabsf(fx) < 1.0e9f && fabsf(fy) < 1.0e9f;
  const float flx = floorf(fx), fly = floorf(fy);
  f.x0 = f.finite ? (int)flx : -(1 << 30);
  f.y0 = f.finite ? (int)fly : -(1 << 30);
  const float x0f = (float)f.x0, y0f = (float)f.y0;
  const float x1f = (float)(f.x0 + 1), y1f = (float)(f.y0 + 1);
  f.w[0] = __fmul_rn(__fsub_rn(x1f, fx), __fsub_rn(y1f, fy));
  f.w[1] = __fmul_rn(__fsub_rn(fx, x0f), __fsub_rn(y1f, fy));
  f.w[2] = __fmul_rn(__fsub_rn(x1f, fx), __fsub_rn(fy, y0f));
  f.w[3] = __fmul_rn(__fsub_rn(fx, x0f), __fsub_rn(fy, y0f));
  return f;
}

__device__ __forceinline__ bool corner_inside(const Footprint& f, int corner, int w, int h, int& cx, int& cy) {
  cx = f.x0 + (corner & 1);
  cy = f.y0 + (corner >> 1);
  return (cx >= 0) & (cx < w) & (cy >= 0) & (cy < h);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// max of non-negative floats through their (order-preserving) int patterns
__device__ __forceinline__ void red_max_nonneg(float* addr, float a) {
  asm volatile("red.global.max.s32 [%0], %1;" ::"l"(addr), "r"(__float_as_int(a)) : "memory");
}

}  // namespace motif
