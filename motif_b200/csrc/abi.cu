// Error reporting, launch accounting and the decode entry point of libmotif_b200.
#include <stdarg.h>
#include <string.h>

#include "decoder_common.cuh"

namespace motif {

thread_local char g_last_error[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

// ---- per-kernel event timing ----------------------------------------------------------------------
constexpr int kProfMax = 8192;
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_prof_ev[kProfMax][2];
static const char* g_prof_name[kProfMax];
static bool g_prof_created[kProfMax];

ProfScope::ProfScope(const char* name, cudaStream_t stream) : slot(-1), st(stream) {
  if (!g_prof_on || g_prof_n >= kProfMax) return;
  slot = g_prof_n++;
  if (!g_prof_created[slot]) {
    cudaEventCreate(&g_prof_ev[slot][0]);
    cudaEventCreate(&g_prof_ev[slot][1]);
    g_prof_created[slot] = true;
  }
  g_prof_name[slot] = name;
  cudaEventRecord(g_prof_ev[slot][0], st);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof_ev[slot][1], st);
}

}  // namespace motif

using namespace motif;

extern "C" void motif_prof_enable(int on) {
  g_prof_on = on != 0;
  g_prof_n = 0;
}

// Synchronises the device, then sums the recorded intervals per kernel name into `out_ms` / `out_count`
// for the names given (exact string match).  Returns the number of recorded launches.
extern "C" int motif_prof_collect(const char* const* names, int n_names, double* out_ms, long long* out_count) {
  cudaDeviceSynchronize();
  for (int i = 0; i < n_names; ++i) {
    out_ms[i] = 0.0;
    out_count[i] = 0;
  }
  for (int s = 0; s < g_prof_n; ++s) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[s][0], g_prof_ev[s][1]) != cudaSuccess) continue;
    for (int i = 0; i < n_names; ++i)
      if (strcmp(names[i], g_prof_name[s]) == 0) {
        out_ms[i] += ms;
        out_count[i] += 1;
      }
  }
  const int n = g_prof_n;
  g_prof_n = 0;
  return n;
}

extern "C" int motif_abi_version(void) { return MOTIF_ABI_VERSION; }
extern "C" size_t motif_sizeof_decode_t(void) { return sizeof(motif_decode_t); }
extern "C" const char* motif_last_error(void) { return g_last_error; }
extern "C" long long motif_launch_count(void) { return g_launches.load(); }
extern "C" void motif_reset_launch_count(void) { g_launches.store(0); }

// Strided host <-> device copy on a stream (cudaMemcpy2DAsync): `height` runs of `width` bytes, `spitch` / `dpitch` bytes apart.
// The host-buffer pipeline uses it to pull only the LR rows a destination row band reads out of pinned NCHW latents.
extern "C" int motif_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, int to_device,
                                    void* stream) {
  MOTIF_REQUIRE(dst && src && width > 0 && height > 0 && dpitch >= width && spitch >= width, "memcpy2d: bad arguments");
  MOTIF_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  return 0;
}

extern "C" int motif_decode(const motif_decode_t* args, void* stream) {
  if (int rc = check_decode(args)) return rc;
  if (args->local_ensemble != 0 && args->precision == MOTIF_PRECISION_TF32X3)
    return fail(MOTIF_E_UNSUPPORTED, "decode: local_ensemble is implemented by MOTIF_PRECISION_F16X3 and MOTIF_PRECISION_FP32 (precision %d given)", args->precision);
  if (args->latents_nchw != 0 && args->precision != MOTIF_PRECISION_F16X3)
    return fail(MOTIF_E_UNSUPPORTED, "decode: NCHW latents are read by MOTIF_PRECISION_F16X3 only (precision %d given): pack them with motif_pack_latents", args->precision);
  if (args->row_end != 0 && args->precision != MOTIF_PRECISION_F16X3)
    return fail(MOTIF_E_UNSUPPORTED, "decode: destination row bands are implemented by MOTIF_PRECISION_F16X3 only (precision %d given)", args->precision);
  switch (args->precision) {
    case MOTIF_PRECISION_TF32X3: return decode_tc(args, (cudaStream_t)stream);
    case MOTIF_PRECISION_FP32: return decode_simt(args, (cudaStream_t)stream);
    case MOTIF_PRECISION_F16X3: return decode_f16(args, (cudaStream_t)stream);
    default: return fail(MOTIF_E_BADARG, "decode: unknown precision %d", args->precision);
  }
}

extern "C" int motif_tc_set_trace(long long* buf, int capacity) {
  if (int rc = tc_set_trace(buf, capacity)) return rc;
  return f16_set_trace(buf, capacity);
}

// Debug aid: returns a host pointer to a 4 KB buffer mapped into the device (allocated on first use) that expired mbarrier
// waits of the f16x3 decoder kernels report into: word 0 = number of reports, then (barrier shared address, parity, block,
// thread) per report from word 4 on.  Host memory survives the trap that follows an expired wait.
extern "C" unsigned int* motif_tc_wait_debug_buffer(void) {
  static unsigned int* host = nullptr;
  if (host == nullptr) {
    if (cudaHostAlloc((void**)&host, 4096, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(host, 0, 4096);
    unsigned int* dev = nullptr;
    if (cudaHostGetDevicePointer((void**)&dev, host, 0) != cudaSuccess || f16_set_wait_debug(dev) != 0) return nullptr;
  }
  return host;
}
