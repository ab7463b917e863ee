/*
 * motif_b200 -- C ABI of the Blackwell (sm_100a) implementation of MoTIF's per-pixel
 * inference hot path.  Plain pointers and sizes only; no torch types.  Every pointer is
 * a DEVICE pointer to fp32 (or int32 where stated), contiguous in the layout named.
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on success, a
 * positive cudaError_t on a CUDA failure or a negative MOTIF_E_* code on a bad argument;
 * motif_last_error() gives the message (thread-local).  Kernels never allocate: scratch
 * comes from the caller (`workspace`).  Functions are re-entrant and launch only on the
 * stream they are given (the reference launches on torch.cuda.current_stream(),
 * models/softsplat_cp.py:248).
 *
 * The reference has no native ABI: its operators are CUDA-C strings compiled by cupy and
 * launched as  cupy.RawModule(...).get_function(name)(grid, block, args, stream)
 * (models/softsplat_cp.py:215-249, OpticalFlow/correlation.py:286-341).  Each entry point
 * below names the reference interface it replaces (paths relative to the reference).
 */
#ifndef MOTIF_B200_H_
#define MOTIF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOTIF_ABI_VERSION 5

#define MOTIF_E_BADARG (-1)   /* null pointer, non-positive size, unknown mode        */
#define MOTIF_E_WORKSPACE (-2) /* workspace smaller than the *_workspace_bytes() answer */
#define MOTIF_E_UNSUPPORTED (-3)

int motif_abi_version(void);
const char* motif_last_error(void);
/* Number of kernels this library launched since the last motif_reset_launch_count()
 * (process-wide; used by bench.py for its `gpu_launches` claim). */
long long motif_launch_count(void);
void motif_reset_launch_count(void);
/* Strided copy between pinned host memory and the device on a stream (cudaMemcpy2DAsync): height runs of width bytes.  Plumbing of
 * the multi-GPU host-buffer pipeline (motif_b200/clip_stream.py): a rank pulls only the LR rows its destination row band reads. */
int motif_memcpy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, int to_device, void* stream);
/* Optional per-kernel timing: when enabled, the library brackets each of its main kernels with
 * CUDA events on the launch stream.  motif_prof_collect() synchronises the device and sums the
 * recorded durations (ms) and launch counts for the kernel names given; returns the number of
 * recorded launches and clears the record. */
void motif_prof_enable(int on);
int motif_prof_collect(const char* const* names, int n_names, double* out_ms, long long* out_count);

/* ------------------------------------------------------------------------------------
 * Forward splatting.  Replaces kernel_Softsplat_updateOutput + FunctionSoftsplat of
 * models/softsplat_cp.py:12-52, 221-259, 320-347.
 *   in     [n, c, h, w]      flow [n, 2, h, w] (ch0 = x, ch1 = y, destination pixels)
 *   metric [n, 1, h, w]      (LINEAR / SOFTMAX; NULL otherwise)
 *   out    [n, c_out, h, w]  c_out = c for SUMMATION, c + 1 otherwise; the last channel is
 *                            the normaliser (ones / metric / exp(metric) splatted).  The
 *                            output is UN-normalised, as in the reference (:336-346).
 * `out` needs no initialisation.  Non-finite flow: the source pixel is skipped (the
 * reference traps on a device assert, :25-26).
 * motif_splat_fwd is the destination-centric production kernel (deterministic, no float
 * atomics on the common path); it needs motif_splat_workspace_bytes(n, h, w) bytes.
 * motif_splat_fwd_atomic is the reference-order float-atomic scatter, kept as the overflow
 * path of the former and as an on-device cross-check.
 * ---------------------------------------------------------------------------------- */
enum { MOTIF_SPLAT_SUMMATION = 0, MOTIF_SPLAT_AVERAGE = 1, MOTIF_SPLAT_LINEAR = 2, MOTIF_SPLAT_SOFTMAX = 3 };

size_t motif_splat_workspace_bytes(int n, int h, int w);
int motif_splat_fwd(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h, int w,
                    int mode, void* workspace, size_t workspace_bytes, void* stream);
int motif_splat_fwd_atomic(const float* in, const float* flow, const float* metric, float* out, int n, int c, int h,
                           int w, int mode, void* stream);

/* Max splat: out = max(1.0, max over contributions of in*weight).  Replaces
 * models/softsplat_max_cp.py:12-58, 240-278 (output initialised to ONES, :254).
 *   in [n,c,h,w], flow [n,2,h,w], out [n,c,h,w] (needs no initialisation). */
int motif_splat_max_fwd(const float* in, const float* flow, float* out, int n, int c, int h, int w, void* stream);

/* Count splat: number of source pixels whose 2x2 footprint covers each destination pixel.
 * Replaces models/softsplat_count_cp.py:14-52, 117-165 (the wrapper feeds ones; the input
 * values are ignored).   flow [n,2,h,w], out [n,1,h,w] fp32 holding integers. */
int motif_splat_count_fwd(const float* flow, float* out, int n, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------
 * PWC-Net cost volume.  Replaces kernel_Correlation_rearrange + kernel_Correlation_updateOutput
 * + _FunctionCorrelation.forward of OpticalFlow/correlation.py:17-112, 294-348.
 *   first, second [b, c, h, w]   out [b, 81, h, w]
 *   out[b, 9*(dy+4)+(dx+4), y, x] = (1/c) * sum_k first[b,k,y,x] * second[b,k,y+dy,x+dx], zero outside.
 * ---------------------------------------------------------------------------------- */
int motif_corr_fwd(const float* first, const float* second, float* out, int b, int c, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------
 * Reliability maps + flow-encoder input (the step upstream of the decoder).  Replaces
 * models/modules/Ours.py:562-578 (psi_photo / psi_flow / psi_var through BackWarp, :892-923, and the 3x3 gaussian
 * conv3d with reflect padding) and :613-637 (the tensor handed to flow_process; trans=False, input_Z=True).
 *   fr0, fr1 [B, 3, H, W]   the two LR frames           flow [4B, 2, H, W]  LR flows of the pairs 00, 01, 10, 11
 *   g_filter [9]            LunaTokis.g_filter          out  [2B, 14, H, W] per reference r and pair j = 0, 1:
 *                           [flow / 20 (2) | psi_photo, psi_flow / 10, psi_var | durations / 8 (2)]
 * ---------------------------------------------------------------------------------- */
int motif_flow_front(const float* fr0, const float* fr1, const float* flow, const float* g_filter, float* out, int B, int H, int W,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * RAFT correlation lookup of one pyramid level.  Replaces alt_cuda_corr.forward as
 * AlternateCorrBlock.__call__ uses it (models/core/corr.py:69-87; the module is a binary the reference does not ship);
 * the in-repo definition of the quantity is CorrBlock (corr.py:8-56, utils/utils.py:57-70).
 *   fmap1 [B, H, W, C]   fmap2 [B, H2, W2, C] (this level)   coords [B, H, W, 2] = (x, y) in pixels of this level
 *   out [B, (2r+1)^2, H, W]: out[b, a*(2r+1)+c, y, x] = sum over the bilinear corners (zero outside fmap2) of
 *        <fmap1[b,y,x,:], fmap2[b,yy,xx,:]> at (cx + a - r, cy + c - r)   -- NOT divided by sqrt(C) (corr.py:87 does that)
 * ---------------------------------------------------------------------------------- */
int motif_raft_corr_lookup(const float* fmap1, const float* fmap2, const float* coords, float* out, int B, int H, int W, int H2,
                           int W2, int C, int r, void* stream);
/* The whole AlternateCorrBlock.__call__ (models/core/corr.py:69-87) in one launch: fmap2_levels[l] [B, h2[l], w2[l], C] is fmap2
 * average-pooled l times (corr.py:64-67), coords [B, H, W, 2] the level-0 coordinates (divided by 2^l per level as corr.py:80 does),
 * out the stacked [B, n_levels * (2r+1)^2, H, W] tensor (corr.py:85-86), divided by sqrt(C) when normalize != 0 (corr.py:87).
 * fmap2_levels / h2 / w2 are HOST arrays of n_levels <= 4 entries.  r <= 3, C = 128 or 256 (RAFT small / full). */
int motif_raft_corr_lookup_pyramid(const float* fmap1, const float* const* fmap2_levels, const int* h2, const int* w2, int n_levels,
                                   const float* coords, float* out, int B, int H, int W, int C, int r, int normalize, void* stream);

/* ------------------------------------------------------------------------------------
 * Modulated deformable convolution (DCNv2) forward, 3x3 / stride 1 / padding 1 / dilation 1.  Replaces
 * dcn_v2_forward of the reference's extension (models/modules/DCNv2/dcn_v2.py:13-47, src/cuda/dcn_v2_im2col_cuda.cu:25-55,
 * 125-195 + the SGEMM of src/cuda/dcn_v2_cuda.cu), which cannot be built on torch 2.x (THC).
 *   in [B, Cin, H, W]   offset [B, dg*18, H, W] (per group and tap: dh, dw)   mask [B, dg*9, H, W]
 *   weight [Cout, Cin, 3, 3]   bias [Cout] or NULL   out [B, Cout, H, W];   Cin / dg <= 8
 * The model's configuration (Cin = Cout = 64, dg = 8: every call site of Ours.py:53-172) runs as an implicit GEMM on tcgen05 with
 * fp32-equivalent two-piece fp16 operands (csrc/dcn_v2_tc.cu); every other shape on CUDA cores (csrc/dcn_v2.cu).
 * ---------------------------------------------------------------------------------- */
int motif_dcn_v2_fwd(const float* in, const float* offset, const float* mask, const float* weight, const float* bias, float* out,
                     int B, int Cin, int Cout, int H, int W, int deformable_groups, void* stream);

/* ------------------------------------------------------------------------------------
 * Output path of the evaluation loop.  Replaces the eager crop + L1 + luma + MSE of test.py:187-235.
 *   fake [n_frames, 3, hp, wp] decoded frames (possibly padded), cropped to the top-left h x w
 *   real [n_frames, 3, h, w]   ground truth
 *   out  [n_frames, 2] DOUBLE: sum over RGB of |real - fake|, sum over pixels of (Y(real) - Y(fake))^2 with
 *        Y = ((255 R * 65.481 + 255 G * 128.553 + 255 B * 24.966) / 255 + 16) / 255 in fp32 as test.py:212-217 forms it
 * ---------------------------------------------------------------------------------- */
int motif_frame_metrics(const float* fake, const float* real, double* out, int n_frames, int hp, int wp, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------
 * Space-time local implicit decoder.  Replaces models/modules/Ours.py:659-858 (LunaTokis.forward
 * from make_coord to the clamp) with SIREN MLPs of models/modules/SIREN.py:44-45, 76-79.
 * ---------------------------------------------------------------------------------- */

/* One SIREN MLP in the checkpoint (state_dict) layout: weight[l] is [out_l, in_l] row-major,
 * bias[l] is [out_l]; layers 0..n_layers-2 are sine layers (omega = 30), the last is linear. */
typedef struct {
  int n_layers;
  const float* weight[5];
  const float* bias[5];
} motif_siren_t;

typedef struct {
  int B;      /* clips                                  */
  int N;      /* target timestamps per clip             */
  int H, W;   /* LR latent size                         */
  int HH, WW; /* HR output size                         */
  /* 1-D coordinate sequences exactly as make_coord (Ours.py:874-889) builds them on the
   * host in fp32: seq[i] = fl(fl(-1 + 1/n) + fl(fl(2/n) * i)).  Device pointers. */
  const float* seq_hh; /* [HH] */
  const float* seq_ww; /* [WW] */
  const float* seq_h;  /* [H]  */
  const float* seq_w;  /* [W]  */
  float flow_scale;    /* fl32(HH / H), the python-double ratio of Ours.py:794 cast to fp32 */
} motif_geom_t;

/* Nearest-latent index and relative coordinate of every HR query (Ours.py:667-689, 704,
 * 720-722; ATen grid_sample nearest, align_corners=False).  For bit-exactness tests.
 *   iy, ix [HH*WW] int32; coord [HH*WW, 2] shifted+clamped (y,x); rel [HH*WW, 2] (y,x). */
int motif_query_geometry(const motif_geom_t* g, int32_t* iy, int32_t* ix, float* coord, float* rel, void* stream);

/* LR latents NCHW -> pixel-major ("channel-last") [rows, H*W, 64]; rows = leading dim. */
int motif_pack_latents(const float* nchw, float* packed, int rows, int channels, int hw, void* stream);
/* The same for the pixels [p_begin, p_end) of every plane only (the LR rows a destination row band of a sharded decode reads:
 * rows [floor(s0 * H / HH) - 1, ceil(s1 * H / HH) + 1) for the source rows [s0, s1) = the band widened by its halo). */
int motif_pack_latents_range(const float* nchw, float* packed, int rows, int channels, int hw, int p_begin, int p_end, void* stream);

typedef struct {
  motif_geom_t geom;
  /* pixel-major LR latents (motif_pack_latents) */
  const float* feat;      /* [2B, H*W, 64]  F_0^L, F_1^L  leading index r*B+b (Ours.py:609)  */
  const float* flow_feat; /* [2B, H*W, 64]  T_0^L, T_1^L  (Ours.py:638)                    */
  const float* residual;  /* [B,  H*W, 64]  F_01^L        (Ours.py:607)                    */
  const float* target_t;  /* [B, N] HOST pointer                                            */
  motif_siren_t imnet, flow_imnet, synth_net;
  float alpha;            /* LunaTokis.alpha (Ours.py:509)                                  */
  /* outputs */
  float* rgb;      /* [N, B, 3, HH, WW] clamped to [0,1] (Ours.py:858)                      */
  float* flow_out; /* [2*B*N, 2, HH, WW] = flow_hr / 20 / (HH/H) (Ours.py:858); may be NULL  */
  /* scratch: motif_decode_workspace_bytes() bytes */
  void* workspace;
  size_t workspace_bytes;
  /* optional debug taps (may be NULL), NCHW like the reference tensors:
   *   dbg_splat [B*N, 133, HH, WW] = blended splat (130) + extra (3)  (Ours.py:810-836) */
  float* dbg_synth_in; /* [B*N, 198, HH, WW] the synth_net input (Ours.py:839-844); fp32 / tf32x3 only */
  float* dbg_pre0;     /* [B*N, 64, HH, WW] synth_net layer-0 pre-activation (synth_in * W0^T + b0); f16x3 only */
  int n_begin, n_end;  /* timestamps [n_begin, n_end) of each clip are decoded (sharding) */
  int precision;       /* MOTIF_PRECISION_*: arithmetic of the three SIREN MLPs               */
  int local_ensemble;  /* LunaTokis.local_ensemble (Ours.py:453, 660-663, 754-764): 0 as shipped = one nearest latent;
                        * 1 = the four shifted latents blended by diagonally swapped area weights.  Implemented by
                        * MOTIF_PRECISION_F16X3 and MOTIF_PRECISION_FP32 (TF32X3 returns MOTIF_E_UNSUPPORTED). */
  /* Destination row band of a sharded decode (MOTIF_PRECISION_F16X3 only).  row_end == 0: the whole image.
   * Otherwise only the destination rows [row_begin, row_end) of `rgb` are produced (row_begin and row_end multiples of 16, or
   * row_end == HH) and only the sources of rows [row_begin - halo, row_end + halo) are evaluated (`flow_out` is written for
   * those rows only): correct iff no source outside them lands inside the band, i.e. iff max |flow_y| < halo - 1 HR pixels
   * everywhere.  flow_y_max (device, 64 words, may be NULL) receives the float bit patterns whose maximum is the largest
   * |flow_y| over the sources of the band's OWN rows; the maximum over all bands of a clip checks the halo. */
  int row_begin, row_end, halo;
  float* flow_y_max;
  int weights_ready;   /* 1: `workspace` still holds the weight images a previous motif_decode wrote for THESE weights at THIS
                        * precision (same workspace pointer, nothing else wrote to it): the per-call repacking is skipped.
                        * Honoured by MOTIF_PRECISION_F16X3; 0 is always safe.                                          */
  int latents_nchw;    /* 1: feat / flow_feat / residual are the reference's own NCHW tensors [R, 64, H, W] (no motif_pack_latents
                        * pass): MOTIF_PRECISION_F16X3 only, whose per-LR-pixel tables are the only readers of the latents.     */
} motif_decode_t;

/* MLP arithmetic.  F16X3 (default of the Python mirror): tcgen05 kind::f16, every fp32 operand split into two
 * fp16 pieces (22 significant bits) and three products per term, fp32 TMEM accumulation; layer 0 of each MLP
 * evaluated per LR pixel; list-based (atomic-free) splat.  TF32X3: first-generation tensor-core path,
 * error-compensated 3xTF32 with a float-atomic scatter.  FP32: CUDA-core FFMA + sinf, the numerical yardstick. */
enum { MOTIF_PRECISION_TF32X3 = 0, MOTIF_PRECISION_FP32 = 1, MOTIF_PRECISION_F16X3 = 2 };

size_t motif_decode_workspace_bytes(int B, int N, int H, int W, int HH, int WW);
/* sizeof(motif_decode_t) as this library was compiled: lets a foreign-language binding verify its struct layout. */
size_t motif_sizeof_decode_t(void);
/* Whole hot path for one batch of clips from resident LR latents: imnet once per clip, then
 * per timestamp flow_imnet -> 3 splats of both references -> blend -> synth_net -> clamp. */
int motif_decode(const motif_decode_t* args, void* stream);

/* Self-test of the tcgen05 path on one tile: d[128][64] = x[128][64] * w[64][64]^T with `terms` = 1
 * (plain TF32) or 3 (error-compensated 3xTF32).  scratch: >= 32 KiB of device memory. */
/* Tuning aid: install a device buffer of `capacity` (event id, clock64) pairs that CTA 0 of the tensor-core
 * decoder kernels fills (NULL uninstalls).  Not used by the product path. */
int motif_tc_set_trace(long long* buf, int capacity);
/* Tuning aid: issue `reps` back-to-back tcgen05.mma kind::tf32 (M=128, N=n, K=8; A from TMEM or shared
 * memory, round-robin over n_acc independent accumulators) in one CTA; out[0] = cycles to completion,
 * out[1] = cycles spent issuing. */
int motif_tc_mma_rate(long long* out, int n, int reps, int a_in_tmem, int n_acc, void* stream);
int motif_tc_selftest(const float* x, const float* w, float* d, float* scratch, int terms, void* stream);
/* Debug aid: host pointer to a 4 KB device-mapped buffer that expired mbarrier waits of the f16x3 decoder kernels report
 * into before they trap (word 0 = count, then (barrier shared address, parity, block, thread) from word 4 on). */
unsigned int* motif_tc_wait_debug_buffer(void);

#ifdef __cplusplus
}
#endif
#endif /* MOTIF_B200_H_ */
