/*
 * mra_gan_b200.h -- C ABI of the B200-native 3D-CycleGAN hot-path library (libmra_b200.so).
 *
 * The reference (pedrob37/MRA-GAN) has no FFI: its numerics are torch.nn modules that dispatch to
 * ATen/cuDNN.  Each entry point below replaces one of those library call sites; the citation names
 * the reference line whose arithmetic it reproduces (paths relative to /root/reference).
 *
 * Conventions
 *   - plain C types only; every buffer is BORROWED from the caller (the library never allocates or
 *     frees device memory that outlives a call), all launches go to the `stream` argument, no host
 *     synchronisation inside, re-entrant across streams.
 *   - return value 0 = success, negative = error; mra_last_error() gives the message for the
 *     calling thread.  Nothing throws across the ABI.  Descriptors are validated (sizes, stride,
 *     dtype, consistent output dims) and the mandatory operands of the conv entry points are
 *     checked for NULL on the host BEFORE any CUDA call is made: a rejected call has not touched
 *     the device.  Optional operands (bias, stats, dw / dbias, workspace of size 0) may be NULL.
 *   - activations are channels-last: [N][D][H][W][C] contiguous, element type `dtype`
 *     (MRA_F32 or MRA_BF16); accumulation is always fp32 (statistics fp64).
 *   - conv weights are "packed": [kD*kH*kW][Cout][Cin] (tap-major, Cin contiguous) where Cout/Cin
 *     are the op's output/input channels (for ConvTranspose3d too).  `wT` is the per-tap transpose
 *     [taps][Cin][Cout] used by dgrad.
 *   - a "padded" tensor is simply a tensor whose D/H/W already include a materialised halo.
 */
#ifndef MRA_GAN_B200_H
#define MRA_GAN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mra_stream_t;          /* cudaStream_t */

enum { MRA_F32 = 0, MRA_BF16 = 1 };
enum { MRA_ACT_NONE = 0, MRA_ACT_RELU = 1, MRA_ACT_LRELU = 2, MRA_ACT_TANH = 3, MRA_ACT_SIGMOID = 4 };
enum { MRA_LOSS_L1 = 0, MRA_LOSS_MSE_CONST = 1, MRA_LOSS_BCE_CONST = 2 };
enum {
  MRA_CONV_FORCE_NAIVE = 1,          /* use the CUDA-core kernels even where tcgen05 is eligible */
  MRA_CONV_ACCUMULATE  = 2,          /* wgrad: add into dw/dbias instead of overwriting */
  MRA_CONV_WS_REUSE    = 4           /* the workspace still holds the channel-expanded operand that an earlier call of
                                        this layer on the SAME tensors left at its start (mra_conv3d_lowering() != 0:
                                        fprop -> wgrad for lowerings 1 and 3, wgrad -> dgrad for lowering 2): skip the
                                        expansion pass */
};

/* nn.Conv3d / nn.ConvTranspose3d geometry (cubic kernel, isotropic stride, implicit zero padding).
 * Reference call sites: models/networks3D.py:186,194,205,212,241,256 (generators),
 * :304,312,319,326 (UNet), :392,402,411,417 (PatchGAN). */
typedef struct mra_conv_desc {
  int32_t n, cin, cout;
  int32_t din, hin, win;             /* x spatial dims as stored (incl. any materialised halo) */
  int32_t dout, hout, wout;          /* y spatial dims */
  int32_t k, stride, pad;            /* kernel edge, stride (1|2), implicit zero pad */
  int32_t transposed;                /* 0 = Conv3d, 1 = ConvTranspose3d (output_padding folded into dout) */
  int32_t dtype;                     /* storage type of x / y / dy / dx / w / wT */
  int32_t act;                       /* fprop epilogue activation (after bias) */
  float   slope;                     /* LeakyReLU slope */
  int32_t flags;
} mra_conv_desc;

/* y = act(conv(x, w) + bias); optional per-(n,cout) {sum, sum of squares} of the pre-activation
 * output accumulated into `stats` ([n][cout][2] doubles, zeroed by the call) for the InstanceNorm
 * that follows (models/networks3D.py:188,196,209,242,257,404,413). */
int mra_conv3d_fprop(const mra_conv_desc* d, const void* x, const void* w, const float* bias,
                     void* y, double* stats, void* workspace, size_t workspace_bytes,
                     mra_stream_t stream);
/* dx = conv^T(dy, w): gradient wrt the op's input (ATen convolution_backward, input part). */
int mra_conv3d_dgrad(const mra_conv_desc* d, const void* dy, const void* wT, void* dx,
                     void* workspace, size_t workspace_bytes, mra_stream_t stream);
/* dgrad fused with the STATISTICS PASS of the InstanceNorm backward in front of this conv (models/networks3D.py:188-197,
 * 233-243: conv <- [pad <-] ReLU/LeakyReLU <- InstanceNorm).  y_act = that fused norm's stored output, i.e. this conv's
 * own forward input (same shape as dx); norm_act / norm_slope = its activation (NONE, RELU, LRELU with 0 <= slope <= 1).
 * Besides dx the epilogue accumulates sums[n][cin][2] = {sum dy, sum dy * xhat} (dy = act'(xhat) * fold(dx), fp64) --
 * exactly what mra_inorm_act_pad_bwd_stats computes from (dx, x) in a separate sweep -- so the norm backward is left
 * with its apply pass (mra_inorm_act_pad_bwd_apply): 3 tensor sweeps instead of 5.  Only for layers whose dgrad runs
 * on the tensor-core gather kernels (query first); sums is zeroed by the call. */
int mra_conv3d_dgrad_nstats_supported(const mra_conv_desc* d);
int mra_conv3d_dgrad_nstats(const mra_conv_desc* d, const void* dy, const void* wT, void* dx, const void* y_act,
                            int norm_act, float norm_slope, double* sums, void* workspace, size_t workspace_bytes,
                            mra_stream_t stream);

/* dw[taps][cout][cin] (fp32) and optional dbias[cout] (fp32) (convolution_backward, weight/bias part). */
int mra_conv3d_wgrad(const mra_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                     void* workspace, size_t workspace_bytes, mra_stream_t stream);
/* Bytes of caller-provided scratch the op needs (which: 0 fprop, 1 dgrad, 2 wgrad); 0 for most
 * layers -- only the channel-expanded stem (Cin = 1) / head (Cout = 1) lowerings use any. */
size_t mra_conv3d_workspace_size(const mra_conv_desc* d, int which);
/* 1 if the tcgen05 tensor-core path serves this descriptor's fprop(0)/dgrad(1)/wgrad(2). */
int mra_conv3d_uses_tensor_cores(const mra_conv_desc* d, int which);

/* wT[t][ci][co] = w[t][co][ci] with dtype conversion (src_dtype -> dst_dtype). */
int mra_pack_weight_t(const void* w, int src_dtype, void* wT, int dst_dtype, int taps, int cout,
                      int cin, mra_stream_t stream);
/* dst = convert(src) elementwise. */
int mra_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t numel,
                mra_stream_t stream);

/* nn.InstanceNorm3d(affine=False, track_running_stats=True) (+ in-place ReLU / LeakyReLU(0.2),
 * + residual add, + the nn.ReplicationPad3d that feeds the next conv), fused.
 * Reference: models/networks3D.py:19 and its uses :185-197,205-211,232-263,402-414. */
typedef struct mra_norm_desc {
  int32_t n, c, d, h, w;             /* dims of x (the un-padded conv output) */
  int32_t pad;                       /* replication halo written around y (0 = none) */
  int32_t act;                       /* MRA_ACT_NONE | RELU | LRELU, applied after normalisation */
  float   slope;
  int32_t res_pad;                   /* residual tensor's own halo; -1 = no residual */
  int32_t dtype;
  float   eps, momentum;
  int32_t use_running;               /* 0: statistics from `stats` (training); 1: eval mode, normalise with
                                      * running_mean / running_var; 2: mean / rstd ([n][c]) are INPUTS filled by the
                                      * caller -- batch norm with its affine folded in: rstd' = gamma rstd_b,
                                      * mean' = mean_b - beta / rstd' (models/networks3D.py:15-24, norm='batch') */
} mra_norm_desc;

/* stats[n][c][2] (double) = {sum x, sum x^2} over D*H*W  -- the exact pass, used when the producing
 * conv did not already emit them. */
int mra_inorm_stats(const mra_norm_desc* d, const void* x, double* stats, mra_stream_t stream);
/* y[n][d+2p][h+2p][w+2p][c] = pad(act((x-mean)*rstd) + residual); writes mean/rstd ([n][c] fp32) for
 * the backward pass and EMA-updates running_mean / running_var (nullable) like torch does:
 * rm = (1-m) rm + m mean_n(mu), rv = (1-m) rv + m mean_n(unbiased var). */
int mra_inorm_act_pad_fwd(const mra_norm_desc* d, const void* x, const double* stats,
                          const void* residual, void* y, float* mean, float* rstd,
                          float* running_mean, float* running_var, mra_stream_t stream);
/* dx = rstd (dy - mean(dy) - xhat mean(dy xhat)) with dy = act'(.) * fold_pad(gy); optional
 * dres = fold_pad(gy) written into the interior of a zero-haloed tensor with the residual's shape.
 * `sums` is an [n][c][2] double workspace. */
int mra_inorm_act_pad_bwd(const mra_norm_desc* d, const void* gy, const void* x, const float* mean,
                          const float* rstd, void* dx, void* dres, double* sums,
                          mra_stream_t stream);

/* The two passes of mra_inorm_act_pad_bwd on their own.  Batch norm reduces `sums` over the batch and folds the affine
 * between them: the apply pass computes dx = rstd (dy - sums[.][0]/V - xhat sums[.][1]/V) from whatever the caller
 * left in `sums`. */
int mra_inorm_act_pad_bwd_stats(const mra_norm_desc* d, const void* gy, const void* x, const float* mean,
                                const float* rstd, double* sums, mra_stream_t stream);
int mra_inorm_act_pad_bwd_apply(const mra_norm_desc* d, const void* gy, const void* x, const float* mean,
                                const float* rstd, const double* sums, void* dx, void* dres,
                                mra_stream_t stream);

/* Stand-alone activations (UNet pre-activations :306,308; PatchGAN layer 0 :393; Tanh :213,316;
 * Sigmoid :420).  fwd: y = act(x).  bwd: dx = dy * act'(.) evaluated from the OUTPUT y. */
int mra_act_fwd(const void* x, void* y, int64_t numel, int act, float slope, int dtype,
                mra_stream_t stream);
int mra_act_bwd(const void* dy, const void* y, void* dx, int64_t numel, int act, float slope,
                int dtype, mra_stream_t stream);
/* nn.Dropout(p) in training mode (models/networks3D.py:244-245, 332-333): y = x * keep / (1 - p) with the caller's
 * keep mask (one byte per element, 0 / 1); pass scale = 1 / (1 - p).  The backward is the same call on the gradient. */
int mra_mask_scale(const void* x, const unsigned char* keep, void* y, int64_t numel, float scale,
                   int dtype, mra_stream_t stream);
/* UNet skip connection (models/networks3D.py:339-343 torch.cat([x, model(x)], 1) + the parent's ReLU(True) :319,326):
 * out[pos][0:ca] = act(a[pos][:]), out[pos][ca:ca+cb] = act(b[pos][:]) for `positions` channels-last positions -- both
 * halves written into one buffer, no concat copy.  bwd: da = dout[:, 0:ca] * act'(.), db = dout[:, ca:] * act'(.) with
 * act' taken from the stored output (NONE / RELU / LRELU); da or db may be null. */
int mra_cat2_act_fwd(const void* a, const void* b, void* out, int64_t positions, int ca, int cb, int act, float slope,
                     int dtype, mra_stream_t stream);
int mra_cat2_act_bwd(const void* dout, const void* out, void* da, void* db, int64_t positions, int ca, int cb, int act,
                     float slope, int dtype, mra_stream_t stream);

/* nn.ReplicationPad3d forward/backward on its own (models/networks3D.py:185,211 when the producer
 * is not a norm): y = pad(x); dx = fold_pad(gy). */
int mra_reppad_fwd(const void* x, void* y, int n, int d, int h, int w, int c, int pad, int dtype,
                   mra_stream_t stream);
int mra_reppad_bwd(const void* gy, void* dx, int n, int d, int h, int w, int c, int pad, int dtype,
                   mra_stream_t stream);

/* Loss reductions (models/networks3D.py:130-150 GANLoss = MSE/BCE against a constant label;
 * models/cycle_gan_model.py:104-105 L1).  fwd accumulates the un-normalised sum into *acc (double,
 * zeroed by the caller); bwd writes da = (*gout) * scale * dloss/da. */
int mra_loss_fwd(int kind, const void* a, const void* b, float target, int64_t numel, int dtype,
                 double* acc, mra_stream_t stream);
int mra_loss_bwd(int kind, const void* a, const void* b, float target, int64_t numel, int dtype,
                 const float* gout, float scale, void* da, mra_stream_t stream);
/* Cor_CoeLoss sums (models/networks3D.py:156-166): acc[5] += {Sx, Sy, Sxy, Sxx, Syy}. */
int mra_corr_sums(const void* x, const void* y, int64_t numel, int dtype, double* acc,
                  mra_stream_t stream);

/* torch.optim.Adam step (models/cycle_gan_model.py:107-110,234,240), multi-tensor, fp32 state,
 * optional compute-dtype shadow copy of the updated parameter. */
typedef struct mra_adam_tensor {
  float* p; const float* g; float* m; float* v;
  void*  shadow;                     /* nullable: bf16 copy of p in the same element order */
  int64_t numel;
} mra_adam_tensor;
int mra_adam_multi(const mra_adam_tensor* tensors, int count, float lr, float beta1, float beta2,
                   float eps, int step, mra_stream_t stream);
/* Same sweep with the hyper-parameters read from DEVICE memory: hyper = {lr, beta1, beta2, eps, lr / (1 - beta1^t),
 * sqrt(1 - beta2^t)} (6 floats).  Lets a captured CUDA graph of the training step be replayed while the host advances
 * the step count and the learning-rate schedule. */
int mra_adam_multi_dev(const mra_adam_tensor* tensors, int count, const float* hyper, mra_stream_t stream);
/* Device-resident step counter for mra_adam_multi_dev: state = {lr, beta1, beta2, eps, t, -, -, -} (8 doubles in
 * device memory, written by the caller when a rate changes or a checkpoint is loaded).  One launch does t += 1 and
 * writes hyper[0..5] with the bias corrections of step t computed in double (torch.optim.Adam's host arithmetic).
 * Captured together with the sweep, a replayed graph advances exactly one step per replay, so the host may run any
 * number of steps ahead of the device without a staging buffer to race on. */
int mra_adam_advance(double* state, float* hyper, mra_stream_t stream);

/* Sliding-window inference helpers (test.py:147-178): window extraction with the (x-127.5)/127.5
 * scaling, and label += pred*127.5+127.5 ; weight += 1 accumulation; final label/weight + 0.01. */
int mra_window_extract(const float* vol, int X, int Y, int Z, int i0, int j0, int k0, int px, int py,
                       int pz, void* patch, int dtype, mra_stream_t stream);
int mra_window_accumulate(const void* pred, int dtype, float* label, float* weight, int X, int Y,
                          int Z, int i0, int j0, int k0, int px, int py, int pz, mra_stream_t stream);
int mra_window_finalize(float* label, const float* weight, int64_t numel, mra_stream_t stream);

/* Which channel-expanded lowering (csrc/conv_special.cuh) serves this layer: 0 none, 1 stem (Cin = 1, k x k x k),
 * 2 head (Cout = 1), 3 im2col (Cin = 1, strided).  Callers use it to share one workspace between the calls of a
 * layer (MRA_CONV_WS_REUSE); the workspace must then be max over the calls' mra_conv3d_workspace_size(). */
int mra_conv3d_lowering(const mra_conv_desc* d);

/* Host-only introspection of the implicit-GEMM plan (no GPU needed; used by the CPU tests).
 * Fills `out` (capacity `cap` int32 words) with the launch list and returns the number of words
 * written, or a negative error.  Layout documented in csrc/conv_plan.h. */
int mra_conv_plan_describe(const mra_conv_desc* d, int which, int32_t* out, int cap);

/* Reads (reset != 0: and clears) the device-side error flag the tensor-core kernels raise when a
 * bounded mbarrier wait expires.  Synchronises the device; for tests / debugging only. */
int mra_debug_tc_error(int reset);

/* Reads (reset != 0: and clears) 8 device-side cycle counters that the tensor-core kernels fill when run
 * with the MRA_WGRAD_DEBUG / MRA_GATHER_DEBUG environment bit 1 set (pipeline diagnosis; synchronises). */
int mra_debug_counters(unsigned long long* out, int reset);

/* Host-side walk of the PERSISTENT SCHEDULES of the tensor-core kernels for one op (which: 0 fprop, 1 dgrad, 2 wgrad), run
 * on the CPU with the very functions the kernels execute (halo_decode / halo_stats_key / halo_plane_live for
 * gather_halo_kernel, wseg_begin / wseg_next for wgrad_tc_kernel's stream-K ranges).  No device work: usable without a
 * GPU; tests/test_schedule_cpu.py checks coverage, channel bounds of the statistics flushes and the stream-K partition.
 * units: CTAs (or CTA pairs) of the persistent grid, 0 = what the launch would use (148 SMs without a device).
 * single != 0: gather_halo_kernel without CTA pairs.  Layout of out[] (int32 words):
 *   which 0/1: {1, n_launches} then per halo-capable launch {li, pair, mode, N, Dl, Hl, Wl, Wb, Cn, n_tile, n_tiles,
 *     total_tiles, split_from, total_work, units, skip, kd, nsub, n_records} + n_records x {unit, rank, work, n, d, n0,
 *     width, h0, w0, f0, tb, coff, live_mask, sub, rotation}
 *   which 2:   {2, pair, m_tiles, n_tiles, n_groups, n_items, kblocks, units, cost_lo, cost_hi} + n_groups x {tap0, ntaps,
 *     gpi, item0, n_items} + {n_segments} + n_segments x {unit, item, mt, nt, g, tap0, ntap, kb0, kb1}
 * Returns the number of words written, or a negative error code (cap too small, unsupported geometry). */
int mra_debug_schedule(const mra_conv_desc* d, int which, int units, int single, int32_t* out, int cap);

/* Number of kernels this library has launched in this process (host-side counter). */
long long mra_debug_launch_count(void);

const char* mra_last_error(void);
int mra_version(void);

#ifdef __cplusplus
}
#endif
#endif
