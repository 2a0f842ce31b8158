/*
 * npml_conv.h -- C ABI of libnpmlconv.so, the B200 (sm_100a) implementation of the
 * numpy-ml convolution hot path.
 *
 * The reference (ddbourgin/numpy-ml) is pure Python and has no FFI of its own; the drop-in
 * surface is the Python layer API (numpy_ml/neural_nets/{utils/utils.py,layers/layers.py}).
 * Every entry point below therefore names the reference function whose arithmetic it
 * replaces (file:line relative to numpy_ml/neural_nets/).  The Python host layer under
 * numpy-ml_b200/ resolves all geometry ('same' pads, output dims, error cases) exactly as
 * the reference does and passes plain ints here; INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - return 0 on success, non-zero on error; npml_last_error() returns a thread-local message.
 *   - every pointer is a DEVICE pointer owned by the caller (inputs, outputs, workspace).  The
 *     library never allocates or frees device memory on these calls and never synchronises:
 *     all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - activations are NHWC (n_ex, rows, cols, ch); weights are (fr, fc, in_ch, out_ch) fp32;
 *     bias / dW / db are fp32.  The element type of activation tensors is selected by
 *     desc.mode: NPML_MODE_FP32 and NPML_MODE_TF32 -> float, NPML_MODE_BF16 -> bfloat16.
 *   - "dil" is the tap spacing D = dilation + 1 (numpy-ml's `dilation` counts inserted zeros,
 *     utils/utils.py:95).
 *   - alignment: activation tensors and the workspace 16 / 128 bytes (the tensor-core paths report an error
 *     otherwise); outputs that are also 32-byte aligned let the epilogues use 32-byte stores.
 */
#ifndef NPML_CONV_H_
#define NPML_CONV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NPML_MODE_FP32 = 0, NPML_MODE_TF32 = 1, NPML_MODE_BF16 = 2 };

/* activation ids (activations/activations.py); p0/p1 = alpha, or slope/intercept for AFFINE */
enum {
  NPML_ACT_IDENTITY = 0, NPML_ACT_AFFINE = 1, NPML_ACT_RELU = 2, NPML_ACT_LEAKY_RELU = 3,
  NPML_ACT_TANH = 4, NPML_ACT_SIGMOID = 5, NPML_ACT_ELU = 6, NPML_ACT_SELU = 7,
  NPML_ACT_HARD_SIGMOID = 8, NPML_ACT_SOFTPLUS = 9, NPML_ACT_GELU = 10, NPML_ACT_EXPONENTIAL = 11
};

/* element types for the dtype-explicit elementwise entry points */
enum { NPML_F32 = 0, NPML_BF16 = 1, NPML_F64 = 2 };

/* operation ids for npml_workspace_bytes */
enum {
  NPML_OP_CONV_FPROP = 0, NPML_OP_CONV_DGRAD = 1, NPML_OP_CONV_WGRAD = 2,
  NPML_OP_DECONV_FPROP = 3, NPML_OP_DECONV_DGRAD = 4, NPML_OP_DECONV_WGRAD = 5,
  NPML_OP_DZ_PREP = 6, NPML_OP_BN = 7, NPML_OP_CONV_BWD = 8
};

/*
 * Geometry of one Conv2D / Deconv2D layer call.  For both layer kinds (N,H,W,C) is the layer
 * INPUT volume, K the layer's out_ch and (OH,OW) the layer OUTPUT rows/cols:
 *   Conv2D  : OH = (H + pr1 + pr2 - (fr*dil - dil + 1)) / stride + 1    (utils/utils.py:653-657)
 *   Deconv2D: OH = stride*(H-1) + fr - pr1 - pr2, dil must be 1          (utils/utils.py:780-786)
 */
typedef struct npml_conv_desc {
  int32_t N, H, W, C;          /* input volume, NHWC                                   */
  int32_t K;                   /* out_ch                                               */
  int32_t fr, fc;              /* kernel rows / cols                                   */
  int32_t stride;
  int32_t dil;                 /* tap spacing D = dilation + 1                         */
  int32_t pr1, pr2, pc1, pc2;  /* (top, bottom, left, right) zero padding              */
  int32_t OH, OW;              /* output rows / cols                                   */
  int32_t mode;                /* NPML_MODE_*                                          */
  int32_t act;                 /* NPML_ACT_* applied in the forward epilogue           */
  float act_p0, act_p1;
  int32_t accumulate;          /* wgrad: dW += result (layers/layers.py:3087-3089)      */
  int32_t flags;               /* NPML_FLAG_*                                           */
} npml_conv_desc;

/* fprop / dgrad: the workspace handed to this call ALREADY holds the operand-type copy of the weights that an
 * earlier call with the same (desc, op) and the same W wrote there -- skip the repack kernel.  The host layer keeps one
 * such workspace per (layer, op) and drops it whenever W changes (assignment, optimizer step), so weights are
 * repacked once per update(), not once per call.  Never set it on a workspace the library has not filled. */
enum { NPML_FLAG_WEIGHTS_PACKED = 1 };

/* ---- Conv2D (layers/layers.py:2895-3178) --------------------------------------------------
 * fprop replaces pad2D + im2col + W_col @ X_col + b + act_fn  (utils/utils.py:603-665,
 * layers/layers.py:3037-3038).  Z (pre-activation) and/or Y (= act(Z)) may be NULL. */
int npml_conv2d_fprop(const npml_conv_desc* d, const void* X, const float* W, const float* b,
                      void* Z, void* Y, void* ws, size_t ws_bytes, void* stream);
/* fprop that also leaves the statistics of its OUTPUT for the training-mode BatchNorm2D behind it (the
 * "optional per-channel sum / sum of squares for BN" of npml_conv2d_fprop): bn_sums[2*K] (double, device) = per-channel
 * [sum y, sum y^2] over (N, OH, OW) of y = act(conv + b) (of Z when Y is NULL), i.e. what layers/layers.py:1146-1147
 * reduces.  One call: the convolution kernel, then a fixed-order column pass over the tensor it has just written (still
 * in L2 for the benchmark shapes) -- forming the sums inside the tcgen05 epilogue was measured and is slower, see
 * DESIGN.md.  Feed them to npml_bn2d_apply / npml_bn2d_finalize (all-reduced first under sync-BN). */
int npml_conv2d_fprop_stats(const npml_conv_desc* d, const void* X, const float* W, const float* b, void* Z, void* Y,
                            double* bn_sums, void* ws, size_t ws_bytes, void* stream);
/* fprop with a FROZEN BatchNorm2D folded into the same epilogue (north_star: "the BatchNorm2D affine [is a] fused
 * epilogue"; modules/modules.py:905-937 runs conv -> BatchNorm2D back to back, and with retain_derived=False or a frozen
 * layer BatchNorm2D.forward is the per-channel affine of its running statistics, layers/layers.py:1141-1156):
 *   Y = bn_scale[k] * act_fn(conv(X, W) + b) + bn_shift[k]        (Z, if given, still receives conv(X, W) + b)
 * bn_scale / bn_shift: fp32 device vectors of K entries, produced on the device by npml_bn2d_fold.  Y is required. */
int npml_conv2d_fprop_bn(const npml_conv_desc* d, const void* X, const float* W, const float* b,
                         const float* bn_scale, const float* bn_shift, void* Z, void* Y, void* ws, size_t ws_bytes,
                         void* stream);
/* dgrad replaces W_col @ dLdZ_col + col2im / np.add.at  (layers/layers.py:3113-3114,
 * utils/utils.py:546-595); gather form, no atomics.  dZ already holds dLdy * act'(Z). */
int npml_conv2d_dgrad(const npml_conv_desc* d, const void* dZ, const float* W, void* dX,
                      void* ws, size_t ws_bytes, void* stream);
/* wgrad replaces im2col + dLdZ_col @ X_col.T  (layers/layers.py:3106-3110); deterministic
 * fixed-order split reduction.  dW is (fr, fc, C, K) fp32. */
int npml_conv2d_wgrad(const npml_conv_desc* d, const void* X, const void* dZ, float* dW,
                      void* ws, size_t ws_bytes, void* stream);
/* wgrad + db in one call: dW as above and db[k] (+)= sum_{n,oh,ow} dZ[n,oh,ow,k] (layers/layers.py:3109);
 * db (fp32, K entries) may be NULL.  On the slab path db is one more accumulator row of the same
 * tensor-core pass (no extra read of dZ); desc.accumulate applies to both outputs.  Workspace:
 * npml_workspace_bytes(d, NPML_OP_CONV_WGRAD). */
int npml_conv2d_wgrad_db(const npml_conv_desc* d, const void* X, const void* dZ, float* dW, float* db,
                         void* ws, size_t ws_bytes, void* stream);

/* The whole of Conv2D._bwd in one call (layers/layers.py:3093-3116): dW (+)= ..., db (+)= ... (desc.accumulate) and
 * dX = ... from dZ = dLdy * act'(Z).  For thin stride-1 "same" shapes (in_ch <= 4, out_ch <= 64: cfg1) it is ONE kernel
 * launch that reads dZ once (csrc/thin_bwd.cu; npml_conv2d_bwd_is_fused says whether a shape takes it); every other shape
 * runs npml_conv2d_wgrad_db followed by npml_conv2d_dgrad.  db may be NULL.  Workspace:
 * npml_workspace_bytes(d, NPML_OP_CONV_BWD). */
int npml_conv2d_bwd(const npml_conv_desc* d, const void* X, const void* dZ, const float* W, float* dW, float* db,
                    void* dX, void* ws, size_t ws_bytes, void* stream);
int npml_conv2d_bwd_is_fused(const npml_conv_desc* d);

/* ---- Deconv2D (layers/layers.py:3353-3568, utils/utils.py:720-794) ------------------------
 * fprop: Z[n, iy*s + r - pr1, ix*s + t - pc1, k] += X[n,iy,ix,c] * W[r,t,c,k]  (no zero-stuffing)
 * dgrad: dX[n,iy,ix,c] = sum dZ[n, iy*s + r - pr1, ix*s + t - pc1, k] * W[r,t,c,k]
 * wgrad: dW[r,t,c,k]  = sum X[n,iy,ix,c] * dZ[n, iy*s + r - pr1, ix*s + t - pc1, k] */
int npml_deconv2d_fprop(const npml_conv_desc* d, const void* X, const float* W, const float* b,
                        void* Z, void* Y, void* ws, size_t ws_bytes, void* stream);
int npml_deconv2d_fprop_bn(const npml_conv_desc* d, const void* X, const float* W, const float* b,
                           const float* bn_scale, const float* bn_shift, void* Z, void* Y, void* ws, size_t ws_bytes,
                           void* stream);
int npml_deconv2d_dgrad(const npml_conv_desc* d, const void* dZ, const float* W, void* dX,
                        void* ws, size_t ws_bytes, void* stream);
int npml_deconv2d_wgrad(const npml_conv_desc* d, const void* X, const void* dZ, float* dW,
                        void* ws, size_t ws_bytes, void* stream);

size_t npml_workspace_bytes(const npml_conv_desc* d, int op);

/* ---- backward prologue: dZ = dLdy * act'(Z) and db = sum_{n,oh,ow} dZ --------------------
 * (layers/layers.py:3103, 3109; Add._bwd layers/layers.py:739-742 when db == NULL).
 * rows = N*OH*OW, ch = K.  Z may be NULL for NPML_ACT_IDENTITY.  dZ may alias dY.
 * db (fp32, ch entries) may be NULL; accumulate != 0 adds into db. */
int npml_dz_prep(int64_t rows, int32_t ch, int32_t act, float p0, float p1, const void* dY,
                 int32_t dy_dtype, const void* Z, int32_t z_dtype, void* dZ, int32_t dz_dtype,
                 float* db, int32_t accumulate, void* ws, size_t ws_bytes, void* stream);

/* ---- elementwise helpers ------------------------------------------------------------------ */
/* dst = (dst_dtype) src, n elements; used at the NumPy(float64) <-> device boundary. */
int npml_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n,
              void* stream);
/* Add.forward (layers/layers.py:690-711): sum = x0 + ... + x_{n_in-1}; Y = act(sum).
 * `sum` may be NULL when act is identity. */
int npml_add_act_fwd(int32_t n_in, const void* const* xs, int32_t dtype, int64_t n, int32_t act,
                     float p0, float p1, void* sum, void* Y, void* stream);
/* Multiply.forward / _bwd (layers/layers.py:800-856): prod = x0 * ... * x_{n_in-1}, Y = act(prod);
 * grads[i] = dY * act'(prod) * prod_{j != i} x_j.  `prod` may be NULL in the forward when act is identity (then Y is
 * the product) and in the backward when act is identity. */
int npml_mul_act_fwd(int32_t n_in, const void* const* xs, int32_t dtype, int64_t n, int32_t act,
                     float p0, float p1, void* prod, void* Y, void* stream);
int npml_mul_act_bwd(int32_t n_in, const void* const* xs, void* const* grads, int32_t dtype, int64_t n,
                     int32_t act, float p0, float p1, const void* dY, const void* prod, void* stream);
/* activation fn on its own (activations/activations.py `fn`): Y = act(Z) */
int npml_act_fwd(const void* Z, void* Y, int32_t dtype, int64_t n, int32_t act, float p0, float p1,
                 void* stream);

/* ---- BatchNorm2D (layers/layers.py:969-1215) ---------------------------------------------
 * fwd, use_batch_stats != 0: per-channel mean and BIASED variance over rows = N*H*W;
 *   running = momentum*running + (1-momentum)*batch (layers/layers.py:1146-1150);
 *   y = scaler*(x-mean)/sqrt(var+eps) + intercept.  save_mean/save_rstd (fp32, ch) receive the
 *   statistics used.  use_batch_stats == 0 normalises with running_mean/var and leaves them.
 * bwd (layers/layers.py:1192-1215): d_scaler/d_intercept accumulate (+=) when accumulate != 0.
 * stats buffers, scaler, intercept, gradients are fp32; x/y/dy/dx are `dtype`. */
int npml_bn2d_fwd(int64_t rows, int32_t ch, const void* x, void* y, int32_t dtype,
                  const float* scaler, const float* intercept, float* running_mean,
                  float* running_var, float momentum, float eps, int32_t use_batch_stats,
                  float* save_mean, float* save_rstd, void* ws, size_t ws_bytes, void* stream);
/* frozen BatchNorm2D as a per-channel affine (layers/layers.py:1141-1156 with the running statistics):
 *   scale = scaler / sqrt(running_var + eps), shift = intercept - running_mean * scale
 * for npml_conv2d_fprop_bn / npml_deconv2d_fprop_bn; everything stays on the device. */
int npml_bn2d_fold(int32_t ch, const float* scaler, const float* intercept, const float* running_mean,
                   const float* running_var, float eps, float* scale, float* shift, void* stream);
int npml_bn2d_bwd(int64_t rows, int32_t ch, const void* dy, const void* x, void* dx, int32_t dtype,
                  const float* scaler, const float* save_mean, const float* save_rstd,
                  float* d_scaler, float* d_intercept, int32_t accumulate, void* ws,
                  size_t ws_bytes, void* stream);

/* sync-BN under batch sharding (new): the two passes above split at the point where the per-channel
 * sums cross ranks.  stats: sums[0..ch) = sum x, sums[ch..2ch) = sum x^2 over this rank's rows (double, device).
 * The host all-reduces `sums` (npml_allreduce_sum_f64) and calls apply with rows_total = rows over all ranks.
 * bwd_stats: sums = [sum dy, sum dy*nhat]; bwd_apply accumulates THIS rank's sums into d_intercept / d_scaler
 * (the gradient bucket is all-reduced later) and uses the all-rank sums for dX. */
int npml_bn2d_stats(int64_t rows, int32_t ch, const void* x, int32_t dtype, double* sums, void* ws,
                    size_t ws_bytes, void* stream);
int npml_bn2d_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* x, void* y, int32_t dtype,
                    const float* scaler, const float* intercept, float* running_mean, float* running_var,
                    float momentum, float eps, const double* sums, float* save_mean, float* save_rstd,
                    void* ws, size_t ws_bytes, void* stream);
int npml_bn2d_bwd_stats(int64_t rows, int32_t ch, const void* dy, const void* x, int32_t dtype,
                        const float* save_mean, const float* save_rstd, double* sums, void* ws,
                        size_t ws_bytes, void* stream);
int npml_bn2d_bwd_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* dy, const void* x, void* dx,
                        int32_t dtype, const float* scaler, const float* save_mean, const float* save_rstd,
                        const double* sums_local, const double* sums_global, float* d_scaler,
                        float* d_intercept, int32_t accumulate, void* ws, size_t ws_bytes, void* stream);

/* ---- fused passes of the skip-connection modules (modules/modules.py:905-984) ------------------------------------
 * The residual tail Y = act(BatchNorm2D_a(xa) + BatchNorm2D_b(xb)) (modules.py:931-937: batchnorm2, batchnorm_skip,
 * add3; layers.py:1146-1156, 705-711) and its backward (modules.py:941-953; layers.py:739-742, 1192-1215) read and
 * write every tensor ONCE instead of once per layer.  All tensors are (rows, ch) views of NHWC volumes in `dtype`.
 *
 * npml_bn2d_finalize: (mean, rstd) and the running-statistics update from per-channel sums = [sum x, sum x^2]
 *   (npml_bn2d_stats, all-reduced by the host under sync-BN); nothing is normalised here.
 * npml_bn2d_pair_add_act_fwd: sum = BN_a(xa) + BN_b(xb), Y = act(sum) in one pass.  scaler_b == NULL: xb is added as
 *   it is (SkipConnectionIdentityModule, modules.py:560-590).  `sum` and/or `Y` may be NULL.
 * npml_bn2d_pair_bwd_stats: g = dY * act'(s) is formed on the fly (s = the stored sum; for ReLU the output Y works
 *   as well, for identity s is ignored); sums[3*ch] = [sum g, sum g*nhat_a, sum g*nhat_b].  mean_b == NULL: no second
 *   BatchNorm2D (third block zero).
 * npml_bn2d_pair_bwd_apply: dxa = BN_a'(g), dxb = BN_b'(g) (dxb = g without a second BatchNorm2D, may be NULL then);
 *   dScaler / dIntercept of both accumulate `sums_local`, dX uses `sums_global` (pass the same buffer twice without
 *   sync-BN) and rows_total.
 * npml_bn2d_bwd_act: BatchNorm2D backward whose dX pass also applies the activation derivative of the layer in FRONT
 *   of it (x = act(z)): dz = BN'(dy) * act'(z) -- the prologue of that layer's own backward (layers.py:3103) without a
 *   pass of its own.  zm = z, or NULL for ReLU (x > 0 <=> z > 0) / identity.  sums_local/global: NULL (statistics are
 *   computed here) or the split sync-BN form of npml_bn2d_bwd_stats. */
int npml_bn2d_finalize(int64_t rows_total, int32_t ch, const double* sums, float* running_mean, float* running_var,
                       float momentum, float eps, float* save_mean, float* save_rstd, void* stream);
int npml_bn2d_pair_add_act_fwd(int64_t rows, int32_t ch, const void* xa, const void* xb, int32_t dtype,
                               const float* scaler_a, const float* intercept_a, const float* mean_a, const float* rstd_a,
                               const float* scaler_b, const float* intercept_b, const float* mean_b, const float* rstd_b,
                               int32_t act, float p0, float p1, void* sum, void* Y, void* stream);
int npml_bn2d_pair_bwd_stats(int64_t rows, int32_t ch, const void* dY, const void* s, int32_t act, float p0, float p1,
                             const void* xa, const void* xb, int32_t dtype, const float* mean_a, const float* rstd_a,
                             const float* mean_b, const float* rstd_b, double* sums, void* ws, size_t ws_bytes,
                             void* stream);
int npml_bn2d_pair_bwd_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* dY, const void* s, int32_t act,
                             float p0, float p1, const void* xa, const void* xb, int32_t dtype, const float* scaler_a,
                             const float* mean_a, const float* rstd_a, const float* scaler_b, const float* mean_b,
                             const float* rstd_b, const double* sums_local, const double* sums_global, void* dxa,
                             void* dxb, float* d_scaler_a, float* d_intercept_a, float* d_scaler_b, float* d_intercept_b,
                             int32_t accumulate, void* ws, size_t ws_bytes, void* stream);
int npml_bn2d_bwd_act(int64_t rows, int64_t rows_total, int32_t ch, const void* dy, const void* x, void* dz,
                      int32_t dtype, const float* scaler, const float* save_mean, const float* save_rstd,
                      const double* sums_local, const double* sums_global, float* d_scaler, float* d_intercept,
                      int32_t accumulate, int32_t act, float p0, float p1, const void* zm, void* ws, size_t ws_bytes,
                      void* stream);

/* ---- SGD step on the (all-reduced) gradient (optimizers/optimizers.py:110-148) ----------------
 * g' = grad * grad_scale (grad_scale carries the clip-norm factor t/||g||, 1 when not clipping);
 * u = momentum*velocity + lr*g'; velocity = u; param -= u.  All fp32, n elements. */
int npml_sgd_update(float* param, const float* grad, float* velocity, int64_t n, float lr, float momentum,
                    float grad_scale, void* stream);
/* All four reference optimizers as one elementwise step on the (all-reduced) gradient
 * (optimizers/optimizers.py:133-148 SGD, 243-259 AdaGrad, 345-361 RMSProp, 450-482 Adam); g' = grad*grad_scale.
 *   SGD     : state0 = velocity, a = momentum                      u = a*state0 + lr*g'; state0 = u; p -= u
 *   AdaGrad : state0 = cache                                       state0 += g'^2; p -= lr*g'/(sqrt(state0)+eps)
 *   RMSProp : state0 = cache, a = decay                            state0 = a*state0+(1-a)*g'^2; p -= lr*g'/(sqrt(state0)+eps)
 *   Adam    : state0 = mean, state1 = var, a/b = decay1/decay2, c1 = 1-decay1^t, c2 = 1-decay2^t (host computes them)
 *             p -= lr*(mean/c1)/(sqrt(var/c2)+eps)
 * All buffers fp32, n elements; the scalars are doubles so that 1 - decay is formed in double. */
enum { NPML_OPT_SGD = 0, NPML_OPT_ADAGRAD = 1, NPML_OPT_RMSPROP = 2, NPML_OPT_ADAM = 3 };
int npml_optimizer_update(int32_t kind, float* param, const float* grad, float* state0, float* state1, int64_t n,
                          double lr, double a, double b, double eps, double c1, double c2, double grad_scale,
                          void* stream);
/* The same step with the clip-norm factor formed ON THE DEVICE: grad_sumsq points at ||grad||^2 (double, device, from
 * npml_sumsq_f32 on the same stream); g' = grad * (clip_norm / ||grad|| if ||grad|| > clip_norm else 1)
 * (optimizers/optimizers.py:136-139, 246-249, 348-351, 457-460).  No host read-back, so update() never synchronises. */
int npml_optimizer_update_clipped(int32_t kind, float* param, const float* grad, float* state0, float* state1, int64_t n,
                                  double lr, double a, double b, double eps, double c1, double c2,
                                  const double* grad_sumsq, double clip_norm, void* stream);
/* sum of squares of an fp32 buffer into *out (double, device); deterministic two-stage reduction */
int npml_sumsq_f32(const float* buf, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream);

/* ---- Pool2D (layers/layers.py:3181-3350), SURVEY.md 8f rank 3 -------------------------------
 * Uses the desc's N,H,W,C (input), fr,fc,stride, pr1,pr2,pc1,pc2 and OH,OW = (H+pr1+pr2-fr)/stride+1; K, dil, mode, act are
 * ignored.  pool_mode: 0 = max, 1 = average.  The zero padding takes part in the max / mean, as in the reference
 * (layers.py:3262).  Backward is a gather over the windows covering each input pixel (no atomics); in max mode only
 * the first maximal pixel of a window in row-major order receives its gradient (layers.py:3332-3338).
 * dtype: NPML_F32 or NPML_BF16, the same for every tensor of a call. */
enum { NPML_POOL_MAX = 0, NPML_POOL_AVERAGE = 1 };
int npml_pool2d_fwd(const npml_conv_desc* d, int32_t pool_mode, const void* X, void* Y, int32_t dtype, void* stream);
int npml_pool2d_bwd(const npml_conv_desc* d, int32_t pool_mode, const void* dY, const void* X, void* dX,
                    int32_t dtype, void* stream);

/* ---- im2col / col2im exports (utils/utils.py:486-595), API parity only; fp32 ---------------
 * X_col is (fr*fc*C, OH*OW*N) row-major with q = c*fr*fc + r*fc + t, z = (oh*OW + ow)*N + n.
 * col2im returns NCHW like the reference (utils/utils.py:595). */
int npml_im2col(const npml_conv_desc* d, const float* X, float* X_col, void* stream);
int npml_col2im(const npml_conv_desc* d, const float* X_col, float* X_nchw, void* stream);

/* ---- data-parallel gradient exchange (new; the reference is single process) ---------------
 * One communicator per process (one process per GPU, torch.distributed style).  The 128-byte
 * ncclUniqueId is created on rank 0 with npml_comm_unique_id and broadcast by the host. */
int npml_comm_unique_id(void* id128);
int npml_comm_init_rank(const void* id128, int32_t nranks, int32_t rank);
int npml_allreduce_sum_f32(float* buf, int64_t n, void* stream);
int npml_allreduce_sum_f64(double* buf, int64_t n, void* stream);   /* sync-BN statistics */
int npml_comm_destroy(void);

/* ---- all-reduce over NVLink / NVSwitch peer memory (csrc/peer.cu) -----------------------------------------------
 * The collectives of the data-parallel step are small (dW || db of a layer: 1 KB - 2 MB; sync-BN statistics: 2 KB) and
 * latency-bound, so they are ONE kernel each instead of a NCCL call: every rank pushes each 32-bit word of its buffer,
 * paired with the call number in one 8-byte store, into an inbox in every peer's memory, polls its own inbox and adds the
 * copies in rank order (the same order on every rank: bit-identical results everywhere, run after run).
 * Graph-capturable (the call counter lives in device memory).  Setup: every rank calls npml_peer_alloc (cudaMalloc + cudaIpcGetMemHandle of its region -- the one
 * allocation this library makes, at init), the host gathers the 64-byte handles, every rank calls npml_peer_open and
 * passes a host barrier.  `channel` (0..3): calls on one channel must be issued in the same order by every rank; use
 * different channels for streams that run concurrently.  n * sizeof(element) <= 2.5 MiB. */
size_t npml_peer_region_bytes(void);
int npml_peer_alloc(void* handle64);
int npml_peer_open(const void* handles, int32_t world, int32_t rank);
int npml_peer_ready(void);
int npml_peer_allreduce_sum_f32(float* buf, int64_t n, int32_t channel, void* stream);
int npml_peer_allreduce_sum_f64(double* buf, int64_t n, int32_t channel, void* stream);
int npml_peer_status(void);      /* 0 ok, 1 a block timed out waiting for a peer (synchronises the device) */
int npml_peer_close(void);

/* ---- misc --------------------------------------------------------------------------------- */
const char* npml_last_error(void);
/* 1 if the tcgen05 tensor-core path will be used for (desc, op), 0 if the CUDA-core path. */
int npml_uses_tensor_cores(const npml_conv_desc* d, int op);
/* Test hook (host only, no device work): the decisions of the weight-stationary planner (csrc/tc_ws_conv.cu; variant must
 * be 1) for a forward / dX call, honouring NPML_WS like a real call.  out[0..15] = supported,
 * pair, delta, NT, NC, n_acc, UT, SR, Hp, Wp, pady, padx, ngroups, weight TMEM columns, tiles, dynamic smem bytes, then
 * one (tap offset in slots, tap_a, tap_b) triple per MMA group.  Returns the number of ints written, < 0 on bad arguments.
 * tests/test_ws_plan_cpu.py checks it against the NumPy mirror that is itself checked against the oracle. */
int npml_debug_ws_plan(const npml_conv_desc* d, int op, int variant, int32_t* out, int32_t cap);
/* Host-only test hook: the output parity classes (and their taps) the fp32 register-tiled convolution kernel is launched
 * with for `op` (csrc/direct_fp32.cu: build_rt_classes).  out = [classes, then per class ntaps, y0, x0, so, sin, oh, ow
 * and ntaps triples (dy, dx, filter tap)]; returns the integers written or -1. */
int npml_debug_rt_classes(const npml_conv_desc* d, int op, int32_t* out, int32_t cap);
/* number of kernels launched by this library on the calling thread since the last reset */
int64_t npml_launch_count(int32_t reset);
int npml_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NPML_CONV_H_ */
