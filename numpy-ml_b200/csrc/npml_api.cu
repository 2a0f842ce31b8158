// C-ABI entry points of libnpmlconv.so (declared in include/npml_conv.h): maps the six layer
// operations onto the internal gather-conv / wgrad forms and picks the tcgen05 or CUDA-core kernel.
#include "npml_common.cuh"

namespace npml {

static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }

static int act_dtype(int mode) { return mode == NPML_MODE_BF16 ? NPML_BF16 : NPML_F32; }

static int check_desc(const npml_conv_desc* d, bool deconv) {
  NPML_CHECK(d != nullptr, "desc is NULL");
  NPML_CHECK(d->N >= 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0, "bad tensor dims");
  NPML_CHECK(d->fr > 0 && d->fc > 0 && d->stride > 0 && d->dil > 0, "bad kernel/stride/dilation");
  NPML_CHECK(d->pr1 >= 0 && d->pr2 >= 0 && d->pc1 >= 0 && d->pc2 >= 0, "negative padding");
  NPML_CHECK(d->OH > 0 && d->OW > 0, "non-positive output dims");
  NPML_CHECK(d->mode >= NPML_MODE_FP32 && d->mode <= NPML_MODE_BF16, "bad mode");
  if (deconv) {
    NPML_CHECK(d->dil == 1, "Deconv2D has no dilation");
    NPML_CHECK(d->OH == d->stride * (d->H - 1) + d->fr - d->pr1 - d->pr2, "OH inconsistent with deconv geometry");
    NPML_CHECK(d->OW == d->stride * (d->W - 1) + d->fc - d->pc1 - d->pc2, "OW inconsistent with deconv geometry");
  } else {
    const int er = d->fr * d->dil - d->dil + 1, ec = d->fc * d->dil - d->dil + 1;
    NPML_CHECK(d->OH == (d->H + d->pr1 + d->pr2 - er) / d->stride + 1, "OH inconsistent with conv geometry");
    NPML_CHECK(d->OW == (d->W + d->pc1 + d->pc2 - ec) / d->stride + 1, "OW inconsistent with conv geometry");
  }
  return 0;
}

// op -> internal forms ---------------------------------------------------------------------------------
static GatherConv make_gather(const npml_conv_desc* d, int op) {
  GatherConv g{};
  g.N = d->N; g.fr = d->fr; g.fc = d->fc; g.s = d->stride; g.D = d->dil; g.pr = d->pr1; g.pc = d->pc1;
  g.w_tap = (int64_t)d->C * d->K;
  g.act = NPML_ACT_IDENTITY; g.p0 = 0.f; g.p1 = 0.f;
  g.weights_packed = (d->flags & NPML_FLAG_WEIGHTS_PACKED) != 0;
  switch (op) {
    case NPML_OP_CONV_FPROP:    // Z[n,oh,ow,k] = sum X[n, oh*s + r*D - pr1, ow*s + t*D - pc1, c] W[r,t,c,k]
    case NPML_OP_DECONV_FPROP:  // Z[n,y,x,k]   = sum X[n, (y + pr1 - r)/s, (x + pc1 - t)/s, c] W[r,t,c,k]
      g.IH = d->H; g.IW = d->W; g.CI = d->C; g.OH = d->OH; g.OW = d->OW; g.CO = d->K;
      g.w_ci = d->K; g.w_co = 1;
      g.form = (op == NPML_OP_CONV_FPROP) ? 0 : 1;
      g.act = d->act; g.p0 = d->act_p0; g.p1 = d->act_p1;
      break;
    case NPML_OP_CONV_DGRAD:    // dX[n,h,w,c] = sum dZ[n, (h + pr1 - r*D)/s, (w + pc1 - t*D)/s, k] W[r,t,c,k]
    case NPML_OP_DECONV_DGRAD:  // dX[n,iy,ix,c] = sum dZ[n, iy*s + r - pr1, ix*s + t - pc1, k] W[r,t,c,k]
      g.IH = d->OH; g.IW = d->OW; g.CI = d->K; g.OH = d->H; g.OW = d->W; g.CO = d->C;
      g.w_ci = 1; g.w_co = d->K;
      g.form = (op == NPML_OP_CONV_DGRAD) ? 1 : 0;
      break;
  }
  return g;
}

static GatherWgrad make_wgrad(const npml_conv_desc* d, int op) {
  GatherWgrad g{};
  g.N = d->N; g.fr = d->fr; g.fc = d->fc; g.s = d->stride; g.D = d->dil; g.pr = d->pr1; g.pc = d->pc1;
  g.o_tap = (int64_t)d->C * d->K;
  g.accumulate = d->accumulate;
  if (op == NPML_OP_CONV_WGRAD) {  // dW[r,t,c,k] = sum X[n, oh*s + r*D - pr1, ...] dZ[n,oh,ow,k]
    g.IH = d->H; g.IW = d->W; g.CI = d->C; g.OH = d->OH; g.OW = d->OW; g.CO = d->K;
    g.o_ci = d->K; g.o_co = 1;
  } else {                         // dW[r,t,c,k] = sum dZ[n, iy*s + r - pr1, ...][k] X[n,iy,ix,c]
    g.IH = d->OH; g.IW = d->OW; g.CI = d->K; g.OH = d->H; g.OW = d->W; g.CO = d->C;
    g.o_ci = 1; g.o_co = d->K;
    g.no_db = 1;
  }
  return g;
}

size_t colpass_ws_bytes(int64_t rows, int ch, int nv);

// workspace bytes of the convolution kernel run_gather would pick (the packed weights)
static size_t gather_ws_bytes(const GatherConv& g, int mode) {
  if (mode == NPML_MODE_BF16 && ws_conv_supported(g, mode)) return ws_conv_ws_bytes(g, mode);
  if (mode != NPML_MODE_FP32 && slab_conv_supported(g, mode)) return slab_conv_ws_bytes(g, mode);
  if (mode != NPML_MODE_FP32 && tc_conv_supported(g, mode)) return tc_conv_ws_bytes(g, mode);
  return 0;
}

static int run_gather(const npml_conv_desc* d, int op, const void* in, const float* W, const float* b, void* Z,
                      void* Y, void* ws, size_t ws_bytes, cudaStream_t st, const float* post_scale = nullptr,
                      const float* post_shift = nullptr, double* bn_sums = nullptr) {
  if (check_desc(d, op == NPML_OP_DECONV_FPROP || op == NPML_OP_DECONV_DGRAD)) return 1;
  if (d->N == 0) return 0;
  NPML_CHECK(in && W, "NULL input");
  NPML_CHECK(Z || Y, "no output buffer");
  GatherConv g = make_gather(d, op);
  NPML_CHECK((post_scale == nullptr) == (post_shift == nullptr), "bn_scale and bn_shift go together");
  NPML_CHECK(post_scale == nullptr || Y != nullptr, "the folded BatchNorm2D needs the Y output");
  g.post_scale = post_scale; g.post_shift = post_shift;
  if (g.act == NPML_ACT_IDENTITY && Z == nullptr && post_scale == nullptr) { Z = Y; Y = nullptr; }
  if (bn_sums != nullptr) {
    // statistics of the layer output for the BatchNorm2D behind it: one column pass over the tensor the convolution
    // has just written (L2-hot), inside the same call.  Forming them in the tcgen05 kernels' epilogue was measured and
    // LOSES (cfg4, three convolutions: 208 us with the pass, 263 us with a shuffle transpose-reduce in the epilogue,
    // 308 us on the staged read-back path with one epilogue group; profiles/r2j_cfg4_conv_stats_ab.md): those kernels
    // are bound by their epilogue already.
    const size_t wbytes = align_up(gather_ws_bytes(g, d->mode), 256);
    const int64_t rows = (int64_t)d->N * d->OH * d->OW;
    NPML_CHECK(ws != nullptr && ws_bytes >= wbytes + colpass_ws_bytes(rows, d->K, 2), "workspace too small");
    if (int e = run_gather(d, op, in, W, b, Z, Y, ws, wbytes, st, post_scale, post_shift)) return e;
    const void* out = Y ? Y : Z;
    return npml_bn2d_stats(rows, d->K, out, act_dtype(d->mode), bn_sums, (char*)ws + wbytes, ws_bytes - wbytes, (void*)st);
  }
  if (d->mode == NPML_MODE_BF16 && ws_conv_supported(g, d->mode))
    return ws_gather_conv(g, d->mode, in, W, b, Z, Y, ws, ws_bytes, st);
  if (d->mode != NPML_MODE_FP32 && slab_conv_supported(g, d->mode))
    return slab_gather_conv(g, d->mode, in, W, b, Z, Y, ws, ws_bytes, st);
  if (d->mode != NPML_MODE_FP32 && tc_conv_supported(g, d->mode))
    return tc_gather_conv(g, d->mode, in, W, b, Z, Y, ws, ws_bytes, st);
  return direct_gather_conv(g, in, W, b, Z, Y, act_dtype(d->mode), st);
}

// db != nullptr is only passed for Conv2D (grad == dZ); paths without a fused db fall back to one column pass.
static int run_wgrad(const npml_conv_desc* d, int op, const void* in, const void* grad, float* dW, void* ws,
                     size_t ws_bytes, cudaStream_t st, float* db = nullptr) {
  if (check_desc(d, op == NPML_OP_DECONV_WGRAD)) return 1;
  NPML_CHECK(dW != nullptr, "dW is NULL");
  GatherWgrad g = make_wgrad(d, op);
  if (d->N == 0) {
    if (!d->accumulate) NPML_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * d->fr * d->fc * d->C * d->K, st));
    if (db && !d->accumulate) NPML_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * d->K, st));
    return 0;
  }
  NPML_CHECK(in && grad, "NULL input");
  if (d->mode != NPML_MODE_FP32 && slab_wgrad_supported(g, d->mode))
    return slab_wgrad(g, d->mode, in, grad, dW, db, ws, ws_bytes, st);
  int e;
  bool db_done = false;
  if (d->mode != NPML_MODE_FP32 && tc_wgrad_supported(g, d->mode)) {
    e = tc_wgrad(g, d->mode, in, grad, dW, ws, ws_bytes, st, db);
    db_done = db != nullptr;
  } else
    e = direct_wgrad(g, in, grad, dW, act_dtype(d->mode), ws, ws_bytes, st, db, &db_done);
  if (e || !db || db_done) return e;
  // the workspace is free again once the reduction above has been enqueued (stream order)
  const int dt = act_dtype(d->mode);
  return npml_dz_prep((int64_t)d->N * d->OH * d->OW, d->K, NPML_ACT_IDENTITY, 0.f, 0.f, grad, dt, nullptr, dt, nullptr, dt,
                      db, d->accumulate, ws, ws_bytes, (void*)st);
}

}  // namespace npml

using namespace npml;

extern "C" {

int npml_conv2d_fprop(const npml_conv_desc* d, const void* X, const float* W, const float* b, void* Z, void* Y,
                      void* ws, size_t ws_bytes, void* stream) {
  return run_gather(d, NPML_OP_CONV_FPROP, X, W, b, Z, Y, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_conv2d_fprop_stats(const npml_conv_desc* d, const void* X, const float* W, const float* b, void* Z, void* Y,
                            double* bn_sums, void* ws, size_t ws_bytes, void* stream) {
  NPML_CHECK(bn_sums != nullptr, "bn_sums is NULL");
  return run_gather(d, NPML_OP_CONV_FPROP, X, W, b, Z, Y, ws, ws_bytes, (cudaStream_t)stream, nullptr, nullptr, bn_sums);
}
int npml_conv2d_fprop_bn(const npml_conv_desc* d, const void* X, const float* W, const float* b, const float* bn_scale,
                         const float* bn_shift, void* Z, void* Y, void* ws, size_t ws_bytes, void* stream) {
  return run_gather(d, NPML_OP_CONV_FPROP, X, W, b, Z, Y, ws, ws_bytes, (cudaStream_t)stream, bn_scale, bn_shift);
}
int npml_deconv2d_fprop_bn(const npml_conv_desc* d, const void* X, const float* W, const float* b, const float* bn_scale,
                           const float* bn_shift, void* Z, void* Y, void* ws, size_t ws_bytes, void* stream) {
  return run_gather(d, NPML_OP_DECONV_FPROP, X, W, b, Z, Y, ws, ws_bytes, (cudaStream_t)stream, bn_scale, bn_shift);
}
int npml_conv2d_dgrad(const npml_conv_desc* d, const void* dZ, const float* W, void* dX, void* ws, size_t ws_bytes,
                      void* stream) {
  return run_gather(d, NPML_OP_CONV_DGRAD, dZ, W, nullptr, dX, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_conv2d_wgrad(const npml_conv_desc* d, const void* X, const void* dZ, float* dW, void* ws, size_t ws_bytes,
                      void* stream) {
  return run_wgrad(d, NPML_OP_CONV_WGRAD, X, dZ, dW, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_conv2d_wgrad_db(const npml_conv_desc* d, const void* X, const void* dZ, float* dW, float* db, void* ws,
                         size_t ws_bytes, void* stream) {
  return run_wgrad(d, NPML_OP_CONV_WGRAD, X, dZ, dW, ws, ws_bytes, (cudaStream_t)stream, db);
}
int npml_conv2d_bwd_is_fused(const npml_conv_desc* d) {
  return d != nullptr && d->N > 0 && d->dil > 0 && thin_bwd_supported(d) ? 1 : 0;
}
int npml_conv2d_bwd(const npml_conv_desc* d, const void* X, const void* dZ, const float* W, float* dW, float* db, void* dX,
                    void* ws, size_t ws_bytes, void* stream) {
  if (check_desc(d, false)) return 1;
  NPML_CHECK(dW != nullptr && dX != nullptr, "dW / dX is NULL");
  if (d->N > 0 && thin_bwd_supported(d)) {
    NPML_CHECK(X && dZ && W, "NULL input");
    return thin_bwd(d, X, dZ, W, dW, db, dX, ws, ws_bytes, (cudaStream_t)stream);
  }
  if (int e = run_wgrad(d, NPML_OP_CONV_WGRAD, X, dZ, dW, ws, ws_bytes, (cudaStream_t)stream, db)) return e;
  return run_gather(d, NPML_OP_CONV_DGRAD, dZ, W, nullptr, dX, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_deconv2d_fprop(const npml_conv_desc* d, const void* X, const float* W, const float* b, void* Z, void* Y,
                        void* ws, size_t ws_bytes, void* stream) {
  return run_gather(d, NPML_OP_DECONV_FPROP, X, W, b, Z, Y, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_deconv2d_dgrad(const npml_conv_desc* d, const void* dZ, const float* W, void* dX, void* ws,
                        size_t ws_bytes, void* stream) {
  return run_gather(d, NPML_OP_DECONV_DGRAD, dZ, W, nullptr, dX, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}
int npml_deconv2d_wgrad(const npml_conv_desc* d, const void* X, const void* dZ, float* dW, void* ws,
                        size_t ws_bytes, void* stream) {
  // the shifted (gathered) operand is dZ, the per-pixel operand is X
  return run_wgrad(d, NPML_OP_DECONV_WGRAD, dZ, X, dW, ws, ws_bytes, (cudaStream_t)stream);
}

size_t npml_workspace_bytes(const npml_conv_desc* d, int op) {
  if (!d) return 0;
  size_t bytes = 256;
  switch (op) {
    case NPML_OP_CONV_FPROP: case NPML_OP_CONV_DGRAD: case NPML_OP_DECONV_FPROP: case NPML_OP_DECONV_DGRAD: {
      GatherConv g = make_gather(d, op);
      bytes += align_up(gather_ws_bytes(g, d->mode), 256);
      if (op == NPML_OP_CONV_FPROP)       // npml_conv2d_fprop_stats: the column pass behind the convolution
        bytes += colpass_ws_bytes((int64_t)d->N * d->OH * d->OW, d->K, 2);
      break;
    }
    case NPML_OP_CONV_WGRAD: case NPML_OP_DECONV_WGRAD: {
      GatherWgrad g = make_wgrad(d, op);
      size_t w;
      if (d->mode != NPML_MODE_FP32 && slab_wgrad_supported(g, d->mode)) w = slab_wgrad_ws_bytes(g, d->mode);
      else if (d->mode != NPML_MODE_FP32 && tc_wgrad_supported(g, d->mode)) w = tc_wgrad_ws_bytes(g, d->mode);
      else w = direct_wgrad_ws_bytes(g);
      const size_t c = colpass_ws_bytes((int64_t)d->N * d->OH * d->OW, d->K, 1);   // db fallback of npml_conv2d_wgrad_db
      bytes += w > c ? w : c;
      break;
    }
    case NPML_OP_CONV_BWD: {
      const size_t a = npml_workspace_bytes(d, NPML_OP_CONV_WGRAD), b = npml_workspace_bytes(d, NPML_OP_CONV_DGRAD);
      const size_t c = d->N > 0 && d->dil > 0 && thin_bwd_supported(d) ? thin_bwd_ws_bytes(d) : 0;
      bytes += (a > b ? a : b) > c ? (a > b ? a : b) : c;
      break;
    }
    case NPML_OP_DZ_PREP: {
      const int64_t rows = (int64_t)d->N * d->OH * d->OW;
      bytes += colpass_ws_bytes(rows, d->K, 1);
      break;
    }
    case NPML_OP_BN: {
      const int64_t rows = (int64_t)d->N * d->OH * d->OW;
      bytes += colpass_ws_bytes(rows, d->K, 3);       // up to three column sums per channel (npml_bn2d_pair_bwd_stats)
      break;
    }
  }
  return bytes;
}

int npml_uses_tensor_cores(const npml_conv_desc* d, int op) {
  if (!d || d->mode == NPML_MODE_FP32) return 0;
  if (op == NPML_OP_CONV_WGRAD || op == NPML_OP_DECONV_WGRAD)
    return slab_wgrad_supported(make_wgrad(d, op), d->mode) || tc_wgrad_supported(make_wgrad(d, op), d->mode);
  if (op >= NPML_OP_CONV_FPROP && op <= NPML_OP_DECONV_DGRAD)
    return ws_conv_supported(make_gather(d, op), d->mode) || slab_conv_supported(make_gather(d, op), d->mode) ||
           tc_conv_supported(make_gather(d, op), d->mode);
  return 0;
}

int npml_debug_ws_plan(const npml_conv_desc* d, int op, int variant, int32_t* out, int32_t cap) {
  if (!d || !out || op < NPML_OP_CONV_FPROP || op > NPML_OP_DECONV_DGRAD) return -1;
  const GatherConv g = make_gather(d, op);
  NPML_CHECK(variant == 1, "only variant 1 exists");
  return ws_plan_debug(g, d->mode, out, cap);
}

int npml_debug_rt_classes(const npml_conv_desc* d, int op, int32_t* out, int32_t cap) {
  if (!d || !out || op < NPML_OP_CONV_FPROP || op > NPML_OP_DECONV_DGRAD) return -1;
  return rt_classes_debug(make_gather(d, op), out, cap);
}

const char* npml_last_error(void) { return g_err.c_str(); }

int64_t npml_launch_count(int32_t reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int npml_version(void) { return 100; }

}  // extern "C"
