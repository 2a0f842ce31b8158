// Shared device/host helpers for libnpmlconv (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

#include "../../include/npml_conv.h"

namespace npml {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const std::string& msg);
void count_launch(int n = 1);

#define NPML_CHECK(cond, msg)                                            \
  do {                                                                   \
    if (!(cond)) {                                                       \
      npml::set_error(std::string(__func__) + ": " + (msg));             \
      return 1;                                                          \
    }                                                                    \
  } while (0)

#define NPML_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      npml::set_error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e)); \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)

#define NPML_LAUNCH_OK()                                                               \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      npml::set_error(std::string(__func__) + ": launch -> " + cudaGetErrorString(_e)); \
      return 3;                                                                        \
    }                                                                                  \
    npml::count_launch();                                                              \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  // test hook: NPML_MAX_CTAS caps every persistent grid, so that a miniature problem walks a CTA through many work
  // units (ring wrap-arounds, phase flips, the partial last unit) -- tests/test_gpu_parity.py, steady-state cases
  if (const char* e = getenv("NPML_MAX_CTAS")) {
    const int cap = atoi(e);
    if (cap > 0 && cap < n) return cap;
  }
  return n;
}

// nanoseconds the long-waiting warps of the tcgen05 kernels sleep between two polls of an mbarrier
// (ptx::mbar_wait_sleep); NPML_WAIT_NS overrides, 0 = spin
inline int wait_ns() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NPML_WAIT_NS");
    v = e ? atoi(e) : 0;       // measured on B200 (cfg2..cfg5, 0 / 128 / 512 ns): failed polls are rare (try_wait suspends in hardware), sleeping only adds wake-up latency
    if (v < 0) v = 0;
    if (v > 10000) v = 10000;
  }
  return v;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- element access ---------------------------------------------------------------------------
__device__ __forceinline__ float ld_f(const float* p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ld_f(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ float ld_f(const double* p, int64_t i) { return (float)p[i]; }
__device__ __forceinline__ void st_f(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st_f(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st_f(double* p, int64_t i, float v) { p[i] = (double)v; }

// ---- activations (activations/activations.py) -------------------------------------------------
#define NPML_SELU_ALPHA 1.6732632423543772848170429916717f
#define NPML_SELU_SCALE 1.0507009873554804934193349852946f

__device__ __forceinline__ float act_fn(int act, float z, float p0, float p1) {
  switch (act) {
    case NPML_ACT_IDENTITY: return z;
    case NPML_ACT_AFFINE: return p0 * z + p1;                                  // :362-370
    case NPML_ACT_RELU: return z > 0.f ? z : 0.f;                              // :104-114
    case NPML_ACT_LEAKY_RELU: return z < 0.f ? z * p0 : z;                     // :169-181
    case NPML_ACT_TANH: return tanhf(z);                                       // :313-315
    case NPML_ACT_SIGMOID: return 1.f / (1.f + expf(-z));                      // :39-47
    case NPML_ACT_ELU: return z > 0.f ? z : p0 * (expf(z) - 1.f);              // :449-460
    case NPML_ACT_SELU:                                                        // :568-584
      return NPML_SELU_SCALE * (z > 0.f ? z : NPML_SELU_ALPHA * (expf(z) - 1.f));
    case NPML_ACT_HARD_SIGMOID: return fminf(fmaxf(0.2f * z + 0.5f, 0.f), 1.f);  // :631-642
    case NPML_ACT_SOFTPLUS: return logf(expf(z) + 1.f);                        // :688-696
    case NPML_ACT_GELU: {                                                      // :239-251
      float u = 0.7978845608028654f * (z + 0.044715f * z * z * z);
      return 0.5f * z * (1.f + tanhf(u));
    }
    case NPML_ACT_EXPONENTIAL: return expf(z);                                 // :500-507
  }
  return z;
}

// Epilogue form: act over 32 accumulator columns held in registers.  Identity and ReLU are inlined; every other
// activation goes through ONE out-of-line copy of the switch (8 values per call) so that the tensor-core kernels
// stay small enough for the instruction cache (4 values per call; 32 inlined copies of the full switch made the conv1 + ReLU
// epilogue of cfg4 instruction-fetch bound).
// The tensor-core epilogues (bf16 / tf32 modes, tolerances 2e-2 / 5e-3) evaluate the transcendental activations with the
// hardware approximations: tanh.approx.f32 (max abs error 2^-11), __expf / __logf (2 ulp), one MUFU instead of the
// 30-40 instruction libm paths.  Measured on cfg4's conv1 shape (scripts/epi_probe.py): tanh epilogue 90 -> 76 us
// (identity: 43 us; what remains is the out-of-line call per four values).  The fp32 exactness mode never comes here
// (its CUDA-core kernels call act_fn).
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float act_fn_fast(int act, float z, float p0, float p1) {
  switch (act) {
    case NPML_ACT_TANH: return tanh_fast(z);
    case NPML_ACT_SIGMOID: return fmaf(0.5f, tanh_fast(0.5f * z), 0.5f);       // 1 / (1 + e^-z) = (1 + tanh(z/2)) / 2
    case NPML_ACT_ELU: return z > 0.f ? z : p0 * (__expf(z) - 1.f);
    case NPML_ACT_SELU: return NPML_SELU_SCALE * (z > 0.f ? z : NPML_SELU_ALPHA * (__expf(z) - 1.f));
    case NPML_ACT_SOFTPLUS: return z > 15.f ? z : __logf(__expf(z) + 1.f);     // log(1 + e^z) = z + O(e^-z): 3e-7 at 15
    case NPML_ACT_GELU: {
      const float u = 0.7978845608028654f * (z + 0.044715f * z * z * z);
      return 0.5f * z * (1.f + tanh_fast(u));
    }
    case NPML_ACT_EXPONENTIAL: return __expf(z);
    default: return act_fn(act, z, p0, p1);
  }
}
static __device__ __noinline__ float4 act_fn4(int act, float p0, float p1, float4 x) {
  return make_float4(act_fn_fast(act, x.x, p0, p1), act_fn_fast(act, x.y, p0, p1), act_fn_fast(act, x.z, p0, p1),
                     act_fn_fast(act, x.w, p0, p1));
}
__device__ __forceinline__ void act_apply32(int act, float p0, float p1, float (&v)[32]) {
  if (act == NPML_ACT_IDENTITY) return;
  if (act == NPML_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
    return;
  }
  // one- to three-instruction activations inline, the switch hoisted out of the element loop (scripts/epi_probe.py: the
  // out-of-line call per four values cost a tanh epilogue 33 us on cfg4's conv1 shape)
  if (act == NPML_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = tanh_fast(v[i]);
    return;
  }
  if (act == NPML_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(0.5f, tanh_fast(0.5f * v[i]), 0.5f);
    return;
  }
  if (act == NPML_ACT_LEAKY_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] < 0.f ? v[i] * p0 : v[i];
    return;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 4) {     // values travel in registers (by value), so v[] never needs a stack slot
    const float4 r = act_fn4(act, p0, p1, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
    v[i] = r.x; v[i + 1] = r.y; v[i + 2] = r.z; v[i + 3] = r.w;
  }
}

// derivative conventions at the kinks follow the reference
__device__ __forceinline__ float act_grad(int act, float x, float p0, float p1) {
  switch (act) {
    case NPML_ACT_IDENTITY: return 1.f;
    case NPML_ACT_AFFINE: return p0;                                           // :372-381
    case NPML_ACT_RELU: return x > 0.f ? 1.f : 0.f;                            // :116-126
    case NPML_ACT_LEAKY_RELU: return x < 0.f ? p0 : 1.f;                       // :183-196
    case NPML_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }            // :317-326
    case NPML_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }  // :49-58
    case NPML_ACT_ELU: return x > 0.f ? 1.f : p0 * expf(x);                    // :462-474
    case NPML_ACT_SELU:                                                        // :586-599
      return x >= 0.f ? NPML_SELU_SCALE : expf(x) * NPML_SELU_ALPHA * NPML_SELU_SCALE;
    case NPML_ACT_HARD_SIGMOID: return (x >= -2.5f && x <= 2.5f) ? 0.2f : 0.f;  // :644-655
    case NPML_ACT_SOFTPLUS: { float e = expf(x); return e / (e + 1.f); }       // :698-708
    case NPML_ACT_GELU: {                                                      // :254-277 (literal)
      float s = x * 0.7071067811865475f;
      float erf_prime = 1.1283791670955126f * expf(-(s * s));
      float approx = tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x));
      return 0.5f + 0.5f * approx + (0.5f * x * erf_prime) * 0.7071067811865475f;
    }
    case NPML_ACT_EXPONENTIAL: return expf(x);                                 // :509-518
  }
  return 1.f;
}

// ---- internal "gather convolution" description -------------------------------------------------
// out[n, y, x, co] = sum_{r, t, ci} in[n, iy, ix, ci] * w[(r*fc + t)*w_tap + ci*w_ci + co*w_co]
//   form 0 (fprop-like):  iy = y*s + r*D - pr,            ix = x*s + t*D - pc
//   form 1 (dgrad-like):  iy = (y + pr - r*D) / s if s divides it (else the tap is skipped)
// Conv2D.forward / Deconv2D dX use form 0; Conv2D dX / Deconv2D.forward use form 1.
struct GatherConv {
  int N, IH, IW, CI;    // tensor that is gathered from
  int OH, OW, CO;       // tensor that is produced
  int fr, fc, s, D, pr, pc;
  int form;
  int64_t w_tap, w_ci, w_co;
  int act;
  float p0, p1;
  // optional per-channel affine applied AFTER the activation: Y = post_scale[co] * act(z + b) + post_shift[co]
  // (a frozen BatchNorm2D that follows the convolution, layers/layers.py:1141-1156, folded into the epilogue);
  // device pointers, CO entries each, both NULL when unused.  Z (the pre-activation) is never affected.
  const float* post_scale;
  const float* post_shift;
  // NPML_FLAG_WEIGHTS_PACKED: the workspace already holds Wt for these weights; launch_repack_weights is a no-op
  bool weights_packed;
};

__device__ __forceinline__ float post_affine(const GatherConv& g, int co, float y) {
  return g.post_scale ? fmaf(y, g.post_scale[co], g.post_shift[co]) : y;
}

// wgrad:  dW[(r*fc+t)*o_tap + ci*o_ci + co*o_co] (+)= sum_{n,y,x} in[n, y*s + r*D - pr, x*s + t*D - pc, ci]
//                                                                * g[n, y, x, co]
// where `in` is (N, IH, IW, CI) and `g` is (N, OH, OW, CO).
struct GatherWgrad {
  int N, IH, IW, CI;
  int OH, OW, CO;
  int fr, fc, s, D, pr, pc;
  int64_t o_tap, o_ci, o_co;
  int accumulate;
  int no_db;            // the caller never asks this problem for db (Deconv2D: its db is a column sum of the OTHER operand)
};

// direct (CUDA-core, fp32 accumulate) kernels, direct_fp32.cu
int direct_gather_conv(const GatherConv& g, const void* in, const float* w, const float* bias, void* Z,
                       void* Y, int dtype, cudaStream_t st);
int rt_classes_debug(const GatherConv& g, int32_t* out, int32_t cap);   // host-only test hook
size_t direct_wgrad_ws_bytes(const GatherWgrad& g);
int direct_wgrad(const GatherWgrad& g, const void* in, const void* grad, float* dW, int dtype, void* ws,
                 size_t ws_bytes, cudaStream_t st, float* db = nullptr, bool* db_done = nullptr);

// one-launch backward (dX, dW, db) of thin stride-1 "same" convolutions, thin_bwd.cu
bool thin_bwd_supported(const npml_conv_desc* d);
size_t thin_bwd_ws_bytes(const npml_conv_desc* d);
int thin_bwd(const npml_conv_desc* d, const void* X, const void* dZ, const float* W, float* dW, float* db, void* dX,
             void* ws, size_t ws_bytes, cudaStream_t st);

// tensor-core (tcgen05) kernels, tc_conv.cu / tc_wgrad.cu
bool tc_conv_supported(const GatherConv& g, int mode);
size_t tc_conv_ws_bytes(const GatherConv& g, int mode);
int tc_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z,
                   void* Y, void* ws, size_t ws_bytes, cudaStream_t st);
// persistent slab kernel for stride-1 gather convolutions, tc_slab_conv.cu
bool slab_conv_supported(const GatherConv& g, int mode);
size_t slab_conv_ws_bytes(const GatherConv& g, int mode);
int slab_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z,
                     void* Y, void* ws, size_t ws_bytes, cudaStream_t st);
// weight-stationary variant (weights in tensor memory, pixels along N) for out_ch <= 128, tc_ws_conv.cu
bool ws_conv_supported(const GatherConv& g, int mode);
size_t ws_conv_ws_bytes(const GatherConv& g, int mode);
int ws_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z, void* Y,
                   void* ws, size_t ws_bytes, cudaStream_t st);
int ws_plan_debug(const GatherConv& g, int mode, int32_t* out, int32_t cap);      // host-only test hooks
// row-of-taps variant for narrow outputs (fc * out_ch <= 256), tc_slab3_conv.cu; slab_gather_conv dispatches to it
bool slab3_conv_supported(const GatherConv& g, int mode);
int slab3_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z,
                      void* Y, void* ws, size_t ws_bytes, cudaStream_t st);
// persistent slab wgrad (+ db) for stride-1 convolutions, tc_slab_wgrad.cu
bool slab_wgrad_supported(const GatherWgrad& g, int mode);
size_t slab_wgrad_ws_bytes(const GatherWgrad& g, int mode);
int slab_wgrad(const GatherWgrad& g, int mode, const void* in, const void* grad, float* dW, float* db, void* ws,
               size_t ws_bytes, cudaStream_t st);
bool tc_wgrad_supported(const GatherWgrad& g, int mode);
size_t tc_wgrad_ws_bytes(const GatherWgrad& g, int mode);
int tc_wgrad(const GatherWgrad& g, int mode, const void* in, const void* grad, float* dW, void* ws,
             size_t ws_bytes, cudaStream_t st, float* db = nullptr);

}  // namespace npml
