// Pool2D forward / backward on NHWC tensors (numpy_ml/neural_nets/layers/layers.py:3181-3350), the layer that sits
// next to Conv2D in the reference's models.  HBM-bound gathers:
//   forward : one thread per (output pixel, VEC channels); the zero padding takes part in the max / mean like the
//             reference's np.pad (layers.py:3262).
//   backward: one thread per (INPUT pixel, VEC channels) that visits the windows covering the pixel -- a gather, so no
//             atomics and a fixed summation order.  max mode re-scans each covering window in row-major order and
//             credits the pixel only if it is the FIRST maximal one (layers.py:3332-3338).
#include "npml_common.cuh"

namespace npml {
namespace {

template <typename T, int VEC>
__device__ __forceinline__ void ldv(const T* __restrict__ p, int64_t i, float (&v)[VEC]) {
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if constexpr (VEC == 4 && sizeof(T) == 2) {
    const uint2 t = *reinterpret_cast<const uint2*>(p + i);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = ld_f(p, i + j);
  }
}

template <typename T, int VEC>
__device__ __forceinline__ void stv(T* __restrict__ p, int64_t i, const float (&v)[VEC]) {
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
  } else if constexpr (VEC == 4 && sizeof(T) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p + i) = t;
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) st_f(p, i + j, v[j]);
  }
}

struct PoolGeom {
  int N, H, W, C, OH, OW, fr, fc, s, pr, pc, is_max;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256) pool_fwd_kernel(PoolGeom g, const T* __restrict__ X, T* __restrict__ Y) {
  const int cv = g.C / VEC;
  const int64_t total = (int64_t)g.N * g.OH * g.OW * cv;
  const float inv = 1.f / (float)(g.fr * g.fc);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    int64_t pix = idx / cv;
    const int ox = (int)(pix % g.OW);
    pix /= g.OW;
    const int oy = (int)(pix % g.OH);
    const int n = (int)(pix / g.OH);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = g.is_max ? -INFINITY : 0.f;
    for (int r = 0; r < g.fr; ++r) {
      const int iy = oy * g.s + r - g.pr;
      for (int t = 0; t < g.fc; ++t) {
        const int ix = ox * g.s + t - g.pc;
        float v[VEC];
        if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
          ldv<T, VEC>(X, (((int64_t)n * g.H + iy) * g.W + ix) * g.C + c, v);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) v[j] = 0.f;          // a zero pad pixel
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = g.is_max ? fmaxf(acc[j], v[j]) : acc[j] + v[j];
      }
    }
    if (!g.is_max) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] *= inv;
    }
    stv<T, VEC>(Y, idx * VEC, acc);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) pool_bwd_kernel(PoolGeom g, const T* __restrict__ dY, const T* __restrict__ X,
                                                       T* __restrict__ dX) {
  const int cv = g.C / VEC;
  const int64_t total = (int64_t)g.N * g.H * g.W * cv;
  const float inv = 1.f / (float)(g.fr * g.fc);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * VEC;
    int64_t pix = idx / cv;
    const int x = (int)(pix % g.W);
    pix /= g.W;
    const int y = (int)(pix % g.H);
    const int n = (int)(pix / g.H);
    const int yp = y + g.pr, xp = x + g.pc;                    // position in the padded volume
    // windows (oy, ox) with oy*s <= yp < oy*s + fr
    const int oy_lo = yp - g.fr + 1 <= 0 ? 0 : (yp - g.fr + g.s) / g.s;
    const int oy_hi = min(yp / g.s, g.OH - 1);
    const int ox_lo = xp - g.fc + 1 <= 0 ? 0 : (xp - g.fc + g.s) / g.s;
    const int ox_hi = min(xp / g.s, g.OW - 1);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    float mine[VEC];
    if (g.is_max) ldv<T, VEC>(X, idx * VEC, mine);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        float d[VEC];
        ldv<T, VEC>(dY, (((int64_t)n * g.OH + oy) * g.OW + ox) * g.C + c, d);
        if (!g.is_max) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] += d[j] * inv;
          continue;
        }
        // this pixel takes the window's gradient iff nothing before it (row-major) is >= it and nothing after it
        // is > it, i.e. it is the first maximal element
        const int my_r = yp - oy * g.s, my_t = xp - ox * g.s;
        bool win[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) win[j] = true;
        for (int r = 0; r < g.fr; ++r) {
          const int iy = oy * g.s + r - g.pr;
          for (int t = 0; t < g.fc; ++t) {
            if (r == my_r && t == my_t) continue;
            const int ix = ox * g.s + t - g.pc;
            float v[VEC];
            if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) {
              ldv<T, VEC>(X, (((int64_t)n * g.H + iy) * g.W + ix) * g.C + c, v);
            } else {
#pragma unroll
              for (int j = 0; j < VEC; ++j) v[j] = 0.f;
            }
            const bool before = r < my_r || (r == my_r && t < my_t);
#pragma unroll
            for (int j = 0; j < VEC; ++j) win[j] = win[j] && (before ? v[j] < mine[j] : v[j] <= mine[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += win[j] ? d[j] : 0.f;
      }
    }
    stv<T, VEC>(dX, idx * VEC, acc);
  }
}

inline int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 16LL * num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int check_geom(const npml_conv_desc* d, int mode, PoolGeom& g) {
  NPML_CHECK(d != nullptr, "NULL descriptor");
  NPML_CHECK(mode == 0 || mode == 1, "pool mode must be 0 (max) or 1 (average)");
  NPML_CHECK(d->fr >= 1 && d->fc >= 1 && d->stride >= 1 && d->N >= 0 && d->C >= 1, "bad geometry");
  NPML_CHECK(d->OH == (d->H + d->pr1 + d->pr2 - d->fr) / d->stride + 1 && d->OW == (d->W + d->pc1 + d->pc2 - d->fc) / d->stride + 1,
             "OH / OW do not match the geometry");
  g = PoolGeom{d->N, d->H, d->W, d->C, d->OH, d->OW, d->fr, d->fc, d->stride, d->pr1, d->pc1, mode == 0 ? 1 : 0};
  return 0;
}

template <typename T>
bool vec4_ok(int C, const void* a, const void* b, const void* c) {
  const uintptr_t m = 4 * sizeof(T) - 1;
  return C % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & m) == 0;
}

}  // namespace
}  // namespace npml

using namespace npml;

extern "C" int npml_pool2d_fwd(const npml_conv_desc* d, int32_t pool_mode, const void* X, void* Y, int32_t dtype,
                               void* stream) {
  PoolGeom g;
  if (int e = check_geom(d, pool_mode, g)) return e;
  NPML_CHECK(dtype == NPML_F32 || dtype == NPML_BF16, "dtype must be NPML_F32 or NPML_BF16");
  const int64_t pixels = (int64_t)g.N * g.OH * g.OW;
  if (pixels * g.C == 0) return 0;
  NPML_CHECK(X && Y, "NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NPML_F32) {
    if (vec4_ok<float>(g.C, X, Y, nullptr))
      pool_fwd_kernel<float, 4><<<blocks_for(pixels * g.C / 4), 256, 0, st>>>(g, (const float*)X, (float*)Y);
    else
      pool_fwd_kernel<float, 1><<<blocks_for(pixels * g.C), 256, 0, st>>>(g, (const float*)X, (float*)Y);
  } else {
    if (vec4_ok<__nv_bfloat16>(g.C, X, Y, nullptr))
      pool_fwd_kernel<__nv_bfloat16, 4><<<blocks_for(pixels * g.C / 4), 256, 0, st>>>(g, (const __nv_bfloat16*)X, (__nv_bfloat16*)Y);
    else
      pool_fwd_kernel<__nv_bfloat16, 1><<<blocks_for(pixels * g.C), 256, 0, st>>>(g, (const __nv_bfloat16*)X, (__nv_bfloat16*)Y);
  }
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_pool2d_bwd(const npml_conv_desc* d, int32_t pool_mode, const void* dY, const void* X, void* dX,
                               int32_t dtype, void* stream) {
  PoolGeom g;
  if (int e = check_geom(d, pool_mode, g)) return e;
  NPML_CHECK(dtype == NPML_F32 || dtype == NPML_BF16, "dtype must be NPML_F32 or NPML_BF16");
  const int64_t pixels = (int64_t)g.N * g.H * g.W;
  if (pixels * g.C == 0) return 0;
  NPML_CHECK(dY && dX && (X || !g.is_max), "NULL tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NPML_F32) {
    if (vec4_ok<float>(g.C, dY, X, dX))
      pool_bwd_kernel<float, 4><<<blocks_for(pixels * g.C / 4), 256, 0, st>>>(g, (const float*)dY, (const float*)X, (float*)dX);
    else
      pool_bwd_kernel<float, 1><<<blocks_for(pixels * g.C), 256, 0, st>>>(g, (const float*)dY, (const float*)X, (float*)dX);
  } else {
    typedef __nv_bfloat16 B;
    if (vec4_ok<B>(g.C, dY, X, dX))
      pool_bwd_kernel<B, 4><<<blocks_for(pixels * g.C / 4), 256, 0, st>>>(g, (const B*)dY, (const B*)X, (B*)dX);
    else
      pool_bwd_kernel<B, 1><<<blocks_for(pixels * g.C), 256, 0, st>>>(g, (const B*)dY, (const B*)X, (B*)dX);
  }
  NPML_LAUNCH_OK();
  return 0;
}
