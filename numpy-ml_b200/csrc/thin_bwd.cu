// One-launch backward of a thin stride-1 "same" convolution (cfg1: 28x28x1 -> 32, N = 32): dX, dW and db of
// Conv2D.backward (layers/layers.py:3093-3116) from ONE pass over dZ.
//
// Why.  On such a shape every kernel is a chain of a few memory latencies (10 MB of traffic per step); the step was five
// dependent graph nodes -- memset, fprop, wgrad, fixed-order reduce, dgrad -- of 8 us each (0.014 of HBM).  Here a block
// owns TR output rows of one image, stages its dZ rows (+ halo) and X rows (+ halo) in shared memory ONCE and produces
//   dX[iy, ix, ci]   = sum_{r,t,co} dZ[iy + pr - r D, ix + pc - t D, co] W[r,t,ci,co]      (its own rows)
//   part[blk][m][co] = sum_{pixels of the tile} X[y + r D - pr, x + t D - pc, ci] dZ[y, x, co]   (m = tap*CI + ci; one
//                      more row of ones gives db)
// A block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps its dW / db partial in registers; the grid is
// sized so that every block is resident, the blocks meet at ONE grid barrier (grid_barrier.cuh) and then each adds its
// own slice of the outputs over all partials in a FIXED order (no atomics on data, no second launch, no block that runs
// alone at the end: the first version handed the reduction to "the last block to finish" through two levels of tickets
// and spent 40 % of its 27 us in that serial tail).  Two launches of THIS kernel must not overlap in time.
#include "grid_barrier.cuh"
#include "npml_common.cuh"

namespace npml {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxOutPerThread = 4;         // (taps*CI + 1) * CO <= 4 * 256 outputs of the wgrad partial
__device__ GridBar g_bar;

struct ThinBwd {
  int N, H, W, CI, CO;            // X (N,H,W,CI), dZ (N,H,W,CO): stride 1, output size == input size
  int fr, fc, D, pr, pc;
  int TR, tiles_y;                // output rows per block, blocks per image
  int GR, GW, hb, hl;             // dZ tile rows / columns with halo; rows above / columns left of the tile
  int XR, XW;                     // X tile rows / columns with halo
  int M, M1;                      // taps*CI, M + 1 (ones row)
  int ntiles;
  int64_t w_tap, w_ci, w_co;      // strides of W and dW (the reference layout (fr, fc, in_ch, out_ch))
  int accumulate;
};

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
};
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, int CI_T>
__global__ void __launch_bounds__(kThreads, 3) thin_bwd_kernel(const ThinBwd P, const T* __restrict__ X, const T* __restrict__ dZ,
                                                            const float* __restrict__ Wt, T* __restrict__ dX,
                                                            float* __restrict__ dW, float* __restrict__ db,
                                                            float* __restrict__ part) {
  extern __shared__ __align__(16) float sm[];
  const int CS = P.CO + 4;                                   // padded pixel stride of the dZ tile (bank spread)
  float* Gs = sm;                                            // [GR][GW][CS]
  float* Xs = Gs + (size_t)P.GR * P.GW * CS;                 // [XR][XW][CI]
  float* Ws = Xs + (((size_t)P.XR * P.XW * P.CI + 3) & ~(size_t)3);   // [tap][ci][co]
  __shared__ float s_red[kThreads / 32];
  const int tid = threadIdx.x;
  const int taps = P.fr * P.fc, cog = P.CO >> 2;
  const int total = P.M1 * P.CO;
  float wacc[kMaxOutPerThread];                              // this block's dW / db partial, outputs tid + 256 k
#pragma unroll
  for (int k = 0; k < kMaxOutPerThread; ++k) wacc[k] = 0.f;
  const int cog_sh = 31 - __clz(cog);                        // cog is a power of two
  for (int i = tid; i < taps * P.CI * P.CO; i += kThreads) Ws[i] = Wt[i];        // (fr, fc, CI, CO) is already [tap][ci][co]

  for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
  const int n = tile / P.tiles_y, ty = tile - n * P.tiles_y;
  const int y0 = ty * P.TR;
  const int rows = min(P.TR, P.H - y0);
  __syncthreads();                                           // the previous tile's readers are done with Gs / Xs

  // ---- stage: weights, X rows [y0 - pr, ...) and dZ rows [y0 - hb, ...), zero outside the image -----------------
  // (row loops outside, a thread keeps its column: no integer division anywhere -- the first version spent 4 400
  // instructions per warp, most of them on `i % CO`-style index arithmetic)
  {
    const int xw = P.XW * P.CI;                              // one X row of the tile, channels innermost
    for (int yy = 0; yy < P.XR; ++yy) {
      const int iy = y0 - P.pr + yy;
      const bool row_ok = iy >= 0 && iy < P.H;
      const T* src = X + (((int64_t)n * P.H + iy) * P.W - P.pc) * P.CI;
      for (int j = tid; j < xw; j += kThreads) {
        const int e = j - P.pc * P.CI;                       // element of the image row
        Xs[yy * xw + j] = (row_ok && e >= 0 && e < P.W * P.CI) ? to_f(src[j]) : 0.f;
      }
    }
  }
  {
    const int c4 = tid & (cog - 1), px0 = tid >> cog_sh, ppb = kThreads >> cog_sh;
    for (int yy0 = 0; yy0 < P.GR; yy0 += 4) {                // four rows' loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int yy = yy0 + u, gy = y0 - P.hb + yy;
        const bool row_ok = yy < P.GR && gy >= 0 && gy < P.H;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gx = px0 - P.hl;
        if (row_ok && px0 < P.GW && gx >= 0 && gx < P.W)
          v[u] = Vec4<T>::ld(dZ + (((int64_t)n * P.H + gy) * P.W + gx) * P.CO + 4 * c4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (yy0 + u < P.GR && px0 < P.GW)
          *reinterpret_cast<float4*>(Gs + ((yy0 + u) * P.GW + px0) * CS + 4 * c4) = v[u];
      // columns beyond the first ppb pixels of the row (wide images)
      for (int xx = px0 + ppb; xx < P.GW; xx += ppb) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int yy = yy0 + u, gy = y0 - P.hb + yy;
          const int gx = xx - P.hl;
          float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (yy < P.GR && gy >= 0 && gy < P.H && gx >= 0 && gx < P.W)
            w = Vec4<T>::ld(dZ + (((int64_t)n * P.H + gy) * P.W + gx) * P.CO + 4 * c4);
          if (yy < P.GR) *reinterpret_cast<float4*>(Gs + (yy * P.GW + xx) * CS + 4 * c4) = w;
        }
      }
    }
  }
  __syncthreads();

  // ---- dX of the tile: cog lanes per pixel, lane l owns output channels 4l .. 4l+3 of every tap ------------------
  {
    const int l = tid & (cog - 1), px0 = tid >> cog_sh, ppb = kThreads >> cog_sh;   // pixels per pass
    const int xpasses = (P.W + ppb - 1) / ppb;
    for (int yy = 0; yy < rows; ++yy)
      for (int xp = 0; xp < xpasses; ++xp) {
        const int xx = px0 + xp * ppb;
        const bool live = xx < P.W;
        const int xc = live ? xx : 0;
        float4 a4[CI_T];                                        // four independent chains per input channel
#pragma unroll
        for (int ci = 0; ci < CI_T; ++ci) a4[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P.fr; ++r) {
          // dZ[iy + pr - r D][ix + pc - t D]; tile row = that - (y0 - hb), tile column = that + hl
          const float* grow = Gs + ((yy + (P.fr - 1 - r) * P.D) * P.GW + xc + (P.fc - 1) * P.D) * CS + 4 * l;
          const float* w = Ws + (r * P.fc) * P.CI * P.CO + 4 * l;
          for (int t = 0; t < P.fc; ++t, grow -= P.D * CS, w += P.CI * P.CO) {
            const float4 g4 = *reinterpret_cast<const float4*>(grow);
#pragma unroll
            for (int ci = 0; ci < CI_T; ++ci) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + ci * P.CO);
                a4[ci].x = fmaf(g4.x, w4.x, a4[ci].x); a4[ci].y = fmaf(g4.y, w4.y, a4[ci].y);
                a4[ci].z = fmaf(g4.z, w4.z, a4[ci].z); a4[ci].w = fmaf(g4.w, w4.w, a4[ci].w);
              }
          }
        }
        float acc[CI_T];
#pragma unroll
        for (int ci = 0; ci < CI_T; ++ci) acc[ci] = (a4[ci].x + a4[ci].y) + (a4[ci].z + a4[ci].w);
        for (int o = cog >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int ci = 0; ci < CI_T; ++ci) acc[ci] += __shfl_xor_sync(0xffffffffu, acc[ci], o);
        }
        if (live && l == 0) {
          T* o = dX + (((int64_t)n * P.H + y0 + yy) * P.W + xx) * P.CI;
#pragma unroll
          for (int ci = 0; ci < CI_T; ++ci) from_f(o + ci, acc[ci]);
        }
      }
  }

  // ---- dW / db partial of the tile: thread = (m, co), m = M is the row of ones ----------------------------------
#pragma unroll
  for (int k = 0; k < kMaxOutPerThread; ++k) {
    const int i = tid + k * kThreads;
    if (i >= total) break;
    const int m = i / P.CO, co = i - m * P.CO;
    float a0 = 0.f, a1 = 0.f;                                // two chains: the loop is FMA latency otherwise
    const float* gs = Gs + (P.hb * P.GW + P.hl) * CS + co;
    const int gstep = P.GW * CS, xstep = P.XW * P.CI, ci_st = P.CI, W2 = P.W & ~1;
    if (m < P.M) {
      const int tp = m / P.CI, ci = m - tp * P.CI;
      const int r = tp / P.fc, t = tp - r * P.fc;
      const float* xs = Xs + (r * P.D * P.XW + t * P.D) * P.CI + ci;
      for (int yy = 0; yy < rows; ++yy, gs += gstep, xs += xstep) {
        const float *gp = gs, *xq = xs;
#pragma unroll 4
        for (int xx = 0; xx < W2; xx += 2, gp += 2 * CS, xq += 2 * ci_st) {
          a0 = fmaf(xq[0], gp[0], a0);
          a1 = fmaf(xq[ci_st], gp[CS], a1);
        }
        if (P.W & 1) a0 = fmaf(xq[0], gp[0], a0);
      }
    } else {
      for (int yy = 0; yy < rows; ++yy, gs += gstep) {
        const float* gp = gs;
#pragma unroll 4
        for (int xx = 0; xx < W2; xx += 2, gp += 2 * CS) { a0 += gp[0]; a1 += gp[CS]; }
        if (P.W & 1) a0 += gp[0];
      }
    }
    wacc[k] += a0 + a1;
  }
  }  // tiles

  // ---- fixed-order reduction: partials out, one grid barrier, every block adds its slice of the outputs ------------
  {
    float* mypart = part + (size_t)blockIdx.x * total;
#pragma unroll
    for (int k = 0; k < kMaxOutPerThread; ++k) {
      const int i = tid + k * kThreads;
      if (i < total) mypart[i] = wacc[k];
    }
  }
  grid_barrier(&g_bar, gridDim.x);
  const int nb = (int)gridDim.x, warp = tid >> 5, lane = tid & 31;
  for (int i = blockIdx.x; i < total; i += nb) {
    // thread t adds the partials t, t + 256, ... of output i (all loads in flight), then a fixed-order block sum
    float s = 0.f;
    for (int p0 = tid; p0 < nb; p0 += 4 * kThreads) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = p0 + u * kThreads < nb ? __ldcg(part + (size_t)(p0 + u * kThreads) * total + i) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) s += v[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
      const int m = i / P.CO, co = i - m * P.CO;
      if (m < P.M) {
        const int tp = m / P.CI, ci = m - tp * P.CI;
        const int64_t o = (int64_t)tp * P.w_tap + ci * P.w_ci + co * P.w_co;
        dW[o] = P.accumulate ? dW[o] + t : t;
      } else if (db != nullptr) {
        db[co] = P.accumulate ? db[co] + t : t;
      }
    }
  }
}

constexpr size_t kSmemMax = 96 * 1024;

// the instantiation for (storage type, in_ch) as an untyped kernel pointer (attributes, occupancy, launch)
const void* kernel_of(bool bf16, int ci) {
#define NPML_TB(T)                                                                              \
  switch (ci) {                                                                                 \
    case 1: return (const void*)thin_bwd_kernel<T, 1>;                                          \
    case 2: return (const void*)thin_bwd_kernel<T, 2>;                                          \
    case 3: return (const void*)thin_bwd_kernel<T, 3>;                                          \
    default: return (const void*)thin_bwd_kernel<T, 4>;                                         \
  }
  if (bf16) { NPML_TB(__nv_bfloat16) }
  NPML_TB(float)
#undef NPML_TB
}

struct ThinBwdPlan {
  ThinBwd P;
  unsigned blocks = 0;
  size_t smem = 0, part_floats = 0;
};

bool plan_thin_bwd(const npml_conv_desc* d, ThinBwdPlan& pl) {
  if (getenv("NPML_DISABLE_THIN_BWD")) return false;
  if (d->stride != 1 || d->OH != d->H || d->OW != d->W || d->N < 1) return false;
  if (d->pr1 > (d->fr - 1) * d->dil || d->pc1 > (d->fc - 1) * d->dil) return false;
  if (d->C > 4 || d->K > 64 || d->K % 4) return false;
  const int cog = d->K / 4;
  if (cog & (cog - 1)) return false;                     // lanes per pixel must divide a warp
  if (d->fr * d->fc > 25) return false;
  ThinBwd& P = pl.P;
  P.N = d->N; P.H = d->H; P.W = d->W; P.CI = d->C; P.CO = d->K;
  P.fr = d->fr; P.fc = d->fc; P.D = d->dil; P.pr = d->pr1; P.pc = d->pc1;
  P.M = d->fr * d->fc * d->C; P.M1 = P.M + 1;
  if ((int64_t)P.M1 * P.CO > kMaxOutPerThread * kThreads) return false;
  P.w_tap = (int64_t)d->C * d->K; P.w_ci = d->K; P.w_co = 1;
  P.accumulate = d->accumulate;
  P.hb = (d->fr - 1) * d->dil - d->pr1;                  // rows of dZ above the tile that its dX rows read
  P.hl = (d->fc - 1) * d->dil - d->pc1;
  const int halo_y = (d->fr - 1) * d->dil, halo_x = (d->fc - 1) * d->dil;
  P.GW = d->W + halo_x; P.XW = d->W + halo_x;
  const int64_t rows_total = (int64_t)d->N * d->H;
  static bool attr_set = false;
  if (!attr_set) {                       // (fails harmlessly on a box without a GPU: the queries below then say 1)
    for (int b = 0; b < 2; ++b)
      for (int c = 1; c <= 4; ++c) cudaFuncSetAttribute(kernel_of(b != 0, c), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    cudaGetLastError();
    attr_set = true;
  }
  // Every block must be resident (grid barrier) and the kernel is a chain of latencies, so: as many blocks as fit (at
  // most four per SM), one tile each when possible.  TR = rows per tile follows from the number of blocks; shared memory
  // depends on TR and the number of resident blocks on shared memory, hence two rounds.
  int per_sm = 4;
  for (int round = 0; round < 3; ++round) {
    const int64_t want = (int64_t)per_sm * num_sms();
    int TR = (int)((rows_total + want - 1) / want);
    if (TR < 1) TR = 1;
    if (TR > d->H) TR = d->H;
    for (;; --TR) {
      if (TR < 1) return false;
      P.TR = TR; P.GR = TR + halo_y; P.XR = TR + halo_y;
      const size_t fl = (size_t)P.GR * P.GW * (P.CO + 4) + (((size_t)P.XR * P.XW * P.CI + 3) & ~(size_t)3) +
                        (size_t)d->fr * d->fc * d->C * d->K;
      pl.smem = fl * sizeof(float);
      if (pl.smem <= kSmemMax) break;
    }
    int fit = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, kernel_of(d->mode == NPML_MODE_BF16, d->C), kThreads, pl.smem);
    if (e != cudaSuccess || fit < 1) { cudaGetLastError(); fit = 1; }
    if (fit >= per_sm) break;
    per_sm = fit;
  }
  P.tiles_y = (d->H + P.TR - 1) / P.TR;
  const int64_t tiles = (int64_t)d->N * P.tiles_y;
  if (tiles >= (1LL << 30)) return false;
  P.ntiles = (int)tiles;
  const int64_t cap = (int64_t)per_sm * num_sms();
  pl.blocks = (unsigned)(tiles < cap ? tiles : cap);
  pl.part_floats = (size_t)pl.blocks * P.M1 * P.CO;
  return true;
}

}  // namespace

bool thin_bwd_supported(const npml_conv_desc* d) {
  ThinBwdPlan pl;
  return plan_thin_bwd(d, pl);
}

size_t thin_bwd_ws_bytes(const npml_conv_desc* d) {
  ThinBwdPlan pl;
  if (!plan_thin_bwd(d, pl)) return 0;
  return align_up(pl.part_floats * sizeof(float), 256);
}

int thin_bwd(const npml_conv_desc* d, const void* X, const void* dZ, const float* W, float* dW, float* db, void* dX,
             void* ws, size_t ws_bytes, cudaStream_t st) {
  ThinBwdPlan pl;
  NPML_CHECK(plan_thin_bwd(d, pl), "shape not supported by the fused thin backward");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.part_floats * sizeof(float), "workspace too small");
  float* part = (float*)ws;
  const bool bf16 = d->mode == NPML_MODE_BF16;
  NPML_CHECK((reinterpret_cast<uintptr_t>(dZ) & (bf16 ? 7 : 15)) == 0, "dZ must be 16-byte aligned");
  void* args[] = {(void*)&pl.P, (void*)&X, (void*)&dZ, (void*)&W, (void*)&dX, (void*)&dW, (void*)&db, (void*)&part};
  NPML_CUDA(cudaLaunchKernel(kernel_of(bf16, d->C), dim3(pl.blocks), dim3(kThreads), args, pl.smem, st));
  NPML_LAUNCH_OK();
  return 0;
}

}  // namespace npml
