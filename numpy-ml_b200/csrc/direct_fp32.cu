// CUDA-core implicit-GEMM convolution kernels with fp32 accumulation: the exactness mode
// (NPML_MODE_FP32) and the path for shapes the tensor-core kernels do not take (in_ch = 1, odd
// channel counts).  No im2col is materialised: the gather of utils/utils.py:541 and the
// scatter-add of utils/utils.py:591 are folded into the tile loads.
#include "npml_common.cuh"

namespace npml {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256, PADM = 4;

// ---- fprop / dgrad -------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(NT) gather_conv_kernel(GatherConv g, const TIn* __restrict__ in,
                                                         const float* __restrict__ w,
                                                         const float* __restrict__ bias, TOut* __restrict__ Z,
                                                         TOut* __restrict__ Y) {
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Ktot = g.fr * g.fc * g.CI;

  // A-load mapping: one k column, four pixel rows per thread
  const int a_k = tid % BK;
  const int a_r = tid / BK;  // 0..15
  int pn[4], py[4], px[4];
  bool pv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t p = m0 + a_r + 16 * j;
    pv[j] = p < P;
    int64_t q = pv[j] ? p : 0;
    px[j] = (int)(q % g.OW);
    q /= g.OW;
    py[j] = (int)(q % g.OH);
    pn[j] = (int)(q / g.OH);
  }
  // B-load mapping: one out channel, four k rows per thread
  const int b_c = tid % BN;
  const int b_r = tid / BN;  // 0..3

  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
    {
      const int kk = k0 + a_k;
      const bool kv = kk < Ktot;
      const int tap = kv ? kk / g.CI : 0;
      const int ci = kv ? kk - tap * g.CI : 0;
      const int r = tap / g.fc, t = tap - r * g.fc;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (kv && pv[j]) {
          int iy, ix;
          bool ok = true;
          if (g.form == 0) {
            iy = py[j] * g.s + r * g.D - g.pr;
            ix = px[j] * g.s + t * g.D - g.pc;
          } else {
            int ny = py[j] + g.pr - r * g.D, nx = px[j] + g.pc - t * g.D;
            ok = ny >= 0 && nx >= 0 && (ny % g.s) == 0 && (nx % g.s) == 0;
            iy = ny / g.s;
            ix = nx / g.s;
          }
          if (ok && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW)
            v = ld_f(in, (((int64_t)pn[j] * g.IH + iy) * g.IW + ix) * g.CI + ci);
        }
        As[a_k][a_r + 16 * j] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kl = b_r + 4 * j;
      const int kk = k0 + kl;
      const int co = n0 + b_c;
      float v = 0.f;
      if (kk < Ktot && co < g.CO) {
        const int tap = kk / g.CI;
        const int ci = kk - tap * g.CI;
        v = w[tap * g.w_tap + ci * g.w_ci + co * g.w_co];
      }
      Bs[kl][b_c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t p = m0 + ty * 4 + i;
    if (p >= P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= g.CO) continue;
      float z = acc[i][j] + (bias ? bias[co] : 0.f);
      if (Z) st_f(Z, p * g.CO + co, z);
      if (Y) st_f(Y, p * g.CO + co, act_fn(g.act, z, g.p0, g.p1));
    }
  }
}

// ---- wgrad: split over pixel chunks, partials to workspace -----------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) wgrad_partial_kernel(GatherWgrad g, const T* __restrict__ in,
                                                           const T* __restrict__ grad, float* __restrict__ part,
                                                           int64_t chunk) {
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int M = g.fr * g.fc * g.CI;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  const int64_t p_end = (p_begin + chunk < P) ? p_begin + chunk : P;

  // A mapping: one (tap, ci) column per thread, four pixel rows
  const int a_m = tid % BM, a_r = tid / BM;  // a_r 0..3
  const int kk = m0 + a_m;
  const bool mv = kk < M;
  const int tap = mv ? kk / g.CI : 0;
  const int ci = mv ? kk - tap * g.CI : 0;
  const int r = tap / g.fc, t = tap - r * g.fc;
  const int b_c = tid % BN, b_r = tid / BN;
  const int co_l = n0 + b_c;

  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kl = a_r + 4 * j;
      const int64_t p = p0 + kl;
      float va = 0.f, vb = 0.f;
      if (p < p_end) {
        int64_t q = p;
        const int x = (int)(q % g.OW);
        q /= g.OW;
        const int y = (int)(q % g.OH);
        const int n = (int)(q / g.OH);
        if (mv) {
          const int iy = y * g.s + r * g.D - g.pr, ix = x * g.s + t * g.D - g.pc;
          if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW)
            va = ld_f(in, (((int64_t)n * g.IH + iy) * g.IW + ix) * g.CI + ci);
        }
        if (co_l < g.CO) vb = ld_f(grad, p * g.CO + co_l);
      }
      As[kl][a_m] = va;
      Bs[kl][b_c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = part + (int64_t)blockIdx.z * M * g.CO;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co < g.CO) out[(int64_t)m * g.CO + co] = acc[i][j];
    }
  }
}

// ---- thin shapes (cfg1: in_ch = 1): the 64x64 tiles above would be almost empty ---------------------------
// gather conv with CO <= 4 (Conv2D dX when in_ch is tiny): one THREAD per output pixel (the index arithmetic is paid
// once per pixel, not once per lane), the taps' channel vectors are read with 16-byte loads, the weights sit in
// shared memory (every thread reads the same address: broadcast).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128) thin_co_conv_kernel(GatherConv g, const TIn* __restrict__ in,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           TOut* __restrict__ Z, TOut* __restrict__ Y, int w_in_smem) {
  extern __shared__ float ws_[];            // [tap][ci][co] when it fits
  const int taps = g.fr * g.fc;
  if (w_in_smem) {
    for (int i = threadIdx.x; i < taps * g.CI * g.CO; i += blockDim.x) {
      const int co = i % g.CO, ci = (i / g.CO) % g.CI, tp = i / (g.CO * g.CI);
      ws_[i] = w[(int64_t)tp * g.w_tap + ci * g.w_ci + co * g.w_co];
    }
    __syncthreads();
  }
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int pi = (int)p;
  const int x = pi % g.OW, y = (pi / g.OW) % g.OH, n = pi / (g.OW * g.OH);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int V = 16 / (int)sizeof(TIn);                 // elements per 16-byte load
  const bool vec = (g.CI % V) == 0;
  for (int r = 0; r < g.fr; ++r) {
    int iy;
    if (g.form == 0) iy = y * g.s + r * g.D - g.pr;
    else { const int ny = y + g.pr - r * g.D; if (ny < 0 || ny % g.s) continue; iy = ny / g.s; }
    if (iy < 0 || iy >= g.IH) continue;
    for (int t = 0; t < g.fc; ++t) {
      int ix;
      if (g.form == 0) ix = x * g.s + t * g.D - g.pc;
      else { const int nx = x + g.pc - t * g.D; if (nx < 0 || nx % g.s) continue; ix = nx / g.s; }
      if (ix < 0 || ix >= g.IW) continue;
      const int64_t base = (((int64_t)n * g.IH + iy) * g.IW + ix) * g.CI;
      const int tp = r * g.fc + t;
      if (vec) {
        for (int c0 = 0; c0 < g.CI; c0 += V) {
          float v[V];
          const uint4 u = *reinterpret_cast<const uint4*>(in + base + c0);
          if constexpr (sizeof(TIn) == 4) {
            v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
          } else {
            const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&q[j]);
              v[2 * j] = __low2float(h); v[2 * j + 1] = __high2float(h);
            }
          }
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const int ci = c0 + j;
#pragma unroll
            for (int co = 0; co < 4; ++co)
              if (co < g.CO) {
                const float wv = w_in_smem ? ws_[(tp * g.CI + ci) * g.CO + co]
                                           : __ldg(w + (int64_t)tp * g.w_tap + ci * g.w_ci + co * g.w_co);
                acc[co] = fmaf(v[j], wv, acc[co]);
              }
          }
        }
      } else {
        for (int ci = 0; ci < g.CI; ++ci) {
          const float v = ld_f(in, base + ci);
#pragma unroll
          for (int co = 0; co < 4; ++co)
            if (co < g.CO) {
              const float wv = w_in_smem ? ws_[(tp * g.CI + ci) * g.CO + co]
                                         : __ldg(w + (int64_t)tp * g.w_tap + ci * g.w_ci + co * g.w_co);
              acc[co] = fmaf(v, wv, acc[co]);
            }
        }
      }
    }
  }
  for (int co = 0; co < g.CO; ++co) {
    const float z = acc[co] + (bias ? bias[co] : 0.f);
    if (Z) st_f(Z, p * g.CO + co, z);
    if (Y) st_f(Y, p * g.CO + co, act_fn(g.act, z, g.p0, g.p1));
  }
}

// wgrad with (taps*CI + 1) * CO <= 1024: one thread per (row m, out channel), a block per pixel chunk; the extra row
// m = M has the constant input 1 and yields db.  Partials [chunk][M][CO] (+ [chunk][CO]) go to the fixed-order reduce.
template <typename T>
__global__ void __launch_bounds__(1024) thin_m_wgrad_kernel(GatherWgrad g, const T* __restrict__ in,
                                                            const T* __restrict__ grad, float* __restrict__ part,
                                                            float* __restrict__ part_db, int64_t chunk) {
  const int M = g.fr * g.fc * g.CI;
  const int m = threadIdx.x / g.CO, co = threadIdx.x - m * g.CO;
  if (m > M) return;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int64_t p_begin = (int64_t)blockIdx.x * chunk;
  const int64_t p_end = (p_begin + chunk < P) ? p_begin + chunk : P;
  const int tap = m < M ? m / g.CI : 0, ci = m < M ? m - tap * g.CI : 0;
  const int r = tap / g.fc, t = tap - r * g.fc;
  float acc = 0.f;
  const int q0 = (int)p_begin;
  int x = q0 % g.OW, y = (q0 / g.OW) % g.OH, n = q0 / (g.OW * g.OH);
#pragma unroll 4
  for (int64_t p = p_begin; p < p_end; ++p) {
    float a = 1.f;
    if (m < M) {
      const int iy = y * g.s + r * g.D - g.pr, ix = x * g.s + t * g.D - g.pc;
      a = (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) ? ld_f(in, (((int64_t)n * g.IH + iy) * g.IW + ix) * g.CI + ci) : 0.f;
    }
    acc = fmaf(a, ld_f(grad, p * g.CO + co), acc);
    if (++x == g.OW) { x = 0; if (++y == g.OH) { y = 0; ++n; } }
  }
  if (m < M) part[((int64_t)blockIdx.x * M + m) * g.CO + co] = acc;
  else part_db[(int64_t)blockIdx.x * g.CO + co] = acc;
}

}  // namespace

// Fixed-order reduction of split partials [splits][M][CO] -> dW (strided), shared with the tensor-core wgrad
// kernels.  When part_db != nullptr the same launch also reduces [splits][CO] -> db.  Block = (64 outputs, 8 split
// groups): thread (x, y) adds the splits z = y, y + 8, ... of output x (reads coalesced along x), the 8 group sums
// are then added in a fixed order -- deterministic, and 8x the memory-level parallelism of one thread per output.
__global__ void __launch_bounds__(512) wgrad_reduce_kernel(const float* __restrict__ part, int splits, int M, int CO,
                                                           int CI, int64_t o_tap, int64_t o_ci, int64_t o_co,
                                                           int accumulate, float* __restrict__ dW,
                                                           const float* __restrict__ part_db, float* __restrict__ db) {
  __shared__ float red[8][64];
  const int64_t total = (int64_t)M * CO;
  const int64_t total_all = total + (part_db != nullptr ? CO : 0);
  for (int64_t i0 = (int64_t)blockIdx.x * 64; i0 < total_all; i0 += (int64_t)gridDim.x * 64) {
    const int64_t i = i0 + threadIdx.x;
    float s = 0.f;
    if (i < total) {
      for (int z = threadIdx.y; z < splits; z += 8) s += part[(int64_t)z * total + i];
    } else if (i < total_all) {
      for (int z = threadIdx.y; z < splits; z += 8) s += part_db[(int64_t)z * CO + (i - total)];
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && i < total_all) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += red[y][threadIdx.x];
      if (i < total) {
        const int m = (int)(i / CO), co = (int)(i - (int64_t)m * CO);
        const int tap = m / CI, ci = m - tap * CI;
        const int64_t o = tap * o_tap + ci * o_ci + co * o_co;
        dW[o] = accumulate ? dW[o] + t : t;
      } else {
        const int co = (int)(i - total);
        db[co] = accumulate ? db[co] + t : t;
      }
    }
    __syncthreads();
  }
}

// few splits: one thread per output element is enough
__global__ void wgrad_reduce_simple_kernel(const float* __restrict__ part, int splits, int M, int CO, int CI,
                                           int64_t o_tap, int64_t o_ci, int64_t o_co, int accumulate,
                                           float* __restrict__ dW, const float* __restrict__ part_db, float* __restrict__ db) {
  const int64_t total = (int64_t)M * CO;
  const int64_t total_all = total + (part_db != nullptr ? CO : 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_all;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < total) {
      for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + i];
      const int m = (int)(i / CO), co = (int)(i - (int64_t)m * CO);
      const int tap = m / CI, ci = m - tap * CI;
      const int64_t o = tap * o_tap + ci * o_ci + co * o_co;
      dW[o] = accumulate ? dW[o] + s : s;
    } else {
      const int co = (int)(i - total);
      for (int z = 0; z < splits; ++z) s += part_db[(int64_t)z * CO + co];
      db[co] = accumulate ? db[co] + s : s;
    }
  }
}

// Few outputs, many splits (thin shapes): one warp per output element, lanes stride over the splits, fixed-order
// shuffle tree (deterministic).
__global__ void wgrad_reduce_warp_kernel(const float* __restrict__ part, int splits, int M, int CO, int CI,
                                         int64_t o_tap, int64_t o_ci, int64_t o_co, int accumulate,
                                         float* __restrict__ dW, const float* __restrict__ part_db, float* __restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)M * CO;
  const int64_t total_all = total + (part_db != nullptr ? CO : 0);
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= total_all) return;
  float s = 0.f;
  if (i < total) for (int z = lane; z < splits; z += 32) s += part[(int64_t)z * total + i];
  else for (int z = lane; z < splits; z += 32) s += part_db[(int64_t)z * CO + (i - total)];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  if (i < total) {
    const int m = (int)(i / CO), co = (int)(i - (int64_t)m * CO);
    const int tap = m / CI, ci = m - tap * CI;
    const int64_t o = tap * o_tap + ci * o_ci + co * o_co;
    dW[o] = accumulate ? dW[o] + s : s;
  } else {
    const int co = (int)(i - total);
    db[co] = accumulate ? db[co] + s : s;
  }
}

// Same reduction when the output's fastest dimension is ci (Deconv2D: dW[r,t,c,k] with the roles of c and k
// swapped): a 32x32 (ci, co) tile is summed with reads coalesced along co, transposed through shared memory and
// written coalesced along ci.
__global__ void wgrad_reduce_t_kernel(const float* __restrict__ part, int splits, int taps, int CO, int CI,
                                      int64_t o_tap, int64_t o_ci, int64_t o_co, int accumulate,
                                      float* __restrict__ dW) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const int64_t total = (int64_t)taps * CI * CO;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    float s = 0.f;
    if (ci < CI && co < CO) {
      const int64_t i = ((int64_t)tap * CI + ci) * CO + co;
      for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + i];
    }
    tile[r][threadIdx.x] = s;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (ci < CI && co < CO) {
      const int64_t o = tap * o_tap + ci * o_ci + co * o_co;
      const float s = tile[threadIdx.x][r];
      dW[o] = accumulate ? dW[o] + s : s;
    }
  }
}

int launch_wgrad_reduce(const float* part, int splits, int M, int CO, int CI, int64_t o_tap, int64_t o_ci,
                        int64_t o_co, int accumulate, float* dW, cudaStream_t st, const float* part_db = nullptr,
                        float* db = nullptr) {
  if (o_ci == 1 && o_co != 1 && part_db == nullptr) {
    const int taps = M / CI;
    dim3 grid((unsigned)((CO + 31) / 32), (unsigned)((CI + 31) / 32), (unsigned)taps);
    wgrad_reduce_t_kernel<<<grid, dim3(32, 8), 0, st>>>(part, splits, taps, CO, CI, o_tap, o_ci, o_co, accumulate, dW);
    NPML_LAUNCH_OK();
    return 0;
  }
  const int64_t total = (int64_t)M * CO + (part_db != nullptr ? CO : 0);
  if (total <= 8192 && splits >= 64) {
    wgrad_reduce_warp_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(part, splits, M, CO, CI, o_tap, o_ci, o_co,
                                                                         accumulate, dW, part_db, db);
    NPML_LAUNCH_OK();
    return 0;
  }
  if (splits < 32) {
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    wgrad_reduce_simple_kernel<<<blocks, 256, 0, st>>>(part, splits, M, CO, CI, o_tap, o_ci, o_co, accumulate, dW, part_db,
                                                       db);
    NPML_LAUNCH_OK();
    return 0;
  }
  int blocks = (int)((total + 63) / 64);
  if (blocks > 8192) blocks = 8192;
  wgrad_reduce_kernel<<<blocks, dim3(64, 8), 0, st>>>(part, splits, M, CO, CI, o_tap, o_ci, o_co, accumulate, dW, part_db, db);
  NPML_LAUNCH_OK();
  return 0;
}

int direct_gather_conv(const GatherConv& g, const void* in, const float* w, const float* bias, void* Z,
                       void* Y, int dtype, cudaStream_t st) {
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  if (P == 0 || g.CO == 0) return 0;
  if (g.CO <= 4 && P < (1LL << 31) && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    const unsigned blocks = (unsigned)((P + 127) / 128);
    const size_t wn = (size_t)g.fr * g.fc * g.CI * g.CO * sizeof(float);
    const int w_smem = wn <= 40 * 1024;
    if (dtype == NPML_F32)
      thin_co_conv_kernel<float, float><<<blocks, 128, w_smem ? wn : 0, st>>>(g, (const float*)in, w, bias, (float*)Z,
                                                                             (float*)Y, w_smem);
    else
      thin_co_conv_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 128, w_smem ? wn : 0, st>>>(
          g, (const __nv_bfloat16*)in, w, bias, (__nv_bfloat16*)Z, (__nv_bfloat16*)Y, w_smem);
    NPML_LAUNCH_OK();
    return 0;
  }
  dim3 grid((unsigned)((P + BM - 1) / BM), (unsigned)((g.CO + BN - 1) / BN));
  if (dtype == NPML_F32)
    gather_conv_kernel<float, float><<<grid, NT, 0, st>>>(g, (const float*)in, w, bias, (float*)Z, (float*)Y);
  else
    gather_conv_kernel<__nv_bfloat16, __nv_bfloat16>
        <<<grid, NT, 0, st>>>(g, (const __nv_bfloat16*)in, w, bias, (__nv_bfloat16*)Z, (__nv_bfloat16*)Y);
  NPML_LAUNCH_OK();
  return 0;
}

static bool thin_wgrad(const GatherWgrad& g) {
  return (int64_t)(g.fr * g.fc * g.CI + 1) * g.CO <= 1024;
}

static int direct_wgrad_splits(const GatherWgrad& g) {
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int M = g.fr * g.fc * g.CI;
  if (thin_wgrad(g)) {               // one block per pixel chunk; the kernel is latency-bound, so many short chunks
    int64_t splits = 8LL * num_sms();
    if (splits * 16 > P) splits = (P + 15) / 16;
    return (int)(splits < 1 ? 1 : splits);
  }
  const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((g.CO + BN - 1) / BN);
  int64_t splits = (4LL * num_sms() + tiles - 1) / tiles;
  const int64_t by_len = (P + 4095) / 4096;  // keep every sequential fp32 sum short
  if (splits < by_len) splits = by_len;
  const int64_t max_splits = (P + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits > 1024) splits = 1024;
  if (splits < 1) splits = 1;
  return (int)splits;
}

size_t direct_wgrad_ws_bytes(const GatherWgrad& g) {
  const int M = g.fr * g.fc * g.CI;
  return (size_t)direct_wgrad_splits(g) * (M + 1) * g.CO * sizeof(float);   // + one db row per split (thin shapes)
}

// *db_done (when not null) tells the caller whether db was produced here (thin shapes) or still has to be computed
int direct_wgrad(const GatherWgrad& g, const void* in, const void* grad, float* dW, int dtype, void* ws,
                 size_t ws_bytes, cudaStream_t st, float* db, bool* db_done) {
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int M = g.fr * g.fc * g.CI;
  if (db_done) *db_done = false;
  if (M == 0 || g.CO == 0) return 0;
  const int splits = direct_wgrad_splits(g);
  NPML_CHECK(ws_bytes >= direct_wgrad_ws_bytes(g), "workspace too small");
  if (thin_wgrad(g)) {
    const int64_t chunk = (P + splits - 1) / splits;
    const int nsplit = (int)((P + chunk - 1) / chunk);
    float* part = (float*)ws;
    float* part_db = part + (size_t)nsplit * M * g.CO;
    const int threads = (M + 1) * g.CO;
    if (dtype == NPML_F32)
      thin_m_wgrad_kernel<float><<<nsplit, threads, 0, st>>>(g, (const float*)in, (const float*)grad, part, part_db, chunk);
    else
      thin_m_wgrad_kernel<__nv_bfloat16>
          <<<nsplit, threads, 0, st>>>(g, (const __nv_bfloat16*)in, (const __nv_bfloat16*)grad, part, part_db, chunk);
    NPML_LAUNCH_OK();
    if (db_done) *db_done = db != nullptr;
    return launch_wgrad_reduce(part, nsplit, M, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate, dW, st,
                               db != nullptr ? part_db : nullptr, db);
  }
  int64_t chunk = (P + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  dim3 grid((M + BM - 1) / BM, (g.CO + BN - 1) / BN, splits);
  if (dtype == NPML_F32)
    wgrad_partial_kernel<float><<<grid, NT, 0, st>>>(g, (const float*)in, (const float*)grad, (float*)ws, chunk);
  else
    wgrad_partial_kernel<__nv_bfloat16>
        <<<grid, NT, 0, st>>>(g, (const __nv_bfloat16*)in, (const __nv_bfloat16*)grad, (float*)ws, chunk);
  NPML_LAUNCH_OK();
  return launch_wgrad_reduce((const float*)ws, splits, M, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate,
                             dW, st);
}

}  // namespace npml
