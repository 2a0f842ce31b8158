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
      if (Y) st_f(Y, p * g.CO + co, post_affine(g, co, act_fn(g.act, z, g.p0, g.p1)));
    }
  }
}

// ---- fprop / dgrad, register-tiled (fp32 storage, in_ch % 4 == 0, out_ch % 4 == 0) --------------------------------
// The kernel above is a 64 x 64 tile with 4 x 4 outputs per thread, scalar gathers and two barriers per K block:
// 16 TFLOP/s at cfg2 (12.5 ms per fp32-mode step; the exactness mode is the constructor default).  Here a 128 x (16 TN)
// tile, 8 x TN outputs per thread (four 16-byte shared loads per 16 TN FMAs), 16-byte gathers along in_ch with the tap
// walked incrementally (no division in the loop), operands of the next K block in registers while the current one is
// multiplied, and two shared-memory buffers (one barrier per K block).
// One launch covers one output parity class: the outputs (y0 + so yy, x0 + so xx) and the taps that reach them, as
// offsets into the input: iy = yy * sin + dy[j].  The forward form (Conv2D.forward, Deconv2D dX) is ONE class with every
// tap (so = 1, sin = stride); the transposed form (Conv2D dX, Deconv2D.forward) with stride s has s^2 classes with a
// fraction of the taps each and sin = 1 -- walking all taps under a divisibility predicate instead wasted 3/4 of the
// FMAs of a stride-2 problem (fp32 mode: cfg5 forward 4.0 ms, cfg3 dX 2.0 ms against 1.0 / 0.3 ms for their adjoints).
constexpr int kRtMaxTaps = 64;
struct RtClass {
  int ntaps, y0, x0, so, sin, oh, ow;
  int16_t dy[kRtMaxTaps], dx[kRtMaxTaps], wt[kRtMaxTaps];    // input offset and filter tap (index into W) of tap j
};

template <int TN>                                       // output channels per thread: 4 (64-wide tile) or 8 (128-wide)
__global__ void __launch_bounds__(256, 2) gather_conv_rt_kernel(GatherConv g, const RtClass c, const float* __restrict__ in,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             float* __restrict__ Z, float* __restrict__ Y) {
  constexpr int TBM = 128, TBN = 16 * TN, TBK = 16;
  __shared__ __align__(16) float As[2][TBK][TBM];
  __shared__ __align__(16) float Bs[2][TBK][TBN];
  const int tid = threadIdx.x;
  const int64_t P = (int64_t)g.N * c.oh * c.ow;           // output pixels of this class
  const int64_t m0 = (int64_t)blockIdx.x * TBM;
  const int n0 = blockIdx.y * TBN;
  const int Ktot = c.ntaps * g.CI;
  const int nkb = (Ktot + TBK - 1) / TBK;

  // A gather: thread -> pixel am = tid % 128, k groups ag and ag + 2 (four consecutive in_ch each)
  const int am = tid & 127, ag = tid >> 7;
  const int64_t ap = m0 + am;
  const bool apv = ap < P;
  int ax, ay, an;
  {
    int64_t q = apv ? ap : 0;
    ax = (int)(q % c.ow); q /= c.ow;
    ay = (int)(q % c.oh); an = (int)(q / c.oh);
  }
  const float* aimg = in + (int64_t)an * g.IH * g.IW * g.CI;
  // (tap, ci) of this thread's two k groups, advanced by 16 per K block without dividing
  int kci[2], ktap[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int kk = 4 * (ag + 2 * u);
    ktap[u] = kk / g.CI; kci[u] = kk - ktap[u] * g.CI;
  }
  // B load: thread -> k row bk = tid / 16 (16 rows), columns 4 * (tid % 16) + 64 * v
  const int bk = tid >> 4, bc = (tid & 15) * 4;
  int bci = bk % g.CI, btap = bk / g.CI;
  // dgrad reads W[tap][c][k] with the REDUCTION channel k contiguous (w_ci == 1): 16-byte loads along k for one output
  // column, stored transposed like the A operand.  Thread -> column tn = tid % TBN, k groups tg + (256 / TBN) u
  const bool b_trans = g.w_ci == 1 && g.w_co != 1;
  constexpr int kGroupsPerPass = 256 / TBN;               // 4 (64-wide tile) or 2 (128-wide)
  const int tn = tid % TBN, tg = tid / TBN;
  int tci[TN / 4], ttap[TN / 4];
#pragma unroll
  for (int u = 0; u < TN / 4; ++u) {
    const int kk = 4 * (tg + kGroupsPerPass * u);
    ttap[u] = kk / g.CI; tci[u] = kk - ttap[u] * g.CI;
  }

  auto load_a = [&](int u) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (apv && ktap[u] < c.ntaps) {
      const int iy = ay * c.sin + c.dy[ktap[u]], ix = ax * c.sin + c.dx[ktap[u]];
      if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW)
        v = *reinterpret_cast<const float4*>(aimg + ((int64_t)iy * g.IW + ix) * g.CI + kci[u]);
    }
    return v;
  };
  auto advance_a = [&]() {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      kci[u] += TBK;
      while (kci[u] >= g.CI) { kci[u] -= g.CI; ++ktap[u]; }
    }
  };
  auto load_b = [&](float (&bv)[TN]) {
    if (b_trans) {
#pragma unroll
      for (int u = 0; u < TN / 4; ++u) {
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ttap[u] < c.ntaps && n0 + tn < g.CO)
          q = *reinterpret_cast<const float4*>(w + (int64_t)c.wt[ttap[u]] * g.w_tap + (int64_t)(n0 + tn) * g.w_co + tci[u]);
        bv[4 * u] = q.x; bv[4 * u + 1] = q.y; bv[4 * u + 2] = q.z; bv[4 * u + 3] = q.w;
      }
      return;
    }
    const bool kv = btap < c.ntaps;
    const float* src = w + (int64_t)(kv ? c.wt[btap] : 0) * g.w_tap + (int64_t)bci * g.w_ci;
#pragma unroll
    for (int v = 0; v < TN / 4; ++v) {
      const int co = n0 + bc + 64 * v;
      if (g.w_co == 1 && kv && co + 3 < g.CO) {
        const float4 q = *reinterpret_cast<const float4*>(src + co);
        bv[4 * v] = q.x; bv[4 * v + 1] = q.y; bv[4 * v + 2] = q.z; bv[4 * v + 3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) bv[4 * v + e] = (kv && co + e < g.CO) ? src[(int64_t)(co + e) * g.w_co] : 0.f;
      }
    }
  };
  auto advance_b = [&]() {
    bci += TBK;
    while (bci >= g.CI) { bci -= g.CI; ++btap; }
#pragma unroll
    for (int u = 0; u < TN / 4; ++u) {
      tci[u] += TBK;
      while (tci[u] >= g.CI) { tci[u] -= g.CI; ++ttap[u]; }
    }
  };
  auto store_tile = [&](int buf, const float4 (&av)[2], const float (&bv)[TN]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = 4 * (ag + 2 * u);
      As[buf][k][am] = av[u].x; As[buf][k + 1][am] = av[u].y; As[buf][k + 2][am] = av[u].z; As[buf][k + 3][am] = av[u].w;
    }
    if (b_trans) {
#pragma unroll
      for (int u = 0; u < TN / 4; ++u) {
        const int k = 4 * (tg + kGroupsPerPass * u);
        Bs[buf][k][tn] = bv[4 * u]; Bs[buf][k + 1][tn] = bv[4 * u + 1]; Bs[buf][k + 2][tn] = bv[4 * u + 2]; Bs[buf][k + 3][tn] = bv[4 * u + 3];
      }
      return;
    }
#pragma unroll
    for (int v = 0; v < TN / 4; ++v)
      *reinterpret_cast<float4*>(&Bs[buf][bk][bc + 64 * v]) = make_float4(bv[4 * v], bv[4 * v + 1], bv[4 * v + 2], bv[4 * v + 3]);
  };

  const int tx = tid & 15, ty = tid >> 4;                 // outputs: rows 4 ty + {0..3} (+ 64), columns 4 tx + {0..3} (+ 64)
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 av[2];
  float bv[TN];
  av[0] = load_a(0); av[1] = load_a(1);
  load_b(bv);
  store_tile(0, av, bv);
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    const bool more = kb + 1 < nkb;
    if (more) {
      advance_a(); advance_b();
      av[0] = load_a(0); av[1] = load_a(1);
      load_b(bv);
    }
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int v = 0; v < TN / 4; ++v) {
        const float4 q = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64 * v]);
        b[4 * v] = q.x; b[4 * v + 1] = q.y; b[4 * v + 2] = q.z; b[4 * v + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tile(buf ^ 1, av, bv);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t pc = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (pc >= P) continue;
    int64_t p = pc;                                      // class pixel -> pixel of the whole output
    if (c.so != 1 || c.oh != g.OH || c.ow != g.OW) {
      const int xx = (int)(pc % c.ow);
      const int64_t q = pc / c.ow;
      const int yy = (int)(q % c.oh), nn = (int)(q / c.oh);
      p = ((int64_t)nn * g.OH + c.y0 + c.so * yy) * g.OW + c.x0 + c.so * xx;
    }
#pragma unroll
    for (int v = 0; v < TN / 4; ++v) {
      const int co = n0 + tx * 4 + 64 * v;
      if (co >= g.CO) continue;                          // CO % 4 == 0: a group of four is all in or all out
      float z[4], y[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        z[e] = acc[i][4 * v + e] + (bias ? bias[co + e] : 0.f);
        y[e] = Y ? post_affine(g, co + e, act_fn(g.act, z[e], g.p0, g.p1)) : 0.f;
      }
      if (Z) *reinterpret_cast<float4*>(Z + p * g.CO + co) = make_float4(z[0], z[1], z[2], z[3]);
      if (Y) *reinterpret_cast<float4*>(Y + p * g.CO + co) = make_float4(y[0], y[1], y[2], y[3]);
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
// gather conv with CO <= 4 (Conv2D dX when in_ch is tiny).  Four lanes share one output pixel and split the input
// channels of every tap (16-byte loads); the index arithmetic is paid once per pixel, the weights sit in shared
// memory as [tap][ci][co] (broadcast reads), CO is a template parameter so the inner loop is pure FMA.
template <typename TIn, typename TOut, int CO_T>
__global__ void __launch_bounds__(128) thin_co_conv_kernel(GatherConv g, const TIn* __restrict__ in,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           TOut* __restrict__ Z, TOut* __restrict__ Y) {
  extern __shared__ __align__(16) float ws_[];          // [tap][ci][co]
  const int taps = g.fr * g.fc;
  for (int i = threadIdx.x; i < taps * g.CI * CO_T; i += blockDim.x) {
    const int co = i % CO_T, ci = (i / CO_T) % g.CI, tp = i / (CO_T * g.CI);
    ws_[i] = w[(int64_t)tp * g.w_tap + ci * g.w_ci + co * g.w_co];
  }
  __syncthreads();
  constexpr int V = 16 / (int)sizeof(TIn);                 // elements per 16-byte load
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int sub = threadIdx.x & 3;                         // channel quarter of this lane
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2);
  const bool live = p < P;
  const int pi = live ? (int)p : 0;
  const int x = pi % g.OW, y = (pi / g.OW) % g.OH, n = pi / (g.OW * g.OH);
  // channels [c_lo, c_hi) of every tap belong to this lane (multiples of V)
  const int per = ((g.CI / V + 3) / 4) * V;
  const int c_lo = sub * per, c_hi = min(g.CI, c_lo + per);
  float acc[CO_T];
#pragma unroll
  for (int co = 0; co < CO_T; ++co) acc[co] = 0.f;
  // The taps are walked in batches of kTB: all loads of a batch are issued before the first FMA, so a pixel costs
  // ceil(taps / kTB) global-memory round trips instead of one per tap (measured on cfg1: this kernel is pure latency,
  // 12.5 us for 1.6 MB with the one-tap-at-a-time loop).
  constexpr int kTB = 9;
  for (int c0 = c_lo; c0 < c_hi; c0 += V) {
    for (int tp0 = 0; tp0 < taps; tp0 += kTB) {
      uint4 u[kTB];
      bool ok[kTB];
#pragma unroll
      for (int j = 0; j < kTB; ++j) {
        const int tp = tp0 + j;
        const int r = tp / g.fc, t = tp - r * g.fc;
        int iy, ix;
        bool v = tp < taps;
        if (g.form == 0) {
          iy = y * g.s + r * g.D - g.pr; ix = x * g.s + t * g.D - g.pc;
        } else {
          const int ny = y + g.pr - r * g.D, nx = x + g.pc - t * g.D;
          v = v && ny >= 0 && nx >= 0 && ny % g.s == 0 && nx % g.s == 0;
          iy = ny / g.s; ix = nx / g.s;
        }
        v = v && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
        ok[j] = v;
        u[j] = make_uint4(0u, 0u, 0u, 0u);
        if (v) u[j] = *reinterpret_cast<const uint4*>(in + (((int64_t)n * g.IH + iy) * g.IW + ix) * g.CI + c0);
      }
#pragma unroll
      for (int j = 0; j < kTB; ++j) {
        if (!ok[j]) continue;
        const float* wt = ws_ + (size_t)(tp0 + j) * g.CI * CO_T;
        float v[V];
        if constexpr (sizeof(TIn) == 4) {
          v[0] = __uint_as_float(u[j].x); v[1] = __uint_as_float(u[j].y); v[2] = __uint_as_float(u[j].z); v[3] = __uint_as_float(u[j].w);
        } else {
          const uint32_t q[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&q[k]);
            v[2 * k] = __low2float(h); v[2 * k + 1] = __high2float(h);
          }
        }
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
          for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(v[k], wt[(c0 + k) * CO_T + co], acc[co]);
      }
    }
  }
#pragma unroll
  for (int co = 0; co < CO_T; ++co) {
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 1);
    acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], 2);
  }
  if (live && sub == 0) {
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
      const float z = acc[co] + (bias ? bias[co] : 0.f);
      if (Z) st_f(Z, p * CO_T + co, z);
      if (Y) st_f(Y, p * CO_T + co, post_affine(g, co, act_fn(g.act, z, g.p0, g.p1)));
    }
  }
}

// gather conv with a tiny reduction (taps * CI <= 64, e.g. Conv2D.forward with in_ch = 1): thread = (pixel, 4 output
// channels); the K gathered inputs are read once per thread (neighbouring threads share them through L1), the weights
// sit in shared memory as [k][co] and are read 16 bytes at a time; bias + activation in the same pass.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) thin_ci_conv_kernel(GatherConv g, const TIn* __restrict__ in,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           TOut* __restrict__ Z, TOut* __restrict__ Y) {
  extern __shared__ __align__(16) float ws_[];          // [tap*CI + ci][co], then the per-k offset table
  const int K = g.fr * g.fc * g.CI;
  // k -> (r*D, t*D, ci) once per block: the gather loop below was two integer divisions per input (ncu: 60 % of
  // this kernel's instructions were index arithmetic)
  int* ktab = reinterpret_cast<int*>(ws_ + K * g.CO);
  const bool w_dense = g.w_co == 1 && g.w_ci == g.CO && g.w_tap == (int64_t)g.CI * g.CO;
  if (w_dense) {
    for (int i = threadIdx.x; i < K * g.CO; i += blockDim.x) ws_[i] = w[i];
  } else {
    for (int i = threadIdx.x; i < K * g.CO; i += blockDim.x) {
      const int co = i % g.CO, k = i / g.CO;
      const int tp = k / g.CI, ci = k - tp * g.CI;
      ws_[i] = w[(int64_t)tp * g.w_tap + ci * g.w_ci + co * g.w_co];
    }
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const int tp = k / g.CI, ci = k - tp * g.CI;
    const int r = tp / g.fc, t = tp - r * g.fc;
    ktab[k] = (r * g.D) | ((t * g.D) << 10) | (ci << 20);
  }
  __syncthreads();
  const int cog = g.CO >> 2;                              // groups of 4 output channels
  const int64_t total = (int64_t)g.N * g.OH * g.OW * cog;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % cog) * 4;
  const int pi = (int)(idx / cog);
  const int x = pi % g.OW, y = (pi / g.OW) % g.OH, n = pi / (g.OW * g.OH);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  // K = taps * CI gathered inputs, kKB at a time: every load of a batch is in flight before the first FMA (one
  // round trip per batch instead of one per input; this kernel is latency, not bandwidth, on cfg1)
  constexpr int kKB = 9;
  const TIn* img = in + (int64_t)n * g.IH * g.IW * g.CI;
  for (int k0 = 0; k0 < K; k0 += kKB) {
    float a[kKB];
#pragma unroll
    for (int j = 0; j < kKB; ++j) {
      const int k = k0 + j;
      bool v = k < K;
      const int e = ktab[v ? k : 0];
      const int rD = e & 1023, tD = (e >> 10) & 1023, ci = e >> 20;
      int iy, ix;
      if (g.form == 0) {
        iy = y * g.s + rD - g.pr; ix = x * g.s + tD - g.pc;
      } else {
        const int ny = y + g.pr - rD, nx = x + g.pc - tD;
        v = v && ny >= 0 && nx >= 0 && ny % g.s == 0 && nx % g.s == 0;
        iy = ny / g.s; ix = nx / g.s;
      }
      v = v && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
      a[j] = v ? ld_f(img, (iy * g.IW + ix) * g.CI + ci) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kKB; ++j) {
      if (k0 + j >= K) break;
      const float4 wv = *reinterpret_cast<const float4*>(ws_ + (size_t)(k0 + j) * g.CO + c4);
      acc[0] = fmaf(a[j], wv.x, acc[0]); acc[1] = fmaf(a[j], wv.y, acc[1]);
      acc[2] = fmaf(a[j], wv.z, acc[2]); acc[3] = fmaf(a[j], wv.w, acc[3]);
    }
  }
  const int64_t o = (int64_t)pi * g.CO + c4;
  float zv[4], yv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    zv[j] = acc[j] + (bias ? bias[c4 + j] : 0.f);
    yv[j] = Y ? post_affine(g, c4 + j, act_fn(g.act, zv[j], g.p0, g.p1)) : 0.f;
  }
  // four channels per thread: one 8-byte (bf16) / 16-byte (fp32) store instead of four scalar ones (o % 4 == 0)
  if constexpr (sizeof(TOut) == 2) {
    if (Z) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(zv[0], zv[1]), b = __floats2bfloat162_rn(zv[2], zv[3]);
      *reinterpret_cast<uint2*>(Z + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
    if (Y) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(yv[0], yv[1]), b = __floats2bfloat162_rn(yv[2], yv[3]);
      *reinterpret_cast<uint2*>(Y + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
  } else {
    if (Z) *reinterpret_cast<float4*>(Z + o) = make_float4(zv[0], zv[1], zv[2], zv[3]);
    if (Y) *reinterpret_cast<float4*>(Y + o) = make_float4(yv[0], yv[1], yv[2], yv[3]);
  }
}

// wgrad with (taps*CI + 1) * CO <= 1024: a block owns a chunk of pixels.  The chunk's gathered inputs (an im2col tile
// of kThinPix pixels x M rows, plus a row of ones that yields db) and its dZ rows are staged in shared memory once,
// then thread (m, co) does 2 shared loads + 1 FMA per pixel.  Partials [chunk][M][CO] (+ [chunk][CO]) go to the
// fixed-order reduce.
constexpr int kThinPix = 64;
template <typename T>
__global__ void __launch_bounds__(1024) thin_m_wgrad_kernel(GatherWgrad g, const T* __restrict__ in,
                                                            const T* __restrict__ grad, float* __restrict__ part,
                                                            float* __restrict__ part_db, int64_t chunk) {
  extern __shared__ float sm_[];
  const int M = g.fr * g.fc * g.CI, M1 = M + 1;
  float* As = sm_;                       // [kThinPix][M1]
  float* Gs = sm_ + kThinPix * M1;       // [kThinPix][CO]
  const int nthr = M1 * g.CO;
  const int m = threadIdx.x / g.CO, co = threadIdx.x - m * g.CO;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int64_t p_begin = (int64_t)blockIdx.x * chunk;
  const int64_t p_end = (p_begin + chunk < P) ? p_begin + chunk : P;
  float acc = 0.f;
  for (int64_t p0 = p_begin; p0 < p_end; p0 += kThinPix) {
    const int np = (int)((p_end - p0 < kThinPix) ? p_end - p0 : kThinPix);
    __syncthreads();
    // staging loads four at a time: issued back to back, stored afterwards (one round trip per four elements; with a
    // load -> store per iteration the 64-pixel round was a chain of ~9 global-memory latencies)
    for (int i0 = threadIdx.x; i0 < np * M1; i0 += 4 * nthr) {
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        a[u] = 1.f;
        if (i < np * M1) {
          const int pp = i / M1, mm = i - pp * M1;
          if (mm < M) {
            const int q = (int)(p0 + pp);
            const int x = q % g.OW, y = (q / g.OW) % g.OH, n = q / (g.OW * g.OH);
            const int tap = mm / g.CI, ci = mm - tap * g.CI;
            const int r = tap / g.fc, t = tap - r * g.fc;
            const int iy = y * g.s + r * g.D - g.pr, ix = x * g.s + t * g.D - g.pc;
            a[u] = (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) ? ld_f(in, (((int64_t)n * g.IH + iy) * g.IW + ix) * g.CI + ci) : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < np * M1) As[i] = a[u];
      }
    }
    for (int i0 = threadIdx.x; i0 < np * g.CO; i0 += 4 * nthr) {
      float gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        gv[u] = i < np * g.CO ? ld_f(grad, p0 * g.CO + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < np * g.CO) Gs[i] = gv[u];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < np; ++pp) acc = fmaf(As[pp * M1 + m], Gs[pp * g.CO + co], acc);
  }
  if (m < M) part[((int64_t)blockIdx.x * M + m) * g.CO + co] = acc;
  else part_db[(int64_t)blockIdx.x * g.CO + co] = acc;
}

// ---- wgrad, register-tiled (fp32 storage, in_ch % 4 == 0, out_ch % 4 == 0): the same 128 x (16 TN) tile as
// gather_conv_rt_kernel with the pixels as the reduction dimension.  Both operands are read 16 bytes at a time along
// their channel dimension and land in shared memory without a transpose; a thread's pixel coordinates advance
// incrementally (no division in the loop).
template <int TN>
__global__ void __launch_bounds__(256, 2) wgrad_partial_rt_kernel(GatherWgrad g, const float* __restrict__ in,
                                                                  const float* __restrict__ grad, float* __restrict__ part,
                                                                  int64_t chunk) {
  constexpr int TBM = 128, TBN = 16 * TN, TBK = 16;
  __shared__ __align__(16) float As[2][TBK][TBM];
  __shared__ __align__(16) float Bs[2][TBK][TBN];
  const int tid = threadIdx.x;
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int M = g.fr * g.fc * g.CI;
  const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * TBN;
  const int64_t p_begin = (int64_t)blockIdx.z * chunk;
  const int64_t p_end = (p_begin + chunk < P) ? p_begin + chunk : P;
  const int nkb = (int)((p_end - p_begin + TBK - 1) / TBK);

  // A: thread -> four consecutive rows m (one tap, four in_ch) of pixels ka and ka + 8 of the K block
  const int am = m0 + (tid & 31) * 4, ka = tid >> 5;
  const bool mv = am < M;
  const int tap = mv ? am / g.CI : 0, ci = mv ? am - tap * g.CI : 0;
  const int r = tap / g.fc, t = tap - r * g.fc;
  int px[2], py[2], pn[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    int64_t q = p_begin + ka + 8 * u;
    if (q >= P) q = P - 1;                                // (never loaded: p < p_end is checked)
    px[u] = (int)(q % g.OW); q /= g.OW;
    py[u] = (int)(q % g.OH); pn[u] = (int)(q / g.OH);
  }
  // B: thread -> pixel bk = tid / 16 of the K block, columns 4 (tid % 16) + 64 v
  const int bk = tid >> 4, bc = (tid & 15) * 4;

  auto load_a = [&](int kb, int u) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t p = p_begin + (int64_t)kb * TBK + ka + 8 * u;
    if (mv && p < p_end) {
      const int iy = py[u] * g.s + r * g.D - g.pr, ix = px[u] * g.s + t * g.D - g.pc;
      if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW)
        v = *reinterpret_cast<const float4*>(in + (((int64_t)pn[u] * g.IH + iy) * g.IW + ix) * g.CI + ci);
    }
    return v;
  };
  auto advance_a = [&]() {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      px[u] += TBK;
      while (px[u] >= g.OW) {
        px[u] -= g.OW;
        if (++py[u] == g.OH) { py[u] = 0; ++pn[u]; }
      }
    }
  };
  auto load_b = [&](int kb, float (&bv)[TN]) {
    const int64_t p = p_begin + (int64_t)kb * TBK + bk;
#pragma unroll
    for (int v = 0; v < TN / 4; ++v) {
      const int co = n0 + bc + 64 * v;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < p_end && co < g.CO) q = *reinterpret_cast<const float4*>(grad + p * g.CO + co);
      bv[4 * v] = q.x; bv[4 * v + 1] = q.y; bv[4 * v + 2] = q.z; bv[4 * v + 3] = q.w;
    }
  };
  auto store_tile = [&](int buf, const float4 (&av)[2], const float (&bv)[TN]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) *reinterpret_cast<float4*>(&As[buf][ka + 8 * u][(tid & 31) * 4]) = av[u];
#pragma unroll
    for (int v = 0; v < TN / 4; ++v)
      *reinterpret_cast<float4*>(&Bs[buf][bk][bc + 64 * v]) = make_float4(bv[4 * v], bv[4 * v + 1], bv[4 * v + 2], bv[4 * v + 3]);
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float4 av[2];
  float bv[TN];
  if (nkb > 0) {
    av[0] = load_a(0, 0); av[1] = load_a(0, 1);
    load_b(0, bv);
    store_tile(0, av, bv);
  }
  __syncthreads();
  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    const bool more = kb + 1 < nkb;
    if (more) {
      advance_a();
      av[0] = load_a(kb + 1, 0); av[1] = load_a(kb + 1, 1);
      load_b(kb + 1, bv);
    }
#pragma unroll
    for (int k = 0; k < TBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int v = 0; v < TN / 4; ++v) {
        const float4 q = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4 + 64 * v]);
        b[4 * v] = q.x; b[4 * v + 1] = q.y; b[4 * v + 2] = q.z; b[4 * v + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tile(buf ^ 1, av, bv);
    __syncthreads();
  }
  float* out = part + (int64_t)blockIdx.z * M * g.CO;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int v = 0; v < TN / 4; ++v) {
      const int co = n0 + tx * 4 + 64 * v;
      if (co < g.CO)
        *reinterpret_cast<float4*>(out + (int64_t)m * g.CO + co) =
            make_float4(acc[i][4 * v], acc[i][4 * v + 1], acc[i][4 * v + 2], acc[i][4 * v + 3]);
    }
  }
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

// The same reduction with 16-byte loads: thread (x, y) of a (32, 16) block adds the splits z = y, y + 16, ... of FOUR
// consecutive outputs, four loads in flight at a time; the 16 group sums are added in a fixed order.  The partials
// were just written and sit in L2: 148 splits x 147 KB (cfg2) took 11.3 us with 4-byte loads and 19 dependent
// round trips per thread (2 TB/s).  Needs CO % 4 == 0 (then M * CO and the db tail are multiples of 4 too).
__global__ void __launch_bounds__(512) wgrad_reduce_v4_kernel(const float* __restrict__ part, int splits, int M, int CO,
                                                              int CI, int64_t o_tap, int64_t o_ci, int64_t o_co,
                                                              int accumulate, float* __restrict__ dW,
                                                              const float* __restrict__ part_db, float* __restrict__ db) {
  __shared__ float4 red[16][32];
  const int64_t total = (int64_t)M * CO;
  const int64_t total_all = total + (part_db != nullptr ? CO : 0);
  for (int64_t i0 = (int64_t)blockIdx.x * 128; i0 < total_all; i0 += (int64_t)gridDim.x * 128) {
    const int64_t i = i0 + 4 * threadIdx.x;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total_all) {
      const float* src = i < total ? part + i : part_db + (i - total);
      const int64_t stride = i < total ? total : (int64_t)CO;
      for (int z = threadIdx.y; z < splits; z += 64) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          v[u] = (z + 16 * u < splits) ? *reinterpret_cast<const float4*>(src + (int64_t)(z + 16 * u) * stride)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
      }
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y < 4) {                               // thread (x, e) finishes element e of the four
      const int64_t ie = i + threadIdx.y;
      if (ie < total_all) {
        float t = 0.f;
#pragma unroll
        for (int y = 0; y < 16; ++y) t += reinterpret_cast<const float*>(&red[y][threadIdx.x])[threadIdx.y];
        if (ie < total) {
          const int m = (int)(ie / CO), co = (int)(ie - (int64_t)m * CO);
          const int tap = m / CI, ci = m - tap * CI;
          const int64_t o = tap * o_tap + ci * o_ci + co * o_co;
          dW[o] = accumulate ? dW[o] + t : t;
        } else {
          const int co = (int)(ie - total);
          db[co] = accumulate ? db[co] + t : t;
        }
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
  if (CO % 4 == 0 && (reinterpret_cast<uintptr_t>(part) & 15) == 0 &&
      (part_db == nullptr || (reinterpret_cast<uintptr_t>(part_db) & 15) == 0)) {
    int blocks = (int)((total + 127) / 128);
    if (blocks > 8192) blocks = 8192;
    wgrad_reduce_v4_kernel<<<blocks, dim3(32, 16), 0, st>>>(part, splits, M, CO, CI, o_tap, o_ci, o_co, accumulate, dW,
                                                            part_db, db);
    NPML_LAUNCH_OK();
    return 0;
  }
  int blocks = (int)((total + 63) / 64);
  if (blocks > 8192) blocks = 8192;
  wgrad_reduce_kernel<<<blocks, dim3(64, 8), 0, st>>>(part, splits, M, CO, CI, o_tap, o_ci, o_co, accumulate, dW, part_db, db);
  NPML_LAUNCH_OK();
  return 0;
}

// The output parity classes of a gather convolution for gather_conv_rt_kernel (see RtClass): one class with every tap
// for the forward form, up to s^2 classes with the taps that reach them for the transposed form.
constexpr int kRtMaxClasses = 9;
static int build_rt_classes(const GatherConv& g, RtClass* out) {
  auto pm = [](int a, int b) { const int m = a % b; return m < 0 ? m + b : m; };
  const int nq = g.form == 1 ? g.s : 1;
  int n = 0;
  for (int qy = 0; qy < nq; ++qy)
    for (int qx = 0; qx < nq; ++qx) {
      if (n >= kRtMaxClasses) return n;
      RtClass& c = out[n];
      memset(&c, 0, sizeof(c));
      if (g.form == 0) {
        c.y0 = 0; c.x0 = 0; c.so = 1; c.sin = g.s; c.oh = g.OH; c.ow = g.OW;
        for (int r = 0; r < g.fr; ++r)
          for (int t = 0; t < g.fc; ++t) {
            c.dy[c.ntaps] = (int16_t)(r * g.D - g.pr); c.dx[c.ntaps] = (int16_t)(t * g.D - g.pc);
            c.wt[c.ntaps++] = (int16_t)(r * g.fc + t);
          }
      } else {
        // outputs y = y0 + s yy see the taps with (y + pr - r D) % s == 0, i.e. (r D) % s == qy where y0 = (qy - pr) mod s
        c.y0 = pm(qy - g.pr, g.s); c.x0 = pm(qx - g.pc, g.s); c.so = g.s; c.sin = 1;
        if (c.y0 >= g.OH || c.x0 >= g.OW) continue;
        c.oh = (g.OH - c.y0 + g.s - 1) / g.s; c.ow = (g.OW - c.x0 + g.s - 1) / g.s;
        for (int r = 0; r < g.fr; ++r) {
          if (pm(r * g.D, g.s) != qy) continue;
          for (int t = 0; t < g.fc; ++t) {
            if (pm(t * g.D, g.s) != qx) continue;
            c.dy[c.ntaps] = (int16_t)((c.y0 + g.pr - r * g.D) / g.s);      // exact by construction
            c.dx[c.ntaps] = (int16_t)((c.x0 + g.pc - t * g.D) / g.s);
            c.wt[c.ntaps++] = (int16_t)(r * g.fc + t);
          }
        }
      }
      ++n;
    }
  return n;
}

// Host-only test hook (npml_debug_rt_classes): out = [classes, then per class ntaps, y0, x0, so, sin, oh, ow and ntaps
// triples (dy, dx, filter tap)]; returns the number of integers written or -1 (tests/test_rt_classes_cpu.py).
int rt_classes_debug(const GatherConv& g, int32_t* out, int32_t cap) {
  if (g.fr * g.fc > kRtMaxTaps || g.s < 1 || g.s > 3) return -1;
  RtClass cls[kRtMaxClasses];
  const int ncls = build_rt_classes(g, cls);
  int n = 0;
  if (cap < 1) return -1;
  out[n++] = ncls;
  for (int q = 0; q < ncls; ++q) {
    const RtClass& c = cls[q];
    if (n + 7 + 3 * c.ntaps > cap) return -1;
    const int32_t head[7] = {c.ntaps, c.y0, c.x0, c.so, c.sin, c.oh, c.ow};
    for (int i = 0; i < 7; ++i) out[n++] = head[i];
    for (int j = 0; j < c.ntaps; ++j) { out[n++] = c.dy[j]; out[n++] = c.dx[j]; out[n++] = c.wt[j]; }
  }
  return n;
}

int direct_gather_conv(const GatherConv& g, const void* in, const float* w, const float* bias, void* Z,
                       void* Y, int dtype, cudaStream_t st) {
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  if (P == 0 || g.CO == 0) return 0;
  const int vin = dtype == NPML_F32 ? 4 : 8;                  // elements per 16-byte load
  const size_t wn = (size_t)g.fr * g.fc * g.CI * g.CO * sizeof(float);
  if (g.CO <= 4 && P < (1LL << 31) && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && g.CI % vin == 0 && wn <= 40 * 1024) {
    const unsigned blocks = (unsigned)((P + 31) / 32);         // 32 pixels x 4 channel quarters per block
#define NPML_THIN_CO(COV)                                                                                        \
    if (dtype == NPML_F32)                                                                                         \
      thin_co_conv_kernel<float, float, COV><<<blocks, 128, wn, st>>>(g, (const float*)in, w, bias, (float*)Z, (float*)Y); \
    else                                                                                                           \
      thin_co_conv_kernel<__nv_bfloat16, __nv_bfloat16, COV><<<blocks, 128, wn, st>>>(                             \
          g, (const __nv_bfloat16*)in, w, bias, (__nv_bfloat16*)Z, (__nv_bfloat16*)Y);
    switch (g.CO) {
      case 1: NPML_THIN_CO(1) break;
      case 2: NPML_THIN_CO(2) break;
      case 3: NPML_THIN_CO(3) break;
      default: NPML_THIN_CO(4) break;
    }
#undef NPML_THIN_CO
    NPML_LAUNCH_OK();
    return 0;
  }
  if (g.fr * g.fc * g.CI <= 64 && g.CO % 4 == 0 && wn <= 40 * 1024 && P * (g.CO / 4) < (1LL << 31) &&
      (g.fr - 1) * g.D < 1024 && (g.fc - 1) * g.D < 1024 && (int64_t)g.IH * g.IW * g.CI < (1LL << 31) &&
      (reinterpret_cast<uintptr_t>(Z) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0) {
    const int64_t total = P * (g.CO / 4);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const size_t wk = wn + (size_t)g.fr * g.fc * g.CI * sizeof(int);        // weights + the per-k offset table
    if (dtype == NPML_F32)
      thin_ci_conv_kernel<float, float><<<blocks, 256, wk, st>>>(g, (const float*)in, w, bias, (float*)Z, (float*)Y);
    else
      thin_ci_conv_kernel<__nv_bfloat16, __nv_bfloat16>
          <<<blocks, 256, wk, st>>>(g, (const __nv_bfloat16*)in, w, bias, (__nv_bfloat16*)Z, (__nv_bfloat16*)Y);
    NPML_LAUNCH_OK();
    return 0;
  }
  if (dtype == NPML_F32 && g.CI % 4 == 0 && g.CO % 4 == 0 && g.CI >= 16 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(Z) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(Y) & 15) == 0 && P >= 256 && g.fr * g.fc <= kRtMaxTaps && g.s <= 3 &&
      !getenv("NPML_DISABLE_FP32_RT")) {
    // register-tiled kernel, one launch per output parity class: 128-wide N tile for wide outputs, 64-wide otherwise
    RtClass cls[kRtMaxClasses];
    const int ncls = build_rt_classes(g, cls);
    for (int q = 0; q < ncls; ++q) {
      const RtClass& c = cls[q];
      const int64_t Pc = (int64_t)g.N * c.oh * c.ow;
      if (Pc == 0) continue;
      if (g.CO > 64) {
        dim3 grid((unsigned)((Pc + 127) / 128), (unsigned)((g.CO + 127) / 128));
        gather_conv_rt_kernel<8><<<grid, 256, 0, st>>>(g, c, (const float*)in, w, bias, (float*)Z, (float*)Y);
      } else {
        dim3 grid((unsigned)((Pc + 127) / 128), (unsigned)((g.CO + 63) / 64));
        gather_conv_rt_kernel<4><<<grid, 256, 0, st>>>(g, c, (const float*)in, w, bias, (float*)Z, (float*)Y);
      }
      NPML_LAUNCH_OK();
    }
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
  const int64_t m1 = (int64_t)g.fr * g.fc * g.CI + 1;
  return m1 * g.CO <= 1024 && (m1 + g.CO) * kThinPix * (int64_t)sizeof(float) <= 40 * 1024;
}

static int direct_wgrad_splits(const GatherWgrad& g) {
  const int64_t P = (int64_t)g.N * g.OH * g.OW;
  const int M = g.fr * g.fc * g.CI;
  if (thin_wgrad(g)) {               // one block per pixel chunk, about two blocks per SM
    int64_t splits = 2LL * num_sms();
    if (splits * kThinPix > P) splits = (P + kThinPix - 1) / kThinPix;
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
    const size_t smem = (size_t)kThinPix * (M + 1 + g.CO) * sizeof(float);
    if (dtype == NPML_F32)
      thin_m_wgrad_kernel<float><<<nsplit, threads, smem, st>>>(g, (const float*)in, (const float*)grad, part, part_db, chunk);
    else
      thin_m_wgrad_kernel<__nv_bfloat16>
          <<<nsplit, threads, smem, st>>>(g, (const __nv_bfloat16*)in, (const __nv_bfloat16*)grad, part, part_db, chunk);
    NPML_LAUNCH_OK();
    if (db_done) *db_done = db != nullptr;
    return launch_wgrad_reduce(part, nsplit, M, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate, dW, st,
                               db != nullptr ? part_db : nullptr, db);
  }
  int64_t chunk = (P + splits - 1) / splits;
  chunk = (chunk + BK - 1) / BK * BK;
  if (dtype == NPML_F32 && g.CI % 4 == 0 && g.CO % 4 == 0 && g.CI >= 4 && M >= 64 && P >= 256 &&
      (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(ws) & 15) == 0 && !getenv("NPML_DISABLE_FP32_RT")) {
    if (g.CO > 64) {
      dim3 grid((M + 127) / 128, (g.CO + 127) / 128, splits);
      wgrad_partial_rt_kernel<8><<<grid, 256, 0, st>>>(g, (const float*)in, (const float*)grad, (float*)ws, chunk);
    } else {
      dim3 grid((M + 127) / 128, (g.CO + 63) / 64, splits);
      wgrad_partial_rt_kernel<4><<<grid, 256, 0, st>>>(g, (const float*)in, (const float*)grad, (float*)ws, chunk);
    }
    NPML_LAUNCH_OK();
    return launch_wgrad_reduce((const float*)ws, splits, M, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate,
                               dW, st);
  }
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
