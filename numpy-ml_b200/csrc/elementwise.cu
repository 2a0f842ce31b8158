// HBM-bound passes around the convolution GEMMs: the backward prologue dZ = dLdy * act'(Z) with
// the db column sum (layers/layers.py:3103, 3109), Add (layers/layers.py:633-742), BatchNorm2D
// (layers/layers.py:969-1215), dtype casts, and the im2col / col2im exports.
// All column reductions are two-stage with a fixed order (deterministic, no atomics).
#include "npml_common.cuh"

namespace npml {

namespace {

// ---- 4-wide typed access -----------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void load_vec(const void* p, int dtype, int64_t i, float (&v)[VEC]) {
  if constexpr (VEC == 8) {        // 16-byte (bf16) / 2 x 16-byte (fp32) accesses
    if (dtype == NPML_F32) {
      const float4 t0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
      const float4 t1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
      v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else if (dtype == NPML_BF16) {
      const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[2 * j] = __low2float(a); v[2 * j + 1] = __high2float(a);
      }
    } else {
      const double* d = reinterpret_cast<const double*>(p) + i;
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = (float)d[j];
    }
  } else if constexpr (VEC == 4) {
    if (dtype == NPML_F32) {
      const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if (dtype == NPML_BF16) {
      const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
      v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    } else {
      const double* d = reinterpret_cast<const double*>(p) + i;
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[j] = (float)d[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (dtype == NPML_F32) v[j] = reinterpret_cast<const float*>(p)[i + j];
      else if (dtype == NPML_BF16) v[j] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i + j]);
      else v[j] = (float)reinterpret_cast<const double*>(p)[i + j];
    }
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec(void* p, int dtype, int64_t i, const float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    if (dtype == NPML_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (dtype == NPML_BF16) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&a);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      double* d = reinterpret_cast<double*>(p) + i;
#pragma unroll
      for (int j = 0; j < VEC; ++j) d[j] = (double)v[j];
    }
  } else if constexpr (VEC == 4) {
    if (dtype == NPML_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (dtype == NPML_BF16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
      uint2 t;
      t.x = *reinterpret_cast<uint32_t*>(&a);
      t.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = t;
    } else {
      double* d = reinterpret_cast<double*>(p) + i;
#pragma unroll
      for (int j = 0; j < VEC; ++j) d[j] = (double)v[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      if (dtype == NPML_F32) reinterpret_cast<float*>(p)[i + j] = v[j];
      else if (dtype == NPML_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i + j] = __float2bfloat16_rn(v[j]);
      else reinterpret_cast<double*>(p)[i + j] = (double)v[j];
    }
  }
}

inline int dtype_size(int dt) { return dt == NPML_F32 ? 4 : (dt == NPML_BF16 ? 2 : 8); }
inline bool aligned_for_vec4(const void* p, int dt) {
  return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (size_t)(4 * dtype_size(dt))) == 0;
}
inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// widest per-thread vector (elements): 8 when every tensor is 16-byte aligned and n % 8 == 0, else 4, else 1
inline int pick_vec(int64_t n, bool ok4, bool ok16) { return (ok16 && n % 8 == 0) ? 8 : ((ok4 && n % 4 == 0) ? 4 : 1); }

// ---- column-reduction skeleton ----------------------------------------------------------------------
// A [rows][ch] tensor is cut into gridDim.x row slabs; block (TX, TY): thread (tx, ty) owns VEC
// consecutive channels and rows ty, ty+TY, ... of the slab.  The functor both does the elementwise
// work and returns NV values per element to be column-summed.  part: [gridDim.x][NV][ch].
struct ColGeom {
  int64_t rows;
  int ch;
  int64_t rows_per_block;
  int TX, TY;  // threads over channel groups / over rows
};

// DT >= 0: every tensor of the pass has element type DT, known at compile time -- the dtype switch of load_vec /
// store_vec then folds away, which is what lets the unrolled row loop keep several 16-byte loads in flight (with a
// run-time dtype there is a branch in front of every load).  DT = -1: run-time element types.
template <int VEC, int NV, typename Acc, typename F, int DT>
__global__ void __launch_bounds__(256) col_pass_kernel(F f, ColGeom cg, Acc* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Acc* red = reinterpret_cast<Acc*>(smem_raw);  // [TY][TX*VEC*NV]
  const int tx = threadIdx.x % cg.TX, ty = threadIdx.x / cg.TX;
  const int c0 = (blockIdx.y * cg.TX + tx) * VEC;
  const bool active = ty < cg.TY && c0 < cg.ch;
  // per-thread partial sums stay in fp32 (a thread adds at most rows_per_block / TY ~ tens of terms); the block
  // and grid levels of the reduction run in Acc
  float acc[NV > 0 ? NV : 1][VEC];
#pragma unroll
  for (int a = 0; a < NV; ++a)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[a][j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * cg.rows_per_block;
  const int64_t r1 = (r0 + cg.rows_per_block < cg.rows) ? r0 + cg.rows_per_block : cg.rows;
  if (active) {
    // a thread keeps the same VEC channels for all of its rows: per-channel constants are loaded once
    float kc[F::kConsts > 0 ? F::kConsts : 1][VEC];
    f.template init<VEC>(c0, cg.ch, kc);
    // kBatch rows per trip: ALL their loads are issued before the first store.  The output of a pass may alias an
    // input (dZ on dY), so the compiler will not move a load above an earlier row's store by itself, and one row of
    // loads in flight per thread leaves these passes at ~3.8 TB/s on 67 MB tensors.
    constexpr int kBatch = F::kBatch;
    int64_t r = r0 + ty;
#pragma unroll(kBatch == 1 ? 4 : 1)
    for (; r + (int64_t)(kBatch - 1) * cg.TY < r1; r += (int64_t)kBatch * cg.TY) {
      float in[kBatch][F::kIn][VEC];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) f.template load<VEC, DT>((r + (int64_t)u * cg.TY) * cg.ch + c0, in[u]);
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        float out[NV > 0 ? NV : 1][VEC];
        f.template finish<VEC, DT>((r + (int64_t)u * cg.TY) * cg.ch + c0, kc, in[u], out);
#pragma unroll
        for (int a = 0; a < NV; ++a)
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[a][j] += out[a][j];
      }
    }
    for (; r < r1; r += cg.TY) {
      float in[F::kIn][VEC];
      float out[NV > 0 ? NV : 1][VEC];
      f.template load<VEC, DT>(r * cg.ch + c0, in);
      f.template finish<VEC, DT>(r * cg.ch + c0, kc, in, out);
#pragma unroll
      for (int a = 0; a < NV; ++a)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[a][j] += out[a][j];
    }
  }
  if (NV == 0 || part == nullptr) return;
  const int W = cg.TX * VEC * NV;
  if (ty < cg.TY) {
#pragma unroll
    for (int a = 0; a < NV; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) red[ty * W + (a * cg.TX + tx) * VEC + j] = active ? (Acc)acc[a][j] : (Acc)0;
  }
  __syncthreads();
  // every thread adds one column of the TY x W staging matrix (fixed order over the rows)
  for (int col = threadIdx.x; col < W; col += blockDim.x) {
    Acc s = (Acc)0;
    for (int y = 0; y < cg.TY; ++y) s += red[y * W + col];
    const int a = col / (cg.TX * VEC), rem = col - a * cg.TX * VEC;
    const int c = (blockIdx.y * cg.TX + rem / VEC) * VEC + rem % VEC;
    if (c < cg.ch) part[((int64_t)blockIdx.x * NV + a) * cg.ch + c] = s;
  }
}

// Fixed-order sum of column `c` over the `nblk` partial rows, for an (8, 128) thread block: thread (cx, sy) adds rows
// sy, sy + 128, ... (5 for the usual 592 partial rows; the (32, 32) shape of round 1 walked 19 per thread on a quarter of
// the blocks: these kernels are nothing but that chain of L2 round trips), then 128 -> 8 lanes -> 1 in a fixed order.
constexpr int kFinCols = 8, kFinRows = 128;
template <typename Acc>
__device__ __forceinline__ double block_col_sum(const Acc* __restrict__ part, int nblk, int width, int c, bool ok,
                                                double (*red)[kFinCols + 1]) {
  const int cx = threadIdx.x, sy = threadIdx.y;
  double s = 0.0;
  if (ok)
    for (int b = sy; b < nblk; b += kFinRows) s += (double)part[(int64_t)b * width + c];
  __syncthreads();
  red[sy][cx] = s;
  __syncthreads();
  if (sy < 8) {
    double t = 0.0;
    for (int y = 0; y < 16; ++y) t += red[sy * 16 + y][cx];
    red[sy * 16][cx] = t;                      // (row sy * 16 is read by this thread only)
  }
  __syncthreads();
  double t = 0.0;
  for (int y = 0; y < 8; ++y) t += red[y * 16][cx];
  return t;
}

// sums[NV*ch] (double) = fixed-order sum over `nblk` partial rows
template <typename Acc>
__global__ void col_final_kernel(const Acc* __restrict__ part, int nblk, int width, double* __restrict__ sums) {
  __shared__ double red[8][33];
  const int cx = threadIdx.x, sy = threadIdx.y;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (c < width)
    for (int b = sy; b < nblk; b += 8) s += (double)part[(int64_t)b * width + c];
  red[sy][cx] = s;
  __syncthreads();
  if (sy == 0 && c < width) {
    double t = 0.0;
    for (int y = 0; y < 8; ++y) t += red[y][cx];
    sums[c] = t;
  }
}

struct ColPlan {
  ColGeom cg;
  dim3 grid;
  int vec;
  size_t smem_elems;  // per NV
};

ColPlan plan_cols(int64_t rows, int ch, int vec, bool reduce = true) {
  ColPlan p;
  p.vec = vec;
  const int groups = (ch + p.vec - 1) / p.vec;
  int TX = 1;
  while (TX < groups && TX < 256) TX <<= 1;
  if (TX > 256) TX = 256;
  const int TY = 256 / TX;
  // passes that end in a block reduction keep the grid small (the reduction is a fixed cost per block) and get
  // their memory-level parallelism from the unrolled row loop; pure elementwise passes use a larger grid
  int64_t nblk = (reduce ? 4LL : 8LL) * num_sms();
  const int64_t min_rows = 4LL * TY;
  if (nblk * min_rows > rows) nblk = (rows + min_rows - 1) / min_rows;
  if (nblk < 1) nblk = 1;
  p.cg.rows = rows;
  p.cg.ch = ch;
  p.cg.rows_per_block = (rows + nblk - 1) / nblk;
  if (p.cg.rows_per_block < 1) p.cg.rows_per_block = 1;      // rows == 0 (an empty batch): one idle block
  p.cg.TX = TX;
  p.cg.TY = TY;
  nblk = (rows + p.cg.rows_per_block - 1) / p.cg.rows_per_block;
  if (nblk < 1) nblk = 1;
  p.grid = dim3((unsigned)nblk, (unsigned)((groups + TX - 1) / TX));
  p.smem_elems = (size_t)TY * TX * p.vec;
  return p;
}

template <int NV, typename Acc, typename F, int DT>
int launch_col_pass_dt(const ColPlan& p, F f, Acc* part, cudaStream_t st) {
  const size_t smem = NV * p.smem_elems * sizeof(Acc);
  if (p.vec == 8)
    col_pass_kernel<8, NV, Acc, F, DT><<<p.grid, 256, smem, st>>>(f, p.cg, part);
  else if (p.vec == 4)
    col_pass_kernel<4, NV, Acc, F, DT><<<p.grid, 256, smem, st>>>(f, p.cg, part);
  else
    col_pass_kernel<1, NV, Acc, F, DT><<<p.grid, 256, smem, st>>>(f, p.cg, part);
  NPML_LAUNCH_OK();
  return 0;
}

// dt: the element type shared by every tensor of the pass (NPML_F32 / NPML_BF16), or -1 when they differ
template <int NV, typename Acc, typename F>
int launch_col_pass(const ColPlan& p, F f, Acc* part, cudaStream_t st, int dt = -1) {
  if (dt == NPML_F32) return launch_col_pass_dt<NV, Acc, F, NPML_F32>(p, f, part, st);
  if (dt == NPML_BF16) return launch_col_pass_dt<NV, Acc, F, NPML_BF16>(p, f, part, st);
  return launch_col_pass_dt<NV, Acc, F, -1>(p, f, part, st);
}

template <typename Acc>
int launch_col_final(const Acc* part, int nblk, int width, double* sums, cudaStream_t st) {
  col_final_kernel<Acc><<<(width + 31) / 32, dim3(32, 8), 0, st>>>(part, nblk, width, sums);
  NPML_LAUNCH_OK();
  return 0;
}

// ---- functors -----------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ int sel(int runtime_dt) { return DT >= 0 ? DT : runtime_dt; }

// Functor protocol: kConsts per-channel constants are fetched once per thread by init(c0, ch, kc); load(i, in) reads
// the kIn input vectors at flat index i, finish(i, kc, in, out) does the arithmetic and the store and returns the NV
// values to be column-summed.
static __device__ __noinline__ float act_grad_ool(int act, float x, float p0, float p1) { return act_grad(act, x, p0, p1); }

struct DzPrepF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 2;
  static constexpr int kBatch = 2;      // rows whose loads are issued together (registers: kBatch*kIn*VEC)
  const void* dY; const void* Z; void* dZ;
  int dy_dt, z_dt, dz_dt, act;
  float p0, p1;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[2][VEC]) const {
    load_vec<VEC>(dY, sel<DT>(dy_dt), i, in[0]);
    if (act != NPML_ACT_IDENTITY) load_vec<VEC>(Z, sel<DT>(z_dt), i, in[1]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t i, const float (&)[1][VEC], float (&in)[2][VEC], float (&out)[1][VEC]) const {
    float (&g)[VEC] = in[0];
    if (act == NPML_ACT_RELU) {                 // the common case stays inline; one copy of the 12-way switch otherwise
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] = in[1][j] > 0.f ? g[j] : 0.f;
    } else if (act != NPML_ACT_IDENTITY) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] *= act_grad_ool(act, in[1][j], p0, p1);
    }
    // the sum is taken over the values the GEMMs will see (after rounding to the dZ dtype)
    if (sel<DT>(dz_dt) == NPML_BF16) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] = __bfloat162float(__float2bfloat16_rn(g[j]));
    }
    if (dZ) store_vec<VEC>(dZ, sel<DT>(dz_dt), i, g);
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[0][j] = g[j];
  }
};

// column sums of one tensor (db of an identity-activation layer whose dZ is dLdy itself)
struct ColSumF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 1;
  static constexpr int kBatch = 4;
  const void* x; int dt;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[1][VEC]) const { load_vec<VEC>(x, sel<DT>(dt), i, in[0]); }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t, const float (&)[1][VEC], float (&in)[1][VEC], float (&out)[1][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[0][j] = in[0][j];
  }
};

struct BnStatsF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 1;
  static constexpr int kBatch = 4;
  const void* x; int dt;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[1][VEC]) const { load_vec<VEC>(x, sel<DT>(dt), i, in[0]); }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t, const float (&)[1][VEC], float (&in)[1][VEC], float (&out)[2][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) { out[0][j] = in[0][j]; out[1][j] = in[0][j] * in[0][j]; }
  }
};

// y = a*(x - mean) + h with a = scaler*rstd (coef: [3][ch] = mean, a, h, written by the finalize kernels)
struct BnNormF {
  static constexpr int kConsts = 3;
  static constexpr int kIn = 1;
  static constexpr int kBatch = 4;
  const void* x; void* y; int dt;
  const float* coef;
  template <int VEC>
  __device__ __forceinline__ void init(int c0, int ch, float (&kc)[3][VEC]) const {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) kc[a][j] = coef[a * ch + (c0 + j < ch ? c0 + j : ch - 1)];
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[1][VEC]) const { load_vec<VEC>(x, sel<DT>(dt), i, in[0]); }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t i, const float (&kc)[3][VEC], float (&in)[1][VEC], float (&)[1][VEC]) const {
    float (&v)[VEC] = in[0];
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = fmaf(kc[1][j], v[j] - kc[0][j], kc[2][j]);
    store_vec<VEC>(y, sel<DT>(dt), i, v);
  }
};

// raw column sums [sum dy, sum dy*x]; the finalize kernels turn them into sum dy*nhat = rstd*(sum dy*x - mean*sum dy)
// in double, so the pass itself carries no per-channel constants (registers -> occupancy -> bytes in flight)
struct BnBwdStatsF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 2;
  static constexpr int kBatch = 2;
  const void* dy; const void* x; int dt;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[2][VEC]) const {
    load_vec<VEC>(dy, sel<DT>(dt), i, in[0]);
    load_vec<VEC>(x, sel<DT>(dt), i, in[1]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t, const float (&)[1][VEC], float (&in)[2][VEC], float (&out)[2][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      out[0][j] = in[0][j];
      out[1][j] = in[0][j] * in[1][j];
    }
  }
};

// dx = g*rstd*(dy - m1 - nhat*m2), nhat = (x - mean)*rstd, regrouped as dx = A1*dy + A2*x + A3 with
// coef: [3][ch] = A1 = g*rstd, A2 = -A1*m2*rstd, A3 = -A1*m1 - A2*mean (formed in double by the finalize kernel).
// act != identity: the layer in front of this BatchNorm2D ends in an activation (x = act(z)); its backward prologue
// dZ = dx * act'(z) (layers/layers.py:3103) is applied here, in the same pass, instead of in a pass of its own.  ReLU
// needs no extra tensor (x > 0 <=> z > 0); every other activation reads z from `zm`.
struct BnBwdDxF {
  static constexpr int kConsts = 3;
  static constexpr int kIn = 3;
  static constexpr int kBatch = 1;
  const void* dy; const void* x; void* dx; int dt;
  const float* coef;
  int act = NPML_ACT_IDENTITY;
  float p0 = 0.f, p1 = 0.f;
  const void* zm = nullptr;
  template <int VEC>
  __device__ __forceinline__ void init(int c0, int ch, float (&kc)[3][VEC]) const {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) kc[a][j] = coef[a * ch + (c0 + j < ch ? c0 + j : ch - 1)];
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[3][VEC]) const {
    load_vec<VEC>(dy, sel<DT>(dt), i, in[0]);
    load_vec<VEC>(x, sel<DT>(dt), i, in[1]);
    if (zm != nullptr) load_vec<VEC>(zm, sel<DT>(dt), i, in[2]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t i, const float (&kc)[3][VEC], float (&in)[3][VEC], float (&)[1][VEC]) const {
    float (&v)[VEC] = in[1];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float xv = v[j];
      float d = fmaf(kc[0][j], in[0][j], fmaf(kc[1][j], xv, kc[2][j]));
      if (act == NPML_ACT_RELU) {
        d = (zm != nullptr ? in[2][j] : xv) > 0.f ? d : 0.f;
      } else if (act != NPML_ACT_IDENTITY) {
        // the unfused path rounds dx to the storage type before the prologue multiplies it
        if (sel<DT>(dt) == NPML_BF16) d = __bfloat162float(__float2bfloat16_rn(d));
        d *= act_grad_ool(act, in[2][j], p0, p1);
      }
      v[j] = d;
    }
    store_vec<VEC>(dx, sel<DT>(dt), i, v);
  }
};

// ---- the residual tail of the skip-connection modules, fused (modules/modules.py:931-937, 941-953) -----------------
// Forward: Y = act(BN_a(xa) + BN_b(xb)) in ONE pass over (xa, xb) instead of two BatchNorm2D passes and an Add pass.
// `b` may be a plain tensor (identity shortcut: scaler_b == nullptr -> xb is added as it is).
struct BnPairAddActF {
  static constexpr int kConsts = 3;      // a_a, a_b, c: sum = a_a*xa + a_b*xb + c
  static constexpr int kIn = 2;
  static constexpr int kBatch = 2;
  const void* xa; const void* xb; void* sum; void* Y; int dt;
  const float* scaler_a; const float* intercept_a; const float* mean_a; const float* rstd_a;
  const float* scaler_b; const float* intercept_b; const float* mean_b; const float* rstd_b;
  int act; float p0, p1;
  template <int VEC>
  __device__ __forceinline__ void init(int c0, int ch, float (&kc)[3][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int c = c0 + j < ch ? c0 + j : ch - 1;
      const double aa = (double)scaler_a[c] * (double)rstd_a[c];
      double ab = 1.0, k = (double)intercept_a[c] - aa * (double)mean_a[c];
      if (scaler_b != nullptr) {
        ab = (double)scaler_b[c] * (double)rstd_b[c];
        k += (double)intercept_b[c] - ab * (double)mean_b[c];
      }
      kc[0][j] = (float)aa; kc[1][j] = (float)ab; kc[2][j] = (float)k;
    }
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[2][VEC]) const {
    load_vec<VEC>(xa, sel<DT>(dt), i, in[0]);
    load_vec<VEC>(xb, sel<DT>(dt), i, in[1]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t i, const float (&kc)[3][VEC], float (&in)[2][VEC], float (&)[1][VEC]) const {
    float (&v)[VEC] = in[0];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float t = fmaf(kc[0][j], v[j], fmaf(kc[1][j], in[1][j], kc[2][j]));
      if (sel<DT>(dt) == NPML_BF16) t = __bfloat162float(__float2bfloat16_rn(t));   // activations see the stored sum
      v[j] = t;
    }
    if (sum) store_vec<VEC>(sum, sel<DT>(dt), i, v);
    if (Y) {
      if (act == NPML_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
      } else if (act != NPML_ACT_IDENTITY) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = act_fn(act, v[j], p0, p1);
      }
      store_vec<VEC>(Y, sel<DT>(dt), i, v);
    }
  }
};

// Backward, pass 1: g = dY * act'(s) is never written; its raw column sums [sum g, sum g*xa, sum g*xb] are all the two
// BatchNorm backwards need (pair_stats_fix_kernel turns them into sum g*nhat; dIntercept of both, dScaler_a, dScaler_b;
// layers/layers.py:1207-1213).  For ReLU `s` may be the OUTPUT Y (Y > 0 <=> sum > 0); identity reads no s at all.
struct BnPairBwdStatsF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 4;
  static constexpr int kBatch = 1;
  const void* dY; const void* s; const void* xa; const void* xb; int dt;
  int has_b;
  int act; float p0, p1;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[4][VEC]) const {
    load_vec<VEC>(dY, sel<DT>(dt), i, in[0]);
    if (act != NPML_ACT_IDENTITY) load_vec<VEC>(s, sel<DT>(dt), i, in[1]);
    load_vec<VEC>(xa, sel<DT>(dt), i, in[2]);
    if (has_b) load_vec<VEC>(xb, sel<DT>(dt), i, in[3]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t, const float (&)[1][VEC], float (&in)[4][VEC], float (&out)[3][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float g = in[0][j];
      if (act == NPML_ACT_RELU) g = in[1][j] > 0.f ? g : 0.f;
      else if (act != NPML_ACT_IDENTITY) g *= act_grad_ool(act, in[1][j], p0, p1);
      if (sel<DT>(dt) == NPML_BF16) g = __bfloat162float(__float2bfloat16_rn(g));      // the value the unfused path stores
      out[0][j] = g;
      out[1][j] = g * in[2][j];
      out[2][j] = has_b ? g * in[3][j] : 0.f;
    }
  }
};

// raw [sum g, sum g*xa, sum g*xb] -> [sum g, sum g*nhat_a, sum g*nhat_b] (double; linear, so it commutes with the
// all-reduce of sync-BN as long as mean / rstd are the global ones, which they are)
__global__ void pair_stats_fix_kernel(int ch, double* sums, const float* mean_a, const float* rstd_a, const float* mean_b,
                                      const float* rstd_b) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  const double sg = sums[c];
  sums[ch + c] = (double)rstd_a[c] * (sums[ch + c] - (double)mean_a[c] * sg);
  sums[2 * ch + c] = mean_b != nullptr ? (double)rstd_b[c] * (sums[2 * ch + c] - (double)mean_b[c] * sg) : 0.0;
}

// Backward, pass 2: dxa = BN_a'(g), dxb = BN_b'(g) (or dxb = g for an identity shortcut), both from one read of
// (dY, s, xa, xb): dx = A1*g + A2*x + A3 per BatchNorm2D (see BnBwdDxF).  coef: [6][ch] = A1a, A2a, A3a, A1b, A2b, A3b.
struct BnPairBwdDxF {
  static constexpr int kConsts = 6;
  static constexpr int kIn = 4;
  static constexpr int kBatch = 1;
  const void* dY; const void* s; const void* xa; const void* xb; void* dxa; void* dxb; int dt;
  const float* coef;
  int has_b;
  int act; float p0, p1;
  template <int VEC>
  __device__ __forceinline__ void init(int c0, int ch, float (&kc)[6][VEC]) const {
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) kc[a][j] = coef[a * ch + (c0 + j < ch ? c0 + j : ch - 1)];
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[4][VEC]) const {
    load_vec<VEC>(dY, sel<DT>(dt), i, in[0]);
    if (act != NPML_ACT_IDENTITY) load_vec<VEC>(s, sel<DT>(dt), i, in[1]);
    load_vec<VEC>(xa, sel<DT>(dt), i, in[2]);
    if (has_b) load_vec<VEC>(xb, sel<DT>(dt), i, in[3]);
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t i, const float (&kc)[6][VEC], float (&in)[4][VEC], float (&)[1][VEC]) const {
    float (&da)[VEC] = in[2];
    float (&db)[VEC] = in[3];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float g = in[0][j];
      if (act == NPML_ACT_RELU) g = in[1][j] > 0.f ? g : 0.f;
      else if (act != NPML_ACT_IDENTITY) g *= act_grad_ool(act, in[1][j], p0, p1);
      if (sel<DT>(dt) == NPML_BF16) g = __bfloat162float(__float2bfloat16_rn(g));
      da[j] = fmaf(kc[0][j], g, fmaf(kc[1][j], da[j], kc[2][j]));
      db[j] = has_b ? fmaf(kc[3][j], g, fmaf(kc[4][j], db[j], kc[5][j])) : g;
    }
    store_vec<VEC>(dxa, sel<DT>(dt), i, da);
    if (dxb) store_vec<VEC>(dxb, sel<DT>(dt), i, db);
  }
};

// per-channel epilogue of the pair backward: parameter gradients from this rank's sums, dX coefficients from the
// sums over all ranks (sync-BN; the same buffer otherwise).  sums = [sum g, sum g*nhat_a, sum g*nhat_b].
__global__ void bn_pair_bwd_finalize_kernel(int ch, int accumulate, double inv_m, const double* __restrict__ local,
                                            const double* __restrict__ global, const float* scaler_a, const float* mean_a,
                                            const float* rstd_a, const float* scaler_b, const float* mean_b,
                                            const float* rstd_b, float* d_scaler_a, float* d_intercept_a,
                                            float* d_scaler_b, float* d_intercept_b, float* coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  const float l0 = (float)local[c], la = (float)local[ch + c], lb = (float)local[2 * ch + c];
  if (d_intercept_a) d_intercept_a[c] = accumulate ? d_intercept_a[c] + l0 : l0;
  if (d_scaler_a) d_scaler_a[c] = accumulate ? d_scaler_a[c] + la : la;
  if (d_intercept_b) d_intercept_b[c] = accumulate ? d_intercept_b[c] + l0 : l0;
  if (d_scaler_b) d_scaler_b[c] = accumulate ? d_scaler_b[c] + lb : lb;
  const double m1 = global[c] * inv_m;
  {
    const double a1 = (double)scaler_a[c] * (double)rstd_a[c];
    const double a2 = -a1 * (global[ch + c] * inv_m) * (double)rstd_a[c];
    coef[c] = (float)a1;
    coef[ch + c] = (float)a2;
    coef[2 * ch + c] = (float)(-a1 * m1 - a2 * (double)mean_a[c]);
  }
  if (scaler_b != nullptr) {
    const double b1 = (double)scaler_b[c] * (double)rstd_b[c];
    const double b2 = -b1 * (global[2 * ch + c] * inv_m) * (double)rstd_b[c];
    coef[3 * ch + c] = (float)b1;
    coef[4 * ch + c] = (float)b2;
    coef[5 * ch + c] = (float)(-b1 * m1 - b2 * (double)mean_b[c]);
  } else {
    coef[3 * ch + c] = 0.f; coef[4 * ch + c] = 0.f; coef[5 * ch + c] = 0.f;
  }
}

// part [nblk][width] -> sums[width] (double), fixed order
__global__ void col_sums_kernel(const double* __restrict__ part, int nblk, int width, double* __restrict__ sums) {
  __shared__ double red[kFinRows][kFinCols + 1];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  const double s = block_col_sum<double>(part, nblk, width, c, c < width, red);
  if (threadIdx.y == 0 && c < width) sums[c] = s;
}

// grid = ceil(ch / 8), block = (8, 128): 8 columns per block and 128 row lanes, so that a thread walks nblk / 128 partial
// rows (5 for the usual 592) instead of nblk / 32 on a quarter of the blocks; fixed order: lane sums, then lanes 0..127
__global__ void __launch_bounds__(1024) db_finalize_kernel(const float* __restrict__ part, int nblk, int ch, int accumulate,
                                                           float* db) {
  __shared__ double red[128][9];
  const int cx = threadIdx.x, sy = threadIdx.y;
  const int c = blockIdx.x * 8 + cx;
  double s = 0.0;
  if (c < ch)
    for (int b = sy; b < nblk; b += 128) s += (double)part[(int64_t)b * ch + c];
  red[sy][cx] = s;
  __syncthreads();
  // 128 -> 8 lanes (each adds 16 lanes in order), then the last 8 in order
  if (sy < 8) {
    double t = 0.0;
    for (int y = 0; y < 16; ++y) t += red[sy * 16 + y][cx];
    red[sy * 16][cx] = t;
  }
  __syncthreads();
  if (sy == 0 && c < ch) {
    double t = 0.0;
    for (int y = 0; y < 8; ++y) t += red[y * 16][cx];
    db[c] = accumulate ? db[c] + (float)t : (float)t;
  }
}

// sums_in != nullptr: the column sums are already reduced (possibly over several ranks, sync-BN) and `part` is unused
__global__ void bn_finalize_kernel(const double* __restrict__ part, int nblk, int ch, double inv_m, float momentum,
                                   float eps, float* running_mean, float* running_var, float* save_mean,
                                   float* save_rstd, const float* scaler, const float* intercept, float* coef,
                                   const double* __restrict__ sums_in) {
  __shared__ double red[kFinRows][kFinCols + 1];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  double s1, s2;
  if (sums_in != nullptr) {
    s1 = c < ch ? sums_in[c] : 0.0;
    s2 = c < ch ? sums_in[ch + c] : 0.0;
  } else {
    s1 = block_col_sum<double>(part, nblk, 2 * ch, c, c < ch, red);
    s2 = block_col_sum<double>(part, nblk, 2 * ch, ch + c, c < ch, red);
  }
  if (threadIdx.y != 0 || c >= ch) return;
  const double mean = s1 * inv_m;
  double var = s2 * inv_m - mean * mean;  // biased (layers/layers.py:1147)
  if (var < 0.0) var = 0.0;
  running_mean[c] = (float)((double)momentum * running_mean[c] + (1.0 - (double)momentum) * mean);
  running_var[c] = (float)((double)momentum * running_var[c] + (1.0 - (double)momentum) * var);
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  save_mean[c] = (float)mean;
  save_rstd[c] = rs;
  if (coef == nullptr) return;
  coef[c] = (float)mean;
  coef[ch + c] = scaler[c] * rs;
  coef[2 * ch + c] = intercept[c];
}

__global__ void bn_running_to_saved_kernel(int ch, float eps, const float* running_mean, const float* running_var,
                                           float* save_mean, float* save_rstd, const float* scaler,
                                           const float* intercept, float* coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  const float rs = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  save_mean[c] = running_mean[c];
  save_rstd[c] = rs;
  coef[c] = running_mean[c];
  coef[ch + c] = scaler[c] * rs;
  coef[2 * ch + c] = intercept[c];
}

// Frozen BatchNorm2D as a per-channel affine y = scale * x + shift (layers/layers.py:1141-1156 with the running statistics):
// scale = scaler / sqrt(running_var + eps), shift = intercept - running_mean * scale.  Feeds the convolution epilogues.
__global__ void bn_fold_kernel(int ch, float eps, const float* running_mean, const float* running_var, const float* scaler,
                               const float* intercept, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  const float rs = (float)(1.0 / sqrt((double)running_var[c] + (double)eps));
  const float sc = scaler[c] * rs;
  scale[c] = sc;
  shift[c] = fmaf(-running_mean[c], sc, intercept[c]);
}

// sums_local / sums_global != nullptr (sync-BN): the parameter gradients accumulate this rank's sums (the gradient
// bucket is all-reduced later), the dX coefficients use the sums over all ranks.  `part` holds RAW partial sums
// [sum dy, sum dy*x] (BnBwdStatsF); sums_local / sums_global hold [sum dy, sum dy*nhat] (npml_bn2d_bwd_stats).
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, int nblk, int ch, int accumulate,
                                       float* d_scaler, float* d_intercept, double inv_m, const float* mean,
                                       const float* rstd, const float* scaler, float* coef,
                                       const double* __restrict__ sums_local, const double* __restrict__ sums_global) {
  __shared__ double red[kFinRows][kFinCols + 1];
  const int c = blockIdx.x * kFinCols + threadIdx.x;
  double s1, s2, l1, l2;
  if (sums_global != nullptr) {
    s1 = c < ch ? sums_global[c] : 0.0;
    s2 = c < ch ? sums_global[ch + c] : 0.0;
    l1 = c < ch ? sums_local[c] : 0.0;
    l2 = c < ch ? sums_local[ch + c] : 0.0;
  } else {
    s1 = l1 = block_col_sum<double>(part, nblk, 2 * ch, c, c < ch, red);
    const double raw = block_col_sum<double>(part, nblk, 2 * ch, ch + c, c < ch, red);
    s2 = l2 = c < ch ? (double)rstd[c] * (raw - (double)mean[c] * s1) : 0.0;
  }
  if (threadIdx.y != 0 || c >= ch) return;
  const float di = (float)l1, ds = (float)l2;
  if (d_intercept) d_intercept[c] = accumulate ? d_intercept[c] + di : di;
  if (d_scaler) d_scaler[c] = accumulate ? d_scaler[c] + ds : ds;
  const double a1 = (double)scaler[c] * (double)rstd[c];
  const double a2 = -a1 * (s2 * inv_m) * (double)rstd[c];
  coef[c] = (float)a1;
  coef[ch + c] = (float)a2;
  coef[2 * ch + c] = (float)(-a1 * (s1 * inv_m) - a2 * (double)mean[c]);
}

// raw [sum dy, sum dy*x] -> [sum dy, sum dy*nhat] (the form that crosses ranks under sync-BN)
__global__ void bwd_stats_fix_kernel(int ch, double* sums, const float* mean, const float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ch) return;
  sums[ch + c] = (double)rstd[c] * (sums[ch + c] - (double)mean[c] * sums[c]);
}

// ---- flat elementwise ------------------------------------------------------------------------------
template <int VEC>
__global__ void cast_kernel(const void* src, int sdt, void* dst, int ddt, int64_t n_vec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float v[VEC];
    load_vec<VEC>(src, sdt, i * VEC, v);
    store_vec<VEC>(dst, ddt, i * VEC, v);
  }
}

struct AddArgs {
  const void* xs[8];
  int n_in;
};

template <int VEC, int DT>
__global__ void add_act_kernel(AddArgs a, int dt_rt, int64_t n_vec, int act, float p0, float p1, void* sum, void* Y) {
  const int dt = sel<DT>(dt_rt);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // two positions per trip, all loads before the first store (the outputs may alias an input for the compiler)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += 2 * stride) {
    const bool two = i + stride < n_vec;
    float s[2][VEC];
    load_vec<VEC>(a.xs[0], dt, i * VEC, s[0]);
    if (two) load_vec<VEC>(a.xs[0], dt, (i + stride) * VEC, s[1]);
    for (int k = 1; k < a.n_in; ++k) {
      float v[2][VEC];
      load_vec<VEC>(a.xs[k], dt, i * VEC, v[0]);
      if (two) load_vec<VEC>(a.xs[k], dt, (i + stride) * VEC, v[1]);
#pragma unroll
      for (int j = 0; j < VEC; ++j) { s[0][j] += v[0][j]; s[1][j] += two ? v[1][j] : 0.f; }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const int64_t pos = (i + u * stride) * VEC;
      // later passes read the stored sum, so activations see the stored (rounded) value
      if (dt == NPML_BF16) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) s[u][j] = __bfloat162float(__float2bfloat16_rn(s[u][j]));
      }
      if (sum) store_vec<VEC>(sum, dt, pos, s[u]);
      if (Y) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) s[u][j] = act_fn(act, s[u][j], p0, p1);
        store_vec<VEC>(Y, dt, pos, s[u]);
      }
    }
  }
}

// Multiply.forward (layers/layers.py:800-822): prod = x0 * x1 * ...; Y = act(prod)
template <int VEC>
__global__ void mul_act_kernel(AddArgs a, int dt, int64_t n_vec, int act, float p0, float p1, void* prod, void* Y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float s[VEC];
    load_vec<VEC>(a.xs[0], dt, i * VEC, s);
    for (int k = 1; k < a.n_in; ++k) {
      float v[VEC];
      load_vec<VEC>(a.xs[k], dt, i * VEC, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) s[j] *= v[j];
    }
    if (dt == NPML_BF16) {                 // later passes read the stored product
#pragma unroll
      for (int j = 0; j < VEC; ++j) s[j] = __bfloat162float(__float2bfloat16_rn(s[j]));
    }
    if (prod) store_vec<VEC>(prod, dt, i * VEC, s);
    if (Y) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) s[j] = act_fn(act, s[j], p0, p1);
      store_vec<VEC>(Y, dt, i * VEC, s);
    }
  }
}

struct MulBwdArgs {
  const void* xs[8];
  void* gs[8];
  int n_in;
};

// Multiply._bwd (layers/layers.py:849-856): g_i = dLdY * act'(prod) * prod_{j != i} x_j
template <int VEC>
__global__ void mul_act_bwd_kernel(MulBwdArgs a, int dt, int64_t n_vec, int act, float p0, float p1, const void* dY,
                                   const void* prod) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float g[VEC];
    load_vec<VEC>(dY, dt, i * VEC, g);
    if (act != NPML_ACT_IDENTITY) {
      float pr[VEC];
      load_vec<VEC>(prod, dt, i * VEC, pr);
#pragma unroll
      for (int j = 0; j < VEC; ++j) g[j] *= act_grad(act, pr[j], p0, p1);
    }
    float x[8][VEC];
    for (int k = 0; k < a.n_in; ++k) load_vec<VEC>(a.xs[k], dt, i * VEC, x[k]);
    for (int k = 0; k < a.n_in; ++k) {
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = g[j];
      for (int q = 0; q < a.n_in; ++q) {
        if (q == k) continue;
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] *= x[q][j];
      }
      store_vec<VEC>(a.gs[k], dt, i * VEC, o);
    }
  }
}

inline int flat_blocks(int64_t n_vec) {
  int64_t b = (n_vec + 255) / 256;
  const int64_t cap = 16LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// clip-norm factor formed on the device (no host read-back of ||g||): t / ||g|| if ||g|| > t else 1
// (optimizers/optimizers.py:136-139), ||g||^2 from npml_sumsq_f32
__device__ __forceinline__ float clip_scale(const double* __restrict__ sumsq, float clip, float gscale) {
  if (sumsq == nullptr) return gscale;
  const double nrm = sqrt(*sumsq);
  return nrm > (double)clip ? (float)((double)clip / nrm) : 1.f;
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, int64_t n, float lr,
                           float momentum, float gscale, const double* __restrict__ sumsq = nullptr, float clip = 0.f) {
  gscale = clip_scale(sumsq, clip, gscale);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float u = momentum * v[i] + lr * (g[i] * gscale);
    v[i] = u;
    p[i] -= u;
  }
}

// AdaGrad / RMSProp / Adam steps on the (all-reduced) fp32 gradient (optimizers/optimizers.py:243-259, 345-361,
// 450-482).  g' = grad * gscale carries the clip-norm factor.  KIND: 1 AdaGrad, 2 RMSProp, 3 Adam.
template <int KIND>
__global__ void adaptive_opt_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ s0,
                                    float* __restrict__ s1, int64_t n, float lr, float a, float oma, float b, float omb,
                                    float eps, float c1, float c2, float gscale,
                                    const double* __restrict__ sumsq = nullptr, float clip = 0.f) {
  gscale = clip_scale(sumsq, clip, gscale);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    if (KIND == 1) {                       // cache += g^2 ; p -= lr * g / (sqrt(cache) + eps)
      const float c = s0[i] + gi * gi;
      s0[i] = c;
      p[i] -= lr * gi / (sqrtf(c) + eps);
    } else if (KIND == 2) {                // cache = decay*cache + (1-decay)*g^2
      const float c = a * s0[i] + oma * gi * gi;
      s0[i] = c;
      p[i] -= lr * gi / (sqrtf(c) + eps);
    } else {                               // Adam: s0 = mean, s1 = var; c1 = 1 - d1^t, c2 = 1 - d2^t
      const float m = a * s0[i] + oma * gi;
      const float v = b * s1[i] + omb * gi * gi;
      s0[i] = m;
      s1[i] = v;
      p[i] -= lr * (m / c1) / (sqrtf(v / c2) + eps);
    }
  }
}

struct SumSqF {
  static constexpr int kConsts = 0;
  static constexpr int kIn = 1;
  static constexpr int kBatch = 4;
  const float* x;
  template <int VEC>
  __device__ __forceinline__ void init(int, int, float (&)[1][VEC]) const {}
  template <int VEC, int DT>
  __device__ __forceinline__ void load(int64_t i, float (&in)[1][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) in[0][j] = x[i + j];
  }
  template <int VEC, int DT>
  __device__ __forceinline__ void finish(int64_t, const float (&)[1][VEC], float (&in)[1][VEC], float (&out)[1][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[0][j] = in[0][j] * in[0][j];
  }
};

// ---- im2col / col2im exports -------------------------------------------------------------------------
__global__ void im2col_kernel(npml_conv_desc d, const float* __restrict__ X, float* __restrict__ Xc) {
  const int64_t L = (int64_t)d.OH * d.OW * d.N;
  const int64_t total = (int64_t)d.fr * d.fc * d.C * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = i / L, z = i - q * L;
    const int c = (int)(q / (d.fr * d.fc));
    const int rt = (int)(q - (int64_t)c * d.fr * d.fc);
    const int r = rt / d.fc, t = rt - r * d.fc;
    const int n = (int)(z % d.N);
    const int64_t pix = z / d.N;
    const int oh = (int)(pix / d.OW), ow = (int)(pix - (int64_t)oh * d.OW);
    const int iy = oh * d.stride + r * d.dil - d.pr1, ix = ow * d.stride + t * d.dil - d.pc1;
    float v = 0.f;
    if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) v = X[(((int64_t)n * d.H + iy) * d.W + ix) * d.C + c];
    Xc[i] = v;
  }
}

__global__ void col2im_kernel(npml_conv_desc d, const float* __restrict__ Xc, float* __restrict__ out) {
  const int64_t L = (int64_t)d.OH * d.OW * d.N;
  const int64_t total = (int64_t)d.N * d.C * d.H * d.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i;
    const int w = (int)(q % d.W); q /= d.W;
    const int h = (int)(q % d.H); q /= d.H;
    const int c = (int)(q % d.C);
    const int n = (int)(q / d.C);
    float s = 0.f;
    for (int r = 0; r < d.fr; ++r) {
      const int ny = h + d.pr1 - r * d.dil;
      if (ny < 0 || ny % d.stride) continue;
      const int oh = ny / d.stride;
      if (oh >= d.OH) continue;
      for (int t = 0; t < d.fc; ++t) {
        const int nx = w + d.pc1 - t * d.dil;
        if (nx < 0 || nx % d.stride) continue;
        const int ow = nx / d.stride;
        if (ow >= d.OW) continue;
        const int64_t row = ((int64_t)c * d.fr + r) * d.fc + t;
        s += Xc[row * L + ((int64_t)oh * d.OW + ow) * d.N + n];
      }
    }
    out[i] = s;
  }
}

}  // namespace

size_t colpass_ws_bytes(int64_t rows, int ch, int nv) {
  ColPlan p = plan_cols(rows, ch, 1);     // vec = 1 gives the largest grid.x any plan for this shape can have
  return align_up((size_t)p.grid.x * nv * ch * sizeof(double), 256) + align_up((size_t)nv * ch * sizeof(double), 256) +
         align_up((size_t)9 * ch * sizeof(float), 256);   // partials, sums, per-channel coefficient table
}

}  // namespace npml

using namespace npml;

extern "C" int npml_dz_prep(int64_t rows, int32_t ch, int32_t act, float p0, float p1, const void* dY,
                            int32_t dy_dtype, const void* Z, int32_t z_dtype, void* dZ, int32_t dz_dtype, float* db,
                            int32_t accumulate, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(dY != nullptr, "dY is NULL");
  NPML_CHECK(act == NPML_ACT_IDENTITY || Z != nullptr, "Z is required for a non-identity activation");
  const bool vec_ok = aligned_for_vec4(dY, dy_dtype) && aligned_for_vec4(Z, z_dtype) && aligned_for_vec4(dZ, dz_dtype);
  ColPlan p = plan_cols(rows, ch, pick_vec(ch, vec_ok, aligned16(dY) && aligned16(Z) && aligned16(dZ)), db != nullptr);
  DzPrepF f{dY, Z, dZ, dy_dtype, z_dtype, dz_dtype, act, p0, p1};
  const int same_dt = (dy_dtype == dz_dtype && (act == NPML_ACT_IDENTITY || z_dtype == dy_dtype)) ? dy_dtype : -1;
  if (!db) return launch_col_pass<1, float>(p, f, (float*)nullptr, st, same_dt);
  const size_t part_bytes = align_up((size_t)p.grid.x * ch * sizeof(float), 256);
  NPML_CHECK(ws != nullptr && ws_bytes >= part_bytes + (size_t)ch * sizeof(double), "workspace too small");
  float* part = (float*)ws;
  if (act == NPML_ACT_IDENTITY && dZ == nullptr) {          // nothing to write: a plain column sum of dLdy
    ColSumF cs{dY, dy_dtype};
    if (int e = launch_col_pass<1, float>(p, cs, part, st, dy_dtype == NPML_F64 ? -1 : dy_dtype)) return e;
  } else if (int e = launch_col_pass<1, float>(p, f, part, st, same_dt)) {
    return e;
  }
  db_finalize_kernel<<<(ch + 7) / 8, dim3(8, 128), 0, st>>>(part, (int)p.grid.x, ch, accumulate, db);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  const int vec = pick_vec(n, aligned_for_vec4(src, src_dtype) && aligned_for_vec4(dst, dst_dtype),
                           aligned16(src) && aligned16(dst));
  if (vec == 8)
    cast_kernel<8><<<flat_blocks(n / 8), 256, 0, st>>>(src, src_dtype, dst, dst_dtype, n / 8);
  else if (vec == 4)
    cast_kernel<4><<<flat_blocks(n / 4), 256, 0, st>>>(src, src_dtype, dst, dst_dtype, n / 4);
  else
    cast_kernel<1><<<flat_blocks(n), 256, 0, st>>>(src, src_dtype, dst, dst_dtype, n);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_add_act_fwd(int32_t n_in, const void* const* xs, int32_t dtype, int64_t n, int32_t act, float p0,
                                float p1, void* sum, void* Y, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(n_in >= 1 && n_in <= 8, "between 1 and 8 inputs are supported");
  if (n <= 0) return 0;
  AddArgs a;
  a.n_in = n_in;
  bool al = aligned_for_vec4(sum, dtype) && aligned_for_vec4(Y, dtype);
  for (int i = 0; i < n_in; ++i) {
    a.xs[i] = xs[i];
    al = al && aligned_for_vec4(xs[i], dtype);
  }
  bool al16 = aligned16(sum) && aligned16(Y);
  for (int i = 0; i < n_in; ++i) al16 = al16 && aligned16(xs[i]);
  const int vec = pick_vec(n, al, al16);
  if (vec == 8 && dtype == NPML_BF16)
    add_act_kernel<8, NPML_BF16><<<flat_blocks(n / 8), 256, 0, st>>>(a, dtype, n / 8, act, p0, p1, sum, Y);
  else if (vec == 8 && dtype == NPML_F32)
    add_act_kernel<8, NPML_F32><<<flat_blocks(n / 8), 256, 0, st>>>(a, dtype, n / 8, act, p0, p1, sum, Y);
  else if (vec == 8)
    add_act_kernel<8, -1><<<flat_blocks(n / 8), 256, 0, st>>>(a, dtype, n / 8, act, p0, p1, sum, Y);
  else if (vec == 4)
    add_act_kernel<4, -1><<<flat_blocks(n / 4), 256, 0, st>>>(a, dtype, n / 4, act, p0, p1, sum, Y);
  else
    add_act_kernel<1, -1><<<flat_blocks(n), 256, 0, st>>>(a, dtype, n, act, p0, p1, sum, Y);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_mul_act_fwd(int32_t n_in, const void* const* xs, int32_t dtype, int64_t n, int32_t act, float p0,
                                float p1, void* prod, void* Y, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(n_in >= 1 && n_in <= 8, "between 1 and 8 inputs are supported");
  if (n <= 0) return 0;
  AddArgs a;
  a.n_in = n_in;
  bool al = aligned_for_vec4(prod, dtype) && aligned_for_vec4(Y, dtype);
  for (int i = 0; i < n_in; ++i) {
    a.xs[i] = xs[i];
    al = al && aligned_for_vec4(xs[i], dtype);
  }
  if (al && n % 4 == 0)
    mul_act_kernel<4><<<flat_blocks(n / 4), 256, 0, st>>>(a, dtype, n / 4, act, p0, p1, prod, Y);
  else
    mul_act_kernel<1><<<flat_blocks(n), 256, 0, st>>>(a, dtype, n, act, p0, p1, prod, Y);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_mul_act_bwd(int32_t n_in, const void* const* xs, void* const* grads, int32_t dtype, int64_t n,
                                int32_t act, float p0, float p1, const void* dY, const void* prod, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(n_in >= 1 && n_in <= 8, "between 1 and 8 inputs are supported");
  NPML_CHECK(dY != nullptr && (act == NPML_ACT_IDENTITY || prod != nullptr), "dY (and prod for an activation) required");
  if (n <= 0) return 0;
  MulBwdArgs a;
  a.n_in = n_in;
  bool al = aligned_for_vec4(dY, dtype) && aligned_for_vec4(prod, dtype);
  for (int i = 0; i < n_in; ++i) {
    a.xs[i] = xs[i];
    a.gs[i] = grads[i];
    al = al && aligned_for_vec4(xs[i], dtype) && aligned_for_vec4(grads[i], dtype);
  }
  if (al && n % 4 == 0)
    mul_act_bwd_kernel<4><<<flat_blocks(n / 4), 256, 0, st>>>(a, dtype, n / 4, act, p0, p1, dY, prod);
  else
    mul_act_bwd_kernel<1><<<flat_blocks(n), 256, 0, st>>>(a, dtype, n, act, p0, p1, dY, prod);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_act_fwd(const void* Z, void* Y, int32_t dtype, int64_t n, int32_t act, float p0, float p1,
                            void* stream) {
  const void* xs[1] = {Z};
  return npml_add_act_fwd(1, xs, dtype, n, act, p0, p1, nullptr, Y, stream);
}

extern "C" int npml_bn2d_fwd(int64_t rows, int32_t ch, const void* x, void* y, int32_t dtype, const float* scaler,
                             const float* intercept, float* running_mean, float* running_var, float momentum,
                             float eps, int32_t use_batch_stats, float* save_mean, float* save_rstd, void* ws,
                             size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(save_mean && save_rstd, "save_mean/save_rstd are required");
  const bool vec_ok = aligned_for_vec4(x, dtype) && aligned_for_vec4(y, dtype);
  const int vec = pick_vec(ch, vec_ok, aligned16(x) && aligned16(y));
  ColPlan p = plan_cols(rows, ch, vec, true);
  const size_t part_bytes = align_up((size_t)p.grid.x * 2 * ch * sizeof(double), 256);
  const size_t sums_bytes = align_up((size_t)2 * ch * sizeof(double), 256);
  NPML_CHECK(ws != nullptr && ws_bytes >= part_bytes + sums_bytes + (size_t)3 * ch * sizeof(float), "workspace too small");
  double* part = (double*)ws;
  double* sums = (double*)((char*)ws + part_bytes);
  float* coef = (float*)((char*)ws + part_bytes + sums_bytes);
  if (use_batch_stats) {
    BnStatsF fs{x, dtype};
    if (int e = launch_col_pass<2, double>(p, fs, part, st, dtype)) return e;
    bn_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, ch, 1.0 / (double)rows, momentum, eps,
                                                                running_mean, running_var, save_mean, save_rstd, scaler,
                                                                intercept, coef, nullptr);
    NPML_LAUNCH_OK();
  } else {
    bn_running_to_saved_kernel<<<(ch + 127) / 128, 128, 0, st>>>(ch, eps, running_mean, running_var, save_mean,
                                                                 save_rstd, scaler, intercept, coef);
    NPML_LAUNCH_OK();
  }
  BnNormF fn{x, y, dtype, coef};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), fn, (float*)nullptr, st, dtype);
}

extern "C" int npml_bn2d_bwd(int64_t rows, int32_t ch, const void* dy, const void* x, void* dx, int32_t dtype,
                             const float* scaler, const float* save_mean, const float* save_rstd, float* d_scaler,
                             float* d_intercept, int32_t accumulate, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  const bool vec_ok = aligned_for_vec4(x, dtype) && aligned_for_vec4(dy, dtype) && aligned_for_vec4(dx, dtype);
  const int vec = pick_vec(ch, vec_ok, aligned16(x) && aligned16(dy) && aligned16(dx));
  ColPlan p = plan_cols(rows, ch, vec, true);
  const size_t part_bytes = align_up((size_t)p.grid.x * 2 * ch * sizeof(double), 256);
  const size_t sums_bytes = align_up((size_t)2 * ch * sizeof(double), 256);
  NPML_CHECK(ws != nullptr && ws_bytes >= part_bytes + sums_bytes + (size_t)5 * ch * sizeof(float), "workspace too small");
  double* part = (double*)ws;
  double* sums = (double*)((char*)ws + part_bytes);
  float* coef = (float*)((char*)ws + part_bytes + sums_bytes);
  BnBwdStatsF fs{dy, x, dtype};
  if (int e = launch_col_pass<2, double>(p, fs, part, st, dtype)) return e;
  bn_bwd_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, ch, accumulate, d_scaler, d_intercept,
                                                                  1.0 / (double)rows, save_mean, save_rstd, scaler, coef,
                                                                  nullptr, nullptr);
  NPML_LAUNCH_OK();
  BnBwdDxF fd{dy, x, dx, dtype, coef};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), fd, (float*)nullptr, st, dtype);
}

extern "C" int npml_bn2d_fold(int32_t ch, const float* scaler, const float* intercept, const float* running_mean,
                              const float* running_var, float eps, float* scale, float* shift, void* stream) {
  if (ch <= 0) return 0;
  NPML_CHECK(scaler && intercept && running_mean && running_var && scale && shift, "NULL pointer");
  bn_fold_kernel<<<(ch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ch, eps, running_mean, running_var, scaler, intercept,
                                                                     scale, shift);
  NPML_LAUNCH_OK();
  return 0;
}

// ---- sync-BN: the same passes split at the point where the per-channel sums cross ranks ---------------------
extern "C" int npml_bn2d_stats(int64_t rows, int32_t ch, const void* x, int32_t dtype, double* sums, void* ws,
                               size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(sums != nullptr, "sums is NULL");
  if (rows <= 0 || ch <= 0) return 0;
  ColPlan p = plan_cols(rows, ch, pick_vec(ch, aligned_for_vec4(x, dtype), aligned16(x)), true);
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)p.grid.x * 2 * ch * sizeof(double), "workspace too small");
  double* part = (double*)ws;
  BnStatsF fs{x, dtype};
  if (int e = launch_col_pass<2, double>(p, fs, part, st, dtype)) return e;
  col_sums_kernel<<<(2 * ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, 2 * ch, sums);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_bn2d_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* x, void* y, int32_t dtype,
                               const float* scaler, const float* intercept, float* running_mean, float* running_var,
                               float momentum, float eps, const double* sums, float* save_mean, float* save_rstd,
                               void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(sums && save_mean && save_rstd && rows_total >= rows, "bad arguments");
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)3 * ch * sizeof(float), "workspace too small");
  float* coef = (float*)ws;
  bn_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(nullptr, 0, ch, 1.0 / (double)rows_total, momentum, eps,
                                                              running_mean, running_var, save_mean, save_rstd, scaler,
                                                              intercept, coef, sums);
  NPML_LAUNCH_OK();
  const int vec = pick_vec(ch, aligned_for_vec4(x, dtype) && aligned_for_vec4(y, dtype), aligned16(x) && aligned16(y));
  BnNormF fn{x, y, dtype, coef};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), fn, (float*)nullptr, st, dtype);
}

extern "C" int npml_bn2d_bwd_stats(int64_t rows, int32_t ch, const void* dy, const void* x, int32_t dtype,
                                   const float* save_mean, const float* save_rstd, double* sums, void* ws,
                                   size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(sums != nullptr, "sums is NULL");
  if (rows <= 0 || ch <= 0) return 0;
  const int vec = pick_vec(ch, aligned_for_vec4(x, dtype) && aligned_for_vec4(dy, dtype), aligned16(x) && aligned16(dy));
  ColPlan p = plan_cols(rows, ch, vec, true);
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)p.grid.x * 2 * ch * sizeof(double), "workspace too small");
  double* part = (double*)ws;
  BnBwdStatsF fs{dy, x, dtype};
  if (int e = launch_col_pass<2, double>(p, fs, part, st, dtype)) return e;
  col_sums_kernel<<<(2 * ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, 2 * ch, sums);
  NPML_LAUNCH_OK();
  bwd_stats_fix_kernel<<<(ch + 127) / 128, 128, 0, st>>>(ch, sums, save_mean, save_rstd);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_bn2d_bwd_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* dy, const void* x, void* dx,
                                   int32_t dtype, const float* scaler, const float* save_mean, const float* save_rstd,
                                   const double* sums_local, const double* sums_global, float* d_scaler,
                                   float* d_intercept, int32_t accumulate, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(sums_local && sums_global && rows_total >= rows, "bad arguments");
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)5 * ch * sizeof(float), "workspace too small");
  float* coef = (float*)ws;
  bn_bwd_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(nullptr, 0, ch, accumulate, d_scaler, d_intercept,
                                                                  1.0 / (double)rows_total, save_mean, save_rstd, scaler,
                                                                  coef, sums_local, sums_global);
  NPML_LAUNCH_OK();
  const bool vec_ok = aligned_for_vec4(x, dtype) && aligned_for_vec4(dy, dtype) && aligned_for_vec4(dx, dtype);
  const int vec = pick_vec(ch, vec_ok, aligned16(x) && aligned16(dy) && aligned16(dx));
  BnBwdDxF fd{dy, x, dx, dtype, coef};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), fd, (float*)nullptr, st, dtype);
}

// BatchNorm2D backward with the activation derivative of the layer in front of it folded into the dX pass
// (BnBwdDxF): dZ = BN'(dy) * act'(z).  zm == nullptr with ReLU: the sign of x itself is the mask.
extern "C" int npml_bn2d_bwd_act(int64_t rows, int64_t rows_total, int32_t ch, const void* dy, const void* x, void* dz,
                                 int32_t dtype, const float* scaler, const float* save_mean, const float* save_rstd,
                                 const double* sums_local, const double* sums_global, float* d_scaler, float* d_intercept,
                                 int32_t accumulate, int32_t act, float p0, float p1, const void* zm, void* ws,
                                 size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(rows_total >= rows, "rows_total < rows");
  NPML_CHECK((sums_local == nullptr) == (sums_global == nullptr), "sums_local and sums_global go together");
  NPML_CHECK(act == NPML_ACT_IDENTITY || act == NPML_ACT_RELU || zm != nullptr, "the pre-activation is required");
  const bool vec_ok = aligned_for_vec4(x, dtype) && aligned_for_vec4(dy, dtype) && aligned_for_vec4(dz, dtype) &&
                      aligned_for_vec4(zm, dtype);
  const int vec = pick_vec(ch, vec_ok, aligned16(x) && aligned16(dy) && aligned16(dz) && aligned16(zm));
  ColPlan p = plan_cols(rows, ch, vec, true);
  const size_t part_bytes = align_up((size_t)p.grid.x * 2 * ch * sizeof(double), 256);
  const size_t sums_bytes = align_up((size_t)2 * ch * sizeof(double), 256);
  NPML_CHECK(ws != nullptr && ws_bytes >= part_bytes + sums_bytes + (size_t)5 * ch * sizeof(float), "workspace too small");
  double* part = (double*)ws;
  float* coef = (float*)((char*)ws + part_bytes + sums_bytes);
  if (sums_global == nullptr) {
    BnBwdStatsF fs{dy, x, dtype};
    if (int e = launch_col_pass<2, double>(p, fs, part, st, dtype)) return e;
  }
  bn_bwd_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, ch, accumulate, d_scaler, d_intercept,
                                                                  1.0 / (double)rows_total, save_mean, save_rstd, scaler, coef,
                                                                  sums_local, sums_global);
  NPML_LAUNCH_OK();
  BnBwdDxF fd{dy, x, dz, dtype, coef, act, p0, p1, zm};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), fd, (float*)nullptr, st, dtype);
}

// per-channel statistics -> (mean, rstd) + running-statistics update, without touching the tensor (the normalisation
// itself happens in a fused consumer, e.g. npml_bn2d_pair_add_act_fwd)
extern "C" int npml_bn2d_finalize(int64_t rows_total, int32_t ch, const double* sums, float* running_mean,
                                  float* running_var, float momentum, float eps, float* save_mean, float* save_rstd,
                                  void* stream) {
  if (ch <= 0) return 0;
  NPML_CHECK(sums && running_mean && running_var && save_mean && save_rstd && rows_total > 0, "bad arguments");
  bn_finalize_kernel<<<(ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, (cudaStream_t)stream>>>(nullptr, 0, ch, 1.0 / (double)rows_total,
                                                                               momentum, eps, running_mean, running_var,
                                                                               save_mean, save_rstd, nullptr, nullptr,
                                                                               nullptr, sums);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_bn2d_pair_add_act_fwd(int64_t rows, int32_t ch, const void* xa, const void* xb, int32_t dtype,
                                          const float* scaler_a, const float* intercept_a, const float* mean_a,
                                          const float* rstd_a, const float* scaler_b, const float* intercept_b,
                                          const float* mean_b, const float* rstd_b, int32_t act, float p0, float p1,
                                          void* sum, void* Y, void* stream) {
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(xa && xb && scaler_a && intercept_a && mean_a && rstd_a && (sum || Y), "NULL argument");
  NPML_CHECK(scaler_b == nullptr || (intercept_b && mean_b && rstd_b), "incomplete second BatchNorm2D");
  const bool vec_ok = aligned_for_vec4(xa, dtype) && aligned_for_vec4(xb, dtype) && aligned_for_vec4(sum, dtype) &&
                      aligned_for_vec4(Y, dtype);
  const int vec = pick_vec(ch, vec_ok, aligned16(xa) && aligned16(xb) && aligned16(sum) && aligned16(Y));
  BnPairAddActF f{xa, xb, sum, Y, dtype, scaler_a, intercept_a, mean_a, rstd_a, scaler_b, intercept_b, mean_b, rstd_b,
                  act, p0, p1};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), f, (float*)nullptr, (cudaStream_t)stream, dtype);
}

extern "C" int npml_bn2d_pair_bwd_stats(int64_t rows, int32_t ch, const void* dY, const void* s, int32_t act, float p0,
                                        float p1, const void* xa, const void* xb, int32_t dtype, const float* mean_a,
                                        const float* rstd_a, const float* mean_b, const float* rstd_b, double* sums,
                                        void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(sums != nullptr, "sums is NULL");
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(dY && xa && mean_a && rstd_a && (act == NPML_ACT_IDENTITY || s), "NULL argument");
  NPML_CHECK((mean_b == nullptr) == (rstd_b == nullptr) && (mean_b == nullptr || xb), "incomplete second BatchNorm2D");
  const bool vec_ok = aligned_for_vec4(dY, dtype) && aligned_for_vec4(s, dtype) && aligned_for_vec4(xa, dtype) &&
                      aligned_for_vec4(xb, dtype);
  // four input tensors: 4 channels per thread keep the pass at 64 registers (4 blocks per SM); with 8 it needs 97 and
  // the lower occupancy costs more bytes in flight than the wider loads bring (measured: 300 us -> see profiles/)
  int vec = pick_vec(ch, vec_ok, aligned16(dY) && aligned16(s) && aligned16(xa) && aligned16(xb));
  if (vec > 4) vec = 4;
  ColPlan p = plan_cols(rows, ch, vec, true);
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)p.grid.x * 3 * ch * sizeof(double), "workspace too small");
  double* part = (double*)ws;
  BnPairBwdStatsF f{dY, s, xa, xb, dtype, mean_b != nullptr ? 1 : 0, act, p0, p1};
  if (int e = launch_col_pass<3, double>(p, f, part, st, dtype)) return e;
  col_sums_kernel<<<(3 * ch + kFinCols - 1) / kFinCols, dim3(kFinCols, kFinRows), 0, st>>>(part, (int)p.grid.x, 3 * ch, sums);
  NPML_LAUNCH_OK();
  pair_stats_fix_kernel<<<(ch + 127) / 128, 128, 0, st>>>(ch, sums, mean_a, rstd_a, mean_b, rstd_b);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_bn2d_pair_bwd_apply(int64_t rows, int64_t rows_total, int32_t ch, const void* dY, const void* s,
                                        int32_t act, float p0, float p1, const void* xa, const void* xb, int32_t dtype,
                                        const float* scaler_a, const float* mean_a, const float* rstd_a,
                                        const float* scaler_b, const float* mean_b, const float* rstd_b,
                                        const double* sums_local, const double* sums_global, void* dxa, void* dxb,
                                        float* d_scaler_a, float* d_intercept_a, float* d_scaler_b, float* d_intercept_b,
                                        int32_t accumulate, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0 || ch <= 0) return 0;
  NPML_CHECK(dY && xa && dxa && scaler_a && mean_a && rstd_a && sums_local && sums_global && rows_total >= rows,
             "bad arguments");
  NPML_CHECK(act == NPML_ACT_IDENTITY || s != nullptr, "the activation needs its argument");
  const bool has_b = scaler_b != nullptr;
  NPML_CHECK(!has_b || (xb && mean_b && rstd_b && dxb), "incomplete second BatchNorm2D");
  NPML_CHECK(ws != nullptr && ws_bytes >= (size_t)9 * ch * sizeof(float), "workspace too small");
  float* coef = (float*)ws;
  bn_pair_bwd_finalize_kernel<<<(ch + 127) / 128, 128, 0, st>>>(ch, accumulate, 1.0 / (double)rows_total, sums_local,
                                                                sums_global, scaler_a, mean_a, rstd_a, scaler_b, mean_b,
                                                                rstd_b, d_scaler_a, d_intercept_a, d_scaler_b,
                                                                d_intercept_b, coef);
  NPML_LAUNCH_OK();
  const bool vec_ok = aligned_for_vec4(dY, dtype) && aligned_for_vec4(s, dtype) && aligned_for_vec4(xa, dtype) &&
                      aligned_for_vec4(xb, dtype) && aligned_for_vec4(dxa, dtype) && aligned_for_vec4(dxb, dtype);
  int vec = pick_vec(ch, vec_ok, aligned16(dY) && aligned16(s) && aligned16(xa) && aligned16(xb) && aligned16(dxa) &&
                                     aligned16(dxb));
  if (vec > 4) vec = 4;                                   // see npml_bn2d_pair_bwd_stats
  BnPairBwdDxF f{dY, s, xa, xb, dxa, dxb, dtype, coef, has_b ? 1 : 0, act, p0, p1};
  return launch_col_pass<0, float>(plan_cols(rows, ch, vec, false), f, (float*)nullptr, st, dtype);
}

extern "C" int npml_im2col(const npml_conv_desc* d, const float* X, float* X_col, void* stream) {
  const int64_t total = (int64_t)d->fr * d->fc * d->C * d->OH * d->OW * d->N;
  if (total <= 0) return 0;
  im2col_kernel<<<flat_blocks(total), 256, 0, (cudaStream_t)stream>>>(*d, X, X_col);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_col2im(const npml_conv_desc* d, const float* X_col, float* X_nchw, void* stream) {
  const int64_t total = (int64_t)d->N * d->C * d->H * d->W;
  if (total <= 0) return 0;
  col2im_kernel<<<flat_blocks(total), 256, 0, (cudaStream_t)stream>>>(*d, X_col, X_nchw);
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_sgd_update(float* param, const float* grad, float* velocity, int64_t n, float lr, float momentum,
                               float grad_scale, void* stream) {
  if (n <= 0) return 0;
  sgd_kernel<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(param, grad, velocity, n, lr, momentum, grad_scale);
  NPML_LAUNCH_OK();
  return 0;
}

// hyper-parameters arrive as doubles so that 1 - decay is formed in double (1 - 0.999f is off by 5e-5 in fp32)
static int optimizer_update(int32_t kind, float* param, const float* grad, float* state0, float* state1, int64_t n,
                            double lr_d, double a_d, double b_d, double eps_d, double c1_d, double c2_d,
                            double grad_scale_d, const double* sumsq, double clip_d, void* stream) {
  const float lr = (float)lr_d, a = (float)a_d, b = (float)b_d, eps = (float)eps_d, c1 = (float)c1_d, c2 = (float)c2_d;
  const float oma = (float)(1.0 - a_d), omb = (float)(1.0 - b_d), grad_scale = (float)grad_scale_d, clip = (float)clip_d;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return 0;
  NPML_CHECK(param && grad && state0, "NULL buffer");
  const int blocks = flat_blocks(n);
  switch (kind) {
    case NPML_OPT_SGD:
      sgd_kernel<<<blocks, 256, 0, st>>>(param, grad, state0, n, lr, a, grad_scale, sumsq, clip);
      break;
    case NPML_OPT_ADAGRAD:
      adaptive_opt_kernel<1><<<blocks, 256, 0, st>>>(param, grad, state0, state1, n, lr, a, oma, b, omb, eps, c1, c2,
                                                      grad_scale, sumsq, clip);
      break;
    case NPML_OPT_RMSPROP:
      adaptive_opt_kernel<2><<<blocks, 256, 0, st>>>(param, grad, state0, state1, n, lr, a, oma, b, omb, eps, c1, c2,
                                                      grad_scale, sumsq, clip);
      break;
    case NPML_OPT_ADAM:
      NPML_CHECK(state1 != nullptr, "Adam needs two state buffers");
      adaptive_opt_kernel<3><<<blocks, 256, 0, st>>>(param, grad, state0, state1, n, lr, a, oma, b, omb, eps, c1, c2,
                                                      grad_scale, sumsq, clip);
      break;
    default:
      NPML_CHECK(false, "unknown optimizer kind");
  }
  NPML_LAUNCH_OK();
  return 0;
}

extern "C" int npml_optimizer_update(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                                     int64_t n, double lr, double a, double b, double eps, double c1, double c2,
                                     double grad_scale, void* stream) {
  return optimizer_update(kind, param, grad, state0, state1, n, lr, a, b, eps, c1, c2, grad_scale, nullptr, 0.0, stream);
}

extern "C" int npml_optimizer_update_clipped(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                                             int64_t n, double lr, double a, double b, double eps, double c1, double c2,
                                             const double* grad_sumsq, double clip_norm, void* stream) {
  NPML_CHECK(grad_sumsq != nullptr && clip_norm > 0.0, "clip-norm needs ||g||^2 on the device and a positive threshold");
  return optimizer_update(kind, param, grad, state0, state1, n, lr, a, b, eps, c1, c2, 1.0, grad_sumsq, clip_norm, stream);
}

extern "C" int npml_sumsq_f32(const float* buf, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NPML_CHECK(n > 0 && out, "bad arguments");
  // view the buffer as [n][1]: one column, reduced over rows in a fixed order
  ColPlan p = plan_cols(n, 1, 1);
  const size_t part_bytes = align_up((size_t)p.grid.x * sizeof(double), 256);
  NPML_CHECK(ws != nullptr && ws_bytes >= part_bytes, "workspace too small");
  SumSqF f{buf};
  if (int e = launch_col_pass<1, double>(p, f, (double*)ws, st)) return e;
  return launch_col_final<double>((const double*)ws, (int)p.grid.x, 1, out, st);
}
