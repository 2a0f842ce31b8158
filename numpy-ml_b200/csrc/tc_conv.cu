// tcgen05 implicit-GEMM convolution for sm_100a: Conv2D.forward / Deconv2D dX (form 0) and
// Conv2D dX / Deconv2D.forward (form 1, the transposed gather form -- col2im folded in, no atomics).
//
//   D[pixel, co] = sum_{tap, ci} A_tap[pixel, ci] * Wt[tap][co][ci]
//
// * no im2col is materialised: for every filter tap the TMA loads the 128-pixel x 64-channel (128 B
//   per pixel, 128B-swizzled) window of the NHWC activation tensor that the tap touches; zero padding is
//   the TMA's out-of-bounds fill.  A 128-row M tile is a (TN images) x (TH rows) x (TW cols) patch of
//   output pixels, so the window of a tap is one 4-D box.
// * stride s > 1 is handled by parity decomposition: the input (form 0) or the output (form 1) is
//   viewed as s*s interleaved sub-tensors, each with its own tensor map / pointer, so every box is
//   dense and every output pixel visits only the taps that reach it.
// * tcgen05.mma (M=128, N=BN, K=32 bytes) accumulates in TMEM (fp32); warp 0 = TMA producer,
//   warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue (tcgen05.ld -> +bias -> act -> global).
#include <vector>

#include "npml_common.cuh"
#include "tc_ptx.cuh"

namespace npml {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// rank-`rank` tensor map over `base`; dims/strides innermost first, strides in bytes (stride[0] implicit).
int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b) {
  EncodeTiledFn enc = get_encode_tiled();
  NPML_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_b[i + 1];
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base,
                   gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
             (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    set_error(buf);
    return 5;
  }
  return 0;
}

// also used by tc_wgrad.cu
int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);

namespace {

constexpr int kMaxTaps = 64;
constexpr int kMaxClasses = 9;
constexpr int kMaxMaps = 9;
constexpr int kRowBytes = 128;       // one K block = 128 bytes per row (64 bf16 / 32 fp32)
constexpr int kATileBytes = 128 * kRowBytes;
constexpr int kThreads = 192;

struct Tap {
  int16_t dy, dx;   // coordinate offset (in the tensor-map's own row/col units) relative to the tile origin
  int16_t amap;     // which activation tensor map
  int16_t wtap;     // which filter tap of Wt
};

struct ClassInfo {
  int oh, ow;           // output pixels of this parity class
  int y0, x0, so;       // output pixel (y, x) = (y0 + so*i, x0 + so*j)
  int tap_begin, ntaps;
  int tiles_x, tiles_y;
};

struct TcConvParams {
  CUtensorMap amap[kMaxMaps];
  CUtensorMap wmap;
  ClassInfo cls[kMaxClasses];
  Tap taps[kMaxTaps];
  int N, OH, OW, CO;    // full output tensor
  int TW, TH, TN;       // patch of output pixels per M tile (TW*TH*TN <= 128)
  int tiles_n;
  int BN, stages, nkb;  // N tile, pipeline depth, K blocks per tap
  int kelems;           // elements per K block (64 bf16 / 32 fp32)
  uint32_t idesc, tmem_cols;
  int wait_ns;                 // sleep between polls of the epilogue warps (npml::wait_ns())
  int act;
  float p0, p1;
};

// ---- weight repack: w (strided fp32) -> Wt[tap][co][ci] in the operand type --------------------------
template <typename T>
__global__ void repack_weights_kernel(const float* __restrict__ w, T* __restrict__ wt, int taps, int CO, int CI,
                                      int64_t w_tap, int64_t w_ci, int64_t w_co) {
  const int64_t total = (int64_t)taps * CO * CI;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % CI);
    const int64_t q = i / CI;
    const int co = (int)(q % CO);
    const int tap = (int)(q / CO);
    st_f(wt, i, w[tap * w_tap + ci * w_ci + co * w_co]);
  }
}

// Same repack when the SOURCE is contiguous along co (w_co == 1: the forward convolutions): a per-tap CI x CO -> CO x CI
// transpose.  32 x 32 tiles through shared memory so that both the reads (along co) and the writes (along ci) are
// coalesced; the plain kernel above reads with a stride of CO floats there (6 us instead of 2 for cfg5's 2 MB).
template <typename T>
__global__ void __launch_bounds__(256) repack_weights_t_kernel(const float* __restrict__ w, T* __restrict__ wt, int CO, int CI,
                                                               int64_t w_tap, int64_t w_ci) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z, ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ci = ci0 + ty + 8 * k, co = co0 + tx;
    tile[ty + 8 * k][tx] = (ci < CI && co < CO) ? w[tap * w_tap + ci * w_ci + co] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co = co0 + ty + 8 * k, ci = ci0 + tx;
    if (co < CO && ci < CI) st_f(wt, ((int64_t)tap * CO + co) * CI + ci, tile[tx][ty + 8 * k]);
  }
}

// ---- the kernel ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) tc_gather_conv_kernel(const __grid_constant__ TcConvParams P,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ post_scale,
                                                                  const float* __restrict__ post_shift, T* __restrict__ Z,
                                                                  T* __restrict__ Y) {
  constexpr bool kTf32 = sizeof(T) == 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring, then the barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = P.BN * kRowBytes;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + P.stages;
  uint64_t* tmem_full_bar = empty_bar + P.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  // warp index broadcast from lane 0: tells ptxas it is warp-uniform, so the role branches are uniform branches and
  // everything the MMA issuer computes stays on the uniform datapath (no R2UR per descriptor)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const ClassInfo& C = P.cls[blockIdx.z];

  // tile -> (image block, patch row, patch col) of this class
  int tile = blockIdx.x;
  const int tiles_per_img = C.tiles_x * C.tiles_y;
  if (tile >= tiles_per_img * P.tiles_n) return;           // classes smaller than class 0 (uniform exit)
  const int tn = tile / tiles_per_img;
  tile -= tn * tiles_per_img;
  const int ty = tile / C.tiles_x, tx = tile - ty * C.tiles_x;
  const int n0 = tn * P.TN, oy0 = ty * P.TH, ox0 = tx * P.TW;
  const int co0 = blockIdx.y * P.BN;
  const int nk = C.ntaps * P.nkb;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < P.stages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, P.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.wmap);
      const uint32_t tx_bytes = (uint32_t)(P.TW * P.TH * P.TN * kRowBytes + b_tile_bytes);
      uint32_t s = 0, ph = 0;     // ring position / phase (counters: no division in the loop)
      for (int t = 0; t < C.ntaps; ++t) {
        const Tap tap = P.taps[C.tap_begin + t];
        for (int kb = 0; kb < P.nkb; ++kb) {
          ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          ptx::mbar_expect_tx(&full_bar[s], tx_bytes);
          ptx::tma_load_4d(a_dst, &P.amap[tap.amap], &full_bar[s], kb * P.kelems, ox0 + tap.dx, oy0 + tap.dy, n0);
          ptx::tma_load_3d(a_dst + kATileBytes, &P.wmap, &full_bar[s], kb * P.kelems, co0, tap.wtap);
          if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp walks the loop in convergent control flow (descriptors stay in uniform
    // registers), one elected lane issues =====
    const bool leader = ptx::elect_one();
    constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);
    const uint32_t base16 = ((ptx::smem_u32(smem_raw) + 1023u) & ~1023u) >> 4;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, stages = (uint32_t)P.stages, idesc = P.idesc;
    uint32_t s = 0, ph = 0;
    for (int k = 0; k < nk; ++k) {
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after();
      const uint64_t adesc = kDesc64 + (base16 + s * stage16);
      const uint64_t bdesc = adesc + (uint64_t)(kATileBytes >> 4);
      if (leader) {
#pragma unroll
        for (int j = 0; j < 4; ++j)   // 4 x 32-byte K slices inside the 128-byte swizzle atom
          ptx::mma_ss<kTf32>(tmem_base, adesc + 2 * j, bdesc + 2 * j, idesc, (k | j) ? 1u : 0u);
        ptx::tc_commit(&empty_bar[s]);   // frees the smem slot once these MMAs have read it
      }
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (leader) ptx::tc_commit(tmem_full_bar);     // accumulator complete
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, act) -> global =====
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;                  // row of the M tile = pixel of the patch
    const int pw = m % P.TW;
    const int ph_ = (m / P.TW) % P.TH;
    const int pn = m / (P.TW * P.TH);
    const int n = n0 + pn, oy = oy0 + ph_, ox = ox0 + pw;
    const bool row_ok = (m < P.TW * P.TH * P.TN) && n < P.N && oy < C.oh && ox < C.ow;
    const int64_t pix = ((int64_t)n * P.OH + (C.y0 + (int64_t)C.so * oy)) * P.OW + (C.x0 + (int64_t)C.so * ox);
    if (nk > 0) {
      ptx::mbar_wait_sleep(tmem_full_bar, 0, (uint32_t)P.wait_ns);
      ptx::tc_fence_after();
    }
    for (int c = 0; c < P.BN; c += 32) {
      float v[32];
      if (nk > 0) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      const int co = co0 + c;
      const int n_valid = P.CO - co;              // columns of this chunk inside the tensor
      if (row_ok && n_valid > 0) {
        if (bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (j < n_valid) ? bias[co + j] : 0.f;
        }
        if (Z) store_row_chunk<T>(Z + pix * P.CO + co, v, n_valid);
        if (Y) {
          act_apply32(P.act, P.p0, P.p1, v);
          if (post_scale) {                       // frozen BatchNorm2D folded into the epilogue
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < n_valid) v[j] = fmaf(v[j], post_scale[co + j], post_shift[co + j]);
          }
          store_row_chunk<T>(Y + pix * P.CO + co, v, n_valid);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
inline int posmod(int a, int b) { return ((a % b) + b) % b; }
inline int ceildiv(int a, int b) { return (a + b - 1) / b; }

// patch (TW, TH, TN) with TW*TH*TN <= 128 minimising the number of M tiles
void choose_patch(int ow, int oh, int n, int& TW, int& TH, int& TN) {
  long best = -1;
  for (int tw = 1; tw <= 128 && tw <= ow; ++tw) {
    for (int th = 1; th * tw <= 128 && th <= oh; ++th) {
      int tn = 128 / (tw * th);
      if (tn > n) tn = n;
      if (tn > 256) tn = 256;
      const long tiles = (long)ceildiv(ow, tw) * ceildiv(oh, th) * ceildiv(n, tn);
      // prefer fewer tiles, then wider rows (longer contiguous runs per TMA request)
      const long score = tiles * 1024 - tw;
      if (best < 0 || score < best) { best = score; TW = tw; TH = th; TN = tn; }
    }
  }
}

struct Plan {
  bool ok = false;
  TcConvParams P;
  dim3 grid;
  size_t smem = 0;
  size_t wt_bytes = 0;
  std::vector<int> zero_classes;  // form-1 parity classes no tap reaches (output must be zero + bias)
};

// Fills everything except the tensor maps' base pointers (those need the real buffers).
bool plan_geometry(const GatherConv& g, int mode, Plan& pl, int force_bn = 0) {
  const bool bf16 = mode == NPML_MODE_BF16;
  const int es = bf16 ? 2 : 4;
  const int vec = 16 / es;
  if (mode != NPML_MODE_BF16 && mode != NPML_MODE_TF32) return false;
  if (g.CI % vec || g.CO % vec) return false;
  if (g.s < 1 || g.s > 3 || g.fr * g.fc > kMaxTaps) return false;
  if (g.N < 1) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  TcConvParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.N = g.N; P.OH = g.OH; P.OW = g.OW; P.CO = g.CO;
  P.kelems = kRowBytes / es;
  P.nkb = ceildiv(g.CI, P.kelems);
  P.act = g.act; P.p0 = g.p0; P.p1 = g.p1;
  int bn = (g.CO + 15) / 16 * 16;
  if (bn > 256) bn = (g.CO % 256 == 0 || g.CO > 512) ? 256 : 128;
  if (const char* e = std::getenv("NPML_TC_BN")) {                       // experiment knob: N tile of the per-tile kernel
    const int v = atoi(e);
    if (v >= 16 && v <= 256 && v % 16 == 0 && v < bn) bn = v;
  }
  if (force_bn > 0 && force_bn < bn) bn = force_bn;
  P.BN = bn;
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  P.tmem_cols = cols;
  P.idesc = ptx::instr_desc(bf16 ? 1 : 2, 0, 0, 128, bn);
  P.wait_ns = wait_ns();
  const int stage_bytes = kATileBytes + bn * kRowBytes;
  P.stages = bn <= 64 ? 4 : (bn <= 128 ? 3 : 4);
  pl.smem = (size_t)P.stages * stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  pl.wt_bytes = align_up((size_t)g.fr * g.fc * g.CO * g.CI * es, 256);

  int ncls = 0, ntap = 0, max_tiles = 0;
  if (g.form == 0) {
    ClassInfo& C = P.cls[0];
    C.oh = g.OH; C.ow = g.OW; C.y0 = 0; C.x0 = 0; C.so = 1; C.tap_begin = 0;
    for (int r = 0; r < g.fr; ++r)
      for (int t = 0; t < g.fc; ++t) {
        const int offy = r * g.D - g.pr, offx = t * g.D - g.pc;
        const int py = posmod(offy, g.s), px = posmod(offx, g.s);
        if (py >= g.IH || px >= g.IW) continue;      // that parity plane is empty: the tap only sees padding
        Tap& tp = P.taps[ntap++];
        tp.dy = (int16_t)floordiv(offy, g.s);
        tp.dx = (int16_t)floordiv(offx, g.s);
        tp.amap = (int16_t)(py * g.s + px);
        tp.wtap = (int16_t)(r * g.fc + t);
      }
    C.ntaps = ntap;
    ncls = 1;
  } else {
    for (int qy = 0; qy < g.s; ++qy)
      for (int qx = 0; qx < g.s; ++qx) {
        const int y0 = posmod(qy - g.pr, g.s), x0 = posmod(qx - g.pc, g.s);
        if (y0 >= g.OH || x0 >= g.OW) continue;
        ClassInfo& C = P.cls[ncls];
        C.y0 = y0; C.x0 = x0; C.so = g.s;
        C.oh = ceildiv(g.OH - y0, g.s); C.ow = ceildiv(g.OW - x0, g.s);
        C.tap_begin = ntap;
        for (int r = 0; r < g.fr; ++r) {
          if (posmod(r * g.D, g.s) != qy) continue;
          for (int t = 0; t < g.fc; ++t) {
            if (posmod(t * g.D, g.s) != qx) continue;
            Tap& tp = P.taps[ntap++];
            tp.dy = (int16_t)((y0 + g.pr - r * g.D) / g.s);   // exact by construction
            tp.dx = (int16_t)((x0 + g.pc - t * g.D) / g.s);
            tp.amap = 0;
            tp.wtap = (int16_t)(r * g.fc + t);
          }
        }
        C.ntaps = ntap - C.tap_begin;
        ++ncls;
      }
  }
  if (ncls == 0) return false;
  // one patch shape for all classes, chosen for the largest class
  int mh = 0, mw = 0;
  for (int c = 0; c < ncls; ++c) { if (P.cls[c].oh > mh) mh = P.cls[c].oh; if (P.cls[c].ow > mw) mw = P.cls[c].ow; }
  choose_patch(mw, mh, g.N, P.TW, P.TH, P.TN);
  P.tiles_n = ceildiv(g.N, P.TN);
  for (int c = 0; c < ncls; ++c) {
    P.cls[c].tiles_x = ceildiv(P.cls[c].ow, P.TW);
    P.cls[c].tiles_y = ceildiv(P.cls[c].oh, P.TH);
    const int tiles = P.cls[c].tiles_x * P.cls[c].tiles_y * P.tiles_n;
    if (tiles > max_tiles) max_tiles = tiles;
  }
  pl.grid = dim3((unsigned)max_tiles, (unsigned)ceildiv(g.CO, bn), (unsigned)ncls);
  // N = 256 tiles with fewer CTAs than SMs (cfg3 forward: 121): two N = 128 tiles per M tile fill the machine with two
  // CTAs per SM and six pipeline stages per SM instead of four (measured 24.4 -> 22.4 us; N = 64 loses, 26.5 us)
  if (bn == 256 && force_bn == 0 && (int64_t)max_tiles * ncls < num_sms()) return plan_geometry(g, mode, pl, 128);
  pl.ok = true;
  return true;
}

}  // namespace

bool tc_conv_supported(const GatherConv& g, int mode) {
  Plan pl;
  return plan_geometry(g, mode, pl);
}

size_t tc_conv_ws_bytes(const GatherConv& g, int mode) {
  Plan pl;
  if (!plan_geometry(g, mode, pl)) return 0;
  return pl.wt_bytes + 256;
}

// w (strided fp32) -> Wt[tap][co][ci] in the operand type (also used by tc_slab_conv.cu)
int launch_repack_weights(const GatherConv& g, bool bf16, const float* w, void* wt, cudaStream_t st) {
  if (g.weights_packed) return 0;      // the caller kept the packed copy of an earlier call (NPML_FLAG_WEIGHTS_PACKED)
  const int64_t total = (int64_t)g.fr * g.fc * g.CO * g.CI;
  if (g.w_co == 1 && g.CI >= 16 && g.fr * g.fc <= 65535) {            // source contiguous along co: tiled transpose
    const dim3 grid((unsigned)((g.CO + 31) / 32), (unsigned)((g.CI + 31) / 32), (unsigned)(g.fr * g.fc));
    if (bf16)
      repack_weights_t_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(w, (__nv_bfloat16*)wt, g.CO, g.CI, g.w_tap, g.w_ci);
    else
      repack_weights_t_kernel<float><<<grid, 256, 0, st>>>(w, (float*)wt, g.CO, g.CI, g.w_tap, g.w_ci);
    NPML_LAUNCH_OK();
    return 0;
  }
  int blocks = (int)((total + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  if (bf16)
    repack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)wt, g.fr * g.fc, g.CO, g.CI, g.w_tap,
                                                                 g.w_ci, g.w_co);
  else
    repack_weights_kernel<float><<<blocks, 256, 0, st>>>(w, (float*)wt, g.fr * g.fc, g.CO, g.CI, g.w_tap, g.w_ci, g.w_co);
  NPML_LAUNCH_OK();
  return 0;
}

template <typename T>
static int launch_tc(const GatherConv& g, Plan& pl, const void* in, const float* w, const float* bias, void* Z, void* Y,
                     void* ws, cudaStream_t st) {
  const bool bf16 = sizeof(T) == 2;
  const int es = sizeof(T);
  TcConvParams& P = pl.P;
  T* wt = reinterpret_cast<T*>(ws);
  if (int e = launch_repack_weights(g, bf16, w, wt, st)) return e;
  // weights: (ci, co, tap)
  {
    const uint64_t dims[3] = {(uint64_t)g.CI, (uint64_t)g.CO, (uint64_t)(g.fr * g.fc)};
    const uint64_t strides[3] = {(uint64_t)es, (uint64_t)g.CI * es, (uint64_t)g.CI * g.CO * es};
    const uint32_t box[3] = {(uint32_t)P.kelems, (uint32_t)P.BN, 1};
    if (int e = make_tmap(&P.wmap, bf16, wt, 3, dims, strides, box, false)) return e;
  }
  // activations: one map per input parity plane (form 0) or a single plain map (form 1)
  const int sp = (g.form == 0) ? g.s : 1;
  for (int py = 0; py < sp; ++py)
    for (int px = 0; px < sp; ++px) {
      if (py >= g.IH || px >= g.IW) continue;
      const uint64_t hh = (uint64_t)ceildiv(g.IH - py, sp), ww = (uint64_t)ceildiv(g.IW - px, sp);
      const uint64_t dims[4] = {(uint64_t)g.CI, ww, hh, (uint64_t)g.N};
      const uint64_t strides[4] = {(uint64_t)es, (uint64_t)sp * g.CI * es, (uint64_t)sp * g.IW * g.CI * es,
                                   (uint64_t)g.IH * g.IW * g.CI * es};
      const uint32_t box[4] = {(uint32_t)P.kelems, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
      char* base = (char*)in + ((size_t)py * g.IW + px) * g.CI * es;
      if (int e = make_tmap(&P.amap[py * sp + px], bf16, base, 4, dims, strides, box, false)) return e;
    }
  auto kern = tc_gather_conv_kernel<T>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[bf16 ? 0 : 1]) {
    NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[bf16 ? 0 : 1] = true;
  }
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, bias, g.post_scale, g.post_shift, (T*)Z, (T*)Y);
  NPML_LAUNCH_OK();
  return 0;
}

int tc_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z, void* Y,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  Plan pl;
  NPML_CHECK(plan_geometry(g, mode, pl), "shape not supported by the tensor-core path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.wt_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 127) == 0,
             "activation / workspace pointers must be 16 / 128 byte aligned");
  if (mode == NPML_MODE_BF16) return launch_tc<__nv_bfloat16>(g, pl, in, w, bias, Z, Y, ws, st);
  return launch_tc<float>(g, pl, in, w, bias, Z, Y, ws, st);
}

}  // namespace npml
