// tcgen05 weight-gradient kernel for sm_100a (Conv2D / Deconv2D dW, layers/layers.py:3106-3110, 3557-3559):
//
//   dW[(tap, ci), co] = sum_{pixel} A_tap[pixel, ci] * G[pixel, co]
//
// The reduction runs over output pixels, so both operands are "MN-major": a 128-pixel x 128-byte
// TMA tile (the same shifted activation window the forward uses, and the matching dZ window) is fed to
// tcgen05.mma as K = pixels without any transpose.  An M tile of 128 rows is made of 128/kelems
// (tap, channel-block) windows placed 16 KB apart (the descriptor's leading-dimension offset).
// The pixel range is split over gridDim.z; every split writes an fp32 partial to the workspace and a
// second kernel adds the partials in a fixed order (deterministic, no atomics).
#include "npml_common.cuh"
#include "tc_ptx.cuh"

namespace npml {

int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);
int launch_wgrad_reduce(const float* part, int splits, int M, int CO, int CI, int64_t o_tap, int64_t o_ci,
                        int64_t o_co, int accumulate, float* dW, cudaStream_t st, const float* part_db = nullptr,
                        float* db = nullptr);

namespace {

constexpr int kMaxBlocks = 256;      // (tap, channel-block) windows
constexpr int kMaxMaps = 9;
constexpr int kRowBytes = 128;
constexpr int kTileBytes = 128 * kRowBytes;   // one 128-pixel window
constexpr int kThreads = 192;

struct ABlock {
  int16_t dy, dx;    // window offset in the parity plane's coordinates
  int16_t amap;      // parity plane
  int16_t tap;       // filter tap
  int32_t c0;        // first channel of the block
};

struct TcWgradParams {
  CUtensorMap amap[kMaxMaps];
  CUtensorMap gmap;
  ABlock blocks[kMaxBlocks];
  int nblocks;            // total A windows
  int mb, nb;             // windows per M tile / per N tile
  int kelems;             // channels per window (64 bf16 / 32 fp32)
  int CI, CO, Mrows;      // Mrows = taps * CI
  int N, OH, OW;          // pixel space that is reduced over (the `grad` tensor)
  int TW, TH, TN, tiles_x, tiles_y, tiles_n, ntiles;
  int tiles_per_split;
  int stages, rows_used, kinstr_bytes, mmas_per_tile;
  uint32_t idesc, tmem_cols;
  int wait_ns;                 // sleep between polls of the epilogue warps (npml::wait_ns())
};

template <typename T>
__global__ void __launch_bounds__(kThreads) tc_wgrad_kernel(const __grid_constant__ TcWgradParams P,
                                                            float* __restrict__ part, float* __restrict__ part_db) {
  constexpr bool kTf32 = sizeof(T) == 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = (P.mb + P.nb) * kTileBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + P.stages;
  uint64_t* tmem_full_bar = empty_bar + P.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  // warp index broadcast from lane 0: tells ptxas it is warp-uniform, so the role branches are uniform branches and
  // everything the MMA issuer computes stays on the uniform datapath (no R2UR per descriptor)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int blk0 = blockIdx.x * P.mb;                       // first A window of this M tile
  const int nblk = min(P.mb, P.nblocks - blk0);             // windows that exist
  const int co0 = blockIdx.y * P.nb * P.kelems;
  const int nbn = min(P.nb, (P.CO - co0 + P.kelems - 1) / P.kelems);   // N windows that exist
  const int t_begin = blockIdx.z * P.tiles_per_split;
  const int t_end = min(P.ntiles, t_begin + P.tiles_per_split);
  const int nk = max(0, t_end - t_begin);

  // rows of a window that the TMA never writes must be zero: they are part of the reduction
  if (P.rows_used < 128 || nblk < P.mb || nbn < P.nb) {
    uint4* p = reinterpret_cast<uint4*>(smem);
    const int n16 = P.stages * stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += kThreads) p[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  // the db window (tap < 0) is never loaded: it is a constant block of ones in every stage, so the rows of the
  // accumulator it owns receive sum_pixels 1 * dZ[pixel, co] = db (pixel rows the TMA leaves zero add nothing)
  int n_tma = 0;
  for (int b = 0; b < nblk; ++b) {
    if (P.blocks[blk0 + b].tap >= 0) { ++n_tma; continue; }
    __syncthreads();
    const uint32_t one = kTf32 ? 0x3F800000u : 0x3F803F80u;
    for (int st = 0; st < P.stages; ++st) {
      uint32_t* q = reinterpret_cast<uint32_t*>(smem + (size_t)st * stage_bytes + (size_t)b * kTileBytes);
      for (int i = threadIdx.x; i < kTileBytes / 4; i += kThreads) q[i] = one;
    }
    ptx::fence_proxy_async();
  }
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
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.gmap);
      const uint32_t tx_bytes = (uint32_t)((n_tma + nbn) * P.rows_used * kRowBytes);
      const int per_img = P.tiles_x * P.tiles_y;
      uint32_t s = 0, ph = 0;     // ring position / phase
      for (int k = 0; k < nk; ++k) {
        int tile = t_begin + k;
        const int tn = tile / per_img;
        tile -= tn * per_img;
        const int ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
        const int n0 = tn * P.TN, oy0 = ty * P.TH, ox0 = tx * P.TW;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* dst = smem + (size_t)s * stage_bytes;
        ptx::mbar_expect_tx(&full_bar[s], tx_bytes);
        for (int b = 0; b < nblk; ++b) {
          const ABlock ab = P.blocks[blk0 + b];
          if (ab.tap < 0) continue;
          ptx::tma_load_4d(dst + b * kTileBytes, &P.amap[ab.amap], &full_bar[s], ab.c0, ox0 + ab.dx, oy0 + ab.dy, n0);
        }
        for (int b = 0; b < nbn; ++b)
          ptx::tma_load_4d(dst + (P.mb + b) * kTileBytes, &P.gmap, &full_bar[s], co0 + b * P.kelems, ox0, oy0, n0);
        if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: convergent loop (uniform-register descriptors), one elected lane issues.
    // MN-major: LBO = distance between 128-byte column blocks (one window = 16 KB), SBO = distance between
    // swizzle-atom groups of pixel rows.  16-bit operands use the plain 128B swizzle (8-row atoms, 1 KB);
    // MN-major tf32 operands only exist in the 128B swizzle with 32-byte granularity (4-row atoms, 512 B),
    // which the TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (layout type 1 = SWIZZLE_128B_BASE32B).
    const bool leader = ptx::elect_one();
    constexpr uint64_t kDescMN = ((uint64_t)(((kTf32 ? 512u : 1024u) >> 4) | (1u << 14) | ((kTf32 ? 1u : 2u) << 29)) << 32) |
                                 ((uint64_t)(kTileBytes >> 4) << 16);
    const uint32_t base16 = ((ptx::smem_u32(smem_raw) + 1023u) & ~1023u) >> 4;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, stages = (uint32_t)P.stages, idesc = P.idesc;
    const uint32_t step = (uint32_t)P.kinstr_bytes >> 4, b_off16 = (uint32_t)(P.mb * kTileBytes) >> 4;
    const int mmas = P.mmas_per_tile;
    uint32_t s = 0, ph = 0;
    for (int k = 0; k < nk; ++k) {
      ptx::mbar_wait(&full_bar[s], ph);
      ptx::tc_fence_after();
      uint64_t adesc = kDescMN + (base16 + s * stage16);
      uint64_t bdesc = adesc + b_off16;
      for (int j = 0; j < mmas; ++j) {
        if (leader) ptx::mma_ss<kTf32>(tmem_base, adesc, bdesc, idesc, (k | j) ? 1u : 0u);
        adesc += step;
        bdesc += step;
      }
      if (leader) ptx::tc_commit(&empty_bar[s]);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (leader) ptx::tc_commit(tmem_full_bar);
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int b = m / P.kelems, cl = m - b * P.kelems;
    bool row_ok = b < nblk;
    int64_t row = 0;
    bool is_db = false;
    if (row_ok) {
      const ABlock ab = P.blocks[blk0 + b];
      const int ci = ab.c0 + cl;
      is_db = ab.tap < 0;
      row_ok = is_db ? (cl == 0) : (ci < P.CI);
      row = (int64_t)ab.tap * P.CI + ci;
    }
    float* out = is_db ? part_db + (int64_t)blockIdx.z * P.CO : part + ((int64_t)blockIdx.z * P.Mrows + row) * P.CO;
    if (nk > 0) {
      ptx::mbar_wait_sleep(tmem_full_bar, 0, (uint32_t)P.wait_ns);
      ptx::tc_fence_after();
    }
    const int BN = P.nb * P.kelems;
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      if (nk > 0) {
        ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int co = co0 + c;
      if (row_ok && co < P.CO) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (co + 4 * j < P.CO)
            *reinterpret_cast<uint4*>(out + co + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
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

void choose_patch_w(int ow, int oh, int n, int& TW, int& TH, int& TN) {
  long best = -1;
  for (int tw = 1; tw <= 128 && tw <= ow; ++tw)
    for (int th = 1; th * tw <= 128 && th <= oh; ++th) {
      int tn = 128 / (tw * th);
      if (tn > n) tn = n;
      // K per instruction is 8 or 16 pixels: keep the used rows a multiple of 16 where possible is not
      // required (unused rows are zero), so just minimise tiles
      const long tiles = (long)ceildiv(ow, tw) * ceildiv(oh, th) * ceildiv(n, tn);
      const long score = tiles * 1024 - tw;
      if (best < 0 || score < best) { best = score; TW = tw; TH = th; TN = tn; }
    }
}

struct WPlan {
  TcWgradParams P;
  dim3 grid;
  size_t smem = 0, part_bytes = 0, partdb_bytes = 0;
  int splits = 1;
};

bool plan_wgrad(const GatherWgrad& g, int mode, WPlan& pl, bool want_db = false) {
  const bool bf16 = mode == NPML_MODE_BF16;
  if (mode != NPML_MODE_BF16 && mode != NPML_MODE_TF32) return false;
  const int es = bf16 ? 2 : 4, vec = 16 / es;
  if (g.CI % vec || g.CO % vec || g.CO % 4) return false;
  if (g.s < 1 || g.s > 3 || g.N < 1) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  TcWgradParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.kelems = kRowBytes / es;
  const int cblocks = ceildiv(g.CI, P.kelems);
  if (g.fr * g.fc * cblocks + 1 > kMaxBlocks) return false;
  P.CI = g.CI; P.CO = g.CO; P.Mrows = g.fr * g.fc * g.CI;
  P.N = g.N; P.OH = g.OH; P.OW = g.OW;
  P.mb = 128 / P.kelems;                                   // 2 (bf16) or 4 (tf32) windows per M tile
  const int max_bn = bf16 ? 128 : 64;
  int bn = ceildiv(g.CO, P.kelems) * P.kelems;
  if (bn > max_bn) bn = max_bn;
  P.nb = bn / P.kelems;
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  P.tmem_cols = cols;
  P.idesc = ptx::instr_desc(bf16 ? 1 : 2, 1, 1, 128, bn);
  P.wait_ns = wait_ns();
  P.kinstr_bytes = (bf16 ? 16 : 8) * kRowBytes;             // K pixels per MMA x 128 B
  P.mmas_per_tile = 128 / (bf16 ? 16 : 8);
  int nb = 0;
  for (int r = 0; r < g.fr; ++r)
    for (int t = 0; t < g.fc; ++t) {
      const int offy = r * g.D - g.pr, offx = t * g.D - g.pc;
      const int py = posmod(offy, g.s), px = posmod(offx, g.s);
      for (int cb = 0; cb < cblocks; ++cb) {
        ABlock& b = P.blocks[nb++];
        b.dy = (int16_t)floordiv(offy, g.s);
        b.dx = (int16_t)floordiv(offx, g.s);
        // an empty parity plane means the tap only ever sees padding: point it far outside any map
        if (py >= g.IH || px >= g.IW) { b.amap = 0; b.dy = 30000; b.dx = 30000; }
        else b.amap = (int16_t)(py * g.s + px);
        b.tap = (int16_t)(r * g.fc + t);
        b.c0 = cb * P.kelems;
      }
    }
  if (want_db) {                                           // the constant ones window that yields db
    ABlock& b = P.blocks[nb++];
    b.dy = 0; b.dx = 0; b.amap = 0; b.tap = -1; b.c0 = 0;
  }
  P.nblocks = nb;
  choose_patch_w(g.OW, g.OH, g.N, P.TW, P.TH, P.TN);
  P.rows_used = P.TW * P.TH * P.TN;
  P.tiles_x = ceildiv(g.OW, P.TW); P.tiles_y = ceildiv(g.OH, P.TH); P.tiles_n = ceildiv(g.N, P.TN);
  P.ntiles = P.tiles_x * P.tiles_y * P.tiles_n;
  const int stage_bytes = (P.mb + P.nb) * kTileBytes;
  P.stages = (200 * 1024) / stage_bytes;
  if (P.stages > 4) P.stages = 4;
  if (P.stages < 2) return false;
  pl.smem = (size_t)P.stages * stage_bytes + 1024 + 256;
  const int mt = ceildiv(P.nblocks, P.mb), nt = ceildiv(g.CO, bn);
  // Split the pixel range so that the CTAs fill WHOLE waves: the kernel takes waves x (pixel tiles per CTA + a fixed
  // prologue / epilogue), and one CTA more than a multiple of the resident slots costs a full extra wave (cfg5 used
  // to launch 320 CTAs on 148 slots: three waves for 2.16 waves of work).  Every split also adds one partial to the
  // fixed-order reduction.  Costs are in units of one pixel tile (8 or 16 MMAs).
  const int per_sm = (int)((227 * 1024) / pl.smem) < 1 ? 1 : (int)((227 * 1024) / pl.smem);
  const int slots = per_sm * num_sms();
  const int max_splits = P.ntiles < 4 ? 1 : ceildiv(P.ntiles, 4);   // at least 4 pixel tiles per CTA
  int splits = 1;
  double best = -1.0;
  for (int sp = 1; sp <= max_splits && sp <= 64; ++sp) {
    const int tps = ceildiv(P.ntiles, sp);
    const int eff = ceildiv(P.ntiles, tps);
    if (eff != sp) continue;                                  // same launch as a smaller sp
    const int waves = ceildiv(mt * nt * eff, slots);
    const double cost = (double)waves * (tps + 2.0) + 1.2 * eff;
    if (best < 0.0 || cost < best) { best = cost; splits = sp; }
  }
  P.tiles_per_split = ceildiv(P.ntiles, splits);
  splits = ceildiv(P.ntiles, P.tiles_per_split);
  pl.splits = splits;
  pl.grid = dim3((unsigned)mt, (unsigned)nt, (unsigned)splits);
  pl.part_bytes = align_up((size_t)splits * P.Mrows * g.CO * sizeof(float), 256);
  pl.partdb_bytes = align_up((size_t)splits * g.CO * sizeof(float), 256);
  return true;
}

template <typename T>
int launch_wgrad(const GatherWgrad& g, WPlan& pl, const void* in, const void* grad, float* dW, float* db, void* ws,
                 cudaStream_t st) {
  const bool bf16 = sizeof(T) == 2;
  const int es = sizeof(T);
  TcWgradParams& P = pl.P;
  const uint32_t box[4] = {(uint32_t)P.kelems, (uint32_t)P.TW, (uint32_t)P.TH, (uint32_t)P.TN};
  {
    const uint64_t dims[4] = {(uint64_t)g.CO, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
    const uint64_t strides[4] = {(uint64_t)es, (uint64_t)g.CO * es, (uint64_t)g.OW * g.CO * es,
                                 (uint64_t)g.OH * g.OW * g.CO * es};
    if (int e = make_tmap(&P.gmap, bf16, const_cast<void*>(grad), 4, dims, strides, box, !bf16)) return e;
  }
  for (int py = 0; py < g.s; ++py)
    for (int px = 0; px < g.s; ++px) {
      if (py >= g.IH || px >= g.IW) continue;
      const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)ceildiv(g.IW - px, g.s), (uint64_t)ceildiv(g.IH - py, g.s),
                                (uint64_t)g.N};
      const uint64_t strides[4] = {(uint64_t)es, (uint64_t)g.s * g.CI * es, (uint64_t)g.s * g.IW * g.CI * es,
                                   (uint64_t)g.IH * g.IW * g.CI * es};
      char* base = (char*)in + ((size_t)py * g.IW + px) * g.CI * es;
      if (int e = make_tmap(&P.amap[py * g.s + px], bf16, base, 4, dims, strides, box, !bf16)) return e;
    }
  auto kern = tc_wgrad_kernel<T>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[bf16 ? 0 : 1]) {
    NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[bf16 ? 0 : 1] = true;
  }
  float* part_db = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.part_bytes);
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, (float*)ws, part_db);
  NPML_LAUNCH_OK();
  return launch_wgrad_reduce((const float*)ws, pl.splits, P.Mrows, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate,
                             dW, st, db != nullptr ? part_db : nullptr, db);
}

}  // namespace

bool tc_wgrad_supported(const GatherWgrad& g, int mode) {
  WPlan pl;
  return plan_wgrad(g, mode, pl);
}

size_t tc_wgrad_ws_bytes(const GatherWgrad& g, int mode) {
  WPlan pl;
  // the split count depends on the number of M tiles, which the optional db window can change: cover both plans
  WPlan pl2;
  if (!plan_wgrad(g, mode, pl, true) || !plan_wgrad(g, mode, pl2, false)) return 0;
  const size_t a = pl.part_bytes + pl.partdb_bytes, b = pl2.part_bytes + pl2.partdb_bytes;
  return (a > b ? a : b) + 256;
}

// db != nullptr: db (+)= column sums of `grad` from the same pass (one extra window of ones); only meaningful when
// `grad` is dZ (Conv2D).  The fixed-order reduction of the partials writes dW and db.
int tc_wgrad(const GatherWgrad& g, int mode, const void* in, const void* grad, float* dW, void* ws, size_t ws_bytes,
             cudaStream_t st, float* db) {
  WPlan pl;
  NPML_CHECK(plan_wgrad(g, mode, pl, db != nullptr), "shape not supported by the tensor-core path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.part_bytes + pl.partdb_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0,
             "activation pointers must be 16-byte aligned");
  if (mode == NPML_MODE_BF16) return launch_wgrad<__nv_bfloat16>(g, pl, in, grad, dW, db, ws, st);
  return launch_wgrad<float>(g, pl, in, grad, dW, db, ws, st);
}

}  // namespace npml
