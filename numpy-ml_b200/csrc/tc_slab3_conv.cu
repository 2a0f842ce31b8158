// "Row-of-taps" slab convolution for narrow outputs (fc * out_ch <= 256, e.g. cfg2: 3x3, 64 -> 64) on sm_100a.
//
// tc_slab_conv.cu issues one M=128 x N=out_ch MMA per filter tap.  With out_ch = 64 such an MMA is dominated by its
// operand fetch (4 KB of A + 2 KB of B from shared memory for 32 cycles of math; DESIGN.md finding 2).  Here the fc
// taps of one filter ROW share a single MMA: the B operand is the fc weight blocks of that row stacked along N
// (N = fc * out_ch = 192), the A operand is the slab window of the row's FIRST tap only.  Column block j of the
// accumulator row of slot q then holds
//
//     D_j[q] = sum_{row taps g, ci} X[q + offy_g * Wp] * W[g][j]            (no column shift inside the MMA)
//
// which is the part of output slot q - j*D that comes from column tap j, so the epilogue finishes the sum with
//
//     out[s] = D_0[s] + D_1[s + D] + ... + D_{fc-1}[s + (fc-1) D]
//
// (a shift along accumulator rows = TMEM lanes: warp shuffles inside a 32-lane quadrant, a small shared-memory
// exchange across quadrants).  A tile of 128 accumulator rows therefore yields VR = 128 - (fc-1) D finished
// output slots, and consecutive tiles start VR slots apart.  Per tile of cfg2 this is 12 MMAs of 128 x 192 x 16
// instead of 36 of 128 x 64 x 16: the A slab is fetched 3x less often and the math per fetch triples.
//
// One persistent CTA per SM, one tile per work unit, two TMEM accumulators (ping-pong with the epilogue):
//   warp 0  slab producer (TMA, whole padded rows, 2-stage ring of (tile, channel-block) planes)
//   warp 1  MMA issuer
//   warp 2  weight producer (all taps resident in smem) and TMEM allocator
//   warps 4..11 epilogue, two per TMEM lane quadrant (each pair splits the accumulator columns)
#include <cstdlib>

#include "npml_common.cuh"
#include "tc_ptx.cuh"

namespace npml {

int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);
int launch_repack_weights(const GatherConv& g, bool bf16, const float* w, void* wt, cudaStream_t st);

namespace {

constexpr int kMaxRows = 8;        // filter rows
constexpr int kMaxFC = 4;          // filter columns folded into N
constexpr int kRowBytes = 128;
constexpr int kThreads = 384;      // warps 0..3 producers / issuer, warps 4..11 epilogue (two per TMEM lane quadrant)
constexpr int kSlabStages = 2;
constexpr int kAccCols = 256;      // TMEM columns per accumulator (N <= 256), two accumulators
constexpr int kMaxShift = 8;       // (fc-1)*D rows exchanged between neighbouring epilogue warps
constexpr int kSmemLimit = 227 * 1024;

struct Slab3Params {
  CUtensorMap amap, wmap;
  uint32_t row_off16[kMaxRows];            // offy * Wp slots * 128 B >> 4: descriptor advance of a filter row
  int16_t tap_w[kMaxRows * kMaxFC];        // Wt index of (filter row g, accumulator column block j)
  int nrows, fc, nkb, kelems;
  int N, OH, OW, CO, Hp, Wp, pady, padx;
  int D, VR, SR, BN, NB;
  int plane_bytes, w_bytes;
  int xrows;                               // (fc-1)*D accumulator rows a warp hands to its lower neighbour
  int total_tiles;
  uint32_t idesc, idesc1;                  // N = FOLD*BN (folded taps) and N = BN (the remaining column taps)
  int act;
  float p0, p1;
};

constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);

// named barrier of the four epilogue warps that work on the same accumulator columns (ids 1 and 2)
__device__ __forceinline__ void epi_bar_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// FC column taps per filter row; the first FOLD of them are folded into N (one MMA, epilogue row shifts), the others
// are separate N = BN MMAs whose A window is advanced by t*D slots and which add straight into column block 0.
template <typename T, int FR, int FC, int FOLD>
__global__ void __launch_bounds__(kThreads, 1) tc_slab3_conv_kernel(const __grid_constant__ Slab3Params P,
                                                                    const float* __restrict__ bias, T* __restrict__ Z,
                                                                    T* __restrict__ Y) {
  constexpr bool kTf32 = sizeof(T) == 4;
  // [weights (1024-aligned)] [2 slab planes] [epilogue staging] [row exchange] [bias] [barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* w_area = smem_raw;
  uint8_t* smem = w_area + P.w_bytes;
  uint8_t* staging = smem + (size_t)kSlabStages * P.plane_bytes;                 // 8 warps x 32 rows x 32 columns of T
  float* xch = reinterpret_cast<float*>(staging + 8 * 32 * 32 * sizeof(T));      // [2 groups][2][FOLD-1][4 quadrants][xrows][32]
  float* bias_s = xch + 2 * 2 * (FOLD - 1) * 4 * P.xrows * 32;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(bias_s + 128);
  uint64_t* slab_empty = slab_full + kSlabStages;
  uint64_t* w_full = slab_empty + kSlabStages;
  uint64_t* acc_full = w_full + 1;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nrows = FR > 0 ? FR : P.nrows;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kSlabStages; ++i) { ptx::mbar_init(&slab_full[i], 1); ptx::mbar_init(&slab_empty[i], 1); }
    ptx::mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (warp >= 4 && warp < 8) {
    const int i = threadIdx.x - 128;
    bias_s[i] = (bias != nullptr && i < P.BN && i < P.CO) ? bias[i] : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== slab producer: padded rows [rho0, rho0 + SR) of channel block kb, one TMA box per row =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.amap);
      const uint32_t row_bytes = (uint32_t)P.Wp * kRowBytes;
      uint32_t p = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int rho0 = (tile * P.VR) / P.Wp;
        for (int kb = 0; kb < P.nkb; ++kb, ++p) {
          const int st = p & 1;
          ptx::mbar_wait(&slab_empty[st], ((p >> 1) & 1u) ^ 1u);
          ptx::mbar_expect_tx(&slab_full[st], (uint32_t)P.SR * row_bytes);
          uint8_t* dst = smem + (size_t)st * P.plane_bytes;
          int n = rho0 / P.Hp, yp = rho0 - n * P.Hp;
          for (int i = 0; i < P.SR; ++i) {
            ptx::tma_load_4d(dst + (size_t)i * row_bytes, &P.amap, &slab_full[st], kb * P.kelems, -P.padx, yp - P.pady, n);
            if (++yp == P.Hp) { yp = 0; ++n; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer: every (kb, filter row) block of fc * BN rows, resident for the whole kernel =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.wmap);
      ptx::mbar_expect_tx(w_full, (uint32_t)P.w_bytes);
      const int blk = P.BN * kRowBytes;
      for (int kb = 0; kb < P.nkb; ++kb)
        for (int g = 0; g < nrows; ++g)
          for (int j = 0; j < FC; ++j)
            ptx::tma_load_3d(w_area + (size_t)((kb * nrows + g) * FC + j) * blk, &P.wmap, w_full, kb * P.kelems, 0,
                             P.tap_w[g * FC + j]);
    }
  } else if (warp == 1) {
    // ===== MMA issuer (convergent control flow, descriptors on the uniform datapath) =====
    const bool leader = ptx::elect_one();
    const uint32_t idesc = P.idesc, idesc1 = P.idesc1;
    const uint32_t tapstep16 = (uint32_t)P.D * (kRowBytes >> 4), bn16 = (uint32_t)(P.BN * kRowBytes) >> 4;
    const int nkb = P.nkb, Wp = P.Wp, VR = P.VR, total_tiles = P.total_tiles;
    const uint32_t plane16 = (uint32_t)P.plane_bytes >> 4;
    const uint32_t wblk16 = (uint32_t)(P.NB * kRowBytes) >> 4;
    const uint64_t slab_d = kDesc64 + (ptx::smem_u32(smem) >> 4);
    const uint64_t w_d = kDesc64 + (ptx::smem_u32(w_area) >> 4);
    uint32_t p = 0, it = 0;
    ptx::mbar_wait(w_full, 0);
    ptx::tc_fence_after();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      const uint32_t dt = tmem_base + acc * kAccCols;
      const uint32_t a_off16 = (uint32_t)((tile * VR) % Wp) * (kRowBytes >> 4);
      ptx::mbar_wait(&acc_empty[acc], aph ^ 1u);
      ptx::tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++p) {
        const uint32_t st = p & 1u;
        ptx::mbar_wait(&slab_full[st], (p >> 1) & 1u);
        ptx::tc_fence_after();
        const uint64_t a_base = slab_d + (st * plane16 + a_off16);
#pragma unroll
        for (int g = 0; g < nrows; ++g) {
          const uint64_t ad = a_base + P.row_off16[g];
          const uint64_t bd = w_d + (uint32_t)(kb * nrows + g) * wblk16;
          if (leader) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::mma_ss<kTf32>(dt, ad + (uint64_t)(2 * ks), bd + (uint64_t)(2 * ks), idesc, (kb | g | ks) ? 1u : 0u);
#pragma unroll
            for (int t = FOLD; t < FC; ++t) {     // unfolded column taps: shifted A window, N = BN, into column block 0
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::mma_ss<kTf32>(dt, ad + (uint64_t)(t * tapstep16 + 2 * ks), bd + (uint64_t)(t * bn16 + 2 * ks), idesc1, 1u);
            }
          }
        }
        if (leader) ptx::tc_commit(&slab_empty[st]);
      }
      if (leader) ptx::tc_commit(&acc_full[acc]);
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    constexpr int kRB = 32 * (int)sizeof(T);
    constexpr int kCPR = kRB / 16;
    constexpr int kRPI = 32 / kCPR;
    constexpr int kVec = 16 / (int)sizeof(T);
    const int qd = warp & 3;
    const int m = qd * 32 + lane;
    const int img_slots = P.Hp * P.Wp;
    const int BN = P.BN, CO = P.CO, Wp = P.Wp, D = P.D, VR = P.VR;
    const int64_t slots = (int64_t)P.N * img_slots;
    uint8_t* stg = staging + (size_t)(warp - 4) * 32 * kRB;
    const uint32_t stg_u32 = ptx::smem_u32(stg);
    const int my_swz = kCPR == 4 ? ((lane >> 1) & 3) : (lane & 7);
    const int rd_row0 = lane / kCPR, rd_cc = lane % kCPR;

    auto emit = [&](T* __restrict__ out, const float (&v)[32], int my_pix, int co) {
#pragma unroll
      for (int cc = 0; cc < kCPR; ++cc) {
        uint4 u;
        if constexpr (sizeof(T) == 2) {
          __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * cc], v[8 * cc + 1]);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * cc + 2], v[8 * cc + 3]);
          __nv_bfloat162 c2 = __floats2bfloat162_rn(v[8 * cc + 4], v[8 * cc + 5]);
          __nv_bfloat162 d2 = __floats2bfloat162_rn(v[8 * cc + 6], v[8 * cc + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b2);
          u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d2);
        } else {
          u.x = __float_as_uint(v[4 * cc]); u.y = __float_as_uint(v[4 * cc + 1]);
          u.z = __float_as_uint(v[4 * cc + 2]); u.w = __float_as_uint(v[4 * cc + 3]);
        }
        const uint32_t addr = stg_u32 + (uint32_t)(lane * kRB + ((cc ^ my_swz) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int it2 = 0; it2 < kCPR; ++it2) {
        const int r = it2 * kRPI + rd_row0;
        const int pr = __shfl_sync(0xffffffffu, my_pix, r);
        const int swz = kCPR == 4 ? ((r >> 1) & 3) : (r & 7);
        uint4 u;
        const uint32_t addr = stg_u32 + (uint32_t)(r * kRB + ((rd_cc ^ swz) << 4));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
        const int c_el = co + rd_cc * kVec;
        if (pr >= 0 && c_el < CO) *reinterpret_cast<uint4*>(out + (int64_t)pr * CO + c_el) = u;
      }
      __syncwarp();
    };

    // two warps per TMEM lane quadrant: group 0 finishes accumulator columns [0, 32), [64, 96), ..., group 1 the
    // columns [32, 64), [96, 128), ...  Each group has its own named barrier and exchange buffers.
    const int grp = (warp - 4) >> 2;
    float* xg = xch + (size_t)grp * 2 * (FOLD - 1) * 4 * P.xrows * 32;
    const int xstride = (FOLD - 1) * 4 * P.xrows * 32;      // floats per exchange buffer
    uint32_t it = 0, xb = 0;     // tile counter (accumulator ping-pong), exchange buffer parity
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      const int64_t s = (int64_t)tile * VR + m;
      const int n = (int)(s / img_slots);
      const int rem = (int)(s - (int64_t)n * img_slots);
      const int yp = rem / Wp, xp = rem - yp * Wp;
      const bool row_ok = m < VR && s < slots && yp < P.OH && xp < P.OW;
      const int my_pix = row_ok ? (n * P.OH + yp) * P.OW + xp : -1;
      ptx::mbar_wait(&acc_full[acc], aph);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + acc * kAccCols;
      bool released = false;
      for (int c = grp * 32; c < BN; c += 64, xb ^= 1u) {
        uint32_t r0[32], rs[FOLD > 1 ? FOLD - 1 : 1][32];
        ptx::tmem_ld32(t_addr + (uint32_t)c, r0);
#pragma unroll
        for (int j = 1; j < FOLD; ++j) ptx::tmem_ld32(t_addr + (uint32_t)(j * BN + c), rs[j - 1]);
        ptx::tmem_ld_wait();
        if (c + 64 >= BN) {                      // this warp's last read of the accumulator
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
          released = true;
        }
        // rows the neighbouring quadrant below needs from column block j: my first j*D lanes
        float* xw = xg + (size_t)xb * xstride;
#pragma unroll
        for (int j = 1; j < FOLD; ++j) {
          if (lane < j * D) {
            float4* dst = reinterpret_cast<float4*>(xw + (((j - 1) * 4 + qd) * P.xrows + lane) * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              dst[i] = make_float4(__uint_as_float(rs[j - 1][4 * i]), __uint_as_float(rs[j - 1][4 * i + 1]),
                                   __uint_as_float(rs[j - 1][4 * i + 2]), __uint_as_float(rs[j - 1][4 * i + 3]));
          }
        }
        epi_bar_sync(grp);
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_s + c + i);
          v[i] = __uint_as_float(r0[i]) + bb.x; v[i + 1] = __uint_as_float(r0[i + 1]) + bb.y;
          v[i + 2] = __uint_as_float(r0[i + 2]) + bb.z; v[i + 3] = __uint_as_float(r0[i + 3]) + bb.w;
        }
#pragma unroll
        for (int j = 1; j < FOLD; ++j) {
          const int sh = j * D;
          float a[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = __shfl_down_sync(0xffffffffu, __uint_as_float(rs[j - 1][i]), sh);
          if (lane + sh >= 32) {                 // the row lives in the next quadrant (or beyond the tile: masked row)
            const float4* src = reinterpret_cast<const float4*>(
                xw + (((j - 1) * 4 + ((qd + 1) & 3)) * P.xrows + ((lane + sh) & 31)) * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 x = src[i];
              a[4 * i] = x.x; a[4 * i + 1] = x.y; a[4 * i + 2] = x.z; a[4 * i + 3] = x.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += a[i];
        }
        if (Z) emit(Z, v, my_pix, c);
        if (Y) {
          act_apply32(P.act, P.p0, P.p1, v);
          emit(Y, v, my_pix, c);
        }
      }
      if (!released) {                           // a group without columns of its own still hands the accumulator back
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

inline int ceildiv(int a, int b) { return (a + b - 1) / b; }

struct Slab3Plan {
  Slab3Params P;
  int fold = 3;
  dim3 grid;
  size_t smem = 0, wt_bytes = 0;
};

bool plan_slab3(const GatherConv& g, int mode, Slab3Plan& pl) {
  if (mode != NPML_MODE_BF16 && mode != NPML_MODE_TF32) return false;
  // Opt-in experiment (NPML_SLAB3=2 or 3 = number of column taps folded into one MMA).  Measured on cfg2 (B200, 1 kW
  // cap): 3 folded taps are epilogue-bound (88 us vs 76 us for tc_slab_conv.cu), 2 folded taps need 6 % fewer SM
  // cycles but the board then clocks lower under its power cap (74 us, and the wgrad that follows slows by the same
  // amount), so tc_slab_conv.cu stays the default.  See DESIGN.md section 4, finding 5.
  const char* opt = std::getenv("NPML_SLAB3");
  if (opt == nullptr || (opt[0] != '2' && opt[0] != '3')) return false;
  if (g.post_scale != nullptr) return false;        // the folded BatchNorm2D epilogue lives in tc_slab_conv.cu
  const bool bf16 = mode == NPML_MODE_BF16;
  const int es = bf16 ? 2 : 4, vec = 16 / es;
  if (g.s != 1 || g.N < 1) return false;
  if (g.CI % vec || g.CO % vec) return false;
  if (g.fc != 3 || g.fr > kMaxRows) return false;                     // the instantiated row width
  const int bn = (g.CO + 15) / 16 * 16;
  const int fold = opt[0] - '0';
  pl.fold = fold;
  if (bn < 32 || bn * fold > 256) return false;                       // narrow outputs only: that is where it pays
  if ((fold - 1) * g.D > kMaxShift) return false;
  const int Hp = g.OH + (g.fr - 1) * g.D, Wp = g.OW + (g.fc - 1) * g.D;
  if (Wp > 256) return false;
  const int64_t slots = (int64_t)g.N * Hp * Wp;
  if (slots + 128 * 8 >= (1LL << 31) / 2) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  Slab3Params& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.N = g.N; P.OH = g.OH; P.OW = g.OW; P.CO = g.CO; P.Hp = Hp; P.Wp = Wp;
  P.pady = g.form == 0 ? g.pr : (g.fr - 1) * g.D - g.pr;
  P.padx = g.form == 0 ? g.pc : (g.fc - 1) * g.D - g.pc;
  P.nrows = g.fr; P.fc = g.fc; P.D = g.D;
  // accumulator column block j collects the taps whose column offset is j*D; filter row group gi has row offset gi*D
  for (int gi = 0; gi < g.fr; ++gi) {
    P.row_off16[gi] = (uint32_t)(gi * g.D * Wp) * (kRowBytes >> 4);
    for (int j = 0; j < g.fc; ++j) {
      const int r = g.form == 0 ? gi : g.fr - 1 - gi;
      const int t = g.form == 0 ? j : g.fc - 1 - j;
      P.tap_w[gi * g.fc + j] = (int16_t)(r * g.fc + t);
    }
  }
  P.kelems = kRowBytes / es;
  P.nkb = ceildiv(g.CI, P.kelems);
  P.act = g.act; P.p0 = g.p0; P.p1 = g.p1;
  P.BN = bn;
  P.NB = bn * g.fc;                      // weight rows per (kb, filter row) block in smem: all fc taps, contiguous
  P.VR = 128 - (fold - 1) * g.D;
  P.xrows = (fold - 1) * g.D;
  P.idesc = ptx::instr_desc(bf16 ? 1 : 2, 0, 0, 128, bn * fold);
  P.idesc1 = ptx::instr_desc(bf16 ? 1 : 2, 0, 0, 128, bn);
  P.w_bytes = g.fr * P.nkb * P.NB * kRowBytes;
  const int halo = (g.fr - 1) * g.D * Wp + (g.fc - fold) * g.D + (fold < g.fc ? fold - 1 : 0) * g.D;
  P.SR = (Wp - 1 + 128 + halo - 1) / Wp + 1;
  P.plane_bytes = P.SR * Wp * kRowBytes;
  P.total_tiles = (int)((slots + P.VR - 1) / P.VR);
  const size_t fixed = 8 * 32 * 32 * es + 2 * 2 * (fold - 1) * 4 * P.xrows * 32 * sizeof(float) + 512 + 256;
  pl.smem = (size_t)P.w_bytes + (size_t)kSlabStages * P.plane_bytes + fixed;
  if (pl.smem > (size_t)kSmemLimit) return false;
  int ctas = num_sms();
  if (ctas > P.total_tiles) ctas = P.total_tiles;
  pl.grid = dim3((unsigned)ctas, 1, 1);
  pl.wt_bytes = align_up((size_t)g.fr * g.fc * g.CO * g.CI * es, 256);
  return true;
}

template <typename T>
int launch_slab3(const GatherConv& g, Slab3Plan& pl, const void* in, const float* w, const float* bias, void* Z, void* Y,
                 void* ws, cudaStream_t st) {
  const bool bf16 = sizeof(T) == 2;
  const int es = sizeof(T);
  Slab3Params& P = pl.P;
  if (int e = launch_repack_weights(g, bf16, w, ws, st)) return e;
  {
    const uint64_t dims[3] = {(uint64_t)g.CI, (uint64_t)g.CO, (uint64_t)(g.fr * g.fc)};
    const uint64_t strides[3] = {(uint64_t)es, (uint64_t)g.CI * es, (uint64_t)g.CI * g.CO * es};
    const uint32_t box[3] = {(uint32_t)P.kelems, (uint32_t)P.BN, 1};
    if (int e = make_tmap(&P.wmap, bf16, ws, 3, dims, strides, box, false)) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[4] = {(uint64_t)es, (uint64_t)g.CI * es, (uint64_t)g.IW * g.CI * es,
                                 (uint64_t)g.IH * g.IW * g.CI * es};
    const uint32_t box[4] = {(uint32_t)P.kelems, (uint32_t)P.Wp, 1, 1};
    if (int e = make_tmap(&P.amap, bf16, const_cast<void*>(in), 4, dims, strides, box, false)) return e;
  }
  typedef void (*Kern)(const Slab3Params, const float*, T*, T*);
  Kern kern;
  if (pl.fold == 3) kern = P.nrows == 3 ? (Kern)tc_slab3_conv_kernel<T, 3, 3, 3> : (Kern)tc_slab3_conv_kernel<T, 0, 3, 3>;
  else kern = P.nrows == 3 ? (Kern)tc_slab3_conv_kernel<T, 3, 3, 2> : (Kern)tc_slab3_conv_kernel<T, 0, 3, 2>;
  NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, bias, (T*)Z, (T*)Y);
  NPML_LAUNCH_OK();
  return 0;
}

}  // namespace

bool slab3_conv_supported(const GatherConv& g, int mode) {
  Slab3Plan pl;
  return plan_slab3(g, mode, pl);
}

int slab3_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z, void* Y,
                      void* ws, size_t ws_bytes, cudaStream_t st) {
  Slab3Plan pl;
  NPML_CHECK(plan_slab3(g, mode, pl), "shape not supported by the row-of-taps slab path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.wt_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 127) == 0,
             "activation / workspace pointers must be 16 / 128 byte aligned");
  if (mode == NPML_MODE_BF16) return launch_slab3<__nv_bfloat16>(g, pl, in, w, bias, Z, Y, ws, st);
  return launch_slab3<float>(g, pl, in, w, bias, Z, Y, ws, st);
}

}  // namespace npml
