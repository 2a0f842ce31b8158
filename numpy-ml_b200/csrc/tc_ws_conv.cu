// Weight-stationary ("transposed") slab convolution for stride-1 gather convolutions with few output channels
// (out_ch <= 128, all filter taps small enough to live in tensor memory), bf16, sm_100a.
//
// Why.  scripts/mma_rate.cu (profiles/r1g_mma_rate.txt) measured on B200 that a tcgen05.mma whose two operands come
// from shared memory is paced by the shared-memory port (128 B/clk): 128x64x16 needs 4 KB of A + 2 KB of B = 48 clk for
// 32 clk of math, so a 64-channel convolution (cfg2) cannot pass 2/3 of the tensor peak in the pixel-major form of
// tc_slab_conv.cu, while the same MMA with A in TENSOR MEMORY runs at the math floor.  Here the roles are swapped:
//
//   D[co, pixel] += W_tap[co, ci] * slab[pixel + off(tap), ci]
//
//   * A = the weights, resident in TMEM for the whole kernel (written once with tcgen05.st; layout measured by
//     scripts/ts_probe.cu: lane = row, 32-bit column j = K elements 2j, 2j+1),
//   * B = NT consecutive slots of the same shared-memory slab tc_slab_conv.cu uses as its A operand (whole padded rows
//     by TMA, zero padding = out-of-bounds fill, tap = descriptor start advanced by the tap offset),
//   * the accumulator is [co (TMEM lanes)] x [pixels (columns)]: an epilogue thread owns one channel and writes
//     2-byte elements, the 32 lanes of a warp covering 64 contiguous bytes of a pixel.
//
// out_ch <= 64 would leave half of the M = 128 datapath idle, so two taps (a, b) of a filter row, off(b) = off(a) + delta
// (delta = D or 2D slots, whichever is even), share one MMA.  Both halves multiply the SAME window, so the b half accumulates tap b's contribution to the output D
// slots further left: the MMA is 8 columns wider than the tile and the epilogue adds b[:, q + delta] to a[:, q].  To make that
// a plain register add, the two halves of a channel sit in the SAME 32-lane TMEM quadrant (rows 32w..32w+15 = W_a of
// channels 16w..16w+15, rows 32w+16..32w+31 = W_b of the same channels) and the epilogue reads each 16-lane half with
// tcgen05.ld.16x256b at its own column offset: both loads deliver the same (row, column) fragment to every thread
// (mapping measured by scripts/ts_probe2.cu).  A 3x3 filter needs 6 such MMAs per K block instead of 9.
//
// The epilogue then transposes through a swizzled shared-memory tile so that global stores are 16 bytes per lane and
// 128 contiguous bytes per pixel.
//
//   warp 0      slab producer (TMA, 2-stage ring)                 warp 1  MMA issuer
//   warp 2      TMEM allocator                                    warp 3  idle
//   warps 4..7  epilogue group 0 (and the one-time weight load)   warps 8..11 epilogue group 1 (tiles alternate)
#include <cstdlib>
#include <cstring>

#include "npml_common.cuh"
#include "tc_ptx.cuh"

namespace npml {

int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);
int launch_repack_weights(const GatherConv& g, bool bf16, const float* w, void* wt, cudaStream_t st);

namespace {

constexpr int kMaxGroups = 40;
constexpr int kRowBytes = 128;
constexpr int kThreads = 384;
constexpr int kSlabStages = 2;
constexpr int kMaxAcc = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);

struct WsParams {
  CUtensorMap amap;
  uint32_t tap_off16[kMaxGroups];                 // B-descriptor advance of the group's (first) tap
  int16_t tap_a[kMaxGroups], tap_b[kMaxGroups];   // filter tap (index into Wt) of lanes 0..63 / 64..127; -1 = zeros
  int ngroups, nkb, CI, CO;
  int N, OH, OW, Hp, Wp, pady, padx;
  int pair, delta;           // pair mode (CO <= 64): the bottom half's columns are shifted by delta slots
  int NT, NC;                // output slots per tile, MMA N (= accumulator columns) >= NT + delta
  int UT, SR, n_acc, wcols;  // tiles per unit, slab rows per unit, accumulator ring, TMEM columns of the weights
  int plane_bytes, stage_bytes, total_tiles, num_units;
  uint32_t idesc;
  int act, affine;
  float p0, p1;
  int dbg;                   // NPML_WS_DBG (scripts/ws_probe.py: timing experiments only, results are wrong): 1 = epilogue without shared / global traffic, 2 = no TMA refills
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem], bf16 operands
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 16 lanes x 32 consecutive fp32 columns: register 4j + k of thread t holds row (t / 4 + 8 * (k / 2)) of the 16 lanes at
// the lane base, column 8 * j + 2 * (t % 4) + (k % 2)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Four 8x8 b16 matrices, each stored TRANSPOSED: the fragment a thread holds for matrix m (register m: row t / 4,
// columns 2 (t % 4), 2 (t % 4) + 1 -- exactly the packed tcgen05.ld.16x256b accumulator fragment, channels along rows,
// pixels along columns) lands as 8 rows of [pixel][8 channels] (16 bytes each) at the addresses given by threads
// 8 m .. 8 m + 7.  One instruction replaces 32 two-byte st.shared per thread-quad: the [co][pixel] -> NHWC transpose
// of the weight-stationary epilogue costs nothing.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) tc_ws_conv_kernel(const __grid_constant__ WsParams P,
                                                                 const __nv_bfloat16* __restrict__ wt,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ post_scale,
                                                                 const float* __restrict__ post_shift,
                                                                 __nv_bfloat16* __restrict__ Z, __nv_bfloat16* __restrict__ Y) {
  // [2 slab planes] [2 x (epilogue staging tile: NT pixels x (128 | 256) bytes, pixel table: NT ints)] [barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  uint8_t* stage = smem + (size_t)kSlabStages * P.plane_bytes;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(stage + 2 * (size_t)P.stage_bytes);
  uint64_t* slab_empty = slab_full + kSlabStages;
  uint64_t* acc_full = slab_empty + kSlabStages;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kMaxAcc);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kSlabStages; ++i) { ptx::mbar_init(&slab_full[i], 1); ptx::mbar_init(&slab_empty[i], 1); }
    for (int i = 0; i < kMaxAcc; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- one-time weight load: Wt[tap][co][ci] (bf16, global) -> TMEM, block (group, kb) at columns (g*nkb + kb)*32 ----
  // All eight epilogue warps share it (both sets of four cover the four lane quadrants) and the loads of the next block
  // are in flight while the current one is written: the chain of global-load latencies was 8 % of the cfg2 kernel.
  if (warp >= 4) {
    const int qd = warp & 3, m = qd * 32 + lane;
    const int nitems = P.ngroups * P.nkb;
    auto fetch = [&](int q, uint4 (&u)[8]) {
      const int g = q / P.nkb, kb = q - g * P.nkb;
      int tap, co;
      if (P.pair) { tap = (m & 16) ? P.tap_b[g] : P.tap_a[g]; co = (m >> 5) * 16 + (m & 15); }
      else { tap = P.tap_a[g]; co = m; }
      const bool row_ok = tap >= 0 && co < P.CO;
      const __nv_bfloat16* src = wt + ((int64_t)(row_ok ? tap : 0) * P.CO + (row_ok ? co : 0)) * P.CI + kb * 64;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        u[p] = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && kb * 64 + 8 * p < P.CI) u[p] = *reinterpret_cast<const uint4*>(src + 8 * p);
      }
    };
    uint4 cur[8], nxt[8];
    int q = (warp - 4) >> 2;
    if (q < nitems) fetch(q, cur);
    for (; q < nitems; q += 2) {
      if (q + 2 < nitems) fetch(q + 2, nxt);
      uint32_t v[32];
#pragma unroll
      for (int p = 0; p < 8; ++p) { v[4 * p] = cur[p].x; v[4 * p + 1] = cur[p].y; v[4 * p + 2] = cur[p].z; v[4 * p + 3] = cur[p].w; }
      tmem_st32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(q * 32), v);
#pragma unroll
      for (int p = 0; p < 8; ++p) cur[p] = nxt[p];
    }
    tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  const uint32_t acc_col0 = tmem_base + (uint32_t)P.wcols;

  if (warp == 0) {
    // ===== slab producer: whole padded rows [rho0, rho0 + SR) of channel block kb, one TMA box per row =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.amap);
      const uint32_t row_bytes = (uint32_t)P.Wp * kRowBytes;
      uint32_t p = 0;
      for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x) {
        const int rho0 = (unit * P.UT * P.NT) / P.Wp;
        for (int kb = 0; kb < P.nkb; ++kb, ++p) {
          const int st = p & 1;
          ptx::mbar_wait(&slab_empty[st], ((p >> 1) & 1u) ^ 1u);
          if ((P.dbg & 2) && p >= 2) { ptx::mbar_arrive(&slab_full[st]); continue; }
          ptx::mbar_expect_tx(&slab_full[st], (uint32_t)P.SR * row_bytes);
          uint8_t* dst = smem + (size_t)st * P.plane_bytes;
          int n = rho0 / P.Hp, yp = rho0 - n * P.Hp;
          for (int i = 0; i < P.SR; ++i) {
            ptx::tma_load_4d(dst + (size_t)i * row_bytes, &P.amap, &slab_full[st], kb * 64, -P.padx, yp - P.pady, n);
            if (++yp == P.Hp) { yp = 0; ++n; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: convergent loops, only the tcgen05 instructions under the elected lane =====
    const bool leader = ptx::elect_one();
    const int nkb = P.nkb, ngroups = P.ngroups, Wp = P.Wp, UT = P.UT, NT = P.NT, total_tiles = P.total_tiles;
    const uint32_t n_acc = (uint32_t)P.n_acc, NC = (uint32_t)P.NC, idesc = P.idesc;
    const uint32_t plane16 = (uint32_t)P.plane_bytes >> 4, tile16 = (uint32_t)NT * (kRowBytes >> 4);
    const uint64_t slab_d = kDesc64 + (ptx::smem_u32(smem) >> 4);
    uint32_t p = 0, slot0 = 0, ph0 = 0;          // slab ring sequence; accumulator slot / phase of the unit's first tile
    for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x) {
      const int s0 = unit * UT * NT;
      const int ntile = min(UT, total_tiles - unit * UT);
      const uint32_t a_off16 = (uint32_t)(s0 % Wp) * (kRowBytes >> 4);
      for (int kb = 0; kb < nkb; ++kb, ++p) {
        const uint32_t st = p & 1u;
        ptx::mbar_wait(&slab_full[st], (p >> 1) & 1u);
        ptx::tc_fence_after();
        const uint64_t b_base = slab_d + (st * plane16 + a_off16);
        for (int j = 0; j < ntile; ++j) {
          uint32_t sl = slot0 + (uint32_t)j, ph = ph0;
          if (sl >= n_acc) { sl -= n_acc; ph ^= 1u; }
          if (kb == 0) {
            ptx::mbar_wait(&acc_empty[sl], ph ^ 1u);
            ptx::tc_fence_after();
          }
          const uint32_t d = acc_col0 + sl * NC;
          const uint64_t bj = b_base + (uint64_t)((uint32_t)j * tile16);
          for (int g = 0; g < ngroups; ++g) {
            const uint64_t bd = bj + P.tap_off16[g];
            const uint32_t a = tmem_base + (uint32_t)((g * nkb + kb) * 32);
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_ts(d, a + 8u * ks, bd + (uint64_t)(2 * ks), idesc, (kb == 0 && g == 0 && ks == 0) ? 0u : 1u);
            }
          }
          if (leader && kb == nkb - 1) ptx::tc_commit(&acc_full[sl]);
        }
        if (leader) ptx::tc_commit(&slab_empty[st]);
      }
      slot0 += (uint32_t)ntile;
      if (slot0 >= n_acc) { slot0 -= n_acc; ph0 ^= 1u; }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    // A thread of quadrant qd reads two 16-lane halves (h = 0, 1) of its 32 TMEM lanes; register 4j + k of a half holds
    // row rq + 8 * (k / 2), column 8j + 2cq + (k % 2) of the 32-column chunk.  Pair mode: half 1 is the b tap of the same
    // channels, read delta columns to the right and added.  Single mode: half 1 holds 16 more channels.
    const int qd = warp & 3, grp = (warp - 4) >> 2;
    const int rq = lane >> 2, cq = lane & 3;
    const int tid_g = qd * 32 + lane;                       // thread of the group
    const bool pair = P.pair != 0;
    const int RB = pair ? 128 : 256, cpr_shift = pair ? 3 : 4;      // staging row bytes, log2(16-byte pieces per row)
    const uint32_t stg_u32 = ptx::smem_u32(stage + (size_t)grp * P.stage_bytes);
    const uint32_t ptab_u32 = stg_u32 + (uint32_t)(P.NT * RB);           // pixel table behind the staging tile
    const int bar_id = 1 + grp;
    int cos[4];
    float bvs[4], scs[4], shs[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {                           // q = 2h + k2 (pair mode uses q = 0, 1 only)
      const int co = pair ? 16 * qd + rq + 8 * (q & 1) : 32 * qd + 16 * (q >> 1) + rq + 8 * (q & 1);
      const bool ok = co < P.CO;
      cos[q] = co;
      bvs[q] = (bias != nullptr && ok) ? bias[co] : 0.f;
      scs[q] = (P.affine && ok) ? post_scale[co] : 1.f;
      shs[q] = (P.affine && ok) ? post_shift[co] : 0.f;
    }
    const int img_slots = P.Hp * P.Wp, NT = P.NT, UT = P.UT, Wp = P.Wp, CO = P.CO;
    const uint32_t n_acc = (uint32_t)P.n_acc, NC = (uint32_t)P.NC;
    const int npass = (Z != nullptr ? 1 : 0) + (Y != nullptr ? 1 : 0);
    uint32_t slot = 0, sph = 0, tseq = 0;
    for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x) {
      const int ntile = min(UT, P.total_tiles - unit * UT);
      for (int j = 0; j < ntile; ++j, ++tseq) {
        if ((tseq & 1u) == (uint32_t)grp) {
          const int s_tile = (unit * UT + j) * NT;
          ptx::mbar_wait(&acc_full[slot], sph);
          ptx::tc_fence_after();
          const uint32_t t_addr = acc_col0 + ((uint32_t)(qd * 32) << 16) + slot * NC;
          // output pixel of every slot of the tile, once per tile: ptab[p] = pixel index or -1 (junk slot)
          if (tid_g < NT) {
            const unsigned sidx = (unsigned)(s_tile + tid_g);
            const unsigned n = sidx / (unsigned)img_slots;
            const unsigned rem = sidx - n * (unsigned)img_slots;
            const unsigned yp = rem / (unsigned)Wp, xp = rem - yp * (unsigned)Wp;
            const int pix = (n < (unsigned)P.N && yp < (unsigned)P.OH && xp < (unsigned)P.OW)
                                ? (int)((n * (unsigned)P.OH + yp) * (unsigned)P.OW + xp) : -1;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(ptab_u32 + 4u * (uint32_t)tid_g), "r"(pix) : "memory");
          }
          for (int pass = 0; pass < npass; ++pass) {
            if (P.dbg & 1) {                                  // timing experiment: release the slot, touch nothing else
              if (pass == npass - 1) { ptx::tc_fence_before(); __syncwarp(); if (lane == 0) ptx::mbar_arrive(&acc_empty[slot]); }
              continue;
            }
            const bool is_y = (Z == nullptr) || pass == 1;
            const bool last_pass = pass == npass - 1;
            __nv_bfloat16* __restrict__ out = is_y ? Y : Z;
            auto load_chunk = [&](int cb, uint32_t (&ra)[16], uint32_t (&rb)[16]) {
              tmem_ld_16x256b_x4(t_addr + (uint32_t)cb, ra);
              tmem_ld_16x256b_x4(t_addr + (16u << 16) + (uint32_t)(cb + (pair ? P.delta : 0)), rb);
            };
            // chunk cb is in (ra, rb); the loads of chunk cb + 32 go into (na, nb) before cb is converted
            auto process = [&](int cb, uint32_t (&ra)[16], uint32_t (&rb)[16], uint32_t (&na)[16], uint32_t (&nb)[16]) {
              ptx::tmem_ld_wait();
              if (cb + 32 < NT) {
                load_chunk(cb + 32, na, nb);
              } else if (last_pass) {                         // accumulator fully read: hand the slot back
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&acc_empty[slot]);
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (pair && h == 1) continue;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int q = 2 * h + ((i >> 1) & 1);
                  float x = __uint_as_float(h == 0 ? ra[i] : rb[i]);
                  if (pair) x += __uint_as_float(rb[i]);
                  v[i] = x + bvs[q];
                }
                if (is_y) {
                  if (P.act == NPML_ACT_RELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.f;
                  } else if (P.act != NPML_ACT_IDENTITY) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                      const float4 r4 = act_fn4(P.act, P.p0, P.p1, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                      v[i] = r4.x; v[i + 1] = r4.y; v[i + 2] = r4.z; v[i + 3] = r4.w;
                    }
                  }
                  if (P.affine) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                      const int q = 2 * h + ((i >> 1) & 1);
                      v[i] = fmaf(v[i], scs[q], shs[q]);
                    }
                  }
                }
                // registers -> staging tile [pixel][channel] (bf16): pack the (pixel, pixel + 1) pairs of a channel and
                // store the eight 8 x 8 blocks transposed (stmatrix), 16-byte pieces XOR-swizzled with pixel & 7
                uint32_t pk[8];                               // pk[2 j + kh]: columns 8 j .. 8 j + 7, rows rq + 8 kh
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                  for (int kh = 0; kh < 2; ++kh) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4 * j + 2 * kh], v[4 * j + 2 * kh + 1]);
                    pk[2 * j + kh] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                // thread `lane` addresses row (lane & 7) of matrix (lane >> 3): matrices 0..3 = (j0, kh 0), (j0, kh 1),
                // (j0 + 1, kh 0), (j0 + 1, kh 1)
                const int mr = lane & 7, mm = lane >> 3;
                const int cbase = (pair ? 16 * qd : 32 * qd + 16 * h) + 8 * (mm & 1);
#pragma unroll
                for (int j0 = 0; j0 < 4; j0 += 2) {
                  const int pp = cb + 8 * (j0 + (mm >> 1)) + mr;
                  const uint32_t addr = stg_u32 + (uint32_t)(pp * RB + (((cbase >> 3) ^ (pp & 7)) << 4));
                  stmatrix_x4_trans(addr, pk[2 * j0], pk[2 * j0 + 1], pk[2 * j0 + 2], pk[2 * j0 + 3]);
                }
              }
            };
            {
              uint32_t a0[16], b0[16], a1[16], b1[16];
              load_chunk(0, a0, b0);
              for (int cb = 0; cb < NT; cb += 64) {
                process(cb, a0, b0, a1, b1);
                if (cb + 32 < NT) process(cb + 32, a1, b1, a0, b0);
              }
            }
            bar_sync(bar_id, 128);
            // staging -> global: 16 bytes per lane, a pixel's channels contiguous; a thread keeps its 16-byte column k
            {
              const unsigned k = (unsigned)tid_g & ((1u << cpr_shift) - 1u), pstep = 128u >> cpr_shift;
              if (8u * k < (unsigned)CO) {
                for (unsigned pp = (unsigned)tid_g >> cpr_shift; pp < (unsigned)NT; pp += pstep) {
                  int pix;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pix) : "r"(ptab_u32 + 4u * pp) : "memory");
                  if (pix >= 0) {
                    uint4 u;
                    const uint32_t addr = stg_u32 + pp * (uint32_t)RB + ((k ^ (pp & 7u)) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
                    *reinterpret_cast<uint4*>(out + ((int64_t)pix * CO + 8 * k)) = u;
                  }
                }
              }
            }
            bar_sync(bar_id, 128);
          }
        }
        if (++slot == n_acc) { slot = 0; sph ^= 1u; }
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

struct WsPlan {
  WsParams P;
  dim3 grid;
  size_t smem = 0, wt_bytes = 0;
};

bool plan_ws(const GatherConv& g, int mode, WsPlan& pl) {
  if (mode != NPML_MODE_BF16) return false;
  const char* opt = std::getenv("NPML_WS");
  if (opt == nullptr || opt[0] == '0') return false;       // opt-in while it is being measured
  if (g.s != 1 || g.N < 1) return false;
  if (g.CI % 8 || g.CO % 8 || g.CO > 128) return false;
  const int ntaps = g.fr * g.fc;
  if (ntaps > kMaxGroups) return false;
  WsParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  // tap offsets on the flat padded grid (same construction as tc_slab_conv.cu, stride 1)
  int dy[kMaxGroups], dx[kMaxGroups];
  int lo_y = 1 << 30, hi_y = -(1 << 30), lo_x = 1 << 30, hi_x = -(1 << 30);
  for (int r = 0; r < g.fr; ++r)
    for (int t = 0; t < g.fc; ++t) {
      const int i = r * g.fc + t;
      dy[i] = g.form == 0 ? r * g.D - g.pr : g.pr - r * g.D;
      dx[i] = g.form == 0 ? t * g.D - g.pc : g.pc - t * g.D;
      lo_y = dy[i] < lo_y ? dy[i] : lo_y; hi_y = dy[i] > hi_y ? dy[i] : hi_y;
      lo_x = dx[i] < lo_x ? dx[i] : lo_x; hi_x = dx[i] > hi_x ? dx[i] : hi_x;
    }
  const int Hp = g.OH + (hi_y - lo_y), Wp = g.OW + (hi_x - lo_x);
  if (Wp > 256) return false;
  const int64_t slots = (int64_t)g.N * Hp * Wp;
  if (slots + 4096 >= (1LL << 30)) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  P.N = g.N; P.OH = g.OH; P.OW = g.OW; P.CO = g.CO; P.CI = g.CI; P.Hp = Hp; P.Wp = Wp; P.pady = -lo_y; P.padx = -lo_x;
  P.nkb = ceildiv(g.CI, 64);
  P.pair = g.CO <= 64 ? 1 : 0;
  int halo = 0, ng = 0;
  if (P.pair) {
    // Taps of a filter row are D slots apart.  tcgen05.ld.16x256b needs an EVEN column offset (scripts/ts_probe2.cu),
    // so the partner of a tap is its neighbour when D is even and the tap after next when D is odd (a 3-tap row is
    // then (0, 2) + (1), a 4-tap row (0, 2) + (1, 3)).
    const int dist = (g.D % 2 == 0) ? 1 : 2;
    P.delta = dist * g.D;
    if (P.delta > 8) return false;
    for (int r = 0; r < g.fr; ++r) {
      bool used[kMaxGroups] = {false};
      for (int k = 0; k < g.fc; ++k) {          // k walks the row in the direction of increasing offset
        if (used[k]) continue;
        const int ia = r * g.fc + (g.form == 0 ? k : g.fc - 1 - k);
        int ib = -1;
        if (k + dist < g.fc && !used[k + dist]) {
          ib = r * g.fc + (g.form == 0 ? k + dist : g.fc - 1 - (k + dist));
          used[k + dist] = true;
        }
        used[k] = true;
        const int off = (dy[ia] - lo_y) * Wp + (dx[ia] - lo_x);
        P.tap_off16[ng] = (uint32_t)off * (kRowBytes >> 4);
        P.tap_a[ng] = (int16_t)ia;
        P.tap_b[ng] = (int16_t)ib;
        halo = off > halo ? off : halo;
        ++ng;
      }
    }
  } else {
    for (int i = 0; i < ntaps; ++i) {
      const int off = (dy[i] - lo_y) * Wp + (dx[i] - lo_x);
      P.tap_off16[ng] = (uint32_t)off * (kRowBytes >> 4);
      P.tap_a[ng] = (int16_t)i;
      P.tap_b[ng] = -1;
      halo = off > halo ? off : halo;
      ++ng;
    }
  }
  P.ngroups = ng;
  P.wcols = ng * P.nkb * 32;
  // accumulators: at least two of NC = NT + pad columns next to the weights; NT a multiple of 32 (epilogue chunks)
  const int pad = P.pair ? 8 : 0;
  // three accumulators when possible: with two, the MMAs of tile k + 2 wait for the epilogue to finish READING tile k
  int NT = 128;
  while (NT >= 64 && P.wcols + 3 * (NT + pad) > 512) NT -= 32;
  if (NT < 64) {
    NT = 128;
    while (NT >= 64 && P.wcols + 2 * (NT + pad) > 512) NT -= 32;
  }
  if (NT < 64) return false;
  if (const char* e = std::getenv("NPML_WS_NT")) {          // experiment knob
    const int v = atoi(e);
    if (v >= 32 && v <= 128 && v % 32 == 0 && P.wcols + 2 * (v + pad) <= 512) NT = v;
  }
  P.NT = NT; P.NC = NT + pad;
  P.n_acc = (512 - P.wcols) / P.NC;
  if (P.n_acc > kMaxAcc) P.n_acc = kMaxAcc;
  P.idesc = ptx::instr_desc(1, 0, 0, 128, P.NC);
  P.total_tiles = (int)((slots + NT - 1) / NT);
  P.stage_bytes = NT * (P.pair ? 128 : 256) + NT * 4;        // staging tile + pixel table
  const int fixed = 1024 /*barriers*/ + 2 * P.stage_bytes;
  bool found = false;
  for (int ut = P.n_acc < 4 ? P.n_acc : 4; ut >= 1 && !found; --ut) {
    const int sr = (Wp - 1 + NT * ut + pad + halo - 1) / Wp + 1;
    const int plane = sr * Wp * kRowBytes;
    if (fixed + kSlabStages * plane > kSmemLimit) continue;
    P.UT = ut; P.SR = sr; P.plane_bytes = plane;
    pl.smem = (size_t)fixed + (size_t)kSlabStages * plane;
    found = true;
  }
  if (!found) return false;
  P.num_units = ceildiv(P.total_tiles, P.UT);
  P.act = g.act; P.p0 = g.p0; P.p1 = g.p1;
  if (const char* e = std::getenv("NPML_WS_DBG")) P.dbg = atoi(e);
  P.affine = g.post_scale != nullptr ? 1 : 0;
  int ctas = num_sms();
  if (ctas > P.num_units) ctas = P.num_units;
  pl.grid = dim3((unsigned)ctas, 1, 1);
  pl.wt_bytes = align_up((size_t)ntaps * g.CO * g.CI * 2, 256);
  return true;
}

}  // namespace

bool ws_conv_supported(const GatherConv& g, int mode) {
  WsPlan pl;
  return plan_ws(g, mode, pl);
}

size_t ws_conv_ws_bytes(const GatherConv& g, int mode) {
  if (!ws_conv_supported(g, mode)) return 0;
  return align_up((size_t)g.fr * g.fc * g.CO * g.CI * 2, 256) + 256;
}

int ws_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z, void* Y,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  WsPlan pl;
  NPML_CHECK(plan_ws(g, mode, pl), "shape not supported by the weight-stationary path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.wt_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 127) == 0,
             "activation / workspace pointers must be 16 / 128 byte aligned");
  WsParams& P = pl.P;
  if (int e = launch_repack_weights(g, true, w, ws, st)) return e;
  {
    const uint32_t box[4] = {64u, (uint32_t)P.Wp, 1, 1};
    const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[4] = {2, (uint64_t)g.CI * 2, (uint64_t)g.IW * g.CI * 2, (uint64_t)g.IH * g.IW * g.CI * 2};
    if (int e = make_tmap(&P.amap, true, const_cast<void*>(in), 4, dims, strides, box, false)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    NPML_CUDA(cudaFuncSetAttribute(tc_ws_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  tc_ws_conv_kernel<<<pl.grid, kThreads, pl.smem, st>>>(P, (const __nv_bfloat16*)ws, bias, g.post_scale, g.post_shift,
                                                        (__nv_bfloat16*)Z, (__nv_bfloat16*)Y);
  NPML_LAUNCH_OK();
  return 0;
}

// Host-only test hook (npml_debug_ws_plan): what the planner decided, as plain integers.
//   out[0..15] = supported, pair, delta, NT, NC, n_acc, UT, SR, Hp, Wp, pady, padx, ngroups, wcols, total_tiles, smem bytes
//   then ngroups triples (tap offset in slots, tap_a, tap_b)
int ws_plan_debug(const GatherConv& g, int mode, int32_t* out, int32_t cap) {
  WsPlan pl;
  const bool ok = plan_ws(g, mode, pl);
  if (cap < 16) return -1;
  for (int i = 0; i < cap; ++i) out[i] = 0;
  out[0] = ok ? 1 : 0;
  if (!ok) return 16;
  const WsParams& P = pl.P;
  const int32_t head[16] = {1, P.pair, P.delta, P.NT, P.NC, P.n_acc, P.UT, P.SR, P.Hp, P.Wp, P.pady, P.padx, P.ngroups,
                            P.wcols, P.total_tiles, (int32_t)pl.smem};
  for (int i = 0; i < 16; ++i) out[i] = head[i];
  int n = 16;
  for (int gi = 0; gi < P.ngroups && n + 3 <= cap; ++gi) {
    out[n++] = (int32_t)(P.tap_off16[gi] / (kRowBytes >> 4));
    out[n++] = P.tap_a[gi];
    out[n++] = P.tap_b[gi];
  }
  return n;
}

}  // namespace npml
