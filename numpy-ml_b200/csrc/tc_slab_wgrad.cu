// Persistent "slab" weight-gradient kernel for stride-1 convolutions on sm_100a (bf16 operands):
//
//   dW[tap, ci, co] = sum_slots  X[slot + off(tap), ci] * dZ[slot, co]        (layers/layers.py:3106-3110)
//   db[co]          = sum_slots  1                      * dZ[slot, co]        (layers/layers.py:3109)
//
// Same flat padded-pixel ("slot") view as tc_slab_conv.cu: slot = (n*Hp + yp)*Wp + xp, off(tap) = r*D*Wp + t*D.
// The reduction runs over slots, so both operands are MN-major: a slot is one 128-byte row of 64 channels,
// exactly the K index of a 128B-swizzled MN-major tile.  One CTA keeps R padded rows of dZ and the R + halo
// rows of X in shared memory (TMA, zero padding = out-of-bounds fill, so junk slots of dZ are zero and add
// nothing) and feeds EVERY filter tap from that one X slab: the A descriptor of a tap is the slab's start
// address advanced by off(tap)*128 bytes.  An M = 128 tile stacks two (tap, channel-block) windows; the
// descriptor's leading-dimension offset is simply the byte distance between the two windows.  db rides along
// as one more window made of ones.  Accumulators for all taps stay in TMEM for the CTA's whole slot range;
// each CTA writes one fp32 partial and a second kernel adds the partials in a fixed order (deterministic).
//
// Stride s > 1 (strided Conv2D dW; Deconv2D dW, where the shifted operand is dZ -- cfg5): tap (r, t) reads row
// y*s + r*D - pr = s*(y + dy) + py of X, i.e. row y + dy of the input PARITY PLANE py.  The taps of one plane form a
// stride-1 problem on a strided view of X (a tensor map with s-fold strides), so every plane becomes its own group of
// M tiles with its own X slab, all on one common slot grid, in ONE launch.  The per-tile kernel (tc_wgrad.cu) needs
// 94 B/clk of L2 -> shared-memory traffic per SM at cfg5 against the 42 B/clk the chip delivers to 148 SMs at once
// (0.47 of the tensor peak); here a slot costs 512 B for 4 x 64 clk of MMAs.
#include <cstdlib>
#include <vector>

#include "npml_common.cuh"
#include "tc_ptx.cuh"
#include "wgrad_reduce.cuh"

namespace npml {

int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);
int launch_wgrad_reduce(const float* part, int splits, int M, int CO, int CI, int64_t o_tap, int64_t o_ci,
                        int64_t o_co, int accumulate, float* dW, cudaStream_t st, const float* part_db = nullptr,
                        float* db = nullptr);

namespace {

constexpr int kMaxMT = 40;           // M tiles (pairs of windows) over all groups
constexpr int kMaxMGroups = 16;
constexpr int kMaxPlanes = 9;        // input parity planes of a stride-2 / stride-3 convolution
constexpr int kRowBytes = 128;
constexpr int kKI = 16;              // slots (K) per tcgen05.mma for 16-bit operands
constexpr int kThreads = 256;
constexpr int kSmemLimit = 227 * 1024;
constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B

__device__ GridBar g_bar_slab, g_bar_ns;     // one per kernel (two launches of the same kernel never overlap: one stream)

struct SlabWgradParams {
  CUtensorMap xmap[kMaxPlanes], gmap;   // stride s > 1: one strided view of X per input parity plane (stride 1: one)
  uint32_t a16[kMaxMT];          // start of window 0 inside an X stage, >> 4 (kind 2: unused)
  uint32_t lbo16[kMaxMT];        // kind 0: byte distance window 0 -> window 1, >> 4
  int16_t tap[kMaxMT][2];        // filter tap of each window (-1 = the ones window, i.e. db)
  int16_t cb[kMaxMT][2];         // channel block of each window
  uint8_t kind[kMaxMT];          // 0 = (X, X), 1 = (X, ones), 2 = (ones, ones)
  int nmt_total, n_mgroups;
  // M tiles are cut into groups of at most 512 / BN accumulators; group g owns tiles [g_mt0, g_mt0 + g_nmt) and CTAs
  // [g_cta0, g_cta0 + g_nctas) of grid.x.  The CTA counts are PROPORTIONAL to the tile counts (a group with half the
  // tiles per K step gets half the CTAs), so all groups finish together; max_splits = the largest g_nctas.
  int16_t g_mt0[kMaxMGroups], g_nmt[kMaxMGroups], g_cta0[kMaxMGroups], g_nctas[kMaxMGroups];
  // the X plane a group reads and where the plane's slot (0, 0) sits in it (its taps' smallest offsets, negated)
  int16_t g_plane[kMaxMGroups], g_pady[kMaxMGroups], g_padx[kMaxMGroups];
  int max_splits;
  int ncb, nb, BN;
  int N, Hp, Wp, pady, padx, CI, CO, Mrows;
  int R, SRX, KS, num_units;
  int xplane_bytes, xstage_bytes, gplane_bytes, gstage_bytes;
  uint32_t idesc;
  int wait_ns;                 // sleep between mbarrier polls of the producer / epilogue warps (npml::wait_ns())
  FusedReduce red;             // the split reduction behind a grid barrier, inside this kernel (wgrad_reduce.cuh)
};

// The MMA issue loop for a group of NT accumulator tiles.  The whole warp walks the loops in convergent control
// flow so the descriptors live in uniform registers; one elected lane issues.  Tiles 0 .. NT-2 of any group are
// plain (X, X) pairs whose descriptor advances by a compile-time constant per K step; only the last tile of the
// last group can involve the ones window (kinds 1, 2) and carries run-time deltas.
template <int NT>
__device__ __forceinline__ void issue_units(const SlabWgradParams& P, int mt0, int bidx, int nctas, bool leader,
                                            uint32_t tmem_base, uint64_t* full_bar, uint64_t* empty_bar, uint64_t* acc_bar,
                                            uint32_t xs_u32, uint32_t gs_u32, uint32_t ones_u32) {
  const uint32_t idesc = P.idesc, BN = (uint32_t)P.BN;
  const int KS = P.KS, num_units = P.num_units;
  const uint32_t xs16 = xs_u32 >> 4, gs16 = gs_u32 >> 4, ones16 = ones_u32 >> 4;
  const uint32_t xstage16 = (uint32_t)P.xstage_bytes >> 4, gstage16 = (uint32_t)P.gstage_bytes >> 4;
  constexpr uint64_t kStep = (kKI * kRowBytes) >> 4;                 // descriptor advance of one K step
  constexpr uint64_t kHi64 = (uint64_t)kHi << 32;
  uint64_t d0[NT];                                                    // descriptor at (stage 0, k = 0)
#pragma unroll
  for (int i = 0; i < NT - 1; ++i) d0[i] = kHi64 | (uint64_t)((xs16 + P.a16[mt0 + i]) | (P.lbo16[mt0 + i] << 16));
  int64_t dk_last, ds_last;                                           // last tile: change per K step / per stage
  {
    const int gi = mt0 + NT - 1;
    const uint32_t kind = P.kind[gi];
    const uint32_t s = xs16 + P.a16[gi];
    if (kind == 0) {
      d0[NT - 1] = kHi64 | (uint64_t)(s | (P.lbo16[gi] << 16));
      dk_last = (int64_t)kStep; ds_last = (int64_t)xstage16;
    } else if (kind == 1) {       // window 1 = ones: its distance from window 0 shrinks as window 0 advances
      d0[NT - 1] = kHi64 | (uint64_t)(s | ((ones16 - s) << 16));
      dk_last = (int64_t)kStep - ((int64_t)kStep << 16); ds_last = (int64_t)xstage16 - ((int64_t)xstage16 << 16);
    } else {
      d0[NT - 1] = kHi64 | (uint64_t)ones16;
      dk_last = 0; ds_last = 0;
    }
  }
  const uint64_t g0 = kHi64 | (uint64_t)(gs16 | (((uint32_t)P.gplane_bytes >> 4) << 16));
  uint32_t uc = 0;
  for (int unit = bidx; unit < num_units; unit += nctas, ++uc) {
    const uint32_t st = uc & 1u;
    ptx::mbar_wait(&full_bar[st], (uc >> 1) & 1u);
    ptx::tc_fence_after();
    uint64_t ad[NT];
#pragma unroll
    for (int i = 0; i < NT - 1; ++i) ad[i] = d0[i] + (uint64_t)(st * xstage16);
    ad[NT - 1] = d0[NT - 1] + (uint64_t)(st ? ds_last : (int64_t)0);
    uint64_t bd = g0 + (uint64_t)(st * gstage16);
#pragma unroll 2
    for (int k = 0; k < KS; ++k) {
      if (leader) {
        const uint32_t acc = (uc | (uint32_t)k) ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < NT; ++i) ptx::mma_ss<false>(tmem_base + (uint32_t)i * BN, ad[i], bd, idesc, acc);
      }
#pragma unroll
      for (int i = 0; i < NT - 1; ++i) ad[i] += kStep;
      ad[NT - 1] += (uint64_t)dk_last;
      bd += kStep;
    }
    if (leader) ptx::tc_commit(&empty_bar[st]);
  }
  if (leader) ptx::tc_commit(acc_bar);
}

__global__ void __launch_bounds__(kThreads, 1) tc_slab_wgrad_kernel(const __grid_constant__ SlabWgradParams P,
                                                                    float* __restrict__ part,
                                                                    float* __restrict__ part_db) {
  // [X stage 0][X stage 1][dZ stage 0][dZ stage 1][ones][barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* xs = smem_raw;
  uint8_t* gs = xs + 2 * (size_t)P.xstage_bytes;
  uint8_t* ones = gs + 2 * (size_t)P.gstage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + kKI * kRowBytes);
  uint64_t* empty_bar = full_bar + 2;
  uint64_t* acc_bar = empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");

  // warp index broadcast from lane 0: tells ptxas it is warp-uniform, so the role branches are uniform branches and
  // everything the MMA issuer computes stays on the uniform datapath (no R2UR per descriptor)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int cotile = blockIdx.z;
  const int co0 = cotile * P.BN;
  int mgroup = 0;                                         // (a scan, no division: keeps the indices warp-uniform)
  for (int gq = 1; gq < P.n_mgroups; ++gq)
    if ((int)blockIdx.x >= P.g_cta0[gq]) mgroup = gq;
  const int mt0 = P.g_mt0[mgroup], nmt = P.g_nmt[mgroup];
  const int bidx = (int)blockIdx.x - P.g_cta0[mgroup], nctas = P.g_nctas[mgroup];
  const int rows_g = P.R * P.Wp;                 // dZ slots the TMA writes per plane

  // constant regions: the ones window and the zero slack behind every dZ plane (K steps may overrun the unit)
  for (int i = threadIdx.x; i < kKI * kRowBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  for (int pl = 0; pl < 2 * P.nb; ++pl) {
    uint32_t* z = reinterpret_cast<uint32_t*>(gs + (size_t)pl * P.gplane_bytes + (size_t)rows_g * kRowBytes);
    for (int i = threadIdx.x; i < kKI * kRowBytes / 4; i += kThreads) z[i] = 0u;
  }
  ptx::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    ptx::mbar_init(acc_bar, 1);
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

  if (warp == 0 || warp == 3) {
    // ===== two producers: warp 0 loads the X rows [u*R, u*R + SRX) of every channel block, warp 3 the dZ rows
    // [u*R, u*R + R) of the CO tile.  One elected thread issues a TMA box per padded row and plane; at Wp = 34 (cfg4)
    // that is ~50 boxes of 4 KB per unit, and the issue loop of ONE thread was what the MMAs waited for: cfg4 conv2's
    // wgrad 105.6 -> 94.7 us, conv1's 60.5 -> 53.5 us with two issuing warps (`gpurun_out/r3x_*`).  One three-dimensional
    // box per plane and unit (units cut at image boundaries, 4 boxes instead of 50) was measured too: 97.8 / 54.2 us --
    // the ragged last unit of every image costs what the shorter issue loop saves.  Warp 0 announces the bytes of both.
    if (ptx::elect_one()) {
      const bool xside = warp == 0;
      const CUtensorMap* xmap = &P.xmap[P.g_plane[mgroup]];
      const int pady = P.g_pady[mgroup], padx = P.g_padx[mgroup];
      ptx::prefetch_tmap(xside ? xmap : &P.gmap);
      const uint32_t row_bytes = (uint32_t)P.Wp * kRowBytes;
      const uint32_t tx = (uint32_t)(P.ncb * P.SRX + P.nb * P.R) * row_bytes;
      uint32_t uc = 0;
      for (int unit = bidx; unit < P.num_units; unit += nctas, ++uc) {
        const uint32_t st = uc & 1u;
        ptx::mbar_wait_sleep(&empty_bar[st], ((uc >> 1) & 1u) ^ 1u, (uint32_t)P.wait_ns);
        if (xside) ptx::mbar_expect_tx(&full_bar[st], tx);
        const int rho0 = unit * P.R;
        int n = rho0 / P.Hp, yp = rho0 - n * P.Hp;
        uint8_t* xd = xs + (size_t)st * P.xstage_bytes;
        uint8_t* gd = gs + (size_t)st * P.gstage_bytes;
        if (xside) {
          for (int i = 0; i < P.SRX; ++i) {
            for (int c = 0; c < P.ncb; ++c)
              ptx::tma_load_4d(xd + (size_t)c * P.xplane_bytes + (size_t)i * row_bytes, xmap, &full_bar[st], c * 64, -padx,
                               yp - pady, n);
            if (++yp == P.Hp) { yp = 0; ++n; }
          }
        } else {
          for (int i = 0; i < P.R; ++i) {
            for (int b = 0; b < P.nb; ++b)
              ptx::tma_load_4d(gd + (size_t)b * P.gplane_bytes + (size_t)i * row_bytes, &P.gmap, &full_bar[st], co0 + b * 64, 0,
                               yp, n);
            if (++yp == P.Hp) { yp = 0; ++n; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (issue_units<NT>, NT = tiles of this CTA's group, compile-time) =====
    // shared-window addresses as plain integers off the (link-time constant) base of the dynamic smem array,
    // so that ptxas sees them as warp-uniform (a cvta of a computed generic pointer is not)
    const bool leader = ptx::elect_one();
    const uint32_t xs_u32 = ptx::smem_u32(smem_raw);
    const uint32_t gs_u32 = xs_u32 + 2u * (uint32_t)P.xstage_bytes;
    const uint32_t ones_u32 = gs_u32 + 2u * (uint32_t)P.gstage_bytes;
    switch (nmt) {
      case 1: issue_units<1>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 2: issue_units<2>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 3: issue_units<3>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 4: issue_units<4>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 5: issue_units<5>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 6: issue_units<6>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 7: issue_units<7>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      default: issue_units<8>(P, mt0, bidx, nctas, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
    }
  } else if (warp >= 4) {
    // ===== epilogue (once): TMEM -> this CTA's fp32 partial =====
    const int qd = warp & 3;
    const int l = qd * 32 + lane;                 // accumulator row
    const int w = l >> 6, cl = l & 63;            // window of the pair, channel inside the window
    ptx::mbar_wait_sleep(acc_bar, 0, (uint32_t)P.wait_ns * 4u);      // these warps wait for the whole reduction
    ptx::tc_fence_after();
    for (int i = 0; i < nmt; ++i) {
      const int gi = mt0 + i;
      const int tap = P.tap[gi][w];
      const int ci = P.cb[gi][w] * 64 + cl;
      float* out = nullptr;
      int64_t split_stride = 0;                    // distance between two split slots of this row
      if (tap >= 0) {
        if (ci < P.CI) out = part + ((int64_t)bidx * P.Mrows + (int64_t)tap * P.CI + ci) * P.CO;
        split_stride = (int64_t)P.Mrows * P.CO;
      } else if (tap == -1 && cl == 0 && (w == 0 || P.kind[gi] == 1)) {
        out = part_db + (int64_t)bidx * P.CO;
        split_stride = P.CO;
      }
      for (int c = 0; c < P.BN; c += 32) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(i * P.BN + c), r);
        ptx::tmem_ld_wait();
        const int co = co0 + c;
        if (out != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (co + 4 * j < P.CO)
              *reinterpret_cast<uint4*>(out + co + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          // the fixed-order reduction adds max_splits slots of every row: this group has fewer CTAs than the largest
          // one, so its unused slots are written as zeros (by the CTAs it does have, round robin)
          for (int sidx = bidx + nctas; sidx < P.max_splits; sidx += nctas) {
            float* zo = out + (int64_t)(sidx - bidx) * split_stride;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (co + 4 * j < P.CO) *reinterpret_cast<uint4*>(zo + co + 4 * j) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
  if (P.red.enabled) {
    // the operand stages are free now: their first bytes serve as the reduction's scratch
    const int nctas = (int)(gridDim.x * gridDim.z), cta = (int)(blockIdx.x + gridDim.x * blockIdx.z);
    grid_barrier(&g_bar_slab, (unsigned)nctas);
    wgrad_reduce_slices(P.red, part, part_db, P.Mrows, P.CO, P.CI, cta, nctas, reinterpret_cast<float4*>(smem_raw));
  }
}

inline int ceildiv(int a, int b) { return (a + b - 1) / b; }

// The in-kernel reduction needs every CTA of the grid resident at once (it meets at a grid barrier): one CTA per SM,
// so at most as many CTAs as the device has SMs.  NPML_DISABLE_FUSED_REDUCE=1: the separate reduce kernel (A/B).
inline FusedReduce fused_reduce(const GatherWgrad& g, dim3 grid, int splits, float* dW, float* db) {
  FusedReduce r{};
  r.dW = dW; r.db = db; r.o_tap = g.o_tap; r.o_ci = g.o_ci; r.o_co = g.o_co; r.accumulate = g.accumulate; r.splits = splits;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
  // (o_co != 1: Deconv2D's transposed dW, whose stores want the shared-memory transpose of wgrad_reduce_t_kernel)
  r.enabled = (g.CO % 4 == 0 && g.o_co == 1 && ctas <= sms && std::getenv("NPML_DISABLE_FUSED_REDUCE") == nullptr) ? 1 : 0;
  return r;
}

struct SlabWgradPlan {
  SlabWgradParams P;
  dim3 grid;
  size_t smem = 0, part_bytes = 0, partdb_bytes = 0;
  int splits = 0;
  int nplanes = 1, plane_py[kMaxPlanes] = {0}, plane_px[kMaxPlanes] = {0};
};

inline int posmod(int a, int b) { const int m = a % b; return m < 0 ? m + b : m; }
inline int floordiv(int a, int b) { return (a - posmod(a, b)) / b; }

bool plan_slab_wgrad(const GatherWgrad& g, int mode, SlabWgradPlan& pl) {
  if (mode != NPML_MODE_BF16) return false;       // MN-major tf32 needs the 32B-atom layout: tc_wgrad.cu
  if (std::getenv("NPML_DISABLE_SLAB")) return false;
  if (g.s < 1 || g.s > 3 || g.N < 1) return false;
  // Opt-in (NPML_SLAB_WGRAD_PLANES=1): validated (parity tests run it) but it LOSES to the per-tile kernel at cfg5,
  // 58 + 22 us (kernel + 18-way reduce) against 44 + 9 us.  ncu (profiles/r3l_*): the MMAs need 33 k of the 73 k cycles,
  // the issuer waits for operands -- a strided plane is a gather of 128-byte granules (pixel stride s * CI * 2 bytes) and
  // the TMA delivers those at ~15 B/clk per SM, half of what 4 accumulator tiles consume (the same limit finding 16
  // met at cfg4 conv2, whose 128-channel rows are 128-byte granules 256 bytes apart).
  if (g.s > 1 && !std::getenv("NPML_SLAB_WGRAD_PLANES")) return false;
  if (g.CI % 8 || g.CO % 8) return false;
  // ---- taps by input parity plane: row y*s + r*D - pr = s*(y + dy) + py --------------------------------------
  struct Tap { int id, dy, dx; };
  struct Plane { int py, px, lo_y, lo_x; std::vector<Tap> taps; };
  std::vector<Plane> planes;
  int span_y = 0, span_x = 0;
  for (int py = 0; py < g.s; ++py)
    for (int px = 0; px < g.s; ++px) {
      if (py >= g.IH || px >= g.IW) continue;
      Plane pn{py, px, 1 << 30, 1 << 30, {}};
      int hi_y = -(1 << 30), hi_x = -(1 << 30);
      for (int r = 0; r < g.fr; ++r)
        for (int t = 0; t < g.fc; ++t) {
          const int oy = r * g.D - g.pr, ox = t * g.D - g.pc;
          if (posmod(oy, g.s) != py || posmod(ox, g.s) != px) continue;
          const int dy = floordiv(oy, g.s), dx = floordiv(ox, g.s);
          pn.taps.push_back({r * g.fc + t, dy, dx});
          pn.lo_y = dy < pn.lo_y ? dy : pn.lo_y; hi_y = dy > hi_y ? dy : hi_y;
          pn.lo_x = dx < pn.lo_x ? dx : pn.lo_x; hi_x = dx > hi_x ? dx : hi_x;
        }
      if (pn.taps.empty()) continue;
      span_y = hi_y - pn.lo_y > span_y ? hi_y - pn.lo_y : span_y;
      span_x = hi_x - pn.lo_x > span_x ? hi_x - pn.lo_x : span_x;
      planes.push_back(pn);
    }
  if (planes.empty() || (int)planes.size() > kMaxPlanes) return false;
  const int Hp = g.OH + span_y, Wp = g.OW + span_x;
  if (Wp > 256) return false;
  // the flat slot grid computes Hp*Wp slots per OH*OW outputs: small strided planes under a wide filter footprint waste
  // too much (cfg3: 14x14 slots per 11x11 outputs); the per-tile kernel keeps those
  if (g.s > 1 && (double)Hp * Wp > 1.3 * g.OH * g.OW) return false;
  if ((int64_t)g.N * Hp * Wp >= (1LL << 30)) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  SlabWgradParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.N = g.N; P.Hp = Hp; P.Wp = Wp; P.pady = -planes[0].lo_y; P.padx = -planes[0].lo_x; P.CI = g.CI; P.CO = g.CO;
  P.Mrows = g.fr * g.fc * g.CI;
  P.ncb = ceildiv(g.CI, 64);
  int bn = ceildiv(g.CO, 64) * 64;
  if (bn > 128) bn = 128;
  P.BN = bn;
  P.nb = bn / 64;
  const int ncot = ceildiv(g.CO, bn);
  P.idesc = ptx::instr_desc(1, 1, 1, 128, bn);
  P.wait_ns = wait_ns();
  const int halo = span_y * Wp + span_x;
  // largest unit (R padded rows) whose two stages fit
  int R = 0;
  for (int r = 32; r >= 1; --r) {
    const int ks = ceildiv(r * Wp, kKI);
    const int srx = ceildiv(ks * kKI + halo, Wp);
    const size_t xplane = (size_t)srx * Wp * kRowBytes;
    const size_t gplane = (size_t)(r * Wp + kKI) * kRowBytes;
    const size_t need = 2 * (P.ncb * xplane + P.nb * gplane) + kKI * kRowBytes + 256;
    if (need <= (size_t)kSmemLimit) {
      R = r; P.KS = ks; P.SRX = srx; P.xplane_bytes = (int)xplane; P.gplane_bytes = (int)gplane;
      P.xstage_bytes = (int)(P.ncb * xplane); P.gstage_bytes = (int)(P.nb * gplane);
      pl.smem = need;
      break;
    }
  }
  if (R == 0) return false;
  P.R = R;
  P.num_units = ceildiv(g.N * Hp, R);
  // ---- per plane: every (channel block, tap) window in ascending smem address order, paired into M tiles; the ones
  // window (db) closes the LAST plane's list, so the only tile with run-time descriptor deltas is the last tile of the
  // last group (what issue_units assumes).  A plane's tiles are cut into groups of at most 512 / BN accumulators.
  const int cap = 512 / bn > 8 ? 8 : 512 / bn;   // issue_units<NT> is instantiated up to 8
  int nmt = 0, ngroups = 0;
  int group_tiles[kMaxMGroups];
  for (size_t q = 0; q < planes.size(); ++q) {
    const Plane& pn = planes[q];
    struct Win { int tap, cb; uint32_t a16; };
    std::vector<Win> wins;
    for (int c = 0; c < P.ncb; ++c) {
      std::vector<Win> blk;
      for (const Tap& tp : pn.taps)
        blk.push_back({tp.id, c, (uint32_t)(((size_t)c * P.xplane_bytes + (size_t)((tp.dy - pn.lo_y) * Wp + (tp.dx - pn.lo_x)) * kRowBytes) >> 4)});
      for (size_t i = 1; i < blk.size(); ++i)            // ascending address (insertion sort: a handful of taps)
        for (size_t j = i; j > 0 && blk[j].a16 < blk[j - 1].a16; --j) std::swap(blk[j], blk[j - 1]);
      wins.insert(wins.end(), blk.begin(), blk.end());
    }
    if (q + 1 == planes.size() && !g.no_db) wins.push_back({-1, 0, 0});
    if (wins.size() & 1) wins.push_back({-1, 0, 0});    // (a zero-weight filler: tap -1 rows are never stored twice)
    const int tiles_q = (int)wins.size() / 2;
    if (nmt + tiles_q > kMaxMT) return false;
    for (int i = 0; i < tiles_q; ++i) {
      const Win &a = wins[2 * i], &b = wins[2 * i + 1];
      const int gi = nmt + i;
      P.tap[gi][0] = (int16_t)a.tap; P.tap[gi][1] = (int16_t)b.tap;
      P.cb[gi][0] = (int16_t)a.cb; P.cb[gi][1] = (int16_t)b.cb;
      P.a16[gi] = a.a16;
      if (a.tap >= 0 && b.tap >= 0) { P.kind[gi] = 0; P.lbo16[gi] = b.a16 - a.a16; if (P.lbo16[gi] >= (1u << 14)) return false; }
      else if (a.tap >= 0) P.kind[gi] = 1;
      else P.kind[gi] = 2;
      // a filler / ones window anywhere but in the last tile of the last plane would break issue_units' assumption
      if (P.kind[gi] != 0 && !(q + 1 == planes.size() && i + 1 == tiles_q)) return false;
    }
    const int ng_q = ceildiv(tiles_q, cap);
    if (ngroups + ng_q > kMaxMGroups) return false;
    int t0 = 0;
    for (int k = 0; k < ng_q; ++k) {
      const int tiles = tiles_q / ng_q + (k < tiles_q % ng_q ? 1 : 0);
      P.g_mt0[ngroups] = (int16_t)(nmt + t0); P.g_nmt[ngroups] = (int16_t)tiles;
      P.g_plane[ngroups] = (int16_t)q; P.g_pady[ngroups] = (int16_t)(-pn.lo_y); P.g_padx[ngroups] = (int16_t)(-pn.lo_x);
      group_tiles[ngroups] = tiles;
      t0 += tiles; ++ngroups;
    }
    nmt += tiles_q;
  }
  P.nmt_total = nmt;
  P.n_mgroups = ngroups;
  // CTAs in proportion to the tiles of a group (measured on cfg4 conv2: 10 tiles as 4 + 4 + 2 with 49 CTAs each left a
  // third of the SMs idle for half of the kernel)
  int budget = num_sms() / ncot;
  if (budget < ngroups) budget = ngroups;
  int done = 0, cta0 = 0, max_splits = 0;
  for (int q = 0; q < ngroups; ++q) {
    const int tiles = group_tiles[q];
    int ctas = (int)((int64_t)budget * (done + tiles) / nmt) - (int)((int64_t)budget * done / nmt);
    if (ctas < 1) ctas = 1;
    if (ctas > P.num_units) ctas = P.num_units;
    P.g_cta0[q] = (int16_t)cta0; P.g_nctas[q] = (int16_t)ctas;
    done += tiles; cta0 += ctas;
    max_splits = ctas > max_splits ? ctas : max_splits;
  }
  P.max_splits = max_splits;
  pl.splits = max_splits;
  pl.grid = dim3((unsigned)cta0, 1, (unsigned)ncot);
  pl.part_bytes = align_up((size_t)max_splits * P.Mrows * g.CO * sizeof(float), 256);
  pl.partdb_bytes = align_up((size_t)max_splits * g.CO * sizeof(float), 256);
  pl.nplanes = (int)planes.size();
  for (size_t q = 0; q < planes.size(); ++q) { pl.plane_py[q] = planes[q].py; pl.plane_px[q] = planes[q].px; }
  return true;
}


// =====================================================================================================================
//  Narrow outputs (out_ch <= 64): column-shifted dZ windows stacked along N
// =====================================================================================================================
// With N = out_ch = 64 a 128x64x16 tcgen05.mma fetches 6 KB of operands for 32 clk of math and is paced by the
// shared-memory port at 48 clk (profiles/r1g_mma_rate.txt).  Both wgrad operands are spatial, so shifts compose:
//
//     sum_s X[s + a, ci] * dZ[s + b, co]  =  dW[tap at offset a - b][ci, co]
//
// The B operand of ONE MMA is therefore fc windows of the same dZ slab, shifted by j*D slots (j = 0 .. fc-1) and
// stacked along N through the descriptor's leading-dimension offset (D*128 bytes: the windows overlap in memory),
// while the A operand stacks two windows of the X slab that sit at filter COLUMN fc-1 of two filter rows.  A
// 128 x (fc*64) x 16 MMA then yields the 2 x fc taps {(r0, fc-1-j), (r1, fc-1-j)}: for a 3x3 filter two MMAs per K step
// (rows {0, 1}; row 2 + the ones window for db) of 10 KB operands / 96 clk math each instead of five of 6 KB / 32 clk.
// Work units are ranges of KS*16 consecutive slots of the flat padded-slot sequence (not whole rows), so no K step
// ever overruns its unit; the sequence starts (fc-1)*D slots early so that every shifted window sees every slot.
struct NsParams {
  CUtensorMap xmap, gmap;
  uint32_t a16[kMaxMT];
  uint32_t lbo16[kMaxMT];
  int16_t row[kMaxMT][2];        // filter row of each A window (-1 = the ones window, i.e. db)
  int16_t cb[kMaxMT][2];
  uint8_t kind[kMaxMT];          // 0 = (X, X), 1 = (X, ones), 2 = (ones, ones)
  int nmt_total, mt_per_group, n_mgroups;
  int ncb, fc, NN, D, shift;
  int N, Hp, Wp, pady, padx, CI, CO, Mrows;
  int SRX, SRG, KS, num_units;
  int xplane_bytes, xstage_bytes, gstage_bytes;
  uint32_t idesc;
  int wait_ns;
  FusedReduce red;
};

__device__ __forceinline__ int ns_floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

template <int NT>
__device__ __forceinline__ void ns_issue(const NsParams& P, int mt0, bool leader, uint32_t tmem_base, uint64_t* full_bar,
                                         uint64_t* empty_bar, uint64_t* acc_bar, uint32_t xs_u32, uint32_t gs_u32,
                                         uint32_t ones_u32) {
  const uint32_t idesc = P.idesc, NN = (uint32_t)P.NN;
  const int KS = P.KS, num_units = P.num_units, Wp = P.Wp, shift = P.shift;
  const uint32_t xs16 = xs_u32 >> 4, gs16 = gs_u32 >> 4, ones16 = ones_u32 >> 4;
  const uint32_t xstage16 = (uint32_t)P.xstage_bytes >> 4, gstage16 = (uint32_t)P.gstage_bytes >> 4;
  constexpr uint64_t kStep = (kKI * kRowBytes) >> 4;
  constexpr uint64_t kHi64 = (uint64_t)kHi << 32;
  uint64_t d0[NT];
#pragma unroll
  for (int i = 0; i < NT - 1; ++i) d0[i] = kHi64 | (uint64_t)((xs16 + P.a16[mt0 + i]) | (P.lbo16[mt0 + i] << 16));
  int64_t dk_last, ds_last, do_last;          // last tile: change per K step / per stage / per 16 bytes of unit offset
  {
    const int gi = mt0 + NT - 1;
    const uint32_t kind = P.kind[gi];
    const uint32_t s = xs16 + P.a16[gi];
    if (kind == 0) {
      d0[NT - 1] = kHi64 | (uint64_t)(s | (P.lbo16[gi] << 16));
      dk_last = (int64_t)kStep; ds_last = (int64_t)xstage16; do_last = 1;
    } else if (kind == 1) {       // window 1 = ones: its distance from window 0 shrinks as window 0 advances
      d0[NT - 1] = kHi64 | (uint64_t)(s | ((ones16 - s) << 16));
      dk_last = (int64_t)kStep - ((int64_t)kStep << 16); ds_last = (int64_t)xstage16 - ((int64_t)xstage16 << 16);
      do_last = 1 - ((int64_t)1 << 16);
    } else {
      d0[NT - 1] = kHi64 | (uint64_t)ones16;
      dk_last = 0; ds_last = 0; do_last = 0;
    }
  }
  const uint64_t g0 = kHi64 | (uint64_t)(gs16 | (((uint32_t)P.D * (kRowBytes >> 4)) << 16));
  uint32_t uc = 0;
  for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++uc) {
    const uint32_t st = uc & 1u;
    const int slot0 = unit * KS * kKI - shift;
    const uint32_t off16 = (uint32_t)(slot0 - ns_floordiv(slot0, Wp) * Wp) * (kRowBytes >> 4);
    ptx::mbar_wait(&full_bar[st], (uc >> 1) & 1u);
    ptx::tc_fence_after();
    uint64_t ad[NT];
#pragma unroll
    for (int i = 0; i < NT - 1; ++i) ad[i] = d0[i] + (uint64_t)(st * xstage16 + off16);
    ad[NT - 1] = d0[NT - 1] + (uint64_t)((st ? ds_last : (int64_t)0) + do_last * (int64_t)off16);
    uint64_t bd = g0 + (uint64_t)(st * gstage16 + off16);
#pragma unroll 2
    for (int k = 0; k < KS; ++k) {
      if (leader) {
        const uint32_t acc = (uc | (uint32_t)k) ? 1u : 0u;
#pragma unroll
        for (int i = 0; i < NT; ++i) ptx::mma_ss<false>(tmem_base + (uint32_t)i * NN, ad[i], bd, idesc, acc);
      }
#pragma unroll
      for (int i = 0; i < NT - 1; ++i) ad[i] += kStep;
      ad[NT - 1] += (uint64_t)dk_last;
      bd += kStep;
    }
    if (leader) ptx::tc_commit(&empty_bar[st]);
  }
  if (leader) ptx::tc_commit(acc_bar);
}

__global__ void __launch_bounds__(kThreads, 1) tc_slab_wgrad_ns_kernel(const __grid_constant__ NsParams P,
                                                                       float* __restrict__ part,
                                                                       float* __restrict__ part_db) {
  // [X stage 0][X stage 1][dZ stage 0][dZ stage 1][ones][barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* xs = smem_raw;
  uint8_t* gs = xs + 2 * (size_t)P.xstage_bytes;
  uint8_t* ones = gs + 2 * (size_t)P.gstage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + kKI * kRowBytes);
  uint64_t* empty_bar = full_bar + 2;
  uint64_t* acc_bar = empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int mt0 = blockIdx.y * P.mt_per_group;
  const int nmt = min(P.mt_per_group, P.nmt_total - mt0);

  for (int i = threadIdx.x; i < kKI * kRowBytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  ptx::fence_proxy_async();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&full_bar[i], 1); ptx::mbar_init(&empty_bar[i], 1); }
    ptx::mbar_init(acc_bar, 1);
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

  if (warp == 0) {
    // ===== producer: the padded rows that hold slots [slot0, slot0 + KS*16 + halo) of X and [.., + shift) of dZ =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.xmap);
      ptx::prefetch_tmap(&P.gmap);
      const uint32_t row_bytes = (uint32_t)P.Wp * kRowBytes;
      const uint32_t tx = (uint32_t)(P.ncb * P.SRX + P.SRG) * row_bytes;
      const int nrow = P.SRX > P.SRG ? P.SRX : P.SRG;
      uint32_t uc = 0;
      for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x, ++uc) {
        const uint32_t st = uc & 1u;
        ptx::mbar_wait_sleep(&empty_bar[st], ((uc >> 1) & 1u) ^ 1u, (uint32_t)P.wait_ns);
        ptx::mbar_expect_tx(&full_bar[st], tx);
        const int slot0 = unit * P.KS * kKI - P.shift;
        const int rho0 = ns_floordiv(slot0, P.Wp);
        int n = ns_floordiv(rho0, P.Hp), yp = rho0 - n * P.Hp;      // n = -1 / n >= N: out of bounds, zero-filled
        uint8_t* xd = xs + (size_t)st * P.xstage_bytes;
        uint8_t* gd = gs + (size_t)st * P.gstage_bytes;
        for (int i = 0; i < nrow; ++i) {
          if (i < P.SRX)
            for (int c = 0; c < P.ncb; ++c)
              ptx::tma_load_4d(xd + (size_t)c * P.xplane_bytes + (size_t)i * row_bytes, &P.xmap, &full_bar[st], c * 64, -P.padx,
                               yp - P.pady, n);
          if (i < P.SRG) ptx::tma_load_4d(gd + (size_t)i * row_bytes, &P.gmap, &full_bar[st], 0, 0, yp, n);
          if (++yp == P.Hp) { yp = 0; ++n; }
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t xs_u32 = ptx::smem_u32(smem_raw);
    const uint32_t gs_u32 = xs_u32 + 2u * (uint32_t)P.xstage_bytes;
    const uint32_t ones_u32 = gs_u32 + 2u * (uint32_t)P.gstage_bytes;
    switch (nmt) {
      case 1: ns_issue<1>(P, mt0, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 2: ns_issue<2>(P, mt0, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      case 3: ns_issue<3>(P, mt0, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
      default: ns_issue<4>(P, mt0, leader, tmem_base, full_bar, empty_bar, acc_bar, xs_u32, gs_u32, ones_u32); break;
    }
  } else if (warp >= 4) {
    // ===== epilogue (once): TMEM -> this CTA's fp32 partial; column block j of a tile is filter column fc-1-j =====
    const int qd = warp & 3;
    const int l = qd * 32 + lane;
    const int w = l >> 6, cl = l & 63;
    ptx::mbar_wait_sleep(acc_bar, 0, (uint32_t)P.wait_ns * 4u);
    ptx::tc_fence_after();
    for (int i = 0; i < nmt; ++i) {
      const int gi = mt0 + i;
      const int r = P.row[gi][w];
      const int ci = P.cb[gi][w] * 64 + cl;
      const bool is_db = r == -1 && cl == 0 && (w == 0 || P.kind[gi] == 1);
      for (int c = 0; c < P.NN; c += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(i * P.NN + c), v);
        ptx::tmem_ld_wait();
        const int j = c >> 6, co = c & 63;
        float* out = nullptr;
        if (r >= 0) {
          if (ci < P.CI) out = part + ((int64_t)blockIdx.x * P.Mrows + (int64_t)(r * P.fc + (P.fc - 1 - j)) * P.CI + ci) * P.CO;
        } else if (is_db && j == 0) {
          out = part_db + (int64_t)blockIdx.x * P.CO;
        }
        if (out != nullptr) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (co + 4 * q < P.CO)
              *reinterpret_cast<uint4*>(out + co + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
  if (P.red.enabled) {
    const int nctas = (int)(gridDim.x * gridDim.y), cta = (int)(blockIdx.x + gridDim.x * blockIdx.y);
    grid_barrier(&g_bar_ns, (unsigned)nctas);
    wgrad_reduce_slices(P.red, part, part_db, P.Mrows, P.CO, P.CI, cta, nctas, reinterpret_cast<float4*>(smem_raw));
  }
}

struct NsPlan {
  NsParams P;
  dim3 grid;
  size_t smem = 0, part_bytes = 0, partdb_bytes = 0;
  int splits = 0;
};

bool plan_ns_wgrad(const GatherWgrad& g, int mode, NsPlan& pl) {
  if (mode != NPML_MODE_BF16 || std::getenv("NPML_DISABLE_SLAB") || std::getenv("NPML_DISABLE_WGRAD_NS")) return false;
  if (g.s != 1 || g.N < 1 || g.CI % 8 || g.CO % 8) return false;
  if (g.CO > 64 || g.fc * 64 > 256) return false;       // N = fc * 64 accumulator columns per tile
  if (g.fc < 2) return false;                            // nothing to stack: the plain kernel
  const int Hp = g.OH + (g.fr - 1) * g.D, Wp = g.OW + (g.fc - 1) * g.D;
  if (Wp > 256) return false;
  if ((int64_t)g.N * Hp * Wp >= (1LL << 30)) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  NsParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.N = g.N; P.Hp = Hp; P.Wp = Wp; P.pady = g.pr; P.padx = g.pc; P.CI = g.CI; P.CO = g.CO;
  P.Mrows = g.fr * g.fc * g.CI;
  P.ncb = ceildiv(g.CI, 64);
  P.fc = g.fc; P.NN = g.fc * 64; P.D = g.D; P.shift = (g.fc - 1) * g.D;
  P.idesc = ptx::instr_desc(1, 1, 1, 128, P.NN);
  P.wait_ns = wait_ns();
  const int halo = (g.fr - 1) * g.D * Wp + P.shift;
  int KS = 0;
  for (int ks = 64; ks >= 4; --ks) {
    const int srx = ceildiv(Wp - 1 + ks * kKI + halo, Wp), srg = ceildiv(Wp - 1 + ks * kKI + P.shift, Wp);
    const size_t need = 2 * (size_t)(P.ncb * srx + srg) * Wp * kRowBytes + kKI * kRowBytes + 256;
    if (need <= (size_t)kSmemLimit) {
      KS = ks; P.SRX = srx; P.SRG = srg; pl.smem = need;
      break;
    }
  }
  if (KS == 0) return false;
  P.KS = KS;
  P.xplane_bytes = P.SRX * Wp * kRowBytes;
  P.xstage_bytes = P.ncb * P.xplane_bytes;
  P.gstage_bytes = P.SRG * Wp * kRowBytes;
  const int64_t total_slots = (int64_t)g.N * Hp * Wp + P.shift;
  P.num_units = (int)((total_slots + (int64_t)KS * kKI - 1) / ((int64_t)KS * kKI));
  struct Win { int row, cb; uint32_t a16; };
  std::vector<Win> wins;
  for (int c = 0; c < P.ncb; ++c)
    for (int r = 0; r < g.fr; ++r)
      wins.push_back({r, c, (uint32_t)(((size_t)c * P.xplane_bytes + (size_t)(r * g.D * Wp + P.shift) * kRowBytes) >> 4)});
  wins.push_back({-1, 0, 0});
  if (wins.size() & 1) wins.push_back({-1, 0, 0});
  const int nmt = (int)wins.size() / 2;
  if (nmt > kMaxMT) return false;
  for (int i = 0; i < nmt; ++i) {
    const Win &a = wins[2 * i], &b = wins[2 * i + 1];
    P.row[i][0] = (int16_t)a.row; P.row[i][1] = (int16_t)b.row;
    P.cb[i][0] = (int16_t)a.cb; P.cb[i][1] = (int16_t)b.cb;
    P.a16[i] = a.a16;
    if (a.row >= 0 && b.row >= 0) { P.kind[i] = 0; P.lbo16[i] = b.a16 - a.a16; if (P.lbo16[i] >= (1u << 14)) return false; }
    else if (a.row >= 0) P.kind[i] = 1;
    else P.kind[i] = 2;
  }
  P.nmt_total = nmt;
  int cap = 512 / P.NN;
  if (cap > 4) cap = 4;
  P.n_mgroups = ceildiv(nmt, cap);
  P.mt_per_group = ceildiv(nmt, P.n_mgroups);
  // the ones window sorts last, so the only tile with run-time descriptor deltas is the last tile of the last group,
  // which is what ns_issue assumes
  int ctas = num_sms() / P.n_mgroups;
  if (ctas < 1) ctas = 1;
  if (ctas > P.num_units) ctas = P.num_units;
  pl.splits = ctas;
  pl.grid = dim3((unsigned)ctas, (unsigned)P.n_mgroups, 1);
  pl.part_bytes = align_up((size_t)ctas * P.Mrows * g.CO * sizeof(float), 256);
  pl.partdb_bytes = align_up((size_t)ctas * g.CO * sizeof(float), 256);
  return true;
}

int ns_wgrad(const GatherWgrad& g, NsPlan& pl, const void* in, const void* grad, float* dW, float* db, void* ws,
             size_t ws_bytes, cudaStream_t st) {
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.part_bytes + pl.partdb_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0,
             "activation pointers must be 16-byte aligned");
  NsParams& P = pl.P;
  {
    const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.N};
    const uint64_t strides[4] = {2, (uint64_t)g.CI * 2, (uint64_t)g.IW * g.CI * 2, (uint64_t)g.IH * g.IW * g.CI * 2};
    const uint32_t box[4] = {64, (uint32_t)P.Wp, 1, 1};
    if (int e = make_tmap(&P.xmap, true, const_cast<void*>(in), 4, dims, strides, box, false)) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g.CO, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
    const uint64_t strides[4] = {2, (uint64_t)g.CO * 2, (uint64_t)g.OW * g.CO * 2, (uint64_t)g.OH * g.OW * g.CO * 2};
    const uint32_t box[4] = {64, (uint32_t)P.Wp, 1, 1};
    if (int e = make_tmap(&P.gmap, true, const_cast<void*>(grad), 4, dims, strides, box, false)) return e;
  }
  float* part = reinterpret_cast<float*>(ws);
  float* part_db = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.part_bytes);
  auto kern = tc_slab_wgrad_ns_kernel;
  NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  P.red = fused_reduce(g, pl.grid, pl.splits, dW, db);
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, part, part_db);
  NPML_LAUNCH_OK();
  if (P.red.enabled) return 0;
  return launch_wgrad_reduce(part, pl.splits, P.Mrows, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate, dW, st,
                             db != nullptr ? part_db : nullptr, db);
}

}  // namespace

bool slab_wgrad_supported(const GatherWgrad& g, int mode) {
  SlabWgradPlan pl;
  return plan_slab_wgrad(g, mode, pl);
}

size_t slab_wgrad_ws_bytes(const GatherWgrad& g, int mode) {
  SlabWgradPlan pl;
  if (!plan_slab_wgrad(g, mode, pl)) return 0;
  size_t need = pl.part_bytes + pl.partdb_bytes;
  NsPlan ns;
  if (plan_ns_wgrad(g, mode, ns) && ns.part_bytes + ns.partdb_bytes > need) need = ns.part_bytes + ns.partdb_bytes;
  return need + 256;
}

// dW (+)= X (*) dZ and, when db != nullptr, db (+)= column sums of dZ, both from one pass over X and dZ.
int slab_wgrad(const GatherWgrad& g, int mode, const void* in, const void* grad, float* dW, float* db, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  {
    NsPlan ns;                 // out_ch <= 64: column-shifted dZ windows stacked along N
    if (plan_ns_wgrad(g, mode, ns)) return ns_wgrad(g, ns, in, grad, dW, db, ws, ws_bytes, st);
  }
  SlabWgradPlan pl;
  NPML_CHECK(plan_slab_wgrad(g, mode, pl), "shape not supported by the slab wgrad path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.part_bytes + pl.partdb_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0,
             "activation pointers must be 16-byte aligned");
  NPML_CHECK(db == nullptr || !g.no_db, "db was requested from a problem planned without the ones window");
  SlabWgradParams& P = pl.P;
  for (int q = 0; q < pl.nplanes; ++q) {
    // plane (py, px) of X as a strided view: rows py, py + s, ... and columns px, px + s, ...
    const int py = pl.plane_py[q], px = pl.plane_px[q];
    const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)((g.IW - px + g.s - 1) / g.s), (uint64_t)((g.IH - py + g.s - 1) / g.s),
                              (uint64_t)g.N};
    const uint64_t strides[4] = {2, (uint64_t)g.s * g.CI * 2, (uint64_t)g.s * g.IW * g.CI * 2, (uint64_t)g.IH * g.IW * g.CI * 2};
    const uint32_t box[4] = {64, (uint32_t)P.Wp, 1, 1};
    void* base = const_cast<char*>(reinterpret_cast<const char*>(in)) + ((int64_t)py * g.IW + px) * g.CI * 2;
    if (int e = make_tmap(&P.xmap[q], true, base, 4, dims, strides, box, false)) return e;
  }
  {
    const uint64_t dims[4] = {(uint64_t)g.CO, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)g.N};
    const uint64_t strides[4] = {2, (uint64_t)g.CO * 2, (uint64_t)g.OW * g.CO * 2, (uint64_t)g.OH * g.OW * g.CO * 2};
    const uint32_t box[4] = {64, (uint32_t)P.Wp, 1, 1};
    if (int e = make_tmap(&P.gmap, true, const_cast<void*>(grad), 4, dims, strides, box, false)) return e;
  }
  float* part = reinterpret_cast<float*>(ws);
  float* part_db = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.part_bytes);
  auto kern = tc_slab_wgrad_kernel;
  NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  P.red = fused_reduce(g, pl.grid, pl.splits, dW, db);
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, part, part_db);
  NPML_LAUNCH_OK();
  if (P.red.enabled) return 0;
  return launch_wgrad_reduce(part, pl.splits, P.Mrows, g.CO, g.CI, g.o_tap, g.o_ci, g.o_co, g.accumulate, dW, st,
                             db != nullptr ? part_db : nullptr, db);
}

}  // namespace npml
