// Persistent "slab" implicit-GEMM convolution for stride-1 gather convolutions on sm_100a
// (Conv2D.forward and Conv2D dX with stride 1 -- cfg2, and the three convolutions of cfg4).
//
// The activation tensor is viewed as ONE flat sequence of zero-padded pixels ("slots", 128 bytes of
// channels each): slot = (n*Hp + yp)*Wp + xp with Hp = OH + (fr-1)*D, Wp = OW + (fc-1)*D.  In that view
// the pixel a filter tap (r, t) multiplies with output slot s is simply slot s + r*D*Wp + t*D, so
//
//   * an M tile is 128 CONSECUTIVE output slots (the few slots with xp >= OW or yp >= OH are junk and
//     are masked in the epilogue), and
//   * the A operand of every tap is the SAME shared-memory slab read at a different start address:
//     the UMMA descriptor's start address is advanced by (r*D*Wp + t*D)*128 bytes.  The slab (whole padded
//     rows, one TMA box per row, zero padding = the TMA's out-of-bounds fill, 128B swizzle) is loaded
//     from L2 once per work unit instead of once per tap.
//
// One persistent CTA per SM walks work units of UT consecutive M tiles:
//   warp 0  slab producer (TMA, 2-stage ring of (unit, channel-block) planes)
//   warp 1  MMA issuer (tcgen05.mma M=128 N=BN K=32 B, fp32 accumulators in a ring of TMEM tile slots)
//   warp 2  weight producer (TMA; all taps resident in smem when they fit, else a streamed ring shared by
//           the UT tiles of the unit) and TMEM allocator
//   warps 4..7  epilogue (tcgen05.ld -> +bias -> act -> global), overlapped with the next tiles' MMAs
//   warps 8..11 second epilogue group: the tiles of a CTA alternate between the two groups.  Measured (ncu source
//               page of cfg2): with one group the MMA warp spent ~25 % of the kernel waiting for accumulator slots --
//               a tile's epilogue is one long dependent chain (tcgen05.ld -> math -> st.shared -> ld.shared ->
//               st.global) and four warps cannot hide it
#include <cstdlib>
#include <type_traits>

#include "npml_common.cuh"
#include "tc_ptx.cuh"

namespace npml {

int make_tmap(CUtensorMap* m, bool bf16, void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
              const uint32_t* box, bool atom32b);
int launch_repack_weights(const GatherConv& g, bool bf16, const float* w, void* wt, cudaStream_t st);

namespace {

constexpr int kMaxTaps = 64;
constexpr int kMaxCls = 9;
constexpr int kRowBytes = 128;
constexpr int kThreads = 384;      // warps 0..3 producers / MMA, 4..7 epilogue group 0, 8..11 epilogue group 1
constexpr int kSlabStages = 2;
constexpr int kMaxWStages = 8;
constexpr int kMaxAcc = 16;
constexpr int kSmemLimit = 227 * 1024;

struct SlabParams {
  CUtensorMap amap[kMaxCls], wmap;  // one activation map per input parity plane (a single one unless form 0, s > 1)
  uint32_t tap_off16[kMaxTaps];  // (offy*Wp + offx) slots * 128 B >> 4: descriptor start-address advance of the tap
  int16_t tap_w[kMaxTaps];     // index of the tap in Wt
  int ntaps, nkb, kelems;
  int N, OH, OW, CO, Hp, Wp, pady, padx;
  // Output parity classes (form 1 with stride s > 1: Deconv2D.forward, strided Conv2D dX).  Class c owns the output
  // pixels (y0 + s*i, x0 + s*j); on those the transposed convolution is an ordinary stride-1 gather of the input
  // with the class's own taps, so every class is a slab problem of its own on a common (Hp, Wp) slot grid and the
  // work units are numbered class-major.  ncls == 1 (so = 1, y0 = x0 = 0) is the plain stride-1 convolution.
  int ncls, units_per_cls, so, OHf, OWf;
  // Input parity planes (form 0 with stride s > 1: strided Conv2D.forward, Deconv2D dX).  Plane (py, px) is the
  // strided view in[:, py::s, px::s, :]; on it the taps with that parity are stride-1 taps, so the K loop simply
  // runs over (plane, channel block) pairs, each with its own tensor map and tap slice (the cls_* tables are then
  // indexed by plane).  nplanes == 1 otherwise.  nkg = nplanes * nkb K groups per unit.
  int nplanes, nkg;
  int16_t cls_tap0[kMaxCls], cls_ntaps[kMaxCls];       // the class's slice of tap_off16 / tap_w
  int16_t cls_pady[kMaxCls], cls_padx[kMaxCls], cls_y0[kMaxCls], cls_x0[kMaxCls], cls_oh[kMaxCls], cls_ow[kMaxCls];
  int UT, SR, BN, n_acc;
  int plane_bytes, w_block_bytes, w_resident, w_stages;
  int total_tiles, num_units;
  uint32_t idesc;
  int epi_direct;               // 1 (NPML_SLAB_EPI=direct, experiment): rows go straight from registers to global
                                // memory with 256-bit stores; measured equal to the staged path (the kernel is
                                // bound by the MMA operand fetch, not by the epilogue)
  int act;
  float p0, p1;
  int epi2;                     // 1: warps 8..11 take every other tile (their rows go straight to global memory)
  int wait_ns;                  // sleep between mbarrier polls of the producer / epilogue warps (npml::wait_ns())
  int affine;                   // 1: Y = scale[co] * act(z + b) + shift[co] (frozen BatchNorm2D folded into the epilogue)
};

// Measured on B200: the 128B swizzle of both the TMA write and the UMMA read is a function of the ABSOLUTE
// shared-memory address bits (chunk ^= (addr >> 7) & 7), so an operand start address that is only 128-byte
// aligned needs no correction: the descriptor's base_offset field stays 0 (setting it to (addr >> 7) & 7
// double-counts the phase and gives wrong results).
// K-major 128B-swizzle operand descriptor minus its start address: SBO = 1024 B, LBO field = 1, version 1,
// layout SWIZZLE_128B.  descriptor = kDesc64 + (shared address >> 4).
constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);

template <typename T, int NT, int UT>
__global__ void __launch_bounds__(kThreads, 1) tc_slab_conv_kernel(const __grid_constant__ SlabParams P,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ post_scale,
                                                                   const float* __restrict__ post_shift, T* __restrict__ Z,
                                                                   T* __restrict__ Y) {
  constexpr bool kTf32 = sizeof(T) == 4;
  // [weights (1024-aligned base)] [2 slab planes (128-aligned)] [epilogue staging] [bias | scale | shift] [barriers]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* w_area = smem_raw;
  const int w_bytes = (P.w_resident ? P.ntaps * P.nkb : P.w_stages) * P.w_block_bytes;
  uint8_t* smem = w_area + w_bytes;                                     // slab planes
  uint8_t* staging = smem + (size_t)kSlabStages * P.plane_bytes;         // 4 warps x 32 rows x (32 columns of T)
  float* bias_s = reinterpret_cast<float*>(staging + 4 * 32 * 32 * sizeof(T));
  float* scale_s = bias_s + 256;
  float* shift_s = scale_s + 256;
  uint64_t* slab_full = reinterpret_cast<uint64_t*>(bias_s + (P.affine ? 768 : 256));   // scale / shift only when used
  uint64_t* slab_empty = slab_full + kSlabStages;
  uint64_t* w_full = slab_empty + kSlabStages;
  uint64_t* w_empty = w_full + kMaxWStages;
  uint64_t* acc_full = w_empty + kMaxWStages;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kMaxAcc);
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) asm volatile("trap;");   // the weight tiles need the 1024-byte base

  // warp index broadcast from lane 0: tells ptxas it is warp-uniform, so the role branches are uniform branches and
  // everything the MMA issuer computes stays on the uniform datapath (no R2UR per descriptor)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * P.BN;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kSlabStages; ++i) { ptx::mbar_init(&slab_full[i], 1); ptx::mbar_init(&slab_empty[i], 1); }
    for (int i = 0; i < kMaxWStages; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < kMaxAcc; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < 256; i += kThreads - 128) {    // BN <= 256
      const bool ok = i < P.BN && co0 + i < P.CO;
      bias_s[i] = (bias != nullptr && ok) ? bias[co0 + i] : 0.f;
      if (P.affine) {
        scale_s[i] = ok ? post_scale[co0 + i] : 1.f;
        shift_s[i] = ok ? post_shift[co0 + i] : 0.f;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== slab producer: whole padded rows [rho0, rho0 + SR) of channel block kb, one TMA box per row =====
    if (ptx::elect_one()) {
      for (int i = 0; i < P.nplanes; ++i) ptx::prefetch_tmap(&P.amap[i]);
      const uint32_t row_bytes = (uint32_t)P.Wp * kRowBytes;
      uint32_t p = 0;
      for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x) {
        const int cls = P.ncls > 1 ? unit / P.units_per_cls : 0;
        const int rho0 = ((unit - cls * P.units_per_cls) * UT * 128) / P.Wp;
        for (int kg = 0; kg < P.nkg; ++kg, ++p) {
          const int plane = P.nplanes > 1 ? kg / P.nkb : 0;
          const int kb = kg - plane * P.nkb;
          const int tg = P.nplanes > 1 ? plane : cls;
          const int padx = P.cls_padx[tg], pady = P.cls_pady[tg];
          const int st = p & 1;
          ptx::mbar_wait_sleep(&slab_empty[st], ((p >> 1) & 1u) ^ 1u, (uint32_t)P.wait_ns);
          ptx::mbar_expect_tx(&slab_full[st], (uint32_t)P.SR * row_bytes);
          uint8_t* dst = smem + (size_t)st * P.plane_bytes;
          int n = rho0 / P.Hp, yp = rho0 - n * P.Hp;
          for (int i = 0; i < P.SR; ++i) {
            ptx::tma_load_4d(dst + (size_t)i * row_bytes, &P.amap[plane], &slab_full[st], kb * P.kelems, -padx, yp - pady, n);
            if (++yp == P.Hp) { yp = 0; ++n; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer =====
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&P.wmap);
      if (P.w_resident) {
        ptx::mbar_expect_tx(&w_full[0], (uint32_t)w_bytes);
        for (int t = 0; t < P.ntaps; ++t)
          for (int kb = 0; kb < P.nkb; ++kb)
            ptx::tma_load_3d(w_area + (size_t)(t * P.nkb + kb) * P.w_block_bytes, &P.wmap, &w_full[0], kb * P.kelems, co0,
                             P.tap_w[t]);
      } else {
        uint32_t ws = 0, wph = 0;
        for (int unit = blockIdx.x; unit < P.num_units; unit += gridDim.x) {
          const int cls = P.ncls > 1 ? unit / P.units_per_cls : 0;
          for (int kg = 0; kg < P.nkg; ++kg) {
            const int plane = P.nplanes > 1 ? kg / P.nkb : 0;
            const int kb = kg - plane * P.nkb;
            const int tg = P.nplanes > 1 ? plane : cls;
            const int tap0 = P.cls_tap0[tg], nt = P.cls_ntaps[tg];
            for (int t = 0; t < nt; ++t) {
              ptx::mbar_wait_sleep(&w_empty[ws], wph ^ 1u, (uint32_t)P.wait_ns >> 2);
              ptx::mbar_expect_tx(&w_full[ws], (uint32_t)P.w_block_bytes);
              ptx::tma_load_3d(w_area + (size_t)ws * P.w_block_bytes, &P.wmap, &w_full[ws], kb * P.kelems, co0,
                               P.tap_w[tap0 + t]);
              if (++ws == (uint32_t)P.w_stages) { ws = 0; wph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The whole warp walks the loops in CONVERGENT control flow (so ptxas keeps every descriptor in uniform
    // registers and advances it with one UIADD3.64 per MMA); only the tcgen05.mma / commit instructions sit
    // under the elected-lane branch.  NT (taps) and UT (tiles per unit) are compile-time so the tap loop is
    // straight-line code with immediate constant-bank offsets.  Ring positions are counters (no division).
    const bool leader = ptx::elect_one();
    const uint32_t BN = (uint32_t)P.BN, n_acc = (uint32_t)P.n_acc, idesc = P.idesc;
    const int nkb = P.nkb, nkg = P.nkg, Wp = P.Wp, total_tiles = P.total_tiles, num_units = P.num_units;
    const bool w_res = P.w_resident != 0;
    const uint32_t w_stages = (uint32_t)P.w_stages, wbb16 = (uint32_t)P.w_block_bytes >> 4;
    const uint32_t plane16 = (uint32_t)P.plane_bytes >> 4;
    const uint64_t slab_d = kDesc64 + (ptx::smem_u32(smem) >> 4);
    const uint64_t w_d = kDesc64 + (ptx::smem_u32(w_area) >> 4);
    uint32_t p = 0;               // slab ring sequence number
    uint32_t ws = 0, wph = 0;     // weight ring position / phase
    uint32_t slot = 0, sph = 0;   // accumulator ring position / phase of the unit's first tile
    if (w_res) {
      ptx::mbar_wait(&w_full[0], 0);
      ptx::tc_fence_after();
    }
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      // class of the unit (NT > 0 instantiations are single-class with compile-time taps)
      const int cls = (NT == 0 && P.ncls > 1) ? unit / P.units_per_cls : 0;
      const int lu = unit - cls * P.units_per_cls;          // unit index inside the class (total_tiles is per class)
      const int s0 = lu * UT * 128;
      const int ntile = min(UT, total_tiles - lu * UT);
      const uint32_t a_off16 = (uint32_t)(s0 % Wp) * (kRowBytes >> 4);
      uint32_t sl[UT], ph[UT], dt[UT];
#pragma unroll
      for (int j = 0; j < UT; ++j) {
        sl[j] = slot; ph[j] = sph; dt[j] = tmem_base + slot * BN;
        if (j < ntile && ++slot == n_acc) { slot = 0; sph ^= 1u; }
      }
      auto run_unit = [&](auto full_tag, auto res_tag) {
        constexpr bool kFull = decltype(full_tag)::value;     // every tile of the unit exists
        constexpr bool kRes = decltype(res_tag)::value;       // weights resident in smem (no ring waits)
        for (int kg = 0; kg < nkg; ++kg, ++p) {
          // tap slice of this K group: the unit's class (form 1), the group's input plane (form 0, s > 1), or all
          // NT compile-time taps
          const int plane = (NT == 0 && P.nplanes > 1) ? kg / nkb : 0;
          const int kb = kg - plane * nkb;
          const int tg = (NT == 0 && P.nplanes > 1) ? plane : cls;
          const int ntaps = NT > 0 ? NT : P.cls_ntaps[tg];
          const int tap0 = NT > 0 ? 0 : P.cls_tap0[tg];
          const uint32_t st = p & 1u;
          ptx::mbar_wait(&slab_full[st], (p >> 1) & 1u);
          ptx::tc_fence_after();
          const uint64_t a_base = slab_d + (st * plane16 + a_off16);
#pragma unroll
          for (int t = 0; t < ntaps; ++t) {
            const uint64_t ad = a_base + P.tap_off16[tap0 + t];
            uint64_t bd;
            if (kRes) {
              bd = w_d + (uint32_t)((tap0 + t) * nkb + kb) * wbb16;
            } else {
              ptx::mbar_wait(&w_full[ws], wph);
              ptx::tc_fence_after();
              bd = w_d + ws * wbb16;
            }
            const bool first = (kg == 0 && t == 0), last = (kg == nkg - 1 && t == ntaps - 1);
#pragma unroll
            for (int j = 0; j < UT; ++j) {
              if (kFull || j < ntile) {
                if (first) {            // the accumulator slot of tile j is waited for right before its first MMA
                  ptx::mbar_wait(&acc_empty[sl[j]], ph[j] ^ 1u);
                  ptx::tc_fence_after();
                }
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    ptx::mma_ss<kTf32>(dt[j], ad + (uint64_t)(j * (128 * kRowBytes >> 4) + 2 * ks), bd + (uint64_t)(2 * ks), idesc,
                                       (first && ks == 0) ? 0u : 1u);
                  if (last) ptx::tc_commit(&acc_full[sl[j]]);
                }
              }
            }
            if (leader && !kRes) ptx::tc_commit(&w_empty[ws]);
            if (!kRes && ++ws == w_stages) { ws = 0; wph ^= 1u; }
          }
          if (leader) ptx::tc_commit(&slab_empty[st]);
        }
      };
      if (w_res) {
        if (ntile == UT) run_unit(std::true_type{}, std::true_type{});
        else run_unit(std::false_type{}, std::true_type{});
      } else {
        if (ntile == UT) run_unit(std::true_type{}, std::false_type{});
        else run_unit(std::false_type{}, std::false_type{});
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> (+bias, act) -> swizzled smem staging -> coalesced global stores =====
    // A thread owns one accumulator row (TMEM lane = output slot).  Writing rows straight from registers would
    // touch 32 different 128-byte lines per store instruction, so each 32-column chunk is transposed through a
    // per-warp staging tile (XOR-swizzled, conflict-free both ways) and written with kCPR lanes per pixel.
    constexpr int kRB = 32 * (int)sizeof(T);      // bytes of one row of a 32-column chunk (64 bf16 / 128 fp32)
    constexpr int kCPR = kRB / 16;                // 16-byte pieces per row
    constexpr int kRPI = 32 / kCPR;               // rows written per store instruction
    constexpr int kVec = 16 / (int)sizeof(T);     // elements per 16-byte piece
    const int qd = warp & 3;                      // TMEM lane quadrant of this warp
    const int grp = (warp - 4) >> 2;              // epilogue group: tiles alternate between the groups when P.epi2
    const uint32_t ngrp = P.epi2 ? 2u : 1u;
    const int m = qd * 32 + lane;                 // row of the M tile
    const int img_slots = P.Hp * P.Wp;
    const int BN = P.BN, CO = P.CO, Wp = P.Wp, n_acc = P.n_acc;
    uint8_t* stg = staging + (size_t)((warp - 4) & 3) * 32 * kRB;      // group 1 never stages
    const uint32_t stg_u32 = ptx::smem_u32(stg);
    const int my_swz = kCPR == 4 ? ((lane >> 1) & 3) : (lane & 7);
    const int rd_row0 = lane / kCPR, rd_cc = lane % kCPR;
    uint32_t slot = 0, sph = 0;

    auto emit = [&](T* __restrict__ out, const float (&v)[32], int my_pix, int co) {
      // registers -> staging (row = lane), 16-byte pieces XOR-swizzled
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
      // staging -> global: kCPR lanes cover one pixel's chunk, kRPI pixels per instruction
#pragma unroll
      for (int it = 0; it < kCPR; ++it) {
        const int r = it * kRPI + rd_row0;
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

    // alternative without the staging round trip: every thread writes its own row with 32-byte stores
    auto emit_direct = [&](T* __restrict__ out, const float (&v)[32], int my_pix, int co) {
      if (my_pix < 0) return;
      T* dst = out + (int64_t)my_pix * CO + co;
      constexpr int kPer = 32 / (int)sizeof(T);          // elements per 32-byte store
#pragma unroll
      for (int h = 0; h < 32 / kPer; ++h) {
        uint32_t w[8];
        if constexpr (sizeof(T) == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[h * 16 + 2 * i], v[h * 16 + 2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&a);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[h * 8 + i]);
        }
        if (co + h * kPer < CO)
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + h * kPer), "r"(w[0]), "r"(w[1]),
                       "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                       : "memory");
      }
    };
    const bool direct = P.epi_direct != 0 || grp == 1;
    uint32_t tseq = 0;                            // running tile number of this CTA

    for (int unit = (grp == 0 || P.epi2) ? (int)blockIdx.x : P.num_units; unit < P.num_units; unit += gridDim.x) {
      const int cls = P.ncls > 1 ? unit / P.units_per_cls : 0;
      const int lu = unit - cls * P.units_per_cls;
      const int ntile = min(UT, P.total_tiles - lu * UT);
      for (int j = 0; j < ntile; ++j, ++tseq) {
        if ((tseq & (ngrp - 1u)) != (uint32_t)grp) {          // the other group's tile: only the ring position moves
          if (++slot == (uint32_t)n_acc) { slot = 0; sph ^= 1u; }
          continue;
        }
        const int s = (lu * UT + j) * 128 + m;
        const int n = s / img_slots;
        const int rem = s - n * img_slots;
        const int yp = rem / Wp, xp = rem - yp * Wp;
        const bool row_ok = n < P.N && yp < P.cls_oh[cls] && xp < P.cls_ow[cls];
        // pixel of the FULL output tensor: (y0 + so*yp, x0 + so*xp)
        const int my_pix = row_ok ? (n * P.OHf + P.cls_y0[cls] + P.so * yp) * P.OWf + P.cls_x0[cls] + P.so * xp : -1;
        ptx::mbar_wait_sleep(&acc_full[slot], sph, (uint32_t)P.wait_ns);
        ptx::tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + slot * (uint32_t)BN;
        for (int c = 0; c < BN; c += 32) {
          uint32_t r[32];
          float v[32];
          ptx::tmem_ld32(t_addr + (uint32_t)c, r);
          ptx::tmem_ld_wait();
          if (c + 32 >= BN) {                      // accumulator fully read: hand the TMEM slot back early
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_empty[slot]);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + c + i);
            v[i] = __uint_as_float(r[i]) + bb.x; v[i + 1] = __uint_as_float(r[i + 1]) + bb.y;
            v[i + 2] = __uint_as_float(r[i + 2]) + bb.z; v[i + 3] = __uint_as_float(r[i + 3]) + bb.w;
          }
          const int co = co0 + c;
          if (Z) {
            if (direct) emit_direct(Z, v, my_pix, co);
            else emit(Z, v, my_pix, co);
          }
          if (Y) {
            act_apply32(P.act, P.p0, P.p1, v);
            if (P.affine) {                         // frozen BatchNorm2D: per-channel scale / shift after the activation
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 sc = *reinterpret_cast<const float4*>(scale_s + c + i);
                const float4 sh = *reinterpret_cast<const float4*>(shift_s + c + i);
                v[i] = fmaf(v[i], sc.x, sh.x); v[i + 1] = fmaf(v[i + 1], sc.y, sh.y);
                v[i + 2] = fmaf(v[i + 2], sc.z, sh.z); v[i + 3] = fmaf(v[i + 3], sc.w, sh.w);
              }
            }
            if (direct) emit_direct(Y, v, my_pix, co);
            else emit(Y, v, my_pix, co);
          }
        }
        if (++slot == (uint32_t)n_acc) { slot = 0; sph ^= 1u; }
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

struct SlabPlan {
  SlabParams P;
  dim3 grid;
  size_t smem = 0, wt_bytes = 0;
  int plane_py[kMaxCls], plane_px[kMaxCls];      // input parity of each plane (form 0, s > 1)
};

inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

inline int posmod(int a, int b) { return ((a % b) + b) % b; }

bool plan_slab(const GatherConv& g, int mode, SlabPlan& pl) {
  if (mode != NPML_MODE_BF16 && mode != NPML_MODE_TF32) return false;
  if (std::getenv("NPML_DISABLE_SLAB")) return false;
  const bool bf16 = mode == NPML_MODE_BF16;
  const int es = bf16 ? 2 : 4, vec = 16 / es;
  if (g.N < 1 || g.s < 1) return false;
  // stride 1 (both forms); stride 2 / 3: the transposed form split into output parity classes, the forward form
  // split into input parity planes
  const bool classes = g.s > 1 && g.form == 1, planes = g.s > 1 && g.form == 0;
  if (g.s > 3) return false;
  if (classes && std::getenv("NPML_DISABLE_SLAB_CLASSES")) return false;
  if (planes && std::getenv("NPML_DISABLE_SLAB_PLANES")) return false;
  if (g.CI % vec || g.CO % vec) return false;
  if (g.fr * g.fc > kMaxTaps) return false;
  SlabParams& P = pl.P;
  memset(&P, 0, sizeof(P));
  P.N = g.N; P.OH = g.OH; P.OW = g.OW; P.CO = g.CO;
  P.OHf = g.OH; P.OWf = g.OW; P.so = classes ? g.s : 1;
  // per class: the taps that reach it, as (dy, dx) offsets into the UNSTRIDED input
  struct TapOff { int dy, dx, w; };
  TapOff taps[kMaxTaps];
  int ncls = 0, nt = 0, span_y = 0, span_x = 0, max_oh = 0, max_ow = 0;
  int dymin[kMaxCls], dxmin[kMaxCls];
  for (int qy = 0; qy < g.s; ++qy)
    for (int qx = 0; qx < g.s; ++qx) {
      int y0 = 0, x0 = 0, oh = g.OH, ow = g.OW;
      if (classes) {
        y0 = posmod(qy - g.pr, g.s); x0 = posmod(qx - g.pc, g.s);
        if (y0 >= g.OH || x0 >= g.OW) continue;
        oh = ceildiv(g.OH - y0, g.s); ow = ceildiv(g.OW - x0, g.s);
      }
      if (planes && (qy >= g.IH || qx >= g.IW)) continue;      // empty parity plane
      const int c = ncls;                                       // group index: class (form 1) or plane (form 0)
      int lo_y = 1 << 30, hi_y = -(1 << 30), lo_x = 1 << 30, hi_x = -(1 << 30);
      P.cls_tap0[c] = (int16_t)nt;
      for (int r = 0; r < g.fr; ++r)
        for (int t = 0; t < g.fc; ++t) {
          int dy, dx;
          if (g.form == 0) {                       // in[y*s + r*D - pr]: row (y + dy) of the plane with parity py
            const int offy = r * g.D - g.pr, offx = t * g.D - g.pc;
            if (planes && (posmod(offy, g.s) != qy || posmod(offx, g.s) != qx)) continue;
            dy = floordiv(offy, g.s); dx = floordiv(offx, g.s);
          } else {                                 // in[(y + pr - r*D) / s], only the taps whose parity matches
            if (posmod(r * g.D, g.s) != qy || posmod(t * g.D, g.s) != qx) continue;
            dy = (y0 + g.pr - r * g.D) / g.s; dx = (x0 + g.pc - t * g.D) / g.s;      // exact by construction
          }
          taps[nt++] = TapOff{dy, dx, r * g.fc + t};
          lo_y = dy < lo_y ? dy : lo_y; hi_y = dy > hi_y ? dy : hi_y;
          lo_x = dx < lo_x ? dx : lo_x; hi_x = dx > hi_x ? dx : hi_x;
        }
      P.cls_ntaps[c] = (int16_t)(nt - P.cls_tap0[c]);
      if (P.cls_ntaps[c] == 0) {
        if (planes) continue;                      // a plane no tap reads
        return false;                              // a class no tap reaches (bias only): left to the per-tile kernel
      }
      pl.plane_py[c] = qy; pl.plane_px[c] = qx;
      dymin[c] = lo_y; dxmin[c] = lo_x;
      span_y = hi_y - lo_y > span_y ? hi_y - lo_y : span_y;
      span_x = hi_x - lo_x > span_x ? hi_x - lo_x : span_x;
      max_oh = oh > max_oh ? oh : max_oh; max_ow = ow > max_ow ? ow : max_ow;
      P.cls_y0[c] = (int16_t)y0; P.cls_x0[c] = (int16_t)x0; P.cls_oh[c] = (int16_t)oh; P.cls_ow[c] = (int16_t)ow;
      P.cls_pady[c] = (int16_t)(-lo_y); P.cls_padx[c] = (int16_t)(-lo_x);
      ++ncls;
      if (g.s == 1) { qy = g.s; break; }
    }
  if (ncls == 0) return false;
  const int ngroups = ncls;
  if (planes) ncls = 1;
  const int Hp = max_oh + span_y, Wp = max_ow + span_x;
  if (Wp > 256) return false;
  // the flat slot grid computes Hp*Wp slots per max_oh*max_ow outputs: not worth it when the taps span much more
  // than the outputs (small strided planes); the per-tile kernel takes those.  Measured: cfg5 forward (17x17 slots
  // per 16x16 outputs) 61 -> 45 us on the slab, cfg3 dX (17x17 per 14x14) 32 -> 36 us.
  if ((classes || planes) && (double)Hp * Wp > 1.3 * max_oh * max_ow) return false;
  P.Hp = Hp; P.Wp = Wp; P.ncls = ncls; P.nplanes = planes ? ngroups : 1;
  const int64_t slots = (int64_t)g.N * Hp * Wp;
  if (slots + 128 * 8 >= (1LL << 31) / 2) return false;
  if ((int64_t)g.N * g.IH * g.IW * g.CI >= (1LL << 31) || (int64_t)g.N * g.OH * g.OW * g.CO >= (1LL << 31)) return false;
  P.pady = P.cls_pady[0]; P.padx = P.cls_padx[0];
  int halo = 0;
  for (int c = 0; c < ngroups; ++c)
    for (int i = P.cls_tap0[c]; i < P.cls_tap0[c] + P.cls_ntaps[c]; ++i) {
      const int off = (taps[i].dy - dymin[c]) * Wp + (taps[i].dx - dxmin[c]);
      P.tap_off16[i] = (uint32_t)off * (kRowBytes >> 4);
      P.tap_w[i] = (int16_t)taps[i].w;
      halo = off > halo ? off : halo;
    }
  P.ntaps = nt;
  P.kelems = kRowBytes / es;
  P.nkb = ceildiv(g.CI, P.kelems);
  P.nkg = P.nkb * P.nplanes;
  P.act = g.act; P.p0 = g.p0; P.p1 = g.p1;
  P.affine = g.post_scale != nullptr ? 1 : 0;
  P.wait_ns = wait_ns();
  int bn = (g.CO + 15) / 16 * 16;
  // wide outputs: one N = 192..256 tile (two accumulators) instead of two N = 128 tiles -- the A slab is fetched once
  // per MMA for twice the math, and the CTAs of the second N tile do not reload the same slabs
  if (bn > 256) bn = (g.CO % 256 == 0 || g.CO > 512) ? 256 : 128;
  if (std::getenv("NPML_SLAB_BN128") && bn > 128) bn = 128;
  if (bn < 32) bn = 32;
  P.BN = bn;
  P.n_acc = 512 / bn;
  if (P.n_acc > kMaxAcc) P.n_acc = kMaxAcc;
  P.idesc = ptx::instr_desc(bf16 ? 1 : 2, 0, 0, 128, bn);
  P.w_block_bytes = bn * kRowBytes;
  const int total_tiles = (int)((slots + 127) / 128);         // per class
  P.total_tiles = total_tiles;
  const int fixed = 4 * 32 * 32 * es /*epilogue staging*/ + (P.affine ? 3072 : 1024) /*bias (, scale, shift)*/ + 512 /*barriers*/;
  const int w_res_bytes = nt * P.nkb * P.w_block_bytes;
  bool found = false;
  // weights resident first (no re-streaming), largest unit that fits; then streamed weights
  for (int resident = 1; resident >= 0 && !found; --resident) {
    for (int ut = 4; ut >= (resident ? 2 : 1) && !found; --ut) {
      if (ut * bn > 512) continue;
      const int sr = (Wp - 1 + 128 * ut + halo - 1) / Wp + 1;
      const int plane = sr * Wp * kRowBytes;   // 128-byte aligned is enough (absolute-address swizzle)
      int wst = 4;
      int wbytes = resident ? w_res_bytes : wst * P.w_block_bytes;
      if (!resident && fixed + kSlabStages * plane + wbytes > kSmemLimit) { wst = 3; wbytes = wst * P.w_block_bytes; }
      if (fixed + kSlabStages * plane + wbytes > kSmemLimit) continue;
      P.UT = ut; P.SR = sr; P.plane_bytes = plane; P.w_resident = resident; P.w_stages = wst;
      pl.smem = (size_t)fixed + (size_t)kSlabStages * plane + wbytes;
      found = true;
    }
  }
  if (!found) return false;
  P.units_per_cls = ceildiv(total_tiles, P.UT);
  P.num_units = P.units_per_cls * ncls;
  {
    const char* e = std::getenv("NPML_SLAB_EPI");
    P.epi_direct = (e && e[0] == 'd' && (g.CO * es) % 32 == 0) ? 1 : 0;
    // second epilogue group (32-byte row stores straight from registers: needs 32-byte aligned rows)
    P.epi2 = ((g.CO * es) % 32 == 0 && !std::getenv("NPML_SLAB_EPI1")) ? 1 : 0;
  }
  const int nty = ceildiv(g.CO, bn);
  int ctas = num_sms() / nty;
  if (ctas < 1) ctas = 1;
  if (ctas > P.num_units) ctas = P.num_units;
  pl.grid = dim3((unsigned)ctas, (unsigned)nty, 1);
  pl.wt_bytes = align_up((size_t)g.fr * g.fc * g.CO * g.CI * es, 256);
  return true;
}

template <typename T>
int launch_slab(const GatherConv& g, SlabPlan& pl, const void* in, const float* w, const float* bias, void* Z, void* Y,
                void* ws, cudaStream_t st) {
  const bool bf16 = sizeof(T) == 2;
  const int es = sizeof(T);
  SlabParams& P = pl.P;
  if (int e = launch_repack_weights(g, bf16, w, ws, st)) return e;
  {
    const uint64_t dims[3] = {(uint64_t)g.CI, (uint64_t)g.CO, (uint64_t)(g.fr * g.fc)};
    const uint64_t strides[3] = {(uint64_t)es, (uint64_t)g.CI * es, (uint64_t)g.CI * g.CO * es};
    const uint32_t box[3] = {(uint32_t)P.kelems, (uint32_t)P.BN, 1};
    if (int e = make_tmap(&P.wmap, bf16, ws, 3, dims, strides, box, false)) return e;
  }
  {
    const uint32_t box[4] = {(uint32_t)P.kelems, (uint32_t)P.Wp, 1, 1};
    const int sp = P.nplanes > 1 ? g.s : 1;           // plane (py, px) = in[:, py::sp, px::sp, :]
    for (int i = 0; i < P.nplanes; ++i) {
      const int py = sp > 1 ? pl.plane_py[i] : 0, px = sp > 1 ? pl.plane_px[i] : 0;
      const uint64_t dims[4] = {(uint64_t)g.CI, (uint64_t)ceildiv(g.IW - px, sp), (uint64_t)ceildiv(g.IH - py, sp),
                                (uint64_t)g.N};
      const uint64_t strides[4] = {(uint64_t)es, (uint64_t)sp * g.CI * es, (uint64_t)sp * g.IW * g.CI * es,
                                   (uint64_t)g.IH * g.IW * g.CI * es};
      char* base = (char*)const_cast<void*>(in) + ((size_t)py * g.IW + px) * g.CI * es;
      if (int e = make_tmap(&P.amap[i], bf16, base, 4, dims, strides, box, false)) return e;
    }
  }
  typedef void (*Kern)(const SlabParams, const float*, const float*, const float*, T*, T*);
  Kern kern = nullptr;
#define NPML_SLAB_UT(NTV)                                                          \
  switch (P.UT) {                                                                  \
    case 1: kern = tc_slab_conv_kernel<T, NTV, 1>; break;                          \
    case 2: kern = tc_slab_conv_kernel<T, NTV, 2>; break;                          \
    case 3: kern = tc_slab_conv_kernel<T, NTV, 3>; break;                          \
    default: kern = tc_slab_conv_kernel<T, NTV, 4>; break;                         \
  }
  if (P.ncls == 1 && P.nplanes == 1 && P.ntaps == 9) { NPML_SLAB_UT(9) }
  else if (P.ncls == 1 && P.nplanes == 1 && P.ntaps == 1) { NPML_SLAB_UT(1) }
  else { NPML_SLAB_UT(0) }
#undef NPML_SLAB_UT
  NPML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  kern<<<pl.grid, kThreads, pl.smem, st>>>(P, bias, g.post_scale, g.post_shift, (T*)Z, (T*)Y);
  NPML_LAUNCH_OK();
  return 0;
}

}  // namespace

bool slab_conv_supported(const GatherConv& g, int mode) {
  SlabPlan pl;
  return plan_slab(g, mode, pl) || slab3_conv_supported(g, mode);
}

size_t slab_conv_ws_bytes(const GatherConv& g, int mode) {
  if (!slab_conv_supported(g, mode)) return 0;
  const int es = mode == NPML_MODE_BF16 ? 2 : 4;       // the repacked weights Wt[tap][co][ci], both slab kernels
  return align_up((size_t)g.fr * g.fc * g.CO * g.CI * es, 256) + 256;
}

int slab_gather_conv(const GatherConv& g, int mode, const void* in, const float* w, const float* bias, void* Z, void* Y,
                     void* ws, size_t ws_bytes, cudaStream_t st) {
  // narrow outputs (fc * out_ch <= 256): one MMA per filter ROW, column taps folded into N (tc_slab3_conv.cu)
  if (slab3_conv_supported(g, mode)) return slab3_gather_conv(g, mode, in, w, bias, Z, Y, ws, ws_bytes, st);
  SlabPlan pl;
  NPML_CHECK(plan_slab(g, mode, pl), "shape not supported by the slab path");
  NPML_CHECK(ws != nullptr && ws_bytes >= pl.wt_bytes, "workspace too small");
  NPML_CHECK((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 127) == 0,
             "activation / workspace pointers must be 16 / 128 byte aligned");
  // rows stored straight from registers use 32-byte stores: outputs that are not 32-byte aligned (a caller's own view)
  // keep the staged 16-byte path
  if (((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(Y)) & 31) != 0) { pl.P.epi2 = 0; pl.P.epi_direct = 0; }
  NPML_CHECK(((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0, "Z / Y must be 16-byte aligned");
  if (mode == NPML_MODE_BF16) return launch_slab<__nv_bfloat16>(g, pl, in, w, bias, Z, Y, ws, st);
  return launch_slab<float>(g, pl, in, w, bias, Z, Y, ws, st);
}

}  // namespace npml
