// Micro-benchmark: cycles per tcgen05.mma (M = 128 per CTA, K = 16 bf16) on sm_100a as a function of
//   * where the A operand comes from: shared memory ("SS") or tensor memory ("TS"),
//   * the N extent (64 / 128 / 256),
//   * cta_group::1 vs cta_group::2 (CTA pair: M = 256, each CTA fetches its own A half and HALF of B).
// One elected thread issues `iters` MMAs back to back into one accumulator (4 K-steps of a 128-byte swizzled row,
// like the convolution kernels), commits, waits, and reports clock64() deltas.  Numbers are not checked: this only
// measures the issue/operand-fetch/math pacing that DESIGN.md section 4 (finding 2) reasons about.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_rate scripts/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "../numpy-ml_b200/csrc/tc_ptx.cuh"

namespace cg = cooperative_groups;
using namespace npml;

constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);

template <int CG>
__device__ __forceinline__ void mma_issue_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_issue_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc),
                 "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc),
                 "r"(acc)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     ptx::smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* slot) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ptx::smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ptx::smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t addr) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}

// CG: cta_group; TS: A from tensor memory; N: MMA N (whole pair for CG = 2); NACC: accumulators cycled through
template <int CG, int TS, int N, int NACC, int M = 128>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_s = smem;                        // 128 rows x 128 B, K-major, 128B swizzle
  uint8_t* b_s = smem + 16384;                // up to 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 2; i += blockDim.x)
    reinterpret_cast<uint16_t*>(smem)[i] = (uint16_t)(0x3c00 + ((i * 2654435761u) >> 25));   // bf16 around 0.01
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) tmem_alloc_cg<CG>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) cg::this_cluster().sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool leader_cta = CG == 1 || cg::this_cluster().block_rank() == 0;
  long long dt = 0;
  if (warp == 1) {
    const bool lane0 = ptx::elect_one();
    const uint32_t idesc = ptx::instr_desc(1, 0, 0, M * CG, N);
    const uint64_t ad = kDesc64 + (ptx::smem_u32(a_s) >> 4), bd = kDesc64 + (ptx::smem_u32(b_s) >> 4);
    const uint32_t a_t = tmem + (uint32_t)(NACC * N);      // A (4 K-steps x 8 columns) sits after the accumulators
    if (leader_cta && lane0) {
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + (uint32_t)((it % NACC) * N);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (TS) mma_issue_ts<CG>(d, a_t + 8 * ks, bd + 2 * ks, idesc, (it >= NACC || ks) ? 1u : 0u);
          else mma_issue_ss<CG>(d, ad + 2 * ks, bd + 2 * ks, idesc, (it >= NACC || ks) ? 1u : 0u);
        }
      }
      commit<CG>(&bar);
      ptx::mbar_wait(&bar, 0);
      dt = clock64() - t0;
      out[blockIdx.x] = dt;
    } else if (!leader_cta && lane0) {
      ptx::mbar_wait(&bar, 0);
      out[blockIdx.x] = 0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CG == 2) cg::this_cluster().sync();
  if (warp == 0) {
    ptx::tc_fence_after();
    tmem_dealloc_cg<CG>(tmem);
  }
}

template <int CG, int TS, int N, int NACC, int M = 128>
void run(const char* name, int grid, int iters, long long* d_out) {
  auto kern = mma_rate_kernel<CG, TS, N, NACC, M>;
  const size_t smem = 16384 + 32768;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {            // first repetition warms up
    cudaMemset(d_out, 0, sizeof(long long) * grid);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters, d_out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s grid %3d: %s\n", name, grid, cudaGetErrorString(e)); return; }
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60; int n = 0;
  for (long long v : h) if (v > 0) { mx = v > mx ? v : mx; mn = v < mn ? v : mn; ++n; }
  // math floor per MMA instruction: every SM computes 128 x N x 16 (with cta_group::2 the pair computes 256 x N x 16),
  // i.e. N / 2 cycles at 8192 flop/clk/SM in both cases
  const double per = (double)mx / (4.0 * iters);
  printf("%-28s grid %3d: %7.1f cycles/MMA (min CTA %7.1f), math floor %5.1f -> %4.1f %% of peak\n", name, grid, per,
         (double)mn / (4.0 * iters), N / 2.0, 100.0 * (N / 2.0) / per);
}

int main(int argc, char** argv) {
  // one variant per process (`mma_rate <index> <grid>`): a faulting variant must not poison the others
  const int which = argc > 1 ? atoi(argv[1]) : -1, grid = argc > 2 ? atoi(argv[2]) : 148;
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148);
  const int iters = 2048;
  int v = 0;
#define VARIANT(CG, TS, N, NACC, NAME) \
  if (which < 0 || which == v) run<CG, TS, N, NACC>(NAME, grid, iters, d_out); \
  ++v;
  VARIANT(1, 0, 64, 1, "SS cg1 N=64")
  VARIANT(1, 0, 128, 1, "SS cg1 N=128")
  VARIANT(1, 0, 256, 1, "SS cg1 N=256")
  VARIANT(1, 0, 128, 2, "SS cg1 N=128 2 acc")
  VARIANT(1, 1, 64, 1, "TS cg1 N=64")
  VARIANT(1, 1, 128, 1, "TS cg1 N=128")
  VARIANT(1, 1, 256, 1, "TS cg1 N=256")
  VARIANT(1, 1, 128, 2, "TS cg1 N=128 2 acc")
  VARIANT(2, 0, 64, 1, "SS cg2 N=64")
  VARIANT(2, 0, 128, 1, "SS cg2 N=128")
  VARIANT(2, 0, 256, 1, "SS cg2 N=256")
  VARIANT(2, 1, 128, 1, "TS cg2 N=128")
  VARIANT(2, 1, 256, 1, "TS cg2 N=256")
  // N extents of the weight-stationary convolution (tile of NT slots + 8 columns for the paired tap)
  VARIANT(1, 1, 72, 1, "TS cg1 N=72")
  VARIANT(1, 1, 80, 1, "TS cg1 N=80")
  VARIANT(1, 1, 96, 1, "TS cg1 N=96")
  VARIANT(1, 1, 104, 1, "TS cg1 N=104")
  VARIANT(1, 1, 112, 1, "TS cg1 N=112")
  VARIANT(1, 1, 120, 1, "TS cg1 N=120")
  VARIANT(1, 1, 136, 1, "TS cg1 N=136")
  VARIANT(1, 1, 144, 1, "TS cg1 N=144")
  VARIANT(1, 1, 160, 1, "TS cg1 N=160")
  VARIANT(1, 1, 192, 1, "TS cg1 N=192")
  VARIANT(1, 1, 200, 1, "TS cg1 N=200")
  VARIANT(1, 1, 136, 2, "TS cg1 N=136 2 acc")
  VARIANT(1, 0, 136, 1, "SS cg1 N=136")
  VARIANT(1, 0, 192, 1, "SS cg1 N=192")
  // M = 64: does a half-height MMA cost half the cycles?  (floor printed for M = 128)
  if (which < 0 || which == v) run<1, 1, 136, 1, 64>("TS cg1 M=64 N=136", grid, iters, d_out);
  ++v;
  if (which < 0 || which == v) run<1, 0, 128, 1, 64>("SS cg1 M=64 N=128", grid, iters, d_out);
  ++v;
  if (which < 0 || which == v) run<1, 1, 256, 1, 64>("TS cg1 M=64 N=256", grid, iters, d_out);
  ++v;
  if (which >= v) return 3;
  return 0;
}
