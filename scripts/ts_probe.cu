// Probe: layout of the A operand of tcgen05.mma when it is read from TENSOR MEMORY (kind::f16, bf16, M = 128).
//   A is written with tcgen05.st.32x32b.x32 (thread = TMEM lane = row m, register j = column j) as
//   reg j = bf16(A[m][2j]) | bf16(A[m][2j+1]) << 16, B (N rows, K-major, 128B swizzle) is one-hot: B[n][k] = (k == n % 64)
//   * (1 + n / 64), so D[m][n] = (1 + n / 64) * A[m'][k'] names the element of A the hardware paired with B's k = n % 64.
//   With A[m][k] = 64 * (m % 4) + k + 1 (exact in bf16) the mapping can be read off D.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ts_probe scripts/ts_probe.cu
//   build/ts_probe <N>        (N = 64, 128, 136, 144, 256)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../numpy-ml_b200/csrc/tc_ptx.cuh"

using namespace npml;

constexpr uint64_t kDesc64 = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32) | (1u << 16);

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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

__device__ __forceinline__ uint32_t bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return (uint32_t)*reinterpret_cast<uint16_t*>(&h);
}

__global__ void __launch_bounds__(128, 1) ts_probe_kernel(int N, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];       // B: 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B[n][k] = (k == n % 64) * (1 + n / 64), K-major rows of 64 bf16, 16-byte chunk c of row n stored at chunk c ^ (n & 7)
  for (int i = threadIdx.x; i < 256 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    const float v = (k == (n & 63)) ? (float)(1 + (n >> 6)) : 0.f;
    const int chunk = (k >> 3) ^ (n & 7);
    reinterpret_cast<uint16_t*>(smem)[n * 64 + chunk * 8 + (k & 7)] = (uint16_t)bf16_bits(v);
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  ptx::fence_proxy_async();
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_col = 256;                              // A: columns [256, 288), D: columns [0, N)
  {
    const int m = warp * 32 + lane;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float lo = (float)(64 * (m & 3) + 2 * j + 1), hi = (float)(64 * (m & 3) + 2 * j + 2);
      v[j] = bf16_bits(lo) | (bf16_bits(hi) << 16);
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + a_col, v);
    tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1 && ptx::elect_one()) {
    const uint32_t idesc = ptx::instr_desc(1, 0, 0, 128, N);
    const uint64_t bd = kDesc64 + (ptx::smem_u32(smem) >> 4);
    for (int ks = 0; ks < 4; ++ks) mma_ts(tmem, tmem + a_col + 8 * ks, bd + 2 * ks, idesc, ks ? 1u : 0u);
    ptx::tc_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    ptx::tmem_ld_wait();
    const int m = warp * 32 + lane;
    for (int j = 0; j < 32; ++j)
      if (c + j < N) out[m * 256 + c + j] = __uint_as_float(r[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 64;
  float* d_out;
  cudaMalloc(&d_out, sizeof(float) * 128 * 256);
  cudaMemset(d_out, 0xff, sizeof(float) * 128 * 256);
  cudaFuncSetAttribute(ts_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 128);
  ts_probe_kernel<<<1, 128, 256 * 128>>>(N, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
  std::vector<float> h(128 * 256);
  cudaMemcpy(h.data(), d_out, sizeof(float) * 128 * 256, cudaMemcpyDeviceToHost);
  // expected under the natural layout: D[m][n] = (1 + n / 64) * (64 * (m % 4) + n % 64 + 1)
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      const float want = (float)((1 + n / 64) * (64 * (m & 3) + (n & 63) + 1));
      if (h[m * 256 + n] != want) ++bad;
    }
  printf("N=%d: %d of %d entries differ from the natural layout (lane = row, column j = K elements 2j, 2j+1)\n", N, bad, 128 * N);
  if (bad) {
    for (int m : {0, 1, 2, 3, 4, 37, 64, 127}) {
      printf("  m=%3d:", m);
      for (int n = 0; n < (N < 72 ? N : 72); ++n) printf(" %g", h[m * 256 + n]);
      printf("\n");
    }
  }
  return 0;
}
