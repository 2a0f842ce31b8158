// Probe: register mapping of tcgen05.ld.16x256b.x8 and whether a warp may address the upper 16 lanes of its quadrant
// (lane base 32*w + 16) and odd column offsets.  TMEM is filled with value = lane * 256 + column (tcgen05.st.32x32b), then
// every warp reads 16 lanes x 64 columns four ways: (lane base +0 | +16) x (column 0 | 3).
// Assumed mapping (the m16n8 accumulator fragment, repeated along the columns): register 4j + i of thread t holds
//   row  t / 4 + 8 * (i / 2),   column  8 * j + 2 * (t % 4) + (i % 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ts_probe2 scripts/ts_probe2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../numpy-ml_b200/csrc/tc_ptx.cuh"

using namespace npml;

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
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe2_kernel(float* out, int lane_off, int col_off) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int m = warp * 32 + lane;
  for (int c = 0; c < 96; c += 32) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(m * 256 + c + j));
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  {
    const uint32_t lane_base = (uint32_t)(warp * 32 + lane_off), col = (uint32_t)col_off;
    uint32_t r[32];
    tmem_ld_16x256b_x8(tmem + (lane_base << 16) + col, r);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 32 + i] = __uint_as_float(r[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int main(int argc, char** argv) {
  const int lane_off = argc > 1 ? atoi(argv[1]) : 0, col_off = argc > 2 ? atoi(argv[2]) : 0;
  float* d_out;
  const int n = 4 * 32 * 32;
  cudaMalloc(&d_out, sizeof(float) * n);
  cudaMemset(d_out, 0xff, sizeof(float) * n);
  probe2_kernel<<<1, 128>>>(d_out, lane_off, col_off);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("16x256b.x8 lane base +%d, column +%d: %s\n", lane_off, col_off, cudaGetErrorString(e)); return 1; }
  std::vector<float> h(n);
  cudaMemcpy(h.data(), d_out, sizeof(float) * n, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int t = 0; t < 32; ++t)
      for (int i = 0; i < 32; ++i) {
        const int j = i / 4, k = i % 4;
        const int row = w * 32 + lane_off + t / 4 + 8 * (k / 2);
        const int col = col_off + 8 * j + 2 * (t % 4) + (k % 2);
        if (h[(w * 32 + t) * 32 + i] != (float)(row * 256 + col)) ++bad;
      }
  printf("16x256b.x8 lane base +%d, column +%d: %d of 4096 values differ from the assumed mapping\n", lane_off, col_off, bad);
  if (bad) {
    for (int t : {0, 1, 4, 5, 31}) {
      printf("  warp 1 thread %2d (lane,col):", t);
      for (int i = 0; i < 12; ++i) {
        const int v = (int)h[(32 + t) * 32 + i];
        printf(" (%d,%d)", v / 256, v % 256);
      }
      printf("\n");
    }
  }
  return 0;
}
