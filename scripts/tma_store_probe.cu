// Probe for the round-2 epilogue (finding 8): a TMA tensor STORE of a 128B-swizzled shared-memory
// tile whose box hangs over the edge of the output tensor.
//   * the tile holds W consecutive padded-row slots, one slot = 64 bf16 channels = 128 bytes, 16-byte chunk c of slot p
//     stored at chunk (c ^ (p_abs & 7)) where p_abs = (shared address of the slot) >> 7  (the absolute-address rule
//     measured for TMA loads and UMMA reads, DESIGN.md finding 1), the tile starting at an arbitrary 128-byte offset;
//   * tensor map (C = 64, OW, OH, N) bf16, box (64, W, 1, 1), SWIZZLE_128B; store at (0, x0, y, n) with W > OW - x0 and
//     also at a row y >= OH: columns / rows outside the tensor must be dropped.
// Expected: out[n][y][x][c] = value(p = x - x0, c) for the in-bounds part, everything else untouched (-1).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_store_probe scripts/tma_store_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>

#include "../numpy-ml_b200/csrc/tc_ptx.cuh"

using namespace npml;

__global__ void __launch_bounds__(128, 1) store_probe_kernel(const __grid_constant__ CUtensorMap omap, int W, int slot_off,
                                                             int x0, int y, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tile = smem + (size_t)slot_off * 128;           // deliberately not 1024-byte aligned
  for (int i = threadIdx.x; i < W * 64; i += blockDim.x) {
    const int p = i >> 6, c = i & 63;
    const uint32_t slot_addr = ptx::smem_u32(tile) + (uint32_t)p * 128u;
    const uint32_t swz = (slot_addr >> 7) & 7u;
    // bf16 has 8 significant bits: the slot index p (<= 255) and the channel c (<= 63) are each exact, so rows with an odd
    // y carry c and rows with an even y carry p
    const __nv_bfloat16 h = __float2bfloat16_rn((y & 1) ? (float)c : (float)p);
    *reinterpret_cast<__nv_bfloat16*>(tile + (size_t)p * 128 + ((((uint32_t)c >> 3) ^ swz) << 4) + (c & 7) * 2) = h;
  }
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::tma_store_4d(&omap, tile, 0, x0, y, n);
    ptx::tma_store_commit();
    ptx::tma_store_wait0();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int C = 64, OW = 56, OH = 56, N = 2, W = 58;
  const size_t elems = (size_t)N * OH * OW * C;
  __nv_bfloat16* d_out;
  cudaMalloc(&d_out, elems * 2);
  int bad_total = 0;
  // (slot_off, x0, y, n): aligned / misaligned tile start, box hanging over the right edge, a row outside the tensor
  const int cases[][4] = {{0, 0, 3, 0}, {3, 0, 4, 1}, {5, -1, 7, 0}, {2, 10, 8, 1}, {1, 0, 56, 0}, {7, 0, 57, 1}};
  for (auto& cs : cases) {
    cudaMemset(d_out, 0xff, elems * 2);                    // bf16 0xffff = NaN marker for "untouched"
    CUtensorMap omap;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)OW * C * 2, (cuuint64_t)OH * OW * C * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)W, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&omap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_out, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    const int slot_off = cs[0], x0 = cs[1], y = cs[2], n = cs[3];
    cudaFuncSetAttribute(store_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    store_probe_kernel<<<1, 128, 16384>>>(omap, W, slot_off, x0, y, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case off=%d x0=%d y=%d n=%d: %s\n", slot_off, x0, y, n, cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> h(elems);
    cudaMemcpy(h.data(), d_out, elems * 2, cudaMemcpyDeviceToHost);
    int bad = 0, written = 0;
    for (int nn = 0; nn < N; ++nn)
      for (int yy = 0; yy < OH; ++yy)
        for (int xx = 0; xx < OW; ++xx)
          for (int c = 0; c < C; ++c) {
            const uint16_t got = h[(((size_t)nn * OH + yy) * OW + xx) * C + c];
            const int p = xx - x0;
            const bool inside = nn == n && yy == y && p >= 0 && p < W;
            if (inside) {
              ++written;
              const float want = (y & 1) ? (float)c : (float)p;
              __nv_bfloat16 wb = __float2bfloat16_rn(want);
              if (got != *reinterpret_cast<uint16_t*>(&wb)) ++bad;
            } else if (got != 0xffff) {
              ++bad;
            }
          }
    printf("tile start slot %d, store at (x0=%d, y=%d, n=%d): %d elements expected, %d wrong\n", slot_off, x0, y, n, written, bad);
    bad_total += bad;
  }
  printf("%s\n", bad_total ? "MISMATCH" : "TMA store of the swizzled tile: clipping and absolute-address swizzle as assumed");
  return 0;
}
