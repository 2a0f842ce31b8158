// The fixed-order split-K reduction of the wgrad kernels, run INSIDE the kernel that wrote the partials: after one grid
// barrier (grid_barrier.cuh; every CTA of these persistent kernels is resident, one per SM) CTA c adds chunks c,
// c + nctas, ... of 256 outputs over all `splits` partials -- thread (x, y) of the chunk adds the splits y, y + 4, ... of
// four consecutive outputs with 16-byte loads (sixteen in flight), the four group sums are added in a fixed order.  Same
// additions in the same order run after run, no second launch, and the partials are read back while they are still in L2.
// (The separate wgrad_reduce kernel: 11 us of a 75 us wgrad call at cfg2 -- launch, ramp, a single wave of blocks.)
#pragma once
#include "grid_barrier.cuh"

namespace npml {

struct FusedReduce {
  float* dW;
  float* db;                   // may be NULL
  int64_t o_tap, o_ci, o_co;   // strides of dW
  int accumulate;
  int splits;                  // partial slots per output (part[splits][M][CO], part_db[splits][CO])
  int enabled;
};

// All threads of the CTA (blockDim.x == 256) call this after the grid barrier.  red: 4 * 64 float4 of shared memory.
__device__ __forceinline__ void wgrad_reduce_slices(const FusedReduce& R, const float* __restrict__ part,
                                                    const float* __restrict__ part_db, int M, int CO, int CI, int cta,
                                                    int nctas, float4* red) {
  const int64_t total = (int64_t)M * CO;
  const int64_t total_all = total + (R.db != nullptr ? CO : 0);
  const int64_t nchunks = (total_all + 255) / 256;
  const int x = threadIdx.x & 63, y = threadIdx.x >> 6;
  for (int64_t ch = cta; ch < nchunks; ch += nctas) {
    const int64_t i = ch * 256 + 4 * x;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total_all) {
      const float* src = i < total ? part + i : part_db + (i - total);
      const int64_t stride = i < total ? total : (int64_t)CO;
      for (int z0 = y; z0 < R.splits; z0 += 64) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u)
          v[u] = (z0 + 4 * u < R.splits) ? __ldcg(reinterpret_cast<const float4*>(src + (int64_t)(z0 + 4 * u) * stride))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 16; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
      }
    }
    __syncthreads();
    red[y * 64 + x] = s;
    __syncthreads();
    {                                                        // thread t finishes output ch*256 + t
      const int64_t ie = ch * 256 + threadIdx.x;
      if (ie < total_all) {
        const int xe = threadIdx.x >> 2, e = threadIdx.x & 3;
        float t = 0.f;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) t += reinterpret_cast<const float*>(&red[yy * 64 + xe])[e];
        if (ie < total) {
          const int m = (int)(ie / CO), co = (int)(ie - (int64_t)m * CO);
          const int tap = m / CI, ci = m - tap * CI;
          const int64_t o = tap * R.o_tap + ci * R.o_ci + co * R.o_co;
          R.dW[o] = R.accumulate ? R.dW[o] + t : t;
        } else {
          const int co = (int)(ie - total);
          R.db[co] = R.accumulate ? R.db[co] + t : t;
        }
      }
    }
  }
}

}  // namespace npml
