// One grid-wide barrier per launch, for kernels whose blocks are ALL resident at once (the host side sizes the grid from
// cudaOccupancyMaxActiveBlocksPerMultiprocessor, or launches one block per SM).  Used to finish a fixed-order split
// reduction inside the kernel that produced the partials: every block writes its partial, the grid meets here, then
// every block adds its own slice of the outputs over all partials -- no second launch, no block that runs alone at
// the end, and the same order of additions run after run.
//
// The barrier word pair lives in zero-initialised __device__ memory (one GridBar per kernel that uses it): the last
// block to arrive resets the counter and bumps the generation, so consecutive launches on one stream need no memset.
// Two launches that share a GridBar must not overlap in time (the host layer issues them on one stream).
#pragma once
#include <cuda_runtime.h>

namespace npml {

struct GridBar {
  unsigned int count;
  unsigned int gen;
};

__device__ __forceinline__ void grid_barrier(GridBar* b, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                                              // this block's global writes, before the arrival
    const unsigned int g = *reinterpret_cast<volatile unsigned int*>(&b->gen);
    if (atomicAdd(&b->count, 1u) == nblocks - 1u) {
      b->count = 0u;
      __threadfence();
      atomicAdd(&b->gen, 1u);
    } else {
      // Bounded wait: the blocks of these kernels are resident by construction and a barrier takes microseconds; if
      // that assumption is ever wrong (a smaller device, a grid sized by hand) the launch must FAIL, not hang the GPU:
      // ~2 s of polling, then a trap (the host sees cudaErrorLaunchFailure on its next synchronising call).
      for (unsigned long long polls = 0; *reinterpret_cast<volatile unsigned int*>(&b->gen) == g; ++polls) {
        __nanosleep(polls < 4096 ? 32 : 256);       // short sleeps while a normal barrier would complete (wake-up latency)
        if (polls > (8ull << 20)) asm volatile("trap;");
      }
    }
    __threadfence();                                              // the other blocks' writes, after the release
  }
  __syncthreads();
}

}  // namespace npml
