// All-reduce (SUM) of the small gradient / statistics buffers of the data-parallel step over NVLink / NVSwitch PEER
// MEMORY, written as one kernel per call instead of a NCCL collective (dW || db of a layer is 1 KB - 2 MB,
// the sync-BN statistics 2 KB -- all latency-bound).
//
//   push   every rank writes each 32-bit word of its buffer, paired with the call number, as ONE 8-byte store straight
//          into an inbox in every peer's memory (posted NVLink stores; data and flag arrive together, so there is no
//          fence and no separate flag round trip -- the "LL" idea);
//   sum    every rank polls its OWN (local) inbox until the word of rank r carries this call's number and adds the
//          copies in RANK ORDER -- the same order on every rank, so all ranks hold bit-identical results, run after
//          run (the reference sums the batch in one fixed order too, layers/layers.py:3109-3110).
//
// A thread only ever waits for words that the peers push without waiting for anybody: no co-residency requirement, no
// grid barrier.  The call number ("epoch") lives in device memory and is advanced by the last block of a launch, so a
// captured CUDA graph replays correctly.  Calls alternate between two inbox halves, which is enough: a rank can finish
// call k+1 only after every peer has pushed call k+1, i.e. after that peer finished READING call k (stream order).
// Independent streams (the gradient side stream, the sync-BN statistics on the main stream) use different channels.
// The price of the flag is 2x the bytes on the wire; buffers whose one-shot wire bytes (world-1) * 2 * bytes pass 3 MB
// take the two-shot form below (owner-reduces, then broadcasts), and anything above 2.5 MB goes to ncclAllReduce
// (numpy-ml_b200/dist.py).  First version of this file used 16-byte stores, a
// __threadfence_system and one flag per block: 15 us for 144 KB on 2 GPUs against NCCL's 12 (profiles/r2k_mgpu.log).
//
// Memory: one cudaMalloc'ed region per rank, exported with cudaIpcGetMemHandle and opened by every peer at init time
// (host side: numpy-ml_b200/dist.py).  This is the one allocation the library makes itself -- at init, never on the
// hot path -- because an IPC handle must name a whole cudaMalloc allocation.
#include <cstring>

#include "npml_common.cuh"

namespace {

constexpr int kMaxRanks = 8;
constexpr int kChannels = 4;
constexpr int kThreads = 256;
constexpr size_t kMaxMsgBytes = (size_t)5 << 19;     // 2.5 MB per call (cfg5: 2.1 MB)
constexpr size_t kOneShotWire = (size_t)3 << 20;      // (world - 1) * 2 * bytes up to here: one-shot, beyond: two-shot
constexpr unsigned kSpinLimit = 1u << 23;             // seconds of polling: a peer that never arrives must not hang the box

// region layout (identical on every rank)
//   ctrl  [kChannels] { epoch, arrived, error }                       local bookkeeping
//   area  [kChannels][2] of kAreaBytes, written by peers as (word, epoch) pairs:
//         one-shot: [source rank][n] pairs;   two-shot: reduce-scatter inbox [source rank][slice] + all-gather inbox [n]
struct Ctrl { unsigned epoch, arrived, error, pad; };
constexpr size_t kAreaOff = 4096;
constexpr size_t kAreaBytes = 4 * kMaxMsgBytes + 4096;               // two-shot: 2 x (8-byte pair per 4-byte word) x 2 inboxes
constexpr size_t kRegionBytes = kAreaOff + (size_t)kChannels * 2 * kAreaBytes;

struct PeerTable {
  char* base[kMaxRanks];       // base[r]: rank r's region as mapped into THIS process
  int world, rank;
};

struct Peer {
  PeerTable t{};
  void* local = nullptr;
  void* opened[kMaxRanks] = {};
  bool ready = false;
};
Peer g_peer;

__device__ __forceinline__ uint2* area_ptr(char* base, int ch, int half) {
  return reinterpret_cast<uint2*>(base + kAreaOff + ((size_t)ch * 2 + half) * kAreaBytes);
}
__device__ __forceinline__ void st_pair(uint2* p, unsigned v, unsigned f) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(f) : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
  uint2 r;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
// wait until the pair carries this call's number; `dead` (a wait has timed out before): do not wait at all
__device__ __forceinline__ unsigned poll_word(const uint2* p, unsigned epoch, bool& dead) {
  uint2 v = ld_pair(p);
  unsigned spins = 0;
  while (v.y != epoch && !dead) {
    __nanosleep(40);                       // these threads share SMs with the persistent convolution kernels
    if (++spins > kSpinLimit) dead = true;
    v = ld_pair(p);
  }
  return v.x;
}
template <typename T, int W>
__device__ __forceinline__ T from_words(const unsigned (&w)[W]) {
  if (W == 1) return (T)__uint_as_float(w[0]);
  return (T)__hiloint2double((int)w[W - 1], (int)w[0]);
}

// mine + the copies of the other ranks at inbox[r * stride_r + idx], added in RANK order.  All world-1 loads are issued
// before the first one is looked at (one L2 round trip instead of world-1); only words that had not arrived are re-polled.
template <typename T, int W>
__device__ __forceinline__ T sum_in_rank_order(const uint2* inbox, int64_t stride_r, int64_t idx, T mine, int me, int world,
                                               unsigned epoch, bool& dead) {
  uint2 v[kMaxRanks][W];
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r)
    if (r < world && r != me) {
#pragma unroll
      for (int k = 0; k < W; ++k) v[r][k] = ld_pair(inbox + ((int64_t)r * stride_r + idx) * W + k);
    }
  T acc = (T)0;
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r) {
    if (r >= world) continue;
    if (r == me) { acc += mine; continue; }
    unsigned w[W];
#pragma unroll
    for (int k = 0; k < W; ++k)
      w[k] = v[r][k].y == epoch ? v[r][k].x : poll_word(inbox + ((int64_t)r * stride_r + idx) * W + k, epoch, dead);
    acc += from_words<T, W>(w);
  }
  return acc;
}

// W = 32-bit words per element (1: float, 2: double).  Element e of the buffer is handled by the same thread on every
// rank.  TWO_SHOT = false: every rank pushes everything to everybody and adds all copies itself.  TWO_SHOT = true
// (larger buffers): element e is OWNED by rank e / slice; the other ranks push their copy to the owner, the owner adds
// the copies in rank order and pushes the sum back to everybody -- 2 (world-1)/world of the one-shot wire bytes.
template <typename T, int W, bool TWO_SHOT>
__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(PeerTable tab, T* __restrict__ buf, int64_t n,
                                                                  int64_t slice, int channel) {
  char* mine = tab.base[tab.rank];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(mine) + channel;
  const unsigned epoch = ctrl->epoch + 1u;           // every block reads it before the last block advances it
  const int half = (int)(epoch & 1u);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int64_t e0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int world = tab.world, me = tab.rank;
  unsigned* words = reinterpret_cast<unsigned*>(buf);
  // once a wait has timed out (a peer that never made this call: a bug on the host side) nothing waits any more --
  // this call and all later ones rush through with garbage, npml_peer_status() reports it, nothing hangs
  bool dead = ctrl->error != 0u;
  uint2* my_area = area_ptr(mine, channel, half);
  if (!TWO_SHOT) {
    // ---- push: my words (with the call number) into every peer's inbox [me][e] ----
    for (int64_t e = e0; e < n; e += stride) {
      unsigned w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = words[e * W + k];
      for (int p = 1; p < world; ++p) {
        const int dst = (me + p) % world;             // staggered so that the ranks do not all start on the same peer
        uint2* box = area_ptr(tab.base[dst], channel, half) + ((int64_t)me * n + e) * W;
#pragma unroll
        for (int k = 0; k < W; ++k) st_pair(box + k, w[k], epoch);
      }
    }
    // ---- sum: poll the local inbox, add in rank order (my own contribution at position `me`) ----
    for (int64_t e = e0; e < n; e += stride) buf[e] = sum_in_rank_order<T, W>(my_area, n, e, buf[e], me, world, epoch, dead);
  } else {
    uint2* my_gather = my_area + (int64_t)world * slice * W;      // behind the reduce-scatter inbox [world][slice]
    // ---- reduce-scatter push: my copy of element e goes to its owner's inbox [me][e - owner * slice] ----
    for (int64_t e = e0; e < n; e += stride) {
      const int owner = (int)(e / slice);
      if (owner == me) continue;
      uint2* box = area_ptr(tab.base[owner], channel, half) + ((int64_t)me * slice + (e - (int64_t)owner * slice)) * W;
#pragma unroll
      for (int k = 0; k < W; ++k) st_pair(box + k, words[e * W + k], epoch);
    }
    // ---- owner: add the copies in rank order, keep the sum and push it to every peer's gather inbox [e] ----
    for (int64_t e = e0; e < n; e += stride) {
      if ((int)(e / slice) != me) continue;
      const int64_t el = e - (int64_t)me * slice;
      const T acc = sum_in_rank_order<T, W>(my_area, slice, el, buf[e], me, world, epoch, dead);
      buf[e] = acc;
      const unsigned* aw = reinterpret_cast<const unsigned*>(&acc);
      for (int p = 1; p < world; ++p) {
        const int dst = (me + p) % world;
        uint2* box = area_ptr(tab.base[dst], channel, half) + ((int64_t)world * slice + e) * W;
#pragma unroll
        for (int k = 0; k < W; ++k) st_pair(box + k, aw[k], epoch);
      }
    }
    // ---- everybody else: take the owner's sum ----
    for (int64_t e = e0; e < n; e += stride) {
      if ((int)(e / slice) == me) continue;
      unsigned w[W];
#pragma unroll
      for (int k = 0; k < W; ++k) w[k] = poll_word(my_gather + e * W + k, epoch, dead);
      buf[e] = from_words<T, W>(w);
    }
  }
  if (dead) ctrl->error = 1u;
  // ---- the last block of the launch advances the epoch (stream order makes it visible to the next call) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctrl->arrived, 1u) == gridDim.x - 1) {
      ctrl->arrived = 0u;
      __threadfence();
      ctrl->epoch = epoch;
    }
  }
}

template <typename T, int W>
int peer_allreduce(T* buf, int64_t n, int channel, cudaStream_t st) {
  NPML_CHECK(g_peer.ready, "peer memory is not initialised (npml_peer_open)");
  NPML_CHECK(channel >= 0 && channel < kChannels, "bad channel");
  if (n <= 0) return 0;
  const size_t bytes = (size_t)n * sizeof(T);
  NPML_CHECK(bytes <= kMaxMsgBytes, "message larger than the peer inbox");
  const int world = g_peer.t.world;
  const bool two_shot = (size_t)(world - 1) * 2 * bytes > kOneShotWire && world > 2;
  // a few elements per thread and a small grid: these blocks poll next to the persistent convolution kernels
  // (148 blocks of 256 spinning threads cost the overlapped dgrad 10 % at 8 GPUs; profiles/r2n)
  int64_t blocks = (n + 2 * kThreads - 1) / (2 * kThreads);
  const int64_t cap = two_shot ? 148 : 74;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (two_shot) {
    const int64_t slice = (n + world - 1) / world;
    NPML_CHECK((size_t)(world * slice + n) * W * 8 <= kAreaBytes, "message larger than the peer inbox");
    peer_allreduce_kernel<T, W, true><<<(int)blocks, kThreads, 0, st>>>(g_peer.t, buf, n, slice, channel);
  } else {
    NPML_CHECK((size_t)world * n * W * 8 <= kAreaBytes, "message larger than the peer inbox");
    peer_allreduce_kernel<T, W, false><<<(int)blocks, kThreads, 0, st>>>(g_peer.t, buf, n, n, channel);
  }
  NPML_LAUNCH_OK();
  return 0;
}

}  // namespace

extern "C" {

size_t npml_peer_region_bytes(void) { return kRegionBytes; }

// Allocate this rank's region and return its IPC handle (64 bytes).
int npml_peer_alloc(void* handle64) {
  NPML_CHECK(handle64 != nullptr, "handle buffer is NULL");
  NPML_CHECK(g_peer.local == nullptr, "peer region already allocated");
  NPML_CUDA(cudaMalloc(&g_peer.local, kRegionBytes));
  NPML_CUDA(cudaMemset(g_peer.local, 0, kRegionBytes));       // epoch 0 everywhere: no word of call 1 is "there" yet
  cudaIpcMemHandle_t h;
  NPML_CUDA(cudaIpcGetMemHandle(&h, g_peer.local));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}

// handles: world x 64 bytes (rank order).  Maps every peer's region; the caller must put a host barrier behind this
// call on all ranks before the first all-reduce (every region must be zeroed before a peer writes into it).
int npml_peer_open(const void* handles, int32_t world, int32_t rank) {
  NPML_CHECK(handles != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad arguments");
  NPML_CHECK(g_peer.local != nullptr, "npml_peer_alloc first");
  g_peer.t.world = world;
  g_peer.t.rank = rank;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { g_peer.t.base[r] = (char*)g_peer.local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
    void* p = nullptr;
    NPML_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g_peer.opened[r] = p;
    g_peer.t.base[r] = (char*)p;
  }
  NPML_CUDA(cudaDeviceSynchronize());
  g_peer.ready = true;
  return 0;
}

int npml_peer_ready(void) { return g_peer.ready ? 1 : 0; }

int npml_peer_allreduce_sum_f32(float* buf, int64_t n, int32_t channel, void* stream) {
  return peer_allreduce<float, 1>(buf, n, channel, (cudaStream_t)stream);
}
int npml_peer_allreduce_sum_f64(double* buf, int64_t n, int32_t channel, void* stream) {
  return peer_allreduce<double, 2>(buf, n, channel, (cudaStream_t)stream);
}

// 0 = fine; 1 = a block gave up waiting for a peer (the results of that call are garbage).  Synchronises.
int npml_peer_status(void) {
  if (!g_peer.local) return 0;
  Ctrl c[kChannels];
  if (cudaMemcpy(c, g_peer.local, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return 2;
  for (int i = 0; i < kChannels; ++i)
    if (c[i].error) return 1;
  return 0;
}

int npml_peer_close(void) {
  for (int r = 0; r < kMaxRanks; ++r)
    if (g_peer.opened[r]) { cudaIpcCloseMemHandle(g_peer.opened[r]); g_peer.opened[r] = nullptr; }
  if (g_peer.local) { cudaFree(g_peer.local); g_peer.local = nullptr; }
  g_peer.ready = false;
  return 0;
}

}  // extern "C"
