// All-reduce (SUM) of the small gradient / statistics buffers of the data-parallel step over NVLink / NVSwitch PEER
// MEMORY, written as one kernel per call instead of a NCCL collective (dW || db of a layer is 1 KB - 2 MB,
// the sync-BN statistics 2 KB -- all latency-bound).
//
//   push   every rank writes its slice of the buffer straight into an inbox in every peer's memory (posted NVLink
//          stores, no round trip), then releases one flag per block in every peer's flag array (st.release.sys);
//   sum    every rank waits on its OWN (local) flags, then adds the `world` inbox copies in RANK ORDER -- the same
//          order on every rank, so all ranks hold bit-identical results, run after run (the reference sums the batch in
//          one fixed order too, layers/layers.py:3109-3110).
//
// Block b of rank r only ever waits for block b of the other ranks, which wait for nobody before they push: no
// co-residency requirement, no grid barrier.  The call number ("epoch") lives in device memory and is advanced by the
// last block of a launch, so a captured CUDA graph replays correctly.  Calls alternate between two inbox halves, which
// is enough: a rank can finish call k+1 only after every peer has finished READING call k (stream order on the peer).
// Independent streams (the gradient side stream, the sync-BN statistics on the main stream) use different channels.
//
// Memory: one cudaMalloc'ed region per rank, exported with cudaIpcGetMemHandle and opened by every peer at init time
// (host side: numpy-ml_b200/dist.py).  This is the one allocation the library makes itself -- at init, never on the
// hot path -- because an IPC handle must name a whole cudaMalloc allocation.
#include <cstring>

#include "npml_common.cuh"

namespace {

constexpr int kMaxRanks = 8;
constexpr int kChannels = 4;
constexpr int kMaxBlocks = 1024;
constexpr int kThreads = 256;
constexpr size_t kSlotBytes = 4u << 20;          // largest message per call (cfg5: 2.1 MB)
constexpr unsigned kSpinLimit = 1u << 23;        // a few seconds of polling: a peer that never arrives must not hang the box

// region layout (identical on every rank)
//   ctrl  [kChannels] { epoch, arrived }                              local bookkeeping
//   flags [kChannels][2][kMaxRanks][kMaxBlocks] uint32               written by peers
//   inbox [kChannels][2][kMaxRanks][kSlotBytes]                      written by peers
struct Ctrl { unsigned epoch, arrived, error, pad; };
constexpr size_t kCtrlBytes = 4096;
constexpr size_t kFlagBytes = (size_t)kChannels * 2 * kMaxRanks * kMaxBlocks * sizeof(unsigned);
constexpr size_t kInboxOff = kCtrlBytes + kFlagBytes;
constexpr size_t kRegionBytes = kInboxOff + (size_t)kChannels * 2 * kMaxRanks * kSlotBytes;

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

__device__ __forceinline__ unsigned* flag_ptr(char* base, int ch, int half, int src, int blk) {
  return reinterpret_cast<unsigned*>(base + kCtrlBytes) + (((size_t)ch * 2 + half) * kMaxRanks + src) * kMaxBlocks + blk;
}
__device__ __forceinline__ char* inbox_ptr(char* base, int ch, int half, int src) {
  return base + kInboxOff + (((size_t)ch * 2 + half) * kMaxRanks + src) * kSlotBytes;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// T = float or double; 16-byte vectors of VEC = 16 / sizeof(T) elements.  Block b owns vectors [b*vpb, (b+1)*vpb).
template <typename T>
__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(PeerTable tab, T* __restrict__ buf, int64_t n, int vpb,
                                                                  int channel) {
  constexpr int VEC = 16 / (int)sizeof(T);
  char* mine = tab.base[tab.rank];
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(mine) + channel;
  const unsigned epoch = ctrl->epoch + 1u;           // every block reads it before the last block advances it
  const int half = (int)(epoch & 1u);
  const int b = blockIdx.x;
  const int64_t nvec = (n + VEC - 1) / VEC;
  const int64_t v0 = (int64_t)b * vpb;
  const int64_t v1 = v0 + vpb < nvec ? v0 + vpb : nvec;

  // ---- push my slice into every rank's inbox (my own included: the sum below reads all copies the same way) ----
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    uint4 u;
    if ((v + 1) * VEC <= n) {
      u = *reinterpret_cast<const uint4*>(buf + v * VEC);
    } else {                                          // ragged tail: pad with zeros
      T tmp[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) tmp[j] = v * VEC + j < n ? buf[v * VEC + j] : (T)0;
      u = *reinterpret_cast<const uint4*>(tmp);
    }
    for (int p = 0; p < tab.world; ++p) {
      const int dst = (tab.rank + p) % tab.world;     // stagger the targets so the ranks do not all hit one peer first
      reinterpret_cast<uint4*>(inbox_ptr(tab.base[dst], channel, half, tab.rank))[v] = u;
    }
  }
  __syncthreads();
  if (threadIdx.x < tab.world) {
    __threadfence_system();                           // the block's stores are ordered before the flag
    st_release_sys(flag_ptr(tab.base[threadIdx.x], channel, half, tab.rank, b), epoch);
  }
  // ---- wait for every rank's block b (local flags), then add the copies in rank order ----
  if (threadIdx.x < tab.world) {
    const unsigned* f = flag_ptr(mine, channel, half, threadIdx.x, b);
    unsigned spins = 0;
    while (ld_acquire_sys(f) != epoch) {
      if (++spins > kSpinLimit) { ctrl->error = 1u; break; }
    }
  }
  __syncthreads();
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    T acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = (T)0;
    for (int r = 0; r < tab.world; ++r) {
      const uint4 u = __ldcg(reinterpret_cast<const uint4*>(inbox_ptr(mine, channel, half, r)) + v);
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += e[j];
    }
    if ((v + 1) * VEC <= n) {
      *reinterpret_cast<uint4*>(buf + v * VEC) = *reinterpret_cast<const uint4*>(acc);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        if (v * VEC + j < n) buf[v * VEC + j] = acc[j];
    }
  }
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

template <typename T>
int peer_allreduce(T* buf, int64_t n, int channel, cudaStream_t st) {
  NPML_CHECK(g_peer.ready, "peer memory is not initialised (npml_peer_open)");
  NPML_CHECK(channel >= 0 && channel < kChannels, "bad channel");
  if (n <= 0) return 0;
  NPML_CHECK((size_t)n * sizeof(T) <= kSlotBytes, "message larger than a peer inbox slot");
  NPML_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "buffer must be 16-byte aligned");
  constexpr int VEC = 16 / (int)sizeof(T);
  const int64_t nvec = (n + VEC - 1) / VEC;
  int vpb = kThreads;                                 // one vector per thread ...
  while ((nvec + vpb - 1) / vpb > kMaxBlocks) vpb *= 2;   // ... unless that needs more blocks than there are flags
  const int blocks = (int)((nvec + vpb - 1) / vpb);
  peer_allreduce_kernel<T><<<blocks, kThreads, 0, st>>>(g_peer.t, buf, n, vpb, channel);
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
  NPML_CUDA(cudaMemset(g_peer.local, 0, kInboxOff));          // ctrl + flags
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
  return peer_allreduce<float>(buf, n, channel, (cudaStream_t)stream);
}
int npml_peer_allreduce_sum_f64(double* buf, int64_t n, int32_t channel, void* stream) {
  return peer_allreduce<double>(buf, n, channel, (cudaStream_t)stream);
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
