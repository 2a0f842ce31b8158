// Data-parallel gradient exchange: one SUM all-reduce of the flat fp32 gradient bucket (dW || db of
// every layer) per step over NCCL / NVLink 5.  The reference is single-process; the
// sum (never an average) matches layers/layers.py:3087-3089.  NCCL is resolved with dlopen so that
// the library has no link-time dependency on it: NPML_NCCL_LIB may name the .so explicitly,
// otherwise an already-loaded libnccl.so.2 (torch's) is used.
#include <dlfcn.h>

#include <cstdlib>

#include "npml_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 0;
};
Nccl g_nccl;

int load_nccl() {
  if (g_nccl.h) return 0;
  const char* names[4] = {std::getenv("NPML_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 3 && !g_nccl.h; ++i) {
    if (!names[i]) continue;
    g_nccl.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  }
  NPML_CHECK(g_nccl.h != nullptr, "cannot dlopen libnccl.so.2 (set NPML_NCCL_LIB)");
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.h, "ncclGetErrorString");
  NPML_CHECK(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy,
             "libnccl is missing a required symbol");
  return 0;
}

#define NPML_NCCL(expr)                                                                          \
  do {                                                                                           \
    ncclResult_t _r = (expr);                                                                    \
    if (_r != 0) {                                                                               \
      npml::set_error(std::string(__func__) + ": " #expr " -> " +                                \
                      (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"));       \
      return 4;                                                                                  \
    }                                                                                            \
  } while (0)

}  // namespace

extern "C" int npml_comm_unique_id(void* id128) {
  if (int e = load_nccl()) return e;
  NPML_CHECK(id128 != nullptr, "id buffer is NULL");
  NPML_NCCL(g_nccl.GetUniqueId((ncclUniqueId*)id128));
  return 0;
}

extern "C" int npml_comm_init_rank(const void* id128, int32_t nranks, int32_t rank) {
  if (int e = load_nccl()) return e;
  NPML_CHECK(g_nccl.comm == nullptr, "communicator already initialised");
  NPML_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NPML_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  g_nccl.nranks = nranks;
  return 0;
}

extern "C" int npml_allreduce_sum_f32(float* buf, int64_t n, void* stream) {
  NPML_CHECK(g_nccl.comm != nullptr, "communicator not initialised");
  if (n <= 0) return 0;
  NPML_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, g_nccl.comm, (cudaStream_t)stream));
  npml::count_launch();
  return 0;
}

extern "C" int npml_allreduce_sum_f64(double* buf, int64_t n, void* stream) {
  NPML_CHECK(g_nccl.comm != nullptr, "communicator not initialised");
  if (n <= 0) return 0;
  NPML_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat64, ncclSum, g_nccl.comm, (cudaStream_t)stream));
  npml::count_launch();
  return 0;
}

extern "C" int npml_comm_destroy(void) {
  if (g_nccl.comm) {
    NPML_NCCL(g_nccl.CommDestroy(g_nccl.comm));
    g_nccl.comm = nullptr;
  }
  return 0;
}
