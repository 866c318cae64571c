// NCCL plumbing (see comm.h).  libnccl is dlopen'ed so the library loads on boxes without it.
#include "comm.h"
#include "device.h"

#include <dlfcn.h>

#include <cstring>

namespace shakti {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums below.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_api;
static ncclComm_t g_nccl = nullptr;
static Comm g_comm;
Comm& comm() { return g_comm; }

static void load_api() {
  if (g_api.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_api.lib) break;
  }
  if (!g_api.lib) throw Error(SHAKTI_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
  auto sym = [&](const char* s) {
    void* p = dlsym(g_api.lib, s);
    if (!p) throw Error(SHAKTI_ERR_COMM, std::string("NCCL symbol missing: ") + s);
    return p;
  };
  g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
  g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
  g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
  g_api.Send = (decltype(g_api.Send))sym("ncclSend");
  g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
  g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
  g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
}

#define SHAKTI_NCCL(call)                                                                    \
  do {                                                                                       \
    int r_ = (call);                                                                         \
    if (r_ != ncclSuccess)                                                                   \
      throw Error(SHAKTI_ERR_COMM, std::string(#call) + ": " + g_api.GetErrorString(r_));    \
  } while (0)

void comm_unique_id(uint8_t id[128]) {
  load_api();
  ncclUniqueId u;
  SHAKTI_NCCL(g_api.GetUniqueId(&u));
  std::memcpy(id, u.internal, 128);
}

void comm_init(const uint8_t id[128], int rank, int nranks, int device) {
  SHAKTI_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
  if (nranks == 1) { g_comm.rank = 0; g_comm.nranks = 1; return; }
  load_api();
  if (device >= 0) SHAKTI_CUDA(cudaSetDevice(device));
  ncclUniqueId u;
  std::memcpy(u.internal, id, 128);
  SHAKTI_NCCL(g_api.CommInitRank(&g_nccl, nranks, u, rank));
  g_comm.rank = rank;
  g_comm.nranks = nranks;
}

void comm_finalize() {
  if (g_nccl) g_api.CommDestroy(g_nccl);
  g_nccl = nullptr;
  g_comm = Comm();
}

void comm_allreduce_sum(double* dev, int count, cudaStream_t s) {
  if (!g_comm.active() || count == 0) return;
  SHAKTI_NCCL(g_api.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, g_nccl, s));
}

void HaloPlan::build(const std::vector<Neighbor>& nbrs) {
  peers.clear();
  std::vector<int32_t> idx;
  for (const auto& nb : nbrs) {
    Peer p;
    p.rank = nb.rank;
    p.send_off = (int32_t)idx.size();
    p.send_cnt = (int32_t)nb.send_local.size();
    p.recv_begin = nb.recv_begin;
    p.recv_cnt = nb.recv_count;
    idx.insert(idx.end(), nb.send_local.begin(), nb.send_local.end());
    peers.push_back(p);
  }
  n_send = (int32_t)idx.size();
  if (n_send) {
    send_idx.upload(idx);
    send_buf.alloc((size_t)n_send);
  }
}

void HaloPlan::exchange(double* v, cudaStream_t s) {
  if (!g_comm.active() || peers.empty()) return;
  if (n_send) launch_gather(n_send, send_idx.p, v, send_buf.p, s);
  SHAKTI_NCCL(g_api.GroupStart());
  for (const auto& p : peers) {
    if (p.send_cnt) SHAKTI_NCCL(g_api.Send(send_buf.p + p.send_off, (size_t)p.send_cnt, ncclFloat64, p.rank, g_nccl, s));
    if (p.recv_cnt) SHAKTI_NCCL(g_api.Recv(v + p.recv_begin, (size_t)p.recv_cnt, ncclFloat64, p.rank, g_nccl, s));
  }
  SHAKTI_NCCL(g_api.GroupEnd());
}

}  // namespace shakti
